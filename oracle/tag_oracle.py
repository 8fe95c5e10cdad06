"""CPU restatement of the reference's `tag` subcommand (XM synthesis) — TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module; the product path (metheor_b200/csrc/k_tag.cu behind mth_tag, host/tag.cpp) never does.

Follows /root/reference/src/tag.rs line by line, including its quirks:
  * determine_xm_tag_string (tag.rs:130-384): aligned read/reference columns are built from the CIGAR — only M, I and D
    produce columns (tag.rs:186-233, every other op is `_ => {}`: S does not advance the read, N / = / X do not advance
    the reference), two context columns are put on either side (tag.rs:176-183, 235-239), reverse-strand reads work on
    the reverse complement (tag.rs:244-261) and the tag is reversed back at the end (tag.rs:380-383);
  * a 'C' in the reference is classified by its 3-base context: CG -> z/Z, CHG -> x/X, CHH -> h/H, a context holding
    '-' or 'N' -> u/U, and ANY OTHER context (IUPAC codes) emits nothing at all (tag.rs:300-331, 336-375: the if/else
    chain has no final else), so such a tag is shorter than the read;
  * the look-ahead over deletion columns (tag.rs:268-299) is skipped for the last two read positions;
  * panics of the reference are reported as TagPanic: a base outside the reverse-complement map (tag.rs:23, e.g. '='),
    an unmapped read or a read past the contig end (tag.rs:152, 167-172), a look-ahead that finds no second base
    (tag.rs:301 indexes [1]).
Pinned against the reference's own golden (tests/tag-cli.rs:62-83: 1000 chr19 reads, output == Bismark's XM strings) in
tests/test_tag.py through tests/golden/tag_chr19.json.  That golden holds only plain `<len>M` single-end reads: the
insertion / deletion / paired-end branches are restated from the source and not exercised by any reference test.
"""
import os
import re

RC = dict(zip("ACGTNMRWSYKVHDB-", "TGCANKYWSRMBDHV-"))  # tag.rs:78-99
NT16 = "=ACMGRSVTWYHKDBN"
CHG = {"CAG", "CTG", "CCG"}                                                   # tag.rs:32-34
CHH = {"CAA", "CAT", "CAC", "CTA", "CTT", "CTC", "CCA", "CCT", "CCC"}         # tag.rs:36-41


class TagPanic(Exception):
    pass


def _nt16_code(ch):
    """htslib's seq_nt16_table: how SAM text becomes the 4-bit codes `r.seq()` decodes again."""
    c = ch.upper()
    if c in NT16:
        return NT16.index(c)
    return {"0": 1, "1": 2, "2": 4, "3": 8}.get(ch, 15)


def canon_seq(text):
    """SEQ as `str::from_utf8(&r.seq().as_bytes())` sees it (tag.rs:146-149)."""
    if text == "*":
        return ""
    return "".join(NT16[_nt16_code(c)] for c in text)


def need_reverse_complement(flag, paired):
    rev, first, last = bool(flag & 16), bool(flag & 64), bool(flag & 128)
    if paired:  # tag.rs:15-18
        return not ((not rev and first) or (rev and last))
    return rev  # tag.rs:141-144


def parse_cigar(text):
    return [] if text == "*" else [(int(n), op) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", text)]


def reference_end(pos, cigar):
    """htslib bam_endpos: M D N = X consume the reference; a record without any reports pos + 1."""
    n = sum(ln for ln, op in cigar if op in "MDN=X")
    return pos + (n if n else 1)


def xm_string(flag, pos, cigar, seq, contig, chromsize, paired):
    """tag.rs:130-384.  pos 0-based; cigar [(len, op)]; seq canonical upper-case; contig the reference sequence (any case)."""
    start, end = pos, reference_end(pos, cigar)
    rc = need_reverse_complement(flag, paired)
    if end > chromsize:  # tag.rs:168-172: the padding table has three entries
        raise TagPanic("read past the end of the contig")
    if min(end + 2, chromsize) > len(contig):  # tag.rs:158: slice out of range
        raise TagPanic("reference sequence shorter than the header's LN")
    ref_seq = contig[max(start - 2, 0):min(end + 2, chromsize)].upper()
    ref_seq = "N" * max(2 - start, 0) + ref_seq + "N" * max(end - chromsize + 2, 0)
    tread, tref = ["-", "-"], [ref_seq[0], ref_seq[1]]
    ur, uf = 0, 2
    for ln, op in cigar:
        if op == "M":
            tread += list(seq[ur:ur + ln]); tref += list(ref_seq[uf:uf + ln]); ur += ln; uf += ln
        elif op == "I":
            tread += list(seq[ur:ur + ln]); tref += ["-"] * ln; ur += ln
        elif op == "D":
            tread += ["-"] * ln; tref += list(ref_seq[uf:uf + ln]); uf += ln
    tread += ["-", "-"]
    tref += [ref_seq[-2], ref_seq[-1]]
    if len(tread) != len(tref):
        raise TagPanic("SEQ shorter than the CIGAR's M/I bases")
    if rc:
        try:
            R = [RC[c] for c in reversed(tread[:-2])]
            F = [RC[c] for c in reversed(tref[:-2])]
        except KeyError as e:
            raise TagPanic(f"no reverse complement for {e}")
    else:
        R, F = tread[2:], tref[2:]
    L = len(R)
    xm = []

    def classify(ctx, rd):
        if len(ctx) >= 2 and ctx[0] == "C" and ctx[1] == "G":
            kind = "zZ"
        elif ctx in CHG:
            kind = "xX"
        elif ctx in CHH:
            kind = "hH"
        elif "-" in ctx or "N" in ctx:
            kind = "uU"
        else:
            return  # nothing pushed
        xm.append(kind[1] if rd == "C" else kind[0] if rd == "T" else ".")

    for i in range(L - 2):
        if R[i] == "-":
            continue
        if R[i] == "N":
            xm.append(".")
        elif F[i] == "C":
            if (R[i + 1] == "-" or R[i + 2] == "-") and i != L - 3 and i != L - 4:
                ctx, found, k = [F[i]], 0, 1
                while found != 2:
                    if i + k > L - 1:
                        break
                    if R[i + k] != "-":
                        ctx.append(F[i + k]); found += 1
                    k += 1
                if len(ctx) < 2:
                    raise TagPanic("context look-ahead found no second base")
                classify("".join(ctx), R[i])
            else:
                classify("".join(F[i:i + 3]), R[i])
        else:
            xm.append(".")
    return "".join(reversed(xm)) if rc else "".join(xm)


def read_fasta(path):
    """name -> sequence (as in the file, case kept)."""
    if not os.path.exists(path):
        raise TagPanic(f"Error opening reference genome file: file not found: {path}")
    out, name, parts = {}, None, []
    for ln in open(path):
        ln = ln.rstrip("\r\n")
        if ln.startswith(">"):
            if name is not None:
                out[name] = "".join(parts)
            name, parts = ln[1:].split()[0] if ln[1:].split() else "", []
        elif name is not None:
            parts.append(ln.strip())
    if name is not None:
        out[name] = "".join(parts)
    return out


def tag_sam(in_path, out_path, genome_path):
    """`metheor tag -i in.sam -o out.sam -g genome.fa` for SAM input (tag.rs:386-441): header copied, XM:Z appended last."""
    if not os.path.exists(in_path):
        raise TagPanic(f"Error opening BAM file. file not found: {in_path}")
    text = open(in_path).read().split("\n")
    header = [ln for ln in text if ln.startswith("@")]
    recs = [ln for ln in text if ln and not ln.startswith("@")]
    names, sizes = [], []
    for ln in header:
        if ln.startswith("@SQ"):
            f = dict(x.split(":", 1) for x in ln.split("\t")[1:])
            names.append(f["SN"]); sizes.append(int(f["LN"]))
    d = os.path.dirname(out_path)
    if not os.path.isdir(d):
        raise TagPanic(f"No such directory for output alignment file: {d}")
    fa = read_fasta(genome_path)
    for n in names:
        if n not in fa:
            raise TagPanic("Error fetching reference genome sequence.")
    paired = bool(recs) and bool(int(recs[0].split("\t")[1]) & 1)
    out = list(header)
    for ln in recs:
        f = ln.split("\t")
        if f[2] == "*" or f[2] not in names:
            raise TagPanic("unmapped read")
        tid = names.index(f[2])
        xm = xm_string(int(f[1]), int(f[3]) - 1, parse_cigar(f[5]), canon_seq(f[9]), fa[names[tid]], sizes[tid], paired)
        out.append(ln + "\tXM:Z:" + xm)
    with open(out_path, "w") as fo:
        fo.write("\n".join(out) + "\n")
