"""Minimal BGZF/BAM/SAM reader + writer for test fixtures (pure Python + zlib).

Test-only helper: it lets the suite rebuild the reference's fixture BAMs from the record-level JSON
committed under tests/golden/ (the reference tree is not available on the GPU box) and create
synthetic BAMs (reverse strand, indels, soft clips, multiple contigs) for the host decoder tests.
Format facts: SAM/BAM spec v1 (BGZF = gzip members with a 'BC' extra field; BAM little-endian).
"""
import struct
import zlib

CIGAR_OPS = "MIDNSHP=X"


def parse_cigar(s):
    out, n = [], 0
    if s == "*":
        return out
    for ch in s:
        if ch.isdigit():
            n = n * 10 + int(ch)
        else:
            out.append((n, CIGAR_OPS.index(ch)))
            n = 0
    return out


def cigar_str(ops):
    return "".join(f"{n}{CIGAR_OPS[o]}" for n, o in ops) or "*"


def bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    c = comp.compress(data) + comp.flush()
    bsize = len(c) + 25
    hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    return hdr + c + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_compress(data, block=0xFF00):
    out = bytearray()
    for i in range(0, len(data), block):
        out += bgzf_block(data[i:i + block])
    out += BGZF_EOF
    return bytes(out)


def bgzf_decompress(raw):
    out, off = bytearray(), 0
    while off < len(raw):
        assert raw[off] == 31 and raw[off + 1] == 139
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        x, bsize = off + 12, None
        while x < off + 12 + xlen:
            si1, si2, slen = raw[x], raw[x + 1], struct.unpack_from("<H", raw, x + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", raw, x + 4)[0]
            x += 4 + slen
        total = bsize + 1
        cdata = raw[off + 12 + xlen: off + total - 8]
        out += zlib.decompress(cdata, -15)
        off += total
    return bytes(out)


def write_bam(path, refs, reads, header_text=None, with_seq=False, block=0xFF00):
    """refs: [(name, length)], reads: dicts with tid,pos,flag,mapq,cigar(str),xm(str or None)."""
    if header_text is None:
        header_text = "@HD\tVN:1.0\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    ht = header_text.encode()
    buf = bytearray(b"BAM\x01" + struct.pack("<i", len(ht)) + ht + struct.pack("<i", len(refs)))
    for n, l in refs:
        nb = n.encode() + b"\0"
        buf += struct.pack("<i", len(nb)) + nb + struct.pack("<i", l)
    rec_off = []  # uncompressed offset of every record
    for i, r in enumerate(reads):
        rec_off.append(len(buf))
        name = (r.get("name") or f"read_{i}").encode() + b"\0"
        ops = parse_cigar(r["cigar"])
        qlen = sum(n for n, o in ops if o in (0, 1, 4, 7, 8))
        l_seq = qlen if with_seq else 0
        seq = bytes([0x11] * ((l_seq + 1) // 2))
        qual = bytes([30] * l_seq)
        aux = b""
        if r.get("nm") is not None:
            aux += b"NMC" + bytes([r["nm"]])
        if r.get("xm") is not None:
            aux += b"XMZ" + r["xm"].encode() + b"\0"
        body = struct.pack("<iiBBHHHiiii", r["tid"], r["pos"], len(name), r["mapq"], 4680, len(ops), r["flag"],
                           l_seq, -1, -1, 0)
        body += name + b"".join(struct.pack("<I", (n << 4) | o) for n, o in ops) + seq + qual + aux
        buf += struct.pack("<i", len(body)) + body
    with open(path, "wb") as f:
        f.write(bgzf_compress(bytes(buf), block))
    return rec_off


def write_bai(bam_path, refs, reads, rec_off, block=0xFF00, bai_path=None):
    """A .bai for a BAM written by write_bam(..., block): no bins, only the linear index (smallest virtual offset of the reads
    overlapping each 16 kb window; empty windows stay 0, as in indexes that were not back-filled)."""
    raw = open(bam_path, "rb").read()
    coffs, off = [], 0  # compressed offset of every BGZF member
    while off < len(raw):
        coffs.append(off)
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        x, bsize = off + 12, None
        while x < off + 12 + xlen:
            slen = struct.unpack_from("<H", raw, x + 2)[0]
            if raw[x] == 66 and raw[x + 1] == 67:
                bsize = struct.unpack_from("<H", raw, x + 4)[0]
            x += 4 + slen
        off += bsize + 1
    lin = [dict() for _ in refs]
    for r, uo in zip(reads, rec_off):
        if r["tid"] < 0:
            continue
        voff = (coffs[uo // block] << 16) | (uo % block)
        ops = parse_cigar(r["cigar"])
        rlen = max(1, sum(n for n, o in ops if o in (0, 2, 3, 7, 8)))
        for w in range(r["pos"] >> 14, ((r["pos"] + rlen - 1) >> 14) + 1):
            d = lin[r["tid"]]
            d[w] = min(d.get(w, voff), voff)
    out = bytearray(b"BAI\x01" + struct.pack("<i", len(refs)))
    for d in lin:
        out += struct.pack("<i", 0)  # n_bin
        n_intv = (max(d) + 1) if d else 0
        out += struct.pack("<i", n_intv)
        for w in range(n_intv):
            out += struct.pack("<Q", d.get(w, 0))
    open(bai_path or bam_path + ".bai", "wb").write(bytes(out))


def read_bam(path):
    """-> (refs, reads) with the same record dict layout as write_bam takes."""
    d = bgzf_decompress(open(path, "rb").read())
    assert d[:4] == b"BAM\x01"
    o = 4
    l_text = struct.unpack_from("<i", d, o)[0]; o += 4
    text = d[o:o + l_text].decode(); o += l_text
    n_ref = struct.unpack_from("<i", d, o)[0]; o += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", d, o)[0]; o += 4
        name = d[o:o + ln - 1].decode(); o += ln
        refs.append((name, struct.unpack_from("<i", d, o)[0])); o += 4
    reads = []
    while o < len(d):
        bs = struct.unpack_from("<i", d, o)[0]; o += 4
        p = d[o:o + bs]; o += bs
        tid, pos, lrn, mapq, _bin, ncig, flag, lseq = struct.unpack_from("<iiBBHHHi", p, 0)
        q = 32
        name = p[q:q + lrn - 1].decode(); q += lrn
        ops = []
        for _ in range(ncig):
            v = struct.unpack_from("<I", p, q)[0]; q += 4
            ops.append((v >> 4, v & 15))
        q += (lseq + 1) // 2 + lseq
        xm = None
        while q + 3 <= len(p):
            tag, ty = p[q:q + 2], chr(p[q + 2]); q += 3
            if ty in "AcC":
                ln = 1
            elif ty in "sS":
                ln = 2
            elif ty in "iIf":
                ln = 4
            elif ty in "ZH":
                ln = p.index(b"\0", q) - q + 1
            elif ty == "B":
                st, cnt = chr(p[q]), struct.unpack_from("<i", p, q + 1)[0]
                ln = 5 + cnt * (1 if st in "cC" else 2 if st in "sS" else 4)
            else:
                raise ValueError(ty)
            if tag == b"XM" and ty == "Z":
                xm = p[q:q + ln - 1].decode()
            q += ln
        reads.append(dict(name=name, tid=tid, pos=pos, flag=flag, mapq=mapq, cigar=cigar_str(ops), xm=xm))
    return refs, reads, text


def read_header_len(path):
    """-> (refs, number of uncompressed bytes in front of the first record)."""
    d = bgzf_decompress(open(path, "rb").read())
    o = 4
    l_text = struct.unpack_from("<i", d, o)[0]; o += 4 + l_text
    n_ref = struct.unpack_from("<i", d, o)[0]; o += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", d, o)[0]; o += 4
        name = d[o:o + ln - 1].decode(); o += ln
        refs.append((name, struct.unpack_from("<i", d, o)[0])); o += 4
    return refs, o


def read_sam(path):
    refs, reads = [], []
    for line in open(path):
        line = line.rstrip("\n")
        if not line:
            continue
        f = line.split("\t")
        if line[0] == "@":
            if f[0] == "@SQ":
                sn = [x[3:] for x in f if x.startswith("SN:")][0]
                ln = int([x[3:] for x in f if x.startswith("LN:")][0])
                refs.append((sn, ln))
            continue
        names = [n for n, _ in refs]
        xm = None
        for x in f[11:]:
            if x.startswith("XM:Z:"):
                xm = x[5:]
        reads.append(dict(name=f[0], tid=names.index(f[2]) if f[2] != "*" else -1, pos=int(f[3]) - 1, flag=int(f[1]),
                          mapq=int(f[4]), cigar=f[5], xm=xm))
    return refs, reads


def write_sam(path, refs, reads):
    with open(path, "w") as f:
        f.write("@HD\tVN:1.0\tSO:coordinate\n")
        for n, l in refs:
            f.write(f"@SQ\tSN:{n}\tLN:{l}\n")
        for i, r in enumerate(reads):
            ops = parse_cigar(r["cigar"])
            qlen = sum(n for n, o in ops if o in (0, 1, 4, 7, 8))
            chrom = refs[r["tid"]][0] if r["tid"] >= 0 else "*"
            tags = "" if r.get("xm") is None else f"\tXM:Z:{r['xm']}"
            f.write(f"{r.get('name') or f'read_{i}'}\t{r['flag']}\t{chrom}\t{r['pos'] + 1}\t{r['mapq']}\t{r['cigar']}\t*\t0\t0\t"
                    f"{'A' * qlen or '*'}\t{'I' * qlen or '*'}{tags}\n")
