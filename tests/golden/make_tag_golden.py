"""Regenerates tests/golden/tag_chr19.json from the reference's own `tag` fixtures.

Run HERE (container with /root/reference mounted read-only):  python tests/golden/make_tag_golden.py
The reference's CLI test (tests/tag-cli.rs:62-83) runs `metheor tag` on tests/test.chr19.noXM.sam with tests/hg38.chr19.fa
and requires the output to equal tests/test.chr19.XM.sam byte for byte (the committed tests/test.chr19.metheor_tag_out.sam
is that output).  The genome is 59 MB, so this script keeps only what `tag` can look at: the reference bases of
[start - 2, end + 2) of every read (tag.rs:155-161), merged into windows.  The tests rebuild a chr19-sized FASTA with N
everywhere else.  No reference source code is copied, only test data.
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/tests"


def main():
    header, lines = [], []
    for ln in open(f"{REF}/test.chr19.noXM.sam").read().split("\n"):
        if ln.startswith("@"):
            header.append(ln)
        elif ln:
            lines.append(ln)
    expect = [ln for ln in open(f"{REF}/test.chr19.metheor_tag_out.sam").read().split("\n") if ln and not ln.startswith("@")]
    assert open(f"{REF}/test.chr19.metheor_tag_out.sam").read() == open(f"{REF}/test.chr19.XM.sam").read()
    assert len(expect) == len(lines)
    xm = []
    for a, b in zip(lines, expect):
        assert b.startswith(a + "\tXM:Z:"), "the reference appends XM:Z as the last field (tag.rs:417)"
        xm.append(b[len(a) + 6:])
    # chr19 sequence, one string (the .fai says: offset 7, 60 bases per 61-byte line)
    name, length, off, lb, lw = open(f"{REF}/hg38.chr19.fa.fai").read().split()
    length, off = int(length), int(off)
    with open(f"{REF}/hg38.chr19.fa", "rb") as f:
        f.seek(off)
        seq = f.read().replace(b"\n", b"")[:length].decode()
    assert len(seq) == length
    spans = []
    for ln in lines:
        f = ln.split("\t")
        pos = int(f[3]) - 1
        ref_len = sum(int(n) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", f[5]) if op in "MDN=X")
        spans.append((max(pos - 2, 0), min(pos + ref_len + 2, length)))
    spans.sort()
    merged = []
    for s, e in spans:
        if merged and s <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([s, e])
    out = dict(source=dict(input="tests/test.chr19.noXM.sam", expected="tests/test.chr19.metheor_tag_out.sam == tests/test.chr19.XM.sam",
                           genome="tests/hg38.chr19.fa"),
               header=header, contig=name, contig_length=length, records=lines, xm=xm,
               windows=[[s, seq[s:e]] for s, e in merged])
    with open(os.path.join(HERE, "tag_chr19.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(len(lines), "records,", len(merged), "windows,", sum(e - s for s, e in merged), "reference bases,",
          os.path.getsize(os.path.join(HERE, "tag_chr19.json")), "bytes")


if __name__ == "__main__":
    main()
