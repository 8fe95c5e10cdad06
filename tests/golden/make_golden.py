"""Regenerates tests/golden/fixtures.json from the reference's own test inputs.

Run HERE (container with /root/reference mounted read-only):  python tests/golden/make_golden.py
The GPU box has no /root/reference, so the suite never reads it at run time: it rebuilds the fixture BAMs from the
record-level description written by this script (tid/pos/flag/mapq/CIGAR/XM per read — the only fields the measures
consume, readutil.rs:28,35,326-338, pdr.rs:150).  No reference source code is copied, only decoded test data.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import bamio  # noqa: E402

REF = "/root/reference/tests"


def main():
    out = {}
    for k in range(1, 7):
        refs, reads, text = bamio.read_bam(f"{REF}/test{k}.bam")
        out[f"test{k}"] = dict(source=f"tests/test{k}.bam", header_text=text, refs=refs,
                               reads=[dict(pos=r["pos"], flag=r["flag"], mapq=r["mapq"], cigar=r["cigar"], xm=r["xm"],
                                           tid=r["tid"], name=r["name"]) for r in reads])
    refs, reads = bamio.read_sam(f"{REF}/test.chr19.XM.sam")
    out["chr19_1000"] = dict(source="tests/test.chr19.XM.sam", refs=refs,
                             reads=[dict(pos=r["pos"], flag=r["flag"], mapq=r["mapq"], cigar=r["cigar"], xm=r["xm"],
                                         tid=r["tid"]) for r in reads])
    with open(os.path.join(HERE, "fixtures.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print({k: len(v["reads"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
