#!/bin/bash
# A/B of the tile-kernel instances (FDRP: 64 sites x 1024 reads | 32 x 1024 | 64 x 2048; MHL: 2048 | 4608 reads) on chr1 + chr19-like density
O=gpurun_out/${1:-r2e}; mkdir -p $O
for v in dense sparse32 sparse64; do
  METHEOR_FDRP_TILE=$v python profiles/wg_pass.py --contigs 0,1 --sets fdrp,qfdrp,fdrp+qfdrp --warm 1 --profile 2>/dev/null | tail -1 > $O/fdrp_$v.json
done
for v in dense sparse; do
  METHEOR_MHL_TILE=$v python profiles/wg_pass.py --contigs 0,1 --sets mhl --warm 1 --profile 2>/dev/null | tail -1 > $O/mhl_$v.json
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    d=json.load(open(f))
    for s in d["sets"]:
        print(f.split("/")[-1], s["name"], s["fallback_sites"], {k:v for k,v in s["kernels_ms"].items() if k.startswith("~") or k.startswith("k_fdrp") or k.startswith("k_qfdrp") or k=="k_mhl"})
PY
