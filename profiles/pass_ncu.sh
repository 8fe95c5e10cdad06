# usage: bash profiles/pass_ncu.sh TAG 'regex' [lib]
TAG=$1; RX=$2; LIB=$3
OUT=gpurun_out/$TAG; mkdir -p $OUT
METHEOR_B200_LIB=$LIB timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 6 -c 2 \
    -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log | cut -c1-300
