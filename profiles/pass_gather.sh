TAG=${1:-r01m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_mhl|k_quartet|k_fdrp" -c 10 -o $OUT/prof python profiles/gather_prof.py 30 20000000 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
