TAG=${1:-r01m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${3:-k_mhl|k_quartet|k_fdrp}" -c ${4:-10} -o $OUT/prof python profiles/gather_prof.py 30 20000000 ${2:-mhl,pm,fdrp,qfdrp} > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
