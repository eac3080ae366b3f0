mkdir -p gpurun_out
O=gpurun_out/r02ar_stage_knockout.txt; : > $O
for m in 0 1 2 4 8 16 6 3 24; do LATE_ENV=ORB_B200_SKIP=$m python tools/probes/resident_probe.py 2>&1 | tail -n 1 >> $O; done
cat $O
