mkdir -p gpurun_out
P=gpurun_out/r02h
timeout 600 python tools/probes/mma_probe.py --kinds i8,f8 --variants 0,20 > ${P}_mma_probe.txt 2>&1
timeout 300 python tools/probes/match_bench.py > ${P}_match_bench.txt 2>&1
cat ${P}_mma_probe.txt ${P}_match_bench.txt
