mkdir -p gpurun_out
P=gpurun_out/r02r
python tools/probes/match_bench.py > ${P}_match_bench.txt 2>&1
python tools/probes/match_bench.py --pairs 37 --n 1900 >> ${P}_match_bench.txt 2>&1
python tools/probes/match_bench.py --pairs 1 --n 2000 --reps 50 >> ${P}_match_bench.txt 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_match_mma3 -s 3 -c 1 -o ${P}_prof_match python tools/probes/match_bench.py --only mma --reps 2 > ${P}_ncu3.log 2>&1
cat ${P}_match_bench.txt; tail -n 2 ${P}_ncu3.log
