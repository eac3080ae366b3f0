mkdir -p gpurun_out
P=gpurun_out/r02n
timeout 900 python -m pytest tests -m gpu -q > ${P}_pytest.txt 2>&1; tail -n 3 ${P}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.txt 2>&1; tail -n 1 ${P}_smoke.txt
timeout 900 python bench.py > ${P}_bench_default.json 2> ${P}_bench_default.err; echo "bench rc=$?"; tail -n 2 ${P}_bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference.json 2>> ${P}_bench_default.err
timeout 300 python tools/probes/match_bench.py > ${P}_match_bench.txt 2>&1
SHORT="python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-next-rows --no-other-shapes"
timeout 600 $SHORT > ${P}_bench_short.json 2> ${P}_bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file ${P}_launches.csv $SHORT > ${P}_ncu1.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:'k_(detect|octree|blur|describe|resize|repitch)' -c 24 -o ${P}_prof_extract $SHORT > ${P}_ncu2.log 2>&1
MB="python tools/probes/match_bench.py --only mma --reps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_match_mma2 -s 2 -c 2 -o ${P}_prof_match $MB > ${P}_ncu3.log 2>&1
du -sh gpurun_out
