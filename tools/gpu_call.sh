mkdir -p gpurun_out
P=gpurun_out/r02t
SHORT="python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-next-rows --no-other-shapes"
timeout 600 $SHORT > ${P}_bench_short.json 2> ${P}_bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv $SHORT > ${P}_ncu1.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:'k_(detect|octree|blur|describe|resize|repitch)' -c 26 -o ${P}_prof_extract $SHORT > ${P}_ncu2.log 2>&1
tail -n 2 ${P}_ncu1.log ${P}_ncu2.log | cut -c1-300; ls -la gpurun_out | grep r02t
