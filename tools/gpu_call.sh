mkdir -p gpurun_out
O=gpurun_out/r02s_mma3.txt; : > $O
timeout 600 python -m pytest tests/test_gpu_match.py -m gpu -q -x > gpurun_out/r02s_pytest.txt 2>&1
tail -n 5 gpurun_out/r02s_pytest.txt
python tools/probes/mma_probe.py --kinds i8,f8 --variants 0,30 >> $O 2>&1
python tools/probes/match_bench.py --only mma >> $O 2>&1
cat $O
ncu --set full --import-source on --clock-control none -k regex:k_match_mma3 -s 3 -c 1 -o gpurun_out/r02s_mma3 python tools/probes/match_bench.py --only mma --reps 2 > gpurun_out/r02s_ncu.log 2>&1
tail -n 3 gpurun_out/r02s_ncu.log
ls -la gpurun_out/r02s_mma3.ncu-rep
