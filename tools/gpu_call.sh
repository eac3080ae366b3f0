mkdir -p gpurun_out
O=gpurun_out/r02ad_bk.txt; : > $O
python tools/probes/mma_probe.py --kinds i8,f8 --variants 0 >> $O 2>&1
python tools/probes/match_sweep.py --env ORB_B200_MMA_BK=0,1 --env ORB_B200_MMA_DEBUG=0,3,14 >> $O 2>&1
timeout 600 python -m pytest tests/test_gpu_match.py -m gpu -q -x 2>&1 | tail -n 3 >> $O
cat $O
