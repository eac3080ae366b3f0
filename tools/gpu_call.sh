mkdir -p gpurun_out
O=gpurun_out/r02am_graph.txt; : > $O
ORB_B200_GRAPH=1 timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_ingest.py -m gpu -q -x 2>&1 | tail -n 4 >> $O
for g in 0 1; do ORB_B200_GRAPH=$g python tools/probes/resident_probe.py 2>&1 | tail -n 1 >> $O; done
for g in 0 1; do ORB_B200_GRAPH=$g python tools/e2e_probe.py 2>&1 | tail -n 1 >> $O; done
for g in 0 1; do ORB_B200_GRAPH=$g python tools/e2e_probe.py 2>&1 | tail -n 1 >> $O; done
cat $O
