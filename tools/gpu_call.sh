mkdir -p gpurun_out
O=gpurun_out/r02at_blur.txt; : > $O
python tools/probes/resident_probe.py 2>&1 | tail -n 1 >> $O
python tools/probes/resident_probe.py 2>&1 | tail -n 1 >> $O
timeout 900 python -m pytest tests/test_gpu_extract.py -m gpu -q -x 2>&1 | tail -n 3 >> $O
cat $O
