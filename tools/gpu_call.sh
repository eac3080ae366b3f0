mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py -m gpu -q -x -k "graph or topped or dense_frames or pipelined" > gpurun_out/r02aq_pytest.txt 2>&1
tail -n 6 gpurun_out/r02aq_pytest.txt
