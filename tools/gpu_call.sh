mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_match.py -m gpu -q -x > gpurun_out/r02v_pytest.txt 2>&1
tail -n 6 gpurun_out/r02v_pytest.txt
