mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_adapter.py -m gpu -q -x > gpurun_out/r02w_pytest.txt 2>&1
tail -n 3 gpurun_out/r02w_pytest.txt
