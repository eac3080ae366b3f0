mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_frame_reference.py tests/test_search_reference.py -m gpu -q > gpurun_out/r02k_pytest.txt 2>&1
tail -n 15 gpurun_out/r02k_pytest.txt
