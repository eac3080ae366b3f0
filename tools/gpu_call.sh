mkdir -p gpurun_out
O=gpurun_out/r02q_resident.txt; : > $O
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_adapter.py tests/test_gpu_ingest.py tests/test_gpu_match.py -m gpu -q > gpurun_out/r02q_pytest.txt 2>&1
python tools/probes/resident_probe.py >> $O 2>&1
tail -n 4 gpurun_out/r02q_pytest.txt; cat $O
