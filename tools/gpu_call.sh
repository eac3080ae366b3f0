mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02af_pytest.txt 2>&1
tail -n 5 gpurun_out/r02af_pytest.txt
timeout 900 python bench.py > gpurun_out/r02af_bench_default.json 2> gpurun_out/r02af_bench_default.err
tail -c 6000 gpurun_out/r02af_bench_default.json
tail -n 5 gpurun_out/r02af_bench_default.err
