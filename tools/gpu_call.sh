mkdir -p gpurun_out
P=gpurun_out/r02o
timeout 300 python -m pytest tests/test_gpu_adapter.py tests/test_gpu_multi.py -m gpu -q > ${P}_pytest.txt 2>&1; tail -n 3 ${P}_pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 8 > ${P}_bench_n8.json 2> ${P}_bench_n8.err; echo "bench8 rc=$?"
timeout 600 python bench.py --no-other-shapes --no-cpu-baseline > ${P}_bench_n1.json 2> ${P}_bench_n1.err; echo "bench1 rc=$?"
grep -c '^{' ${P}_bench_n8.json ${P}_bench_n1.json
