mkdir -p gpurun_out
P=gpurun_out/r02d
nvidia-smi topo -m > ${P}_topo.txt 2>&1; nproc >> ${P}_topo.txt; free -g >> ${P}_topo.txt
timeout 600 python tools/probes/pcie_ceiling.py --ns 1,2,4,8 --seconds 1.0 --out ${P}_pcie_ceiling.json > ${P}_pcie_ceiling.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 8 > ${P}_bench_n8.json 2> ${P}_bench_n8.err; echo "bench8 rc=$?"
ORB_B200_EAGER_D2H=0 timeout 600 $TR bench.py --gpus 8 --no-other-shapes > ${P}_bench_n8_lazy.json 2> ${P}_bench_n8_lazy.err; echo "bench8 lazy rc=$?"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > ${P}_pytest_multi.txt 2>&1
tail -n 3 ${P}_pytest_multi.txt; cat ${P}_pcie_ceiling.txt | cut -c1-400
