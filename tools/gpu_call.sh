mkdir -p gpurun_out
P=gpurun_out/r02r
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > ${P}_bench_n8.json 2> ${P}_bench_n8.err
timeout 200 python bench.py --no-cpu-baseline --no-next-rows --no-other-shapes > ${P}_bench_n1_samebox.json 2> ${P}_bench_n1_samebox.err
tail -c 1500 ${P}_bench_n8.json; tail -n 3 ${P}_bench_n8.err; tail -c 300 ${P}_bench_n1_samebox.json
