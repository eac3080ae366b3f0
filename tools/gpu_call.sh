mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02u_pytest.txt 2>&1
tail -n 3 gpurun_out/r02u_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/r02u_bench_default.json 2> gpurun_out/r02u_bench_default.err; tail -n 2 gpurun_out/r02u_bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/r02u_bench_reference.json 2> gpurun_out/r02u_bench_reference.err; tail -c 600 gpurun_out/r02u_bench_reference.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02u_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['steady_state_value'], d['e2e']['platform_ceiling_frames_s'], d['roofline']['issue']['frac'], d['hamming']['value'], d['cpu_baseline']['value'])
PY
