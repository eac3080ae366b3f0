mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02ao_pytest.txt 2>&1
tail -n 4 gpurun_out/r02ao_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 600 python bench.py --no-other-shapes > gpurun_out/r02ao_bench.json 2> gpurun_out/r02ao_bench.err; tail -n 3 gpurun_out/r02ao_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ao_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['gpu_launches'], d['next_rows']['adapter_latency'], d['next_rows']['stereo_euroc']['ms_per_pair'])
PY
