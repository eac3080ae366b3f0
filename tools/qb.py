"""Quick bench summary: python tools/qb.py [tag] -- runs bench.py without the CPU / side rows and prints the headline numbers."""
import json, subprocess, sys, os
tag = sys.argv[1] if len(sys.argv) > 1 else ""
out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--no-next-rows"] + sys.argv[2:], capture_output=True, text=True)
try:
    j = json.loads(out.stdout.strip().splitlines()[-1])
    st = j["roofline"]["stage_ms_per_step"]
    print(f"{tag:14s} value {j['value']:9.0f} f/s  {j['ms_per_step']:.4f} ms  e2e {j['e2e']['value']:9.0f} sync {j['e2e'].get('sync_call_value', 0):9.0f}  stages " +
          " ".join(f"{k}={v:.3f}" for k, v in st.items()) + f" sum={sum(st.values()):.3f} launches {j['gpu_launches']} host_enq {j.get('host_enqueue_ms_per_step', 0):.3f} ms  hamming {j['hamming']['value'] / 1e9:.0f} G pairs/s")
except Exception as e:
    print(tag, "FAILED", e, out.stdout[-500:], out.stderr[-1500:])
