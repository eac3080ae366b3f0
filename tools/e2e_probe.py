"""End-to-end timing of the async host API (submit/wait, `depth` batches in flight, pinned buffers) with the host
time spent inside submit and wait, for the settings given in ORB_B200_CHUNK / ORB_B200_LANES / ORB_B200_EAGER_D2H /
CUDA_DEVICE_MAX_CONNECTIONS (the latter must be in the environment before CUDA starts)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orb_slam_system_b200 import ORBextractor
from orb_slam_system_b200.synth import synth_frame
B, R, ROWS, COLS = 64, 3, 376, 1241
DEPTH = int(os.environ.get("DEPTH", "3"))
frames = np.stack([synth_frame(ROWS, COLS, frame=f // 2, right=f & 1) for f in range(B)])
pin = torch.empty((R, B, ROWS, COLS), dtype=torch.uint8, pin_memory=True)
for r in range(R): pin.numpy()[r] = np.roll(frames, r, axis=0)
ex = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=B, max_rows=ROWS, max_cols=COLS)
cap = ex.keypoint_bound(ROWS, COLS)
bufs = [(torch.empty((B, cap, 28), dtype=torch.uint8, pin_memory=True), torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True),
         torch.empty((B,), dtype=torch.int32, pin_memory=True)) for _ in range(DEPTH)]
def run(K):
    ts = tw = 0.0
    pend = []
    t0 = time.perf_counter()
    for i in range(K):
        a = time.perf_counter()
        pend.append(ex.submit_batch_pinned(pin[i % R], *bufs[i % DEPTH], cap))
        b = time.perf_counter()
        if len(pend) >= DEPTH:
            ex.wait_batch(pend.pop(0))
        c = time.perf_counter()
        ts += b - a; tw += c - b
    while pend:
        ex.wait_batch(pend.pop(0))
    dt = (time.perf_counter() - t0) / K
    return dt, ts / K, tw / K
run(6)
best = min(run(60) for _ in range(3))
dt, ts, tw = best
print(f"depth {DEPTH} eager {os.environ.get('ORB_B200_EAGER_D2H', '1')} conns {os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS', 'default')} "
      f"chunk {os.environ.get('ORB_B200_CHUNK', 'auto')} lanes {os.environ.get('ORB_B200_LANES', '2')}: {dt*1e3:.3f} ms/step {B/dt:.0f} frames/s; "
      f"host in submit {ts*1e3:.3f} ms, in wait {tw*1e3:.3f} ms", flush=True)
