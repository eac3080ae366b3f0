"""Device-resident extract rate (64 KITTI-shape frames per step) for the environment it is started in: median of 5
regions of 20 steps.  For A/B runs of run-time knobs (ORB_B200_LANES, ORB_B200_PRIO, ...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from orb_slam_system_b200 import ORBextractor
from orb_slam_system_b200.synth import synth_frame
B, R, ROWS, COLS = 64, 5, 376, 1241
pitch = (COLS + 63) // 64 * 64
frames = np.stack([synth_frame(ROWS, COLS, frame=f // 2, right=f & 1) for f in range(B)])
d_in = torch.zeros((R, B, ROWS, pitch), dtype=torch.uint8, device="cuda")
for r in range(R):
    d_in[r, :, :, :COLS] = torch.from_numpy(np.roll(frames, r, axis=0)).cuda()
ex = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=B, max_rows=ROWS, max_cols=COLS)
cap = ex.keypoint_bound(ROWS, COLS)
dk = torch.zeros((B, cap, 28), dtype=torch.uint8, device="cuda"); dd = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
dc = torch.zeros((B,), dtype=torch.int32, device="cuda")
st = torch.cuda.ExternalStream(ex.stream)
def step(i): ex.extract_batch_device(d_in[i % R][:, :, :COLS], dk, dd, dc, cap)
for i in range(5): step(i)
ex.sync()
if os.environ.get("LATE_ENV"):  # KEY=VALUE set only after the warm-up (timing probes that must start from filled buffers)
    k, v = os.environ["LATE_ENV"].split("=")
    os.environ[k] = v
ms = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(20): step(i)
    e1.record(st); ex.sync()
    ms.append(e0.elapsed_time(e1) / 20)
tag = " ".join(f"{k}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("ORB_B200_"))
print(f"[{tag or 'defaults'}] {np.median(ms):.4f} ms/step  {B / np.median(ms) * 1e3:.0f} frames/s  (min {min(ms):.4f} max {max(ms):.4f}) K={float(dc.float().mean()):.1f}", flush=True)
