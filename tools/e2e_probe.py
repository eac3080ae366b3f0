"""Quick end-to-end timing of orb_extract_batch with pinned host buffers (no device-resident pass)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orb_slam_system_b200 import ORBextractor
from orb_slam_system_b200.synth import synth_frame
B, R, ROWS, COLS = 64, 3, 376, 1241
frames = np.stack([synth_frame(ROWS, COLS, frame=f // 2, right=f & 1) for f in range(B)])
pin = torch.empty((R, B, ROWS, COLS), dtype=torch.uint8, pin_memory=True)
for r in range(R): pin.numpy()[r] = np.roll(frames, r, axis=0)
ex = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=B, max_rows=ROWS, max_cols=COLS)
cap = ex.keypoint_bound(ROWS, COLS)
hk = torch.empty((B, cap, 28), dtype=torch.uint8, pin_memory=True)
hd = torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True)
hc = torch.empty((B,), dtype=torch.int32, pin_memory=True)
for i in range(5): ex.extract_batch_pinned(pin[i % R], hk, hd, hc, cap)
t0 = time.perf_counter(); K = 60
for i in range(K): ex.extract_batch_pinned(pin[i % R], hk, hd, hc, cap)
dt = (time.perf_counter() - t0) / K
print(os.environ.get("ORB_B200_CHUNK", "16"), os.environ.get("ORB_B200_LANES", "2"), f"{dt*1e3:.3f} ms/step  {B/dt:.0f} frames/s")
