"""Latency of one call through the C ABI with preallocated pinned buffers: 1 frame, a stereo pair (2 frames), 8 frames."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from orb_slam_system_b200 import ORBextractor
from orb_slam_system_b200.synth import synth_frame
for (rows, cols, nf, name) in ((376, 1241, 2000, "KITTI"), (480, 752, 1200, "EuRoC")):
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=8, max_rows=rows, max_cols=cols)
    cap = ex.keypoint_bound(rows, cols)
    for n in (1, 2, 8):
        frames = torch.from_numpy(np.stack([synth_frame(rows, cols, frame=f // 2, right=f & 1) for f in range(n)])).pin_memory()
        kps = torch.empty((n, cap, 28), dtype=torch.uint8).pin_memory()
        desc = torch.empty((n, cap, 32), dtype=torch.uint8).pin_memory()
        cnt = torch.empty((n,), dtype=torch.int32).pin_memory()
        for _ in range(5): ex.extract_batch_pinned(frames, kps, desc, cnt, cap)
        ts = []
        for _ in range(50):
            t0 = time.perf_counter(); ex.extract_batch_pinned(frames, kps, desc, cnt, cap); ts.append(time.perf_counter() - t0)
        ts.sort()
        print(f"{name} {n} frame(s): median {ts[25] * 1e3:.3f} ms  min {ts[0] * 1e3:.3f} ms  ({int(cnt[0])} keypoints in frame 0)")
    ex.close()
