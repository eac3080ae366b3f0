"""Latency form: one frame per orb_extract call (page-locked buffers), the way ORB_SLAM2::Frame drives the extractor.
Median / best wall-clock time per call for the shapes of the three datasets, in the environment it is started in
(ORB_B200_GRAPH=0|2 ...)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from orb_slam_system_b200 import ORBextractor
from orb_slam_system_b200.synth import synth_frame

for name, rows, cols, nf in (("kitti", 376, 1241, 2000), ("euroc", 480, 752, 1200), ("tum", 480, 640, 1000)):
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=2, max_rows=rows, max_cols=cols)
    cap = ex.keypoint_bound(rows, cols)
    img = torch.empty((1, rows, cols), dtype=torch.uint8, pin_memory=True)
    img.numpy()[0] = synth_frame(rows, cols, frame=3)
    kk = torch.empty((1, cap, 28), dtype=torch.uint8, pin_memory=True)
    dd = torch.empty((1, cap, 32), dtype=torch.uint8, pin_memory=True)
    cc = torch.empty((1,), dtype=torch.int32, pin_memory=True)
    for _ in range(20):
        ex.extract_batch_pinned(img, kk, dd, cc, cap)
    ts = []
    for _ in range(300):
        t0 = time.perf_counter()
        ex.extract_batch_pinned(img, kk, dd, cc, cap)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    tag = " ".join(f"{k}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("ORB_B200_"))
    print(f"[{tag or 'defaults'}] {name}: median {np.median(ts):.4f} ms  best {ts.min():.4f} ms  p90 {np.percentile(ts, 90):.4f} ms  keypoints {int(cc[0])}", flush=True)
    ex.close()
