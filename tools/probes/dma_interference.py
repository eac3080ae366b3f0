"""Which kernel stages slow down while the copy engines are busy?  Per-stage event times of the resident path
without and with unrelated H2D / D2H transfers running on other streams."""
import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from orb_slam_system_b200 import ORBextractor
from orb_slam_system_b200.synth import synth_frame
B, ROWS, COLS = 64, 376, 1241
ex = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=B, max_rows=ROWS, max_cols=COLS)
cap = ex.keypoint_bound(ROWS, COLS)
frames = np.stack([synth_frame(ROWS, COLS, frame=f // 2, right=f & 1) for f in range(B)])
pitch = (COLS + 63) // 64 * 64
din = torch.zeros((B, ROWS, pitch), dtype=torch.uint8, device="cuda"); din[:, :, :COLS] = torch.from_numpy(frames).cuda()
dk = torch.empty((B, cap, 28), dtype=torch.uint8, device="cuda"); dd = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda"); dc = torch.empty((B,), dtype=torch.int32, device="cuda")
hin = torch.empty((B * ROWS * COLS,), dtype=torch.uint8, pin_memory=True); dland = torch.empty_like(hin, device="cuda")
dout = torch.empty((17_200_000,), dtype=torch.uint8, device="cuda"); hout = torch.empty((17_200_000,), dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(tag, h2d, d2h, K=30):
    stop = threading.Event()
    def pump():
        while not stop.is_set():
            if h2d:
                with torch.cuda.stream(s1): dland.copy_(hin, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
            s1.synchronize(); s2.synchronize()
    t = threading.Thread(target=pump); t.start()
    time.sleep(0.05)
    ex.set_profiling(True)
    for _ in range(K): ex.extract_batch_device(din[:, :, :COLS], dk, dd, dc, cap)
    st, n = ex.stage_times()
    ex.set_profiling(False)
    t0 = time.perf_counter()
    for _ in range(K): ex.extract_batch_device(din[:, :, :COLS], dk, dd, dc, cap)
    ex.sync(); dt = (time.perf_counter() - t0) / K * 1e3
    stop.set(); t.join()
    print(f"{tag:10s} step {dt:.3f} ms  " + " ".join(f"{k}={v / n:.3f}" for k, v in st.items()))
for _ in range(3): ex.extract_batch_device(din[:, :, :COLS], dk, dd, dc, cap)
ex.sync()
run("quiet", False, False)
run("H2D", True, False)
run("D2H", False, True)
run("both", True, True)
