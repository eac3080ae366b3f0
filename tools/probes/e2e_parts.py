"""Where does the end-to-end step time go?  H2D alone, D2H alone, both, and the pipelined API at depth 1..3."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from orb_slam_system_b200 import ORBextractor
from orb_slam_system_b200.synth import synth_frame
B, ROWS, COLS = 64, 376, 1241
ex = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=B, max_rows=ROWS, max_cols=COLS)
cap = ex.keypoint_bound(ROWS, COLS)
frames = np.stack([synth_frame(ROWS, COLS, frame=f // 2, right=f & 1) for f in range(B)])
R = 4
pin = torch.empty((R, B, ROWS, COLS), dtype=torch.uint8, pin_memory=True)
for r in range(R): pin.numpy()[r] = np.roll(frames, r, axis=0)
outs = [(torch.empty((B, cap, 28), dtype=torch.uint8, pin_memory=True), torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True),
         torch.empty((B,), dtype=torch.int32, pin_memory=True)) for _ in range(3)]
d = torch.empty((B, ROWS, COLS), dtype=torch.uint8, device="cuda")
dk = torch.empty((B, cap, 60), dtype=torch.uint8, device="cuda")
hk = torch.empty((B, 4480, 60), dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1): d.copy_(pin[0], non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hk.copy_(dk[:, :4480], non_blocking=True)
print("H2D 29.9 MB: %.3f ms" % timeit(h2d))
print("D2H 17.2 MB (2D): %.3f ms" % timeit(d2h))
print("both: %.3f ms" % timeit(lambda: (h2d(), d2h())))
for depth in (1, 2, 3):
    K = 30
    def run():
        pend = []
        for i in range(K):
            pend.append(ex.submit_batch_pinned(pin[i % R], *outs[i % 3], cap))
            if len(pend) >= depth: ex.wait_batch(pend.pop(0))
        while pend: ex.wait_batch(pend.pop(0))
    run(); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("depth %d: %.3f ms/step  %.0f frames/s" % (depth, dt / K * 1e3, B * K / dt))
# device only, lanes as configured
dkp = torch.empty((B, cap, 28), dtype=torch.uint8, device="cuda"); dd = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda"); dc = torch.empty((B,), dtype=torch.int32, device="cuda")
pitch = (COLS + 63) // 64 * 64
din = torch.zeros((B, ROWS, pitch), dtype=torch.uint8, device="cuda"); din[:, :, :COLS] = pin[0].cuda()
def dev():
    ex.extract_batch_device(din[:, :, :COLS], dkp, dd, dc, cap)
dev(); ex.sync()
t0 = time.perf_counter()
for _ in range(20): dev()
ex.sync(); print("device path: %.3f ms/step" % ((time.perf_counter() - t0) / 20 * 1e3))
# device path with an independent H2D stream running concurrently (is the slowdown physical?)
def dev_with_h2d(n=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(s1): d.copy_(pin[i % R], non_blocking=True)
        dev()
    ex.sync(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) / n * 1e3
dev_with_h2d(3)
print("device path + concurrent unrelated H2D: %.3f ms/step" % dev_with_h2d())
def dev_with_d2h(n=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(s2): hk.copy_(dk[:, :4480], non_blocking=True)
        dev()
    ex.sync(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) / n * 1e3
dev_with_d2h(3)
print("device path + concurrent unrelated D2H: %.3f ms/step" % dev_with_d2h())
