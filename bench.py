#!/usr/bin/env python3
"""Benchmark of the B200-native ORB front end (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the extract + describe hot path over one batch of synthetic
KITTI-shape frames (BASELINE config 3: 1241x376, 2000 features, 8 levels, 64 frames per GPU).
`value` = frames/s with the frames already resident in HBM (device-timed, CUDA events on the
extractor's stream, max over ranks).  `e2e` = the same metric through the C-ABI host-buffer
call orb_extract_batch: pinned host frames in, keypoints + descriptors back in host memory,
copies inside the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS, NFEAT, NLEVELS = 376, 1241, 2000, 8
SCALE, INI_TH, MIN_TH = 1.2, 20, 7
METRIC = "ORB frames/s @KITTI 1241x376 2k feats"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def pin_to_gpu_numa_node(index):
    """Run this rank on the CPUs NVML lists as local to its GPU, so that the pinned host buffers (first touch) and the
    thread that feeds the copy engines sit on the GPU's own NUMA node.  Best effort: any failure leaves the affinity alone."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every few milliseconds while the timed
    regions run (same counters as `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*`)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: report it instead of inventing numbers
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        masks = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, m in masks.items():
                    if r & m:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self._stop.set()
        self._t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def level_pixels():
    """P_l of the BASELINE shape."""
    sc = [1.0, 1.0]
    acc = np.float32(1.0)
    for _ in range(NLEVELS - 2):
        acc = np.float32(np.float64(acc) * np.float64(np.float32(SCALE)))
        sc.append(float(acc))
    P = []
    for s in sc:
        inv = np.float32(1.0) / np.float32(s)
        P.append(int(np.rint(np.float32(COLS) * inv)) * int(np.rint(np.float32(ROWS) * inv)))
    return P


def make_frames(n, first):
    from orb_slam_system_b200.synth import synth_frame
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        # stereo pairs: even = left, odd = right image of the same scene
        frames = list(ex.map(lambda f: synth_frame(ROWS, COLS, seed=7, frame=(first + f) // 2, right=(first + f) & 1), range(n)))
    return np.stack(frames)


def next_rows(device):
    """The rows of SURVEY 8f / the search methods of 8a served through the C ABI, each timed through its public
    host-buffer call (copies inside) next to the oracle on one host core: stereo (BASELINE config 2), windowed search,
    DBoW2 transform.  Small, bounded samples: a few seconds in total."""
    import oracle
    from orb_slam_system_b200 import KP_DTYPE, FrameView, ORBextractor, ORBmatcher, ORBVocabulary
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from voc_cases import make_vocabulary

    def best_of(fn, reps):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    out = {}
    # ---- BASELINE config 2: EuRoC-shape stereo pair, 1200 features per image, extraction of both images in one
    # batch + Frame::ComputeStereoMatches whole (Hamming search, SAD refinement on the resident pyramids, median cut)
    rows, cols, nf = 480, 752, 1200
    bf, fx = 47.90639384423901, 435.2046959714599
    il = oracle.synth_frame(rows, cols, frame=1)
    ir = oracle.synth_frame(rows, cols, frame=1, right=1)
    ex = ORBextractor(nf, SCALE, NLEVELS, INI_TH, MIN_TH, max_batch=2, device=device)
    m = ORBmatcher(0.6, True, device=device)
    pair = np.stack([il, ir])
    res = ex.extract_batch(pair)
    (kl, dl), (kr, dr) = res

    # caller-owned pinned buffers, as the C++ adapter keeps them (no allocation inside the timed call)
    import torch
    cap2 = ex.keypoint_bound(rows, cols)
    p_in = torch.from_numpy(pair).pin_memory()
    p_k = torch.empty((2, cap2, 28), dtype=torch.uint8).pin_memory()
    p_d = torch.empty((2, cap2, 32), dtype=torch.uint8).pin_memory()
    p_c = torch.empty((2,), dtype=torch.int32).pin_memory()

    def gpu_extract_pair():
        ex.extract_batch_pinned(p_in, p_k, p_d, p_c, cap2)

    def gpu_pair():
        gpu_extract_pair()
        c0, c1 = int(p_c[0]), int(p_c[1])
        return m.ComputeStereoMatches(ex, ex, p_k[0, :c0].numpy().view(KP_DTYPE).ravel(), p_d[0, :c0].numpy(),
                                      p_k[1, :c1].numpy().view(KP_DTYPE).ravel(), p_d[1, :c1].numpy(), bf, fx, 0, 1)

    t_extract_pair = best_of(gpu_extract_pair, 20)
    t_pair = best_of(gpu_pair, 10)
    t_stereo = best_of(lambda: m.ComputeStereoMatches(ex, ex, kl, dl, kr, dr, bf, fx, 0, 1), 10)
    ur, _ = m.ComputeStereoMatches(ex, ex, kl, dl, kr, dr, bf, fx, 0, 1)
    t0 = time.perf_counter()
    okl, odl = oracle.extract(il, nfeatures=nf, cap=16000)
    okr, odr = oracle.extract(ir, nfeatures=nf, cap=16000)
    t_cpu_ex = time.perf_counter() - t0
    t0 = time.perf_counter()
    our, _ = oracle.compute_stereo_matches(il, ir, okl, odl, okr, odr, bf, fx)
    t_cpu_st = time.perf_counter() - t0
    out["stereo_euroc"] = {
        "workload": "752x480 stereo pair, 1200 features per image: extract both + Frame::ComputeStereoMatches (BASELINE config 2)",
        "pairs_per_s": 1.0 / t_pair, "ms_per_pair": 1e3 * t_pair, "extract_pair_ms": 1e3 * t_extract_pair, "stereo_match_ms": 1e3 * t_stereo,
        "keypoints_left": int(len(kl)), "stereo_matches": int((ur >= 0).sum()), "identical_to_oracle": bool(ur.tobytes() == our.tobytes()),
        "cpu_ms_per_pair": 1e3 * (t_cpu_ex + t_cpu_st), "cpu_stereo_match_ms": 1e3 * t_cpu_st, "cpu_cores": 1,
        "note": "one pair per call through the host API (latency form); the frames/s metric above is the batched form"}
    # ---- windowed search: SearchByProjection(Frame, MapPoints) over the frame grid built on the device
    F = FrameView(kl, dl, 0, cols, 0, rows, scale_factors=ex.GetScaleFactors())
    rng = np.random.default_rng(0)
    nq = 2000
    src = rng.integers(0, len(kr), nq)
    q = dr[src]
    u = kr["x"][src] + rng.normal(0, 3, nq).astype(np.float32) + 10
    v = kr["y"][src] + rng.normal(0, 2, nq).astype(np.float32)
    lvl = kr["octave"][src].astype(np.int32)
    vc = np.full(nq, 0.9, np.float32)
    occ0 = np.zeros(len(kl), np.uint8)
    t_sp = best_of(lambda: m.SearchByProjection(F, occ0.copy(), q, u, v, u, lvl, vc, 3.0), 10)
    occ_g, occ_o = occ0.copy(), occ0.copy()
    n_g, f_g = m.SearchByProjection(F, occ_g, q, u, v, u, lvl, vc, 3.0)
    t0 = time.perf_counter()
    n_o, f_o = oracle.search_by_projection_map(F, occ_o, q, u, v, u, lvl, vc, 3.0, 0.6)
    t_cpu_sp = time.perf_counter() - t0
    out["search_by_projection"] = {"workload": f"{nq} projected map points against a {len(kl)}-keypoint frame, th = 3",
                                   "ms_per_call": 1e3 * t_sp, "queries_per_s": nq / t_sp, "matches": int(n_g),
                                   "identical_to_oracle": bool(n_g == n_o and (f_g == f_o).all()), "cpu_ms_per_call": 1e3 * t_cpu_sp,
                                   "cpu_cores": 1}
    # ---- DBoW2 transform (Frame::ComputeBoW): ORBvoc-shaped synthetic tree, k = 10, L = 5 (111 111 nodes), levelsup 4
    voc_a = make_vocabulary(np.random.default_rng(1), 10, 5)
    voc = ORBVocabulary(voc_a.child_off, voc_a.children, voc_a.node_desc, voc_a.node_weight, voc_a.node_word, 10, 5, device=device)
    t_voc = best_of(lambda: voc.transform(dl, 4), 10)
    g = voc.transform(dl, 4)
    t0 = time.perf_counter()
    o = oracle.voc_transform(voc_a, dl, 4)
    t_cpu_voc = time.perf_counter() - t0
    out["dbow2_transform"] = {"workload": f"{len(dl)} descriptors through a k=10, L=5 vocabulary ({len(voc_a.parent)} nodes), levelsup 4",
                              "ms_per_call": 1e3 * t_voc, "features_per_s": len(dl) / t_voc,
                              "identical_to_oracle": bool((g["bow_ids"] == o["bow_ids"]).all() and g["bow_values"].tobytes() == o["bow_values"].tobytes()
                                                          and g["fv"] == o["fv"]),
                              "cpu_ms_per_call": 1e3 * t_cpu_voc, "cpu_cores": 1}
    voc.close()
    # ---- image ingest: raw colour EuRoC-shape frames -> cv::remap (rectification maps) -> cv::cvtColor
    # gray, fused into the level-0 load, + extraction; 16 frames per call through the host-buffer API
    nfr = 16
    raw = np.stack([np.stack([oracle.synth_frame(rows, cols, frame=50 + f, seed=7 + c) for c in range(3)], -1) for f in range(nfr)])
    xs, ys = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    xn, yn = (xs - cols / 2) / (0.6 * cols), (ys - rows / 2) / (0.6 * cols)
    fr = 1 - 0.12 * (xn * xn + yn * yn)
    mx = (xn * fr * 0.6 * cols + cols / 2 + 3.25).astype(np.float32)
    my = (yn * fr * 0.6 * cols + rows / 2 - 2.5).astype(np.float32)
    exi = ORBextractor(nf, SCALE, NLEVELS, INI_TH, MIN_TH, max_batch=nfr, device=device)
    exi.set_ingest((rows, cols, 3), maps=(mx, my))
    t_ing = best_of(lambda: exi.ingest_extract_batch(raw), 5)
    gi = exi.ingest_extract_batch(raw)
    level0 = exi.pyramid_level(0)  # the reference's mImGray of frame 0
    t0 = time.perf_counter()
    gray0 = oracle.cvt_gray(oracle.remap_linear(raw[0], mx, my))
    t_cpu_ing = time.perf_counter() - t0
    t0 = time.perf_counter()
    ok0, od0 = oracle.extract(gray0, nfeatures=nf, cap=16000)
    t_cpu_ex0 = time.perf_counter() - t0
    plain = np.ascontiguousarray(raw[..., 1])
    t_plain = best_of(lambda: exi.extract_batch(plain), 5)
    same = len(gi[0][0]) == len(ok0) and all(np.array_equal(gi[0][0][k], ok0[k]) for k in ("x", "y", "size", "angle", "response", "octave"))
    out["ingest"] = {"workload": f"{nfr} raw 752x480x3 frames: remap (CV_32FC1 maps, INTER_LINEAR) + RGB2GRAY fused into the level-0 load, then extraction",
                     "frames_per_s": nfr / t_ing, "ms_per_call": 1e3 * t_ing, "ms_per_call_gray_frames_no_ingest": 1e3 * t_plain,
                     "identical_to_oracle": bool(same and np.array_equal(gi[0][1], od0) and np.array_equal(level0, gray0)),
                     "cpu_ms_per_frame": 1e3 * (t_cpu_ing + t_cpu_ex0), "cpu_ingest_ms_per_frame": 1e3 * t_cpu_ing, "cpu_cores": 1}
    exi.close()
    ex.close()
    m.close()
    return out


def cpu_reference_rate(nframes, threads):
    """The reference's own CPU extractor (oracle/_ref, compiled from the reference sources
    against oracle/cvshim) when present, else the oracle port; one frame per host thread."""
    import oracle
    if oracle.ref_lib() is not None:
        secs, kp = oracle.ref_extract_many(ROWS, COLS, nframes, threads, nfeatures=NFEAT, scaleFactor=SCALE, nlevels=NLEVELS,
                                           iniThFAST=INI_TH, minThFAST=MIN_TH)
        kind = "reference"
    else:
        secs, kp = oracle.extract_many(ROWS, COLS, nframes, threads, nfeatures=NFEAT, scaleFactor=SCALE, nlevels=NLEVELS,
                                       iniThFAST=INI_TH, minThFAST=MIN_TH)
        kind = "port"
    return nframes / secs, kind, kp


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(8, cores)  # bounded sample: one frame per host thread per step
    for _ in range(args.warmup):
        cpu_reference_rate(per_step, cores)
    t0 = time.perf_counter()
    kind, secs = "port", 0.0
    for _ in range(args.steps):
        rate, kind, _ = cpu_reference_rate(per_step, cores)
        secs += per_step / rate  # extraction time only; frame synthesis is outside the clock
    el = time.perf_counter() - t0
    value = args.steps * per_step / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * per_step / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "KITTI-shape stereo 1241x376, 2000 features, 8 levels, scale 1.2, FAST 20/7 (BASELINE config 3)",
                   "frames_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{per_step} synthetic frames per step, one frame per host thread, {args.steps} steps"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": el,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=64)
    ap.add_argument("--rotate", type=int, default=5, help="distinct input batches cycled through (5 x 30 MB > L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the stereo / search / vocabulary side measurements")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    pin_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from orb_slam_system_b200 import KP_DTYPE, ORBextractor, kernel_launch_count

    B, R, K, W = args.frames_per_gpu, args.rotate, args.steps, max(args.warmup, 3)
    ex = ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, max_batch=B, device=local_rank, max_rows=ROWS, max_cols=COLS)
    cap = ex.keypoint_bound(ROWS, COLS)
    pitch = (COLS + 63) // 64 * 64

    # ---- synthetic input: R distinct batches per rank, frames differ across ranks
    host = make_frames(B * R, first=rank * B * R).reshape(R, B, ROWS, COLS)
    pinned_in = torch.empty((R, B, ROWS, COLS), dtype=torch.uint8, pin_memory=True)
    pinned_in.numpy()[:] = host
    d_in = torch.zeros((R, B, ROWS, pitch), dtype=torch.uint8, device="cuda")
    d_in[:, :, :, :COLS] = pinned_in.cuda()
    d_kps = torch.zeros((B, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
    h_kps = torch.empty((B, cap, 28), dtype=torch.uint8, pin_memory=True)
    h_desc = torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True)
    h_counts = torch.empty((B,), dtype=torch.int32, pin_memory=True)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ex.stream, device=torch.device("cuda", local_rank))

    def step_device(i):
        ex.extract_batch_device(d_in[i % R][:, :, :COLS], d_kps, d_desc, d_counts, cap)

    def step_host(i):
        ex.extract_batch_pinned(pinned_in[i % R], h_kps, h_desc, h_counts, cap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for i in range(W):
        step_device(i)
    ex.sync()
    k_mean = float(d_counts.float().mean().item())
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    th0 = time.perf_counter()
    for i in range(K):
        step_device(W + i)
    host_enqueue_ms = 1e3 * (time.perf_counter() - th0) / K
    e1.record(stream)
    ex.sync()
    barrier()
    launches = kernel_launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    # same K steps again with CUDA events around every kernel stage (stages back to back on one
    # stream, so each duration is the kernel's own): per-kernel launch times for the roofline
    ex.set_profiling(True)
    for i in range(K):
        step_device(W + i)
    stage_ms, ncalls = ex.stage_times()
    ex.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-buffer C ABI (e2e): pinned host frames in, keypoints + descriptors
    # back in pinned host memory, every step.  (a) the synchronous call orb_extract_batch, one step at a
    # time; (b) the asynchronous pair orb_extract_batch_submit / _wait with three steps in flight (the uploads of
    # steps i+1, i+2 overlap step i's kernels, step i's download overlaps step i+1's kernels) -- the serving form,
    # reported as e2e; both move the same bytes per step inside the timed region.
    for i in range(W):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        step_host(W + i)
    barrier()
    e2e_sync_s = time.perf_counter() - t0
    h_kps2 = torch.empty_like(h_kps).pin_memory()
    h_desc2 = torch.empty_like(h_desc).pin_memory()
    h_counts2 = torch.empty_like(h_counts).pin_memory()
    DEPTH = 3  # ORB_MAX_IN_FLIGHT
    outs = [(h_kps, h_desc, h_counts), (h_kps2, h_desc2, h_counts2),
            (torch.empty_like(h_kps).pin_memory(), torch.empty_like(h_desc).pin_memory(), torch.empty_like(h_counts).pin_memory())]

    def submit(i):
        o = outs[i % DEPTH]
        return ex.submit_batch_pinned(pinned_in[i % R], o[0], o[1], o[2], cap)

    for i in range(W):
        ex.wait_batch(submit(i))
    barrier()
    sampler2 = ClockSampler(local_rank)
    if rank == 0:
        sampler2.start()
    t0 = time.perf_counter()
    pending = [submit(i) for i in range(min(DEPTH - 1, K))]
    for i in range(len(pending), K):
        pending.append(submit(i))
        ex.wait_batch(pending.pop(0))
    while pending:
        ex.wait_batch(pending.pop(0))
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_clocks = sampler2.stop() if rank == 0 else None
    maxc = int(h_counts.max().item())
    h2d = B * ROWS * COLS
    d2h = B * 4 + B * maxc * 28 + B * maxc * 32

    # ---- second half of the metric: brute-force Hamming matching (BASELINE config 4): 256 keyframe
    # pairs of 2000 x 2000 descriptors drawn from the extractor's own output, device resident
    from orb_slam_system_b200 import ORBmatcher
    NP, NQ = 256, 2000
    m = ORBmatcher(0.6, True, device=local_rank)
    pool = d_desc[:, :NQ, :].contiguous()  # [B, 2000, 32] real descriptors of the last step
    qsel = torch.arange(NP, device="cuda") % B
    tsel = (torch.arange(NP, device="cuda") + 1) % B
    dq, dt = pool[qsel].contiguous(), pool[tsel].contiguous()
    nq = torch.full((NP,), NQ, dtype=torch.int32, device="cuda")
    nt = torch.full((NP,), NQ, dtype=torch.int32, device="cuda")
    obi = torch.empty((NP, NQ), dtype=torch.int32, device="cuda")
    obd = torch.empty_like(obi)
    osd = torch.empty_like(obi)
    torch.cuda.synchronize()
    mstream = torch.cuda.ExternalStream(m.stream, device=torch.device("cuda", local_rank))
    for _ in range(3):
        m.match_all_batch_device(dq, nq, dt, nt, obi, obd, osd)
    m.sync()
    barrier()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    MREP = 10
    m0.record(mstream)
    for _ in range(MREP):
        m.match_all_batch_device(dq, nq, dt, nt, obi, obd, osd)
    m1.record(mstream)
    m.sync()
    barrier()
    match_ms = m0.elapsed_time(m1) / MREP

    # ---- BASELINE config 5 (N > 1): NCCL all-gather of every rank's keyframe descriptor block over NVLink,
    # then each rank brute-force matches its local frames against the same-index keyframes of the next rank
    shard = None
    if world > 1:
        from orb_slam_system_b200.sharding import all_gather_descriptors, cross_shard_pairs
        KF = 2000  # descriptor rows exchanged per keyframe (fixed stride, counts travel with them)
        block = d_desc[:, :KF, :].contiguous()
        cnts = torch.clamp(d_counts, max=KF).to(torch.int32)
        for _ in range(2):
            all_d, all_c = all_gather_descriptors(block, cnts)
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        GREP = 5
        for _ in range(GREP):
            all_d, all_c = all_gather_descriptors(block, cnts)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1) / GREP
        pairs = torch.from_numpy(cross_shard_pairs(B, rank, world, neighbours=1)).cuda()
        tq = block[pairs[:, 0]].contiguous()
        tt = all_d[pairs[:, 1]].contiguous()
        tnq = cnts[pairs[:, 0]].contiguous()
        tnt = all_c[pairs[:, 1]].contiguous()
        sb = torch.empty((len(pairs), KF), dtype=torch.int32, device="cuda")
        sd1, sd2 = torch.empty_like(sb), torch.empty_like(sb)
        torch.cuda.synchronize()
        for _ in range(2):
            m.match_all_batch_device(tq, tnq, tt, tnt, sb, sd1, sd2)
        m.sync()
        dist.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(mstream)
        for _ in range(GREP):
            m.match_all_batch_device(tq, tnq, tt, tnt, sb, sd1, sd2)
        s1.record(mstream)
        m.sync()
        smatch_ms = s0.elapsed_time(s1) / GREP
        tt2 = torch.tensor([gather_ms, smatch_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt2, op=dist.ReduceOp.MAX)
        gather_ms, smatch_ms = float(tt2[0].item()), float(tt2[1].item())
        recv_bytes = (world - 1) * (block.numel() + cnts.numel() * 4)
        npairs_total = float((tnq.double() * tnt.double()).sum().item())
        tp = torch.tensor([npairs_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(tp)
        shard = {"workload": f"all-gather of {B} keyframes x {KF} x 32 B per rank + cross-shard brute-force matching (BASELINE config 5)",
                 "allgather_ms": gather_ms, "allgather_recv_GBps_per_gpu": recv_bytes / (gather_ms * 1e-3) / 1e9,
                 "match_ms": smatch_ms, "match_pairs_per_s": float(tp[0].item()) / (smatch_ms * 1e-3)}

    if world > 1:
        t = torch.tensor([ms_total, e2e_s, match_ms, e2e_sync_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, match_ms, e2e_sync_s = float(t[0].item()), float(t[1].item()), float(t[2].item()), float(t[3].item())

    if rank == 0:
        frames = B * K * world
        value = frames / (ms_total * 1e-3)
        peak, peak_kind = measured_peaks()
        P = level_pixels()
        Psum = sum(P)
        # dominant stage and its algorithmic bytes per launch (DESIGN.md "roofline")
        dom = max(stage_ms, key=stage_ms.get)
        per_launch_ms = stage_ms[dom] / max(ncalls, 1)
        stage_bytes = {
            "pyramid": (Psum - P[-1]) + (Psum - P[0]),
            "detect": Psum,
            "octree": 0,
            "blur": 2 * Psum,
            "describe": 60 * k_mean,
        }
        algo = stage_bytes[dom] * B
        achieved = algo / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        path_bytes = (5 * Psum - P[0] - P[-1] + 60 * k_mean) * B
        path_gbs = path_bytes / (ms_total / K * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "KITTI-shape stereo 1241x376, 2000 features, 8 levels, scale 1.2, FAST 20/7, 64 frames/GPU batch (BASELINE config 3)",
                       "frames_per_gpu_per_step": B, "keypoints_per_frame": k_mean,
                       "l2": f"inputs rotate through {R} distinct batches ({R * B * ROWS * pitch / 1e6:.0f} MB > 126 MB L2)"},
            "roofline": {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of k_detect, ncu --set full (profiles/r01j: 2 x 32-frame
                         # launches of 43.25 + 0.53 MB make one 64-frame step)
                         "traffic": 87.6e6 if dom == "detect" else None, "traffic_source": "profiles/r01j_ncu_summary.txt",
                         "peak_source": peak_kind,
                         "algo_bytes_per_launch": algo, "launch_ms": per_launch_ms,
                         "stage_ms_per_step": {k: v / max(ncalls, 1) for k, v in stage_ms.items()},
                         "stage_note": "stage times from a second pass of the same steps with the stages serialised on one stream; "
                                       "in the timed region the blur overlaps detect + octree on a second stream",
                         "path": {"algo_bytes_per_step": path_bytes, "achieved": path_gbs, "frac": path_gbs / peak}},
            "e2e": {"value": frames / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "orb_extract_batch_submit/_wait, three 64-frame steps in flight, pinned host buffers",
                    "sync_call_value": frames / e2e_sync_s, "sync_call_api": "orb_extract_batch, one step at a time",
                    "clocks": e2e_clocks},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clocks,
            "hamming": {"metric": "Hamming pairs/s (brute force, best + second best)", "value": world * NP * NQ * NQ / (match_ms * 1e-3),
                        "unit": "pairs/s", "workload": f"{NP} keyframe pairs x {NQ} x {NQ} descriptors per GPU (BASELINE config 4)",
                        "ms_per_launch": match_ms,
                        "algo_bytes_per_launch": NP * (32 * 2 * NQ + 12 * NQ),
                        "hbm_frac": NP * (32 * 2 * NQ + 12 * NQ) / (match_ms * 1e-3) / 1e9 / peak,
                        "popc_per_s": world * 3 * NP * NQ * NQ / (match_ms * 1e-3),
                        "logic_ops_per_s": world * 19 * NP * NQ * NQ / (match_ms * 1e-3),
                        # pipe peaks per GPU from tools/probes/pipe_probe: POPC 16, LOP3 / VIMNMX 64 lanes per clock and SM
                        "popc_pipe_frac": 3 * NP * NQ * NQ / (match_ms * 1e-3) / (16 * 148 * 1.965e9),
                        "logic_pipe_frac": 19 * NP * NQ * NQ / (match_ms * 1e-3) / (64 * 148 * 1.965e9),
                        "note": "logic-pipe bound: the 6 live XOR words of a pair (words 6-7 of this fork's descriptors are zero, checked on "
                                "the data; 8 otherwise) go through a LOP3 carry-save tree to 3 POPC instead of 6, ~19 logic-pipe "
                                "instructions per pair against a pipe rate of 64 lanes/clk/SM = 0.97e12 pairs/s per GPU"},
        }
        if shard is not None:
            line["shard_match"] = shard
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            nfr = max(64, 8 * cores)  # ~1 s of wall clock, 10-20 s of CPU work on 16 cores
            rate, kind, _ = cpu_reference_rate(nfr, cores)
            line["cpu_baseline"] = {"value": rate, "unit": "frames/s", "cores": cores, "kind": kind,
                                    "sample": f"{nfr} synthetic KITTI-shape frames, one frame per host thread"}
            if not args.no_next_rows:
                try:
                    line["next_rows"] = next_rows(local_rank)
                except Exception as e:  # side measurements must never cost the headline line
                    line["next_rows"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    ex.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
