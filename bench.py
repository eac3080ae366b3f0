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

NLEVELS, SCALE, INI_TH, MIN_TH = 8, 1.2, 20, 7
DEPTH = 3  # ORB_MAX_IN_FLIGHT: host batches in flight in the end-to-end regions
# The three shapes BASELINE.json / SURVEY 8a name (settings files of the reference); kitti is the headline (config 3).
SHAPES = {
    "kitti": dict(rows=376, cols=1241, nfeat=2000, metric="ORB frames/s @KITTI 1241x376 2k feats",
                  label="KITTI-shape stereo 1241x376, 2000 features, 8 levels, scale 1.2, FAST 20/7, 64 frames/GPU batch (BASELINE config 3; "
                        "Examples/Stereo/KITTI00-02.yaml:38-51)"),
    "tum": dict(rows=480, cols=640, nfeat=1000, metric="ORB frames/s @TUM 640x480 1k feats",
                label="TUM1-shape monocular 640x480, 1000 features, 8 levels, scale 1.2, FAST 20/7, 64 frames/GPU batch (BASELINE config 1 "
                      "batched; Examples/Monocular/TUM1.yaml:30-43)"),
    "euroc": dict(rows=480, cols=752, nfeat=1200, metric="ORB frames/s @EuRoC 752x480 1.2k feats", bf=47.90639384423901, fx=435.2046959714599,
                  label="EuRoC-shape stereo 752x480, 1200 features per image, 8 levels, scale 1.2, FAST 20/7, 64 stereo pairs = 128 images/GPU "
                        "batch (BASELINE config 2 batched; Examples/Stereo/EuRoC.yaml:88-101)"),
}


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


def slice_cpus(local_rank, world):
    """With several ranks on one box, give each its own contiguous slice of the CPUs the process may run on, so that the
    ranks' submit / wait threads (which poll CUDA events) do not migrate over each other.  Best effort."""
    if world <= 1:
        return
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = len(cpus) // world
        if per >= 1 and len(cpus) > per:
            os.sched_setaffinity(0, set(cpus[local_rank * per:(local_rank + 1) * per]))
    except Exception:
        pass


def level_pixels(shape="kitti"):
    """P_l of a BASELINE shape."""
    ROWS, COLS = SHAPES[shape]["rows"], SHAPES[shape]["cols"]
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


def make_frames(shape, n, first):
    from orb_slam_system_b200.synth import synth_frame
    ROWS, COLS = SHAPES[shape]["rows"], SHAPES[shape]["cols"]
    with ThreadPoolExecutor(max_workers=min(8, len(os.sched_getaffinity(0)) or 1)) as ex:
        # stereo pairs: even = left, odd = right image of the same scene
        frames = list(ex.map(lambda f: synth_frame(ROWS, COLS, seed=7, frame=(first + f) // 2, right=(first + f) & 1), range(n)))
    return np.stack(frames)


def euroc_stereo(sb, m, K, barrier, max_over_ranks, world):
    """BASELINE config 2 batched: a step = 64 EuRoC-shape stereo pairs = both images of every pair extracted in one batch
    (left = even, right = odd frames) + Frame::ComputeStereoMatches for all 64 pairs in one orb_compute_stereo_matches_batch
    call (row-band table, Hamming search, SAD refinement on the pyramids the extraction left on the device, median cut:
    four launches over all pairs).  resident: frames in HBM, results stay in HBM.  e2e: pinned host frames in, keypoints +
    descriptors + counts + mvuRight + mvDepth of every pair back in pinned host memory, copies inside the clock."""
    import torch
    S, B = sb.S, sb.B
    npairs = B // 2
    cap = sb.cap
    d_ur = torch.empty((npairs, cap), dtype=torch.float32, device="cuda")
    d_dep = torch.empty_like(d_ur)
    d_st = torch.zeros((npairs,), dtype=torch.int32, device="cuda")
    mstream = torch.cuda.ExternalStream(m.stream)

    def step_resident(i):
        sb.step_device(i)
        m.ComputeStereoMatchesBatchDevice(sb.ex, sb.d_kps, sb.d_desc, sb.d_counts, cap, S["bf"], S["fx"], d_ur, d_dep, d_st)

    for i in range(3):
        step_resident(i)
    m.sync()
    assert int(d_st.abs().sum().item()) == 0, "a synthetic pair was refused"
    counts = sb.d_counts.cpu().numpy()
    w = int(counts.max())
    w = min(cap, (w + w // 32 + 63) // 64 * 64)  # rows copied back per frame: the largest count of the warm-up + 3 %
    valid = torch.arange(cap, device="cuda")[None, :] < sb.d_counts[0::2, None]  # rows of pair p up to its left count
    matches = float(((d_ur >= 0) & valid).sum().item()) / npairs
    Ks = max(3, min(K, 10))
    res = []
    for _ in range(3):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sb.stream)
        for i in range(Ks):
            step_resident(i)
        e1.record(mstream)
        m.sync()
        res.append(e0.elapsed_time(e1) * 1e-3)
    res_s = median(max_over_ranks(res))
    # ---- end to end: one step at a time (upload, extract, stereo, download)
    d_land = torch.empty((B, S["rows"], S["cols"]), dtype=torch.uint8, device="cuda")
    h_k = torch.empty((B, w, 28), dtype=torch.uint8, pin_memory=True)
    h_d = torch.empty((B, w, 32), dtype=torch.uint8, pin_memory=True)
    h_c = torch.empty((B,), dtype=torch.int32, pin_memory=True)
    h_ur = torch.empty((npairs, w), dtype=torch.float32, pin_memory=True)
    h_dep = torch.empty((npairs, w), dtype=torch.float32, pin_memory=True)

    ts = torch.cuda.current_stream()  # torch's own stream for the copies (the library's streams die with their handles)

    def step_e2e(i):
        d_land.copy_(sb.pinned_in[i % sb.R], non_blocking=True)
        ts.synchronize()
        sb.ex.extract_batch_device(d_land, sb.d_kps, sb.d_desc, sb.d_counts, cap)  # dense frames: the library re-pitches them
        m.ComputeStereoMatchesBatchDevice(sb.ex, sb.d_kps, sb.d_desc, sb.d_counts, cap, S["bf"], S["fx"], d_ur, d_dep, d_st)
        m.sync()
        h_k.copy_(sb.d_kps[:, :w].contiguous(), non_blocking=True)
        h_d.copy_(sb.d_desc[:, :w].contiguous(), non_blocking=True)
        h_c.copy_(sb.d_counts, non_blocking=True)
        h_ur.copy_(d_ur[:, :w].contiguous(), non_blocking=True)
        h_dep.copy_(d_dep[:, :w].contiguous(), non_blocking=True)
        ts.synchronize()

    step_e2e(0)
    e2e = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        for i in range(Ks):
            step_e2e(i)
        e2e.append(time.perf_counter() - t0)
    e2e_s = median(max_over_ranks(e2e))
    return {"workload": f"{npairs} stereo pairs per step: extraction of both images + Frame::ComputeStereoMatches for all pairs "
                        "(orb_compute_stereo_matches_batch on the device-resident extraction results)",
            "pairs_per_s": world * npairs * Ks / res_s, "ms_per_step": 1e3 * res_s / Ks,
            "e2e_pairs_per_s": world * npairs * Ks / e2e_s, "e2e_ms_per_step": 1e3 * e2e_s / Ks,
            "e2e_note": "one step at a time: pinned frames up, extraction, stereo matching, then keypoints + descriptors + counts + "
                        f"mvuRight + mvDepth down ({w} rows per frame)",
            "steps": Ks, "stereo_matches_per_pair": matches}


def hamming_cpu_baseline(ham_sets, cores):
    """The reference's own DescriptorDistance (src/ORBmatcher.cc:896-908, compiled from source: oracle/_ref/libref_match.so)
    in the brute-force best / second-best scan over the same 2000 x 2000 descriptor pairs, one keyframe pair per host thread."""
    import oracle
    if oracle.ref_match_lib() is None:
        return {"unavailable": "oracle/_ref/libref_match.so not built"}
    dq, dt = ham_sets
    n = max(2, min(cores, dq.shape[0]))
    q = dq[:n].cpu().numpy()
    t = dt[:n].cpu().numpy()
    secs, bi, bd, sd = oracle.ref_bruteforce_many(q, t, cores)
    return {"value": n * q.shape[1] * t.shape[1] / secs, "unit": "pairs/s", "cores": cores, "kind": "reference",
            "sample": f"{n} keyframe pairs of {q.shape[1]} x {t.shape[1]} descriptors, one pair per host thread"}


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
    # ---- one KITTI frame at a time through the C++ drop-in class ORB_SLAM2::ORBextractor::operator() (adapter/ORBextractor_b200.cc
    # compiled against the cv:: stand-in, oracle/Makefile adapter), one persistent instance, host cv::Mat in, std::vector<cv::KeyPoint>
    # + cv::Mat out: with mvImagePyramid refilled after every call (what the reference's Frame::ComputeStereoMatches reads) and without
    try:
        import ctypes as C
        from orb_slam_system_b200 import ORBextractor as _E
        lib_path = os.path.join(ROOT, "oracle", "_ref", "libadapter_orb.so")
        if os.path.exists(lib_path):
            A = C.CDLL(lib_path)
            A.ref_extract_many.restype = C.c_double
            A.ref_extract_many.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(C.c_longlong)]
            tot = C.c_longlong(0)
            row = {"workload": "one 1241x376 frame per call through ORB_SLAM2::ORBextractor::operator() (C++ adapter class, persistent instance, 50 calls)"}
            for key, env in (("ms_per_frame_with_mvImagePyramid", "1"), ("ms_per_frame_without_mvImagePyramid", "0")):
                os.environ["ORB_B200_IMAGE_PYRAMID"] = env
                A.ref_extract_many(2000, SCALE, NLEVELS, INI_TH, MIN_TH, 376, 1241, 7, 0, 10, 1, C.byref(tot))
                secs = A.ref_extract_many(2000, SCALE, NLEVELS, INI_TH, MIN_TH, 376, 1241, 7, 0, 50, 1, C.byref(tot))
                row[key] = 1e3 * secs / 50
            os.environ.pop("ORB_B200_IMAGE_PYRAMID", None)
            exl = _E(2000, SCALE, NLEVELS, INI_TH, MIN_TH, max_batch=1, device=device)
            capl = exl.keypoint_bound(376, 1241)
            p_img = torch.from_numpy(oracle.synth_frame(376, 1241, frame=0)[None]).pin_memory()
            p_kk = torch.empty((1, capl, 28), dtype=torch.uint8).pin_memory()
            p_dd = torch.empty((1, capl, 32), dtype=torch.uint8).pin_memory()
            p_cc = torch.empty((1,), dtype=torch.int32).pin_memory()
            row["ms_per_frame_c_abi_orb_extract"] = 1e3 * best_of(lambda: exl.extract_batch_pinned(p_img, p_kk, p_dd, p_cc, capl), 50)
            row["note"] = ("the class is compiled against the cv:: stand-in of this repo (no OpenCV in the image): its scalar cv::copyMakeBorder "
                           "and cv::Mat allocations dominate the refill of mvImagePyramid; ORBextractor::KeepImagePyramid(false) removes it")
            exl.close()
            out["adapter_latency"] = row
    except Exception as e:
        out["adapter_latency"] = {"error": repr(e)}
    return out


def cpu_reference_rate(nframes, threads, shape="kitti", fast=False):
    """The reference's own CPU extractor (oracle/_ref, compiled from the reference sources against oracle/cvshim) when
    present, else the oracle port; one frame per host thread.  fast: the timing-only -O3 build (oracle/Makefile reffast)
    instead of the -O2 -ffp-contract=off parity build."""
    import oracle
    S = SHAPES[shape]
    kw = dict(nfeatures=S["nfeat"], scaleFactor=SCALE, nlevels=NLEVELS, iniThFAST=INI_TH, minThFAST=MIN_TH)
    if fast and oracle.ref_lib() is not None and oracle.ref_fast_lib() is not None:
        secs, kp = oracle.ref_extract_many(S["rows"], S["cols"], nframes, threads, fast=True, **kw)
        kind = "reference-O3"
    elif oracle.ref_lib() is not None:
        secs, kp = oracle.ref_extract_many(S["rows"], S["cols"], nframes, threads, **kw)
        kind = "reference"
    else:
        secs, kp = oracle.extract_many(S["rows"], S["cols"], nframes, threads, **kw)
        kind = "port"
    return nframes / secs, kind, kp


CPU_NOTE = ("the reference's ORBextractor.cc compiled from source against oracle/cvshim: cv::FAST / resize / GaussianBlur are scalar "
            "models of OpenCV's arithmetic, not OpenCV's SIMD code (OpenCV is not in this image), so this arm is roughly 2x slower than "
            "the reference linked against a real OpenCV 3.4 build would be; kind reference-O3 = -O3 -march=x86-64-v3 "
            "(the reference's CMakeLists.txt:10-11 uses -O3 -march=native), kind reference = the -O2 -ffp-contract=off parity build")


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(8, cores)  # bounded sample: one frame per host thread per step
    for _ in range(args.warmup):
        cpu_reference_rate(per_step, cores, args.workload, fast=True)
    t0 = time.perf_counter()
    kind, secs = "port", 0.0
    for _ in range(args.steps):
        rate, kind, _ = cpu_reference_rate(per_step, cores, args.workload, fast=True)
        secs += per_step / rate  # extraction time only; frame synthesis is outside the clock
    el = time.perf_counter() - t0
    value = args.steps * per_step / secs
    parity_rate, parity_kind, _ = cpu_reference_rate(per_step, cores, args.workload, fast=False)
    S = SHAPES[args.workload]
    line = {
        "impl": "reference", "metric": S["metric"], "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * per_step / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": S["label"], "frames_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{per_step} synthetic frames per step, one frame per host thread, {args.steps} steps",
                         "parity_build": {"value": parity_rate, "kind": parity_kind}, "note": CPU_NOTE},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": el,
    }
    print(json.dumps(line), flush=True)


def newest_profile():
    """The newest profiles/*_kernels.json (tools/ncu_kernels_json.py on an `ncu --set full` capture of this bench): per-frame
    DRAM bytes and executed warp instructions per kernel."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_kernels.json")))
    if not files:
        return None, None
    try:
        return json.load(open(files[-1])), os.path.relpath(files[-1], ROOT)
    except Exception:
        return None, None


def median(v):
    return float(np.median(np.asarray(v, np.float64)))


class ShapeBench:
    """One extractor handle on one workload shape: device-resident and host-buffer (end-to-end) timing of batches of B frames."""

    def __init__(self, shape, B, rank, local_rank, torch):
        self.t = torch
        S = SHAPES[shape]
        self.S, self.B, self.shape = S, B, shape
        rows, cols = S["rows"], S["cols"]
        from orb_slam_system_b200 import ORBextractor
        self.ex = ORBextractor(S["nfeat"], SCALE, NLEVELS, INI_TH, MIN_TH, max_batch=B, device=local_rank, max_rows=rows, max_cols=cols)
        self.cap = self.ex.keypoint_bound(rows, cols)
        self.pitch = (cols + 63) // 64 * 64
        # enough distinct batches that the inputs of consecutive steps cannot sit in the 126 MB L2
        self.R = R = max(3, -(-140_000_000 // (B * rows * self.pitch)))
        host = make_frames(shape, B * R, first=rank * B * R).reshape(R, B, rows, cols)
        self.pinned_in = torch.empty((R, B, rows, cols), dtype=torch.uint8, pin_memory=True)
        self.pinned_in.numpy()[:] = host
        self.d_in = torch.zeros((R, B, rows, self.pitch), dtype=torch.uint8, device="cuda")
        self.d_in[:, :, :, :cols] = self.pinned_in.cuda()
        cap = self.cap
        self.d_kps = torch.zeros((B, cap, 28), dtype=torch.uint8, device="cuda")
        self.d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
        self.d_counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
        self.outs = [(torch.empty((B, cap, 28), dtype=torch.uint8, pin_memory=True), torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True),
                      torch.empty((B,), dtype=torch.int32, pin_memory=True)) for _ in range(DEPTH)]
        torch.cuda.synchronize()
        self.stream = torch.cuda.ExternalStream(self.ex.stream, device=torch.device("cuda", local_rank))

    def close(self):
        self.ex.close()

    def step_device(self, i):
        self.ex.extract_batch_device(self.d_in[i % self.R][:, :, :self.S["cols"]], self.d_kps, self.d_desc, self.d_counts, self.cap)

    def resident(self, K, reps, barrier):
        """reps regions of K device-resident steps, each bracketed by barrier + synchronize; ms per region, host enqueue ms per step."""
        t = self.t
        out, enq = [], []
        for _ in range(reps):
            barrier()
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            th0 = time.perf_counter()
            for i in range(K):
                self.step_device(i)
            enq.append(1e3 * (time.perf_counter() - th0) / K)
            e1.record(self.stream)
            self.ex.sync()
            barrier()
            out.append(e0.elapsed_time(e1))
        return out, median(enq)

    def stage_profile(self, K):
        self.ex.set_profiling(True)
        for i in range(K):
            self.step_device(i)
        stage_ms, ncalls = self.ex.stage_times()
        self.ex.set_profiling(False)
        return stage_ms, ncalls

    def submit(self, i):
        o = self.outs[i % DEPTH]
        return self.ex.submit_batch_pinned(self.pinned_in[i % self.R], o[0], o[1], o[2], self.cap)

    def e2e_sync(self, K, barrier):
        o = self.outs[0]
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            self.ex.extract_batch_pinned(self.pinned_in[i % self.R], o[0], o[1], o[2], self.cap)
        return time.perf_counter() - t0

    def e2e_async(self, K, reps, barrier, after_wait=None):
        """reps regions of K steps through orb_extract_batch_submit / _wait with DEPTH steps in flight.  Every region starts
        after a barrier + synchronize and ends when this rank's last result is in host memory; the clock is the rank's own
        (no collective inside it) and the caller takes the maximum over ranks."""
        out = []
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            pending = []
            for i in range(K):
                pending.append((i, self.submit(i)))
                if len(pending) >= DEPTH:
                    j, tk = pending.pop(0)
                    self.ex.wait_batch(tk)
                    if after_wait:
                        after_wait(j)
            while pending:
                j, tk = pending.pop(0)
                self.ex.wait_batch(tk)
                if after_wait:
                    after_wait(j)
            out.append(time.perf_counter() - t0)
        return out

    def e2e_steady(self, K, regions):
        """One continuous run of (regions + 1) * K steps; seconds of each K-step window between result arrivals, i.e.
        the serving rate without the fill and drain of the three-deep pipeline."""
        stamps, pending = [], []
        n = (regions + 1) * K
        for i in range(n):
            pending.append(self.submit(i))
            if len(pending) >= DEPTH:
                self.ex.wait_batch(pending.pop(0))
                stamps.append(time.perf_counter())
        while pending:
            self.ex.wait_batch(pending.pop(0))
            stamps.append(time.perf_counter())
        return [stamps[(r + 1) * K - 1] - stamps[r * K - 1] for r in range(1, regions + 1)]

    def algo_bytes_per_frame(self, k_mean):
        P = level_pixels(self.shape)
        Psum = sum(P)
        return {"pyramid": (Psum - P[-1]) + (Psum - P[0]), "detect": Psum, "octree": 0, "blur": 2 * Psum, "describe": 60 * k_mean,
                "path": 5 * Psum - P[0] - P[-1] + 60 * k_mean}


def platform_ceiling(sb, seconds, barrier, torch):
    """What the host <-> device link of this box carries when every rank moves exactly one step's bytes in each direction
    concurrently from / to pinned memory with no kernel at all: steps per second of this rank (tools/probes/pcie_ceiling.py
    is the stand-alone form).  The e2e figure cannot exceed the sum over ranks of this number x frames per step."""
    B, S = sb.B, sb.S
    h2d = B * S["rows"] * S["cols"]
    maxc = int(sb.outs[0][2].max().item())
    d2h = B * maxc * 60 + 4 * B
    d_land = torch.empty(h2d, dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(d2h, dtype=torch.uint8, device="cuda")
    h_res = torch.empty(d2h, dtype=torch.uint8, pin_memory=True)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    flat = [sb.pinned_in[r].reshape(-1) for r in range(sb.R)]
    barrier()
    t0 = time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < seconds:
        for i in range(4):
            with torch.cuda.stream(s_in):
                d_land.copy_(flat[(steps + i) % sb.R], non_blocking=True)
            with torch.cuda.stream(s_out):
                h_res.copy_(d_res, non_blocking=True)
        steps += 4
        s_in.synchronize()
        s_out.synchronize()
    return steps / (time.perf_counter() - t0), h2d, d2h


def hamming_block(m, pool, mstream, barrier, world, peak_hbm, torch):
    """Second half of the metric (BASELINE config 4): 256 keyframe pairs of 2000 x 2000 descriptors, device resident.  The
    tensor-core kernel on the extractor's own descriptors (182 live bits) and on 256-live-bit rows, the POPC kernel as
    the comparator of both."""
    NP, NQ = 256, 2000
    B = pool.shape[0]
    qsel = torch.arange(NP, device="cuda") % B
    tsel = (torch.arange(NP, device="cuda") + 1) % B
    sets = {"182": (pool[qsel].contiguous(), pool[tsel].contiguous())}
    g = torch.Generator(device="cuda").manual_seed(11)
    noise = torch.randint(0, 256, (NP + 1, NQ, 32), dtype=torch.uint8, device="cuda", generator=g)
    sets["256"] = (noise[:NP].contiguous(), noise[1:].contiguous())
    nq = torch.full((NP,), NQ, dtype=torch.int32, device="cuda")
    outs = {impl: [torch.empty((NP, NQ), dtype=torch.int32, device="cuda") for _ in range(3)] for impl in ("mma", "popc")}
    res = {}
    MREP = 10
    saved = os.environ.get("ORB_B200_MATCH")
    for live, (dq, dt) in sets.items():
        for impl in ("mma", "popc"):
            os.environ["ORB_B200_MATCH"] = impl
            o = outs[impl]
            for _ in range(2):
                m.match_all_batch_device(dq, nq, dt, nq, *o)
            m.sync()
            barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record(mstream)
            for _ in range(MREP if impl == "mma" else 3):
                m.match_all_batch_device(dq, nq, dt, nq, *o)
            m1.record(mstream)
            m.sync()
            barrier()
            res[f"{impl}_{live}_ms"] = m0.elapsed_time(m1) / (MREP if impl == "mma" else 3)
        res[f"identical_{live}"] = all(bool((a == b).all().item()) for a, b in zip(outs["mma"], outs["popc"]))
    if saved is None:
        os.environ.pop("ORB_B200_MATCH", None)
    else:
        os.environ["ORB_B200_MATCH"] = saved
    return NP, NQ, res, sets["182"]


def config5(sb, m, mstream, rank, world, torch, dist):
    """BASELINE config 5 as SURVEY 8(e) states it: 4096 KITTI-shape frames over 8 ranks = 512 per rank, extracted as 8
    launches of 64 straight into the rank's keyframe block (fixed stride Kmax x 32 B per frame + an int32 count per
    frame), ONE all-gather of the blocks over NCCL / NVLink, then every rank brute-force matches its local frames against
    the gathered keyframes of every other rank (same frame index: world - 1 keyframes per local frame)."""
    from orb_slam_system_b200.sharding import all_gather_descriptors
    F, KMAX = 512, 4480  # the KITTI-shape octree keeps at most nIni * 4^p = 4480 keypoints per frame
    B = sb.B
    launches = F // B
    block = torch.zeros((F, KMAX, 32), dtype=torch.uint8, device="cuda")
    cnts = torch.zeros((F,), dtype=torch.int32, device="cuda")
    kps = torch.zeros((B, KMAX, 28), dtype=torch.uint8, device="cuda")

    def extract_all():
        for s in range(launches):
            sb.ex.extract_batch_device(sb.d_in[s % sb.R][:, :, :sb.S["cols"]], kps, block[s * B:(s + 1) * B], cnts[s * B:(s + 1) * B], KMAX)

    extract_all()
    sb.ex.sync()
    assert int(cnts.max().item()) <= KMAX, "a frame kept more than Kmax keypoints"
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sb.stream)
    extract_all()
    e1.record(sb.stream)
    sb.ex.sync()
    extract_ms = e0.elapsed_time(e1)
    for _ in range(2):
        all_d, all_c = all_gather_descriptors(block, cnts)
    torch.cuda.synchronize()
    dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    GREP = 5
    g0.record()
    for _ in range(GREP):
        all_d, all_c = all_gather_descriptors(block, cnts)
    g1.record()
    torch.cuda.synchronize()
    gather_ms = g0.elapsed_time(g1) / GREP
    sb_ = torch.empty((world - 1, F, KMAX), dtype=torch.int32, device="cuda")
    sd1, sd2 = torch.empty_like(sb_), torch.empty_like(sb_)

    def match_all_neighbours():
        for d in range(1, world):
            o = (rank + d) % world
            m.match_all_batch_device(block, cnts, all_d[o * F:(o + 1) * F], all_c[o * F:(o + 1) * F], sb_[d - 1], sd1[d - 1], sd2[d - 1])

    match_all_neighbours()
    m.sync()
    dist.barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(mstream)
    match_all_neighbours()
    s1.record(mstream)
    m.sync()
    match_ms = s0.elapsed_time(s1)
    # ---- checks, outside every clock: (1) every rank's block arrived intact (a checksum travels beside it);
    # (2) one cross-shard pair against the CPU oracle's brute-force scan
    w = (torch.arange(32, device="cuda", dtype=torch.int64) + 1)
    mine = (block.to(torch.int64) * w).sum().reshape(1)
    sums = torch.empty((world,), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(sums, mine)
    got = torch.stack([(all_d[r * F:(r + 1) * F].to(torch.int64) * w).sum() for r in range(world)])
    gathered_ok = bool((got == sums).all().item()) and bool((all_c.view(world, F)[rank] == cnts).all().item())
    oracle_ok = None
    if rank == 0:
        try:
            import oracle
            o = 1 % world
            nq0, nt0 = int(cnts[0].item()), int(all_c[o * F].item())
            oi, od, os_ = oracle.match_all(block[0, :nq0].cpu().numpy(), all_d[o * F, :nt0].cpu().numpy())
            oracle_ok = bool((sb_[0, 0, :nq0].cpu().numpy() == oi).all() and (sd1[0, 0, :nq0].cpu().numpy() == od).all()
                             and (sd2[0, 0, :nq0].cpu().numpy() == os_).all())
        except Exception as e:
            oracle_ok = repr(e)
    npairs = float((cnts.double().unsqueeze(0) * torch.stack([all_c[((rank + d) % world) * F:((rank + d) % world + 1) * F] for d in range(1, world)]).double()).sum().item())
    t = torch.tensor([extract_ms, gather_ms, match_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tp = torch.tensor([npairs, 1.0 if gathered_ok else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(tp)
    recv = (world - 1) * (block.numel() + cnts.numel() * 4)
    extract_ms, gather_ms, match_ms = (float(x) for x in t.tolist())
    return {"workload": f"{world * F} KITTI-shape frames = {F} per rank ({launches} launches of {B} over the rank's {sb.R * B} resident synthetic frames, cycled); "
                        f"keyframe block {F} x Kmax {KMAX} x 32 B + counts per rank; one all-gather; every local frame against the same-index "
                        f"keyframe of each of the {world - 1} other ranks (BASELINE config 5, SURVEY 8e)",
            "extract_ms": extract_ms, "extract_frames_per_s": world * F / (extract_ms * 1e-3),
            "allgather_ms": gather_ms, "allgather_bytes_per_rank": block.numel() + cnts.numel() * 4,
            "allgather_recv_GBps_per_gpu": recv / (gather_ms * 1e-3) / 1e9, "nvlink_peak_GBps_per_direction": 900.0,
            "allgather_frac_of_nvlink": recv / (gather_ms * 1e-3) / 1e9 / 900.0,
            "match_ms": match_ms, "match_frame_pairs": world * (world - 1) * F, "match_pairs_per_s": float(tp[0].item()) / (match_ms * 1e-3),
            "gathered_blocks_match_their_checksums_on_all_ranks": bool(tp[1].item() == world),
            "cross_shard_pair_identical_to_oracle": oracle_ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti", choices=sorted(SHAPES), help="shape of the headline line (default: BASELINE config 3)")
    ap.add_argument("--frames-per-gpu", type=int, default=64)
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of --steps steps each; the line reports the median")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the stereo / search / vocabulary side measurements")
    ap.add_argument("--no-other-shapes", action="store_true", help="skip the TUM / EuRoC rows")
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
    torch.set_num_threads(1)
    pin_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from orb_slam_system_b200 import ORBmatcher, kernel_launch_count

    B, K, W, REPS = args.frames_per_gpu, args.steps, max(args.warmup, 3), max(1, args.repeats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sum_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    def measure_shape(shape, Bs, reps, full):
        """Resident and end-to-end rates of one workload shape (max over ranks per region, median over regions)."""
        sb = ShapeBench(shape, Bs, rank, local_rank, torch)
        slice_cpus(local_rank, world)  # after the frame synthesis (which uses a thread pool)
        for i in range(W):
            sb.step_device(i)
        sb.ex.sync()
        k_mean = float(sb.d_counts.float().mean().item())
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        barrier()
        l0 = kernel_launch_count()
        res_ms, enq_ms = sb.resident(K, reps, barrier)
        launches = (kernel_launch_count() - l0) // reps
        res_ms = max_over_ranks(res_ms)
        stage_ms, ncalls = sb.stage_profile(K) if full else ({}, 0)
        clocks = sampler.stop() if rank == 0 else None
        # ---- end to end through the host-buffer C ABI: pinned host frames in, keypoints + descriptors back in pinned
        # host memory, every step.  (a) the synchronous call orb_extract_batch, one step at a time; (b) the asynchronous
        # pair orb_extract_batch_submit / _wait with three steps in flight -- the serving form, reported as e2e; both
        # move the same bytes per step inside the timed region.
        for i in range(W):
            sb.ex.wait_batch(sb.submit(i))
        sync_s = max_over_ranks([sb.e2e_sync(K, barrier)])[0] if full else None
        sampler2 = ClockSampler(local_rank)
        if rank == 0:
            sampler2.start()
        e2e_s = max_over_ranks(sb.e2e_async(K, reps, barrier))
        barrier()
        steady_s = max_over_ranks(sb.e2e_steady(K, reps))
        e2e_clocks = sampler2.stop() if rank == 0 else None
        ceil_steps, h2d, d2h = platform_ceiling(sb, 0.25, barrier, torch)
        ceil_total = sum_over_ranks([ceil_steps])[0]
        frames = Bs * K * world
        r = {"sb": sb, "k_mean": k_mean, "res_ms": res_ms, "enq_ms": enq_ms, "launches": int(launches), "stage_ms": stage_ms, "ncalls": ncalls,
             "clocks": clocks, "e2e_clocks": e2e_clocks, "sync_s": sync_s, "e2e_s": e2e_s, "steady_s": steady_s, "frames": frames,
             "h2d": h2d, "d2h": d2h, "ceiling_frames_s": ceil_total * Bs}
        return r

    def shape_row(shape, r, peak):
        """The compact row of a non-headline shape."""
        sb = r["sb"]
        ab = sb.algo_bytes_per_frame(r["k_mean"])
        ms = median(r["res_ms"])
        value = r["frames"] / (ms * 1e-3)
        e2e = r["frames"] / median(r["e2e_s"])
        gbs = ab["path"] * r["frames"] / world / (ms * 1e-3) / 1e9  # per GPU: the peak it is compared with is one GPU's
        return {"workload": sb.S["label"], "frames_per_gpu_per_step": sb.B, "keypoints_per_frame": r["k_mean"],
                "value": value, "unit": "frames/s", "ms_per_step": ms / K, "value_min_max": [r["frames"] / (max(r["res_ms"]) * 1e-3), r["frames"] / (min(r["res_ms"]) * 1e-3)],
                "e2e": {"value": e2e, "steady_state_value": r["frames"] / median(r["steady_s"]), "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                        "platform_ceiling_frames_s": r["ceiling_frames_s"], "frac_of_platform_ceiling": e2e / r["ceiling_frames_s"]},
                "algo_bytes_per_frame": ab["path"], "path_GBps_per_gpu": gbs, "hbm_frac": gbs / peak}

    peak, peak_kind = measured_peaks()
    head = measure_shape(args.workload, B, REPS, True)
    sb = head["sb"]

    # ---- second half of the metric: brute-force Hamming matching (BASELINE config 4)
    m = ORBmatcher(0.6, True, device=local_rank)
    mstream = torch.cuda.ExternalStream(m.stream, device=torch.device("cuda", local_rank))
    sb.step_device(0)
    sb.ex.sync()
    pool = sb.d_desc[:, :2000, :].contiguous()  # real descriptors of one step
    NP, NQ, ham, ham_sets = hamming_block(m, pool, mstream, barrier, world, peak, torch)
    ham_ms = max_over_ranks([ham[k] for k in ("mma_182_ms", "mma_256_ms", "popc_182_ms", "popc_256_ms")])

    shard = config5(sb, m, mstream, rank, world, torch, dist) if (world > 1 and args.workload == "kitti") else None

    others = {}
    if not args.no_other_shapes:
        for shape in ("tum", "euroc", "kitti"):
            if shape == args.workload:
                continue
            Bo = 128 if shape == "euroc" else 64  # EuRoC: 64 stereo pairs = 128 images per step
            r = measure_shape(shape, Bo, 3, False)
            row = shape_row(shape, r, peak)
            if shape == "euroc":
                row["stereo"] = euroc_stereo(r["sb"], m, K, barrier, max_over_ranks, world)
            others[shape] = row
            r["sb"].close()

    if rank == 0:
        S = sb.S
        frames = head["frames"]
        ms_med = median(head["res_ms"])
        value = frames / (ms_med * 1e-3)
        ab = sb.algo_bytes_per_frame(head["k_mean"])
        stage_ms, ncalls = head["stage_ms"], max(head["ncalls"], 1)
        dom = max(stage_ms, key=stage_ms.get)
        per_launch_ms = stage_ms[dom] / ncalls
        algo = ab[dom] * B
        achieved = algo / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        path_gbs = ab["path"] * B / (ms_med / K * 1e-3) / 1e9
        prof, prof_path = newest_profile() if args.workload == "kitti" else (None, None)
        kname = {"pyramid": "k_resize_tile", "detect": "k_detect", "octree": "k_octree_fast", "blur": "k_blur", "describe": "k_describe_tile"}[dom]
        traffic = issue = None
        if prof and prof.get("per_frame"):
            pf = prof["per_frame"]
            if kname in pf:  # per launch of the dominant stage = one 64-frame step's worth
                traffic = (pf[kname]["dram_read_bytes"] + pf[kname]["dram_write_bytes"]) * B
            sm_hz = ((head["clocks"] or {}).get("sm_mhz") or 1965.0) * 1e6
            inst = prof["per_frame_total"]["warp_inst"] * B
            peak_issue = 148 * 4 * sm_hz  # one warp instruction per SM sub-partition and clock
            issue = {"warp_instr_per_step": inst, "peak_warp_instr_per_s": peak_issue, "floor_ms_per_step": 1e3 * inst / peak_issue,
                     "frac": inst / (ms_med / K * 1e-3) / peak_issue, "source": prof_path,
                     "note": "smsp__inst_executed.sum of every kernel of the step (ncu) / measured step time, against 148 SMs x 4 issue "
                             "slots x the SM clock sampled during the run: the path is bound by instruction issue, not by HBM"}
        e2e_med = median(head["e2e_s"])
        e2e_val = frames / e2e_med
        line = {
            "metric": S["metric"], "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_med / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": S["label"], "frames_per_gpu_per_step": B, "keypoints_per_frame": head["k_mean"],
                       "l2": f"inputs rotate through {sb.R} distinct batches ({sb.R * B * S['rows'] * sb.pitch / 1e6:.0f} MB > 126 MB L2)",
                       "repeats": f"{REPS} timed regions of {K} steps, each bracketed by barrier + synchronize; value, ms_per_step and e2e are "
                                  "the median region (max over ranks per region)"},
            "value_min_max": [frames / (max(head["res_ms"]) * 1e-3), frames / (min(head["res_ms"]) * 1e-3)],
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": prof_path, "peak_source": peak_kind,
                         "algo_bytes_per_launch": algo, "launch_ms": per_launch_ms,
                         "stage_ms_per_step": {k: v / ncalls for k, v in stage_ms.items()},
                         "stage_note": "stage times from a second pass of the same steps with the stages serialised on one stream; "
                                       "in the timed region the blur overlaps detect + octree on a second stream",
                         "path": {"algo_bytes_per_step": ab["path"] * B, "achieved": path_gbs, "frac": path_gbs / peak},
                         "issue": issue},
            "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": head["d2h"],
                    "api": "orb_extract_batch_submit/_wait, three 64-frame steps in flight, pinned host buffers; every region starts after a "
                           "barrier + synchronize and ends when the rank's last result is in host memory (the rank's own clock, max over ranks)",
                    "value_min_max": [frames / max(head["e2e_s"]), frames / min(head["e2e_s"])],
                    "steady_state_value": frames / median(head["steady_s"]),
                    "steady_state_note": "K-step windows between result arrivals inside one continuous run: the serving rate without the "
                                         "fill and drain of the three-deep pipeline that every bracketed region pays",
                    "platform_ceiling_frames_s": head["ceiling_frames_s"], "frac_of_platform_ceiling": e2e_val / head["ceiling_frames_s"],
                    "platform_ceiling_note": "all ranks copying one step's bytes host -> device and device -> host concurrently from pinned memory "
                                             "with no kernels (same run, same box): no e2e figure on this box can exceed it",
                    "sync_call_value": frames / head["sync_s"], "sync_call_api": "orb_extract_batch, one step at a time",
                    "clocks": head["e2e_clocks"]},
            "gpu_launches": head["launches"] * K // K, "host_enqueue_ms_per_step": head["enq_ms"], "clocks": head["clocks"],
        }
        line["gpu_launches"] = head["launches"]
        pairs = NP * NQ * NQ
        mma182, mma256, popc182, popc256 = ham_ms
        # tensor pipe: one tcgen05.mma 128 x 128 x 32 (8-bit operands) per 64 clocks and SM (B300_MICROARCH: max(M, 128) * N / 256)
        kfrac = lambda ms, ksteps: (pairs / (128 * 128)) * ksteps * 64 / (148 * 1.965e9) / (ms * 1e-3)
        line["hamming"] = {
            "metric": "Hamming pairs/s (brute force, best + second best)", "value": world * pairs / (mma182 * 1e-3), "unit": "pairs/s",
            "workload": f"{NP} keyframe pairs x {NQ} x {NQ} descriptors per GPU (BASELINE config 4), the extractor's own descriptors (182 live bits)",
            "kernel": "k_match_mma3 (persistent; tcgen05.mma kind::i8 128x128x32, accumulators in TMEM, in-kernel bit expansion)",
            "ms_per_launch": mma182, "value_256_live_bits": world * pairs / (mma256 * 1e-3), "ms_per_launch_256_live_bits": mma256,
            "popc_kernel_value": world * pairs / (popc182 * 1e-3), "popc_kernel_value_256_live_bits": world * pairs / (popc256 * 1e-3),
            "identical_to_popc_kernel": bool(ham["identical_182"] and ham["identical_256"]),
            "algo_bytes_per_launch": NP * (32 * 2 * NQ + 12 * NQ), "hbm_frac": NP * (32 * 2 * NQ + 12 * NQ) / (mma182 * 1e-3) / 1e9 / peak,
            "mma_pipe_frac": kfrac(mma182, 7), "mma_pipe_frac_256_live_bits": kfrac(mma256, 9),
            "mma_ops_per_s": world * pairs * 2 * 192 / (mma182 * 1e-3),
            "note": "d(a, b) = |a| - a'.b with a' = 2a - 1 in {-1, +1}, b in {0, 1}: exact in int8 with int32 accumulation.  K = 192 (6 K-steps) "
                    "when words 6-7 of a train tile are zero (checked on the data), else 256, plus one K-step that adds the key's bias and "
                    "column so that the accumulator is the packed 16-bit key.  mma_pipe_frac = issued tcgen05.mma time / kernel time at "
                    "64 clocks per 128 x 128 x 32 instruction (7 / 9 K-steps); the rest is the epilogue (tcgen05.ld, one IMAD and two "
                    "ALU-pipe instructions per key pair, both half-rate pipes) and the bit expansion, which share the SM with the tensor pipe"}
        if shard is not None:
            line["shard_match"] = shard
        if others:
            line["other_shapes"] = others
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            nfr = max(64, 8 * cores)  # ~1 s of wall clock, 10-20 s of CPU work on 16 cores
            rate, kind, _ = cpu_reference_rate(nfr, cores, args.workload, fast=True)
            rate2, kind2, _ = cpu_reference_rate(nfr // 2, cores, args.workload, fast=False)
            line["cpu_baseline"] = {"value": rate, "unit": "frames/s", "cores": cores, "kind": kind,
                                    "sample": f"{nfr} synthetic {args.workload}-shape frames, one frame per host thread",
                                    "parity_build": {"value": rate2, "kind": kind2}, "note": CPU_NOTE}
            try:
                line["hamming"]["cpu_baseline"] = hamming_cpu_baseline(ham_sets, cores)
            except Exception as e:
                line["hamming"]["cpu_baseline"] = {"error": repr(e)}
            if not args.no_next_rows:
                try:
                    line["next_rows"] = next_rows(local_rank)
                except Exception as e:  # side measurements must never cost the headline line
                    line["next_rows"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    sb.close()
    m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
