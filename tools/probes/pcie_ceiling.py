#!/usr/bin/env python3
"""Platform ceiling of the end-to-end path: what the host <-> device link of this box carries when N processes, one per
GPU, move exactly the bytes a bench step moves -- 29.9 MB of frames host -> device and 17.2 MB of keypoints +
descriptors device -> host per 64-frame KITTI step -- concurrently on two streams from pinned memory, with no kernels
at all.  frames/s ceiling = 64 * steps/s of the slower direction pair.  Run for N = 1, 2, 4, 8 (as many GPUs as the
box has); the e2e figure of bench.py cannot exceed this number on this box whatever the kernels do.

    python tools/probes/pcie_ceiling.py [--seconds 1.5] [--out gpurun_out/pcie_ceiling.json]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

H2D_BYTES = 64 * 376 * 1241          # one bench step of frames
D2H_BYTES = 64 * 4479 * 60 + 64 * 4  # its keypoints (28 B) + descriptors (32 B) + counts


def child(rank, n, seconds, sync_dir, pin_cpus):
    import torch
    if pin_cpus:
        cpus = sorted(os.sched_getaffinity(0))
        per = max(1, len(cpus) // n)
        os.sched_setaffinity(0, set(cpus[rank * per:(rank + 1) * per]) or set(cpus))
    torch.set_num_threads(1)
    torch.cuda.set_device(rank)
    R = 3
    h_in = [torch.empty(H2D_BYTES, dtype=torch.uint8).pin_memory() for _ in range(R)]
    h_out = [torch.empty(D2H_BYTES, dtype=torch.uint8).pin_memory() for _ in range(R)]
    for t in h_in:
        t.fill_(rank + 1)
    d_in = [torch.empty(H2D_BYTES, dtype=torch.uint8, device="cuda") for _ in range(R)]
    d_out = [torch.zeros(D2H_BYTES, dtype=torch.uint8, device="cuda") for _ in range(R)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def run(mode, secs):
        """mode: 'h2d', 'd2h' or 'both'.  Copies are enqueued a few at a time so the queues never run dry."""
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps = 0
        while time.perf_counter() - t0 < secs:
            for i in range(8):
                if mode != "d2h":
                    with torch.cuda.stream(s_in):
                        d_in[i % R].copy_(h_in[i % R], non_blocking=True)
                if mode != "h2d":
                    with torch.cuda.stream(s_out):
                        h_out[i % R].copy_(d_out[i % R], non_blocking=True)
            steps += 8
            s_in.synchronize()
            s_out.synchronize()
        return steps / (time.perf_counter() - t0)

    run("both", 0.2)
    open(os.path.join(sync_dir, f"ready{rank}"), "w").close()
    go = os.path.join(sync_dir, "go")
    while not os.path.exists(go):
        time.sleep(0.001)
    res = {}
    for mode in ("both", "h2d", "d2h"):
        res[mode] = run(mode, seconds)
        # crude re-alignment between modes: everybody runs the same wall-clock length
    print(json.dumps({"rank": rank, "steps_per_s": res}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--out", default=None)
    ap.add_argument("--ns", default="1,2,4,8")
    ap.add_argument("--no-pin", action="store_true")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--n", type=int, default=1)
    ap.add_argument("--sync-dir", default="")
    args = ap.parse_args()
    if args.child:
        child(args.rank, args.n, args.seconds, args.sync_dir, not args.no_pin)
        return
    import torch
    ngpu = torch.cuda.device_count()
    table = []
    for n in [int(x) for x in args.ns.split(",")]:
        if n > ngpu:
            continue
        with tempfile.TemporaryDirectory() as d:
            procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", "--rank", str(r), "--n", str(n),
                                       "--seconds", str(args.seconds), "--sync-dir", d] + (["--no-pin"] if args.no_pin else []),
                                      stdout=subprocess.PIPE, text=True) for r in range(n)]
            t0 = time.time()
            while sum(os.path.exists(os.path.join(d, f"ready{r}")) for r in range(n)) < n:
                if time.time() - t0 > 300 or any(p.poll() not in (None, 0) for p in procs):
                    break
                time.sleep(0.01)
            open(os.path.join(d, "go"), "w").close()
            outs = [p.communicate(timeout=300)[0] for p in procs]
        rows = [json.loads(o.strip().splitlines()[-1]) for o in outs if o.strip()]
        row = {"n_gpus": n, "cpus": len(os.sched_getaffinity(0))}
        for mode in ("both", "h2d", "d2h"):
            sps = [r["steps_per_s"][mode] for r in rows]
            tot = sum(sps)
            row[mode] = {"steps_per_s_total": tot, "steps_per_s_min_rank": min(sps),
                         "frames_per_s_total": 64 * tot,
                         "h2d_GBps_per_gpu": (tot / n) * H2D_BYTES / 1e9 if mode != "d2h" else 0.0,
                         "d2h_GBps_per_gpu": (tot / n) * D2H_BYTES / 1e9 if mode != "h2d" else 0.0}
        table.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({"h2d_bytes_per_step": H2D_BYTES, "d2h_bytes_per_step": D2H_BYTES, "seconds_per_mode": args.seconds,
                   "table": table}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
