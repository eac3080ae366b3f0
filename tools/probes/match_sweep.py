#!/usr/bin/env python3
"""Sweep of the tensor-core matcher's tuning knobs in one process: every combination of the --env KEY=v1,v2,... lists is timed
on config 4's shape (NP pairs of N x N descriptors, device resident) and checked against the popcount kernel.

    python tools/probes/match_sweep.py --env ORB_B200_MMA_FM=0,16,31 --env ORB_B200_MMA_PARK=0,2000
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--env", action="append", default=[])
    a = ap.parse_args()
    import torch
    from orb_slam_system_b200 import ORBmatcher
    NP, N = a.pairs, a.n
    m = ORBmatcher(0.6, True)
    g = torch.Generator(device="cuda").manual_seed(3)
    full = torch.randint(0, 256, (NP + 1, N, 32), dtype=torch.uint8, device="cuda", generator=g)
    mask = torch.zeros(32, dtype=torch.uint8, device="cuda")
    mask[:22] = 255
    mask[22] = 0x3F
    nq = torch.full((NP,), N, dtype=torch.int32, device="cuda")
    out = [torch.empty((NP, N), dtype=torch.int32, device="cuda") for _ in range(3)]
    ref = [torch.empty((NP, N), dtype=torch.int32, device="cuda") for _ in range(3)]
    st = torch.cuda.ExternalStream(m.stream)
    keys = [e.split("=")[0] for e in a.env]
    vals = [e.split("=")[1].split(",") for e in a.env]
    for live, data in (("182", full & mask), ("256", full)):
        q, t = data[:NP].contiguous(), data[1:].contiguous()
        os.environ["ORB_B200_MATCH"] = "popc"
        m.match_all_batch_device(q, nq, t, nq, *ref)
        m.sync()
        os.environ["ORB_B200_MATCH"] = "mma"
        for combo in itertools.product(*vals):
            for k, v in zip(keys, combo):
                os.environ[k] = v
            for x in out:
                x.fill_(-7)
            for _ in range(3):
                m.match_all_batch_device(q, nq, t, nq, *out)
            m.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(a.reps):
                m.match_all_batch_device(q, nq, t, nq, *out)
            e1.record(st)
            m.sync()
            ms = e0.elapsed_time(e1) / a.reps
            same = all(bool((x == y).all().item()) for x, y in zip(out, ref))
            print(json.dumps({"live": live, **dict(zip(keys, combo)), "ms": round(ms, 4), "Tpairs_per_s": round(NP * N * N / ms / 1e9, 3), "identical": same}), flush=True)


if __name__ == "__main__":
    main()
