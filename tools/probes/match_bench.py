#!/usr/bin/env python3
"""Brute-force Hamming rate (BASELINE config 4: 2000 x 2000 descriptors per keyframe pair, 256 pairs, device resident)
of the tensor-core kernel and of the POPC kernel, on descriptors with 182 live bits (this fork) and with 256.

    python tools/probes/match_bench.py [--pairs 256] [--n 2000] [--reps 10] [--only mma|popc]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from orb_slam_system_b200 import ORBmatcher  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    NP, N = a.pairs, a.n
    m = ORBmatcher(0.6, True)
    g = torch.Generator(device="cuda").manual_seed(3)
    full = torch.randint(0, 256, (NP + 1, N, 32), dtype=torch.uint8, device="cuda", generator=g)
    mask = torch.zeros(32, dtype=torch.uint8, device="cuda")
    mask[:22] = 255
    mask[22] = 0x3F  # bits 176..181
    nq = torch.full((NP,), N, dtype=torch.int32, device="cuda")
    out = [torch.empty((NP, N), dtype=torch.int32, device="cuda") for _ in range(3)]
    ref = [torch.empty((NP, N), dtype=torch.int32, device="cuda") for _ in range(3)]
    st = torch.cuda.ExternalStream(m.stream)
    res = {}
    for live, data in (("182", full & mask), ("256", full)):
        q, t = data[:NP].contiguous(), data[1:].contiguous()
        # mma1: the first form of the tensor-core kernel (ORB_B200_MMA_VARIANT=10); mma16 / mma8: warp-specialised, one CTA per 256 queries, 16 / 8 epilogue warps (20 / 30); mma: the persistent kernel that ships
        for impl in ("popc", "mma1", "mma16", "mma8", "mma"):
            if a.only and impl != a.only:
                continue
            os.environ["ORB_B200_MATCH"] = "popc" if impl == "popc" else "mma"
            os.environ["ORB_B200_MMA_VARIANT"] = {"mma1": "10", "mma16": "20", "mma8": "30"}.get(impl, "0")
            o = ref if impl == "popc" else out
            for _ in range(2):
                m.match_all_batch_device(q, nq, t, nq, *o)
            m.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(a.reps):
                m.match_all_batch_device(q, nq, t, nq, *o)
            e1.record(st)
            m.sync()
            ms = e0.elapsed_time(e1) / a.reps
            res[f"{impl}_{live}"] = {"ms": ms, "Gpairs_per_s": NP * N * N / ms / 1e6}
            if impl != "popc" and not a.only:
                res[f"{impl}_{live}"]["identical"] = all(bool((x == y).all().item()) for x, y in zip(out, ref))
                for x in out:
                    x.fill_(-7)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
