#!/usr/bin/env python3
"""Which encodings of the tcgen05 operands does the device accept for k_match_mma?  Runs the brute-force search on
seeded descriptors with the POPC kernel (ORB_B200_MATCH=popc) and with the tensor-core kernel for each
(ORB_B200_MMA_KIND, ORB_B200_MMA_VARIANT), every combination in its own process (a faulting launch poisons its context),
and prints how many rows agree.  Variant 0 / kind i8 is what the library ships (the warp-specialised kernel); 10 = the
first form of the kernel; 1, 2, 11, 12 = deliberately different descriptor encodings.

    python tools/probes/mma_probe.py [--kinds i8,f8] [--variants 0,1,2]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    sys.path.insert(0, ROOT)
    import numpy as np
    from orb_slam_system_b200 import ORBmatcher
    rng = np.random.default_rng(5)
    out = {}
    m = ORBmatcher(0.6, True)
    for name, nq, nt, live in (("small", 200, 300, 256), ("ragged", 333, 1111, 256), ("fork182", 500, 2000, 182), ("one", 5, 1, 256)):
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        if live < 256:
            mask = np.packbits((np.arange(256) < live).astype(np.uint8), bitorder="little")
            q &= mask
            t &= mask
        t[nt // 2:] = t[:nt - nt // 2]  # duplicates: ties between equal rows, first index must win
        os.environ["ORB_B200_MATCH"] = "popc"
        want = m.match_all(q, t)
        os.environ["ORB_B200_MATCH"] = "mma"
        got = m.match_all(q, t)
        out[name] = {k: int((a == b).sum()) for k, a, b in zip(("idx", "best", "second"), got, want)}
        out[name]["rows"] = nq
        if out[name]["best"] != nq:
            bad = np.nonzero(got[1] != want[1])[0][:4]
            out[name]["sample"] = [[int(got[1][i]), int(want[1][i]), int(got[0][i]), int(want[0][i])] for i in bad]
    print("RESULT " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kinds", default="i8,f8")
    ap.add_argument("--variants", default="0,10")
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        return child()
    for kind in a.kinds.split(","):
        for var in a.variants.split(","):
            env = dict(os.environ, ORB_B200_MMA_KIND=kind, ORB_B200_MMA_VARIANT=var)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=180)
                lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                print(f"kind={kind} variant={var} rc={r.returncode} " + (lines[-1][7:] if lines else "no result: " + (r.stderr.strip().splitlines() or ["?"])[-1]), flush=True)
            except subprocess.TimeoutExpired:
                print(f"kind={kind} variant={var} TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
