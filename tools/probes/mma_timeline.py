#!/usr/bin/env python3
"""Hand-over timeline of k_match_mma3's three roles on SM 0 (a -DORB_B200_MMA_KNOCKOUT build, ORB_B200_MMA_DEBUG with bit 16):
clock() at producer {data ready, stage free, stage full}, MMA warp {stage full seen, accumulator free seen, MMAs issued} and
epilogue {accumulator full seen, accumulator released}, printed per tile relative to the first event shown."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from orb_slam_system_b200 import ORBmatcher
    dbg = sys.argv[1] if len(sys.argv) > 1 else "16"
    NP, N = 256, 2000
    m = ORBmatcher(0.6, True)
    g = torch.Generator(device="cuda").manual_seed(3)
    full = torch.randint(0, 256, (NP + 1, N, 32), dtype=torch.uint8, device="cuda", generator=g)
    mask = torch.zeros(32, dtype=torch.uint8, device="cuda")
    mask[:22] = 255
    mask[22] = 0x3F
    data = full & mask
    q, t = data[:NP].contiguous(), data[1:].contiguous()
    nq = torch.full((NP,), N, dtype=torch.int32, device="cuda")
    out = [torch.zeros((NP, N), dtype=torch.int32, device="cuda") for _ in range(3)]
    os.environ["ORB_B200_MATCH"] = "mma"
    os.environ["ORB_B200_MMA_DEBUG"] = "0"
    for _ in range(3):
        m.match_all_batch_device(q, nq, t, nq, *out)
    m.sync()
    os.environ["ORB_B200_MMA_DEBUG"] = dbg
    m.match_all_batch_device(q, nq, t, nq, *out)
    m.sync()
    tr = out[2].flatten()[: 16 * 256].cpu().numpy().reshape(16, 256).astype("int64") & 0xFFFFFFFF
    names = {0: "P.data", 1: "P.free", 2: "P.full", 4: "M.full", 5: "M.accfree", 6: "M.issued", 8: "E.accfull", 9: "E.release"}
    lo, hi = 40, 60
    base = min(int(tr[k, lo]) for k in names)
    print(f"debug={dbg}; cycles relative to tile {lo}'s first event")
    print("tile " + " ".join(f"{v:>10s}" for v in names.values()))
    for n in range(lo, hi):
        print(f"{n:4d} " + " ".join(f"{(int(tr[k, n]) - base) & 0xFFFFFFFF:10d}" for k in names))
    per = (int(tr[6, hi]) - int(tr[6, lo])) / (hi - lo)
    print(f"cycles per tile (M.issued): {per:.0f}")


if __name__ == "__main__":
    main()
