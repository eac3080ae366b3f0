"""One rank of the NCCL cross-shard test (launched by tests/test_gpu_multi.py under torch.distributed.run, one rank per
GPU): every rank extracts its own shard of frames on its GPU, the keyframe descriptor blocks are all-gathered over NCCL
(orb_slam_system_b200.sharding, BASELINE config 5 / SURVEY 8e), and every local frame is brute-force matched against the
same-index keyframe of every other rank.  Checked against the CPU oracle, which extracts the OTHER ranks' frames itself:
a wrong byte anywhere in the exchange or a wrong match shows up."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import oracle
    from orb_slam_system_b200 import ORBextractor, ORBmatcher
    from orb_slam_system_b200.sharding import all_gather_descriptors, shard_range

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows, cols, nf, F = 240, 320, 500, 3  # F frames per rank
    total = world * F
    b, e = shard_range(total, rank, world)
    assert e - b == F
    frame = lambda g: oracle.synth_frame(rows, cols, frame=700 + g // 2, right=g & 1)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=F, device=local)
    kmax = ex.keypoint_bound(rows, cols)
    pitch = (cols + 63) // 64 * 64
    d_in = torch.zeros((F, rows, pitch), dtype=torch.uint8, device="cuda")
    d_in[:, :, :cols] = torch.from_numpy(np.stack([frame(g) for g in range(b, e)])).cuda()
    d_k = torch.zeros((F, kmax, 28), dtype=torch.uint8, device="cuda")
    block = torch.zeros((F, kmax, 32), dtype=torch.uint8, device="cuda")
    cnts = torch.zeros((F,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ex.extract_batch_device(d_in[:, :, :cols], d_k, block, cnts, kmax)
    ex.sync()
    all_d, all_c = all_gather_descriptors(block, cnts)
    torch.cuda.synchronize()
    # the gathered block against the oracle's own extraction of every frame of every rank
    want = [oracle.extract(frame(g), nfeatures=nf, cap=16 * nf)[1] for g in range(total)]
    for g in range(total):
        n = int(all_c[g].item())
        assert n == len(want[g]), (rank, g, n, len(want[g]))
        assert np.array_equal(all_d[g, :n].cpu().numpy(), want[g]), (rank, g)
    m = ORBmatcher(0.6, True, device=local)
    for d in range(1, world):
        o = (rank + d) % world
        bi = torch.empty((F, kmax), dtype=torch.int32, device="cuda")
        bd, sd = torch.empty_like(bi), torch.empty_like(bi)
        m.match_all_batch_device(block, cnts, all_d[o * F:(o + 1) * F], all_c[o * F:(o + 1) * F], bi, bd, sd)
        m.sync()
        for f in range(F):
            q, t = want[b + f], want[o * F + f]
            oi, od, os_ = oracle.match_all(q, t)
            n = len(q)
            assert np.array_equal(bi[f, :n].cpu().numpy(), oi), (rank, o, f)
            assert np.array_equal(bd[f, :n].cpu().numpy(), od), (rank, o, f)
            assert np.array_equal(sd[f, :n].cpu().numpy(), os_), (rank, o, f)
    dist.barrier()
    ex.close()
    m.close()
    dist.destroy_process_group()
    print(f"rank {rank}: cross-shard gather and matches identical to the oracle", flush=True)


if __name__ == "__main__":
    main()
