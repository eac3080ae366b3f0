"""Deterministic synthetic frames: counter-based splitmix64 block texture.

Library-free definition; the CPU checker restates the same generator and the
benchmarks; tests assert the two agree byte for byte.  variant 0: blocks (5,11,23,47);
variant 1: blocks (4,8,16,32) with the left half at quarter contrast (exercises the
minThFAST retry); right=1: the right image of a stereo pair, i.e. the left texture sampled
at x + d with a per-block disparity d in 0..40.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(k):
    with np.errstate(over="ignore"):
        z = (k + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def synth_frame(rows, cols, seed=7, frame=0, variant=0, right=0):
    blocks = (4, 8, 16, 32) if variant == 1 else (5, 11, 23, 47)
    weights = (0.4, 0.3, 0.2, 0.1)
    base = (np.uint64(seed) << np.uint64(40)) ^ (np.uint64(frame) << np.uint64(24))
    ys = np.arange(rows, dtype=np.uint64)[:, None]
    xs = np.arange(cols, dtype=np.uint64)[None, :]
    if right:
        dk = base ^ (np.uint64(7) << np.uint64(60)) ^ ((ys // np.uint64(47)) << np.uint64(12)) ^ (xs // np.uint64(94))
        sx = xs + (_splitmix64(dk) & np.uint64(0xFF)) % np.uint64(41)
    else:
        sx = np.broadcast_to(xs, (rows, cols))
    acc = np.zeros((rows, cols), np.float64)
    for o, (b, w) in enumerate(zip(blocks, weights)):
        k = base ^ (np.uint64(o) << np.uint64(60)) ^ ((ys // np.uint64(b)) << np.uint64(12)) ^ (sx // np.uint64(b))
        acc = acc + w * (_splitmix64(k) & np.uint64(0xFF)).astype(np.float64)
    if variant == 1:
        half = cols // 2
        acc[:, :half] = 128.0 + (acc[:, :half] - 128.0) * 0.25
    return np.clip(np.floor(acc + 0.5), 0, 255).astype(np.uint8)
