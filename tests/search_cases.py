"""Shared builders for the search-method tests: synthetic Frame / KeyFrame views, projected queries and
DBoW2-style feature vectors.  Deterministic (seeded) so the CPU and GPU suites see identical inputs."""
import numpy as np

from orb_slam_system_b200 import FeatureVector, FrameView
from orb_slam_system_b200._lib import KP_DTYPE

SCALE = np.array([1, 1, 1.2, 1.44, 1.728, 2.0736, 2.48832, 2.985984], np.float32)  # the fork's table


def rand_desc(rng, n, live_bits=182):
    bits = rng.integers(0, 2, size=(n, 256), dtype=np.uint8)
    bits[:, live_bits:] = 0
    return np.packbits(bits, axis=1, bitorder="little")


def noisy_copy(rng, desc, max_flips):
    bits = np.unpackbits(desc, axis=1, bitorder="little")
    for i in range(len(bits)):
        k = int(rng.integers(0, max_flips + 1))
        if k:
            bits[i, rng.choice(182, size=k, replace=False)] ^= 1
    return np.packbits(bits, axis=1, bitorder="little")


def make_frame(rng, n, cols=752, rows=480, stereo=False, clustered=False, min_x=0.0, min_y=0.0):
    """n keypoints; some outside the image bounds (undistorted points can leave it), integer and fractional
    coordinates, 8 octaves, angles in [0, 360)."""
    k = np.zeros(n, KP_DTYPE)
    if clustered:  # many keypoints per grid cell
        cx, cy = rng.uniform(50, cols - 50, 12), rng.uniform(50, rows - 50, 12)
        w = rng.integers(0, 12, n)
        k["x"] = cx[w] + rng.normal(0, 9, n)
        k["y"] = cy[w] + rng.normal(0, 9, n)
    else:
        k["x"] = rng.uniform(min_x - 6, min_x + cols + 6, n)
        k["y"] = rng.uniform(min_y - 6, min_y + rows + 6, n)
    whole = rng.random(n) < 0.4  # level-0/1 keypoints have integer coordinates
    k["x"][whole] = np.round(k["x"][whole])
    k["y"][whole] = np.round(k["y"][whole])
    k["octave"] = rng.integers(0, 8, n)
    k["octave"][whole] = rng.integers(0, 2, whole.sum())
    k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    k["angle"][k["angle"] >= 360] = 0
    k["size"] = 31 * SCALE[k["octave"]]
    k["response"] = rng.integers(20, 120, n)
    k["class_id"] = -1
    desc = rand_desc(rng, n)
    ur = None
    if stereo:
        ur = np.where(rng.random(n) < 0.6, k["x"] - rng.uniform(0.5, 40, n), -1).astype(np.float32)
    return FrameView(k, desc, min_x, min_x + cols, min_y, min_y + rows, scale_factors=SCALE, u_right=ur)


def projected_queries(rng, F, nq, jitter=3.0, max_flips=45, outside=0.05):
    """Map points that project near features of F: descriptor = noisy copy of that feature's, position =
    the feature's plus jitter; a few far outside the image."""
    src = rng.integers(0, F.N, nq)
    q = noisy_copy(rng, F.desc[src], max_flips)
    u = (F.keys_un["x"][src] + rng.normal(0, jitter, nq)).astype(np.float32)
    v = (F.keys_un["y"][src] + rng.normal(0, jitter, nq)).astype(np.float32)
    far = rng.random(nq) < outside
    u[far] += rng.choice([-3000, 3000], far.sum())
    level = np.clip(F.keys_un["octave"][src] + rng.integers(-1, 2, nq), 0, 7).astype(np.int32)
    angle = np.mod(F.keys_un["angle"][src] + rng.normal(0, 4, nq), 360).astype(np.float32)
    return src, q, u, v, level, angle


def feature_vector(rng, words, n_nodes):
    """Feature vector from per-feature vocabulary words: node = word // 10, indices in ascending order like
    DBoW2's addFeature calls (TemplatedVocabulary::transform visits features in index order)."""
    fv = {}
    for i, w in enumerate(words):
        if w < 0:
            continue  # some features get no node
        fv.setdefault(int(w) % n_nodes, []).append(i)
    return fv


def bow_pair(rng, n1=900, n2=1000, n_nodes=40):
    """Two keyframes that share structure: KF2 features are noisy copies of KF1's (plus unrelated ones),
    binned into the same node when the copy is close."""
    F1 = make_frame(rng, n1)
    src = rng.integers(0, n1, n2)
    d2 = noisy_copy(rng, F1.desc[src], 60)
    fresh = rng.random(n2) < 0.25
    d2[fresh] = rand_desc(rng, int(fresh.sum()))
    k2 = F1.keys_un[src].copy()
    k2["x"] += rng.normal(0, 2, n2).astype(np.float32)
    k2["y"] += rng.normal(0, 2, n2).astype(np.float32)
    k2["angle"] = np.mod(k2["angle"] + rng.normal(0, 5, n2), 360).astype(np.float32)
    k2["angle"][k2["angle"] >= 360] = 0
    F2 = FrameView(k2, d2, 0, 752, 0, 480, scale_factors=SCALE)
    w1 = rng.integers(0, n_nodes, n1)
    w2 = np.where(rng.random(n2) < 0.8, w1[src], rng.integers(0, n_nodes, n2))
    w1[rng.random(n1) < 0.03] = -1
    # exact duplicates inside a bucket so that best == second and tie rules are exercised
    for j in range(0, n2 - 1, 50):
        d2[j + 1] = d2[j]
        w2[j + 1] = w2[j]
    F2.desc[:] = d2
    fv1 = feature_vector(rng, w1, n_nodes)
    fv2 = feature_vector(rng, w2, n_nodes)
    fv2.pop(3, None)  # nodes present on one side only -> the lower_bound jumps
    fv1.pop(7, None)
    has1 = (rng.random(n1) < 0.7).astype(np.uint8)
    has2 = (rng.random(n2) < 0.7).astype(np.uint8)
    return F1, F2, fv1, fv2, has1, has2


def numpy_features_in_area(F, x, y, r, min_level=-1, max_level=-1):
    """Independent, direct restatement of Frame::AssignFeaturesToGrid + GetFeaturesInArea in numpy float32
    (one query), used to pin the oracle's own restatement."""
    f = np.float32
    k = F.keys_un
    # C round() is half-away-from-zero (np.round is half-to-even)
    vx = ((k["x"] - F.mnMinX) * F.mfGridElementWidthInv).astype(np.float64)
    vy = ((k["y"] - F.mnMinY) * F.mfGridElementHeightInv).astype(np.float64)
    px = np.sign(vx) * np.floor(np.abs(vx) + 0.5)
    py = np.sign(vy) * np.floor(np.abs(vy) + 0.5)
    x, y, r = f(x), f(y), f(r)
    x0 = max(0, int(np.floor(f(f(f(x - F.mnMinX) - r) * F.mfGridElementWidthInv))))
    x1 = min(63, int(np.ceil(f(f(f(x - F.mnMinX) + r) * F.mfGridElementWidthInv))))
    y0 = max(0, int(np.floor(f(f(f(y - F.mnMinY) - r) * F.mfGridElementHeightInv))))
    y1 = min(47, int(np.ceil(f(f(f(y - F.mnMinY) + r) * F.mfGridElementHeightInv))))
    if x0 >= 64 or x1 < 0 or y0 >= 48 or y1 < 0:
        return []
    check = (min_level > 0) or (max_level >= 0)
    out = []
    for ix in range(x0, x1 + 1):
        for iy in range(y0, y1 + 1):
            for i in np.nonzero((px == ix) & (py == iy))[0]:
                if check:
                    if k["octave"][i] < min_level:
                        continue
                    if max_level >= 0 and k["octave"][i] > max_level:
                        continue
                if abs(f(k["x"][i] - x)) < r and abs(f(k["y"][i] - y)) < r:
                    out.append(int(i))
    return out
