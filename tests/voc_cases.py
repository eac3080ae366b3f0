"""Synthetic DBoW2 vocabularies for the transform tests: a k-ary tree of depth L whose node descriptors are bit-flipped
copies of their parent's (so descents are meaningful), idf-like leaf weights with some stopped words (weight 0), a few
leaves above the bottom level, written in the fork's text format."""
import numpy as np


class VocArrays:
    pass


def make_vocabulary(rng, k=10, L=4, weighting=0, scoring=0, early_leaf=0.0, stopped=0.05, dup_children=True):
    parent, desc, weight, leaf = [0], [np.zeros(32, np.uint8)], [0.0], [False]
    root_bits = rng.integers(0, 2, 256, dtype=np.uint8)
    root_bits[182:] = 0
    frontier = [(0, root_bits, 0)]
    bits_of = {0: root_bits}
    while frontier:
        nxt = []
        for pid, pbits, lvl in frontier:
            for c in range(k):
                b = pbits.copy()
                nflip = max(2, 60 >> lvl)
                b[rng.choice(182, nflip, replace=False)] ^= 1
                if dup_children and c == 1 and rng.random() < 0.2:
                    b = bits_of[len(parent) - 1].copy()  # twin of the previous child: equal distances -> first wins
                nid = len(parent)
                parent.append(pid)
                desc.append(np.packbits(b, bitorder="little"))
                bits_of[nid] = b
                is_leaf = lvl + 1 == L or (lvl + 1 >= 2 and rng.random() < early_leaf)
                leaf.append(is_leaf)
                weight.append(0.0 if (is_leaf and rng.random() < stopped) else (float(rng.uniform(0.5, 9.0)) if is_leaf else 0.0))
                if not is_leaf:
                    nxt.append((nid, b, lvl + 1))
        frontier = nxt
    n = len(parent)
    kids = [[] for _ in range(n)]
    for i in range(1, n):
        kids[parent[i]].append(i)
    v = VocArrays()
    v.k, v.L, v.weighting, v.scoring = k, L, weighting, scoring
    v.child_off = np.zeros(n + 1, np.int32)
    v.child_off[1:] = np.cumsum([len(c) for c in kids])
    v.children = np.array([c for cs in kids for c in cs], np.int32)
    v.node_desc = np.stack(desc).astype(np.uint8)
    v.node_weight = np.array(weight, np.float64)
    v.node_word = np.zeros(n, np.int32)
    w = 0
    for i in range(1, n):
        if leaf[i]:
            v.node_word[i] = w
            w += 1
    v.parent = np.array(parent)
    v.leaf = np.array(leaf)
    return v


def write_text(v, path):
    """The fork's text vocabulary format (TemplatedVocabulary.h:1338-1450); no trailing newline."""
    lines = [f"{v.k} {v.L}  {v.scoring} {v.weighting}"]
    for i in range(1, len(v.parent)):
        lines.append(f"{v.parent[i]} {1 if v.leaf[i] else 0} " + " ".join(str(int(b)) for b in v.node_desc[i]) + f" {float(v.node_weight[i])!r}")
    with open(path, "w") as f:
        f.write("\n".join(lines))


def features_near_leaves(rng, v, n):
    """Descriptors that are noisy copies of random node descriptors (plus some unrelated ones)."""
    src = rng.integers(1, len(v.node_desc), n)
    bits = np.unpackbits(v.node_desc[src], axis=1, bitorder="little")
    for i in range(n):
        kf = int(rng.integers(0, 25))
        if kf:
            bits[i, rng.choice(182, kf, replace=False)] ^= 1
    fresh = rng.random(n) < 0.1
    bits[fresh] = rng.integers(0, 2, (int(fresh.sum()), 256), dtype=np.uint8)
    bits[:, 182:] = 0
    return np.packbits(bits, axis=1, bitorder="little")
