"""Host-side mirror of ``ORB_SLAM2::ORBVocabulary`` = ``DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>``
(reference include/ORBVocabulary.h, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) for the one call the hot path
makes: ``transform(features, BowVector, FeatureVector, levelsup)`` from Frame::ComputeBoW (src/Frame.cc:375-382)
and KeyFrame::ComputeBoW (src/KeyFrame.cc:39-47).  The tree lives on the device; every descriptor's descent runs
there (liborb_b200.so, k_voc_descent)."""
import ctypes as C

import numpy as np

from ._lib import ORB_ERR_CAPACITY, check, lib, ptr

TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3                                         # BowVector.h:36-42
L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = range(6)       # BowVector.h:45-53


class ORBVocabulary:
    def __init__(self, child_off, children, node_desc, node_weight, node_word, k, L, weighting=TF_IDF, scoring=L1_NORM, device=0):
        self.child_off = np.ascontiguousarray(child_off, np.int32)
        self.children = np.ascontiguousarray(children, np.int32)
        self.node_desc = np.ascontiguousarray(node_desc, np.uint8).reshape(-1, 32)
        self.node_weight = np.ascontiguousarray(node_weight, np.float64)
        self.node_word = np.ascontiguousarray(node_word, np.int32)
        self.k, self.L, self.weighting, self.scoring = int(k), int(L), int(weighting), int(scoring)
        n = len(self.child_off) - 1
        assert len(self.node_desc) == n and len(self.node_weight) == n and len(self.node_word) == n
        self._h = C.c_void_p()
        check(lib().orb_vocabulary_create(device, n, ptr(self.child_off), ptr(self.children), ptr(self.node_desc), ptr(self.node_weight),
                                          ptr(self.node_word), self.L, self.weighting, self.scoring, C.byref(self._h)))

    @classmethod
    def loadFromTextFile(cls, path, device=0):
        """The fork's text format (TemplatedVocabulary.h:1338-1422): first line `k L scoring weighting`, then one line
        per node `parent isLeaf d0 .. d31 weight`; node ids = line numbers, word ids = leaf order."""
        with open(path) as f:
            lines = f.read().split("\n")
        k, L, n1, n2 = (int(t) for t in lines[0].split())
        if k < 0 or k > 20 or L < 1 or L > 10 or n1 < 0 or n1 > 5 or n2 < 0 or n2 > 3:
            raise ValueError("Vocabulary loading failure: This is not a correct text file!")
        rows = [ln.split() for ln in lines[1:] if ln.strip()]
        n = len(rows) + 1
        parent = np.zeros(n, np.int64)
        desc = np.zeros((n, 32), np.uint8)
        weight = np.zeros(n, np.float64)
        word = np.zeros(n, np.int32)
        kids = [[] for _ in range(n)]
        nwords = 0
        for i, t in enumerate(rows, start=1):
            parent[i] = int(t[0])
            kids[parent[i]].append(i)
            desc[i] = [int(x) for x in t[2:34]]
            weight[i] = float(t[34])
            if int(t[1]) > 0:
                word[i] = nwords
                nwords += 1
        off = np.zeros(n + 1, np.int32)
        off[1:] = np.cumsum([len(c) for c in kids])
        flat = np.array([c for cs in kids for c in cs], np.int32)
        return cls(off, flat, desc, weight, word, k, L, weighting=n2, scoring=n1, device=device)

    def saveToTextFile(self, path):
        """TemplatedVocabulary::saveToTextFile (:1426-1450), without a trailing newline (the fork's loader would
        turn one into a spurious extra node)."""
        n = len(self.child_off) - 1
        parent = np.zeros(n, np.int64)
        for i in range(n):
            parent[self.children[self.child_off[i]:self.child_off[i + 1]]] = i
        out = [f"{self.k} {self.L}  {self.scoring} {self.weighting}"]
        for i in range(1, n):
            leaf = 1 if self.child_off[i] == self.child_off[i + 1] else 0
            out.append(f"{parent[i]} {leaf} " + " ".join(str(int(b)) for b in self.node_desc[i]) + f" {float(self.node_weight[i])!r}")
        with open(path, "w") as f:
            f.write("\n".join(out))

    def size(self):
        """Number of words."""
        return int((np.diff(self.child_off)[1:] == 0).sum()) if len(self.child_off) > 2 else 0

    def transform(self, desc, levelsup=4):
        """Returns dict(words, nodes, bow_ids, bow_values, fv) -- BowVector as parallel arrays in ascending word id,
        FeatureVector as {node id: [feature indices]}."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        words, nodes = np.zeros(n, np.int32), np.zeros(n, np.int32)
        bi, bv = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.float64)
        fn, fo, fi = np.zeros(n + 1, np.int32), np.zeros(n + 2, np.int32), np.zeros(n + 1, np.int32)
        nb, nf = C.c_int(0), C.c_int(0)
        check(lib().orb_vocabulary_transform(self._h, ptr(d), n, int(levelsup), ptr(words), ptr(nodes), ptr(bi), ptr(bv), n + 1,
                                             C.byref(nb), ptr(fn), ptr(fo), ptr(fi), n + 1, C.byref(nf)))
        fv = {int(fn[k]): fi[fo[k]:fo[k + 1]].tolist() for k in range(nf.value)}
        return dict(words=words, nodes=nodes, bow_ids=bi[: nb.value].copy(), bow_values=bv[: nb.value].copy(), fv=fv)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().orb_vocabulary_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
