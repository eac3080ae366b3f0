"""Host-side mirror of the Hamming searches of the reference's ``ORB_SLAM2::ORBmatcher``
(include/ORBmatcher.h:17-83, src/ORBmatcher.cc) and of ``Frame::ComputeStereoMatches``
(src/Frame.cc:446-529) over the C ABI.  All distances are computed on the GPU; the accept
rules stay on the host exactly as worded in the reference.
"""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, check, lib, ptr

INT_MAX = 2147483647


def _desc(a):
    a = np.ascontiguousarray(a, np.uint8)
    return a.reshape(-1, 32)


class ORBmatcher:
    TH_HIGH = 100      # src/ORBmatcher.cc:13
    TH_LOW = 50        # src/ORBmatcher.cc:14
    HISTO_LENGTH = 30  # src/ORBmatcher.cc:15

    _shared = {}

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio = np.float32(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self.device = device
        self._h = C.c_void_p()
        check(lib().orb_matcher_create(device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().orb_matcher_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the unit: ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:896-908) ----
    @classmethod
    def DescriptorDistance(cls, a, b, device=0):
        m = cls._shared.get(device)
        if m is None:
            m = cls._shared[device] = ORBmatcher(device=device)
        _, bd, _ = m.match_all(_desc(a)[:1], _desc(b)[:1])
        return int(bd[0])

    # ---- raw scans ----
    def match_all(self, q, t):
        """Brute-force scan of every query over all train rows in index order.
        Returns (best_idx, best_dist, second_dist) int32 arrays."""
        q, t = _desc(q), _desc(t)
        bi = np.full(len(q), -1, np.int32)
        bd = np.full(len(q), INT_MAX, np.int32)
        sd = np.full(len(q), INT_MAX, np.int32)
        check(lib().orb_match_all(self._h, ptr(q), len(q), ptr(t), len(t), ptr(bi), ptr(bd), ptr(sd)))
        return bi, bd, sd

    def match_all_batch(self, q, nq, t, nt):
        """q uint8 [P, Qmax, 32], t uint8 [P, Tmax, 32], nq/nt int32 [P] -> three int32 [P, Qmax]."""
        q = np.ascontiguousarray(q, np.uint8)
        t = np.ascontiguousarray(t, np.uint8)
        nq = np.ascontiguousarray(nq, np.int32)
        nt = np.ascontiguousarray(nt, np.int32)
        P, Q = q.shape[0], q.shape[1]
        bi = np.full((P, Q), -1, np.int32)
        bd = np.full((P, Q), INT_MAX, np.int32)
        sd = np.full((P, Q), INT_MAX, np.int32)
        check(lib().orb_match_all_batch(self._h, P, ptr(q), ptr(nq), q.strides[0], ptr(t), ptr(nt), t.strides[0],
                                        ptr(bi), ptr(bd), ptr(sd), Q, 0))
        return bi, bd, sd

    def match_all_batch_device(self, q, nq, t, nt, best_idx, best_dist, second_dist):
        """torch CUDA tensors: q [P,Q,32] u8, t [P,T,32] u8, nq/nt int32 [P], outputs int32 [P,Q]."""
        P, Q = q.shape[0], q.shape[1]
        check(lib().orb_match_all_batch(self._h, P, ptr(q), ptr(nq), q.stride(0), ptr(t), ptr(nt), t.stride(0),
                                        ptr(best_idx), ptr(best_dist), ptr(second_dist), Q, 1))

    def sync(self):
        check(lib().orb_matcher_sync(self._h))

    @property
    def stream(self):
        return lib().orb_matcher_stream(self._h)

    def match_csr(self, q, t, offsets, cand, tie_last=False, max_dist=None):
        """Windowed scan: query i over train rows cand[offsets[i]:offsets[i+1]] in that order."""
        q, t = _desc(q), _desc(t)
        offsets = np.ascontiguousarray(offsets, np.int32)
        cand = np.ascontiguousarray(cand, np.int32)
        assert len(offsets) == len(q) + 1
        bi = np.full(len(q), -1, np.int32)
        bd = np.full(len(q), INT_MAX, np.int32)
        sd = np.full(len(q), INT_MAX, np.int32)
        md = self.TH_LOW if max_dist is None else int(max_dist)
        check(lib().orb_match_csr(self._h, ptr(q), len(q), ptr(t), len(t), ptr(offsets), ptr(cand),
                                  1 if tie_last else 0, md, ptr(bi), ptr(bd), ptr(sd)))
        return bi, bd, sd

    # ---- BASELINE config 4: brute force + the SearchByProjection accept rule (:58) ----
    def BruteForceRatio(self, q, t):
        """match_all + `best <= TH_HIGH && best <= nnratio * second` (src/ORBmatcher.cc:58).
        Returns int32 matches[q] = train index or -1."""
        bi, bd, sd = self.match_all(q, t)
        ok = (bd <= self.TH_HIGH) & (bd.astype(np.float32) <= self.mfNNratio * sd.astype(np.float32))
        return np.where(ok, bi, -1).astype(np.int32)

    # ---- Frame::ComputeStereoMatches, Hamming part (src/Frame.cc:446-529) ----
    def stereo_match(self, kps_left, desc_left, kps_right, desc_right, scale_factors, rows, bf, fx):
        kl = np.ascontiguousarray(kps_left, KP_DTYPE)
        kr = np.ascontiguousarray(kps_right, KP_DTYPE)
        dl, dr = _desc(desc_left), _desc(desc_right)
        sc = np.ascontiguousarray(scale_factors, np.float32)
        br = np.full(len(kl), -1, np.int32)
        bd = np.full(len(kl), self.TH_HIGH, np.int32)
        check(lib().orb_stereo_match(self._h, ptr(kl), ptr(dl), len(kl), ptr(kr), ptr(dr), len(kr), ptr(sc), len(sc),
                                     rows, C.c_float(bf), C.c_float(fx), ptr(br), ptr(bd)))
        return br, bd
