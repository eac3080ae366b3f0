"""Host-side mirror of the Hamming searches of the reference's ``ORB_SLAM2::ORBmatcher``
(include/ORBmatcher.h:17-83, src/ORBmatcher.cc) and of ``Frame::ComputeStereoMatches``
(src/Frame.cc:446-529) over the C ABI.  All distances are computed on the GPU; the accept
rules stay on the host exactly as worded in the reference.
"""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, ORB_ERR_CAPACITY, OrbFeatureVector, OrbFrameView, check, lib, ptr

INT_MAX = 2147483647


def _desc(a):
    a = np.ascontiguousarray(a, np.uint8)
    return a.reshape(-1, 32)


FRAME_GRID_ROWS = 48  # include/Frame.h:17
FRAME_GRID_COLS = 64  # include/Frame.h:18


class FrameView:
    """The members of a reference ``Frame`` / ``KeyFrame`` that the searches read: mvKeysUn, mDescriptors,
    the image bounds and grid pitch (src/Frame.cc:74-83: mfGridElementWidthInv = 64 / (mnMaxX - mnMinX), ...),
    mvuRight and mvScaleFactors.  The 64 x 48 grid itself is rebuilt on the device."""

    def __init__(self, keys_un, desc, min_x, max_x, min_y, max_y, scale_factors=None, u_right=None):
        self.keys_un = np.ascontiguousarray(keys_un, KP_DTYPE)
        self.desc = _desc(desc)
        assert len(self.keys_un) == len(self.desc)
        self.mnMinX, self.mnMaxX, self.mnMinY, self.mnMaxY = (np.float32(v) for v in (min_x, max_x, min_y, max_y))
        self.mfGridElementWidthInv = np.float32(FRAME_GRID_COLS) / (self.mnMaxX - self.mnMinX)
        self.mfGridElementHeightInv = np.float32(FRAME_GRID_ROWS) / (self.mnMaxY - self.mnMinY)
        self.mvScaleFactors = None if scale_factors is None else np.ascontiguousarray(scale_factors, np.float32)
        self.mvuRight = None if u_right is None else np.ascontiguousarray(u_right, np.float32)

    @property
    def N(self):
        return len(self.keys_un)

    def c_view(self):
        return OrbFrameView(self.keys_un.ctypes.data, self.desc.ctypes.data, len(self.keys_un), self.mnMinX, self.mnMinY,
                            self.mfGridElementWidthInv, self.mfGridElementHeightInv)


class FeatureVector:
    """DBoW2::FeatureVector (node id -> feature indices) flattened to CSR with ascending node ids."""

    def __init__(self, node_to_indices):
        self.nodes = np.array(sorted(node_to_indices), np.int32)
        self.off = np.zeros(len(self.nodes) + 1, np.int32)
        flat = []
        for k, nd in enumerate(self.nodes):
            flat.extend(node_to_indices[int(nd)])
            self.off[k + 1] = len(flat)
        self.idx = np.array(flat, np.int32)

    def c_struct(self):
        return OrbFeatureVector(self.nodes.ctypes.data, self.off.ctypes.data, self.idx.ctypes.data, len(self.nodes))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, np.int32)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, np.uint8)


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

    # ---- Frame::ComputeStereoMatches whole (src/Frame.cc:446-619) ----
    def ComputeStereoMatches(self, ex_left, ex_right, kps_left, desc_left, kps_right, desc_right, bf, fx, frame_left=0, frame_right=0):
        """Hamming search + SAD sub-pixel refinement + median cut; the pyramids are read on the device from the two
        ORBextractor objects (their last call).  Returns (mvuRight, mvDepth)."""
        kl = np.ascontiguousarray(kps_left, KP_DTYPE)
        kr = np.ascontiguousarray(kps_right, KP_DTYPE)
        dl, dr = _desc(desc_left), _desc(desc_right)
        ur = np.full(len(kl), -1, np.float32)
        dep = np.full(len(kl), -1, np.float32)
        check(lib().orb_compute_stereo_matches(self._h, ex_left._h, frame_left, ex_right._h, frame_right, ptr(kl), ptr(dl), len(kl),
                                               ptr(kr), ptr(dr), len(kr), C.c_float(bf), C.c_float(fx), ptr(ur), ptr(dep)))
        return ur, dep

    def ComputeStereoMatchesBatch(self, ex, kps, desc, counts, cap, bf, fx):
        """Frame::ComputeStereoMatches for every stereo pair of one extractor batch (pair p = frames 2p, 2p + 1 of ex's last
        call).  Host arrays in the batch layout: kps [2P, cap] KP_DTYPE (or bytes [2P, cap, 28]), desc [2P, cap, 32], counts [2P].
        Returns (u_right [P, cap], depth [P, cap], status [P]); rows of pair p are valid up to counts[2p]."""
        counts = np.ascontiguousarray(counts, np.int32)
        P = len(counts) // 2
        ur = np.full((P, cap), -1, np.float32)
        dep = np.full((P, cap), -1, np.float32)
        status = np.zeros(P, np.int32)
        check(lib().orb_compute_stereo_matches_batch(self._h, ex._h, P, ptr(kps), ptr(desc), cap, ptr(counts), C.c_float(bf), C.c_float(fx),
                                                     ptr(ur), ptr(dep), ptr(status), 0))
        return ur, dep, status

    def ComputeStereoMatchesBatchDevice(self, ex, d_kps, d_desc, d_counts, cap, bf, fx, d_u_right, d_depth, d_status=None):
        """The same on the device-resident outputs of ORBextractor.extract_batch_device (torch CUDA tensors), asynchronous on
        the matcher's stream."""
        P = int(d_counts.shape[0]) // 2
        check(lib().orb_compute_stereo_matches_batch(self._h, ex._h, P, ptr(d_kps), ptr(d_desc), cap, ptr(d_counts), C.c_float(bf), C.c_float(fx),
                                                     ptr(d_u_right), ptr(d_depth), ptr(d_status) if d_status is not None else None, 1))

    # ---- candidate windows: Frame::GetFeaturesInArea for many queries (src/Frame.cc:307-360) ----
    def window_search(self, F, qdesc, x, y, r, min_level=None, max_level=None):
        """Returns (offsets[nq+1], cand, dist): reference candidate order and DescriptorDistance of each."""
        q = _desc(qdesc)
        x, y, r = _f32(x), _f32(y), _f32(r)
        lo, hi = _i32(min_level), _i32(max_level)
        nq = len(q)
        off = np.zeros(nq + 1, np.int32)
        total = C.c_int(0)
        view = F.c_view()
        cap = max(64, 32 * nq)
        while True:
            cand = np.zeros(cap, np.int32)
            dist = np.zeros(cap, np.int32)
            rc = lib().orb_window_search(self._h, C.byref(view), nq, ptr(q), ptr(x), ptr(y), ptr(r), ptr(lo), ptr(hi),
                                         ptr(off), ptr(cand), ptr(dist), cap, C.byref(total))
            if rc == ORB_ERR_CAPACITY:
                cap = total.value
                continue
            check(rc)
            return off, cand[: total.value].copy(), dist[: total.value].copy()

    # ---- ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:19-65 ----
    def SearchByProjection(self, F, occupied, qdesc, proj_x, proj_y, proj_xr, level, view_cos, th=1.0, q_observed=None):
        """Queries = map points passing :24.  occupied (uint8, updated in place) = F.mvpMapPoints[i] with
        observations.  Returns (nmatches, feature_of_query)."""
        q = _desc(qdesc)
        out = np.full(len(q), -1, np.int32)
        n = C.c_int(0)
        view = F.c_view()
        assert occupied.dtype == np.uint8 and occupied.flags.c_contiguous
        check(lib().orb_search_by_projection_map(self._h, C.byref(view), ptr(F.mvuRight), ptr(occupied), ptr(F.mvScaleFactors),
                                                 len(F.mvScaleFactors), len(q), ptr(q), ptr(_f32(proj_x)), ptr(_f32(proj_y)),
                                                 ptr(_f32(proj_xr)), ptr(_i32(level)), ptr(_f32(view_cos)), ptr(_u8(q_observed)),
                                                 C.c_float(th), C.c_float(self.mfNNratio), ptr(out), C.byref(n)))
        return n.value, out

    def _best(self, F, claimed, qdesc, u, v, radius, lo, hi, q_angle, rot_mode, max_dist):
        q = _desc(qdesc)
        out = np.full(len(q), -1, np.int32)
        n = C.c_int(0)
        view = F.c_view()
        if claimed is not None:
            assert claimed.dtype == np.uint8 and claimed.flags.c_contiguous
        check(lib().orb_search_by_projection_best(self._h, C.byref(view), ptr(claimed), len(q), ptr(q), ptr(_f32(u)), ptr(_f32(v)),
                                                  ptr(_f32(radius)), ptr(_i32(lo)), ptr(_i32(hi)), ptr(_f32(q_angle)), rot_mode,
                                                  int(max_dist), ptr(out), C.byref(n)))
        return n.value, out

    # ---- SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono), :732-818 ----
    def SearchByProjectionLast(self, Cur, claimed, qdesc, u, v, last_octave, last_angle, th, forward=False, backward=False):
        """From the projected (u, v) on.  claimed = CurrentFrame.mvpMapPoints[i] != NULL (updated in place)."""
        oc = _i32(last_octave)
        radius = np.float32(th) * Cur.mvScaleFactors[oc]  # :768
        if forward:      # :770-771  GetFeaturesInArea(u, v, radius, nLastOctave)
            lo, hi = oc, np.full_like(oc, -1)
        elif backward:   # :772-773  (…, 0, nLastOctave)
            lo, hi = np.zeros_like(oc), oc
        else:            # :774-775  (…, nLastOctave-1, nLastOctave+1)
            lo, hi = oc - 1, oc + 1
        return self._best(Cur, claimed, qdesc, u, v, radius, lo, hi, last_angle, 2 if self.mbCheckOrientation else 0, self.TH_HIGH)

    # ---- SearchByProjection(Frame &CurrentFrame, KeyFrame*, set&, th, ORBdist), :820-894 ----
    def SearchByProjectionReloc(self, Cur, claimed, qdesc, u, v, predicted_level, kf_angle, th, ORBdist):
        pl = _i32(predicted_level)
        radius = np.float32(th) * Cur.mvScaleFactors[pl]  # :847
        return self._best(Cur, claimed, qdesc, u, v, radius, pl - 1, pl + 1, kf_angle, 1 if self.mbCheckOrientation else 0, ORBdist)

    # ---- SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), :121-195 ----
    def SearchByProjectionKF(self, KF, matched, qdesc, u, v, radius):
        """matched = vpMatched[i] != NULL (updated in place); accept <= TH_LOW (:188)."""
        return self._best(KF, matched, qdesc, u, v, radius, None, None, None, 0, self.TH_LOW)

    # ---- SearchBySim3 (:636-730, one direction) and the search inside Fuse (:504-634) ----
    def SearchBySim3(self, KF2, qdesc, u, v, radius, predicted_level):
        pl = _i32(predicted_level)
        return self._best(KF2, None, qdesc, u, v, radius, pl - 1, pl, None, 0, self.TH_HIGH)  # :699, :711

    def FuseSearch(self, KF, qdesc, u, v, radius, predicted_level):
        pl = _i32(predicted_level)
        return self._best(KF, None, qdesc, u, v, radius, pl - 1, pl, None, 0, self.TH_LOW)  # :541, :547

    # ---- SearchForInitialization, :197-276 ----
    def SearchForInitialization(self, keys1, desc1, F2, prev_matched, window_size=10):
        """prev_matched float32 [n1, 2] is updated in place (:270-273).  Returns (nmatches, vnMatches12)."""
        k1 = np.ascontiguousarray(keys1, KP_DTYPE)
        d1 = _desc(desc1)
        assert prev_matched.dtype == np.float32 and prev_matched.flags.c_contiguous and prev_matched.shape == (len(k1), 2)
        out = np.full(len(k1), -1, np.int32)
        n = C.c_int(0)
        view = F2.c_view()
        check(lib().orb_search_for_initialization(self._h, ptr(k1), ptr(d1), len(k1), C.byref(view), ptr(prev_matched), int(window_size),
                                                  C.c_float(self.mfNNratio), int(self.mbCheckOrientation), ptr(out), C.byref(n)))
        return n.value, out

    # ---- SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), :278-366 ----
    def SearchByBoW(self, desc1, angle1, has_mp1, fv1, desc2, angle2, has_mp2, fv2):
        d1, d2 = _desc(desc1), _desc(desc2)
        out = np.full(len(d1), -1, np.int32)
        n = C.c_int(0)
        s1, s2 = fv1.c_struct(), fv2.c_struct()
        check(lib().orb_search_by_bow(self._h, ptr(d1), ptr(_f32(angle1)), ptr(_u8(has_mp1)), len(d1), ptr(d2), ptr(_f32(angle2)),
                                      ptr(_u8(has_mp2)), len(d2), C.byref(s1), C.byref(s2), C.c_float(self.mfNNratio),
                                      int(self.mbCheckOrientation), ptr(out), C.byref(n)))
        return n.value, out

    # ---- SearchForTriangulation (bOnlyStereo = false), :368-467 ----
    def SearchForTriangulation(self, keys1, desc1, has_mp1, fv1, keys2, desc2, has_mp2, fv2, F12, sigma2):
        k1, k2 = np.ascontiguousarray(keys1, KP_DTYPE), np.ascontiguousarray(keys2, KP_DTYPE)
        d1, d2 = _desc(desc1), _desc(desc2)
        f = np.ascontiguousarray(np.asarray(F12, np.float32).reshape(9))
        s2 = _f32(sigma2)
        out = np.full(len(d1), -1, np.int32)
        n = C.c_int(0)
        a, b = fv1.c_struct(), fv2.c_struct()
        check(lib().orb_search_for_triangulation(self._h, ptr(k1), ptr(d1), ptr(_u8(has_mp1)), len(d1), ptr(k2), ptr(d2), ptr(_u8(has_mp2)),
                                                 len(d2), C.byref(a), C.byref(b), ptr(f), ptr(s2), len(s2), int(self.mbCheckOrientation),
                                                 ptr(out), C.byref(n)))
        return n.value, out

    # ---- SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), :88-119: a no-op stub in this fork ----
    def SearchByBoWFrame(self, n_frame_features):
        """Returns (0, vpMapPointMatches resized to F.N nulls) exactly like the reference stub."""
        return 0, np.full(int(n_frame_features), -1, np.int32)
