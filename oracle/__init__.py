"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.  The product package
(orb_slam_system_b200) never does.

The oracle restates reference src/ORBextractor.cc, src/ORBmatcher.cc:896-908
and src/Frame.cc:446-529; see oracle/orb_oracle.h for the pinning story.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborb_oracle.so")
_REF_LIB = os.path.join(_HERE, "_ref", "libref_orb.so")

KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
assert KP_DTYPE.itemsize == 28

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)


def build(force=False):
    """Compile liborb_oracle.so (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("orb_oracle.cpp", "orb_oracle.h")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, _LIB])
    return _LIB


def build_ref():
    """Compile the reference's own ORBextractor.cc against oracle/cvshim (needs /root/reference)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return _REF_LIB if os.path.exists(_REF_LIB) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_ic_angle.restype = C.c_float
        _lib.orc_extract_many.restype = C.c_double
        _lib.orc_extract_many.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_longlong)]
        _lib.orc_synth_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                         C.c_uint64, C.c_int, C.c_int]
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


DEFAULT = dict(nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7)


def synth_frame(rows, cols, seed=7, frame=0, variant=0, right=0):
    out = np.empty((rows, cols), np.uint8)
    lib().orc_synth_frame(_p(out), rows, cols, cols, seed, frame, variant, right)
    return out


def tables(nfeatures=1000, scaleFactor=1.2, nlevels=8):
    sc = np.zeros(nlevels, np.float32)
    inv = np.zeros(nlevels, np.float32)
    s2 = np.zeros(nlevels, np.float32)
    is2 = np.zeros(nlevels, np.float32)
    nf = np.zeros(nlevels, np.int32)
    um = np.zeros(16, np.int32)
    lib().orc_tables(nfeatures, C.c_float(scaleFactor), nlevels, _p(sc), _p(inv), _p(s2), _p(is2),
                     _p(nf), _p(um))
    return dict(scale=sc, inv_scale=inv, sigma2=s2, inv_sigma2=is2, features_per_level=nf, umax=um)


def extract(img, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, cap=None,
            info=None):
    """Oracle ORBextractor::operator(). Returns (keypoints[KP_DTYPE], descriptors[K,32])."""
    img = _img(img)
    cap = cap or 8 * max(nfeatures, 64)
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    cnt = C.c_int(0)
    ncand = np.zeros(nlevels, np.int32)
    nkept = np.zeros(nlevels, np.int32)
    nretry = C.c_int(0)
    rc = lib().orc_extract(nfeatures, C.c_float(scaleFactor), nlevels, iniThFAST, minThFAST,
                           _p(img), img.shape[0], img.shape[1], img.strides[0], _p(kps), _p(desc),
                           cap, C.byref(cnt), _p(ncand), _p(nkept), C.byref(nretry))
    if rc != 0:
        raise RuntimeError(f"oracle extract failed rc={rc}")
    if info is not None:
        info.update(candidates=ncand, kept=nkept, retry_cells=nretry.value)
    n = cnt.value
    return kps[:n].copy(), desc[:n].copy()


def pyramid_level(img, level, scaleFactor=1.2, nlevels=8):
    img = _img(img)
    dst = np.zeros(img.size, np.uint8)
    r = C.c_int(0)
    c = C.c_int(0)
    rc = lib().orc_pyramid(C.c_float(scaleFactor), nlevels, _p(img), img.shape[0], img.shape[1],
                           img.strides[0], level, _p(dst), dst.size, C.byref(r), C.byref(c))
    assert rc == 0
    return dst[: r.value * c.value].reshape(r.value, c.value).copy()


def resize(img, drows, dcols):
    img = _img(img)
    dst = np.zeros((drows, dcols), np.uint8)
    lib().orc_resize(_p(img), img.shape[0], img.shape[1], img.strides[0], _p(dst), drows, dcols, dcols)
    return dst


def blur7(img):
    img = _img(img)
    dst = np.zeros_like(img)
    lib().orc_blur7(_p(img), img.shape[0], img.shape[1], img.strides[0], _p(dst), dst.strides[0])
    return dst


def fast(img, threshold):
    """cv::FAST(img, threshold, nms=True) model: int32 [n,3] rows (x, y, score), row-major order."""
    img = _img(img)
    cap = max(16, img.size // 4 + 16)
    out = np.zeros((cap, 3), np.int32)
    n = lib().orc_fast(_p(img), img.shape[0], img.shape[1], img.strides[0], threshold, _p(out), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def octree(xyr, minX, maxX, minY, maxY, N):
    """DistributeOctTree: xyr float32 [n,3] (x, y, response) -> kept indices in reference order."""
    xyr = np.ascontiguousarray(xyr, np.float32).reshape(-1, 3)
    out = np.zeros(max(1, len(xyr)), np.int32)
    k = lib().orc_octree(_p(xyr), len(xyr), minX, maxX, minY, maxY, N, _p(out), len(out))
    if k < 0:
        raise RuntimeError("octree does not terminate in the reference")
    return out[:k].copy()


def ic_angle(img, x, y):
    img = _img(img)
    return float(lib().orc_ic_angle(_p(img), img.shape[0], img.shape[1], img.strides[0], x, y))


def describe(blurred, x, y, angle):
    blurred = _img(blurred)
    d = np.zeros(32, np.uint8)
    lib().orc_describe(_p(blurred), blurred.shape[0], blurred.shape[1], blurred.strides[0], x, y,
                       C.c_float(angle), _p(d))
    return d


def distance(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return int(lib().orc_distance(_p(a), _p(b)))


def match_all(q, t):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    bi = np.zeros(len(q), np.int32)
    bd = np.zeros(len(q), np.int32)
    sd = np.zeros(len(q), np.int32)
    lib().orc_match_all(_p(q), len(q), _p(t), len(t), _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


def match_csr(q, t, offsets, cand, tie_last=False, max_dist=50):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    cand = np.ascontiguousarray(cand, np.int32)
    assert len(offsets) == len(q) + 1
    bi = np.zeros(len(q), np.int32)
    bd = np.zeros(len(q), np.int32)
    sd = np.zeros(len(q), np.int32)
    lib().orc_match_csr(_p(q), len(q), _p(t), len(t), _p(offsets), _p(cand), int(tie_last),
                        int(max_dist), _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


def stereo_match(kl, dl, kr, dr, scale, rows, bf, fx):
    kl = np.ascontiguousarray(kl, KP_DTYPE)
    kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8)
    dr = np.ascontiguousarray(dr, np.uint8)
    scale = np.ascontiguousarray(scale, np.float32)
    br = np.zeros(len(kl), np.int32)
    bd = np.zeros(len(kl), np.int32)
    rc = lib().orc_stereo_match(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), _p(scale),
                                len(scale), rows, C.c_float(bf), C.c_float(fx), _p(br), _p(bd))
    if rc != 0:
        raise RuntimeError("stereo row band out of range (reference UB)")
    return br, bd


def compute_stereo_matches(img_l, img_r, kl, dl, kr, dr, bf, fx, scaleFactor=1.2, nlevels=8):
    """Frame::ComputeStereoMatches whole (Hamming search, SAD refinement, parabola, median cut): (mvuRight, mvDepth)."""
    img_l, img_r = _img(img_l), _img(img_r)
    assert img_l.shape == img_r.shape and img_l.strides == img_r.strides
    kl = np.ascontiguousarray(kl, KP_DTYPE)
    kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8)
    dr = np.ascontiguousarray(dr, np.uint8)
    ur = np.zeros(len(kl), np.float32)
    dep = np.zeros(len(kl), np.float32)
    lib().orc_compute_stereo_matches.restype = C.c_int
    rc = lib().orc_compute_stereo_matches(C.c_float(scaleFactor), nlevels, _p(img_l), _p(img_r), img_l.shape[0], img_l.shape[1],
                                          img_l.strides[0], _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), C.c_float(bf),
                                          C.c_float(fx), _p(ur), _p(dep))
    if rc != 0:
        raise RuntimeError(f"reference faults on this input (rc={rc})")
    return ur, dep


_REF_FRAME = os.path.join(_HERE, "_ref", "libref_frame.so")
_ref_frame = None


def ref_frame_lib():
    """CDLL of the reference's own src/Frame.cc compiled unmodified behind oracle/mshim/frame_objects.h (oracle/Makefile
    refframe), or None when it is neither prebuilt nor buildable (no /root/reference)."""
    global _ref_frame
    if _ref_frame is None:
        if not os.path.exists(_REF_FRAME) and os.path.exists("/root/reference/src/Frame.cc"):
            subprocess.check_call(["make", "-s", "-C", _HERE, "refframe"])
        if not os.path.exists(_REF_FRAME):
            return None
        _ref_frame = C.CDLL(_REF_FRAME)
    return _ref_frame


def ref_compute_stereo_matches(img_l, img_r, kl, dl, kr, dr, bf, fx, scaleFactor=1.2, nlevels=8):
    """Frame::ComputeStereoMatches of the COMPILED reference (its stereo constructor run on stub extractors that hand out these
    keypoints / descriptors and the pyramids of the two images): (mvuRight, mvDepth)."""
    kl = np.ascontiguousarray(kl, KP_DTYPE)
    kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8)
    dr = np.ascontiguousarray(dr, np.uint8)
    pl = [pyramid_level(img_l, l, scaleFactor, nlevels) for l in range(nlevels)]
    pr = [pyramid_level(img_r, l, scaleFactor, nlevels) for l in range(nlevels)]
    rows = np.array([p.shape[0] for p in pl], np.int32)
    cols = np.array([p.shape[1] for p in pl], np.int32)
    bl = np.concatenate([np.ascontiguousarray(p).ravel() for p in pl])
    br = np.concatenate([np.ascontiguousarray(p).ravel() for p in pr])
    scale = tables(1000, scaleFactor, nlevels)["scale"].astype(np.float32)
    ur = np.zeros(len(kl), np.float32)
    dep = np.zeros(len(kl), np.float32)
    fn = ref_frame_lib().refm_compute_stereo_matches
    fn.restype = C.c_int
    rc = fn(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), _p(scale), nlevels, _p(bl), _p(br), _p(rows), _p(cols), C.c_float(bf),
            C.c_float(fx), _p(ur), _p(dep))
    if rc != 0:
        raise RuntimeError(f"compiled reference threw (rc={rc})")
    return ur, dep


def ref_frame_features_in_area(keys, rows, cols, x, y, r, min_level=None, max_level=None):
    """Frame::AssignFeaturesToGrid + Frame::GetFeaturesInArea of the COMPILED reference on a rows x cols frame: (offsets, cand)."""
    k = np.ascontiguousarray(keys, KP_DTYPE)
    x, y, r = (np.ascontiguousarray(a, np.float32) for a in (x, y, r))
    lo, hi = _opt(min_level, np.int32), _opt(max_level, np.int32)
    off = np.zeros(len(x) + 1, np.int32)
    cap = max(64, 64 * len(x))
    fn = ref_frame_lib().refm_frame_features_in_area
    fn.restype = C.c_int
    while True:
        cand = np.zeros(cap, np.int32)
        tot = fn(_p(k), len(k), int(rows), int(cols), len(x), _p(x), _p(y), _p(r), _pp(lo), _pp(hi), _p(off), _p(cand), cap)
        if tot <= cap:
            return off, cand[:tot].copy()
        cap = tot


_REF_KEYFRAME = os.path.join(_HERE, "_ref", "libref_keyframe.so")
_ref_keyframe = None


def ref_keyframe_features_in_area(keys, rows, cols, x, y, r):
    """KeyFrame::GetFeaturesInArea of the COMPILED reference (src/KeyFrame.cc:549-588) on the KeyFrame it builds from a
    rows x cols Frame holding these keypoints: (offsets, cand); None when oracle/_ref/libref_keyframe.so is unavailable."""
    global _ref_keyframe
    if _ref_keyframe is None:
        if not os.path.exists(_REF_KEYFRAME) and os.path.exists("/root/reference/src/KeyFrame.cc"):
            subprocess.check_call(["make", "-s", "-C", _HERE, "refkeyframe"])
        if not os.path.exists(_REF_KEYFRAME):
            return None
        _ref_keyframe = C.CDLL(_REF_KEYFRAME)
    k = np.ascontiguousarray(keys, KP_DTYPE)
    x, y, r = (np.ascontiguousarray(a, np.float32) for a in (x, y, r))
    off = np.zeros(len(x) + 1, np.int32)
    cap = max(64, 64 * len(x))
    fn = _ref_keyframe.refm_keyframe_features_in_area
    fn.restype = C.c_int
    while True:
        cand = np.zeros(cap, np.int32)
        tot = fn(_p(k), len(k), int(rows), int(cols), len(x), _p(x), _p(y), _p(r), _p(off), _p(cand), cap)
        if tot <= cap:
            return off, cand[:tot].copy()
        cap = tot


def _featvec_csr(fv):
    """dict node -> list of feature indices  ->  (sorted nodes, offsets, flat indices) int32 arrays."""
    nodes = np.array(sorted(fv), np.int32)
    off = np.zeros(len(nodes) + 1, np.int32)
    flat = []
    for i, nd in enumerate(nodes):
        flat.extend(fv[int(nd)])
        off[i + 1] = len(flat)
    return nodes, off, np.array(flat, np.int32)


def search_by_bow_kf(kf1, kf2, nnratio=0.6, checkOri=True, impl=None):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*): kf = dict(desc, keys[KP_DTYPE], has_mp[bool], featvec{node: [idx]})."""
    d1, d2 = np.ascontiguousarray(kf1["desc"], np.uint8), np.ascontiguousarray(kf2["desc"], np.uint8)
    a1, a2 = np.ascontiguousarray(kf1["keys"]["angle"], np.float32), np.ascontiguousarray(kf2["keys"]["angle"], np.float32)
    h1, h2 = np.ascontiguousarray(kf1["has_mp"], np.uint8), np.ascontiguousarray(kf2["has_mp"], np.uint8)
    n1, o1, i1 = _featvec_csr(kf1["featvec"])
    n2, o2, i2 = _featvec_csr(kf2["featvec"])
    out = np.zeros(len(d1), np.int32)
    n = _search_fn("search_by_bow_kf", impl)(_p(d1), _p(a1), _p(h1), len(d1), _p(d2), _p(a2), _p(h2), len(d2), _p(n1), _p(o1), _p(i1), len(n1),
                                   _p(n2), _p(o2), _p(i2), len(n2), C.c_float(nnratio), int(checkOri), _p(out))
    return int(n), out


def search_for_triangulation(kf1, kf2, F12, checkOri=False, impl=None):
    """ORBmatcher::SearchForTriangulation (bOnlyStereo=false): kf as above plus 'sigma2' (mvLevelSigma2) for kf2."""
    d1, d2 = np.ascontiguousarray(kf1["desc"], np.uint8), np.ascontiguousarray(kf2["desc"], np.uint8)
    k1, k2 = kf1["keys"], kf2["keys"]
    f = lambda a, t=np.float32: np.ascontiguousarray(a, t)
    h1, h2 = f(kf1["has_mp"], np.uint8), f(kf2["has_mp"], np.uint8)
    n1, o1, i1 = _featvec_csr(kf1["featvec"])
    n2, o2, i2 = _featvec_csr(kf2["featvec"])
    x1, y1, a1 = f(k1["x"]), f(k1["y"]), f(k1["angle"])
    x2, y2, a2, oc2 = f(k2["x"]), f(k2["y"]), f(k2["angle"]), f(k2["octave"], np.int32)
    F = f(np.asarray(F12, np.float32).reshape(9))
    s2 = f(kf2["sigma2"])
    out = np.zeros(len(d1), np.int32)
    n = _search_fn("search_for_triangulation", impl)(_p(d1), _p(x1), _p(y1), _p(a1), _p(h1), len(d1), _p(d2), _p(x2), _p(y2), _p(a2), _p(oc2),
                                           _p(h2), len(d2), _p(n1), _p(o1), _p(i1), len(n1), _p(n2), _p(o2), _p(i2), len(n2), _p(F),
                                           _p(s2), int(checkOri), _p(out))
    return int(n), out


_REF_MATCH = os.path.join(_HERE, "_ref", "libref_match.so")
_ref_match = None


def ref_match_lib():
    """CDLL of the reference's own src/ORBmatcher.cc compiled unmodified against oracle/mshim (oracle/Makefile refmatch),
    or None when it is neither prebuilt nor buildable (no /root/reference)."""
    global _ref_match
    if _ref_match is None:
        if not os.path.exists(_REF_MATCH) and os.path.exists("/root/reference/src/ORBmatcher.cc"):
            subprocess.check_call(["make", "-s", "-C", _HERE, "refmatch"])
        if not os.path.exists(_REF_MATCH):
            return None
        _ref_match = C.CDLL(_REF_MATCH)
    return _ref_match


_ADAPTER_MATCH = os.path.join(_HERE, "_ref", "libadapter_match.so")
_adapter_match = None


def adapter_match_lib():
    """CDLL of the drop-in ORB_SLAM2::ORBmatcher (orb_slam_system_b200/adapter/ORBmatcher_b200.cc over liborb_b200.so) behind the
    same bridge as the compiled reference (oracle/Makefile adaptermatch), or None.  Needs a GPU to do anything."""
    global _adapter_match
    if _adapter_match is None:
        if not os.path.exists(_ADAPTER_MATCH):
            return None
        _adapter_match = C.CDLL(_ADAPTER_MATCH)
    return _adapter_match


def _bridge_lib(impl):
    """The library whose refm_* entry points run the ORBmatcher class: the compiled reference or the GPU adapter."""
    L = adapter_match_lib() if impl == "adapter" else ref_match_lib()
    assert L is not None, f"oracle/_ref library for impl={impl!r} is not available"
    return L


def _search_fn(name, impl):
    """orc_<name> of the oracle restatement, or refm_<name> of the compiled reference (impl="reference") / of the drop-in
    adapter class behind the same bridge (impl="adapter")."""
    if impl in ("reference", "adapter"):
        fn = getattr(_bridge_lib(impl), "refm_" + name)
    else:
        fn = getattr(lib(), "orc_" + name)
    fn.restype = C.c_int
    return fn


def ref_bruteforce_many(q, t, nthreads):
    """BASELINE config 4 with the reference's own DescriptorDistance on the host: q [npairs, nq, 32], t [npairs, nt, 32].
    Returns (seconds, best_idx, best_dist, second_dist)."""
    L = ref_match_lib()
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    npairs, nq, _ = q.shape
    nt = t.shape[1]
    bi = np.zeros((npairs, nq), np.int32)
    bd = np.zeros((npairs, nq), np.int32)
    sd = np.zeros((npairs, nq), np.int32)
    L.refm_bruteforce_many.restype = C.c_double
    secs = L.refm_bruteforce_many(_p(q), _p(t), npairs, nq, nt, int(nthreads), _p(bi), _p(bd), _p(sd))
    return secs, bi, bd, sd


def ref_descriptor_distance(a, b):
    """ORBmatcher::DescriptorDistance of the compiled reference."""
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return int(ref_match_lib().refm_descriptor_distance(_p(a), _p(b)))


class _OrcFrame(C.Structure):
    _fields_ = [("keysUn", C.c_void_p), ("desc", C.c_void_p), ("N", C.c_int), ("mnMinX", C.c_float), ("mnMinY", C.c_float),
                ("mfGridElementWidthInv", C.c_float), ("mfGridElementHeightInv", C.c_float)]


def _frame(F):
    """F: any object with keys_un, desc, mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv."""
    k = np.ascontiguousarray(F.keys_un, KP_DTYPE)
    d = np.ascontiguousarray(F.desc, np.uint8)
    fr = _OrcFrame(k.ctypes.data, d.ctypes.data, len(k), F.mnMinX, F.mnMinY, F.mfGridElementWidthInv, F.mfGridElementHeightInv)
    fr._keep = (k, d)
    return fr


def _opt(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


def _pp(a):
    return None if a is None else _p(a)


def features_in_area(F, x, y, r, min_level=None, max_level=None):
    """Frame::GetFeaturesInArea per query: (offsets, cand) in the reference's order."""
    x, y, r = (np.ascontiguousarray(a, np.float32) for a in (x, y, r))
    lo, hi = _opt(min_level, np.int32), _opt(max_level, np.int32)
    fr = _frame(F)
    off = np.zeros(len(x) + 1, np.int32)
    cap = 64
    while True:
        cand = np.zeros(cap, np.int32)
        lib().orc_features_in_area.restype = C.c_int
        tot = lib().orc_features_in_area(C.byref(fr), len(x), _p(x), _p(y), _p(r), _pp(lo), _pp(hi), _p(off), _p(cand), cap)
        if tot <= cap:
            return off, cand[:tot].copy()
        cap = tot


def search_by_projection_map(F, occupied, qdesc, proj_x, proj_y, proj_xr, level, view_cos, th, nnratio, q_observed=None, impl=None):
    q = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
    out = np.zeros(len(q), np.int32)
    fr = _frame(F)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    px, py, pxr, vc = f(proj_x), f(proj_y), f(proj_xr), f(view_cos)
    lv = np.ascontiguousarray(level, np.int32)
    qo = _opt(q_observed, np.uint8)
    n = _search_fn("search_by_projection_map", impl)(C.byref(fr), _pp(_opt(F.mvuRight, np.float32)), _p(occupied), _p(f(F.mvScaleFactors)), len(q),
                                           _p(q), _p(px), _p(py), _p(pxr), _p(lv), _p(vc), _pp(qo), C.c_float(th), C.c_float(nnratio),
                                           _p(out))
    return int(n), out


def search_by_projection_last(Cur, claimed, qdesc, u, v, last_octave, last_angle, th, forward, backward, checkOri, impl=None):
    """Returns (nmatches, feature_of_query); nmatches is None when the reference would index rotHist out of bounds (D9)."""
    q = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
    out = np.zeros(len(q), np.int32)
    fr = _frame(Cur)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    uu, vv, la = f(u), f(v), f(last_angle)
    lo = np.ascontiguousarray(last_octave, np.int32)
    n = _search_fn("search_by_projection_last", impl)(C.byref(fr), _p(claimed), _p(f(Cur.mvScaleFactors)), len(q), _p(q), _p(uu), _p(vv), _p(lo),
                                            _p(la), C.c_float(th), int(forward), int(backward), int(checkOri), _p(out))
    return (None if n == -2147483648 else int(n)), out


def search_by_projection_reloc(Cur, claimed, qdesc, u, v, predicted_level, kf_angle, th, ORBdist, checkOri, impl=None):
    q = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
    out = np.zeros(len(q), np.int32)
    fr = _frame(Cur)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    uu, vv, ka = f(u), f(v), f(kf_angle)
    pl = np.ascontiguousarray(predicted_level, np.int32)
    n = _search_fn("search_by_projection_reloc", impl)(C.byref(fr), _p(claimed), _p(f(Cur.mvScaleFactors)), len(q), _p(q), _p(uu), _p(vv), _p(pl),
                                             _p(ka), C.c_float(th), int(ORBdist), int(checkOri), _p(out))
    return int(n), out


def search_kf_window(KF, claimed, qdesc, u, v, radius, level, max_dist):
    q = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
    out = np.zeros(len(q), np.int32)
    fr = _frame(KF)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    uu, vv, rr = f(u), f(v), f(radius)
    lv = _opt(level, np.int32)
    lib().orc_search_kf_window.restype = C.c_int
    n = lib().orc_search_kf_window(C.byref(fr), _pp(claimed), len(q), _p(q), _p(uu), _p(vv), _p(rr), _pp(lv), int(max_dist), _p(out))
    return int(n), out


def ref_search_by_projection_loop(KF, claimed, qdesc, u, v, radius, impl="reference"):
    """The compiled reference's SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (src/ORBmatcher.cc:121-195) searching
    the windows (u, v, radius); the oracle's counterpart is search_kf_window(KF, claimed, ..., level=None, max_dist=TH_LOW)."""
    q = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
    out = np.zeros(len(q), np.int32)
    fr = _frame(KF)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    fn = _bridge_lib(impl).refm_search_by_projection_loop
    fn.restype = C.c_int
    n = fn(C.byref(fr), _p(claimed), len(q), _p(q), _p(f(u)), _p(f(v)), _p(f(radius)), _p(out))
    return int(n), out


def ref_fuse_search(KF, qdesc, u, v, level, th, variant=0, impl="reference"):
    """The compiled reference's Fuse (src/ORBmatcher.cc:504-568, variant 1: the Scw overload :570-634), search part; the oracle's
    counterpart is search_kf_window(KF, None, ..., radius = th * scale[level], level, max_dist=TH_LOW)."""
    q = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
    out = np.zeros(len(q), np.int32)
    fr = _frame(KF)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    fn = _bridge_lib(impl).refm_fuse_search
    fn.restype = C.c_int
    n = fn(C.byref(fr), _p(f(KF.mvScaleFactors)), len(q), _p(q), _p(f(u)), _p(f(v)), _p(np.ascontiguousarray(level, np.int32)), C.c_float(th),
           int(variant), _p(out))
    return int(n), out


def ref_search_by_sim3(KF2, qdesc, u, v, level, th, impl="reference"):
    """The compiled reference's SearchBySim3 (src/ORBmatcher.cc:636-730) with s12 = 1, R12 = I, t12 = 0; the oracle's counterpart is
    search_kf_window(KF2, None, ..., radius = th * scale[level], level, max_dist=TH_HIGH)."""
    q = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
    out = np.zeros(len(q), np.int32)
    fr = _frame(KF2)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    fn = _bridge_lib(impl).refm_search_by_sim3
    fn.restype = C.c_int
    n = fn(C.byref(fr), _p(f(KF2.mvScaleFactors)), len(q), _p(q), _p(f(u)), _p(f(v)), _p(np.ascontiguousarray(level, np.int32)), C.c_float(th),
           _p(out))
    return int(n), out


def search_for_initialization(keys1, desc1, F2, prev_matched, window_size, nnratio, checkOri, impl=None):
    k1 = np.ascontiguousarray(keys1, KP_DTYPE)
    d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32)
    assert prev_matched.dtype == np.float32 and prev_matched.flags.c_contiguous
    out = np.zeros(len(k1), np.int32)
    fr = _frame(F2)
    n = _search_fn("search_for_initialization", impl)(_p(k1), _p(d1), len(k1), C.byref(fr), _p(prev_matched), int(window_size), C.c_float(nnratio),
                                            int(checkOri), _p(out))
    return int(n), out


_REF_DBOW = os.path.join(_HERE, "_ref", "libref_dbow.so")


def _voc_arrays(voc):
    """voc: object with child_off, children, node_desc, node_weight, node_word, L, weighting, scoring."""
    return (np.ascontiguousarray(voc.child_off, np.int32), np.ascontiguousarray(voc.children, np.int32),
            np.ascontiguousarray(voc.node_desc, np.uint8), np.ascontiguousarray(voc.node_weight, np.float64),
            np.ascontiguousarray(voc.node_word, np.int32))


def voc_transform(voc, desc, levelsup=4):
    """Oracle restatement of TemplatedVocabulary::transform: returns dict(words, nodes, bow_ids, bow_values, fv)."""
    co, ch, nd, nw, wd = _voc_arrays(voc)
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(d)
    words, nodes = np.zeros(n, np.int32), np.zeros(n, np.int32)
    bi, bv = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.float64)
    fn, fo, fi = np.zeros(n + 1, np.int32), np.zeros(n + 2, np.int32), np.zeros(n + 1, np.int32)
    nb, nf = C.c_int(0), C.c_int(0)
    lib().orc_voc_transform.restype = C.c_int
    rc = lib().orc_voc_transform(len(co) - 1, _p(co), _p(ch), _p(nd), _p(nw), _p(wd), int(voc.L), int(voc.weighting), int(voc.scoring),
                                 _p(d), n, int(levelsup), _p(words), _p(nodes), _p(bi), _p(bv), n + 1, C.byref(nb), _p(fn), _p(fo),
                                 _p(fi), n + 1, C.byref(nf))
    if rc != 0:
        raise RuntimeError(f"oracle voc_transform rc={rc}")
    fv = {int(fn[k]): fi[fo[k]:fo[k + 1]].tolist() for k in range(nf.value)}
    return dict(words=words, nodes=nodes, bow_ids=bi[: nb.value].copy(), bow_values=bv[: nb.value].copy(), fv=fv)


class RefVocabulary:
    """The reference's own DBoW2 TemplatedVocabulary<FORB> (oracle/_ref/libref_dbow.so), loaded from the fork's
    text format exactly as ORB_SLAM2::System does (src/System.cc:43-51: loadFromTextFile)."""

    def __init__(self, text_path):
        if not os.path.exists(_REF_DBOW):
            subprocess.check_call(["make", "-s", "-C", _HERE, "refdbow"])
        if not os.path.exists(_REF_DBOW):
            raise FileNotFoundError(_REF_DBOW)
        self._L = C.CDLL(_REF_DBOW)
        self._L.refvoc_load_text.restype = C.c_void_p
        self._L.refvoc_load_text.argtypes = [C.c_char_p]
        self._L.refvoc_free.argtypes = [C.c_void_p]
        self._L.refvoc_size.argtypes = [C.c_void_p]
        self._h = self._L.refvoc_load_text(text_path.encode())
        if not self._h:
            raise RuntimeError("loadFromTextFile failed")

    def size(self):
        return self._L.refvoc_size(self._h)

    def transform(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        bi, bv = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.float64)
        fn, fo, fi = np.zeros(n + 1, np.int32), np.zeros(n + 2, np.int32), np.zeros(n + 1, np.int32)
        nb, nf = C.c_int(0), C.c_int(0)
        self._L.refvoc_transform.argtypes = [C.c_void_p] + [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        rc = self._L.refvoc_transform(self._h, _p(d), n, int(levelsup), _p(bi), _p(bv), n + 1, C.byref(nb), _p(fn), _p(fo), _p(fi), n + 1,
                                      C.byref(nf))
        assert rc == 0
        fv = {int(fn[k]): fi[fo[k]:fo[k + 1]].tolist() for k in range(nf.value)}
        return dict(bow_ids=bi[: nb.value].copy(), bow_values=bv[: nb.value].copy(), fv=fv)

    def __del__(self):
        try:
            self._L.refvoc_free(self._h)
        except Exception:
            pass


def cvt_gray(img, bgr=False, variant=4):
    """cv::cvtColor(img, CV_RGB2GRAY / CV_BGR2GRAY) model on an 8-bit [rows, cols, 3|4] image."""
    img = np.ascontiguousarray(img, np.uint8)
    rows, cols, ch = img.shape
    dst = np.zeros((rows, cols), np.uint8)
    lib().orc_cvt_gray(_p(img), rows, cols, img.strides[0], ch, int(bgr), int(variant), _p(dst), cols)
    return dst


def remap_linear(img, mapx, mapy):
    """cv::remap(img, map1, map2, INTER_LINEAR) model (CV_32FC1 maps, BORDER_CONSTANT 0); img [rows, cols] or [rows, cols, ch]."""
    img = np.ascontiguousarray(img, np.uint8)
    ch = 1 if img.ndim == 2 else img.shape[2]
    mx = np.ascontiguousarray(mapx, np.float32)
    my = np.ascontiguousarray(mapy, np.float32)
    assert mx.shape == my.shape
    dr, dc = mx.shape
    dst = np.zeros((dr, dc) if img.ndim == 2 else (dr, dc, ch), np.uint8)
    lib().orc_remap_linear(_p(img), img.shape[0], img.shape[1], img.strides[0], ch, _p(mx), _p(my), dc, dr, dc, _p(dst), dc * ch)
    return dst


def extract_many(rows, cols, nframes, nthreads, first_frame=0, seed=7, **params):
    p = dict(DEFAULT)
    p.update(params)
    tot = C.c_longlong(0)
    secs = lib().orc_extract_many(p["nfeatures"], C.c_float(p["scaleFactor"]), p["nlevels"],
                                  p["iniThFAST"], p["minThFAST"], rows, cols, seed, first_frame,
                                  nframes, nthreads, C.byref(tot))
    return secs, tot.value


# ---- the reference's own extractor (oracle/_ref/libref_orb.so), when built --------------------
_ref = None


def ref_lib():
    """CDLL of the reference ORBextractor.cc compiled against oracle/cvshim, or None."""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_LIB) and os.path.exists("/root/reference/src/ORBextractor.cc"):
            build_ref()
        if not os.path.exists(_REF_LIB):
            return None
        _ref = C.CDLL(_REF_LIB)
        _ref.ref_extract_many.restype = C.c_double
        _ref.ref_extract_many.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_longlong)]
    return _ref


_ref_fast = None
_REF_FAST_LIB = os.path.join(os.path.dirname(_REF_LIB), "libref_orb_fast.so")


def ref_fast_lib():
    """CDLL of the timing-only build of the reference extractor (oracle/Makefile reffast: -O3 -march=x86-64-v3), or None."""
    global _ref_fast
    if _ref_fast is None:
        if not os.path.exists(_REF_FAST_LIB):
            return None
        _ref_fast = C.CDLL(_REF_FAST_LIB)
        _ref_fast.ref_extract_many.restype = C.c_double
        _ref_fast.ref_extract_many.argtypes = ref_lib().ref_extract_many.argtypes if ref_lib() is not None else None
    return _ref_fast


def ref_extract(img, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, cap=None):
    img = _img(img)
    cap = cap or 8 * max(nfeatures, 64)
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    cnt = C.c_int(0)
    rc = ref_lib().ref_extract(nfeatures, C.c_float(scaleFactor), nlevels, iniThFAST, minThFAST,
                               _p(img), img.shape[0], img.shape[1], img.strides[0], _p(kps),
                               _p(desc), cap, C.byref(cnt))
    if rc != 0:
        raise RuntimeError(f"reference extract failed rc={rc}")
    return kps[: cnt.value].copy(), desc[: cnt.value].copy()


def ref_extract_many(rows, cols, nframes, nthreads, first_frame=0, seed=7, fast=False, **params):
    """Seconds the reference extractor needs for nframes synthetic frames on nthreads host threads (frame synthesis is
    outside the clock) and the keypoint total.  fast: the -O3 timing-only build (ref_fast_lib) instead of the parity build."""
    p = dict(DEFAULT)
    p.update(params)
    tot = C.c_longlong(0)
    secs = (ref_fast_lib() if fast else ref_lib()).ref_extract_many(p["nfeatures"], C.c_float(p["scaleFactor"]), p["nlevels"],
                                      p["iniThFAST"], p["minThFAST"], rows, cols, seed, first_frame,
                                      nframes, nthreads, C.byref(tot))
    return secs, tot.value
