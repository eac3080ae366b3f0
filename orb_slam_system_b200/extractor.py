"""Host-side mirror of the reference's ``ORB_SLAM2::ORBextractor`` over the C ABI.

Same constructor arguments, getters and call semantics as reference
include/ORBextractor.h:26-93 / src/ORBextractor.cc:442-495; the work happens in
hand-written sm_100a kernels behind ``include/orb_b200.h``.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, OrbIngestConfig, OrbParams, check, lib, ptr


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_batch=1, device=0,
                 max_rows=0, max_cols=0):
        self.nfeatures = int(nfeatures)
        self.scaleFactor = float(np.float32(scaleFactor))
        self.nlevels = int(nlevels)
        self.iniThFAST = int(iniThFAST)
        self.minThFAST = int(minThFAST)
        self.max_batch = int(max_batch)
        self.device = int(device)
        self._h = C.c_void_p()
        p = OrbParams(self.nfeatures, self.scaleFactor, self.nlevels, self.iniThFAST, self.minThFAST)
        check(lib().orb_extractor_create(C.byref(p), max_rows, max_cols, self.max_batch, self.device,
                                         C.byref(self._h)))
        n = self.nlevels
        self._scale = np.zeros(n, np.float32)
        self._inv_scale = np.zeros(n, np.float32)
        self._sigma2 = np.zeros(n, np.float32)
        self._inv_sigma2 = np.zeros(n, np.float32)
        self._nfeat = np.zeros(n, np.int32)
        check(lib().orb_extractor_tables(self._h, ptr(self._scale), ptr(self._inv_scale), ptr(self._sigma2),
                                         ptr(self._inv_sigma2), ptr(self._nfeat)))
        self._last_frames = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().orb_extractor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (include/ORBextractor.h:43-63) ----
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return self.scaleFactor

    def GetScaleFactors(self):
        return self._scale.copy()

    def GetInverseScaleFactors(self):
        return self._inv_scale.copy()

    def GetScaleSigmaSquares(self):
        return self._sigma2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._inv_sigma2.copy()

    @property
    def mnFeaturesPerLevel(self):
        return self._nfeat.copy()

    def keypoint_bound(self, rows, cols):
        b = C.c_int(0)
        check(lib().orb_extractor_keypoint_bound(self._h, rows, cols, C.byref(b)))
        return b.value

    # ---- operator() ----
    def __call__(self, image, mask=None, cap=None):
        """ORBextractor::operator()(image, mask, keypoints, descriptors); mask is ignored
        (include/ORBextractor.h:38).  Returns (keypoints[KP_DTYPE], descriptors uint8[K,32]).
        An empty image returns empty outputs (src/ORBextractor.cc:444-445)."""
        image = np.ascontiguousarray(image, np.uint8) if image is not None else None
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if image.ndim != 2:
            raise ValueError("CV_8UC1 image expected")
        res = self.extract_batch(image[None], cap=cap)
        return res[0]

    def extract_batch(self, images, cap=None):
        """n same-shape host frames [n, rows, cols] -> list of (keypoints, descriptors)."""
        images = np.ascontiguousarray(images, np.uint8)
        if images.ndim != 3:
            raise ValueError("images must be [n, rows, cols] uint8 (CV_8UC1 frames)")
        n, rows, cols = images.shape
        if n > self.max_batch:
            raise ValueError(f"{n} frames exceed max_batch={self.max_batch}")
        cap = int(cap) if cap is not None else max(1, self.keypoint_bound(rows, cols))
        if cap < 1:
            raise ValueError("cap must be positive")
        kps = np.zeros((n, cap), KP_DTYPE)
        desc = np.zeros((n, cap, 32), np.uint8)
        counts = np.zeros(n, np.int32)
        check(lib().orb_extract_batch(self._h, n, ptr(images), rows, cols, images.strides[1], images.strides[0],
                                      ptr(kps), ptr(desc), cap, ptr(counts)))
        self._last_frames = n
        return [(kps[f, :counts[f]].copy(), desc[f, :counts[f]].copy()) for f in range(n)]

    def extract_batch_pinned(self, images, kps, desc, counts, cap):
        """Same as extract_batch but with caller-owned (ideally pinned) host buffers, no copies
        on the Python side.  images uint8 [n, rows, cols]; kps [n, cap] KP_DTYPE (or bytes
        [n, cap, 28]); desc [n, cap, 32]; counts int32 [n]."""
        n, rows, cols = images.shape
        check(lib().orb_extract_batch(self._h, n, ptr(images), rows, cols, int(images.stride(1) if hasattr(images, "stride") else images.strides[1]),
                                      int(images.stride(0) if hasattr(images, "stride") else images.strides[0]),
                                      ptr(kps), ptr(desc), cap, ptr(counts)))
        self._last_frames = n

    def submit_batch_pinned(self, images, kps, desc, counts, cap):
        """Asynchronous extract_batch_pinned: returns a ticket for wait_batch; three batches may be in flight."""
        n, rows, cols = images.shape
        t = C.c_int(-1)
        check(lib().orb_extract_batch_submit(self._h, n, ptr(images), rows, cols,
                                             int(images.stride(1) if hasattr(images, "stride") else images.strides[1]),
                                             int(images.stride(0) if hasattr(images, "stride") else images.strides[0]),
                                             ptr(kps), ptr(desc), cap, ptr(counts), C.byref(t)))
        self._last_frames = n
        return t.value

    def wait_batch(self, ticket):
        check(lib().orb_extract_batch_wait(self._h, ticket))

    def extract_batch_device(self, d_images, d_kps, d_desc, d_counts, cap):
        """Device-resident batch: torch CUDA tensors (uint8 [n, rows, pitch>=cols] view with the
        true `cols` given by d_images.shape[2]); asynchronous on the handle's stream."""
        n, rows, cols = d_images.shape
        check(lib().orb_extract_batch_device(self._h, n, ptr(d_images), rows, cols, d_images.stride(1),
                                             d_images.stride(0), ptr(d_kps), ptr(d_desc), cap, ptr(d_counts)))
        self._last_frames = n

    def sync(self):
        check(lib().orb_extractor_sync(self._h))

    # ---- image ingest: cv::remap + cv::cvtColor fused into the level-0 load ----
    def set_ingest(self, src_shape=None, maps=None, bgr=False, gray_variant=4):
        """What the reference does to a raw frame before operator(): cv::remap(raw, M1, M2, INTER_LINEAR)
        (Examples/Stereo/stereo_euroc.cc:136-137) and cv::cvtColor(..., CV_RGB2GRAY / CV_BGR2GRAY)
        (src/Tracking.cc:118-126).  src_shape = (rows, cols[, channels]); maps = (map_x, map_y) float32
        [dst_rows, dst_cols] from cv::initUndistortRectifyMap, or None; bgr = not Camera.RGB.
        src_shape None clears the configuration."""
        if src_shape is None:
            check(lib().orb_extractor_set_ingest(self._h, None))
            self._ingest = None
            return
        rows, cols = int(src_shape[0]), int(src_shape[1])
        ch = int(src_shape[2]) if len(src_shape) > 2 else 1
        cfg = OrbIngestConfig(rows, cols, ch, int(bool(bgr)), int(gray_variant), rows, cols, None, None)
        keep = None
        if maps is not None:
            mx = np.ascontiguousarray(maps[0], np.float32)
            my = np.ascontiguousarray(maps[1], np.float32)
            if mx.ndim != 2 or mx.shape != my.shape:
                raise ValueError("map_x and map_y must be 2-D arrays of the same shape")
            cfg.dst_rows, cfg.dst_cols = mx.shape
            cfg.map_x, cfg.map_y = mx.ctypes.data, my.ctypes.data
            keep = (mx, my)
        check(lib().orb_extractor_set_ingest(self._h, C.byref(cfg)))
        del keep
        self._ingest = (rows, cols, ch, cfg.dst_rows, cfg.dst_cols)

    def ingest_extract_batch(self, raw, cap=None):
        """n raw host frames [n, rows, cols] or [n, rows, cols, channels] -> list of (keypoints, descriptors)
        of the remapped / gray-converted frames."""
        raw = np.ascontiguousarray(raw, np.uint8)
        n = raw.shape[0]
        if getattr(self, "_ingest", None) is None:
            raise ValueError("set_ingest first")
        if n > self.max_batch or raw.shape[1:3] != self._ingest[:2]:
            raise ValueError(f"raw frames {raw.shape} do not match the ingest configuration {self._ingest[:2]} / max_batch={self.max_batch}")
        drows, dcols = self._ingest[3:5]
        cap = int(cap) if cap is not None else max(1, self.keypoint_bound(drows, dcols))
        if cap < 1:
            raise ValueError("cap must be positive")
        kps = np.zeros((n, cap), KP_DTYPE)
        desc = np.zeros((n, cap, 32), np.uint8)
        counts = np.zeros(n, np.int32)
        check(lib().orb_ingest_extract_batch(self._h, n, ptr(raw), raw.strides[1], raw.strides[0], ptr(kps), ptr(desc),
                                             cap, ptr(counts)))
        self._last_frames = n
        return [(kps[f, :counts[f]].copy(), desc[f, :counts[f]].copy()) for f in range(n)]

    def submit_ingest_pinned(self, raw, kps, desc, counts, cap):
        """Asynchronous form on caller-owned (pinned) buffers; raw [n, rows, cols(, channels)]; wait with wait_batch."""
        n = raw.shape[0]
        st = (lambda i: int(raw.stride(i))) if hasattr(raw, "stride") else (lambda i: int(raw.strides[i]))
        t = C.c_int(-1)
        check(lib().orb_ingest_extract_batch_submit(self._h, n, ptr(raw), st(1), st(0), ptr(kps), ptr(desc), cap, ptr(counts),
                                                    C.byref(t)))
        self._last_frames = n
        return t.value

    def ingest_extract_batch_device(self, d_raw, d_kps, d_desc, d_counts, cap):
        """Device-resident raw frames (torch CUDA uint8 [n, rows, cols(, channels)]); asynchronous on the handle's stream."""
        n = d_raw.shape[0]
        check(lib().orb_ingest_extract_batch_device(self._h, n, ptr(d_raw), d_raw.stride(1), d_raw.stride(0), ptr(d_kps),
                                                    ptr(d_desc), cap, ptr(d_counts)))
        self._last_frames = n

    @property
    def stream(self):
        return lib().orb_extractor_stream(self._h)

    def pyramid_level(self, level, frame=0):
        r, c = C.c_int(0), C.c_int(0)
        check(lib().orb_get_pyramid_level(self._h, frame, level, None, 0, C.byref(r), C.byref(c)))
        out = np.zeros((r.value, c.value), np.uint8)
        check(lib().orb_get_pyramid_level(self._h, frame, level, ptr(out), out.strides[0], C.byref(r), C.byref(c)))
        return out

    def pyramid_levels(self, frame=0):
        """All level images of one frame of the last call with one set of copies and one synchronisation."""
        r, c = C.c_int(0), C.c_int(0)
        outs = []
        for l in range(self.nlevels):
            check(lib().orb_get_pyramid_level(self._h, frame, l, None, 0, C.byref(r), C.byref(c)))
            outs.append(np.zeros((r.value, c.value), np.uint8))
        dst = (C.c_void_p * self.nlevels)(*[o.ctypes.data for o in outs])
        stride = (C.c_size_t * self.nlevels)(*[o.strides[0] for o in outs])
        check(lib().orb_get_pyramid_levels(self._h, frame, dst, stride))
        return outs

    @property
    def mvImagePyramid(self):
        """Level images of frame 0 of the last call (tight crops; the reference keeps a 19-px
        reflected border around each, which nothing on this path reads)."""
        return self.pyramid_levels(0)

    def set_profiling(self, on=True):
        check(lib().orb_extractor_set_profiling(self._h, 1 if on else 0))

    def stage_times(self):
        """{stage name: summed ms} over the calls recorded since set_profiling(True), and the call count."""
        ms = np.zeros(5, np.float64)
        n = C.c_int(0)
        check(lib().orb_extractor_stage_times(self._h, ptr(ms), C.byref(n)))
        return {lib().orb_stage_name(i).decode(): float(ms[i]) for i in range(5)}, n.value

    def level_stats(self, frame=0):
        cand = np.zeros(self.nlevels, np.int32)
        kept = np.zeros(self.nlevels, np.int32)
        check(lib().orb_extractor_level_stats(self._h, frame, ptr(cand), ptr(kept)))
        return cand, kept
