"""ctypes binding of liborb_b200.so (include/orb_b200.h).

The shared library is built in-tree by ``orb_slam_system_b200/csrc/Makefile``
(``__graft_entry__.build()``).  There is no CPU fallback: if the library is
missing, or no CUDA device is usable, every entry point raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborb_b200.so")

ORB_OK = 0
ORB_ERR_INVALID = -1
ORB_ERR_SHAPE = -2
ORB_ERR_CAPACITY = -3
ORB_ERR_CUDA = -4
ORB_ERR_UNSEPARABLE = -5

# cv::KeyPoint layout (28 bytes), reference include/Frame.h:117-118
KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
assert KP_DTYPE.itemsize == 28


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class OrbFrameView(C.Structure):
    """orb_frame_view: what Frame::GetFeaturesInArea reads (src/Frame.cc:307-360)."""
    _fields_ = [("keys_un", C.c_void_p), ("desc", C.c_void_p), ("n", C.c_int32), ("min_x", C.c_float),
                ("min_y", C.c_float), ("grid_w_inv", C.c_float), ("grid_h_inv", C.c_float)]


class OrbFeatureVector(C.Structure):
    """orb_feature_vector: DBoW2::FeatureVector as CSR."""
    _fields_ = [("nodes", C.c_void_p), ("off", C.c_void_p), ("idx", C.c_void_p), ("n_nodes", C.c_int32)]


class OrbIngestConfig(C.Structure):
    """orb_ingest_config: raw frame -> cv::remap -> cv::cvtColor gray (include/orb_b200.h)."""
    _fields_ = [("src_rows", C.c_int32), ("src_cols", C.c_int32), ("channels", C.c_int32), ("bgr", C.c_int32),
                ("gray_variant", C.c_int32), ("dst_rows", C.c_int32), ("dst_cols", C.c_int32),
                ("map_x", C.c_void_p), ("map_y", C.c_void_p)]


class OrbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"orb_b200 error {code}: {msg}")
        self.code = code


def build(verbose=False):
    """Compile liborb_b200.so for sm_100a with nvcc (in-tree)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None

# every symbol include/orb_b200.h declares
EXPORTS = [
    "orb_extractor_create", "orb_extractor_destroy", "orb_extractor_tables",
    "orb_extractor_keypoint_bound", "orb_extract", "orb_extract_batch", "orb_extract_batch_submit", "orb_extract_batch_wait", "orb_extract_batch_device",
    "orb_extractor_sync", "orb_extractor_stream", "orb_get_pyramid_level", "orb_get_pyramid_levels",
    "orb_extractor_set_ingest", "orb_ingest_extract_batch", "orb_ingest_extract_batch_submit", "orb_ingest_extract_batch_device",
    "orb_extractor_level_stats", "orb_extractor_set_profiling", "orb_extractor_stage_times",
    "orb_stage_name", "orb_matcher_create", "orb_matcher_destroy", "orb_match_all",
    "orb_match_all_batch", "orb_match_csr", "orb_distances_csr", "orb_stereo_match", "orb_compute_stereo_matches", "orb_compute_stereo_matches_mb", "orb_compute_stereo_matches_batch", "orb_matcher_sync",
    "orb_matcher_stream", "orb_window_search", "orb_search_by_projection_map", "orb_search_by_projection_best",
    "orb_search_for_initialization", "orb_search_by_bow", "orb_search_for_triangulation", "orb_vocabulary_create", "orb_vocabulary_destroy", "orb_vocabulary_transform",
    "orb_host_alloc", "orb_host_free", "orb_last_error", "orb_kernel_launch_count", "orb_version",
]


def lib():
    """The loaded CDLL.  Raises (loudly) if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(orb_slam_system_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        L.orb_extractor_create.argtypes = [C.POINTER(OrbParams), i32, i32, i32, i32, C.POINTER(vp)]
        L.orb_extractor_destroy.argtypes = [vp]
        L.orb_extractor_destroy.restype = None
        L.orb_extractor_tables.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orb_extractor_keypoint_bound.argtypes = [vp, i32, i32, C.POINTER(i32)]
        L.orb_extract.argtypes = [vp, vp, i32, i32, sz, vp, vp, i32, C.POINTER(i32)]
        L.orb_extract_batch.argtypes = [vp, i32, vp, i32, i32, sz, sz, vp, vp, i32, vp]
        L.orb_extract_batch_submit.argtypes = [vp, i32, vp, i32, i32, sz, sz, vp, vp, i32, vp, C.POINTER(i32)]
        L.orb_extract_batch_wait.argtypes = [vp, i32]
        L.orb_extract_batch_device.argtypes = [vp, i32, vp, i32, i32, sz, sz, vp, vp, i32, vp]
        L.orb_extractor_sync.argtypes = [vp]
        L.orb_extractor_set_ingest.argtypes = [vp, C.POINTER(OrbIngestConfig)]
        L.orb_ingest_extract_batch.argtypes = [vp, i32, vp, sz, sz, vp, vp, i32, vp]
        L.orb_ingest_extract_batch_submit.argtypes = [vp, i32, vp, sz, sz, vp, vp, i32, vp, C.POINTER(i32)]
        L.orb_ingest_extract_batch_device.argtypes = [vp, i32, vp, sz, sz, vp, vp, i32, vp]
        L.orb_extractor_stream.argtypes = [vp]
        L.orb_extractor_stream.restype = vp
        L.orb_get_pyramid_level.argtypes = [vp, i32, i32, vp, sz, C.POINTER(i32), C.POINTER(i32)]
        L.orb_get_pyramid_levels.argtypes = [vp, i32, vp, vp]
        L.orb_extractor_level_stats.argtypes = [vp, i32, vp, vp]
        L.orb_extractor_set_profiling.argtypes = [vp, i32]
        L.orb_extractor_stage_times.argtypes = [vp, vp, C.POINTER(i32)]
        L.orb_stage_name.argtypes = [i32]
        L.orb_stage_name.restype = C.c_char_p
        L.orb_matcher_create.argtypes = [i32, C.POINTER(vp)]
        L.orb_matcher_destroy.argtypes = [vp]
        L.orb_matcher_destroy.restype = None
        L.orb_match_all.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp]
        L.orb_match_all_batch.argtypes = [vp, i32, vp, vp, sz, vp, vp, sz, vp, vp, vp, sz, i32]
        L.orb_match_csr.argtypes = [vp, vp, i32, vp, i32, vp, vp, i32, i32, vp, vp, vp]
        L.orb_distances_csr.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp]
        L.orb_stereo_match.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, i32, i32, C.c_float, C.c_float, vp, vp]
        L.orb_compute_stereo_matches.argtypes = [vp, vp, i32, vp, i32, vp, vp, i32, vp, vp, i32, C.c_float, C.c_float, vp, vp]
        L.orb_compute_stereo_matches_batch.argtypes = [vp, vp, i32, vp, vp, i32, vp, C.c_float, C.c_float, vp, vp, vp, i32]
        L.orb_compute_stereo_matches_mb.argtypes = [vp, vp, i32, vp, i32, vp, vp, i32, vp, vp, i32, C.c_float, C.c_float, vp, vp]
        L.orb_matcher_sync.argtypes = [vp]
        L.orb_matcher_stream.argtypes = [vp]
        L.orb_matcher_stream.restype = vp
        f32 = C.c_float
        L.orb_window_search.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]
        L.orb_search_by_projection_map.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, f32, f32, vp, vp]
        L.orb_search_by_projection_best.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp]
        L.orb_search_for_initialization.argtypes = [vp, vp, vp, i32, vp, vp, i32, f32, i32, vp, vp]
        L.orb_search_by_bow.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, f32, i32, vp, vp]
        L.orb_search_for_triangulation.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, vp, vp]
        L.orb_vocabulary_create.argtypes = [i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, vp]
        L.orb_vocabulary_destroy.argtypes = [vp]
        L.orb_vocabulary_destroy.restype = None
        L.orb_vocabulary_transform.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp]
        L.orb_last_error.restype = C.c_char_p
        L.orb_kernel_launch_count.restype = C.c_uint64
        L.orb_version.restype = C.c_char_p
        _lib = L
    return _lib


def check(rc):
    if rc != ORB_OK:
        raise OrbError(rc, lib().orb_last_error().decode())


def ptr(a):
    """Raw pointer of a numpy array or torch tensor (or an int address / None)."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())  # torch tensor


def kernel_launch_count():
    return int(lib().orb_kernel_launch_count())
