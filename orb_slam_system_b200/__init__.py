"""B200-native ORB front end: ORBextractor / ORBmatcher hot paths of
WangHewei16/ORB-SLAM-System behind a C ABI (include/orb_b200.h).

Importing the package does not load CUDA; the first call into the library does and
raises if liborb_b200.so has not been built or no GPU is usable (no CPU fallback).
"""
from ._lib import KP_DTYPE, LIB_PATH, OrbError, build, kernel_launch_count, lib  # noqa: F401
from .extractor import ORBextractor  # noqa: F401
from .matcher import FeatureVector, FrameView, ORBmatcher  # noqa: F401
from .vocabulary import ORBVocabulary  # noqa: F401
