"""cooper-mapper-b200: B200-native LOAM hot path (scan registration + scan-to-map registration).

Python is only the thin host mirror of the reference's operator interface above the C ABI
(include/coopermap.h -> libcoopermap.so, hand-written sm_100a CUDA).  There is NO CPU fallback: if the
library is missing or no CUDA device is visible, everything here raises.

Import with importlib (the directory name carries a hyphen):
    cmb = importlib.import_module("the-cooper-mapper_b200")
"""
from .api import (CM_OK, CM_TOO_FEW_REF, CM_TOO_FEW_MATCHES, CM_NOT_CONVERGED, CM_LOW_SCORE, Config, Context,  # noqa: F401
                  CoopermapError, LaserLocalization, LaserMapping, LaserMappingLocal, LaserOdometry, MatchStats, OdomStats, ScanMatch, lib_path, load_library)
