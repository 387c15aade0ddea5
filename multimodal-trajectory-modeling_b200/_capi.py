"""ctypes binding of include/mtm_b200.h — the only way the Python host reaches the CUDA code.

There is no CPU fallback: if the shared library is missing or no B200 is present the calls
raise (``MtmError``), they never route anywhere else.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# MTM_B200_LIB: alternative build of the same library (kernel-tuning experiments)
LIB_PATH = os.environ.get("MTM_B200_LIB") or os.path.join(HERE, "libmtm_b200.so")

MTM_MAX_DIM = 8
MTM_MAX_CLUSTERS = 64

STATUS_NONFINITE = 1
STATUS_SINGULAR = 2
STATUS_EMPTY_FIT = 4
STATUS_CHOL_FAIL = 8
STATUS_COMM_TIMEOUT = 16

EXIT_RUNNING, EXIT_CONVERGED, EXIT_GUARD, EXIT_EXHAUSTED, EXIT_GUARD0, EXIT_ERROR, EXIT_EMPTY = 0, 1, 2, 3, 4, 5, 6
EXIT_NAMES = {0: "running", 1: "converged", 2: "guard", 3: "exhausted", 4: "guard0", 5: "error", 6: "empty"}


class MtmError(RuntimeError):
    pass


class MtmData(ctypes.Structure):
    _fields_ = [("z_dev", ctypes.c_void_p), ("x_dev", ctypes.c_void_p), ("n", ctypes.c_int64),
                ("T", ctypes.c_int32), ("d", ctypes.c_int32), ("l", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("ld", ctypes.c_int64)]


class EStepInfo(ctypes.Structure):
    _fields_ = [("n_switches", ctypes.c_int64), ("min_count", ctypes.c_int64), ("status", ctypes.c_int32),
                ("stats_mode", ctypes.c_int32)]


class KMeansInfo(ctypes.Structure):
    _fields_ = [("n_iter", ctypes.c_int32), ("converged", ctypes.c_int32), ("empty_cluster", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("tol", ctypes.c_double)]


class FitInfo(ctypes.Structure):
    _fields_ = [("exit_code", ctypes.c_int32), ("n_e_steps", ctypes.c_int32), ("status", ctypes.c_int32),
                ("n_m_steps", ctypes.c_int32), ("last_switches", ctypes.c_int64), ("min_count", ctypes.c_int64)]


# name -> (restype, argtypes); must list every symbol include/mtm_b200.h declares
_vp, _i, _i64, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
SIGNATURES = {
    "mtm_abi_version": (_i, []),
    "mtm_create": (_i, [_i, _vp, ctypes.POINTER(_vp)]),
    "mtm_destroy": (_i, [_vp]),
    "mtm_set_stream": (_i, [_vp, _vp]),
    "mtm_last_error": (ctypes.c_char_p, [_vp]),
    "mtm_param_stride": (_i, [_i, _i]),
    "mtm_stats_stride": (_i, [_i, _i]),
    "mtm_set_data": (_i, [_vp, ctypes.POINTER(MtmData)]),
    "mtm_estep": (_i, [_vp, ctypes.POINTER(MtmData), _vp, _i, _i, _vp, _i, _vp, _vp, _vp, ctypes.POINTER(EStepInfo)]),
    "mtm_estep_prefix": (_i, [_vp, ctypes.POINTER(MtmData), _vp, _i, _i, _vp, ctypes.POINTER(ctypes.c_int32)]),
    "mtm_score": (_i, [_vp, ctypes.POINTER(MtmData), _vp, _i, ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_int32)]),
    "mtm_mstats": (_i, [_vp, _vp, _i, _vp]),
    "mtm_finalize": (_i, [_vp, _vp, _i, _i, _i, _i64, _d, _vp, _vp]),
    "mtm_mstep": (_i, [_vp, _vp, _i, _d, _vp, ctypes.POINTER(EStepInfo)]),
    "mtm_em_step": (_i, [_vp, _i, _d, _vp, _vp, _i, ctypes.POINTER(EStepInfo)]),
    "mtm_em_step_host": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _i, _d, _vp, _vp, _vp, _vp, _i,
                              ctypes.POINTER(EStepInfo)]),
    "mtm_em_fit": (_i, [_vp, _i, _d, _i, _vp, _vp, ctypes.POINTER(FitInfo), ctypes.POINTER(_i64), _i]),
    "mtm_mstep_ext": (_i, [_vp, _vp, _i, _d, _vp, ctypes.POINTER(EStepInfo)]),
    "mtm_em_fit_ext": (_i, [_vp, _i, _d, _i, _vp, _vp, _vp, ctypes.POINTER(FitInfo), ctypes.POINTER(_i64), _i]),
    "mtm_sweep_fit": (_i, [_vp, _i, _i, _d, _i, _vp, _vp, ctypes.POINTER(FitInfo), ctypes.POINTER(_d),
                           ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(_i64), _i]),
    "mtm_comm_unique_id": (_i, [_vp]),
    "mtm_comm_init": (_i, [_vp, _vp, _i, _i]),
    "mtm_comm_attach": (_i, [_vp, _vp, _i, _i]),
    "mtm_comm_destroy": (_i, [_vp]),
    "mtm_p2p_export": (_i, [_vp, _vp]),
    "mtm_p2p_attach": (_i, [_vp, _vp, _i, _i]),
    "mtm_random_labels": (_i, [_vp, ctypes.POINTER(ctypes.c_uint64), _i, _i64, _vp, ctypes.POINTER(_i)]),
    "mtm_pattern_order": (_i, [_vp, ctypes.POINTER(MtmData), _vp]),
    "mtm_release_workspace": (_i, [_vp]),
    "mtm_set_delta_stats": (_i, [_vp, _i]),
    "mtm_permute_f64": (_i, [_vp, _vp, _vp, _i64, _i64, _i, _vp, _i]),
    "mtm_permute_i32": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _i]),
    "mtm_kmeans_init": (_i, [_vp, ctypes.POINTER(MtmData), _vp, _i, _i64, ctypes.POINTER(_d), _i, _i, _d, _vp,
                             ctypes.POINTER(KMeansInfo)]),
    "mtm_fingerprint": (_i, [_vp, ctypes.POINTER(MtmData), _vp, ctypes.POINTER(ctypes.c_uint64)]),
    "mtm_fp64_peak": (_i, [_vp, _i, ctypes.POINTER(_d)]),
    "mtm_launch_count": (_i64, [_vp]),
}

_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libmtm_b200.so (in-tree)."""
    csrc = os.path.join(HERE, "csrc")
    cmd = ["make", "-C", csrc, "-j8"] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise MtmError("building libmtm_b200.so failed")
    return LIB_PATH


def load():
    """dlopen libmtm_b200.so and set the prototypes.  Fails loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MtmError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.mtm_abi_version() != 1:
        raise MtmError("ABI version mismatch between _capi.py and libmtm_b200.so")
    _lib = lib
    return lib


def check(rc: int, handle=None):
    if rc != 0:
        msg = load().mtm_last_error(handle)
        raise MtmError(f"mtm_b200 call failed ({rc}): {msg.decode() if msg else '?'}")
