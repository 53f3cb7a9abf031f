"""ctypes binding of libtrmps_b200.so (include/trmps_b200.h).

This is the only place the Python mirror touches native code.  There is no CPU or
TensorFlow fallback: if the shared library is missing, or no sm_100 GPU is present
when a compute entry point is called, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "libtrmps_b200.so"))

OK = 0
FM_PRECOMPUTED, FM_ONE_SQRT3, FM_ONE_X, FM_COS_SIN = 0, 1, 2, 3
LOSS_SOFTMAX_XENT, LOSS_SQUARED = 0, 1
DT_F32, DT_U8, DT_U8_DIV = 0, 1, 2
POOL_NONE, POOL_AVG, POOL_MAX = 0, 1, 2
S_COST0, S_GDC, S_LR_STAR, S_COST1, S_LR, S_COUNT = 0, 1, 2, 3, 4, 32
I_ACCEPT, I_TRIALS, I_M, I_SVD_SWEEPS, I_N_REVERT, I_N_EXHAUST, I_N_UPDATES, I_N_SWEEPS = 0, 1, 2, 3, 4, 5, 6, 7
I_COUNT = 96

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


class Opt(C.Structure):
    _fields_ = [("lr", C.c_float), ("reg", C.c_float), ("armijo_coeff", C.c_float),
                ("min_singular_value", C.c_float), ("max_size", C.c_int32), ("min_size", C.c_int32),
                ("updates_per_step", C.c_int32), ("use_hessian", C.c_int32),
                ("max_svd_sweeps", C.c_int32), ("single_site", C.c_int32),
                ("split_impl", C.c_int32), ("reserved", C.c_int32)]


class Layout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("bondM", "gradM", "deltaM", "hessM", "gpart", "f0", "fd", "fpart", "resid", "sres", "hres",
                 "phi2", "Ut", "sv", "red", "scal", "bimg", "split_ws", "env_img", "total")] + \
               [(n, C.c_int32) for n in ("gsplit", "nsplit", "red_rows", "svd_block", "gsplit_tc", "reserved0")]


class Sweep(C.Structure):
    _fields_ = [("N", C.c_int32), ("d", C.c_int32), ("L", C.c_int32), ("Ds", C.c_int32),
                ("B", C.c_int64), ("B_total", C.c_int64),
                ("feature_map", C.c_int32), ("loss_kind", C.c_int32),
                ("feat", C.c_void_p), ("labels", C.c_void_p), ("nodes", C.c_void_p), ("special", C.c_void_p),
                ("dims", C.c_void_p), ("envL", C.c_void_p), ("envR", C.c_void_p), ("work", C.c_void_p),
                ("iwork", C.c_void_p), ("comm", C.c_void_p), ("work_floats", C.c_int64)]


class PredictDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("d", C.c_int32), ("L", C.c_int32), ("Ds", C.c_int32),
                ("B", C.c_int64), ("loc", C.c_int32), ("feature_map", C.c_int32),
                ("round_output", C.c_int32), ("reserved", C.c_int32),
                ("feat", C.c_void_p), ("nodes", C.c_void_p), ("special", C.c_void_p), ("dims", C.c_void_p),
                ("logits", C.c_void_p), ("work", C.c_void_p)]


_SIGNATURES = {
    "trmps_abi_version": (C.c_int, []),
    "trmps_last_error": (C.c_char_p, []),
    "trmps_device_check": (C.c_int, [C.c_int]),
    "trmps_set_option": (C.c_int, [C.c_int32, C.c_int32]),
    "trmps_launch_count": (C.c_int64, []),
    "trmps_debug_chain": (C.c_int, [C.POINTER(C.c_float)]),
    "trmps_profile_begin": (C.c_int, [C.c_int32]),
    "trmps_profile_end": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int32]),
    "trmps_sweep_layout": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Layout)]),
    "trmps_predict_work_floats": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "trmps_feature_map": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "trmps_preprocess_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                         C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "trmps_batch_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                     C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "trmps_linreg_pretrain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                        C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "trmps_env_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "trmps_predict": (C.c_int, [C.POINTER(PredictDesc), C.c_void_p]),
    "trmps_sweep_init_env": (C.c_int, [C.POINTER(Sweep), C.c_int32, C.c_void_p]),
    "trmps_bond_step": (C.c_int, [C.POINTER(Sweep), C.c_int32, C.c_int32, C.POINTER(Opt), C.c_void_p]),
    "trmps_sweep_run": (C.c_int, [C.POINTER(Sweep), C.c_int32, C.c_int32, C.c_int32, C.POINTER(Opt), C.c_void_p]),
    "trmps_train_step": (C.c_int, [C.POINTER(Sweep), C.c_int32, C.POINTER(Opt), C.c_void_p]),
    "trmps_bond_merge": (C.c_int, [C.POINTER(Sweep), C.c_int32, C.c_int32, C.c_void_p]),
    "trmps_bond_forward": (C.c_int, [C.POINTER(Sweep), C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "trmps_bond_gradient": (C.c_int, [C.POINTER(Sweep), C.c_int32, C.c_int32, C.POINTER(Opt), C.c_void_p]),
    "trmps_bond_split": (C.c_int, [C.POINTER(Sweep), C.c_int32, C.c_int32, C.POINTER(Opt), C.c_void_p]),
    "trmps_comm_unique_id": (C.c_int, [C.c_void_p]),
    "trmps_comm_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "trmps_comm_destroy": (C.c_int, [C.c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError("libtrmps_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; "
                               "g.build()'` (or make -C mps-mnist_b200/csrc); there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.trmps_abi_version() != 1:
            raise RuntimeError("libtrmps_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def last_error():
    return load().trmps_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != OK:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, last_error()))


def require_gpu(device=0):
    """Fail loudly when the product path cannot run (no sm_100 device)."""
    check(load().trmps_device_check(int(device)), "trmps_device_check")


PROF_CATEGORIES = ("env", "forward", "gradient", "svd", "merge", "scalar", "allreduce")


def profile_begin(max_records=65536):
    check(load().trmps_profile_begin(max_records), "trmps_profile_begin")


def profile_end():
    """-> {category: (milliseconds, number of timed regions)}"""
    n = len(PROF_CATEGORIES)
    ms = (C.c_float * n)()
    cnt = (C.c_int32 * n)()
    check(load().trmps_profile_end(ms, cnt, n), "trmps_profile_end")
    return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(PROF_CATEGORIES)}


OPT_GRAD_IMPL, GRAD_TCGEN05, GRAD_SIMT, GRAD_TCGEN05_F16 = 0, 0, 1, 2      # 0: 3xTF32 on tcgen05, 2: fp16 hi/lo on tcgen05 (default)
GRAD_DEFAULT = GRAD_TCGEN05_F16
OPT_FWD_IMPL, FWD_TCGEN05, FWD_SIMT, FWD_TCGEN05_TF32 = 1, 0, 1, 2      # 0: fp16 hi/lo on tcgen05 (default), 2: 3xTF32 on tcgen05
OPT_CHAIN_IMPL, CHAIN_TCGEN05, CHAIN_SIMT, CHAIN_TCGEN05_ALWAYS = 2, 0, 1, 2     # 0 falls back to 1 for chains whose rounding-bias estimate exceeds 9e-5


OPT_ENV_IMPL, ENV_TCGEN05, ENV_SIMT = 4, 0, 1       # environments of the sweep: tensor-core chain (while its rounding-bias estimate allows) / FP32 SIMT steps


def set_option(option, value):
    check(load().trmps_set_option(option, value), "trmps_set_option")


def launch_count():
    return int(load().trmps_launch_count())


def sweep_layout(N, B, L, Ds, sm_count=148):
    lay = Layout()
    check(load().trmps_sweep_layout(N, B, L, Ds, sm_count, C.byref(lay)), "trmps_sweep_layout")
    return lay
