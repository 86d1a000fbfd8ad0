"""ctypes mirror of include/epi_b200.h (the C ABI of libepi_b200.so).

This is the binding the tests, the benchmark and the MATLAB-signature mirror
(api.py) call.  There is no CPU fallback: a missing library or a missing CUDA
device raises.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EPI_B200_LIB") or os.path.join(PKG, "libepi_b200.so")  # env: kernel-variant experiments

LMAX = 12
MEM_HOST, MEM_DEVICE = 0, 1
OK, ERR_ARG, ERR_ORDER, ERR_OBS_TYPE, ERR_QR_SHAPE = 0, -1, -2, -3, -4
ERR_CUDA, ERR_NO_DEVICE, ERR_NOMEM = -10, -11, -12
(MODEL_SIALPHA, MODEL_SIALPHA_FLIPPED, MODEL_OPTCTRL, MODEL_OPTCTRL_FLIPPED, MODEL_LEGACY_TOOLS,
 MODEL_LEGACY_CODEGEN) = range(6)
OBS_NEWCASES, OBS_TOTALCASES = 0, 1
Q_CONST, Q_PERDAY_SCALAR, Q_PERDAY_FULL = 0, 1, 2
R_CONST, R_PERDAY = 0, 1
RATES_CONST, RATES_SHARED_SERIES, RATES_SERIES = 0, 1, 2
SEIRP_OUT_FULL, SEIRP_OUT_FINAL = 0, 1
U_F64, U_U8, U_PHILOX = 0, 1, 3

_dp = C.c_void_p  # every array pointer is passed as a raw address (host or device)


class ModelParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("beta", C.c_double), ("gamma", C.c_double), ("b", C.c_double),
                ("alpha_min", C.c_double), ("alpha_max", C.c_double), ("s_min", C.c_double),
                ("i_min", C.c_double), ("epsilon", C.c_double), ("sigma", C.c_double),
                ("a", C.c_double * LMAX), ("u_min", C.c_double * LMAX),
                ("u_max", C.c_double * LMAX), ("w", C.c_double * LMAX),
                ("L", C.c_int), ("obs_type", C.c_int)]


class SeirpArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("B", C.c_int), ("K", C.c_int), ("dt", C.c_double),
                ("rate_mode", C.c_int), ("rates", _dp), ("ic", _dp), ("saturated", C.c_int),
                ("beta_0", C.c_double), ("beta_s", C.c_double), ("mu_0", C.c_double),
                ("mu_s", C.c_double), ("sigma", C.c_double), ("i_0", C.c_double),
                ("out_mode", C.c_int), ("out", _dp)]


class RolloutArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("B", C.c_int), ("K", C.c_int), ("L", C.c_int), ("G", C.c_int),
                ("prm", _dp), ("x0", _dp), ("noise_std", _dp), ("u_kind", C.c_int), ("u", _dp),
                ("noise", _dp), ("s", _dp), ("i", _dp), ("alpha", _dp), ("T_total", C.c_int),
                ("j0_prefix", _dp), ("j1_prefix", _dp), ("w", _dp), ("J0", _dp), ("J1", _dp),
                ("seed", C.c_ulonglong), ("first", C.c_longlong)]


class SchedulesArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("B", C.c_int), ("K", C.c_int), ("L", C.c_int), ("G", C.c_int),
                ("prm", _dp), ("seed", C.c_ulonglong), ("first", C.c_longlong), ("u", _dp)]


class RtExpFitArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("B", C.c_int), ("T", C.c_int), ("G", C.c_int), ("x", _dp), ("s_init", _dp),
                ("params", _dp), ("w_bar", _dp), ("Ps_init", _dp), ("Q", _dp), ("R", _dp),
                ("v_bar", C.c_double), ("beta", C.c_double), ("gamma", C.c_double), ("W", C.c_int),
                ("order", C.c_int), ("S_MINUS", _dp), ("S_PLUS", _dp), ("P_MINUS", _dp), ("P_PLUS", _dp),
                ("K_GAIN", _dp), ("S_SMOOTH", _dp), ("P_SMOOTH", _dp), ("innovations", _dp), ("rho", _dp)]


class PreprocessArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("B", C.c_int), ("T", C.c_int), ("L", C.c_int), ("W", C.c_int),
                ("n_first", C.c_int), ("min_cases", C.c_double), ("cc", _dp), ("population", _dp), ("ip", _dp),
                ("ip_filled", _dp), ("refined", _dp), ("smoothed", _dp), ("zerolag", _dp), ("normalized", _dp),
                ("confirmed_norm", _dp), ("R_v", _dp), ("I0", _dp)]


class NnlsArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("B", C.c_int), ("n", C.c_int), ("p", C.c_int), ("max_alt", C.c_int),
                ("X", _dp), ("y", _dp), ("a", _dp), ("b", _dp), ("n_alt", _dp)]


class NpiCostArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("B", C.c_int), ("T", C.c_int), ("L", C.c_int), ("G", C.c_int),
                ("newcases", _dp), ("inputs", _dp), ("weights", _dp), ("J0", _dp), ("J1", _dp)]


class SiArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("B", C.c_int), ("K", C.c_int), ("dt", C.c_double),
                ("alpha", _dp), ("beta", _dp), ("s0", _dp), ("i0", _dp), ("s", _dp), ("i", _dp)]


class EkfArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("model", C.c_int), ("B", C.c_int), ("T", C.c_int), ("L", C.c_int),
                ("G", C.c_int), ("prm", _dp), ("epsilon", _dp), ("u_per_traj", C.c_int), ("u", _dp),
                ("x_per_traj", C.c_int), ("x", _dp), ("r_mode", C.c_int), ("fixed_R", C.c_int),
                ("r_per_traj", C.c_int), ("R", _dp), ("q_mode", C.c_int), ("Q", _dp),
                ("init_per_traj", C.c_int), ("s_init", _dp), ("Ps_init", _dp), ("s_final", _dp),
                ("Ps_final", _dp), ("v_bar", C.c_double), ("beta", C.c_double), ("gamma", C.c_double),
                ("W", C.c_int), ("order", C.c_int), ("u_opt", _dp), ("u_opt_smooth", _dp),
                ("S_MINUS", _dp), ("S_PLUS", _dp), ("S_SMOOTH", _dp), ("P_MINUS", _dp),
                ("P_PLUS", _dp), ("P_SMOOTH", _dp), ("K_GAIN", _dp), ("innovations", _dp),
                ("rho", _dp), ("status", _dp)]


class ParetoArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("n_sets", C.c_int), ("n", C.c_int), ("J0", _dp), ("J1", _dp),
                ("on_front", _dp), ("I_opt", _dp)]


class SweepArgs(C.Structure):
    _fields_ = [("mem", C.c_int), ("n_regions", C.c_int), ("n_eps", C.c_int), ("T", C.c_int),
                ("T_hist", C.c_int), ("L", C.c_int), ("prm", _dp), ("eps", _dp), ("u", _dp),
                ("x", _dp), ("R", _dp), ("s_init", _dp), ("Ps_init", _dp), ("s_final", _dp),
                ("Ps_final", _dp), ("Q", _dp), ("beta_ekf", C.c_double), ("gamma_ekf", C.c_double),
                ("W", C.c_int), ("x0", _dp), ("newcases_hist", _dp), ("weights", _dp),
                ("noise_std", _dp), ("noise", _dp), ("J0", _dp), ("J1", _dp), ("on_front", _dp),
                ("I_opt", _dp), ("u_knee", _dp), ("u_fore", _dp), ("P_first", _dp), ("lean", C.c_int)]


# every symbol include/epi_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "epi_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "epi_destroy": (None, [C.c_void_p]),
    "epi_last_error": (C.c_char_p, [C.c_void_p]),
    "epi_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "epi_sync": (C.c_int, [C.c_void_p]),
    "epi_set_scratch_limit": (C.c_int, [C.c_void_p, C.c_size_t]),
    "epi_release_cache": (C.c_int, [C.c_void_p]),
    "epi_launch_count": (C.c_longlong, [C.c_void_p]),
    "epi_last_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.c_int]),
    "epi_fp64_probe": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "epi_seirp_batch": (C.c_int, [C.c_void_p, C.POINTER(SeirpArgs)]),
    "epi_rollout_cost_batch": (C.c_int, [C.c_void_p, C.POINTER(RolloutArgs)]),
    "epi_random_schedules": (C.c_int, [C.c_void_p, C.POINTER(SchedulesArgs)]),
    "epi_rt_expfit_batch": (C.c_int, [C.c_void_p, C.POINTER(RtExpFitArgs)]),
    "epi_preprocess_batch": (C.c_int, [C.c_void_p, C.POINTER(PreprocessArgs)]),
    "epi_nnls_affine_batch": (C.c_int, [C.c_void_p, C.POINTER(NnlsArgs)]),
    "epi_npicost_batch": (C.c_int, [C.c_void_p, C.POINTER(NpiCostArgs)]),
    "epi_si_controlled_batch": (C.c_int, [C.c_void_p, C.POINTER(SiArgs)]),
    "epi_ekf_eks_batch": (C.c_int, [C.c_void_p, C.POINTER(EkfArgs)]),
    "epi_pareto_batch": (C.c_int, [C.c_void_p, C.POINTER(ParetoArgs)]),
    "epi_sweep": (C.c_int, [C.c_void_p, C.POINTER(SweepArgs)]),
    "epi_sweep_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(SweepArgs)]),
}

_lib = None


class EpiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libepi_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


def load():
    """dlopen libepi_b200.so and bind every declared symbol; raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m epidemicmodeling_b200._build` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
