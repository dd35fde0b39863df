"""ctypes binding of libsaigegds_b200.so (the C-ABI declared in include/saigegds_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present,
every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsaigegds_b200.so")
CSRC = os.path.join(HERE, "csrc")

SGB_OK, SGB_ERR_INVALID, SGB_ERR_CUDA, SGB_ERR_OVERFLOW, SGB_ERR_COMM, SGB_ERR_STATE = range(6)
FAMILY = {"binomial": 0, "gaussian": 1}
KERNEL = {"auto": 0, "simt": 1, "imma": 2, "imma2": 3, "umma": 4}


class SgbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class InvalidArgument(SgbError, ValueError):
    """std::invalid_argument in the reference."""


class OverflowErrorSGB(SgbError, OverflowError):
    """std::overflow_error in the reference ('Large variance estimate ...', 'Sigma_E = 0 ...')."""


class Param(C.Structure):
    _fields_ = [("tol", C.c_double), ("tolPCG", C.c_double), ("seed", C.c_int), ("maxiter", C.c_int),
                ("maxiterPCG", C.c_int), ("no_iteration", C.c_int), ("nrun", C.c_int), ("num_marker", C.c_int),
                ("traceCVcutoff", C.c_double), ("ratioCVcutoff", C.c_double), ("verbose", C.c_int),
                ("indent", C.c_char_p)]


class Fit0(C.Structure):
    _fields_ = [("n", C.c_int64), ("p", C.c_int), ("y", C.POINTER(C.c_double)), ("offset", C.POINTER(C.c_double)),
                ("linear_predictors", C.POINTER(C.c_double)), ("fitted_values", C.POINTER(C.c_double)),
                ("coefficients", C.POINTER(C.c_double)), ("family", C.c_int)]


class Glmm(C.Structure):
    _fields_ = [("coefficients", C.POINTER(C.c_double)), ("tau", C.c_double * 2),
                ("linear_predictors", C.POINTER(C.c_double)), ("fitted_values", C.POINTER(C.c_double)),
                ("residuals", C.POINTER(C.c_double)), ("cov", C.POINTER(C.c_double)), ("converged", C.c_int)]


class NoK(C.Structure):
    _fields_ = [("p", C.c_int), ("X1", C.POINTER(C.c_double)), ("XV", C.POINTER(C.c_double)),
                ("XXVX_inv", C.POINTER(C.c_double))]


class VarRatio(C.Structure):
    _fields_ = [("capacity", C.c_int), ("n", C.c_int), ("id", C.POINTER(C.c_int)), ("maf", C.POINTER(C.c_double)),
                ("mac", C.POINTER(C.c_double)), ("var1", C.POINTER(C.c_double)), ("var2", C.POINTER(C.c_double)),
                ("ratio", C.POINTER(C.c_double))]


class ScoreModel(C.Structure):
    _fields_ = [("trait", C.c_int), ("n", C.c_int64), ("K", C.c_int), ("tau", C.POINTER(C.c_double)),
                ("y", C.POINTER(C.c_double)), ("mu", C.POINTER(C.c_double)), ("y_mu", C.POINTER(C.c_double)),
                ("mu2", C.POINTER(C.c_double)), ("t_XXVX_inv", C.POINTER(C.c_double)), ("XV", C.POINTER(C.c_double)),
                ("t_XVX_inv_XV", C.POINTER(C.c_double)), ("t_X", C.POINTER(C.c_double)), ("XVX", C.POINTER(C.c_double)),
                ("S_a", C.POINTER(C.c_double)), ("var_ratio", C.c_double)]


class GxG(C.Structure):
    _fields_ = [("beta", C.c_double), ("SE", C.c_double), ("pval", C.c_double), ("p_norm", C.c_double),
                ("tau_G", C.c_double), ("n_nonzero", C.c_int64), ("converged", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("n_products", C.c_int64), ("n_product_launches", C.c_int64), ("n_kernel_launches", C.c_int64),
                ("n_pcg_solves", C.c_int64), ("n_pcg_iterations", C.c_int64), ("last_product_ms", C.c_double),
                ("n_host_syncs", C.c_int64), ("host_wait_s", C.c_double)]


# every symbol declared in include/saigegds_b200.h
SYMBOLS = [
    "sgb_ctx_create", "sgb_ctx_destroy", "sgb_last_error", "sgb_set_callbacks", "sgb_set_kernel",
    "sgb_comm_unique_id", "sgb_comm_init", "sgb_store_2b_geno", "sgb_store_2b_geno_device", "sgb_allele_counts",
    "sgb_get_geno_ds", "sgb_grm_mv", "sgb_grm_mv_device", "sgb_diag_sigma", "sgb_pcg", "sgb_fit_AI_PCG_binary",
    "sgb_fit_AI_PCG_quant", "sgb_calc_var_ratio_binary", "sgb_calc_var_ratio_quant", "sgb_r_set_seed",
    "sgb_r_unif_rand", "sgb_r_sample_int", "sgb_get_stats", "sgb_reset_stats", "sgb_synth_geno_device",
    "sgb_copy_from_device", "sgb_free_device", "sgb_time_products_device", "sgb_malloc_device", "sgb_copy_to_device",
    "sgb_set_profiling", "sgb_kernel_times", "sgb_malloc_host", "sgb_free_host",
    "sgb_get_sparse", "sgb_store_sp_geno", "sgb_sparse_to_packed",
    "sgb_score_test_init", "sgb_score_test_packed", "sgb_score_test_dosage", "sgb_score_test_stored", "sgb_score_test_set_path", "sgb_GxG_snp_bin", "sgb_store_gds_geno",
]


def build(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"]
    if not verbose:
        cmd.append("-s")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.sgb_last_error.restype = C.c_char_p
        for s in SYMBOLS:
            getattr(_lib, s)  # AttributeError if a declared symbol is not exported
    return _lib


def check(rc):
    if rc == SGB_OK:
        return
    msg = lib().sgb_last_error().decode("utf-8", "replace")
    if rc == SGB_ERR_INVALID:
        raise InvalidArgument(rc, msg)
    if rc == SGB_ERR_OVERFLOW:
        raise OverflowErrorSGB(rc, msg)
    raise SgbError(rc, msg)
