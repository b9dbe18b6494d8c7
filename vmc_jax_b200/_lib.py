"""ctypes binding of libjvmc_b200.so (the C ABI declared in include/jvmc_b200.h).

There is no CPU fallback: if the library is missing, or no CUDA device is present when a kernel is
requested, the call raises."""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JVMC_B200_LIB") or os.path.join(_HERE, "libjvmc_b200.so")   # override: kernel variants (tools/)
HEADER = os.path.join(os.path.dirname(_HERE), "include", "jvmc_b200.h")

_lib = None

c_ll = ctypes.c_longlong
c_ull = ctypes.c_ulonglong
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_ptr = ctypes.c_void_p

_SIGS = {
    "jvmc_version": (c_int, []),
    "jvmc_error_string": (ctypes.c_char_p, [c_int]),
    "jvmc_rbm_logpsi": (c_int, [c_ptr, c_ll, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "jvmc_rbm_tables_elems": (c_ll, [c_int, c_int]),
    "jvmc_rbm_tables": (c_int, [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "jvmc_rbm_mcmc": (c_int, [c_ptr, c_ll, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ull, c_ull, c_ll, c_int, c_dbl,
                              c_int, c_ll, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "jvmc_bfo_matels": (c_int, [c_ptr, c_ll, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int,
                                c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "jvmc_bfo_emit": (c_int, [c_ptr, c_ll, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                              c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr]),
    "jvmc_oloc_reduce": (c_int, [c_ptr, c_ptr, c_ptr, c_ll, c_int, c_ptr, c_ptr]),
    "jvmc_rbm_eloc_bfo": (c_int, [c_ptr, c_ptr, c_ll, c_int, c_int, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr,
                                  c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "jvmc_rbm_grad": (c_int, [c_ptr, c_ptr, c_ll, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "jvmc_rbm_moments_chunks": (c_int, [c_ll]),
    "jvmc_rbm_moments": (c_int, [c_ptr, c_ptr, c_ptr, c_ll, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "jvmc_rbm_krmatvec": (c_int, [c_ptr, c_ptr, c_ptr, c_ll, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "jvmc_i8_layout": (c_int, [c_ll, c_int, ctypes.POINTER(c_ll), ctypes.POINTER(c_int), ctypes.POINTER(c_ll)]),
    "jvmc_tdvp_solve_workspace": (c_int, [c_int, c_int, c_ll, ctypes.POINTER(c_ll)]),
    "jvmc_tdvp_solve": (c_int, [c_int, c_int, c_ptr, c_ptr, c_ll, c_ptr, c_ptr, c_ptr, c_dbl, c_dbl, c_dbl, c_int, c_dbl,
                                c_dbl, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ll, c_ptr]),
    "jvmc_minsr_solve_workspace": (c_int, [c_int, c_int, ctypes.POINTER(c_ll)]),
    "jvmc_minsr_solve": (c_int, [c_int, c_int, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ll, c_ptr]),
    "jvmc_hetrd_workspace": (c_int, [c_int, c_int, ctypes.POINTER(c_ll)]),
    "jvmc_hetrd": (c_int, [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ll, c_ptr, c_ptr]),
    "jvmc_unmtr_workspace": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_ll)]),
    "jvmc_unmtr": (c_int, [c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ll, c_ptr, c_ptr]),
    "jvmc_unmqr_workspace": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_ll)]),
    "jvmc_unmqr": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_int, c_ptr, c_ptr, c_int, c_ptr, c_ll, c_ptr, c_ptr]),
    "jvmc_tridiag_dense": (c_int, [c_int, c_ptr, c_ptr, c_dbl, c_dbl, c_ptr, c_ptr]),
    "jvmc_real_to_complex": (c_int, [c_ll, c_ptr, c_ptr, c_ptr]),
    "jvmc_secular_roots": (c_int, [c_int, c_ptr, c_ptr, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr]),
    "jvmc_secular_vectors": (c_int, [c_int, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ll, c_ptr, c_ptr]),
    "jvmc_apply_row_rotations": (c_int, [c_int, c_ll, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "jvmc_comm_nccl_version": (c_int, []),
    "jvmc_comm_unique_id": (c_int, [c_ptr]),
    "jvmc_comm_init": (c_int, [c_ptr, c_int, c_int, ctypes.POINTER(c_ptr)]),
    "jvmc_comm_destroy": (c_int, [c_ptr]),
    "jvmc_comm_allreduce_sum_f64": (c_int, [c_ptr, c_ptr, c_ll, c_ptr]),
    "jvmc_comm_allgather_bytes": (c_int, [c_ptr, c_ptr, c_ptr, c_ll, c_ptr]),
    "jvmc_comm_bcast_bytes": (c_int, [c_ptr, c_ptr, c_ll, c_int, c_ptr]),
    "jvmc_comm_reduce_scatter_sum_f64": (c_int, [c_ptr, c_ptr, c_ptr, c_ll, c_ptr]),
    "jvmc_i8_set_debug": (c_int, [c_int]),
    "jvmc_i8_set_trace": (c_int, [c_ptr, c_int]),
    "jvmc_hermitian_mirror_blocks": (c_int, [c_ptr, c_int, c_int, c_ptr]),
    "jvmc_hermitian_packed_elems": (c_ll, [c_int, c_int]),
    "jvmc_hermitian_pack_blocks": (c_int, [c_ptr, c_int, c_int, c_ptr, c_int, c_ptr]),
    "jvmc_hermitian_pack_rows": (c_int, [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_int, c_ptr]),
    "jvmc_i8_tile_shape": (c_int, [ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "jvmc_mcmc_set_generic": (c_int, [c_int]),
    "jvmc_cnn_num_parameters": (c_int, [c_ptr, c_int, ctypes.POINTER(c_int)]),
    "jvmc_cnn_logpsi": (c_int, [c_ptr, c_int, c_ptr, c_ptr, c_ll, c_ptr, c_ptr]),
    "jvmc_cnn_grad": (c_int, [c_ptr, c_int, c_ptr, c_ptr, c_ll, c_ptr, c_ptr]),
    "jvmc_cnn_mcmc": (c_int, [c_ptr, c_int, c_ptr, c_ptr, c_ll, c_ull, c_ull, c_ll, c_int, c_dbl, c_int, c_ll, c_int,
                              c_ptr, c_ptr, c_ptr]),
    "jvmc_cnn_mcmc_inc": (c_int, [c_ptr, c_int, c_ptr, c_ptr, c_ll, c_ull, c_ull, c_ll, c_int, c_dbl, c_int, c_ll, c_int,
                              c_ptr, c_ptr, c_ptr]),
    "jvmc_cnn_eloc_bfo": (c_int, [c_ptr, c_int, c_ptr, c_ptr, c_ll, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int,
                                  c_ptr, c_ptr, c_ptr, c_ptr]),
    "jvmc_cnn_set_generic": (c_int, [c_int]),
    "jvmc_symrbm_logpsi": (c_int, [c_ptr, c_ll, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "jvmc_symrbm_grad": (c_int, [c_ptr, c_ll, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr,
                                 c_ptr]),
    "jvmc_symrbm_mcmc": (c_int, [c_ptr, c_ll, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ull,
                                 c_ull, c_ll, c_dbl, c_int, c_ll, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "jvmc_i8_slice": (c_int, [c_ptr, c_ll, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "jvmc_i8_tail_ratios": (c_int, [c_ptr, c_ll, c_int, c_ptr, c_ptr, c_ptr]),
    "jvmc_i8_outlier_rows": (c_int, [c_ptr, c_ll, c_int, c_ptr, c_dbl, c_ptr, c_ptr]),
    "jvmc_rbm_gram_S_i8": (c_int, [c_ptr, c_ptr, c_ll, c_int, c_int, c_ptr, c_ptr, c_int, c_ptr, c_dbl, c_dbl, c_int, c_ll, c_ll,
                                   c_ptr, c_ptr]),
    "jvmc_pack_sigma_rows": (c_int, [c_ptr, c_ll, c_int, c_int, c_ptr, c_ptr]),
    "jvmc_rbm_gram_T_tiles": (c_ll, [c_ll]),
    "jvmc_rbm_gram_T": (c_int, [c_ptr, c_ll, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ll, c_ll, c_ptr, c_ptr]),
    "jvmc_pack_sigma": (c_int, [c_ptr, c_ll, c_int, c_int, c_ptr, c_ptr]),
    "jvmc_rbm_gram_S": (c_int, [c_ptr, c_ll, c_int, c_int, c_ptr, c_ptr, c_dbl, c_dbl, c_ptr, c_int, c_ptr]),
    "jvmc_expand_S": (c_int, [c_ptr, c_int, c_int, c_int, c_int, c_dbl, c_ptr, c_ptr]),
    "jvmc_eigh_workspace": (c_int, [c_int, c_int, ctypes.POINTER(c_ll), ctypes.POINTER(c_ll)]),
    "jvmc_eigh": (c_int, [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ll, c_ptr, c_ptr]),
    "jvmc_tdvp_regularize": (c_int, [c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr]),
}


def declared_symbols():
    """Names of every function declared in include/jvmc_b200.h."""
    with open(HEADER) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(jvmc_[a-zA-Z0-9_]+)\s*\(", txt)))


def load():
    """Load the shared library and bind every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libjvmc_b200.so not found at %s: build it with `python __graft_entry__.py` "
            "(there is no CPU fallback for the VMC hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name in declared_symbols():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        if name not in _SIGS:
            raise RuntimeError("no ctypes signature registered for " + name)
        fn.restype, fn.argtypes = _SIGS[name]
    _lib = lib
    return lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("vmc_jax_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def ptr(t):
    """Device pointer of a (contiguous) tensor, or None."""
    if t is None:
        return None
    assert t.is_contiguous(), "kernel arguments must be contiguous"
    assert not (t.is_complex() and t.is_conj()), "kernel arguments must have their conjugation materialised"
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc, what=""):
    if rc != 0:
        msg = load().jvmc_error_string(rc).decode()
        raise RuntimeError("libjvmc_b200 %s failed: %s (code %d)" % (what, msg, rc))


# kernels launched per entry point (for the gpu_launches figure of bench.py)
_KERNELS_PER_CALL = {"jvmc_bfo_matels": 2, "jvmc_rbm_moments": 2, "jvmc_i8_slice": 3, "jvmc_i8_tail_ratios": 2, "jvmc_eigh": 0,
                     "jvmc_tdvp_solve": 6, "jvmc_minsr_solve": 3,
                     "jvmc_hetrd": 0, "jvmc_unmtr": 0, "jvmc_unmqr": 0, "jvmc_secular_vectors": 2}
LAUNCHES = 0


def call(name, *args):
    """Invoke an int-returning entry point on the current stream (appended as last argument)."""
    global LAUNCHES
    require_cuda()
    lib = load()
    rc = getattr(lib, name)(*args, stream())
    check(rc, name)
    if name == "jvmc_rbm_gram_S_i8":
        LAUNCHES += 3 * max(1, ((int(args[2]) + 31) // 32 + 799) // 800)    # per <= 25600 samples: column sums, correction, Gram
    else:
        LAUNCHES += _KERNELS_PER_CALL.get(name, 1)
