"""ctypes binding of libnhmc.so (include/nhmc.h).  No fallback: if the library is missing or a
call fails, this raises -- the CUDA extension IS the product path."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NHMC_LIB_PATH") or os.path.join(HERE, "libnhmc.so")      # override: tuning sweeps of tools/ only
CSRC = os.path.join(HERE, "csrc")

# constants of include/nhmc.h
F32, BF16, F16 = 0, 1, 2
LF_FIRST, LF_MIDDLE, LF_LAST = 0, 1, 2
SCHED_PIXEL, SCHED_LATENT, SCHED_COND = 0, 1, 2
OP_INPAINT, OP_SR, OP_HDR, OP_DEBLUR2D, OP_PHASE, OP_DENOISE, OP_COLOR, OP_SEPMAT, OP_CS = 0, 1, 2, 3, 4, 5, 6, 7, 8
GN_PLAIN, GN_SILU, GN_FILM_SILU = 0, 1, 2
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
# per-call flag word of the U-Net epilogue entry points
F_NO_PDL = 1 << 6
GN_F_NO_RESIDENT, GN_F_NO_CLUSTER, GN_F_SLICE_BWD, GN_F_SLICE_ALL, GN_F_NO_SOLO, GN_F_NO_SOLO_CLUSTER = 1, 2, 4, 8, 16, 32
GN_F_LANES_SHIFT = 8


class NhmcError(RuntimeError):
    pass


class OpDesc(C.Structure):
    """struct nhmc_op"""
    _fields_ = [("kind", C.c_int32), ("channels", C.c_int32), ("img_dim", C.c_int32), ("ratio", C.c_int32),
                ("n_kept", C.c_int32), ("pad", C.c_int32), ("no_clamp", C.c_int32), ("reserved", C.c_int32),
                ("kept_pix", C.c_void_p), ("pix_to_kept", C.c_void_p),
                ("U1", C.c_void_p), ("V1", C.c_void_p), ("U2", C.c_void_p), ("V2", C.c_void_p),
                ("gain", C.c_void_p), ("gain_pinv", C.c_void_p), ("coef", C.c_float * 12),
                ("tc_tables", C.c_void_p)]


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libnhmc.so (in-tree; nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise NhmcError("building libnhmc.so failed")
    return LIB_PATH


_SIGS = {
    "nhmc_version": (C.c_int, []),
    "nhmc_last_error": (C.c_char_p, []),
    "nhmc_sizeof_chain": (C.c_size_t, []),
    "nhmc_sizeof_op": (C.c_size_t, []),
    "nhmc_begin_iteration": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "nhmc_leapfrog_workspace_bytes": (C.c_size_t, [C.c_int]),
    "nhmc_leapfrog": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_double, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "nhmc_leapfrog_diag": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "nhmc_ddim_mask_bytes": (C.c_size_t, [C.c_int64]),
    "nhmc_ddim_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_float,
                                C.c_float, C.c_float, C.c_float, C.c_int, C.c_int64, C.c_void_p]),
    "nhmc_ddim_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int64, C.c_void_p]),
    "nhmc_clip_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "nhmc_clip_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "nhmc_op_meas_size": (C.c_int64, [C.POINTER(OpDesc)]),
    "nhmc_op_workspace_bytes": (C.c_size_t, [C.POINTER(OpDesc), C.c_int]),
    "nhmc_op_tables_bytes": (C.c_size_t, [C.POINTER(OpDesc)]),
    "nhmc_op_prepare_tables": (C.c_int, [C.POINTER(OpDesc), C.c_void_p, C.c_void_p]),
    "nhmc_op_forward": (C.c_int, [C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "nhmc_op_adjoint": (C.c_int, [C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "nhmc_op_pinv": (C.c_int, [C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "nhmc_op_loss_grad": (C.c_int, [C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_void_p]),
    "nhmc_op_prepare_measurement": (C.c_int, [C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "nhmc_op_loss_grad_prepared": (C.c_int, [C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_int, C.c_void_p, C.c_void_p]),
    "nhmc_accept": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nhmc_commit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                              C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    "nhmc_commit_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                   C.c_void_p]),
    "nhmc_gn_fwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int,
                              C.c_float, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p]),
    "nhmc_gn_bwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_uint,
                              C.c_void_p]),
    "nhmc_gn_cat_fwd": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_void_p,
                                  C.c_size_t, C.c_uint, C.c_void_p]),
    "nhmc_gn_cat_bwd": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p]),
    "nhmc_gn_workspace_bytes": (C.c_size_t, [C.c_int]),
    "nhmc_gn_check": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "nhmc_bias_add": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint, C.c_void_p]),
    "nhmc_pool2_gather": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint, C.c_void_p]),
    "nhmc_pool2_scatter": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint, C.c_void_p]),
    "nhmc_split2": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "nhmc_vq_nearest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "nhmc_softmax_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int, C.c_void_p]),
    "nhmc_attn_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                C.c_uint, C.c_void_p]),
    "nhmc_softmax_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int,
                                   C.c_void_p]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def lib():
    """The loaded library; raises NhmcError (never falls back) when it cannot be loaded."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise NhmcError("libnhmc.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(make -C noise-space-hmc_b200/csrc). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)      # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(status, what=""):
    if status != 0:
        msg = lib().nhmc_last_error().decode("utf-8", "replace")
        raise NhmcError("%s failed (status %d): %s" % (what or "nhmc call", status, msg))
