"""ctypes binding of libdynamorph_b200.so (the C ABI in include/dynamorph_b200.h).

No CPU fallback: if the library is missing it is built in-tree with nvcc; if that fails the
import of any op raises.  Signatures carry plain pointers and sizes only."""
from __future__ import annotations

import ctypes as C
import os

from . import _build


class DmbModel(C.Structure):
    """struct dmb_model -- the constructor arguments of VQ_VAE that shape the computation."""
    _fields_ = [
        ("arch", C.c_int32), ("num_inputs", C.c_int32), ("num_hiddens", C.c_int32),
        ("num_residual_hiddens", C.c_int32), ("num_residual_layers", C.c_int32),
        ("num_embeddings", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("commitment_cost", C.c_float), ("weight_recon", C.c_float), ("weight_commitment", C.c_float),
        ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
    ]


class DmbTimeMatching(C.Structure):
    """struct dmb_time_matching -- the optional time-matching term of VQ_VAE.forward."""
    _fields_ = [("mat", C.c_void_p), ("variant", C.c_int32), ("w_a", C.c_float), ("w_t", C.c_float),
                ("w_n", C.c_float), ("margin", C.c_float), ("weight", C.c_float)]


SYNC_ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


class DmbSyncBN(C.Structure):
    """struct dmb_sync_bn -- the cross-rank sum the library calls for every BatchNorm under synchronised statistics."""
    _fields_ = [("allreduce", SYNC_ALLREDUCE_FN), ("user", C.c_void_p), ("world", C.c_int32)]


ARCH_Z16, ARCH_Z32 = 0, 1
BN_EVAL, BN_BATCH, BN_PER_SAMPLE = 0, 1, 2
BN_MODES = {"eval": BN_EVAL, "batch": BN_BATCH, "per_sample": BN_PER_SAMPLE}

_P = C.c_void_p
_M = C.POINTER(DmbModel)
_I64 = C.c_int64
_I32 = C.c_int32
_F = C.c_float

# name -> argtypes; every function returns int (0 ok)
SIGNATURES = {
    "dmb_param_count": [_M, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I32)],
    "dmb_param_lookup": [_M, C.c_char_p, C.POINTER(_I32), C.POINTER(_I64), C.POINTER(_I64)],
    "dmb_latent_shape": [_M, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32)],
    "dmb_packed_floats": [_M, C.POINTER(_I64)],
    "dmb_pack_weights": [_M, _P, _P, _I32, _P, _P],
    "dmb_workspace_bytes": [_M, _I64, _I32, _I32, C.POINTER(C.c_size_t)],
    "dmb_encoder_forward": [_M, _P, _P, _I64, _I32, _P, _P, _P, C.c_size_t, _P],
    "dmb_vq_forward": [_P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P],
    "dmb_vq_reset": [_P, _I32, _P],
    "dmb_vq_finalize": [_P, _I32, _I32, _F, _P, _P],
    "dmb_vq_gather": [_P, _P, _I64, _I32, _I32, _I32, _P, _P],
    "dmb_vq_backward": [_P, _P, _P, _P, _P, _F, _F, _I64, _I32, _I32, _I32, _P, _P, _P],
    "dmb_encode": [_M, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P, C.c_size_t, _P],
    "dmb_decoder_forward": [_M, _P, _P, _I64, _I32, _P, _P, _P, C.c_size_t, _P],
    "dmb_residual_block_sizes": [_I32, _I32, _I32, _I64, _I32, _I32, _I32, C.POINTER(_I64), C.POINTER(_I64),
                                 C.POINTER(C.c_size_t)],
    "dmb_residual_block_forward": [_I32, _I32, _I32, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, C.c_size_t, _P],
    "dmb_conv2d_forward": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _P, _I32, _P],
    "dmb_conv2d_tc_scratch_floats": [_I64, _I32, _I32, _I32, _I32, _I32, C.POINTER(_I64)],
    "dmb_conv2d_tc": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _I32, _I32, _P, _P],
    "dmb_conv2d_wino": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P],
    "dmb_conv2d_tm_scratch_floats": [_I32, _I32, _I32, C.POINTER(_I64)],
    "dmb_conv2d_tm": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _I32, _P, _P],
    "dmb_conv2d_tm_bn": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _P, C.POINTER(_I32), _P, _P],
    "dmb_bn_count_batch": [_P, _I32, _P],
    "dmb_conv2d_tm_batch_stat_rows": [C.POINTER(_I32)],
    "dmb_conv2d_tm_dgrad": [_P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                            C.POINTER(_I32), _P, _P],
    "dmb_conv_transpose2d_tm": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                C.POINTER(_I32), _P, _P],
    "dmb_residual_layer_tm_scratch_floats": [C.POINTER(_I64)],
    "dmb_residual_layer_tm": [_P, _P, _P, _P, _P, _P, _I64, _P, _P],
    "dmb_conv2d_weight_grad_scratch_floats": [_I64, _I32, _I32, _I32, _I32, _I32, _I32, C.POINTER(_I64)],
    "dmb_conv2d_weight_grad": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _I32, _P, _P, _P, _P, _P, _P],
    "dmb_conv_transpose2d_forward": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _I32, _P],
    "dmb_bench_fp32_fma": [_I32, _I32, _I32, _P, C.POINTER(C.c_double), _P],
    "dmb_bench_fma_tile": [_I32, _I32, _I32, _P, C.POINTER(C.c_double), _P],
    "dmb_bench_fma2_tile": [_I32, _I32, _I32, _P, C.POINTER(C.c_double), _P],
    "dmb_bench_fma_conv": [_I32, _I32, _I32, _P, C.POINTER(C.c_double), _P],
    "dmb_recon_loss": [_P, _P, _P, _I32, _P, _I64, _I32, _I32, _P, _P],
    "dmb_train_forward": [_M, _P, _P, _P, _P, _I32, _P, _I64, _P, _P, _P, _P, C.c_size_t, _P],
    "dmb_train_backward": [_M, _P, _P, _P, _P, _I32, _P, _P, _I64, _F, _P, _P, C.c_size_t, _P],
    "dmb_train_forward_tm": [_M, _P, _P, _P, _P, _I32, _P, _I64, C.POINTER(DmbTimeMatching), _P, _P, _P, _P, C.c_size_t, _P],
    "dmb_train_backward_tm": [_M, _P, _P, _P, _P, _I32, _P, _P, _I64, C.POINTER(DmbTimeMatching), _F, _P, _P, C.c_size_t, _P],
    "dmb_train_forward_sync": [_M, _P, _P, _P, _P, _I32, _P, _I64, C.POINTER(DmbTimeMatching), C.POINTER(DmbSyncBN), _P, _P,
                               _P, _P, C.c_size_t, _P],
    "dmb_train_backward_sync": [_M, _P, _P, _P, _P, _I32, _P, _P, _I64, C.POINTER(DmbTimeMatching), C.POINTER(DmbSyncBN),
                                _F, _P, _P, C.c_size_t, _P],
    "dmb_time_matching_scratch_floats": [_I64, _I64, C.POINTER(C.c_size_t)],
    "dmb_time_matching_forward": [_P, _I64, _I64, C.POINTER(DmbTimeMatching), _P, _P, _P],
    "dmb_time_matching_backward": [_P, _I64, _I64, _P, _F, _P, _I32, _P],
    "dmb_pca_transform": [_P, _I64, _I32, _P, _P, _I32, _P, _P, _P],
    "dmb_pca_transform_scratch_floats": [_I64, _I32, C.POINTER(_I64)],
    "dmb_pca_transform_tc": [_P, _I64, _I32, _P, _P, _I32, _P, _P, _P, _P],
    "dmb_augment_batch": [_P, _P, _I64, _I32, _I32, _I32, _P, _P],
    "dmb_adam_step": [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _I32, _F, _P],
    "dmb_adam_step_dev": [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _P, _P, _F, _P],
    "dmb_zscore_patch": [_P, _I32, _I64, _I32, _P, _P],
}

_lib = None
ABI_VERSION = 1             # include/dynamorph_b200.h: DMB_ABI_VERSION


class DmbError(RuntimeError):
    pass


class _DevPtr(C.c_void_p):
    """A device pointer that remembers which GPU it lives on (`call` makes that GPU current)."""
    dev = None


class _CurrentStream:
    """Placeholder for "the current stream of the device the pointers live on"; resolved inside `call`."""


STREAM = _CurrentStream()


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load (building first if needed) the shared library; raises if it is unavailable, stale or of another ABI."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and _build.needs_build():
        if os.path.isdir(_build.CSRC) and os.environ.get("DMB_NO_BUILD") != "1":
            try:
                _build.build()
            except Exception as e:
                # a library older than its sources may no longer match the signatures below: never run it silently
                if not os.path.exists(path) or os.environ.get("DMB_ALLOW_STALE") != "1":
                    raise DmbError("libdynamorph_b200.so is %s and could not be rebuilt (set DMB_ALLOW_STALE=1 to "
                                   "load the stale binary anyway): %s"
                                   % ("older than csrc/" if os.path.exists(path) else "missing", e)) from e
                import warnings
                warnings.warn(f"dynamorph_b200: loading a STALE libdynamorph_b200.so (rebuild failed: {e})")
    if not os.path.exists(path):
        raise DmbError("libdynamorph_b200.so not found (no CPU fallback exists); run "
                       "`python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    lib.dmb_last_error.restype = C.c_char_p
    lib.dmb_last_error.argtypes = []
    lib.dmb_abi_version.restype = C.c_int
    lib.dmb_abi_version.argtypes = []
    if lib.dmb_abi_version() != ABI_VERSION:
        raise DmbError(f"libdynamorph_b200.so has ABI version {lib.dmb_abi_version()}, this package binds "
                       f"{ABI_VERSION}; rebuild it (python -m dynamorph_b200._build --force)")
    lib.dmb_launch_count.restype = C.c_longlong
    lib.dmb_launch_count.argtypes = [C.c_int]
    for name, args in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise DmbError(f"libdynamorph_b200.so does not export {name}: stale binary, rebuild it") from e
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = lib
    return lib


def call(name: str, *args):
    """Call one C-ABI entry point.  The GPU that owns the pointer arguments is made current for the call (the library
    keeps per-device state keyed by cudaGetDevice), and a `STREAM` argument becomes that GPU's current stream."""
    lib = load()
    dev = None
    for a in args:
        dev = getattr(a, "dev", None)
        if dev is not None:
            break
    if dev is None and not any(a is STREAM for a in args):
        rc = getattr(lib, name)(*args)
    else:
        import torch
        if dev is None or dev == torch.cuda.current_device():
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            rc = getattr(lib, name)(*[st if a is STREAM else a for a in args])
        else:
            with torch.cuda.device(dev):
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                rc = getattr(lib, name)(*[st if a is STREAM else a for a in args])
    if rc != 0:
        raise DmbError(f"{name} failed ({rc}): {lib.dmb_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    p = _DevPtr(t.data_ptr())
    if t.is_cuda:
        p.dev = t.device.index
    return p
