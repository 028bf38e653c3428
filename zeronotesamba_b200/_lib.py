"""ctypes binding of libzns_sm100.so (include/zns.h).  No fallback: if the library is missing or a
call fails, an exception is raised -- nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZNS_LIB_PATH: load another build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get("ZNS_LIB_PATH") or os.path.join(_HERE, "libzns_sm100.so")

c_void_p, c_int, c_float, c_double, c_ll, c_u32 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_longlong, C.c_uint32


class ZnsError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """zns_conv_desc (include/zns.h)."""
    _fields_ = [
        ("batch", c_int), ("H", c_int), ("W", c_int), ("c_in", c_int), ("c_out", c_int), ("kh", c_int), ("kw", c_int),
        ("relu", c_int), ("dropout_p", c_float), ("seed", c_u32), ("rng_stream", c_u32), ("seed_dev", c_void_p),
        ("out_scale", c_float), ("fmt", c_int),
    ]


FMT_IN_F16, FMT_W_F16, FMT_OUT_F16 = 1, 2, 4
FMT_FORWARD_F16 = FMT_IN_F16 | FMT_W_F16 | FMT_OUT_F16


_PROTOS = {
    "zns_version": (c_int, []),
    "zns_last_error": (C.c_char_p, []),
    "zns_device_check": (c_int, []),
    "zns_vqt_basis_host": (c_int, [c_int, c_int, c_int, c_double, c_double, c_int, c_void_p, c_void_p, C.POINTER(c_int)]),
    "zns_vqt_decimator_taps_host": (c_int, [c_void_p]),
    "zns_vqt_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_double, c_double, c_int, c_int, C.POINTER(c_void_p)]),
    "zns_vqt_plan_destroy": (c_int, [c_void_p]),
    "zns_vqt_num_frames": (c_int, [c_int, c_int]),
    "zns_vqt_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "zns_vqt_forward_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "zns_crop_gather": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "zns_rms_gate": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "zns_conv1_fwd": (c_int, [c_void_p, c_ll, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_u32,
                              c_void_p, c_u32, c_int, c_void_p, c_void_p]),
    "zns_conv1_wgrad": (c_int, [c_void_p, c_void_p, c_ll, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "zns_conv_fwd": (c_int, [C.POINTER(ConvDesc), c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p),
                             C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), c_void_p]),
    "zns_conv_wgrad": (c_int, [C.POINTER(ConvDesc), c_int, C.POINTER(c_void_p), C.POINTER(c_void_p),
                               C.POINTER(c_void_p), c_void_p]),
    "zns_bias_grad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "zns_pack_weights": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "zns_unpack_grads": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    "zns_pool_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_u32, c_void_p, c_u32,
                             c_int, c_void_p, c_void_p]),
    "zns_pool_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "zns_head_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "zns_head_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                             c_float, c_int, c_void_p]),
    "zns_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "zns_conv1_fwd_nbr": (c_int, [c_int, C.POINTER(c_void_p), c_ll, c_ll, C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_int, c_int, c_float, c_u32, c_void_p, c_u32,
                                  c_int, C.POINTER(c_void_p), c_void_p]),
    "zns_conv1_wgrad_nbr": (c_int, [c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), c_ll, c_ll, C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_int, c_int, c_void_p]),
    "zns_pool_fwd_nbr": (c_int, [c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_float, c_u32, c_void_p, c_u32, c_int,
                                 C.POINTER(c_void_p), c_void_p]),
    "zns_pool_bwd_nbr": (c_int, [c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "zns_head_fwd_nbr": (c_int, [c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_int, c_int, c_void_p]),
    "zns_head_bwd_nbr": (c_int, [c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_int, c_float, c_int, c_void_p]),
    "zns_bias_grad_nbr": (c_int, [c_int, C.POINTER(c_void_p), c_int, c_int, c_int, c_int, C.POINTER(c_void_p), c_void_p]),
    "zns_pack_weights_multi": (c_int, [c_int, C.POINTER(c_void_p), C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_void_p]),
    "zns_unpack_grads_multi": (c_int, [c_int, C.POINTER(c_void_p), C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int), c_float, c_int, c_int, C.POINTER(c_void_p), c_void_p]),
    "zns_zero": (c_int, [c_void_p, c_ll, c_void_p]),
    "zns_conv_pool_fwd": (c_int, [C.POINTER(ConvDesc), c_int, c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), c_void_p]),
    "zns_pool_bwd_arg_nbr": (c_int, [c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "zns_act_from_nchw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "zns_act_to_nchw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "zns_ntxent_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "zns_bce_fwd_bwd": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p]),
    "zns_adam_flat": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float, c_float, c_int,
                              c_void_p, c_float, c_void_p]),
    "zns_adam_p2p": (c_int, [c_int, c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), c_void_p, c_void_p, c_ll, c_float, c_float,
                             c_float, c_float, c_int, c_void_p, c_void_p]),
    "zns_counter_add": (c_int, [c_void_p, c_u32, c_void_p]),
    "zns_dbg_conv_fwd_simt": (c_int, [C.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "zns_dbg_conv_wgrad_simt": (c_int, [C.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p]),
    "zns_dbg_umma_probe": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "zns_dbg_umma_rate": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "zns_dbg_umma_raw": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                 c_int, c_void_p, c_void_p]),
    "zns_dbg_vqt_level_plan": (c_int, [c_int, c_int, c_int, c_int, c_double, c_double, c_int, c_void_p, c_int, c_void_p, c_int]),
    "zns_dbg_vqt_timing": (c_int, [c_void_p]),
    "zns_dbg_conv_fwd_plan": (c_int, [C.POINTER(ConvDesc), c_int, C.POINTER(c_int)]),
    "zns_dbg_conv_wgrad_plan": (c_int, [C.POINTER(ConvDesc), c_int, C.POINTER(c_int), C.POINTER(c_u32)]),
}

EXPORTS = tuple(_PROTOS.keys())

# CUDA kernels one call launches (for the launch accounting bench.py reports); 0 = host only
KERNELS_PER_CALL = {
    "zns_vqt_forward": 9, "zns_vqt_forward_host": 9, "zns_crop_gather": 1, "zns_rms_gate": 1, "zns_conv1_fwd": 1, "zns_conv1_wgrad": 1,
    "zns_conv_fwd": 1, "zns_conv_wgrad": 1, "zns_bias_grad": 1, "zns_pack_weights": 1, "zns_unpack_grads": 1,
    "zns_conv1_fwd_nbr": 1, "zns_conv1_wgrad_nbr": 1, "zns_pool_fwd_nbr": 1, "zns_pool_bwd_nbr": 1, "zns_head_fwd_nbr": 1,
    "zns_head_bwd_nbr": 1, "zns_bias_grad_nbr": 1, "zns_pack_weights_multi": 1, "zns_unpack_grads_multi": 1, "zns_zero": 0, "zns_conv_pool_fwd": 1, "zns_pool_bwd_arg_nbr": 1,
    "zns_pool_fwd": 1, "zns_pool_bwd": 1, "zns_head_fwd": 1, "zns_head_bwd": 1, "zns_merge": 1, "zns_act_from_nchw": 1,
    "zns_act_to_nchw": 1, "zns_ntxent_fwd_bwd": 1, "zns_bce_fwd_bwd": 1, "zns_adam_flat": 1, "zns_adam_p2p": 1, "zns_counter_add": 1, "zns_dbg_conv_fwd_simt": 1,
    "zns_dbg_conv_wgrad_simt": 1, "zns_dbg_umma_probe": 1, "zns_dbg_umma_rate": 1, "zns_dbg_umma_raw": 1,
}
CALL_COUNTS: dict = {}


class _LibProxy:
    """Attribute access returns the ctypes function wrapped with a call counter."""

    def __init__(self, handle: C.CDLL):
        self._h = handle
        self._fns = {}

    def __getattr__(self, name):
        fn = self._fns.get(name)
        if fn is None:
            raw = getattr(self._h, name)

            def fn(*a, _raw=raw, _name=name):
                CALL_COUNTS[_name] = CALL_COUNTS.get(_name, 0) + 1
                return _raw(*a)

            self._fns[name] = fn
        return fn


def kernel_launches(counts: Optional[dict] = None) -> int:
    """Kernels launched by the calls recorded in ``counts`` (default: all calls so far)."""
    c = CALL_COUNTS if counts is None else counts
    return sum(n * KERNELS_PER_CALL.get(k, 0) for k, n in c.items())


_lib: Optional[_LibProxy] = None


def lib() -> _LibProxy:
    """Load the shared library once.  Raises ZnsError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZnsError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C zeronotesamba_b200/csrc`.  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = _LibProxy(handle)
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise ZnsError(f"libzns_sm100 error {rc}: {lib().zns_last_error().decode(errors='replace')}")


def ptr(t) -> int:
    """Device (or host) address of a torch tensor / numpy array; None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def ptr_array(ts: Sequence) -> "C.Array":
    arr = (c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = ptr(t)
    return arr


def int_array(vals: Sequence[int]) -> "C.Array":
    return (c_int * len(vals))(*[int(v) for v in vals])


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def conv_desc(batch, H, W, c_in, c_out, kh, kw, relu=0, dropout_p=0.0, seed=0, rng_stream=0, seed_dev=None,
              out_scale=1.0, fmt=0) -> ConvDesc:
    return ConvDesc(batch, H, W, c_in, c_out, kh, kw, int(relu), float(dropout_p), int(seed) & 0xFFFFFFFF,
                    int(rng_stream), ptr(seed_dev), float(out_scale), int(fmt))
