"""ctypes binding of libbflow_b200.so (the C ABI declared in include/bflow_b200.h).

There is no fallback: if the library is missing it is built with nvcc (in-tree); if that is impossible,
or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

MAX_SLOTS, MAX_TARGETS, MAX_DEGREE = 16, 8, 16
ABI_VERSION = 2
ACT = {'none': 0, 'relu': 1, 'sigmoid': 2, 'tanh': 3}
EPI = {'std': 0, 'gru_zr': 1, 'gru_q': 2}

c_float_p = C.c_void_p   # device pointers travel as integers


PREC = {'f32x3': 0, 'f16': 1}      # BFLOW_PREC_SPLIT3 / BFLOW_PREC_F16


class _Desc(C.Structure):
    """Descriptors start with struct_size (ABI version 2): filled in here so that no caller can forget it."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.struct_size = C.sizeof(type(self))


class ConvDesc(_Desc):
    _fields_ = [('struct_size', C.c_int), ('precision', C.c_int),
                ('x0', C.c_void_p), ('c0', C.c_int), ('ld0', C.c_int),
                ('x1', C.c_void_p), ('c1', C.c_int), ('ld1', C.c_int),
                ('w', C.c_void_p), ('ldw', C.c_int),
                ('bias', C.c_void_p),
                ('res', C.c_void_p), ('ldr', C.c_int),
                ('y', C.c_void_p), ('ldy', C.c_int),
                ('N', C.c_int), ('H', C.c_int), ('W', C.c_int), ('Ho', C.c_int), ('Wo', C.c_int), ('Cout', C.c_int),
                ('KH', C.c_int), ('KW', C.c_int), ('stride', C.c_int), ('pad_h', C.c_int), ('pad_w', C.c_int),
                ('act1', C.c_int), ('act2', C.c_int),
                ('scale', C.c_float),
                ('epi', C.c_int),
                ('aux0', C.c_void_p), ('ld_aux0', C.c_int),
                ('aux1', C.c_void_p), ('ld_aux1', C.c_int),
                ('y16_hi', C.c_void_p), ('y16_lo', C.c_void_p), ('ldy16', C.c_int),
                ('res16_hi', C.c_void_p), ('res16_lo', C.c_void_p), ('ldr16', C.c_int),
                ('aux1_16_hi', C.c_void_p), ('aux1_16_lo', C.c_void_p), ('ld_aux1_16', C.c_int),
                ('stats', C.c_void_p), ('stats_hw', C.c_int), ('max_ctas', C.c_int)]


class LookupDesc(_Desc):
    _fields_ = [('struct_size', C.c_int),
                ('n_slots', C.c_int), ('n_targets', C.c_int), ('B', C.c_int), ('h', C.c_int), ('w', C.c_int), ('radius', C.c_int),
                ('vol', C.c_void_p * MAX_SLOTS),
                ('hl', C.c_int * MAX_SLOTS), ('wl', C.c_int * MAX_SLOTS),
                ('target', C.c_int * MAX_SLOTS),
                ('inv_scale', C.c_float * MAX_SLOTS),
                ('coords', C.c_void_p),
                ('params', C.c_void_p), ('params_ld', C.c_int), ('degree', C.c_int),
                ('coef', (C.c_float * MAX_DEGREE) * MAX_TARGETS),
                ('out', C.c_void_p),
                ('out_nhwc', C.c_int),
                ('out_ld', C.c_int),
                ('out16_hi', C.c_void_p), ('out16_lo', C.c_void_p), ('out16_ld', C.c_int),
                ('tiled', C.c_int)]


class LookupOtfDesc(_Desc):
    _fields_ = [('struct_size', C.c_int),
                ('n_slots', C.c_int), ('n_targets', C.c_int), ('B', C.c_int), ('h', C.c_int), ('w', C.c_int), ('radius', C.c_int), ('D', C.c_int),
                ('f1', C.c_void_p * MAX_SLOTS), ('ld1', C.c_int),
                ('f2', C.c_void_p * MAX_SLOTS), ('ld2', C.c_int),
                ('hl', C.c_int * MAX_SLOTS), ('wl', C.c_int * MAX_SLOTS),
                ('target', C.c_int * MAX_SLOTS),
                ('inv_scale', C.c_float * MAX_SLOTS),
                ('scale', C.c_float),
                ('coords', C.c_void_p),
                ('params', C.c_void_p), ('params_ld', C.c_int), ('degree', C.c_int),
                ('coef', (C.c_float * MAX_DEGREE) * MAX_TARGETS),
                ('out', C.c_void_p), ('out_ld', C.c_int),
                ('out16_hi', C.c_void_p), ('out16_lo', C.c_void_p), ('out16_ld', C.c_int)]


_SIGNATURES = {
    'bflow_abi_version': (C.c_int, []),
    'bflow_last_error': (C.c_char_p, []),
    'bflow_built_for_sm': (C.c_int, []),
    'bflow_sizeof_conv_desc': (C.c_int, []),
    'bflow_sizeof_lookup_desc': (C.c_int, []),
    'bflow_sizeof_lookup_otf_desc': (C.c_int, []),
    'bflow_corr_lookup_otf': (C.c_int, [C.POINTER(LookupOtfDesc), C.c_void_p]),
    'bflow_feat_pool': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    'bflow_source_hash': (C.c_char_p, []),
    'bflow_zero': (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_void_p]),
    'bflow_nchw_to_nhwc': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_float, C.c_float, C.c_void_p]),
    'bflow_nhwc_to_nchw': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    'bflow_conv2d_nhwc': (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    'bflow_tma_im2col_map': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 10),
    'bflow_conv2d_nhwc_tc3': (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    'bflow_tma_tile_map': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 8),
    'bflow_tma_out_map': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int]),
    'bflow_conv2d_nhwc_tc3o': (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    'bflow_conv2d_nhwc_tc3s': (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    'bflow_conv2d_slab64': (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    'bflow_conv2d_stem7': (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    'bflow_im2col_split16': (C.c_int, [C.c_void_p] + [C.c_int] * 11 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'bflow_split_f16': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p]),
    'bflow_pack_b_tc': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    'bflow_conv2d_small_n': (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    'bflow_conv2d_thin7': (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    'bflow_plane_sums': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'bflow_instnorm_relu': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    'bflow_instnorm_relu16': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    'bflow_corr_volume': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'bflow_corr_pool': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    'bflow_corr_pool_tiled': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    'bflow_corr_lookup': (C.c_int, [C.POINTER(LookupDesc), C.c_void_p]),
    'bflow_voxelize': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    'bflow_flow_metrics': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_float, C.POINTER(C.c_float), C.c_int,
                                     C.c_void_p, C.c_void_p]),
    'bflow_voxel_norm': (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    'bflow_epe_masked': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]),
    'bflow_forward_load': (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    'bflow_forward_info': (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    'bflow_forward_run': (C.c_int, [C.c_void_p] * 8),
    'bflow_forward_destroy': (None, [C.c_void_p]),
    'bflow_arena_open': (C.c_int, [C.c_ulonglong, C.c_ulonglong]),
    'bflow_arena_alloc': (C.c_void_p, [C.c_long, C.c_int, C.c_void_p]),
    'bflow_arena_free': (None, [C.c_void_p, C.c_long, C.c_int, C.c_void_p]),
    'bflow_arena_used': (C.c_ulonglong, []),
    'bflow_arena_read': (C.c_int, [C.c_ulonglong, C.c_void_p, C.c_ulonglong]),
    'bflow_arena_close': (C.c_int, []),
    'bflow_gru_rh': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p]),
    'bflow_gru_update': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p]),
    'bflow_bezier_eval': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    'bflow_cvx_upsample': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = _build.LIB
        if _build.built_hash() != _build.source_hash():
            # missing, or compiled from other sources than the ones in the tree: rebuild (raises where nvcc is unavailable -- there is
            # no other path, and a binary that does not match its sources is never loaded silently)
            path = _build.build(force=True)
        handle = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)     # AttributeError = ABI mismatch, also loud
            fn.restype = res
            fn.argtypes = args
        if handle.bflow_abi_version() != ABI_VERSION:
            raise RuntimeError(f'libbflow_b200.so has ABI version {handle.bflow_abi_version()}, this binding needs {ABI_VERSION}')
        if ((handle.bflow_sizeof_conv_desc(), handle.bflow_sizeof_lookup_desc(), handle.bflow_sizeof_lookup_otf_desc()) !=
                (C.sizeof(ConvDesc), C.sizeof(LookupDesc), C.sizeof(LookupOtfDesc))):
            raise RuntimeError('descriptor layouts of libbflow_b200.so and bflow_b200/_lib.py differ')
        if handle.bflow_source_hash().decode() != _build.source_hash():
            raise RuntimeError('libbflow_b200.so was not built from the sources in this tree (python -m bflow_b200.build --force)')
        _lib = handle
    return _lib


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        msg = lib().bflow_last_error().decode(errors='replace')
        if rc == 1:
            raise AssertionError(f'bflow_b200 {what}: {msg}')     # the reference's convention for contract violations
        raise RuntimeError(f'bflow_b200 {what}: CUDA failure: {msg}')
