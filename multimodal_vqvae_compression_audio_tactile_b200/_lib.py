"""ctypes binding of libb2c.so (include/b2c.h).  There is no CPU or PyTorch fallback:
if the library is missing, fails to load, or reports an error, this raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2c.so")

NULL_REF = 0xFFFFFFFFFFFFFFFF

ACT_NONE, ACT_SNAKE, ACT_GELU, ACT_TANH = 0, 1, 2, 3
PREC_F32, PREC_BF16X3, PREC_BF16 = 0, 1, 2
FMT_F32, FMT_BF16X2, FMT_BF16 = 0, 1, 2
ABI_VERSION = 4
ROWS_DENSE, ROWS_HEAD, ROWS_HEAD_PREV, ROWS_ZERO = 0, 1, 2, 3
PE_NONE, PE_CHUNK_POS, PE_ROW0, PE_ROW_N = 0, 1, 2, 3

KIND_NAMES = {1: "conv_f32", 2: "conv_tc", 3: "stem", 4: "head", 5: "layernorm", 6: "attention", 7: "rvq",
              8: "nearest", 9: "dac_rvq", 10: "move", 11: "conv_tc_x3"}

PRECISIONS = {"f32": PREC_F32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}
#: activation storage a contraction of this precision reads
FMT_OF_PREC = {PREC_F32: FMT_F32, PREC_BF16X3: FMT_BF16X2, PREC_BF16: FMT_BF16}
#: named precision plans: stage -> arithmetic.  "tc": everything that decides a code index keeps >= 16
#: mantissa bits (bf16 hi/lo split, 3 MMAs); the decoder, downstream of the quantizer, runs single-pass bf16.
PLANS = {
    "f32": {"enc": "f32", "pred": "f32", "dec": "f32"},
    "tc": {"enc": "bf16x3", "pred": "bf16x3", "dec": "bf16"},
    "tc_exact": {"enc": "bf16x3", "pred": "bf16x3", "dec": "bf16x3"},
    "bf16": {"enc": "bf16", "pred": "bf16", "dec": "bf16"},
    "bf16x3": {"enc": "bf16x3", "pred": "bf16x3", "dec": "bf16x3"},
}


def ref(slot: int, off: int = 0) -> int:
    return (slot << 56) | off


class HostCopy(C.Structure):
    _fields_ = [("host", C.c_void_p), ("slot", C.c_int), ("bytes", C.c_size_t)]


class B2CError(RuntimeError):
    pass


_lib = None

_fp = C.POINTER(C.c_float)
_fpp = C.POINTER(_fp)
_ref = C.c_uint64
_i = C.c_int

# name -> (restype, argtypes); the list the "library exports every symbol" test walks
SIGNATURES = {
    "b2c_last_error": (C.c_char_p, []),
    "b2c_abi_version": (_i, []),
    "b2c_debug_ru_trace": (_i, [C.POINTER(C.c_uint64), _i]),
    "b2c_ctx_create": (_i, [_i, C.POINTER(C.c_void_p)]),
    "b2c_ctx_destroy": (_i, [C.c_void_p]),
    "b2c_ctx_weight_bytes": (C.c_size_t, [C.c_void_p]),
    "b2c_pack_conv": (_i, [C.c_void_p, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i]),
    "b2c_pack_vector": (_i, [C.c_void_p, _fp, C.c_size_t]),
    "b2c_pack_codebooks": (_i, [C.c_void_p, _fpp, _i, _i, _i]),
    "b2c_pack_dac_rvq": (_i, [C.c_void_p, _i, _i, _i, _i, _fpp, _fpp, _fpp, _fpp, _fpp, _fpp, _fpp]),
    "b2c_prog_create": (_i, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "b2c_prog_destroy": (_i, [C.c_void_p]),
    "b2c_prog_num_launches": (_i, [C.c_void_p]),
    "b2c_prog_num_ops": (_i, [C.c_void_p]),
    "b2c_conv_tc_eligible": (_i, [C.c_void_p, _i, _i, _i, _i]),
    "b2c_prog_stem": (_i, [C.c_void_p, _i, _ref, _ref, _ref, _i, _i, _i, _i, _i]),
    "b2c_prog_conv": (_i, [C.c_void_p, _i, _ref, _ref, _ref, _ref, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "b2c_prog_convT": (_i, [C.c_void_p, _i, _ref, _ref, _ref, _i, _i, _i, _i, _i, _i, _i]),
    "b2c_ru_tc_eligible": (_i, [C.c_void_p, _i, _i, _i]),
    "b2c_prog_ru": (_i, [C.c_void_p, _i, _i, _i, _ref, _ref, _ref, _ref, _i, _i, _i, _i, _i, _i]),
    "b2c_prog_conv_dsnake": (_i, [C.c_void_p, _i, _ref, _ref, _i, _ref, _ref, _ref, _i, _i, _i, _i, _i, _i, _i, _i]),
    "b2c_prog_head_bwd": (_i, [C.c_void_p, _i, _i, _ref, _ref, _ref, _ref, _ref, _i, _i, _i]),
    "b2c_prog_head": (_i, [C.c_void_p, _i, _ref, _ref, _i, _i, _i]),
    "b2c_prog_layernorm": (_i, [C.c_void_p, _i, _i, _ref, _i, _ref, _i, _i, _i, C.c_float, _ref, _i, _i, _i, _i, _i]),
    "b2c_prog_layernorm_masked": (_i, [C.c_void_p, _i, _i, _ref, _ref, _i, _i, _ref, _i, _i, _i, _i, _i]),
    "b2c_prog_attention_full": (_i, [C.c_void_p, _ref, _ref, _ref, _i, _i, _i, _i]),
    "b2c_prog_select_rows": (_i, [C.c_void_p, _ref, _ref, _ref, _ref, _i, _i]),
    "b2c_prog_ema_update": (_i, [C.c_void_p, _ref, _ref, _ref, _ref, _i, _i, _i, C.c_float, C.c_float]),
    "b2c_codebooks_refresh": (_i, [C.c_void_p, _i, _i, C.c_void_p, C.c_void_p]),
    "b2c_prog_convert": (_i, [C.c_void_p, _ref, _i, _ref, _i, C.c_size_t]),
    "b2c_prog_attention": (_i, [C.c_void_p, _ref, _i, _ref, _ref, _i, _i, _i, _i, _i]),
    "b2c_rvq_scratch_bytes": (C.c_size_t, [_i, _i]),
    "b2c_prog_rvq": (_i, [C.c_void_p, _i, _i, _ref, _ref, _ref, _ref, _i, _i, _i, _i, _i, _i]),
    "b2c_prog_rvq_lookup": (_i, [C.c_void_p, _i, _i, _ref, _ref, _i, _i, _i, _i, _i]),
    "b2c_nearest_scratch_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "b2c_nearest_tc_eligible": (_i, [_i, _i, _i]),
    "b2c_prog_nearest": (_i, [C.c_void_p, _ref, _ref, _ref, _ref, _i, _i, _i, _i]),
    "b2c_prog_dac_rvq": (_i, [C.c_void_p, _i, _i, _ref, _ref, _ref, _i, _i]),
    "b2c_prog_scatter_heads": (_i, [C.c_void_p, _ref, _ref, _i, _i, _i, _i]),
    "b2c_prog_transpose": (_i, [C.c_void_p, _ref, _ref, _i, _i, _i]),
    "b2c_prog_i32_to_i64": (_i, [C.c_void_p, _ref, _ref, C.c_size_t]),
    "b2c_prog_set_lane": (_i, [C.c_void_p, _i]),
    "b2c_prog_join": (_i, [C.c_void_p]),
    "b2c_prog_run": (_i, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), _i]),
    "b2c_prog_profile": (_i, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), _i,
                             C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                             _i]),
    "b2c_prog_run_host_pipelined": (_i, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), _i,
                                         C.POINTER(HostCopy), _i, C.POINTER(HostCopy), _i, _i]),
    "b2c_metric_xcorr_align": (_i, [_i, C.c_void_p, C.c_void_p, C.c_void_p, _i, _i, _i, C.c_void_p, C.c_void_p]),
    "b2c_metric_resample": (_i, [_i, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i, _i, _i, _i, _i, _i]),
    "b2c_metric_psnr": (_i, [_i, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i, _i, C.c_float]),
    "b2c_metric_psnr_resampled": (_i, [_i, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i, _i,
                                       _i, _i, _i, C.c_float]),
    "b2c_metric_stsim_scratch_bytes": (C.c_size_t, [_i, _i, _i]),
    "b2c_metric_stsim": (_i, [_i, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i, _i,
                              _i]),
    "b2c_prog_run_host": (_i, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), _i,
                               C.POINTER(HostCopy), _i, C.POINTER(HostCopy), _i]),
}


def load():
    """Load libb2c.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise B2CError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.b2c_abi_version() != ABI_VERSION:
        raise B2CError(f"libb2c.so ABI {lib.b2c_abi_version()} != {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        msg = load().b2c_last_error().decode("utf-8", "replace")
        raise B2CError(f"{what}: {msg} (code {rc})" if what else f"{msg} (code {rc})")
    return rc
