"""ctypes binding of the C-ABI library (include/onedc_b200.h).

There is deliberately no fallback: if libonedc_b200.so is missing or a call fails, this raises.
ctypes releases the GIL around every foreign call, so the host rANS coder can run in worker threads.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libonedc_b200.so")

BF16, F32 = 0, 1
ACT_NONE, ACT_LRELU, ACT_SILU, ACT_GELU = 0, 1, 2, 3
EPI_PLAIN, EPI_PAIR_LRELU, EPI_GEGLU = 0, 1, 2
ST_NORMAL, ST_PIXSHUF, ST_TRANSPOSED, ST_QUAD = 0, 1, 2, 3


class IgemmDesc(C.Structure):
    _fields_ = [
        ("a_ptr", C.c_void_p * 2), ("a_c", C.c_int32 * 2), ("a_pix_stride", C.c_int64 * 2),
        ("n_img", C.c_int32), ("h_in", C.c_int32), ("w_in", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32),
        ("w_ptr", C.c_void_p), ("cout", C.c_int32), ("ktot", C.c_int32),
        ("w_row_stride", C.c_int64), ("w_z_stride", C.c_int64), ("w_batched", C.c_int32),
        ("bias", C.c_void_p), ("epi_mode", C.c_int32), ("act", C.c_int32), ("slope", C.c_float),
        ("res", C.c_void_p), ("res_dtype", C.c_int32), ("res_ld", C.c_int64),
        ("out", C.c_void_p), ("out_dtype", C.c_int32), ("out_ld", C.c_int64), ("out_col_off", C.c_int32),
        ("store_mode", C.c_int32), ("ps_c", C.c_int32), ("bn", C.c_int32), ("impl", C.c_int32),
        ("splitk_ws", C.c_void_p), ("splitk_ws_floats", C.c_int64), ("splitk_counters", C.c_void_p),
        ("splitk_max_tiles", C.c_int32),
        ("ntaps", C.c_int32), ("tap_dy", C.c_int32 * 9), ("tap_dx", C.c_int32 * 9), ("quad", C.c_int32),
        ("gn_acc", C.c_void_p), ("gn_groups", C.c_int32), ("gn_fused_out", C.c_int32),
        ("prefetch_ptr", C.c_void_p), ("prefetch_bytes", C.c_int64),
        ("deterministic", C.c_int32), ("w_identity_tap", C.c_int32),
    ]


_i32, _i64, _f32, _vp, _sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); also the list of symbols the header declares (checked by tests)
PROTOTYPES = {
    "onedc_last_error": (C.c_char_p, []),
    "onedc_version": (C.c_int, []),
    "onedc_launch_count": (_i64, [C.c_int]),
    "onedc_set_pdl": (C.c_int, [C.c_int]),
    "onedc_igemm": (C.c_int, [C.POINTER(IgemmDesc), _vp]),
    "onedc_igemm_set_debug": (None, [_vp]),
    "onedc_attention_set_plan": (None, [_i32, _i32]),
    "onedc_attention_ws_floats": (_i64, [_i32, _i32, _i32, _i32, _i32]),
    "onedc_attention": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _i64, _vp]),
    "onedc_groupnorm_ws_floats": (_i64, [_i32, _i64, _i32]),
    "onedc_groupnorm_stats": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i32, _i32, _i64, _i32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "onedc_groupnorm_apply": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _i32, _f32,
                                        _vp, _vp, _i32, _vp, _i64, _vp]),
    "onedc_tap_gather": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "onedc_layernorm": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _f32, _vp, _i64, _vp]),
    "onedc_softmax_rows": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _f32, _vp, _i64, _vp]),
    "onedc_softmax_rows_batched": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _i32, _f32, _vp, _i64, _vp]),
    "onedc_dwconv3x3": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "onedc_upsample2x": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "onedc_window_partition": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "onedc_window_merge": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "onedc_x0_prepare": (C.c_int, [_vp, _vp, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "onedc_scale_to_index": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "onedc_build_indexes": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "onedc_dequant_accum": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "onedc_quantize_residual": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "onedc_fsq_codes": (C.c_int, [_vp, _vp, _i64, _vp]),
    "onedc_pmf_to_quantized_cdf": (C.c_int, [_vp, _i32, _i32, _vp]),
    "onedc_rans_tables_create": (_vp, [_vp, _i32, _i32, _vp, _vp]),
    "onedc_rans_tables_destroy": (None, [_vp]),
    "onedc_rans_decoder_create": (_vp, []),
    "onedc_rans_decoder_destroy": (None, [_vp]),
    "onedc_rans_decoder_set_stream": (C.c_int, [_vp, _vp, _sz]),
    "onedc_rans_decoder_decode": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "onedc_rans_encoder_create": (_vp, []),
    "onedc_rans_encoder_destroy": (None, [_vp]),
    "onedc_rans_encoder_reset": (None, [_vp]),
    "onedc_rans_encoder_encode": (C.c_int, [_vp, _vp, _vp, _vp, _i32]),
    "onedc_rans_encoder_flush": (_i64, [_vp]),
    "onedc_rans_encoder_get_stream": (C.c_int, [_vp, _vp, _sz]),
}

_lib = None


class OnedcError(RuntimeError):
    pass


def load():
    """Loads the shared library (building nothing: use onedc_b200.build / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OnedcError(f"{LIB_PATH} not found: build it with `python -m onedc_b200.build` "
                             "(the CUDA extension is the product; there is no fallback path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().onedc_last_error()
        raise OnedcError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def set_pdl(on):
    """Programmatic dependent launch on/off for subsequent launches and graph captures; returns the old setting."""
    return bool(load().onedc_set_pdl(1 if on else 0))


def launch_count(reset=False):
    return int(load().onedc_launch_count(1 if reset else 0))
