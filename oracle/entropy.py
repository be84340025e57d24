"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by onedc_b200/).

CPU restatement of the reference's entropy-model interface for the y/z streams.
Every function cites the reference lines it follows (paths relative to
/root/reference/src).  Float steps use the *same torch CPU op sequence* as the
reference so results are bit-identical to it; integer/byte steps are numpy or the
plain-C library oracle/_build/librans_oracle.so (rans_oracle.c).

Pinned by tests/test_oracle_pinned.py against the reference imported unchanged
(oracle/ref_import.py) and against tests/golden/*.
"""
import ctypes
import math
import os
import struct

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "librans_oracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(path)
        L.oracle_pmf_to_quantized_cdf.restype = ctypes.c_int
        L.oracle_rans_encode.restype = ctypes.c_size_t
        L.oracle_rans_dec_open.restype = ctypes.c_void_p
        _LIB = L
    return _LIB


# ----------------------------------------------------------------------------- CDF tables
SCALE_MIN, SCALE_MAX, SCALE_LEVELS = 0.11, 64.0, 256      # modules/entropy/entropy_models.py:262-266
LOG_SCALE_MIN = math.log(SCALE_MIN)                          # :269
LOG_SCALE_STEP = (math.log(SCALE_MAX) - LOG_SCALE_MIN) / (SCALE_LEVELS - 1)   # :271


def pmf_to_quantized_cdf(pmf, precision=16):
    """cpp/ops/ops.cpp:24-82 via rans_oracle.c."""
    pmf = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(len(pmf) + 1, dtype=np.uint32)
    rc = _lib().oracle_pmf_to_quantized_cdf(pmf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(pmf)),
                                            ctypes.c_int(precision), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out.astype(np.int32)


def gaussian_cdf_table():
    """GaussianEncoder.update, modules/entropy/entropy_models.py:313-353 (+ EntropyCoder.pmf_to_cdf :47-55).
    Returns (cdf int32[256,103], cdf_length int32[256], offset int32[256])."""
    scale_table = torch.exp(torch.linspace(math.log(SCALE_MIN), math.log(SCALE_MAX), SCALE_LEVELS))  # :274-275
    pmf_center = torch.zeros_like(scale_table) + 50
    dist = torch.distributions.normal.Normal(torch.zeros_like(scale_table), scale_table)
    for i in range(50, 1, -1):
        probs = dist.cdf(torch.zeros_like(pmf_center) + i)
        pmf_center = torch.where(probs > torch.zeros_like(pmf_center) + 0.9999,
                                 torch.zeros_like(pmf_center) + i, pmf_center)
    pmf_center = pmf_center.int()
    pmf_length = 2 * pmf_center + 1
    max_length = int(torch.max(pmf_length).item())
    samples = (torch.arange(max_length) - pmf_center[:, None]).float()
    scales = torch.zeros_like(samples) + scale_table[:, None]
    dist = torch.distributions.normal.Normal(torch.zeros_like(scales), scales)
    upper = dist.cdf(samples + 0.5)
    lower = dist.cdf(samples - 0.5)
    pmf = upper - lower
    tail_mass = 2 * lower[:, :1]
    cdf = np.zeros((SCALE_LEVELS, max_length + 2), dtype=np.int32)
    for i in range(SCALE_LEVELS):
        prob = torch.cat((pmf[i, : int(pmf_length[i])], tail_mass[i]), dim=0)
        q = pmf_to_quantized_cdf(prob.tolist())
        cdf[i, : len(q)] = q
    return cdf, (pmf_length + 2).numpy().astype(np.int32), (-pmf_center).numpy().astype(np.int32)


def build_indexes(scales):
    """GaussianEncoder.build_indexes, entropy_models.py:355-362 (skip_thres is None on this path,
    compression_model.py:39).  `scales` is a float tensor; bf16 input is up-cast first, which is
    what autocast does to torch.log on the reference's GPU path."""
    scales = scales.float()
    scales = torch.maximum(scales, torch.zeros_like(scales) + 1e-5)
    indexes = (torch.log(scales) - LOG_SCALE_MIN) / LOG_SCALE_STEP
    indexes = indexes.clamp_(0, SCALE_LEVELS - 1)
    return indexes.int()


def bf16_index_lut():
    """build_indexes evaluated on all 65536 bf16 bit patterns -> uint8[65536].
    NaN patterns propagate through maximum/log/clamp and `.int()` of NaN is
    platform-defined in torch; the product defines them as index 0 and the
    comparison tests mask them out (a NaN scale never occurs on the path)."""
    bits = torch.arange(65536, dtype=torch.int32)
    vals = (bits << 16).view(torch.float32)
    idx = build_indexes(torch.nan_to_num(vals, nan=0.0))
    return idx.to(torch.uint8).numpy()


# ----------------------------------------------------------------------------- four-part prior
# step k, position parity p = 2*(h%2) + (w%2):  active channel group g = p ^ XOR_K[k]
# (get_mask_four_parts, compression_model.py:269-283 + :241-267).
def four_part_masks(B, C, H, W, dtype=torch.float32):
    """compression_model.py:241-283 restated: returns [mask_0..mask_3], each (B,C,H,W)."""
    assert C % 4 == 0
    hh = torch.arange(H).view(1, 1, H, 1)
    ww = torch.arange(W).view(1, 1, 1, W)
    m = [((hh % 2 == a) & (ww % 2 == b)).to(dtype) for a, b in ((0, 0), (0, 1), (1, 0), (1, 1))]
    ones = torch.ones(B, C // 4, H, W, dtype=dtype)
    order = ((0, 1, 2, 3), (3, 2, 1, 0), (2, 3, 0, 1), (1, 0, 3, 2))
    return [torch.cat([ones * m[o] for o in od], dim=1) for od in order]


def combine_for_writing(x):
    """compression_model.py:296-301."""
    x0, x1, x2, x3 = x.chunk(4, 1)
    return (x0 + x1) + (x2 + x3)


# ----------------------------------------------------------------------------- z indices / container
def pack_z_indices(idx, unit=14):
    """codec_module.py:403-409: MSB-first fixed-width fields, left-padded to whole bytes."""
    idx = np.asarray(idx).reshape(-1)
    s = "".join(bin(int(v))[2:].zfill(unit) for v in idx)
    nbytes = (len(s) + 7) // 8
    return int(s, 2).to_bytes(nbytes, "big")


def unpack_z_indices(data, count, unit=14):
    """codec_module.py:426-428."""
    s = bin(int.from_bytes(data, "big"))[2:].zfill(count * unit)
    return np.array([int(s[unit * i: unit * (i + 1)], 2) for i in range(count)], dtype=np.int64)


def padding_size(h, w, p=64):
    """modules/entropy/utils.py:7-16 (pad right/bottom only)."""
    nh, nw = (h + p - 1) // p * p, (w + p - 1) // p * p
    return 0, nw - w, 0, nh - h


def encode_container(height, width, stream_y, stream_z, caption=b""):
    """modules/entropy/utils.py:95-105: BE u32 H, W, len_y, len_caption; y; z; caption."""
    return struct.pack(">4I", height, width, len(stream_y), len(caption)) + stream_y + stream_z + caption


def decode_container(data, unit=14, ds=64):
    """modules/entropy/utils.py:108-132."""
    h, w, ly, lc = struct.unpack(">4I", data[:16])
    pl, pr, pt, pb = padding_size(h, w, ds)
    ph, pw = h + pt + pb, w + pl + pr
    lz = math.ceil((ph // ds) * (pw // ds) * unit / 8.0)
    o = 16
    y = data[o:o + ly]; o += ly
    z = data[o:o + lz]; o += lz
    cap = data[o:o + lc]
    return dict(height=h, width=w, pad_height=ph, pad_width=pw, pad_tuple=(pl, pr, pt, pb),
                bit_stream_y=y, bit_stream_z=z, bit_stream_caption=cap)


# ----------------------------------------------------------------------------- rANS (C restatement)
class RansOracle:
    """encode: list of (symbols int16, indexes int16) groups -> stream bytes (flag 0x01 + payload).
    decode: stateful cursor over a stream; decode(indexes) -> int16 symbols (rans.cpp:303-362)."""

    def __init__(self, cdf=None, lengths=None, offsets=None):
        if cdf is None:
            cdf, lengths, offsets = gaussian_cdf_table()
        self.cdf = np.ascontiguousarray(cdf, dtype=np.int32)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        self._h = None
        self._stream = None

    def _tabs(self):
        return (self.cdf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(self.cdf.shape[1]),
                self.lengths.ctypes.data_as(ctypes.c_void_p), self.offsets.ctypes.data_as(ctypes.c_void_p))

    def encode(self, groups):
        n = len(groups)
        syms = [np.ascontiguousarray(np.clip(np.asarray(s).reshape(-1), -30000, 30000).astype(np.int16)) for s, _ in groups]
        idxs = [np.ascontiguousarray(np.asarray(i).reshape(-1).astype(np.int16)) for _, i in groups]
        sp = (ctypes.c_void_p * n)(*[s.ctypes.data for s in syms])
        ip = (ctypes.c_void_p * n)(*[i.ctypes.data for i in idxs])
        cnt = (ctypes.c_int * n)(*[len(s) for s in syms])
        out = ctypes.c_void_p()
        c, stride, l, o = self._tabs()
        nbytes = _lib().oracle_rans_encode(sp, ip, cnt, ctypes.c_int(n), c, stride, l, o, ctypes.byref(out))
        data = ctypes.string_at(out.value, nbytes)
        _lib().oracle_free(out)
        return data

    def set_stream(self, stream):
        self.close()
        self._stream = np.frombuffer(bytes(stream) + b"\0" * 8, dtype=np.uint8).copy()
        self._h = ctypes.c_void_p(_lib().oracle_rans_dec_open(self._stream.ctypes.data_as(ctypes.c_void_p),
                                                              ctypes.c_size_t(len(stream))))

    def decode(self, indexes):
        idx = np.ascontiguousarray(np.asarray(indexes).reshape(-1).astype(np.int16))
        out = np.empty_like(idx)
        c, stride, l, o = self._tabs()
        _lib().oracle_rans_decode(self._h, idx.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(idx)),
                                  c, stride, l, o, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def close(self):
        if self._h is not None:
            _lib().oracle_rans_dec_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
