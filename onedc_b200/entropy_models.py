"""Entropy-model interface of the y stream (host side), mirroring the reference classes
`EntropyCoder` / `GaussianEncoder` (/root/reference/src/modules/entropy/entropy_models.py:32-94, 252-374)
with the same method names and argument meaning, on top of the C-ABI library:

  * the rANS coder is the library's own host implementation (csrc/rans_host.cpp), called through ctypes
    (GIL released) instead of the pybind11 module MLCodec_rans;
  * `build_indexes` runs on the GPU: bf16 scales go through a 65536-entry table, fp32 scales through 255
    monotone thresholds.  Both tables are produced at `update()` time by evaluating the reference formula
    (entropy_models.py:355-362: max(s,1e-5) -> log -> (x - log 0.11)/step -> clamp -> int) with torch CPU
    fp32 ops, so the device never evaluates `log` and cannot disagree with the host by an ulp.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import lib as L


class EntropyCoder:
    def __init__(self, ec_thread=False, stream_part=1):
        assert stream_part == 1, "the OneDC codec uses a single stream part (codec_module.py:192)"
        self._lib = L.load()
        self._tables = []                         # cdf groups
        self._enc = self._lib.onedc_rans_encoder_create()
        self._dec = self._lib.onedc_rans_decoder_create()

    def __del__(self):
        try:
            self._lib.onedc_rans_encoder_destroy(self._enc)
            self._lib.onedc_rans_decoder_destroy(self._dec)
            for t, *_ in self._tables:
                self._lib.onedc_rans_tables_destroy(t)
        except Exception:
            pass

    @staticmethod
    def pmf_to_quantized_cdf(pmf, precision=16):
        p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
        out = np.zeros(len(p) + 1, dtype=np.int32)
        rc = L.load().onedc_pmf_to_quantized_cdf(p.ctypes.data, len(p), precision, out.ctypes.data)
        if rc != 0:
            raise L.OnedcError("pmf_to_quantized_cdf failed")
        return torch.from_numpy(out)

    @staticmethod
    def pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            _cdf = EntropyCoder.pmf_to_quantized_cdf(prob, 16)
            cdf[i, : _cdf.size(0)] = _cdf
        return cdf

    def reset(self):
        self._lib.onedc_rans_encoder_reset(self._enc)

    def add_cdf(self, cdf, cdf_length, offset):
        cdf = np.ascontiguousarray(cdf, dtype=np.int32)
        cdf_length = np.ascontiguousarray(cdf_length, dtype=np.int32)
        offset = np.ascontiguousarray(offset, dtype=np.int32)
        t = self._lib.onedc_rans_tables_create(cdf.ctypes.data, cdf.shape[0], cdf.shape[1], cdf_length.ctypes.data,
                                               offset.ctypes.data)
        self._tables.append((t, cdf, cdf_length, offset))
        return len(self._tables) - 1

    def tables(self, cdf_group_index=0):
        return self._tables[cdf_group_index][0]

    # ---- encoder -------------------------------------------------------------------------------
    def encode_with_indexes_np(self, symbols, indexes, cdf_group_index):
        s = np.ascontiguousarray(np.clip(symbols, -30000, 30000).astype(np.int16).reshape(-1))
        i = np.ascontiguousarray(indexes.astype(np.int16).reshape(-1))
        L.check(self._lib.onedc_rans_encoder_encode(self._enc, self.tables(cdf_group_index), s.ctypes.data,
                                                    i.ctypes.data, len(s)), "rans encode")

    def encode_with_indexes(self, symbols, indexes, cdf_group_index):
        self.encode_with_indexes_np(symbols.detach().cpu().numpy(), indexes.detach().cpu().numpy(), cdf_group_index)

    def flush(self):
        self._n = int(self._lib.onedc_rans_encoder_flush(self._enc))

    def get_encoded_stream(self):
        buf = np.empty(self._n, dtype=np.uint8)
        L.check(self._lib.onedc_rans_encoder_get_stream(self._enc, buf.ctypes.data, self._n), "rans get_stream")
        return buf.tobytes()

    # ---- decoder -------------------------------------------------------------------------------
    def set_stream(self, stream):
        a = np.frombuffer(stream, dtype=np.uint8)
        L.check(self._lib.onedc_rans_decoder_set_stream(self._dec, a.ctypes.data, len(a)), "rans set_stream")

    def decode_stream_np(self, indexes, cdf_group_index, out=None):
        idx = np.ascontiguousarray(indexes.reshape(-1), dtype=np.int16)
        if out is None:
            out = np.empty_like(idx)
        L.check(self._lib.onedc_rans_decoder_decode(self._dec, self.tables(cdf_group_index), idx.ctypes.data, len(idx),
                                                    out.ctypes.data), "rans decode")
        return out

    def decode_into(self, idx_ptr, n, out_ptr, cdf_group_index=0):
        """Raw-pointer form on the coder's own cursor (pinned host buffers of the 4-step loop); GIL released."""
        L.check(self._lib.onedc_rans_decoder_decode(self._dec, self.tables(cdf_group_index), idx_ptr, n, out_ptr),
                "rans decode")

    def decode_stream(self, indexes, cdf_group_index):
        rv = self.decode_stream_np(indexes.detach().to(torch.int16).cpu().numpy(), cdf_group_index)
        return torch.from_numpy(rv.astype(np.float32))


class StreamDecoder:
    """An independent decoder cursor (one per image of a batch) sharing the coder's CDF tables."""

    def __init__(self, coder, stream, cdf_group_index=0):
        self._lib = coder._lib
        self._tables = coder.tables(cdf_group_index)
        self._dec = self._lib.onedc_rans_decoder_create()
        a = np.frombuffer(stream, dtype=np.uint8)
        L.check(self._lib.onedc_rans_decoder_set_stream(self._dec, a.ctypes.data, len(a)), "rans set_stream")

    def decode_into(self, idx_ptr, n, out_ptr):
        L.check(self._lib.onedc_rans_decoder_decode(self._dec, self._tables, idx_ptr, n, out_ptr), "rans decode")

    def __del__(self):
        try:
            self._lib.onedc_rans_decoder_destroy(self._dec)
        except Exception:
            pass


class GaussianEncoder:
    def __init__(self, distribution="gaussian"):
        assert distribution == "gaussian"
        self.distribution = distribution
        self.scale_min, self.scale_max, self.scale_level = 0.11, 64.0, 256
        self.scale_table = torch.exp(torch.linspace(math.log(self.scale_min), math.log(self.scale_max), self.scale_level))
        self.log_scale_min = math.log(self.scale_min)
        self.log_scale_max = math.log(self.scale_max)
        self.log_scale_step = (self.log_scale_max - self.log_scale_min) / (self.scale_level - 1)
        self.entropy_coder = None
        self.cdf_group_index = None
        self._quantized_cdf = self._cdf_length = self._offset = None
        self._lut_host = self._thr_host = None
        self._dev_tables = {}

    # ---- CDF tables (init time, host) ----------------------------------------------------------
    def set_cdf_info(self, quantized_cdf, cdf_length, offset):
        self._quantized_cdf = quantized_cdf.cpu().numpy()
        self._cdf_length = cdf_length.reshape(-1).int().cpu().numpy()
        self._offset = offset.reshape(-1).int().cpu().numpy()

    def get_cdf_info(self):
        return self._quantized_cdf, self._cdf_length, self._offset

    def update(self, force=False, entropy_coder=None):
        assert entropy_coder is not None
        self.entropy_coder = entropy_coder
        if not force and self._offset is not None:
            return
        st = self.scale_table
        normal = torch.distributions.normal.Normal
        pmf_center = torch.zeros_like(st) + 50
        dist = normal(torch.zeros_like(st), st)
        for i in range(50, 1, -1):
            probs = dist.cdf(torch.zeros_like(pmf_center) + i)
            pmf_center = torch.where(probs > torch.zeros_like(pmf_center) + 0.9999, torch.zeros_like(pmf_center) + i,
                                     pmf_center)
        pmf_center = pmf_center.int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length).item())
        samples = (torch.arange(max_length) - pmf_center[:, None]).float()
        scales = torch.zeros_like(samples) + st[:, None]
        dist = normal(torch.zeros_like(scales), scales)
        upper, lower = dist.cdf(samples + 0.5), dist.cdf(samples - 0.5)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        quantized_cdf = EntropyCoder.pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self.set_cdf_info(quantized_cdf, pmf_length + 2, -pmf_center)
        self.cdf_group_index = self.entropy_coder.add_cdf(*self.get_cdf_info())
        self._build_index_tables()

    # ---- scale -> index tables -----------------------------------------------------------------
    def _formula(self, scales):
        """The reference's build_indexes arithmetic on a CPU fp32 tensor (entropy_models.py:355-362)."""
        scales = torch.maximum(scales, torch.zeros_like(scales) + 1e-5)
        indexes = (torch.log(scales) - self.log_scale_min) / self.log_scale_step
        return indexes.clamp_(0, self.scale_level - 1).int()

    def _build_index_tables(self):
        bits = torch.arange(65536, dtype=torch.int32)
        vals = (bits << 16).view(torch.float32)
        lut = self._formula(torch.nan_to_num(vals, nan=0.0)).to(torch.uint8)      # NaN scales -> index 0
        self._lut_host = lut.contiguous()
        # fp32: thr[i] = smallest positive float whose index is >= i+1 (monotone => bisection on bit patterns)
        lo = torch.full((255,), torch.tensor(1e-6).view(torch.int32).item(), dtype=torch.int64)
        hi = torch.full((255,), torch.tensor(128.0).view(torch.int32).item(), dtype=torch.int64)
        target = torch.arange(1, 256, dtype=torch.int32)
        while int((hi - lo).max()) > 1:
            mid = (lo + hi) // 2
            ok = self._formula(mid.to(torch.int32).view(torch.float32)) >= target
            hi = torch.where(ok, mid, hi)
            lo = torch.where(ok, lo, mid)
        self._thr_host = hi.to(torch.int32).view(torch.float32).contiguous()

    def device_tables(self, device):
        key = str(device)
        if key not in self._dev_tables:
            self._dev_tables[key] = (self._lut_host.to(device), self._thr_host.to(device))
        return self._dev_tables[key]

    def build_indexes(self, scales, skip_thres=None):
        assert skip_thres is None, "skip_thres is never set on the OneDC path (compression_model.py:39)"
        if not scales.is_cuda:
            raise L.OnedcError("build_indexes runs on the GPU only (no CPU fallback)")
        lut, thr = self.device_tables(scales.device)
        s = scales.contiguous()
        out = torch.empty(s.shape, device=s.device, dtype=torch.int32)
        from .ops import _dt, _stream
        L.check(L.load().onedc_build_indexes(s.data_ptr(), _dt(s), lut.data_ptr(), thr.data_ptr(), out.data_ptr(),
                                             s.numel(), _stream()), "build_indexes")
        return out

    def encode(self, x, scales, skip_thres=None):
        indexes = self.build_indexes(scales, skip_thres)
        return self.entropy_coder.encode_with_indexes(x.reshape(-1), indexes.reshape(-1), self.cdf_group_index)

    def decode_stream(self, scales, dtype, device, skip_thres=None):
        indexes = self.build_indexes(scales, skip_thres)
        val = self.entropy_coder.decode_stream(indexes.reshape(-1), self.cdf_group_index)
        return val.reshape(scales.shape).to(device).to(dtype)
