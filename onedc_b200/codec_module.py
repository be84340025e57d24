"""Codec half of the decode path: the B200 drop-in for the reference's `IntraNoAR`
(/root/reference/src/models/sd15_onedc_codec_stage1/codec_module.py:184-454) and the
`CompressionModel` four-part prior (modules/entropy/compression_model.py:369-465).

Same constructor arguments, entry points and return values (`IntraNoAR(cond_ch, ctrl_ch, internal_ch, bottleneck_ch,
unet_ch_config, z_fsq_levels)`, `load_state_dict(sd, strict=True)`, `decode(fp=, stream=)`, `_decompress(...)`,
`decompress_four_part_prior(common_params, adaptor_1, adaptor_2, adaptor_3, y_spatial_prior, reduction)`,
`update(force)`, attributes `ds`, `cond_ds`, `index_unit_length`, `entropy_coder`, `gaussian_encoder`, the sub-module
names `hyper_dec`, `y_prior_fusion`, `y_spatial_prior*`, `semantic_adaptor`, `dec`), same wire format; the z-only
model's `forward(..., fix_codec=True)` result dict (models/sd15_onedc_codec_z_only/codec_module.py:263-308) from the
z indices on (the analysis transform that would produce them is out of scope).
Added, without changing the old calls: `decode_batch(streams)` (same-size images batched through every
kernel, host rANS of the images on a GIL-free thread pool) and `compress_synthetic(...)`, the encode-side
twin of the 4-step loop used to produce decodable streams (the analysis transform is out of scope).

Per prior step the device<->host traffic is one int16 index plane set down and one int16 symbol plane set
up (pinned buffers); everything else stays in HBM.
"""
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np
import torch

from . import bitstream, ops
from .entropy_models import EntropyCoder, GaussianEncoder, StreamDecoder
from .nets import HyperSynthesis, LatentSynthesisNet, SemanticAdaptorNet, SpatialPrior
from .weights import LazyNet, codec_spec


class CompressionModel:
    """Holds the entropy coder + Gaussian conditional of the y stream (reference CompressionModel.__init__/update,
    compression_model.py:25-52,169-171)."""

    def __init__(self, y_distribution="gaussian", z_channel=128, ec_thread=False, stream_part=1):
        self.y_distribution = y_distribution
        self.z_channel = z_channel
        self.entropy_coder = None
        self.gaussian_encoder = GaussianEncoder(distribution=y_distribution)
        self.force_zero_thres = None
        self.ec_thread, self.stream_part = ec_thread, stream_part

    def update(self, force=False):
        self.entropy_coder = EntropyCoder(self.ec_thread, self.stream_part)
        self.gaussian_encoder.update(force=force, entropy_coder=self.entropy_coder)

    def get_y_cdf_info(self):
        return self.gaussian_encoder.get_cdf_info()


class _Seq:
    """callable chain of kernel-launch blocks (the reference's nn.Sequential attributes)"""

    def __init__(self, mods):
        self.mods = list(mods)

    def __call__(self, t):
        for m in self.mods:
            t = m(t)
        return t


def _nhwc_view(t):
    """Accepts the internal NHWC tensor or the reference-shaped logical NCHW view of it; returns (NHWC, was_nchw)."""
    if t.dim() == 4 and t.stride(-1) != 1 and t.stride(1) == 1:
        return t.permute(0, 2, 3, 1), True
    return t, False


class IntraNoAR(LazyNet, CompressionModel):
    # analysis-side modules of the reference class (codec_module.py:196-201): present in model_1.safetensors,
    # not on the decode path
    IGNORED_PREFIXES = ("enc.", "hyper_enc.", "z_vq.")

    def __init__(self, cond_ch=4, ctrl_ch=320, internal_ch=512, bottleneck_ch=128, unet_ch_config=(512, 768, 768),
                 z_fsq_levels=(4,) * 7, state_dict=None, device="cuda", rans_threads=None):
        CompressionModel.__init__(self, y_distribution="gaussian", z_channel=bottleneck_ch, ec_thread=False, stream_part=1)
        assert (cond_ch, ctrl_ch, internal_ch, bottleneck_ch, tuple(unet_ch_config), tuple(z_fsq_levels)) == \
            (4, 320, 512, 128, (512, 768, 768), (4,) * 7), "only the published OneDC configuration is built"
        self.device = torch.device(device)
        self.z_fsq_levels = list(z_fsq_levels)
        self.index_unit_length = 14                  # log2(4^7)
        self.ds, self.cond_ds = 64, 8
        self.debug = False
        self._pool = ThreadPoolExecutor(max_workers=rans_threads or max(1, min(16, (os.cpu_count() or 2) - 1)))
        self._pinned = {}
        self.last_trace = None
        self._lazy_init(state_dict)

    def _spec(self):
        return codec_spec()

    def _build(self, sd):
        self.hyper = HyperSynthesis(sd, self.device)
        self.prior = SpatialPrior(sd, self.device)
        self.semantic_adaptor = SemanticAdaptorNet(sd, self.device)
        self.dec = LatentSynthesisNet(sd, self.device)
        # the reference's attribute names (codec_module.py:203-217), as callables over NHWC bf16 tensors
        self.hyper_dec = self.hyper.hyper_dec
        self.y_prior_fusion = self.hyper.y_prior_fusion
        self.y_spatial_prior_adaptor_1, self.y_spatial_prior_adaptor_2, self.y_spatial_prior_adaptor_3 = self.prior.adaptors[1:]
        self.y_spatial_prior = _Seq(self.prior.prior)
        self.y_spatial_prior_reduction = self.prior.reduce

    # ------------------------------------------------------------------------------------------
    def _pin(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        if key not in self._pinned:
            self._pinned[key] = torch.empty(shape, dtype=dtype, pin_memory=True)
        return self._pinned[key]

    def _lut(self):
        return self.gaussian_encoder.device_tables(self.device)[0]

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, fp=None, stream=None):
        assert fp or stream
        if not stream:
            with open(fp, "rb") as f:
                stream = f.read()
        d = bitstream.decode_i(stream, self.index_unit_length, self.ds)
        x_hat, y_sem = self._decompress(**d)
        return x_hat, y_sem, (d["height"], d["width"]), (d["pad_height"], d["pad_width"]), d["pad_tuple"]

    @torch.no_grad()
    def _decompress(self, bit_stream_y, bit_stream_z, pad_height, pad_width, bit_stream_caption=None, **kwargs):
        """codec_module.py:418-454, call for call; tensors are logical NCHW views over NHWC memory like the reference's."""
        self.entropy_coder.set_stream(bit_stream_y)
        z_idx = self.parse_z([bit_stream_z], pad_height, pad_width)
        params, z_semantic = self.hyper_dec(ops.fsq_codes(z_idx))
        params = self.y_prior_fusion(params)
        y_hat = self.decompress_four_part_prior(params.permute(0, 3, 1, 2),
                                                self.y_spatial_prior_adaptor_1, self.y_spatial_prior_adaptor_2,
                                                self.y_spatial_prior_adaptor_3, self.y_spatial_prior,
                                                self.y_spatial_prior_reduction)
        y_semantic = self.semantic_adaptor(z_semantic)
        x_hat = self.dec(y_hat.permute(0, 2, 3, 1), y_semantic)
        return x_hat.permute(0, 3, 1, 2), y_semantic.permute(0, 3, 1, 2)

    @torch.no_grad()
    def decode_batch(self, streams):
        """Same-size streams -> (x_hat [B,h8,w8,320] NHWC, y_sem [B,hz,wz,768] NHWC, headers)."""
        ds = [bitstream.decode_i(s, self.index_unit_length, self.ds) for s in streams]
        ph, pw = ds[0]["pad_height"], ds[0]["pad_width"]
        assert all(d["pad_height"] == ph and d["pad_width"] == pw for d in ds), "decode_batch needs equal padded sizes"
        x_hat, y_sem = self._decompress_batch([d["bit_stream_y"] for d in ds], [d["bit_stream_z"] for d in ds], ph, pw)
        return x_hat, y_sem, ds

    def parse_z(self, z_streams, pad_height, pad_width):
        hz, wz = pad_height // self.ds, pad_width // self.ds
        idx = np.stack([bitstream.unpack_indices(z, hz * wz, self.index_unit_length).reshape(hz, wz) for z in z_streams])
        host = self._pin("z", idx.shape, torch.int32)
        host.copy_(torch.from_numpy(idx))
        return host.to(self.device, non_blocking=True)

    def _decompress_batch(self, y_streams, z_streams, pad_height, pad_width, trace=None):
        z_idx = self.parse_z(z_streams, pad_height, pad_width)
        common, z_sem = self.hyper(z_idx)
        decoders = [StreamDecoder(self.entropy_coder, s, self.gaussian_encoder.cdf_group_index) for s in y_streams]
        y_hat = self._four_part_loop(common, decoders, trace)
        y_sem = self.semantic_adaptor(z_sem)
        x_hat = self.dec(y_hat, y_sem)
        return x_hat, y_sem

    # ---- the 4-step loop (compression_model.py:369-407) ------------------------------------------
    @torch.no_grad()
    def decompress_four_part_prior(self, common_params, y_spatial_prior_adaptor_1, y_spatial_prior_adaptor_2,
                                   y_spatial_prior_adaptor_3, y_spatial_prior, y_spatial_prior_reduction=None):
        """Reference signature (compression_model.py:369-373).  Reads the y stream from `self.entropy_coder`
        (`set_stream` first, as `_decompress` does); one image, like the reference.  `common_params` may be the
        internal NHWC tensor or its logical NCHW view; y_hat comes back in the same form."""
        assert y_spatial_prior_reduction is not None, "the OneDC codec always reduces the common params (codec_module.py:213)"
        common, nchw = _nhwc_view(common_params)
        assert common.shape[0] == 1, "one stream per call; batches go through decode_batch"
        y_hat = self._four_part_loop(common, [self.entropy_coder], None,
                                     (None, y_spatial_prior_adaptor_1, y_spatial_prior_adaptor_2, y_spatial_prior_adaptor_3),
                                     y_spatial_prior, y_spatial_prior_reduction)
        return y_hat.permute(0, 3, 1, 2) if nchw else y_hat

    def _four_part_loop(self, common_params, decoders, trace=None, adaptors=None, prior=None, reduction=None):
        """decoders: one rANS cursor per image (objects with `decode_into(idx_ptr, n, out_ptr)`)."""
        adaptors = adaptors or self.prior.adaptors
        prior = prior or self.y_spatial_prior
        reduction = reduction or self.y_spatial_prior_reduction
        n, h, w, _ = common_params.shape
        params = torch.empty((n, h, w, 256), device=self.device, dtype=torch.bfloat16)
        reduction(common_params, out=params[..., 128:])          # [..., :128] = y_hat_so_far, [..., 128:] = reduced
        y_hat = params[..., :128]
        nsym = 32 * h * w
        idx_dev = torch.empty((n, 32, h, w), device=self.device, dtype=torch.int16)
        sym_dev = torch.empty((n, 32, h, w), device=self.device, dtype=torch.int16)
        idx_host = self._pin("idx", (n, nsym), torch.int16)
        sym_host = self._pin("sym", (n, nsym), torch.int16)
        lut = self._lut()
        sm = common_params
        stream = torch.cuda.current_stream()
        for k in range(4):
            if k > 0:
                sm = prior(adaptors[k](params))
            ops.scale_to_index(sm[..., :128], lut, k, idx_dev)
            idx_host.copy_(idx_dev.view(n, nsym), non_blocking=True)
            stream.synchronize()
            ip, sp = idx_host.data_ptr(), sym_host.data_ptr()
            if n == 1:
                decoders[0].decode_into(ip, nsym, sp)
            else:
                list(self._pool.map(lambda i: decoders[i].decode_into(ip + 2 * i * nsym, nsym, sp + 2 * i * nsym), range(n)))
            sym_dev.view(n, nsym).copy_(sym_host, non_blocking=True)
            ops.dequant_accum(sym_dev, sm[..., 128:], y_hat, k)
            if trace is not None:
                trace.append(dict(scales=sm[..., :128].float().cpu(), means=sm[..., 128:].float().cpu(),
                                  idx=idx_host.clone(), sym=sym_host.clone(), y_hat=y_hat.float().cpu()))
        return y_hat

    @torch.no_grad()
    def decompress_resident(self, z_idx, syms):
        """GPU-only leg of the decode (bench `value`): z indices and the four decoded symbol planes are already
        resident in HBM, so no host rANS and no synchronisation happens here.  syms: 4 x int16 [B,32,h,w]."""
        common, z_sem = self.hyper(z_idx)
        params = self.prior.init_params(common)
        y_hat = params[..., :128]
        lut = self._lut()
        n, h, w, _ = common.shape
        idx_dev = torch.empty((n, 32, h, w), device=self.device, dtype=torch.int16)
        sm = common
        for k in range(4):
            if k > 0:
                sm = self.prior.step(k, params)
            ops.scale_to_index(sm[..., :128], lut, k, idx_dev)       # the index kernel still runs: it is on the path
            ops.dequant_accum(syms[k], sm[..., 128:], y_hat, k)
        y_sem = self.semantic_adaptor(z_sem)
        return self.dec(y_hat, y_sem), y_sem

    # ---- z-only model (compression_model.py:410-465): y_hat = prior means, no y stream ----------------
    @torch.no_grad()
    def decode_z_only(self, z_idx):
        """z_idx int32 [B,hz,wz] on the device -> (x_hat, y_sem) NHWC."""
        common, z_sem = self.hyper(z_idx)
        y_hat = self.forward_four_part_prior_recon_with_z(None, common)
        y_sem = self.semantic_adaptor(z_sem)
        return self.dec(y_hat, y_sem), y_sem

    @torch.no_grad()
    def forward(self, x=None, cond=None, fix_encoder=False, fix_codec=False, z_vq_indices=None):
        """Decoder half of the z-only model's `forward(x, cond, fix_codec=True)`
        (models/sd15_onedc_codec_z_only/codec_module.py:231-308): same result-dict keys.  The analysis transform
        (`enc`, `hyper_enc`, FSQ forward) is out of scope, so the z indices it would produce are passed in as
        `z_vq_indices` (int [B,hz,wz]); with y_q = 0 and scales_hat = 0 the reference's bit terms are exactly 0
        (probs_to_bits of probability 1, LowerBound 0)."""
        assert z_vq_indices is not None, "onedc_b200 has no analysis transform: pass z_vq_indices (SURVEY.md section 6)"
        z_idx = z_vq_indices.to(self.device, dtype=torch.int32)
        params, z_semantic = self.hyper_dec(ops.fsq_codes(z_idx))
        params = self.y_prior_fusion(params)
        y_hat = self.forward_four_part_prior_recon_with_z(None, params)
        y_semantic = self.semantic_adaptor(z_semantic)
        x_hat = self.dec(y_hat, y_semantic)
        zero = torch.zeros((), device=self.device)
        nchw = lambda t: t.permute(0, 3, 1, 2)
        return {"x_hat": nchw(x_hat), "y_hat": nchw(y_hat), "bit": zero, "bpp": zero, "bpp_y": zero, "bpp_hard_y": zero,
                "y_semantic": nchw(y_semantic), "z_semantic": nchw(z_semantic), "z_vq_indices": z_vq_indices,
                "params_hat": nchw(params), "y_orig": None}

    __call__ = forward

    @torch.no_grad()
    def forward_four_part_prior_recon_with_z(self, y, common_params, y_spatial_prior_adaptor_1=None,
                                             y_spatial_prior_adaptor_2=None, y_spatial_prior_adaptor_3=None,
                                             y_spatial_prior=None, y_spatial_prior_reduction=None, write=False):
        """compression_model.py:421-465: y_hat_k = means_k * mask_k (`y` only lends its shape there; unused here)."""
        common, nchw = _nhwc_view(common_params)
        adaptors = (None, y_spatial_prior_adaptor_1 or self.y_spatial_prior_adaptor_1,
                    y_spatial_prior_adaptor_2 or self.y_spatial_prior_adaptor_2,
                    y_spatial_prior_adaptor_3 or self.y_spatial_prior_adaptor_3)
        prior = y_spatial_prior or self.y_spatial_prior
        n, h, w, _ = common.shape
        params = torch.empty((n, h, w, 256), device=self.device, dtype=torch.bfloat16)
        (y_spatial_prior_reduction or self.y_spatial_prior_reduction)(common, out=params[..., 128:])
        y_hat = params[..., :128]
        sm = common
        for k in range(4):
            if k > 0:
                sm = prior(adaptors[k](params))
            ops.dequant_accum(None, sm[..., 128:], y_hat, k)
        return y_hat.permute(0, 3, 1, 2) if nchw else y_hat

    # ---- encode-side twin of the loop (compression_model.py:303-358) ---------------------------------
    @torch.no_grad()
    def compress_synthetic(self, height, width, seed, trace=None):
        """Builds a decodable stream for an HxW image without the analysis transform: random z indices,
        y drawn inside the loop from the model's own prediction, y = means + clamp(scales,.11,64)*N(0,1)
        (SURVEY.md section 8d).  Uses the same prior-net kernels as decode, so the decoder sees bit-identical
        scales; symbols are entropy-coded with the library's rANS encoder."""
        pl, pr, pt, pb = bitstream.get_padding_size(height, width, self.ds)
        ph, pw = height + pb, width + pr
        hz, wz = ph // self.ds, pw // self.ds
        g = torch.Generator().manual_seed(seed)
        z_idx_h = torch.randint(0, 16384, (1, hz, wz), generator=g, dtype=torch.int32)
        common, _ = self.hyper(z_idx_h.to(self.device))
        n, h, w, _ = common.shape
        params = self.prior.init_params(common)
        y_hat = params[..., :128]
        lut = self._lut()
        gd = torch.Generator(device=self.device).manual_seed(seed + 7919)
        sym = torch.empty((1, 32, h, w), device=self.device, dtype=torch.int16)
        self.entropy_coder.reset()
        sm = common
        for k in range(4):
            if k > 0:
                sm = self.prior.step(k, params)
            scales, means = sm[..., :128], sm[..., 128:]
            y = (means.float() + scales.float().clamp(0.11, 64.0)
                 * torch.randn(means.shape, device=self.device, generator=gd)).to(torch.bfloat16)
            idx = ops.scale_to_index(scales, lut, k)
            ops.quantize_residual(y, means, sym, y_hat, k)
            torch.cuda.current_stream().synchronize()
            self.entropy_coder.encode_with_indexes_np(sym.cpu().numpy(), idx.cpu().numpy(),
                                                      self.gaussian_encoder.cdf_group_index)
            if trace is not None:
                trace.append(dict(idx=idx.cpu().clone(), sym=sym.cpu().clone(), y_hat=y_hat.float().cpu()))
        self.entropy_coder.flush()
        stream_y = self.entropy_coder.get_encoded_stream()
        stream_z = bitstream.pack_indices(z_idx_h.numpy(), self.index_unit_length)
        return bitstream.encode_i(height, width, stream_y, stream_z, b"", 0), z_idx_h
