"""Top-level decode model: the B200 drop-in for `SD15_1step_codec_stage1`
(/root/reference/src/models/sd15_onedc_codec_stage1/model_sd15_with_codec_stage1.py:17-330) on the decode
path that src/inference.py drives (:102-108, :118-119):

    model = SD15_1step_codec_stage1(args=None, accelerator=None, state_dicts=(unet_sd, codec_sd, vae_sd))
    model.eval(); model.codec_model.update(force=True)
    image = model.decode(fp=path)            # or stream=bytes  ->  fp32 [1,3,H,W], un-clamped, on the GPU

`decode()` = codec decode -> one-step UNet at t=999 (context = hyperprior tokens) -> x0 = (reduced -
sqrt(1-a) eps)/sqrt(a) -> VAE decode -> crop.  Everything after the bitstream parse and the host rANS calls is
libonedc_b200 kernels.  Extra entry points (the old ones unchanged): `decode_batch(streams)` and
`decode_z_only(z_idx)` for the hyperprior-only (0.0034 bpp) model of models/sd15_onedc_codec_z_only.
"""
import math

import torch

from . import ops
from .codec_module import IntraNoAR
from .nets import UNet, VAEDecoder, alphas_cumprod_sd15

VAE_SCALING = 0.18215


class SD15_1step_codec_stage1:
    def __init__(self, args=None, accelerator=None, state_dicts=None, device="cuda", vae_attn_patch=16,
                 conditioning_timestep=999):
        assert state_dicts is not None, "pass (unet_sd, codec_sd, vae_sd) reference-named state dicts"
        if not torch.cuda.is_available():
            raise RuntimeError("onedc_b200 needs a CUDA device (sm_100a); there is no CPU path")
        unet_sd, codec_sd, vae_sd = state_dicts
        self.args, self.accelerator = args, accelerator
        self.device = torch.device(device)
        self.conditioning_timestep = conditioning_timestep
        self.vae_attn_patch = vae_attn_patch
        self.codec_model = IntraNoAR(codec_sd, self.device)
        self.feedforward_model = UNet(unet_sd, self.device, conditioning_timestep)
        self.vae_large = VAEDecoder(vae_sd, self.device, vae_attn_patch)
        a = alphas_cumprod_sd15().double()[conditioning_timestep]
        self.sqrt_alpha = float(a.sqrt())
        self.sqrt_1m_alpha = float((1 - a).sqrt())
        self.last_stages = None
        # CUDA-graph replay of the fixed per-size launch sequence (onedc_b200/graphs.py); eager when tracing stages
        self.use_graphs = True
        self._graphed = {}

    def eval(self):
        return self

    def prepare(self):
        return self

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, x_hat, y_sem, stages=None, ctx_kv=None):
        """x_hat [B,h8,w8,320], y_sem [B,hz,wz,768] (NHWC bf16) -> padded image fp32 [B,3,H,W]."""
        b, hz, wz, cs = y_sem.shape
        ctx = y_sem.reshape(b, hz * wz, cs)                 # 'b c h w -> b (h w) c' is a no-op in NHWC
        eps, reduced = self.feedforward_model(x_hat, ctx, ctx_kv)     # ctx_kv: precomputed cross-attention K | V
        z, x0 = ops.x0_prepare(reduced, eps, self.sqrt_alpha, self.sqrt_1m_alpha, 1.0 / VAE_SCALING,
                               self.vae_large.pq_w, self.vae_large.pq_b, want_x0=stages is not None)
        img = self.vae_large(z)
        if stages is not None:
            stages.update(eps=eps.float().permute(0, 3, 1, 2).cpu(), reduced=reduced.float().permute(0, 3, 1, 2).cpu(),
                          x0=x0.permute(0, 3, 1, 2).cpu(), x_hat=x_hat.float().permute(0, 3, 1, 2).cpu(),
                          y_sem=y_sem.float().permute(0, 3, 1, 2).cpu())
        return img

    def graphed(self, batch, pad_h, pad_w):
        from .graphs import GraphedDecoder
        key = (batch, pad_h, pad_w)
        if key not in self._graphed:
            self._graphed[key] = GraphedDecoder(self, batch, pad_h, pad_w)
        return self._graphed[key]

    @torch.no_grad()
    def decode(self, fp=None, stream=None, stages=None):
        assert fp or stream
        if stages is None and self.use_graphs:
            if not stream:
                with open(fp, "rb") as f:
                    stream = f.read()
            return self.decode_batch([stream])[0]
        x_hat, y_sem, (H, W), (pH, pW), pad = self.codec_model.decode(fp=fp, stream=stream)
        img = self.generate(x_hat.permute(0, 2, 3, 1), y_sem.permute(0, 2, 3, 1), stages)
        return img[:, :, :H, :W]

    @torch.no_grad()
    def decode_batch(self, streams):
        """Same-size streams -> list of fp32 [1,3,H,W] images (fresh device tensors; with graphs `last_host_images` additionally
        holds the pinned host copy of the padded batch made by the last graph node)."""
        if self.use_graphs:
            from . import bitstream
            d0 = bitstream.decode_i(streams[0], 14, 64)
            gd = self.graphed(len(streams), d0["pad_height"], d0["pad_width"])
            host, hdrs = gd.decode(streams)
            self.last_host_images = host
            return [gd.img_dev[i:i + 1, :, :d["height"], :d["width"]].clone() for i, d in enumerate(hdrs)]
        x_hat, y_sem, hdrs = self.codec_model.decode_batch(streams)
        img = self.generate(x_hat, y_sem)
        return [img[i:i + 1, :, :d["height"], :d["width"]] for i, d in enumerate(hdrs)]

    def pipelined(self, pad_h, pad_w, depth=2):
        """Decoder that keeps `depth` images of this padded size in flight on the GPU (graphs.PipelinedDecoder)."""
        from .graphs import PipelinedDecoder
        key = ("pipe", pad_h, pad_w, depth)
        if key not in self._graphed:
            self._graphed[key] = PipelinedDecoder(self, pad_h, pad_w, depth)
        return self._graphed[key]

    @torch.no_grad()
    def decode_many(self, streams, depth=2):
        """Throughput API: same-size streams -> list of fp32 [1,3,H,W] HOST tensors, several images in flight."""
        from . import bitstream
        d0 = bitstream.decode_i(streams[0], 14, 64)
        return self.pipelined(d0["pad_height"], d0["pad_width"], depth).decode_many(streams)

    @torch.no_grad()
    def decode_resident(self, z_idx, syms):
        """Device-resident decode (bench `value` leg): inputs already in HBM, padded image left in HBM."""
        x_hat, y_sem = self.codec_model.decompress_resident(z_idx, syms)
        return self.generate(x_hat, y_sem)

    @torch.no_grad()
    def decode_z_only(self, z_idx, stages=None):
        x_hat, y_sem = self.codec_model.decode_z_only(z_idx.to(self.device, dtype=torch.int32))
        return self.generate(x_hat, y_sem, stages)
