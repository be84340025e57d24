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
import os
from collections import OrderedDict

import torch

from . import ops
from .codec_module import IntraNoAR
from .nets import UNet, VAEDecoder, alphas_cumprod_sd15
from .weights import load_checkpoint_file, vae_decoder_state_dict

VAE_SCALING = 0.18215


def _arg(args, name, default):
    """args is the reference's OmegaConf/argparse namespace (or a dict, or None)"""
    if args is None:
        return default
    if isinstance(args, dict):
        v = args.get(name, default)
    else:
        v = getattr(args, name, default)
    return default if v is None else v


class SD15_1step_codec_stage1:
    """Constructed like the reference, `SD15_1step_codec_stage1(args, accelerator)` (model...py:18; the fields read are
    `args.vae_attn_patch`, `args.conditioning_timestep`, `args.codec.*`, `accelerator.device`), without weights; then
    `model.feedforward_model.load_state_dict(sd, strict=True)` / `model.codec_model.load_state_dict(sd, strict=True)`
    exactly as src/inference.py:87-93 does.  The reference pulls the SD-2.1 VAE from the HF hub inside __init__
    (model...py:43-45); there is no network here, so the VAE decoder is loaded from `args.vae_large_path` / the
    ONEDC_VAE_PATH environment variable (a diffusers `vae/diffusion_pytorch_model.safetensors`) or through
    `model.vae_large.load_state_dict(...)`.  A network that is never loaded runs on its seed-0 random initialisation
    (what every test and bench here uses: no checkpoints exist offline).  `state_dicts=(unet, codec, vae)` is the
    shortcut the tests use."""

    def __init__(self, args=None, accelerator=None, state_dicts=None, device=None, vae_attn_patch=None,
                 conditioning_timestep=None):
        if not torch.cuda.is_available():
            raise RuntimeError("onedc_b200 needs a CUDA device (sm_100a); there is no CPU path")
        unet_sd, codec_sd, vae_sd = state_dicts if state_dicts is not None else (None, None, None)
        self.args, self.accelerator = args, accelerator
        if device is None:
            device = getattr(accelerator, "device", None) or "cuda"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("onedc_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.conditioning_timestep = int(conditioning_timestep if conditioning_timestep is not None
                                         else _arg(args, "conditioning_timestep", 999))
        self.vae_attn_patch = int(vae_attn_patch if vae_attn_patch is not None else _arg(args, "vae_attn_patch", 16))
        self.use_large_vae = True
        codec_args = _arg(args, "codec", None)
        self.codec_model = IntraNoAR(cond_ch=4, ctrl_ch=320,
                                     internal_ch=_arg(codec_args, "internal_ch", 512),
                                     bottleneck_ch=_arg(codec_args, "bottleneck_ch", 128),
                                     unet_ch_config=tuple(_arg(codec_args, "unet_ch_config", (512, 768, 768))),
                                     z_fsq_levels=tuple(_arg(codec_args, "z_fsq_levels", (4,) * 7)),
                                     state_dict=codec_sd, device=self.device)
        self.feedforward_model = UNet(unet_sd, self.device, self.conditioning_timestep)
        vae_path = _arg(args, "vae_large_path", None) or os.environ.get("ONEDC_VAE_PATH")
        if vae_sd is None and vae_path:
            vae_sd = vae_decoder_state_dict(load_checkpoint_file(vae_path))
        self.vae_large = VAEDecoder(vae_sd, self.device, self.vae_attn_patch)
        a = alphas_cumprod_sd15().double()[self.conditioning_timestep]
        self.sqrt_alpha = float(a.sqrt())
        self.sqrt_1m_alpha = float((1 - a).sqrt())
        self.last_stages = None
        # CUDA-graph replay of the fixed per-size launch sequence (onedc_b200/graphs.py); eager when tracing stages
        self.use_graphs = True
        self._graphed = OrderedDict()
        self.max_cached_sizes = int(os.environ.get("ONEDC_GRAPH_CACHE", "4"))
        for net in (self.codec_model, self.feedforward_model, self.vae_large):
            net._on_load.append(self._weights_changed)          # captured graphs hold the old weight pointers

    def release_graphs(self):
        """Drops every cached per-size decoder (captured graphs, their memory pools, pinned buffers)."""
        while self._graphed:
            self._graphed.popitem()[1].release()

    _weights_changed = release_graphs

    def eval(self):
        return self

    def prepare(self):
        return self

    def load_part_ckpt(self):
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

    def _cache_get(self, key, make):
        """LRU over (batch, padded size[, depth]): every entry owns two private CUDA-graph memory pools holding all
        activations of one decode (~1.1 GB at 768x768, ~8 GB at 2048x2048, batch 1) plus pinned host buffers, and a
        new size costs a warm-up pass + capture (~1 s), so mixed-size datasets keep at most `max_cached_sizes`
        (ONEDC_GRAPH_CACHE, default 4) entries; the evicted decoder's graphs and pools are released."""
        if key in self._graphed:
            self._graphed.move_to_end(key)
            return self._graphed[key]
        while len(self._graphed) >= max(1, self.max_cached_sizes):
            _, old = self._graphed.popitem(last=False)
            old.release()
        self._graphed[key] = make()
        return self._graphed[key]

    def graphed(self, batch, pad_h, pad_w):
        from .graphs import GraphedDecoder
        return self._cache_get((batch, pad_h, pad_w), lambda: GraphedDecoder(self, batch, pad_h, pad_w))

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
        return self._cache_get(("pipe", pad_h, pad_w, depth), lambda: PipelinedDecoder(self, pad_h, pad_w, depth))

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
        """Hyperprior-only model (models/sd15_onedc_codec_z_only): int [B,hz,wz] z indices (host or device) -> padded
        fp32 image [B,3,64hz,64wz].  One CUDA-graph replay per call unless `stages` asks for the eager trace."""
        if stages is None and self.use_graphs:
            b, hz, wz = z_idx.shape
            return self.graphed(b, hz * 64, wz * 64).decode_z_only(z_idx)
        x_hat, y_sem = self.codec_model.decode_z_only(z_idx.to(self.device, dtype=torch.int32))
        return self.generate(x_hat, y_sem, stages)
