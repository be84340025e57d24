"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by onedc_b200/).

End-to-end CPU fp32 restatement of the decode hot path (and of the encode-side twin of the
4-step prior loop that is needed to produce decodable synthetic streams).

Follows, in order (paths under /root/reference/src):
  models/sd15_onedc_codec_stage1/codec_module.py:357-369   IntraNoAR.decode
  models/sd15_onedc_codec_stage1/codec_module.py:418-454   IntraNoAR._decompress
  modules/entropy/compression_model.py:369-407             decompress_four_part_prior
  modules/entropy/compression_model.py:303-358,224-239     forward_four_part_prior(write=True) / process_with_mask
  modules/entropy/compression_model.py:410-465             forward_four_part_prior_recon_with_z (z-only variant)
  models/sd15_onedc_codec_stage1/model_sd15_with_codec_stage1.py:296-330   SD15_1step_codec_stage1.decode
The reference's CPU path is pure fp32 (autocast("cuda") disables itself without CUDA), and so is this.
"""
import numpy as np
import torch

from . import entropy as E
from .nets import (CodecNets, UNetOracle, VAEOracle, alphas_cumprod_sd15, fsq_indices_to_codes,
                   x0_from_noise)


class CodecOracle:
    def __init__(self, codec_sd):
        self.nets = CodecNets().eval()
        self.nets.load_state_dict(codec_sd, strict=True)
        self.cdf, self.cdf_len, self.cdf_off = E.gaussian_cdf_table()
        self.rans = E.RansOracle(self.cdf, self.cdf_len, self.cdf_off)

    # ---- stage 1: hyper-synthesis ------------------------------------------------------------
    @torch.no_grad()
    def hyper(self, z_idx):
        """z_idx int64 (1,hz,wz) -> (common_params (1,256,h16,w16), z_semantic (1,128,hz,wz))."""
        z_hat = fsq_indices_to_codes(z_idx)
        z_entropy, z_sem = self.nets.hyper_dec(z_hat)
        return self.nets.y_prior_fusion(z_entropy), z_sem

    def _prior(self, k, y_hat_so_far, common_red):
        n = self.nets
        adaptor = (None, n.y_spatial_prior_adaptor_1, n.y_spatial_prior_adaptor_2, n.y_spatial_prior_adaptor_3)[k]
        params = torch.cat((y_hat_so_far, common_red), dim=1)
        return n.y_spatial_prior(adaptor(params)).chunk(2, 1)

    # ---- stage 2: 4-step decode loop -----------------------------------------------------------
    @torch.no_grad()
    def decompress(self, stream_y, common_params, trace=None):
        n = self.nets
        scales, means = common_params.chunk(2, 1)
        common_red = n.y_spatial_prior_reduction(common_params)
        B, C, H, W = means.shape
        masks = E.four_part_masks(B, C, H, W)
        self.rans.set_stream(stream_y)
        y_hat = None
        for k in range(4):
            if k > 0:
                scales, means = self._prior(k, y_hat, common_red)
            scales_r = E.combine_for_writing(scales * masks[k])
            idx = E.build_indexes(scales_r)
            sym = self.rans.decode(idx.reshape(-1).numpy())
            y_q_r = torch.from_numpy(sym.astype(np.float32)).reshape(scales_r.shape)
            cur = (torch.cat((y_q_r,) * 4, dim=1) + means) * masks[k]
            y_hat = cur if y_hat is None else y_hat + cur
            if trace is not None:
                trace.append(dict(scales=scales.clone(), means=means.clone(), idx=idx.clone(),
                                  sym=sym.copy(), y_hat=y_hat.clone()))
        return y_hat

    # ---- encode-side twin (E1) -----------------------------------------------------------------
    @torch.no_grad()
    def compress(self, common_params, y=None, seed=None, trace=None):
        """Runs forward_four_part_prior(write=True) + 4x gaussian_encoder.encode + flush.
        If `y` is None, y is *sampled inside the loop from the model's own prediction*
        (SURVEY.md 8d synthetic stream generator): y = means_k + clamp(scales_k, .11, 64) * N(0,1).
        Returns (stream_y bytes, y_hat)."""
        n = self.nets
        scales, means = common_params.chunk(2, 1)
        common_red = n.y_spatial_prior_reduction(common_params)
        B, C, H, W = means.shape
        masks = E.four_part_masks(B, C, H, W)
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        groups, y_hat = [], None
        for k in range(4):
            if k > 0:
                scales, means = self._prior(k, y_hat, common_red)
            if y is None:
                yk = means + scales.clamp(0.11, 64.0) * torch.randn(means.shape, generator=g)
            else:
                yk = y
            m = masks[k]
            scales_hat, means_hat = scales * m, means * m
            y_q = torch.round((yk - means_hat) * m)
            cur = y_q + means_hat
            y_hat = cur if y_hat is None else y_hat + cur
            y_q_w = E.combine_for_writing(y_q)
            scales_w = E.combine_for_writing(scales_hat)
            idx = E.build_indexes(scales_w)
            groups.append((y_q_w.reshape(-1).numpy(), idx.reshape(-1).numpy()))
            if trace is not None:
                trace.append(dict(scales=scales.clone(), means=means.clone(), idx=idx.clone(),
                                  sym=y_q_w.reshape(-1).numpy().astype(np.int16), y_hat=y_hat.clone()))
        return self.rans.encode(groups), y_hat

    # ---- z-only variant (Z1) -------------------------------------------------------------------
    @torch.no_grad()
    def means_only(self, common_params):
        n = self.nets
        scales, means = common_params.chunk(2, 1)
        common_red = n.y_spatial_prior_reduction(common_params)
        masks = [m.to(means.device) for m in E.four_part_masks(*means.shape)]
        y_hat = None
        for k in range(4):
            if k > 0:
                scales, means = self._prior(k, y_hat, common_red)
            cur = means * masks[k]
            y_hat = cur if y_hat is None else y_hat + cur
        return y_hat

    # ---- stage 3 ---------------------------------------------------------------------------------
    @torch.no_grad()
    def synthesis(self, y_hat, z_sem):
        y_sem = self.nets.semantic_adaptor(z_sem)
        return self.nets.dec(y_hat, y_sem), y_sem

    @torch.no_grad()
    def decode(self, stream, trace=None):
        """IntraNoAR.decode -> (x_hat, y_semantic, (H,W), (padH,padW), pad_tuple)."""
        d = E.decode_container(stream)
        hz, wz = d["pad_height"] // 64, d["pad_width"] // 64
        z_idx = torch.from_numpy(E.unpack_z_indices(d["bit_stream_z"], hz * wz)).reshape(1, hz, wz)
        common, z_sem = self.hyper(z_idx)
        y_hat = self.decompress(d["bit_stream_y"], common, trace)
        x_hat, y_sem = self.synthesis(y_hat, z_sem)
        return x_hat, y_sem, (d["height"], d["width"]), (d["pad_height"], d["pad_width"]), d["pad_tuple"]

    @torch.no_grad()
    def make_stream(self, height, width, seed, trace=None):
        """Synthetic stream for an HxW image (no analysis transform needed): random z indices,
        y sampled from the model's own prior (see compress)."""
        pl, pr, pt, pb = E.padding_size(height, width)
        hz, wz = (height + pb) // 64, (width + pr) // 64
        g = torch.Generator().manual_seed(seed)
        z_idx = torch.randint(0, 16384, (1, hz, wz), generator=g)
        common, _ = self.hyper(z_idx)
        stream_y, y_hat = self.compress(common, seed=seed + 7919, trace=trace)
        stream = E.encode_container(height, width, stream_y, E.pack_z_indices(z_idx.numpy()))
        return stream, z_idx, y_hat


class OneDCOracle:
    """SD15_1step_codec_stage1.decode restated (fp32 CPU)."""

    def __init__(self, codec_sd, unet_sd, vae_sd, attn_patch=16, timestep=999):
        self.codec = CodecOracle(codec_sd)
        self.unet = UNetOracle().eval()
        self.unet.load_state_dict(unet_sd, strict=True)
        self.vae = VAEOracle(attn_patch).eval()
        self.vae.load_state_dict(vae_sd, strict=True)
        self.alphas = alphas_cumprod_sd15()
        self.timestep = timestep

    def to(self, device):
        self.unet.to(device)
        self.vae.to(device)
        return self

    @torch.no_grad()
    def generate(self, x_hat, y_sem, stages=None):
        """UNet(t=999) -> x0 (float64 formula) -> VAE decode.  x_hat (1,320,h8,w8), y_sem (1,768,hz,wz)."""
        dev = next(self.unet.parameters()).device
        ctx = y_sem.flatten(2).transpose(1, 2).contiguous().to(dev)
        t = torch.full((x_hat.shape[0],), self.timestep, dtype=torch.long, device=dev)
        eps, reduced = self.unet(x_hat.to(dev), t, ctx)
        x0 = x0_from_noise(reduced.double(), eps.double(), self.alphas.double().to(dev), t).float()
        img = self.vae(x0)
        if stages is not None:
            stages.update(eps=eps.cpu(), reduced=reduced.cpu(), x0=x0.cpu())
        return img

    @torch.no_grad()
    def decode(self, stream, stages=None):
        x_hat, y_sem, (H, W), (pH, pW), pad = self.codec.decode(stream)
        img = self.generate(x_hat, y_sem, stages)
        if stages is not None:
            stages.update(x_hat=x_hat, y_sem=y_sem)
        return img[:, :, :H, :W]                     # negative F.pad == crop right/bottom (:327-329)

    @torch.no_grad()
    def decode_z_only(self, z_idx, stages=None):
        """z-only model (config 4): decoder half of sd15_onedc_codec_z_only fed z indices."""
        common, z_sem = self.codec.hyper(z_idx)
        y_hat = self.codec.means_only(common)
        x_hat, y_sem = self.codec.synthesis(y_hat, z_sem)
        return self.generate(x_hat, y_sem, stages)
