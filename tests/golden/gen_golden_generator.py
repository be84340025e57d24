"""Generates the generator-side / Z1 / E1 golden fixtures from the UNMODIFIED reference source executed in the build
container (oracle/ref_generator.py, oracle/ref_import.py).  Run there only:
    python tests/golden/gen_golden_generator.py

  generator_128x176.npz  seed-0 weights; a synthetic 128x176 stream decoded by the REFERENCE
                         SD15_1step_codec_stage1.decode (its own forward_unet, get_x0_from_noise in float64,
                         vae_decode_image, windowed_attn_forward with 16x16 and 16x8 edge windows, un-padding):
                         stream bytes + fp32 image.  (diffusers/peft block arithmetic = oracle modules, see
                         oracle/ref_generator.py for exactly what that leaves unpinned.)
  zonly_128x192.npz      z indices (2x3) -> y_hat of the reference forward_four_part_prior_recon_with_z
                         (compression_model.py:410-465) fed hyper_dec/y_prior_fusion of the same indices
  encode_twin_64x128.npz E1: y (seeded) -> the four written symbol groups, scale indices, y_hat and stream bytes of the
                         reference compress_four_part_prior + gaussian_encoder.encode x4 + flush
                         (compression_model.py:303-367, codec_module.py:388-396)
  encode_twin_bf16.npz   the same process_with_mask / quant / combine_for_writing arithmetic on BF16 tensors (what the
                         reference computes on the GPU under autocast; CPU bf16 ops round identically): for two sizes
                         (16x16 and 5x7) and seeded y / means, the four written symbol groups and y_hat (bf16 bits)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from onedc_b200 import weights as W                      # noqa: E402
from oracle.decode import CodecOracle                    # noqa: E402
from oracle.ref_generator import build_reference_model   # noqa: E402


def reference_z_only_y_hat(ref_codec, z_idx):
    z_hat = ref_codec.z_vq.indices_to_codes(z_idx)
    params, z_sem = ref_codec.hyper_dec(z_hat)
    params = ref_codec.y_prior_fusion(params)
    y = torch.zeros_like(params[:, :128])                 # only its shape/dtype is used (y_q * 0. + means_hat)
    return ref_codec.forward_four_part_prior_recon_with_z(
        y, params, ref_codec.y_spatial_prior_adaptor_1, ref_codec.y_spatial_prior_adaptor_2,
        ref_codec.y_spatial_prior_adaptor_3, ref_codec.y_spatial_prior,
        y_spatial_prior_reduction=ref_codec.y_spatial_prior_reduction)


def reference_encode_twin(ref_codec, z_idx, y):
    """codec_module.py:381-396 from `params` on (the analysis transform that makes y is out of scope)."""
    z_hat = ref_codec.z_vq.indices_to_codes(z_idx)
    params, _ = ref_codec.hyper_dec(z_hat)
    params = ref_codec.y_prior_fusion(params)
    out = ref_codec.compress_four_part_prior(
        y, params, ref_codec.y_spatial_prior_adaptor_1, ref_codec.y_spatial_prior_adaptor_2,
        ref_codec.y_spatial_prior_adaptor_3, ref_codec.y_spatial_prior,
        y_spatial_prior_reduction=ref_codec.y_spatial_prior_reduction)
    y_q_w, scales_w, y_hat = out[:4], out[4:8], out[8]
    ref_codec.entropy_coder.reset()
    for q, s in zip(y_q_w, scales_w):
        ref_codec.gaussian_encoder.encode(q, s, skip_thres=ref_codec.force_zero_thres)
    ref_codec.entropy_coder.flush()
    idx = [ref_codec.gaussian_encoder.build_indexes(s) for s in scales_w]
    return y_q_w, idx, y_hat, ref_codec.entropy_coder.get_encoded_stream()


def twin_input(seed, shape):
    """y for the E1 fixture: wide enough to hit escapes now and then"""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * 3.0


def bf16_twin_inputs(h, w, seed):
    g = torch.Generator().manual_seed(seed)
    means = (torch.randn((1, 128, h, w), generator=g) * 2).to(torch.bfloat16)
    y = (torch.randn((1, 128, h, w), generator=g) * 6).to(torch.bfloat16)
    y[:, :, 0, :] = means[:, :, 0, :] + (torch.arange(w).to(torch.bfloat16) - 2.5)       # exact .5 ties (where bf16 allows)
    return y, means


def reference_bf16_twin(ref_codec, y, means):
    """process_with_mask step by step with bf16 tensors; scales are irrelevant to the symbols."""
    B, C, H, W = y.shape
    masks = ref_codec.get_mask_four_parts(B, C, H, W, y.dtype, y.device)
    syms, y_hat = [], None
    for m in masks:
        _, y_q, y_hat_k, _ = ref_codec.process_with_mask(y, torch.ones_like(y), means, m)
        y_hat = y_hat_k if y_hat is None else y_hat + y_hat_k
        syms.append(ref_codec.combine_for_writing(y_q).clamp(-30000, 30000).to(torch.int16).reshape(-1).numpy())
    return np.stack(syms), y_hat


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    sds = (W.random_state_dict(W.unet_spec(), 0), W.random_state_dict(W.codec_spec(), 0),
           W.random_state_dict(W.vae_spec(), 0))
    ref = build_reference_model(*sds)
    orc = CodecOracle(sds[1])
    stream, z_idx, _ = orc.make_stream(128, 176, seed=5)
    img = ref.decode(stream=stream)
    assert img.shape == (1, 3, 128, 176)
    np.savez_compressed(os.path.join(HERE, "generator_128x176.npz"), stream=np.frombuffer(stream, dtype=np.uint8),
                        image=img.numpy().astype(np.float32))

    g = torch.Generator().manual_seed(11)
    z = torch.randint(0, 16384, (1, 2, 3), generator=g)
    y_hat = reference_z_only_y_hat(ref.codec_model, z)
    np.savez_compressed(os.path.join(HERE, "zonly_128x192.npz"), z_idx=z.numpy(), y_hat=y_hat.numpy().astype(np.float32))

    z = torch.randint(0, 16384, (1, 1, 2), generator=g)
    y = twin_input(21, (1, 128, 4, 8))
    y_q_w, idx, y_hat, data = reference_encode_twin(ref.codec_model, z, y)
    np.savez_compressed(os.path.join(HERE, "encode_twin_64x128.npz"), z_idx=z.numpy(),
                        sym=np.stack([q.reshape(-1).numpy().astype(np.int16) for q in y_q_w]),
                        idx=np.stack([i.reshape(-1).numpy().astype(np.int16) for i in idx]),
                        y_hat=y_hat.numpy().astype(np.float32), stream=np.frombuffer(data, dtype=np.uint8))
    out = {}
    for h, w, seed in ((16, 16, 31), (5, 7, 32)):
        yb, mb = bf16_twin_inputs(h, w, seed)
        sym, y_hat = reference_bf16_twin(ref.codec_model, yb, mb)
        out[f"sym_{h}x{w}"] = sym
        out[f"y_hat_{h}x{w}"] = y_hat.view(torch.int16).numpy()
    np.savez_compressed(os.path.join(HERE, "encode_twin_bf16.npz"), **out)
    print("generator-side golden fixtures written")
