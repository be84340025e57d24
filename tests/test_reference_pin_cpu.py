"""Pins the oracle against the reference's OWN source for everything the reference wrote itself.

Two layers:
  * live: the unmodified reference is imported/executed from /root/reference (build container only; skipped elsewhere) --
    SD15_1step_codec_stage1.decode, forward_unet, windowed_attn_forward, get_x0_from_noise, IntraNoAR.decode,
    forward_four_part_prior_recon_with_z (Z1), compress_four_part_prior + gaussian_encoder.encode (E1);
  * golden: fixtures those same functions produced (tests/golden/gen_golden_generator.py), checked everywhere.
All comparisons are bit-exact (torch.equal / byte equality) on CPU fp32.

What remains unpinned is stated in oracle/ref_generator.py: the per-module arithmetic of diffusers 0.32.1 / peft 0.14.0
(neither installed nor vendored).
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)

HAVE_REF = os.path.isdir("/root/reference/src")
live = pytest.mark.skipif(not HAVE_REF, reason="reference tree only exists in the build container")


@pytest.fixture(scope="module")
def sds():
    from onedc_b200 import weights as W
    return (W.random_state_dict(W.unet_spec(), 0), W.random_state_dict(W.codec_spec(), 0),
            W.random_state_dict(W.vae_spec(), 0))


@pytest.fixture(scope="module")
def oracle(sds):
    from oracle.decode import OneDCOracle
    torch.set_grad_enabled(False)
    return OneDCOracle(sds[1], sds[0], sds[2])


@pytest.fixture(scope="module")
def ref(sds):
    from oracle.ref_import import reference_available
    if not reference_available():
        pytest.skip("oracle/_ref not built (make -C oracle ref)")
    from oracle.ref_generator import build_reference_model
    return build_reference_model(*sds)


# ------------------------------------------------------------------------------------------------ golden (everywhere)
def test_oracle_full_decode_matches_reference_golden_image(oracle):
    """The image the REFERENCE decode() produced for this stream (its own forward_unet / x0 / windowed VAE source)."""
    g = np.load(os.path.join(GOLD, "generator_128x176.npz"))
    img = oracle.decode(g["stream"].tobytes())
    assert img.shape == (1, 3, 128, 176)
    assert np.array_equal(img.numpy(), g["image"]), float(np.abs(img.numpy() - g["image"]).max())


def test_oracle_z_only_matches_reference_golden(oracle):
    g = np.load(os.path.join(GOLD, "zonly_128x192.npz"))
    common, _ = oracle.codec.hyper(torch.from_numpy(g["z_idx"]))
    assert np.array_equal(oracle.codec.means_only(common).numpy(), g["y_hat"])


def test_oracle_encode_twin_matches_reference_golden(oracle):
    from gen_golden_generator import twin_input
    g = np.load(os.path.join(GOLD, "encode_twin_64x128.npz"))
    common, _ = oracle.codec.hyper(torch.from_numpy(g["z_idx"]))
    trace = []
    data, y_hat = oracle.codec.compress(common, y=twin_input(21, (1, 128, 4, 8)), trace=trace)
    for k in range(4):
        assert np.array_equal(trace[k]["sym"], g["sym"][k]), f"round-half-even symbols, step {k}"
        assert np.array_equal(trace[k]["idx"].reshape(-1).numpy().astype(np.int16), g["idx"][k])
    assert np.abs(g["sym"]).max() > 3, "fixture should exercise more than the trivial symbols"
    assert np.array_equal(y_hat.numpy(), g["y_hat"])
    assert data == g["stream"].tobytes()


# ------------------------------------------------------------------------------------------------ live (build container)
@live
def test_reference_decode_source_equals_oracle(ref, oracle):
    """reference SD15_1step_codec_stage1.decode, executed unmodified, == oracle decode, bit for bit; size with right
    padding (176 -> 192) and 16x8 edge windows in the VAE mid attention."""
    stream, _, _ = oracle.codec.make_stream(128, 176, seed=5)
    a = oracle.decode(stream)
    b = ref.decode(stream=stream)
    assert a.shape == b.shape == (1, 3, 128, 176)
    assert torch.equal(a, b), float((a - b).abs().max())
    g = np.load(os.path.join(GOLD, "generator_128x176.npz"))
    assert stream == g["stream"].tobytes() and np.array_equal(b.numpy(), g["image"]), "golden fixture is stale"


@live
def test_reference_forward_unet_equals_oracle_unet(ref, oracle):
    """forward_unet (decoder_unet.py:32-305) bound onto the oracle's blocks vs UNetOracle.forward: skip-stack order,
    reduced-sample branch, time embedding plumbing.  Non-square input, batch 2."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 320, 8, 16, generator=g)
    ctx = torch.randn(2, 6, 768, generator=g)
    t = torch.full((2,), 999, dtype=torch.long)
    eps_r, red_r = ref.feedforward_model(sample=x, timestep=t, encoder_hidden_states=ctx, added_cond_kwargs=None)
    eps_o, red_o = oracle.unet(x, t, ctx)
    assert torch.equal(eps_r, eps_o) and torch.equal(red_r, red_o)


@live
def test_reference_x0_formula_equals_oracle(ref):
    from oracle.nets import alphas_cumprod_sd15, x0_from_noise
    from oracle.ref_generator import import_reference_generator
    f = import_reference_generator().get_x0_from_noise
    g = torch.Generator().manual_seed(1)
    s, e = torch.randn(2, 4, 5, 7, generator=g).double(), torch.randn(2, 4, 5, 7, generator=g).double()
    a = alphas_cumprod_sd15().double()
    for ts in ([999, 999], [0, 500]):
        t = torch.tensor(ts, dtype=torch.long)
        assert torch.equal(f(s, e, a, t), x0_from_noise(s, e, a, t))


@live
@pytest.mark.parametrize("patch,h,w", [(16, 16, 24), (16, 20, 12), (8, 16, 16), (64, 16, 24)])
def test_reference_windowed_attention_equals_oracle(ref, sds, patch, h, w):
    """windowed_attn_forward (autoencoders_patch_attn.py:9-62, eval branch, in-place window write-back, smaller edge
    windows) on the oracle's mid block vs the oracle's own window loop."""
    from oracle.nets import VAEOracle
    from oracle.ref_generator import DiffusersShapedVAE, import_reference_generator
    ov = VAEOracle(patch).eval()
    ov.load_state_dict(sds[2], strict=True)
    rv = DiffusersShapedVAE(ov, import_reference_generator().windowed_attn_forward, patch).eval()
    z = torch.randn(1, 4, h, w, generator=torch.Generator().manual_seed(2))
    a = ov.decoder(ov.post_quant_conv(z))
    b = rv.decode(z).sample
    assert torch.equal(a, b)


@live
def test_reference_z_only_loop_equals_oracle(ref, oracle):
    """Z1: forward_four_part_prior_recon_with_z (compression_model.py:410-465) vs CodecOracle.means_only."""
    from gen_golden_generator import reference_z_only_y_hat
    for shape, seed in (((1, 2, 3), 11), ((1, 3, 1), 12)):
        z = torch.randint(0, 16384, shape, generator=torch.Generator().manual_seed(seed))
        y_ref = reference_z_only_y_hat(ref.codec_model, z)
        common, _ = oracle.codec.hyper(z)
        assert torch.equal(oracle.codec.means_only(common), y_ref)


@live
def test_reference_encode_twin_equals_oracle(ref, oracle):
    """E1: process_with_mask / quant (round half even) / combine_for_writing / gaussian_encoder.encode x4 / flush
    (compression_model.py:87-93,224-239,303-367; codec_module.py:388-396) vs CodecOracle.compress, bit-exact:
    symbols, indices, y_hat and the stream bytes.  y includes exact .5 residual ties."""
    from gen_golden_generator import reference_encode_twin, twin_input
    z = torch.randint(0, 16384, (1, 1, 2), generator=torch.Generator().manual_seed(3))
    common, _ = oracle.codec.hyper(z)
    y = twin_input(21, (1, 128, 4, 8))
    # force exact ties: y - means = k + 0.5 on step-0 positions (means_0 known from the hyperprior alone)
    means0 = common[:, 128:]
    y[:, :, 0, 0] = means0[:, :, 0, 0] + torch.arange(128).float().remainder(7) - 3.5
    y_q_w, idx, y_hat, data = reference_encode_twin(ref.codec_model, z, y)
    trace = []
    data_o, y_hat_o = oracle.codec.compress(common, y=y, trace=trace)
    for k in range(4):
        assert np.array_equal(trace[k]["sym"], y_q_w[k].reshape(-1).numpy().astype(np.int16)), f"symbols step {k}"
        assert torch.equal(trace[k]["idx"].reshape(-1).int(), idx[k].reshape(-1).int()), f"indices step {k}"
    assert torch.equal(y_hat_o, y_hat)
    assert data_o == data
