"""End-to-end GPU parity of the decode path against the CPU oracle (oracle/), stage by stage.

A learned codec's stream is only decodable by the prior nets that produced its scales, so the stream is made by
the product's own encode twin and the oracle is applied per stage on the product's inputs:
  * indices: bit-exact given the same scales; rANS symbols / bytes: bit-exact (oracle C coder);
  * y_hat: bit-exact given the same symbols and means; round trip encode -> bytes -> decode lossless;
  * float stages (hyper-synthesis, prior nets, g_s, UNet, x0, VAE): max-abs / rel-L2 vs the fp32 oracle, final
    image PSNR >= 45 dB on [0,1] (test_quality.py:232-233 definition).
"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H, W = 256, 256


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _psnr01(a, b):
    a = a.clamp(-1, 1) * 0.5 + 0.5
    b = b.clamp(-1, 1) * 0.5 + 0.5
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 10 * math.log10(1.0 / max(mse, 1e-20))


@pytest.fixture(scope="module")
def bundle(cuda):
    from onedc_b200 import weights as Wt
    from onedc_b200.model import SD15_1step_codec_stage1
    from oracle.decode import OneDCOracle
    sds = (Wt.random_state_dict(Wt.unet_spec(), 0), Wt.random_state_dict(Wt.codec_spec(), 0),
           Wt.random_state_dict(Wt.vae_spec(), 0))
    model = SD15_1step_codec_stage1(state_dicts=sds, device=cuda)
    model.eval()
    model.codec_model.update(force=True)
    oracle = OneDCOracle(sds[1], sds[0], sds[2])
    if os.environ.get("ONEDC_ORACLE_DEVICE", "cuda") == "cuda":
        oracle.to(cuda)          # same fp32 restatement, executed by torch on the GPU (TF32 off) to save minutes
    enc_trace = []
    stream, z_idx = model.codec_model.compress_synthetic(H, W, seed=1234, trace=enc_trace)
    return dict(model=model, oracle=oracle, stream=stream, z_idx=z_idx, enc_trace=enc_trace, dev=cuda)


def test_roundtrip_lossless_and_entropy_bit_exact(bundle):
    from onedc_b200 import bitstream
    from oracle import entropy as E
    m, orc = bundle["model"].codec_model, bundle["oracle"].codec
    d = bitstream.decode_i(bundle["stream"])
    assert (d["height"], d["width"]) == (H, W)
    dec_trace = []
    z_idx = m.parse_z([d["bit_stream_z"]], d["pad_height"], d["pad_width"])
    assert torch.equal(z_idx.cpu(), bundle["z_idx"]), "z index stream round trip"
    assert np.array_equal(E.unpack_z_indices(d["bit_stream_z"], z_idx.numel()), bundle["z_idx"].numpy().reshape(-1))
    x_hat, y_sem = m._decompress_batch([d["bit_stream_y"]], [d["bit_stream_z"]], d["pad_height"], d["pad_width"], dec_trace)
    masks = E.four_part_masks(1, 128, H // 16, W // 16)
    orc.rans.set_stream(d["bit_stream_y"])
    groups = []
    for k in range(4):
        e, t = bundle["enc_trace"][k], dec_trace[k]
        assert torch.equal(t["idx"].view(-1), e["idx"].view(-1)), f"step {k}: encoder and decoder saw different indices"
        assert torch.equal(t["sym"].view(-1), e["sym"].view(-1)), f"step {k}: symbols not recovered losslessly"
        assert torch.equal(t["y_hat"], e["y_hat"]), f"step {k}: y_hat differs between encoder and decoder"
        # indices bit-exact vs the oracle formula on the SAME (product) scales
        ref_idx = E.build_indexes(E.combine_for_writing(t["scales"].permute(0, 3, 1, 2) * masks[k]))
        assert torch.equal(t["idx"].reshape(-1).int(), ref_idx.reshape(-1)), f"step {k}: index kernel != oracle"
        # oracle C rANS decodes the same symbols from the same bytes
        sym_o = orc.rans.decode(t["idx"].view(-1).numpy())
        assert np.array_equal(sym_o, t["sym"].view(-1).numpy()), f"step {k}: oracle rANS decode differs"
        groups.append((t["sym"].view(-1).numpy(), t["idx"].view(-1).numpy()))
    assert orc.rans.encode(groups) == d["bit_stream_y"], "oracle rANS encoder produces different bytes"
    bundle["dec_trace"] = dec_trace
    bundle["x_hat"], bundle["y_sem"] = x_hat, y_sem


def test_codec_float_stages_vs_oracle(bundle):
    if "dec_trace" not in bundle:
        test_roundtrip_lossless_and_entropy_bit_exact(bundle)
    m, orc = bundle["model"].codec_model, bundle["oracle"].codec
    common_o, z_sem_o = orc.hyper(bundle["z_idx"].long())
    common, z_sem = m.hyper(bundle["z_idx"].to(bundle["dev"]))
    r = _rel_l2(common.float().cpu().permute(0, 3, 1, 2), common_o)
    assert r < 2e-2, f"hyper-synthesis rel-L2 {r}"
    assert _rel_l2(z_sem.float().cpu().permute(0, 3, 1, 2), z_sem_o) < 1e-2
    # prior nets: oracle fed the product's y_hat_so_far
    red_o = orc.nets.y_spatial_prior_reduction(common_o)
    agree = []
    for k in range(1, 4):
        y_prev = bundle["dec_trace"][k - 1]["y_hat"].permute(0, 3, 1, 2)
        s_o, m_o = orc._prior(k, y_prev, red_o)
        t = bundle["dec_trace"][k]
        rs, rm = _rel_l2(t["scales"].permute(0, 3, 1, 2), s_o), _rel_l2(t["means"].permute(0, 3, 1, 2), m_o)
        assert rs < 3e-2 and rm < 3e-2, f"prior step {k}: rel-L2 scales {rs} means {rm}"
    # synthesis on the product's y_hat
    y_hat = bundle["dec_trace"][3]["y_hat"].permute(0, 3, 1, 2)
    x_hat_o, y_sem_o = orc.synthesis(y_hat, z_sem_o)
    r1, r2 = _rel_l2(bundle["x_hat"].float().cpu().permute(0, 3, 1, 2), x_hat_o), \
        _rel_l2(bundle["y_sem"].float().cpu().permute(0, 3, 1, 2), y_sem_o)
    print(f"x_hat rel-L2 {r1:.4g}  y_sem rel-L2 {r2:.4g}")
    assert r1 < 3e-2 and r2 < 3e-2
    bundle["x_hat_o"], bundle["y_sem_o"] = x_hat_o, y_sem_o


def test_full_decode_psnr_vs_oracle(bundle):
    if "x_hat_o" not in bundle:
        test_codec_float_stages_vs_oracle(bundle)
    model, oracle = bundle["model"], bundle["oracle"]
    st = {}
    img = model.decode(stream=bundle["stream"], stages=st)
    assert img.shape == (1, 3, H, W) and img.dtype == torch.float32
    so = {}
    # the oracle generator is fed the product's (bf16) codec outputs -> isolates UNet + x0 + VAE error ...
    img_gen = oracle.generate(st["x_hat"], st["y_sem"], so).cpu()
    for k in ("eps", "reduced", "x0"):
        print(f"{k}: rel-L2 {_rel_l2(st[k], so[k]):.4g} max-abs {float((st[k] - so[k]).abs().max()):.4g}")
    p_gen = _psnr01(img.cpu(), img_gen[:, :, :H, :W])
    # ... and fed the oracle's own fp32 codec outputs -> the whole float path after the entropy decode
    img_all = oracle.generate(bundle["x_hat_o"], bundle["y_sem_o"]).cpu()
    p_all = _psnr01(img.cpu(), img_all[:, :, :H, :W])
    print(f"PSNR vs oracle (generator only) {p_gen:.2f} dB, (codec+generator) {p_all:.2f} dB, "
          f"max-abs {float((img.cpu() - img_all).abs().max()):.4g}")
    assert torch.isfinite(img).all()
    assert p_all >= 45.0, f"bf16 path PSNR {p_all:.2f} dB < 45 dB vs the fp32 oracle"


def test_z_only_decode_vs_oracle(bundle):
    model, oracle = bundle["model"], bundle["oracle"]
    z_idx = torch.randint(0, 16384, (1, 2, 2), generator=torch.Generator().manual_seed(3), dtype=torch.int32)
    img = model.decode_z_only(z_idx).cpu()
    ref = oracle.decode_z_only(z_idx.long()).cpu()
    p = _psnr01(img, ref)
    print(f"z-only PSNR {p:.2f} dB")
    assert p >= 45.0


def test_batch_decode_matches_single(bundle):
    model = bundle["model"]
    s2, _ = model.codec_model.compress_synthetic(H, W, seed=1235)
    a = model.decode(stream=bundle["stream"])
    b = model.decode(stream=s2)
    both = model.decode_batch([bundle["stream"], s2])
    # batch size changes tiling / split-K factors, i.e. fp32 summation order: equal up to bf16 rounding noise
    assert _psnr01(both[0].cpu(), a.cpu()) > 48 and _psnr01(both[1].cpu(), b.cpu()) > 48


def test_graph_replay_matches_eager(bundle):
    """CUDA-graph replay (the default decode route) must be bit-identical to eager launches, run after run."""
    model = bundle["model"]
    eager = model.decode(stream=bundle["stream"], stages={})
    g1 = model.decode(stream=bundle["stream"])
    g2 = model.decode(stream=bundle["stream"])
    assert torch.equal(g1, g2), "graph replays differ run to run"
    # the eager route re-wraps x_hat (NCHW view and back), which drops the fused-statistics hint of one GroupNorm:
    # same math, different summation order => equal up to bf16 rounding noise, not bitwise
    assert _psnr01(eager.cpu(), g1.cpu()) > 55, "graph replay differs from eager"
    host = model.last_host_images
    assert torch.equal(host[:, :, :H, :W], g1.cpu())


@pytest.mark.parametrize("hh,ww", [(192, 320), (100, 70)])
def test_sizes_with_edge_windows_and_padding(bundle, hh, ww):
    """Sizes whose latent is not a multiple of the 16-token VAE attention window (edge windows are smaller,
    autoencoders_patch_attn.py:22-28) and sizes that need right/bottom padding to a multiple of 64."""
    model, oracle = bundle["model"], bundle["oracle"]
    stream, _ = model.codec_model.compress_synthetic(hh, ww, seed=4321)
    st = {}
    img = model.decode(stream=stream, stages=st)
    assert img.shape == (1, 3, hh, ww)
    ref = oracle.generate(st["x_hat"], st["y_sem"]).cpu()[:, :, :hh, :ww]
    p = _psnr01(img.cpu(), ref)
    print(f"{hh}x{ww}: PSNR vs oracle generator {p:.2f} dB")
    assert p >= 45.0
    g = model.decode(stream=stream)                       # graph route
    assert _psnr01(g.cpu(), img.cpu()) > 52 and torch.equal(g, model.decode(stream=stream))


def test_programmatic_dependent_launch_is_bit_identical(bundle):
    """Every kernel is launched with the programmatic-stream-serialization attribute (kernel i+1's prologue overlaps
    kernel i's tail).  Early starts must never change a result: graphs captured with PDL off and on give the same
    bits, replay after replay."""
    from onedc_b200 import lib
    from onedc_b200.graphs import GraphedDecoder
    model = bundle["model"]
    outs = []
    old = lib.set_pdl(False)
    try:
        for on in (False, True):
            lib.set_pdl(on)
            gd = GraphedDecoder(model, 1, H, W)
            for _ in range(3):
                host, _ = gd.decode([bundle["stream"]])
                outs.append(host.clone())
    finally:
        lib.set_pdl(old)
    for o in outs[1:]:
        assert torch.equal(o, outs[0]), "PDL changed the decoded image"


def test_pipelined_decode_matches_one_by_one(bundle):
    """model.decode_many keeps two images in flight (two graph sets on two streams and scratch lanes, host rANS of one
    image under the kernels of the other): every image must equal the one-at-a-time graph decode bit for bit."""
    model = bundle["model"]
    streams = [bundle["stream"]] + [model.codec_model.compress_synthetic(H, W, seed=2000 + i)[0] for i in range(4)]
    singles = [model.decode(stream=s).cpu() for s in streams]
    for depth in (2, 3):
        many = model.decode_many(streams, depth=depth)
        assert len(many) == len(streams)
        for a, b in zip(singles, many):
            assert torch.equal(a, b), "pipelined decode differs from the single-image decode"
