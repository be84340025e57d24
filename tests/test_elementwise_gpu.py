"""GPU parity of the HBM-bound kernels: GroupNorm, LayerNorm, softmax, depthwise conv, upsample, window
(un)partition, x0 combine, FSQ codes and the entropy-index / dequantisation kernels (bit-exact)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(shape, dev, seed, scale=1.0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).to(torch.bfloat16).to(dev)


@pytest.mark.parametrize("n,h,w,c0,c1", [(1, 16, 16, 320, 0), (2, 12, 12, 768, 0), (1, 24, 24, 1280, 640),
                                         (1, 64, 64, 128, 0), (1, 8, 8, 2560, 0), (1, 300, 300, 128, 0)])
@pytest.mark.parametrize("silu", [True, False])
def test_groupnorm(cuda, n, h, w, c0, c1, silu):
    from onedc_b200 import ops
    x = _mk((n, h, w, c0), cuda, 1, 1.5, 0.7)
    x2 = _mk((n, h, w, c1), cuda, 2, 0.5, -1.0) if c1 else None
    c = c0 + c1
    g = torch.Generator().manual_seed(3)
    gamma, beta = 1 + 0.1 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    gn = ops.GroupNorm(gamma, beta, 1e-5, device=cuda)
    out = gn(x, x2, silu=silu)
    xs = x.float() if x2 is None else torch.cat([x.float(), x2.float()], -1)
    ref = F.group_norm(xs.permute(0, 3, 1, 2), 32, gamma.to(cuda), beta.to(cuda), 1e-5)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() < 3e-2
    out2 = gn(x, x2, silu=silu)
    assert torch.equal(out, out2), "GroupNorm must be run-to-run deterministic"


def test_groupnorm_fp32_input(cuda):
    from onedc_b200 import ops
    x = torch.randn((1, 20, 20, 256), generator=torch.Generator().manual_seed(1)).to(cuda)
    gn = ops.GroupNorm(torch.ones(256), torch.zeros(256), 1e-6, device=cuda)
    out = gn(x, silu=False)
    ref = F.group_norm(x.permute(0, 3, 1, 2), 32, eps=1e-6).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("c", [320, 640, 1280])
def test_layernorm(cuda, c):
    from onedc_b200 import ops
    x = _mk((2, 77, c), cuda, 1, 2.0, 0.3)
    g = torch.Generator().manual_seed(3)
    gamma, beta = 1 + 0.1 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    out = ops.LayerNorm(gamma, beta, device=cuda)(x)
    ref = F.layer_norm(x.float(), (c,), gamma.to(cuda), beta.to(cuda), 1e-5)
    assert (out.float() - ref).abs().max().item() < 3e-2


def test_dwconv_upsample_window(cuda):
    from onedc_b200 import ops
    x = _mk((2, 12, 20, 256), cuda, 1)
    w = torch.randn((256, 1, 3, 3), generator=torch.Generator().manual_seed(2)) * 0.3
    b = torch.randn(256, generator=torch.Generator().manual_seed(3)) * 0.1
    out = ops.dwconv3x3(x, ops.DepthwiseW(w, b, cuda))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(cuda), b.to(cuda), padding=1, groups=256).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() < 2e-2
    up = ops.upsample2x(x)
    refu = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), refu)
    xw = _mk((2, 32, 48, 64), cuda, 4)
    p = ops.window_partition(xw, 16)
    refp = xw.view(2, 2, 16, 3, 16, 64).permute(0, 1, 3, 2, 4, 5).reshape(12, 256, 64)
    assert torch.equal(p, refp)
    m = ops.window_merge(p, xw, 16)
    assert torch.equal(m.float(), (xw.float() * 2).to(torch.bfloat16).float())


def test_softmax_rows(cuda):
    from onedc_b200 import lib as L, ops
    s = torch.randn((300, 152), generator=torch.Generator().manual_seed(1)).to(cuda) * 4
    out = torch.empty((300, 152), device=cuda, dtype=torch.bfloat16)
    L.check(L.load().onedc_softmax_rows(s.data_ptr(), 152, 300, 152, 144, 0.5, out.data_ptr(), 152, ops._stream()))
    ref = torch.softmax(s[:, :144] * 0.5, dim=-1)
    assert (out[:, :144].float() - ref).abs().max().item() < 4e-3
    assert out[:, 144:].abs().max().item() == 0


def test_x0_prepare(cuda):
    from onedc_b200 import ops
    from onedc_b200.nets import alphas_cumprod_sd15
    g = torch.Generator().manual_seed(1)
    red, eps = torch.randn((1, 32, 32, 4), generator=g).to(cuda), torch.randn((1, 32, 32, 4), generator=g).to(cuda)
    pq_w, pq_b = torch.randn((4, 4), generator=g) * 0.5, torch.randn(4, generator=g) * 0.1
    a = alphas_cumprod_sd15().double()[999]
    out, x0 = ops.x0_prepare(red, eps, float(a.sqrt()), float((1 - a).sqrt()), 1 / 0.18215, pq_w, pq_b, want_x0=True)
    x0_ref = ((red.double() - (1 - a) ** 0.5 * eps.double()) / a ** 0.5).float()
    assert torch.allclose(x0, x0_ref, rtol=1e-5, atol=1e-4)
    z_ref = (x0_ref / 0.18215) @ pq_w.to(cuda).t() + pq_b.to(cuda)
    z = out[..., :4].float() + out[..., 4:].float()
    assert ((z - z_ref).abs() / z_ref.abs().clamp_min(1.0)).max().item() < 1e-4     # hi+lo split keeps ~16 bits


def test_fsq_codes(cuda):
    from onedc_b200 import ops
    from oracle.nets import fsq_indices_to_codes
    idx = torch.randint(0, 16384, (2, 12, 12), generator=torch.Generator().manual_seed(1), dtype=torch.int32)
    out = ops.fsq_codes(idx.to(cuda)).float().cpu()
    ref = fsq_indices_to_codes(idx).permute(0, 2, 3, 1)
    assert torch.equal(out[..., :7], ref) and out[..., 7].abs().max().item() == 0


def test_build_indexes_exhaustive_bf16_and_fp32(cuda):
    """KAT: every bf16 bit pattern, plus 4M fp32 values dense around the bin edges, against the oracle formula."""
    from onedc_b200.entropy_models import EntropyCoder, GaussianEncoder
    from oracle import entropy as E
    ge = GaussianEncoder()
    ge.update(force=True, entropy_coder=EntropyCoder())
    bits = torch.arange(65536, dtype=torch.int32)
    vals = (bits << 16).view(torch.float32)
    fin = ~torch.isnan(vals)
    got = ge.build_indexes(vals.to(torch.bfloat16).to(cuda)).cpu()
    ref = E.build_indexes(torch.nan_to_num(vals, nan=0.0))
    assert torch.equal(got[fin], ref[fin])
    g = torch.Generator().manual_seed(5)
    s = torch.exp(torch.rand(1 << 22, generator=g) * 24 - 17)                       # log-uniform 4e-8 .. 1e3
    edges = torch.exp(torch.tensor(E.LOG_SCALE_MIN) + torch.arange(0, 256) * E.LOG_SCALE_STEP).float()
    near = (edges[:, None] * (1 + (torch.arange(-2000, 2001) * 6e-8)[None, :])).reshape(-1)
    s = torch.cat([s, near, -s[:1000], torch.zeros(4)])
    got = ge.build_indexes(s.to(cuda)).cpu()
    assert torch.equal(got, E.build_indexes(s))


@pytest.mark.parametrize("n,h,w", [(1, 16, 16), (2, 8, 12), (1, 48, 48), (1, 31, 17)])
def test_scale_to_index_and_dequant_bit_exact(cuda, n, h, w):
    from onedc_b200 import ops
    from onedc_b200.entropy_models import EntropyCoder, GaussianEncoder
    from oracle import entropy as E
    ge = GaussianEncoder()
    ge.update(force=True, entropy_coder=EntropyCoder())
    lut = ge.device_tables(cuda)[0]
    g = torch.Generator().manual_seed(7)
    buf = torch.exp(torch.randn((n, h, w, 256), generator=g) * 2 - 1).to(torch.bfloat16)   # scales | means
    buf[..., 128:] = (torch.randn((n, h, w, 128), generator=g) * 3).to(torch.bfloat16)
    dbuf = buf.to(cuda)
    masks = E.four_part_masks(n, 128, h, w)
    scales_nchw, means_nchw = buf[..., :128].float().permute(0, 3, 1, 2), buf[..., 128:].float().permute(0, 3, 1, 2)
    params = torch.full((n, h, w, 256), 7.0, device=cuda, dtype=torch.bfloat16)
    y_ref = None
    for k in range(4):
        idx = ops.scale_to_index(dbuf[..., :128], lut, k).cpu()
        ref_idx = E.build_indexes(E.combine_for_writing(scales_nchw * masks[k]))
        assert torch.equal(idx.int(), ref_idx), f"step {k}"
        sym = torch.randint(-300, 300, (n, 32, h, w), generator=g, dtype=torch.int16)
        ops.dequant_accum(sym.to(cuda), dbuf[..., 128:], params[..., :128], k)
        # reference: bf16 arithmetic exactly as torch autocast does it (compression_model.py:383-384)
        cur = ((torch.cat((sym.to(torch.bfloat16),) * 4, dim=1) + means_nchw.to(torch.bfloat16)) * masks[k].to(torch.bfloat16))
        y_ref = cur if y_ref is None else y_ref + cur
        got = params[..., :128].float().cpu().permute(0, 3, 1, 2)
        active = sum(masks[: k + 1]) > 0
        assert torch.equal(got[active], y_ref.float()[active]), f"step {k}"
        if k == 0:
            assert got[~active].abs().max().item() == 0
    assert params[..., 128:].float().min().item() == 7.0, "dequant must not touch the other half of the buffer"


def test_quantize_residual_roundtrip(cuda):
    from onedc_b200 import ops
    g = torch.Generator().manual_seed(9)
    n, h, w = 1, 16, 24
    means = (torch.randn((n, h, w, 128), generator=g) * 2).to(torch.bfloat16).to(cuda)
    y = (torch.randn((n, h, w, 128), generator=g) * 6).to(torch.bfloat16).to(cuda)
    enc = torch.zeros((n, h, w, 128), device=cuda, dtype=torch.bfloat16)
    dec = torch.zeros((n, h, w, 128), device=cuda, dtype=torch.bfloat16)
    sym = torch.empty((n, 32, h, w), device=cuda, dtype=torch.int16)
    for k in range(4):
        ops.quantize_residual(y, means, sym, enc, k)
        ops.dequant_accum(sym, means, dec, k)
    assert torch.equal(enc, dec)
    assert (enc.float() - y.float()).abs().max().item() <= 0.5 + 0.07     # half a quantisation step + bf16 rounding


@pytest.mark.parametrize("h,w,seed", [(16, 16, 31), (5, 7, 32)])
def test_quantize_residual_matches_reference_bf16_golden(cuda, h, w, seed):
    """E1 on the device: symbols and y_hat of the encode twin vs the REFERENCE process_with_mask / quant /
    combine_for_writing (compression_model.py:87-93,224-239,296-301) run on bf16 tensors in the build container
    (tests/golden/gen_golden_generator.py): round half to even of the bf16 residual, bit-exact, ties included."""
    import os
    import sys
    import numpy as np
    from onedc_b200 import ops
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold)
    from gen_golden_generator import bf16_twin_inputs
    g = np.load(os.path.join(gold, "encode_twin_bf16.npz"))
    y, means = bf16_twin_inputs(h, w, seed)                                   # NCHW bf16 (CPU, seeded)
    y_d = y.permute(0, 2, 3, 1).contiguous().to(cuda)
    m_d = means.permute(0, 2, 3, 1).contiguous().to(cuda)
    y_hat = torch.full((1, h, w, 128), 7.0, device=cuda, dtype=torch.bfloat16)
    sym = torch.empty((1, 32, h, w), device=cuda, dtype=torch.int16)
    for k in range(4):
        ops.quantize_residual(y_d, m_d, sym, y_hat, k)
        assert np.array_equal(sym.cpu().numpy().reshape(-1), g[f"sym_{h}x{w}"][k]), f"step {k}: symbols != reference"
    ref_y_hat = torch.from_numpy(g[f"y_hat_{h}x{w}"]).view(torch.bfloat16)    # NCHW
    assert torch.equal(y_hat.cpu().permute(0, 3, 1, 2), ref_y_hat)
    assert (np.abs(g[f"sym_{h}x{w}"]) > 8).any()


@pytest.mark.parametrize("n,h,w,cin,cout,two", [(1, 32, 32, 128, 128, False), (2, 64, 48, 128, 256, False), (1, 96, 96, 64, 512, False),
                                                (1, 24, 24, 256, 320, True), (2, 96, 96, 64, 320, False), (1, 48, 48, 128, 640, True),
                                                (1, 40, 24, 64, 1280, False), (1, 24, 24, 1280, 1280, False), (1, 12, 12, 1280, 512, True)])
def test_groupnorm_statistics_fused_into_igemm(cuda, n, h, w, cin, cout, two):
    """conv -> GroupNorm with the statistics accumulated by the conv's epilogue == the two-kernel GroupNorm."""
    from onedc_b200 import ops
    ops.gn_arena_reset(cuda)
    x = _mk((n, h, w, cin), cuda, 1)
    wt = torch.randn((cout, cin, 3, 3), generator=torch.Generator().manual_seed(2)) * (cin * 9) ** -0.5
    cw = ops.ConvW(wt, torch.randn(cout, generator=torch.Generator().manual_seed(3)) * 0.3, cuda)
    y = ops.igemm(x, cw, stats=True)
    # 4/8/16/32 channels per group: per-group sums; other widths (UNet: 10/20/40): per-channel sums, which also cover
    # channel concatenations.  Layers that split K (few tiles, long K) leave the statistics to the GroupNorm kernel.
    if cout % 32 == 0 and cout // 32 in (4, 8, 16, 32):
        assert hasattr(y, "_gn_acc") and not y._gn_chan
    elif hasattr(y, "_gn_acc"):
        assert y._gn_chan
    y2 = ops.igemm(x, cw, stats=True) if two else None
    c = cout * (2 if two else 1)
    g = torch.Generator().manual_seed(4)
    gn = ops.GroupNorm(1 + 0.1 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g), 1e-5, device=cuda)
    fused = gn(y, y2)
    plain = gn(y.clone(), None if y2 is None else y2.clone())          # clones carry no accumulators
    ys = y.float() if y2 is None else torch.cat([y.float(), y2.float()], -1)
    ref = F.silu(F.group_norm(ys.permute(0, 3, 1, 2), 32, gn.gamma, gn.beta, 1e-5)).permute(0, 2, 3, 1)
    assert (fused.float() - ref).abs().max().item() < 3e-2
    assert (fused.float() - plain.float()).abs().max().item() < 4e-2          # one bf16 ulp at |y| >= 4
