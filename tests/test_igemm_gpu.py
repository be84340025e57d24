"""GPU parity of the tcgen05 implicit-GEMM kernel (through the C ABI) against fp32 PyTorch convolutions of the
same bf16-rounded operands, and against the library's SIMT checking kernel."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(shape, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)


def _ref_conv(x, w, b, k, stride, x2=None):
    xs = x.float() if x2 is None else torch.cat([x.float(), x2.float()], dim=-1)
    y = F.conv2d(xs.permute(0, 3, 1, 2), w, b, stride=stride, padding=k // 2)
    return y.permute(0, 2, 3, 1).contiguous()


def _close(a, b, tol=1.5e-2):
    a, b = a.float(), b.float()
    scale = b.abs().max().clamp_min(1e-6)
    err = (a - b).abs().max() / scale
    assert err < tol, f"max rel-to-peak error {err.item():.4g}"


CASES = [
    # n, h, w, cin, cout, k, stride
    (1, 1, 200, 64, 64, 1, 1),
    (1, 1, 128, 128, 256, 1, 1),
    (2, 12, 12, 128, 128, 1, 1),
    (1, 16, 16, 320, 320, 3, 1),
    (1, 24, 20, 64, 48, 3, 1),
    (2, 16, 16, 128, 640, 3, 1),
    (1, 32, 32, 256, 128, 3, 1),
    (1, 16, 16, 320, 320, 3, 2),
    (1, 8, 8, 1280, 1280, 3, 2),
    (1, 4, 4, 8, 128, 1, 1),
    (1, 32, 32, 8, 512, 3, 1),
    (1, 32, 32, 320, 4, 3, 1),
    (1, 48, 48, 512, 2048, 1, 1),
]


@pytest.mark.parametrize("n,h,w,cin,cout,k,stride", CASES)
@pytest.mark.parametrize("impl", [0, 1])
def test_conv_matches_torch(cuda, n, h, w, cin, cout, k, stride, impl):
    from onedc_b200 import ops
    x = _mk((n, h, w, cin), cuda, 1)
    wt = _mk((cout, cin, k, k), "cpu", 2, scale=(cin * k * k) ** -0.5).float()
    b = _mk((cout,), "cpu", 3).float()
    cw = ops.ConvW(wt, b, cuda)
    out = ops.igemm(x, cw, stride=stride, impl=impl)
    ref = _ref_conv(x, cw.w[:k * k].float().reshape(k, k, cout, -1).permute(2, 3, 0, 1)[:, :cin].contiguous(), b.to(cuda), k, stride)
    assert out.shape == ref.shape
    _close(out, ref)


def test_two_sources_residual_act_fp32(cuda):
    from onedc_b200 import ops
    n, h, w = 1, 24, 24
    x, x2 = _mk((n, h, w, 320), cuda, 1), _mk((n, h, w, 640), cuda, 2)
    wt = _mk((320, 960, 3, 3), "cpu", 3, scale=(960 * 9) ** -0.5).float()
    b = _mk((320,), "cpu", 4).float()
    res = _mk((n, h, w, 320), cuda, 5)
    cw = ops.ConvW(wt, b, cuda)
    for impl in (0, 1):
        out = ops.igemm(x, cw, x2=x2, act=ops.ACT_LRELU, slope=0.1, res=res, out_dtype=torch.float32, impl=impl)
        ref = F.leaky_relu(_ref_conv(x, cw.w.float().reshape(3, 3, 320, 960).permute(2, 3, 0, 1).contiguous(),
                                     b.to(cuda), 3, 1, x2), 0.1) + res.float()
        _close(out, ref, 5e-3)
    resf = res.float()
    out = ops.igemm(x, cw, x2=x2, act=ops.ACT_SILU, res=resf)
    ref = F.silu(_ref_conv(x, cw.w.float().reshape(3, 3, 320, 960).permute(2, 3, 0, 1).contiguous(), b.to(cuda), 3, 1, x2)) + resf
    _close(out, ref)


def test_channel_slice_views(cuda):
    """inputs / outputs / residuals that are channel slices of wider buffers (zero-copy concat)."""
    from onedc_b200 import ops
    buf = _mk((1, 16, 16, 256), cuda, 1)
    outbuf = torch.zeros((1, 16, 16, 512), device=cuda, dtype=torch.bfloat16)
    wt = _mk((256, 128, 1, 1), "cpu", 2, scale=128 ** -0.5).float()
    cw = ops.ConvW(wt, None, cuda)
    ops.igemm(buf[..., 128:], cw, out=outbuf[..., 256:], res=buf)
    ref = _ref_conv(buf[..., 128:], cw.w.float().reshape(1, 1, 256, 128).permute(2, 3, 0, 1).contiguous(), None, 1, 1) + buf.float()
    _close(outbuf[..., 256:], ref)
    assert outbuf[..., :256].abs().max().item() == 0


@pytest.mark.parametrize("mode", ["pair", "geglu"])
def test_pair_epilogues(cuda, mode):
    from onedc_b200 import ops
    from onedc_b200.lib import EPI_GEGLU, EPI_PAIR_LRELU
    c = 256
    x = _mk((1, 20, 20, c), cuda, 1)
    wt = _mk((4 * c, c), "cpu", 2, scale=c ** -0.5).float()
    b = _mk((4 * c,), "cpu", 3).float()
    wp, bp = ops.pair_permute(wt, b, 256)
    cw = ops.ConvW(wp, bp, cuda, epi=EPI_PAIR_LRELU if mode == "pair" else EPI_GEGLU, bn=256)
    y = F.linear(x.float(), wt.to(torch.bfloat16).float().to(cuda), b.to(cuda))
    a, g = y.chunk(2, dim=-1)
    ref = F.leaky_relu(a, 0.1) + F.leaky_relu(g, 0.01) if mode == "pair" else a * F.gelu(g)
    for impl in (0, 1):
        out = ops.igemm(x, cw, impl=impl)
        assert out.shape[-1] == 2 * c
        _close(out, ref)


def test_pixel_shuffle_store(cuda):
    from onedc_b200 import ops
    from onedc_b200.lib import ST_PIXSHUF
    cin, cout = 128, 256
    x = _mk((2, 12, 12, cin), cuda, 1)
    wt = _mk((4 * cout, cin), "cpu", 2, scale=cin ** -0.5).float()
    b = _mk((4 * cout,), "cpu", 3).float()
    cw = ops.ConvW(*ops.pixshuf_permute(wt, b), cuda)
    y = F.linear(x.float(), wt.to(torch.bfloat16).float().to(cuda), b.to(cuda)).permute(0, 3, 1, 2)
    ref = F.pixel_shuffle(F.leaky_relu(y, 0.01), 2).permute(0, 2, 3, 1)
    for impl in (0, 1):
        out = ops.igemm(x, cw, act=ops.ACT_LRELU, slope=0.01, store=ST_PIXSHUF, ps_c=cout, impl=impl)
        _close(out, ref)


def test_transposed_store_and_batched_b(cuda):
    from onedc_b200 import ops
    from onedc_b200.lib import ST_TRANSPOSED
    b_, s, c = 3, 144, 768
    x = _mk((b_, 1, s, c), cuda, 1)
    wt = _mk((c, c), "cpu", 2, scale=c ** -0.5).float()
    cw = ops.ConvW(wt, None, cuda)
    vT = torch.zeros((b_, c, 152), device=cuda, dtype=torch.bfloat16)
    ops.igemm(x, cw, store=ST_TRANSPOSED, out=vT)
    ref = F.linear(x.float(), wt.to(torch.bfloat16).float().to(cuda))[:, 0].transpose(1, 2)
    _close(vT[:, :, :s], ref)
    assert vT[:, :, s:].abs().max().item() == 0
    # batched B: scores[b] = q[b] k[b]^T
    q, k = _mk((b_, s, c), cuda, 3, 0.2), _mk((b_, s, c), cuda, 4, 0.2)
    sc = torch.zeros((b_, 1, s, 152), device=cuda, dtype=torch.float32)
    ops.igemm(q[:, None], k, out=sc[..., :s], w_batched=True)
    _close(sc[:, 0, :, :s], torch.bmm(q.float(), k.float().transpose(1, 2)), 2e-3)


def test_large_layer_shapes(cuda):
    """the biggest VAE layer shape (256 ch @ 768^2 is 38 GB of MACs; use a 192^2 crop of it) and N=128 tiles."""
    from onedc_b200 import ops
    for c, hw in ((256, 192), (128, 256)):
        x = _mk((1, hw, hw, c), cuda, 1)
        wt = _mk((c, c, 3, 3), "cpu", 2, scale=(9 * c) ** -0.5).float()
        cw = ops.ConvW(wt, None, cuda)
        out = ops.igemm(x, cw)
        ref = _ref_conv(x, cw.w[:9].float().reshape(3, 3, c, c).permute(2, 3, 0, 1).contiguous(), None, 3, 1)
        _close(out, ref)


@pytest.mark.parametrize("h,w,cin,cout,two", [(24, 24, 1280, 1280, False), (12, 12, 1280, 1280, True), (48, 48, 640, 640, False)])
def test_split_k_small_m_layers(cuda, h, w, cin, cout, two):
    """UNet 12x12 / 24x24 / 48x48 shapes take the split-K route (few tiles, long K): same result as the SIMT checker
    and torch, bit-identical run to run (fixed reduction order)."""
    from onedc_b200 import ops
    x = _mk((1, h, w, cin), cuda, 1)
    x2 = _mk((1, h, w, cin), cuda, 5) if two else None
    ct = cin * (2 if two else 1)
    wt = _mk((cout, ct, 3, 3), "cpu", 2, scale=(ct * 9) ** -0.5).float()
    b = _mk((cout,), "cpu", 3).float()
    res = _mk((1, h, w, cout), cuda, 4)
    cw = ops.ConvW(wt, b, cuda)
    out = ops.igemm(x, cw, x2=x2, res=res)
    out2 = ops.igemm(x, cw, x2=x2, res=res)
    assert torch.equal(out, out2)
    ref = _ref_conv(x, cw.w[:9].float().reshape(3, 3, cout, ct).permute(2, 3, 0, 1).contiguous(), b.to(cuda), 3, 1, x2) + res.float()
    _close(out, ref)
    _close(ops.igemm(x, cw, x2=x2, res=res, impl=1), ref)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 12, 20, 1280, 1280), (1, 24, 24, 1920, 640), (3, 9, 7, 1280, 320)])
def test_split_k_cluster_reduction_with_statistics(cuda, n, h, w, cin, cout):
    """Split-K runs as thread-block clusters that reduce over distributed shared memory: several images, ragged tiles (rows
    that no CTA sends), row slices that do not divide 128, bias + residual, and the fused GroupNorm statistics taken from the
    reduced rows must match torch; bit-identical run to run."""
    from onedc_b200 import ops
    x = _mk((n, h, w, cin), cuda, 1)
    wt = _mk((cout, cin, 3, 3), "cpu", 2, scale=(cin * 9) ** -0.5).float()
    b = _mk((cout,), "cpu", 3).float()
    res = _mk((n, h, w, cout), cuda, 4)
    cw = ops.ConvW(wt, b, cuda)
    ops.gn_arena_reset(cuda)
    out = ops.igemm(x, cw, res=res, stats=True)
    ref = _ref_conv(x, cw.w[:9].float().reshape(3, 3, cout, cin).permute(2, 3, 0, 1).contiguous(), b.to(cuda), 3, 1) + res.float()
    _close(out, ref)
    acc = getattr(out, "_gn_acc", None)
    if acc is not None:
        torch.cuda.synchronize()
        ng = cout if out._gn_chan else 32
        got = acc[: n * ng * 2].view(n, ng, 2)
        o = out.float().view(n, h * w, ng, cout // ng)
        want = torch.stack([o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))], dim=-1).double()
        assert torch.allclose(got, want, rtol=3e-3, atol=1e-3 * float(want.abs().max()))
    ops.gn_arena_reset(cuda)
    assert torch.equal(out, ops.igemm(x, cw, res=res, stats=True)), "not deterministic"


@pytest.mark.parametrize("n,h,w,c,co", [(1, 24, 24, 256, 256), (2, 12, 20, 512, 512), (1, 48, 48, 1280, 1280)])
def test_folded_upsample_conv(cuda, n, h, w, c, co):
    """nearest-2x + conv3x3 folded into four 4-tap convs on the low-res input == the unfolded computation."""
    from onedc_b200 import ops
    x = _mk((n, h, w, c), cuda, 1)
    wt = _mk((co, c, 3, 3), "cpu", 2, scale=(c * 9) ** -0.5).float()
    b = _mk((co,), "cpu", 3).float()
    up = ops.UpConv(wt, b, cuda)
    out = up(x)
    xu = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(xu, wt.to(cuda), b.to(cuda), padding=1).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    _close(out, ref)
    # and the unfolded kernel route agrees too
    ref2 = ops.igemm(ops.upsample2x(x), ops.ConvW(wt, b, cuda))
    _close(out, ref2.float())
    cw = up.quads[3]
    o1 = torch.zeros_like(out)
    ops.igemm(x, cw, out=o1, store=ops.ST_QUAD, quad=3, impl=1)
    _close(o1[:, 1::2, 1::2], ref[:, 1::2, 1::2])


@pytest.mark.parametrize("n,h,w,cin,cin2,cout", [
    (1, 40, 36, 64, 0, 96), (1, 96, 96, 256, 0, 256), (2, 50, 30, 128, 0, 128), (1, 17, 9, 192, 0, 64),
    (1, 192, 160, 128, 0, 128), (1, 96, 96, 320, 320, 320), (1, 64, 64, 72, 40, 48), (3, 33, 65, 64, 0, 512)])
def test_column_copy_mode_3x3(cuda, n, h, w, cin, cin2, cout):
    """Stride-1 3x3 convs without split-K run in column-copy mode (16x8 pixel tiles, one A box per tap column, separate
    A / B rings): ragged edges, several tiles per CTA (ring phase wrap), two concatenated sources, bias, residual and
    the fused GroupNorm statistics must all match torch and the SIMT checker."""
    from onedc_b200 import ops
    x = _mk((n, h, w, cin), cuda, 1)
    x2 = _mk((n, h, w, cin2), cuda, 5) if cin2 else None
    ct = cin + cin2
    wt = _mk((cout, ct, 3, 3), "cpu", 2, scale=(ct * 9) ** -0.5).float()
    b = _mk((cout,), "cpu", 3).float()
    res = _mk((n, h, w, cout), cuda, 4)
    cw = ops.ConvW(wt, b, cuda)
    ops.gn_arena_reset(cuda)
    out = ops.igemm(x, cw, x2=x2, res=res, stats=True)
    ref = _ref_conv(x, cw.w[:9].float().reshape(3, 3, cout, -1).permute(2, 3, 0, 1)[:, :ct].contiguous(), b.to(cuda), 3, 1, x2) + res.float()
    _close(out, ref)
    _close(ops.igemm(x, cw, x2=x2, res=res, impl=1), ref)
    assert torch.equal(out, ops.igemm(x, cw, x2=x2, res=res)), "not deterministic"
    acc = getattr(out, "_gn_acc", None)
    if acc is not None:                          # fused statistics: per (image, group) sum and sum of squares
        torch.cuda.synchronize()
        ng = cout if out._gn_chan else 32
        got = acc[: n * ng * 2].view(n, ng, 2)
        o = out.float().view(n, h * w, ng, cout // ng)
        want = torch.stack([o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))], dim=-1).double()
        assert torch.allclose(got, want, rtol=2e-3, atol=1e-2 * float(want.abs().max()) * 1e-2)


@pytest.mark.parametrize("n,h,w,cin,cout,two,f32", [(1, 384, 384, 128, 128, False, False), (2, 100, 210, 256, 128, False, False),
                                                    (1, 256, 256, 64, 96, True, False), (1, 320, 200, 128, 64, False, True),   # fp32 output: regular tile
                                                    # cout == cin: the residual goes through the tensor core (identity tap)
                                                    (2, 100, 210, 128, 128, False, False), (1, 300, 203, 64, 64, False, False),
                                                    (1, 256, 250, 96, 96, False, False)])
def test_transposed_column_copy_mode(cuda, n, h, w, cin, cout, two, f32):
    """Narrow outputs (cout <= 128) on many tiles run the transposed tile: weights as the M operand, 32x8 pixels as
    N = 256, one bf16 store per lane.  Bias, residual, ragged edges, two sources, fp32 output and the fused GroupNorm
    statistics must match torch / the SIMT checker."""
    from onedc_b200 import ops
    x = _mk((n, h, w, cin), cuda, 1)
    x2 = _mk((n, h, w, cin), cuda, 5) if two else None
    ct = cin * (2 if two else 1)
    wt = _mk((cout, ct, 3, 3), "cpu", 2, scale=(ct * 9) ** -0.5).float()
    b = _mk((cout,), "cpu", 3).float()
    res = _mk((n, h, w, cout), cuda, 4)
    cw = ops.ConvW(wt, b, cuda)
    ops.gn_arena_reset(cuda)
    od = torch.float32 if f32 else torch.bfloat16
    out = ops.igemm(x, cw, x2=x2, res=res, stats=True, out_dtype=od)
    ref = _ref_conv(x, cw.w[:9].float().reshape(3, 3, cout, -1).permute(2, 3, 0, 1)[:, :ct].contiguous(), b.to(cuda), 3, 1, x2) + res.float()
    _close(out, ref)
    _close(ops.igemm(x, cw, x2=x2, res=res, impl=1, out_dtype=od), ref)
    assert torch.equal(out, ops.igemm(x, cw, x2=x2, res=res, out_dtype=od)), "not deterministic"
    acc = getattr(out, "_gn_acc", None)
    assert acc is not None
    torch.cuda.synchronize()
    ng = cout if out._gn_chan else 32
    got = acc[: n * ng * 2].view(n, ng, 2)
    o = out.float().view(n, h * w, ng, cout // ng)
    want = torch.stack([o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))], dim=-1).double()
    assert torch.allclose(got, want, rtol=3e-3, atol=1e-4 * float(want.abs().max()))


@pytest.mark.parametrize("n,h,w,cin,cout,planar,with_res", [(1, 96, 96, 320, 4, False, True), (2, 40, 56, 128, 3, True, False),
                                                          (1, 33, 17, 64, 4, False, False), (1, 128, 128, 128, 3, True, False)])
def test_tap_expanded_small_cout_conv(cuda, n, h, w, cin, cout, planar, with_res):
    """3x3 convs with 3-4 output channels run as a 1x1 GEMM over 9*cout tap-expanded columns + a 9-neighbour gather
    (ops.TapConv3x3): same result as the fp32 torch convolution of the bf16-rounded operands, zero padding included."""
    from onedc_b200 import ops
    x = _mk((n, h, w, cin), cuda, 1)
    wt = _mk((cout, cin, 3, 3), "cpu", 2, scale=(cin * 9) ** -0.5).float()
    b = _mk((cout,), "cpu", 3).float()
    res = torch.randn((n, h, w, cout), generator=torch.Generator().manual_seed(4)).to(cuda) if with_res else None
    tc = ops.TapConv3x3(wt, b, cuda)
    out = tc(x, res=res, planar=planar)
    ref = _ref_conv(x, wt.to(torch.bfloat16).float().to(cuda), b.to(cuda), 3, 1)
    if with_res:
        ref = ref + res
    if planar:
        assert out.shape == (n, cout, h * w)
        out = out.view(n, cout, h, w).permute(0, 2, 3, 1)
    assert out.dtype == torch.float32
    _close(out, ref, tol=2e-3)
    assert torch.equal(tc(x, res=res, planar=planar).reshape(-1), (out if not planar else out.permute(0, 3, 1, 2)).reshape(-1))


@pytest.mark.parametrize("h,w,cin,cout,k", [(48, 48, 128, 128, 3), (24, 24, 128, 128, 3), (48, 48, 256, 1024, 1), (12, 12, 128, 512, 1)])
def test_deterministic_plan_is_batch_independent(cuda, h, w, cin, cout, k):
    """Layers that feed the entropy parameters are launched with det=True: no split-K, no column-copy / transposed tile,
    so the fp32 summation order of every output element -- hence every bf16 bit of the scales -- is the same whether the
    image is decoded alone or inside a batch large enough to flip the default plan (n_img * tiles > SM count)."""
    from onedc_b200 import ops
    nb = 12
    x = _mk((nb, h, w, cin), cuda, 11)
    wt = _mk((cout, cin, k, k), "cpu", 12, scale=(cin * k * k) ** -0.5).float()
    cw = ops.ConvW(wt, _mk((cout,), "cpu", 13).float(), cuda)
    batched = ops.igemm(x, cw, det=True)
    for i in (0, 5, nb - 1):
        single = ops.igemm(x[i:i + 1].contiguous(), cw, det=True)
        assert torch.equal(single[0], batched[i]), "deterministic plan: batched result differs from the single-image one"
