"""GPU parity of the tcgen05 flash-attention kernel and of the unfused (GEMM-softmax-GEMM) route against
torch SDPA in fp32 on the same bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(shape, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)


def _ref(q, k, v, heads, d, scale=None):
    b, sq, _ = q.shape
    qh = q.float().view(b, sq, heads, d).transpose(1, 2)
    kh = k.float().view(b, -1, heads, d).transpose(1, 2)
    vh = v.float().view(b, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(qh, kh, vh, scale=scale)
    return o.transpose(1, 2).reshape(b, sq, heads * d)


CASES = [  # batch, heads, d, sq, skv
    (1, 8, 40, 1024, 1024),
    (1, 8, 40, 1024, 16),
    (2, 8, 40, 256, 144),
    (1, 8, 80, 576, 576),
    (1, 8, 80, 256, 96),
    (1, 8, 160, 144, 144),
    (1, 8, 160, 64, 16),
    (1, 8, 160, 36, 36),
    (1, 2, 64, 300, 200),
]


@pytest.mark.parametrize("b,heads,d,sq,skv", CASES)
@pytest.mark.parametrize("impl", [0, 1])
def test_flash_matches_sdpa(cuda, b, heads, d, sq, skv, impl):
    from onedc_b200 import ops
    c = heads * d
    # fused projection layouts: q inside [b,sq,3c], k/v inside [b,skv,2c]
    qkv = _mk((b, sq, 3 * c), cuda, 1)
    kv = _mk((b, skv, 2 * c), cuda, 2)
    out = torch.zeros((b, sq, c), device=cuda, dtype=torch.bfloat16)
    ops.attention(qkv[:, :, :c], kv[:, :, :c], kv[:, :, c:], out, heads, d, impl=impl)
    ref = _ref(qkv[:, :, :c], kv[:, :, :c], kv[:, :, c:], heads, d)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, f"max abs err {err}"


def test_flash_long_sequence(cuda):
    from onedc_b200 import ops
    b, heads, d, s = 1, 8, 40, 9216
    c = heads * d
    qkv = _mk((b, s, 3 * c), cuda, 3)
    out = torch.empty((b, s, c), device=cuda, dtype=torch.bfloat16)
    ops.attention(qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:], out, heads, d)
    ref = _ref(qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:], heads, d)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, f"max abs err {err}"


@pytest.mark.parametrize("b,heads,d,s", [(1, 1, 768, 144), (4, 1, 512, 256), (1, 8, 40, 200)])
def test_unfused_matches_sdpa(cuda, b, heads, d, s):
    from onedc_b200 import ops
    c = heads * d
    q, k, v = _mk((b, s, c), cuda, 1, 0.3), _mk((b, s, c), cuda, 2, 0.3), _mk((b, s, c), cuda, 3)
    lpad = (s + 7) // 8 * 8
    vT = torch.zeros((b, c, lpad), device=cuda, dtype=torch.bfloat16)
    vT[:, :, :s] = v.transpose(1, 2)
    out = torch.zeros((b, s, c), device=cuda, dtype=torch.bfloat16)
    scale = float(c) ** -0.5 if heads == 1 else None
    ops.attention_unfused(q, k, vT, out, heads, d, scale=scale)
    ref = _ref(q, k, v, heads, d, scale=scale)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, f"max abs err {err}"


@pytest.mark.parametrize("plan", [(1, 1), (2, 3), (1, 3), (2, 4)])
@pytest.mark.parametrize("b,heads,d,sq,skv", [(1, 8, 40, 2304, 5000), (2, 8, 40, 1152, 1152), (1, 8, 64, 2304, 1100), (1, 8, 80, 576, 2304)])
def test_flash_launch_plans(cuda, b, heads, d, sq, skv, plan):
    """The non-default launch plans: one S buffer / three CTAs per SM, and the key range split over several CTAs per
    query tile (fp32 partial O + running max / sum, merged by attention_merge_kernel): ragged last block, batch > 1,
    d = 64 and a head_dim that cannot take the one-buffer variant."""
    from onedc_b200 import lib, ops
    L = lib.load()
    L.onedc_attention_set_plan(*plan)
    try:
        if plan[1] > 1:
            assert L.onedc_attention_ws_floats(b, heads, d, sq, skv) > 0
        c = heads * d
        q = _mk((b, sq, c), cuda, 1)
        kv = _mk((b, skv, 2 * c), cuda, 2)
        out = torch.zeros((b, sq, c), device=cuda, dtype=torch.bfloat16)
        ops.attention(q, kv[:, :, :c], kv[:, :, c:], out, heads, d)
        ref = _ref(q, kv[:, :, :c], kv[:, :, c:], heads, d)
        err = (out.float() - ref).abs().max().item()
        assert err < 2e-2, f"max abs err {err}"
        out2 = torch.zeros_like(out)
        ops.attention(q, kv[:, :, :c], kv[:, :, c:], out2, heads, d)
        assert torch.equal(out, out2)
    finally:
        L.onedc_attention_set_plan(0, 0)


@pytest.mark.parametrize("d,skv", [(40, 1024), (80, 640), (160, 300)])
def test_flash_large_scores_moving_maximum(cuda, d, skv):
    """Large scores (|s| up to ~60 after scaling) with keys ordered so that the row maximum keeps growing from block to
    block for half of the rows and falls for the other half: many rescales of O, running maxima far below / above the
    final one."""
    from onedc_b200 import ops
    b, heads, sq = 1, 4, 256
    c = heads * d
    g = torch.Generator().manual_seed(5)
    q = torch.randn((b, sq, c), generator=g)
    k = torch.randn((b, skv, c), generator=g) * 0.3
    ramp = torch.linspace(-1.0, 1.0, skv)[None, :, None]                    # keys later in the sequence align more with u
    u = torch.randn((1, 1, c), generator=g)
    k = k + 3.0 * ramp * u
    q[:, : sq // 2] += 1.5 * u                                               # rows that follow the ramp
    q[:, sq // 2:] -= 1.5 * u                                                # rows that oppose it
    v = torch.randn((b, skv, c), generator=g)
    q, k, v = (t.to(torch.bfloat16).to(cuda) for t in (q, k, v))
    out = torch.zeros((b, sq, c), device=cuda, dtype=torch.bfloat16)
    ops.attention(q, k, v, out, heads, d)
    ref = _ref(q, k, v, heads, d)
    smax = float((q.float().view(b, sq, heads, d).transpose(1, 2) @ k.float().view(b, skv, heads, d).transpose(1, 2).transpose(2, 3)).abs().max()) * d ** -0.5
    assert smax > 20, f"test should produce large scores (got {smax})"
    err = (out.float() - ref).abs().max().item()
    assert err < 3e-2, f"max abs err {err}"
    chk = torch.zeros_like(out)
    ops.attention(q, k, v, chk, heads, d, impl=1)                           # SIMT checker agrees too
    assert (out.float() - chk.float()).abs().max().item() < 3e-2
