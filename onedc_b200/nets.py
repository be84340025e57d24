"""Networks of the decode hot path, expressed as sequences of libonedc_b200 kernel launches.

Each class is built from a reference-named state dict (onedc_b200/weights.py), repacks the weights into
kernel layouts once (LoRA merged, time embedding folded into conv biases, output channels permuted for the
fused PixelShuffle / GEGLU / ConvFFN3 epilogues) and then only launches kernels.  Activations are NHWC bf16.

Reference structure followed (paths under /root/reference/src):
  DepthConvBlock4 / ResidualBlockUpsample     modules/dcvc.py:183-208,242-265,353-367,424-433
  ResnetBlock / AttnBlock                     modules/vqgan/blocks.py:15-107
  HyperDecoder / SemanticAdaptor / Decoder    models/sd15_onedc_codec_stage1/codec_module.py:88-181
  UNet (diffusers SD1.5 + LoRA + patches)     models/sd15_onedc_codec_stage1/decoder_unet.py:14-29,32-305,331-400
  VAE decoder with windowed mid attention     modules/vae/autoencoders_patch_attn.py:9-81
"""
import math

import torch

from . import ops
from .ops import (ACT_LRELU, ACT_NONE, ConvW, DepthwiseW, GroupNorm, LayerNorm, TapConv3x3, UpConv, igemm, pair_permute,
                  pixshuf_permute)
from .lib import EPI_GEGLU, EPI_PAIR_LRELU, ST_PIXSHUF, ST_TRANSPOSED
from .weights import LazyNet, merge_lora, unet_spec, vae_spec


def _wb(sd, name):
    """(weight fp32, bias or None) with LoRA merged when the layer is peft-wrapped."""
    w, b = merge_lora(sd, name)
    return w.float(), (None if b is None else b.float())


# ============================================================================ codec blocks
class DCB4:
    """DepthConvBlock4 = DepthConv (1x1+LReLU, dw3x3, 1x1 [+1x1 adaptor]) + ConvFFN3.
    5 launches: igemm, dwconv, igemm (adaptor folded in as a second K source), igemm (pair epilogue), igemm."""

    def __init__(self, sd, p, cin, cout, dev, out_stats=False, det=False):
        self.cin, self.cout = cin, cout
        self.out_stats = out_stats            # output feeds a GroupNorm: fuse its statistics into the last epilogue
        self.det = det                        # entropy path: batch-independent launch plan (bit-identical scales)
        w, b = _wb(sd, p + ".block.0.conv1.0")
        self.conv1 = ConvW(w, b, dev)
        self.dw = DepthwiseW(sd[p + ".block.0.depth_conv.weight"].float(), sd[p + ".block.0.depth_conv.bias"].float(), dev)
        w2, b2 = _wb(sd, p + ".block.0.conv2")
        if cin != cout:
            wa, ba = _wb(sd, p + ".block.0.adaptor")
            w2, b2 = torch.cat([w2, wa], dim=1), b2 + ba          # out = [dw_out | x] . [W2 | Wa]^T + (b2 + ba)
        self.conv2 = ConvW(w2, b2, dev)
        wf, bf = pair_permute(*_wb(sd, p + ".block.1.conv"), bn=256)
        self.ffn_in = ConvW(wf, bf, dev, epi=EPI_PAIR_LRELU, bn=256)
        self.ffn_out = ConvW(*_wb(sd, p + ".block.1.conv_out"), dev)

    def __call__(self, x, out=None):
        det = self.det
        t = igemm(x, self.conv1, act=ACT_LRELU, slope=0.01, det=det)
        t = ops.dwconv3x3(t, self.dw)
        if self.cin != self.cout:
            h = igemm(t, self.conv2, x2=x, det=det)
        else:
            h = igemm(t, self.conv2, res=x, det=det)
        f = igemm(h, self.ffn_in, det=det)
        return igemm(f, self.ffn_out, res=h, out=out, stats=self.out_stats, det=det)


class RBU:
    """ResidualBlockUpsample: two 1x1 -> PixelShuffle(2) branches (shuffle fused into the store), LReLU,
    3x3 conv + LReLU(0.1) + identity."""

    def __init__(self, sd, p, cin, cout, dev, det=False):
        self.cout, self.det = cout, det
        self.subpel = ConvW(*pixshuf_permute(*_wb(sd, p + ".subpel_conv.0")), dev)
        self.up = ConvW(*pixshuf_permute(*_wb(sd, p + ".upsample.0")), dev)
        self.conv = ConvW(*_wb(sd, p + ".conv"), dev)

    def __call__(self, x, out=None):
        det = self.det
        a = igemm(x, self.subpel, act=ACT_LRELU, slope=0.01, store=ST_PIXSHUF, ps_c=self.cout, det=det)
        idt = igemm(x, self.up, store=ST_PIXSHUF, ps_c=self.cout, det=det)
        return igemm(a, self.conv, act=ACT_LRELU, slope=0.1, res=idt, out=out, det=det)


class VQRes:
    def __init__(self, sd, p, c, dev):
        self.n1 = GroupNorm(sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-6, device=dev)
        self.c1 = ConvW(sd[p + ".conv1.weight"].float(), None, dev)
        self.n2 = GroupNorm(sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-6, device=dev)
        self.c2 = ConvW(sd[p + ".conv2.weight"].float(), None, dev)

    def __call__(self, x, out=None):
        h = igemm(self.n1(x), self.c1, stats=True)
        return igemm(self.n2(h), self.c2, res=x, out=out, stats=True)


def _pad8(n):
    return (n + 7) // 8 * 8


class VQAttn:
    """vqgan AttnBlock: single head over all H*W tokens, d = C (768), scale C^-0.5."""

    def __init__(self, sd, p, c, dev):
        self.c = c
        self.norm = GroupNorm(sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6, device=dev)
        wq, bq = _wb(sd, p + ".q")
        wk, bk = _wb(sd, p + ".k")
        self.qk = ConvW(torch.cat([wq, wk], 0), torch.cat([bq, bk], 0), dev)
        self.v = ConvW(*_wb(sd, p + ".v"), dev)
        self.proj = ConvW(*_wb(sd, p + ".proj_out"), dev)

    def __call__(self, x):
        n, h, w, c = x.shape
        s = h * w
        hn = self.norm(x, silu=False)
        qk = igemm(hn, self.qk).view(n, s, 2 * c)
        vT = torch.zeros((n, c, _pad8(s)), device=x.device, dtype=torch.bfloat16)
        igemm(hn, self.v, store=ST_TRANSPOSED, out=vT)
        o = torch.empty((n, s, c), device=x.device, dtype=torch.bfloat16)
        ops.attention_unfused(qk[:, :, :c], qk[:, :, c:], vT, o, heads=1, head_dim=c, scale=float(c) ** -0.5)
        return igemm(o.view(n, h, w, c), self.proj, res=x, stats=True)


class HyperSynthesis:
    """HyperDecoder + y_prior_fusion: z codes -> (common_params [N,h16,w16,256], z_semantic [N,hz,wz,128])."""

    def __init__(self, sd, dev):
        self.feat_in = ConvW(*_wb(sd, "hyper_dec.feat_in.0"), dev)
        p = "hyper_dec.to_entropy."
        # det=True everywhere: these layers produce the entropy parameters (scales | means)
        self.seq = [DCB4(sd, p + "0", 128, 128, dev, det=True), RBU(sd, p + "1", 128, 128, dev, det=True),
                    DCB4(sd, p + "2", 128, 128, dev, det=True), RBU(sd, p + "3", 128, 128, dev, det=True),
                    DCB4(sd, p + "4", 128, 128, dev, det=True),
                    DCB4(sd, "y_prior_fusion.0", 128, 256, dev, det=True), DCB4(sd, "y_prior_fusion.1", 256, 256, dev, det=True)]

    def hyper_dec(self, codes):
        """HyperDecoder.forward (codec_module.py:162-166): FSQ codes [N,hz,wz,8] -> (z_entropy, z_semantic)"""
        z_sem = igemm(codes, self.feat_in, act=ACT_LRELU, slope=0.01, det=True)
        t = z_sem
        for m in self.seq[:5]:
            t = m(t)
        return t, z_sem

    def y_prior_fusion(self, t):
        for m in self.seq[5:]:
            t = m(t)
        return t

    def __call__(self, z_idx):
        t, z_sem = self.hyper_dec(ops.fsq_codes(z_idx))
        return self.y_prior_fusion(t), z_sem


class SpatialPrior:
    """y_spatial_prior_reduction + the three adaptors + the shared 3-block prior net.
    `params` is the [N,h,w,256] buffer whose first 128 channels hold y_hat_so_far and whose last 128 hold
    the reduced common params (the reference's torch.cat, compression_model.py:386, done by layout)."""

    def __init__(self, sd, dev):
        self.reduction = ConvW(*_wb(sd, "y_spatial_prior_reduction"), dev)
        self.adaptors = [None] + [DCB4(sd, f"y_spatial_prior_adaptor_{i}", 256, 256, dev, det=True) for i in (1, 2, 3)]
        self.prior = [DCB4(sd, f"y_spatial_prior.{i}", 256, 256, dev, det=True) for i in range(3)]

    def reduce(self, common, out=None):
        """y_spatial_prior_reduction (1x1 conv 256 -> 128)"""
        return igemm(common, self.reduction, out=out, det=True)

    def init_params(self, common):
        n, h, w, _ = common.shape
        params = torch.empty((n, h, w, 256), device=common.device, dtype=torch.bfloat16)
        self.reduce(common, out=params[..., 128:])
        return params

    def step(self, k, params):
        t = self.adaptors[k](params)
        for m in self.prior:
            t = m(t)
        return t                                   # [N,h,w,256] = scales | means


class SemanticAdaptorNet:
    def __init__(self, sd, dev):
        p = "semantic_adaptor.to_semantic."
        self.seq = [DCB4(sd, p + "0", 128, 768, dev, out_stats=True), VQRes(sd, p + "1", 768, dev), VQAttn(sd, p + "2", 768, dev),
                    VQAttn(sd, p + "3", 768, dev), VQRes(sd, p + "4", 768, dev), VQAttn(sd, p + "5", 768, dev),
                    VQAttn(sd, p + "6", 768, dev), DCB4(sd, p + "7", 768, 768, dev)]

    def __call__(self, z_sem):
        ops.gn_arena_reset(z_sem.device)      # one zeroing of the fused-GroupNorm accumulators per pass
        t = z_sem
        for m in self.seq:
            t = m(t)
        return t


class LatentSynthesisNet:
    """Decoder g_s (codec_module.py:88-116): y_hat, y_semantic -> x_hat [N,h8,w8,320]."""

    def __init__(self, sd, dev):
        self.tc = [DCB4(sd, "dec.trans_coding.0", 128, 512, dev), DCB4(sd, "dec.trans_coding.1", 512, 512, dev, out_stats=True)]
        self.res16 = [VQRes(sd, f"dec.blocks.{i}", 512, dev) for i in range(3)]
        self.up = ConvW(*pixshuf_permute(*_wb(sd, "dec.blocks.3")), dev)
        self.up_conv = ConvW(*_wb(sd, "dec.blocks.5"), dev)
        self.res8 = [VQRes(sd, f"dec.blocks.{i}", 256, dev) for i in (6, 7, 8)]
        self.sem = [RBU(sd, "dec.sem_up.0", 768, 512, dev), DCB4(sd, "dec.sem_up.1", 512, 512, dev),
                    RBU(sd, "dec.sem_up.2", 512, 256, dev), DCB4(sd, "dec.sem_up.3", 256, 256, dev),
                    RBU(sd, "dec.sem_up.4", 256, 256, dev)]
        self.conv_out = DCB4(sd, "dec.conv_out", 512, 320, dev, out_stats=True)

    def alloc_cat(self, n, h16, w16, device):
        return torch.empty((n, 2 * h16, 2 * w16, 512), device=device, dtype=torch.bfloat16)

    def sem_path(self, y_sem, cat):
        """semantic branch: independent of y, so it can run while the host decodes the y stream"""
        s = y_sem
        for m in self.sem[:-1]:
            s = m(s)
        self.sem[-1](s, out=cat[..., 256:])

    def main_path(self, y_hat, cat):
        ops.gn_arena_reset(y_hat.device)      # covers g_s, the UNet and the VAE decoder of this pass
        t = y_hat
        for m in self.tc + self.res16:
            t = m(t)
        t = igemm(t, self.up, store=ST_PIXSHUF, ps_c=512)
        t = igemm(t, self.up_conv, stats=True)
        t = self.res8[0](t)
        t = self.res8[1](t)
        self.res8[2](t, out=cat[..., :256])
        return self.conv_out(cat)

    def __call__(self, y_hat, y_sem):
        n, h16, w16, _ = y_hat.shape
        cat = self.alloc_cat(n, h16, w16, y_hat.device)
        self.sem_path(y_sem, cat)
        return self.main_path(y_hat, cat)


# ============================================================================ UNet
def sinusoidal_timestep(t, dim=320):
    half = dim // 2
    e = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    a = torch.tensor([float(t)])[:, None] * e[None, :]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


class UNetRes:
    def __init__(self, sd, p, cin, cout, emb_act, dev):
        self.cin, self.cout = cin, cout
        self.n1 = GroupNorm(sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5, device=dev)
        w1, b1 = _wb(sd, p + ".conv1")
        wt, bt = _wb(sd, p + ".time_emb_proj")
        b1 = b1 + (emb_act @ wt.t() + bt).reshape(-1)           # constant timestep: fold temb into the bias
        self.c1 = ConvW(w1, b1, dev)
        self.n2 = GroupNorm(sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5, device=dev)
        self.c2 = ConvW(*_wb(sd, p + ".conv2"), dev)
        self.sc = ConvW(*_wb(sd, p + ".conv_shortcut"), dev) if cin != cout else None

    def __call__(self, x, skip=None):
        h = igemm(self.n1(x, skip), self.c1, stats=True)
        res = x if self.sc is None else igemm(x, self.sc, x2=skip)
        return igemm(self.n2(h), self.c2, res=res, stats=True)


class UNetTransformer:
    def __init__(self, sd, p, c, dev, heads=8):
        self.c, self.heads, self.d = c, heads, c // heads
        self.norm = GroupNorm(sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6, device=dev)
        self.proj_in = ConvW(*_wb(sd, p + ".proj_in"), dev)
        t = p + ".transformer_blocks.0"
        self.ln1 = LayerNorm(sd[t + ".norm1.weight"], sd[t + ".norm1.bias"], device=dev)
        self.ln2 = LayerNorm(sd[t + ".norm2.weight"], sd[t + ".norm2.bias"], device=dev)
        self.ln3 = LayerNorm(sd[t + ".norm3.weight"], sd[t + ".norm3.bias"], device=dev)
        wq, _ = _wb(sd, t + ".attn1.to_q")
        wk, _ = _wb(sd, t + ".attn1.to_k")
        wv, _ = _wb(sd, t + ".attn1.to_v")
        self.qkv1 = ConvW(torch.cat([wq, wk, wv], 0), None, dev)
        self.v1 = ConvW(wv, None, dev)
        self.out1 = ConvW(*_wb(sd, t + ".attn1.to_out.0"), dev)
        self.q2 = ConvW(_wb(sd, t + ".attn2.to_q")[0], None, dev)
        wk2, _ = _wb(sd, t + ".attn2.to_k")
        wv2, _ = _wb(sd, t + ".attn2.to_v")
        self.kv2 = ConvW(torch.cat([wk2, wv2], 0), None, dev)
        self.v2 = ConvW(wv2, None, dev)
        self.out2 = ConvW(*_wb(sd, t + ".attn2.to_out.0"), dev)
        wg, bg = pair_permute(*_wb(sd, t + ".ff.net.0.proj"), bn=256)
        self.geglu = ConvW(wg, bg, dev, epi=EPI_GEGLU, bn=256)
        self.ff_out = ConvW(*_wb(sd, t + ".ff.net.2"), dev)
        self.proj_out = ConvW(*_wb(sd, p + ".proj_out"), dev)

    def _attend(self, q, k, v, x_for_vT, v_w, b, s):
        c = self.c
        o = torch.empty((b, s, c), device=q.device, dtype=torch.bfloat16)
        if ops.ATTN == "flash":
            ops.attention(q, k, v, o, self.heads, self.d)
        else:
            L = x_for_vT.shape[-2]
            vT = torch.zeros((b, c, _pad8(L)), device=q.device, dtype=torch.bfloat16)
            igemm(x_for_vT, v_w, store=ST_TRANSPOSED, out=vT)
            ops.attention_unfused(q, k, vT, o, self.heads, self.d)
        return o

    def project_ctx(self, ctx):
        """K | V of the cross attention: depends on the hyperprior tokens only, so it can be computed ahead of the UNet
        (GraphedDecoder does it on the semantic side stream, under the host rANS)."""
        b, Lc, cc = ctx.shape
        return igemm(ctx.view(b, 1, Lc, cc), self.kv2).view(b, Lc, 2 * self.c)

    def __call__(self, x, ctx, kv=None):
        b, h, w, c = x.shape
        s = h * w
        t = igemm(self.norm(x, silu=False), self.proj_in).view(b, s, c)
        # self attention
        n1 = self.ln1(t)
        qkv = igemm(n1.view(b, 1, s, c), self.qkv1).view(b, s, 3 * c)
        o = self._attend(qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:], n1.view(b, 1, s, c), self.v1, b, s)
        t = igemm(o.view(b, 1, s, c), self.out1, res=t.view(b, 1, s, c)).view(b, s, c)
        # cross attention on the hyperprior tokens
        n2 = self.ln2(t)
        q = igemm(n2.view(b, 1, s, c), self.q2).view(b, s, c)
        Lc = ctx.shape[1]
        if kv is None:
            kv = self.project_ctx(ctx)
        o = self._attend(q, kv[:, :, :c], kv[:, :, c:], ctx.view(b, 1, Lc, ctx.shape[2]), self.v2, b, s)
        t = igemm(o.view(b, 1, s, c), self.out2, res=t.view(b, 1, s, c)).view(b, s, c)
        # feed forward (GEGLU fused into the first GEMM's epilogue)
        n3 = self.ln3(t)
        g = igemm(n3.view(b, 1, s, c), self.geglu)
        t = igemm(g, self.ff_out, res=t.view(b, 1, s, c))
        return igemm(t.view(b, h, w, c), self.proj_out, res=x, stats=True)


class UNet(LazyNet):
    """One-step SD1.5 UNet at a fixed timestep: (x_hat, ctx tokens) -> (eps fp32, reduced fp32), NHWC, 4 ch.
    `load_state_dict` takes the reference's `model.safetensors` keys (peft-wrapped diffusers UNet, weights.unet_spec)."""

    CH = (320, 640, 1280, 1280)

    def __init__(self, sd=None, dev="cuda", timestep=999):
        self.dev, self.timestep = torch.device(dev), timestep
        self._lazy_init(sd)

    def _spec(self):
        return unet_spec()

    def _build(self, sd):
        dev, timestep = self.dev, self.timestep
        f = lambda n: sd[n].float()
        temb = sinusoidal_timestep(timestep)
        emb = torch.nn.functional.silu(temb @ f("time_embedding.linear_1.weight").t() + f("time_embedding.linear_1.bias"))
        emb = emb @ f("time_embedding.linear_2.weight").t() + f("time_embedding.linear_2.bias")
        emb_act = torch.nn.functional.silu(emb)                 # every resnet applies SiLU before its projection
        self.conv_in = ConvW(*_wb(sd, "conv_in"), dev)
        ch = self.CH
        self.down = []
        cin = ch[0]
        for i, c in enumerate(ch):
            blk = {"res": [UNetRes(sd, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c, emb_act, dev) for j in range(2)]}
            if i < 3:
                blk["attn"] = [UNetTransformer(sd, f"down_blocks.{i}.attentions.{j}", c, dev) for j in range(2)]
                blk["down"] = ConvW(*_wb(sd, f"down_blocks.{i}.downsamplers.0.conv"), dev)
            self.down.append(blk)
            cin = c
        self.mid_res = [UNetRes(sd, f"mid_block.resnets.{j}", 1280, 1280, emb_act, dev) for j in range(2)]
        self.mid_attn = UNetTransformer(sd, "mid_block.attentions.0", 1280, dev)
        skips = [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
        self.up = []
        prev = 1280
        for i, c in enumerate(reversed(ch)):
            blk = {"res": []}
            for j in range(3):
                blk["res"].append(UNetRes(sd, f"up_blocks.{i}.resnets.{j}", prev + skips.pop(), c, emb_act, dev))
                prev = c
            if i > 0:
                blk["attn"] = [UNetTransformer(sd, f"up_blocks.{i}.attentions.{j}", c, dev) for j in range(3)]
            if i < 3:
                blk["up"] = UpConv(*_wb(sd, f"up_blocks.{i}.upsamplers.0.conv"), dev)
            self.up.append(blk)
        self.norm_out = GroupNorm(sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], 1e-5, device=dev)
        self.conv_out = TapConv3x3(*_wb(sd, "conv_out"), dev)
        p = "vae_reduction."
        self.vr_n1 = GroupNorm(sd[p + "blocks.0.weight"], sd[p + "blocks.0.bias"], 1e-6, device=dev)
        self.vr_c1 = ConvW(*_wb(sd, p + "blocks.2"), dev)
        self.vr_n2 = GroupNorm(sd[p + "blocks.3.weight"], sd[p + "blocks.3.bias"], 1e-6, device=dev)
        self.vr_c2 = TapConv3x3(*_wb(sd, p + "blocks.5"), dev)
        self.vr_sc = ConvW(*_wb(sd, p + "short_cut"), dev)

    def transformers(self):
        """the 16 transformer blocks in call order"""
        self._ensure()
        out = []
        for blk in self.down:
            out += blk.get("attn", [])
        out.append(self.mid_attn)
        for blk in self.up:
            out += blk.get("attn", [])
        return out

    def project_ctx(self, ctx):
        """cross-attention K | V of every transformer block (ctx-only work), in call order"""
        return [t.project_ctx(ctx) for t in self.transformers()]

    def __call__(self, x, ctx, kvs=None):
        self._ensure()
        f32 = torch.float32
        kvs = iter(kvs) if kvs is not None else None
        nkv = lambda: next(kvs) if kvs is not None else None
        # vae_reduction (fp32 outputs: this feeds the x0 formula, amplified by 1/sqrt(alpha_bar) ~ 14.6)
        sc = igemm(x, self.vr_sc, out_dtype=f32)
        r = igemm(self.vr_n1(x), self.vr_c1, stats=True)
        reduced = self.vr_c2(self.vr_n2(r), res=sc)
        h = igemm(x, self.conv_in, stats=True)
        stack = [h]
        for blk in self.down:
            for j, r_ in enumerate(blk["res"]):
                h = r_(h)
                if "attn" in blk:
                    h = blk["attn"][j](h, ctx, nkv())
                stack.append(h)
            if "down" in blk:
                h = igemm(h, blk["down"], stride=2, stats=True)
                stack.append(h)
        h = self.mid_res[0](h)
        h = self.mid_attn(h, ctx, nkv())
        h = self.mid_res[1](h)
        for blk in self.up:
            for j, r_ in enumerate(blk["res"]):
                h = r_(h, stack.pop())
                if "attn" in blk:
                    h = blk["attn"][j](h, ctx, nkv())
            if "up" in blk:
                h = blk["up"](h)
        eps = self.conv_out(self.norm_out(h))
        return eps, reduced


def alphas_cumprod_sd15():
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


# ============================================================================ VAE decoder
class VAERes:
    def __init__(self, sd, p, cin, cout, dev):
        self.n1 = GroupNorm(sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-6, device=dev)
        self.c1 = ConvW(*_wb(sd, p + ".conv1"), dev)
        self.n2 = GroupNorm(sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-6, device=dev)
        self.c2 = ConvW(*_wb(sd, p + ".conv2"), dev)
        self.sc = ConvW(*_wb(sd, p + ".conv_shortcut"), dev) if cin != cout else None

    def __call__(self, x):
        h = igemm(self.n1(x), self.c1, stats=True)
        res = x if self.sc is None else igemm(x, self.sc)
        return igemm(self.n2(h), self.c2, res=res, stats=True)


class VAEWindowAttention:
    """Mid-block attention, one head of d=512, applied independently to 16x16 latent windows
    (autoencoders_patch_attn.py:20-29): one batched launch sequence over all windows."""

    def __init__(self, sd, p, c, win, dev):
        self.c, self.win = c, win
        self._valid = {}
        self.norm = GroupNorm(sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], 1e-6, device=dev)
        wq, bq = _wb(sd, p + ".to_q")
        wk, bk = _wb(sd, p + ".to_k")
        self.qk = ConvW(torch.cat([wq, wk], 0), torch.cat([bq, bk], 0), dev)
        self.v = ConvW(*_wb(sd, p + ".to_v"), dev)
        self.out = ConvW(*_wb(sd, p + ".to_out.0"), dev)

    def __call__(self, x):
        n, h, w, c = x.shape
        win = self.win
        nwy, nwx = (h + win - 1) // win, (w + win - 1) // win
        nwin, T = n * nwy * nwx, win * win
        # GroupNorm statistics are per *window* in the reference (the Attention module sees one window at a
        # time), so normalise after partitioning, with each window as its own "image".
        xw = ops.window_partition(x, win)                                  # [nwin, T, c]
        valid = None
        if h % win or w % win:
            # edge windows hold fewer tokens: the partition zero-pads them to T rows; GroupNorm counts and the
            # softmax key range use the real token count of each window
            key = (n, h, w, str(x.device))
            if key not in self._valid:
                vh = torch.tensor([min(win, h - i * win) for i in range(nwy)])
                vw = torch.tensor([min(win, w - j * win) for j in range(nwx)])
                self._valid[key] = (vh[:, None] * vw[None, :]).reshape(-1).repeat(n).to(device=x.device, dtype=torch.int32)
            valid = self._valid[key]
        hn = self.norm(xw.view(nwin, 1, T, c), silu=False, valid=valid)
        qk = igemm(hn, self.qk).view(nwin, T, 2 * c)
        vT = torch.empty((nwin, c, T), device=x.device, dtype=torch.bfloat16)
        igemm(hn, self.v, store=ST_TRANSPOSED, out=vT)
        o = torch.empty((nwin, T, c), device=x.device, dtype=torch.bfloat16)
        ops.attention_unfused(qk[:, :, :c], qk[:, :, c:], vT, o, heads=1, head_dim=c, valid=valid)
        o = igemm(o.view(nwin, 1, T, c), self.out)
        return ops.window_merge(o, x, win)                                 # + residual


class VAEDecoder(LazyNet):
    """post_quant_conv is applied in fp32 by x0_prepare; input here is its [hi|lo] bf16 split (8 channels).
    `load_state_dict` takes the decoder half of the SD-2.1 VAE in diffusers naming (weights.vae_spec)."""

    IGNORED_PREFIXES = ("encoder.", "quant_conv.")

    def __init__(self, sd=None, dev="cuda", attn_patch=16):
        self.dev, self.attn_patch = torch.device(dev), attn_patch
        self._lazy_init(sd)

    def _spec(self):
        return vae_spec()

    def set_attn_patch(self, attn_patch):
        """autoencoders_patch_attn.py:77-81"""
        assert attn_patch > 0, "attn_patch must be greater than 0"
        self.attn_patch = attn_patch
        if self.__dict__["_built"]:
            self.attn.win = attn_patch

    def _build(self, sd):
        dev, attn_patch = self.dev, self.attn_patch
        d = "decoder"
        w, b = _wb(sd, d + ".conv_in")
        self.conv_in = ConvW(torch.cat([w, w], dim=1), b, dev)             # W.(hi + lo)
        self.mid0 = VAERes(sd, d + ".mid_block.resnets.0", 512, 512, dev)
        self.attn = VAEWindowAttention(sd, d + ".mid_block.attentions.0", 512, attn_patch, dev)
        self.mid1 = VAERes(sd, d + ".mid_block.resnets.1", 512, 512, dev)
        self.ups = []
        prev = 512
        for i, c in enumerate((512, 512, 256, 128)):
            res = []
            for j in range(3):
                res.append(VAERes(sd, f"{d}.up_blocks.{i}.resnets.{j}", prev, c, dev))
                prev = c
            up = UpConv(*_wb(sd, f"{d}.up_blocks.{i}.upsamplers.0.conv"), dev) if i < 3 else None
            self.ups.append((res, up))
        self.norm_out = GroupNorm(sd[d + ".conv_norm_out.weight"], sd[d + ".conv_norm_out.bias"], 1e-6, device=dev)
        self.conv_out = TapConv3x3(*_wb(sd, d + ".conv_out"), dev)
        self.pq_w = sd["post_quant_conv.weight"].float().reshape(4, 4)
        self.pq_b = sd["post_quant_conv.bias"].float()

    def __call__(self, z_hilo):
        self._ensure()
        n, h, w, _ = z_hilo.shape
        t = igemm(z_hilo, self.conv_in, stats=True)
        t = self.mid0(t)
        t = self.attn(t)
        t = self.mid1(t)
        for res, up in self.ups:
            for r in res:
                t = r(t)
            if up is not None:
                t = up(t)
        # NCHW fp32 image straight out of the last conv's epilogue
        img = self.conv_out(self.norm_out(t), planar=True)
        return img.view(n, 3, 8 * h, 8 * w)
