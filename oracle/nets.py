"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by onedc_b200/).

fp32 PyTorch restatement of every network on the decode hot path.  Module/parameter
names equal the reference's state-dict keys, so the product's weight inventory
(onedc_b200/weights.py) loads with strict=True -- that is itself a structural check.

  * codec half  : restates modules/dcvc.py:183-208,242-265,353-367,424-433,
                  modules/vqgan/blocks.py:15-107 and
                  models/sd15_onedc_codec_stage1/codec_module.py:88-181,205-217.
                  PINNED: tests/test_oracle_pinned.py compares it with the reference
                  IntraNoAR imported unchanged (bit-exact on CPU) and tests/golden holds
                  reference outputs.
  * UNet/LoRA, VAE, x0 : restated from diffusers==0.32.1 / peft==0.14.0 semantics
                  (SURVEY.md section 8c notes) following the reference call sites
                  decoder_unet.py:14-29,32-305,331-368, autoencoders_patch_attn.py:9-62,
                  model_sd15_with_codec_stage1.py:184-188,296-330, modules/dmd/utils.py:279-284.
                  Those packages are not installed and not vendored => **parity unpinned**
                  for this part (no reference output exists to compare with).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ============================================================================ codec building blocks
class DepthConv(nn.Module):                       # dcvc.py:242-265
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(cin, cin, 1), nn.LeakyReLU(0.01))
        self.depth_conv = nn.Conv2d(cin, cin, 3, padding=1, groups=cin)
        self.conv2 = nn.Conv2d(cin, cout, 1)
        self.adaptor = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        idt = x if self.adaptor is None else self.adaptor(x)
        return self.conv2(self.depth_conv(self.conv1(x))) + idt


class ConvFFN3(nn.Module):                        # dcvc.py:353-367
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c * 4, 1)
        self.conv_out = nn.Conv2d(c * 2, c, 1)

    def forward(self, x):
        x1, x2 = self.conv(x).chunk(2, 1)
        return x + self.conv_out(F.leaky_relu(x1, 0.1) + F.leaky_relu(x2, 0.01))


class DepthConvBlock4(nn.Module):                 # dcvc.py:424-433
    def __init__(self, cin, cout):
        super().__init__()
        self.block = nn.Sequential(DepthConv(cin, cout), ConvFFN3(cout))

    def forward(self, x):
        return self.block(x)


class ResidualBlockUpsample(nn.Module):           # dcvc.py:183-208
    def __init__(self, cin, cout):
        super().__init__()
        self.subpel_conv = nn.Sequential(nn.Conv2d(cin, cout * 4, 1), nn.PixelShuffle(2))
        self.conv = nn.Conv2d(cout, cout, 3, padding=1)
        self.upsample = nn.Sequential(nn.Conv2d(cin, cout * 4, 1), nn.PixelShuffle(2))

    def forward(self, x):
        out = F.leaky_relu(self.subpel_conv(x), 0.01)
        out = F.leaky_relu(self.conv(out), 0.1)
        return out + self.upsample(x)


class VQResnetBlock(nn.Module):                   # vqgan/blocks.py:15-52 (in == out, no shortcut)
    def __init__(self, c):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, c, eps=1e-6)
        self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.norm2 = nn.GroupNorm(32, c, eps=1e-6)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return h + x


class VQAttnBlock(nn.Module):                     # vqgan/blocks.py:55-107
    def __init__(self, c):
        super().__init__()
        self.norm = nn.GroupNorm(32, c, eps=1e-6)
        self.q = nn.Conv2d(c, c, 1)
        self.k = nn.Conv2d(c, c, 1)
        self.v = nn.Conv2d(c, c, 1)
        self.proj_out = nn.Conv2d(c, c, 1)

    def forward(self, x):
        h = self.norm(x)
        b, c, hh, ww = h.shape
        q = self.q(h).reshape(b, c, hh * ww).permute(0, 2, 1)
        k = self.k(h).reshape(b, c, hh * ww)
        v = self.v(h).reshape(b, c, hh * ww)
        w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
        h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
        return x + self.proj_out(h)


class LatentSynthesis(nn.Module):                 # codec_module.py:88-116 (class Decoder)
    def __init__(self, in_ch=128, internal=512, sem=768, out_ch=320):
        super().__init__()
        c8, c16 = internal // 2, internal
        self.trans_coding = nn.Sequential(DepthConvBlock4(in_ch, c16), DepthConvBlock4(c16, c16))
        self.blocks = nn.Sequential(
            VQResnetBlock(c16), VQResnetBlock(c16), VQResnetBlock(c16),
            nn.Conv2d(c16, c16 * 4, 1), nn.PixelShuffle(2), nn.Conv2d(c16, c8, 3, padding=1),
            VQResnetBlock(c8), VQResnetBlock(c8), VQResnetBlock(c8))
        self.sem_up = nn.Sequential(
            ResidualBlockUpsample(sem, c16), DepthConvBlock4(c16, c16),
            ResidualBlockUpsample(c16, c8), DepthConvBlock4(c8, c8),
            ResidualBlockUpsample(c8, c8))
        self.conv_out = DepthConvBlock4(c8 * 2, out_ch)

    def forward(self, y_hat, sem_hat):
        y = self.blocks(self.trans_coding(y_hat))
        s = self.sem_up(sem_hat)
        return self.conv_out(torch.cat([y, s], dim=1))


class HyperDecoder(nn.Module):                    # codec_module.py:145-166
    def __init__(self, ch=128, zdim=7):
        super().__init__()
        self.feat_in = nn.Sequential(nn.Conv2d(zdim, ch, 1), nn.LeakyReLU(0.01))
        self.to_entropy = nn.Sequential(
            DepthConvBlock4(ch, ch), ResidualBlockUpsample(ch, ch), DepthConvBlock4(ch, ch),
            ResidualBlockUpsample(ch, ch), DepthConvBlock4(ch, ch))

    def forward(self, z_hat):
        z_sem = self.feat_in(z_hat)
        return self.to_entropy(z_sem), z_sem


class SemanticAdaptor(nn.Module):                 # codec_module.py:169-181
    def __init__(self, ch=128, sem=768):
        super().__init__()
        self.to_semantic = nn.Sequential(
            DepthConvBlock4(ch, sem),
            VQResnetBlock(sem), VQAttnBlock(sem), VQAttnBlock(sem),
            VQResnetBlock(sem), VQAttnBlock(sem), VQAttnBlock(sem),
            DepthConvBlock4(sem, sem))

    def forward(self, x):
        return self.to_semantic(x)


def fsq_indices_to_codes(indices, levels=(4,) * 7):
    """vector_quantize_pytorch==1.21.2 FSQ.indices_to_codes (call site codec_module.py:431):
    level_i = (idx // 4^i) % 4; code = (level - 2) / 2.  Returns (B, 7, H, W) fp32."""
    basis = torch.tensor([4 ** i for i in range(len(levels))], dtype=torch.int64)
    lv = (indices.long().unsqueeze(-1) // basis) % 4
    return ((lv - 2).float() / 2.0).permute(0, 3, 1, 2).contiguous()


class CodecNets(nn.Module):
    """Decode-side modules of IntraNoAR under the reference's attribute names
    (codec_module.py:196-217)."""

    def __init__(self, N=128, internal=512, sem=768, ctrl=320):
        super().__init__()
        self.dec = LatentSynthesis(N, internal, sem, ctrl)
        self.semantic_adaptor = SemanticAdaptor(N, sem)
        self.hyper_dec = HyperDecoder(N, 7)
        self.y_prior_fusion = nn.Sequential(DepthConvBlock4(N, 2 * N), DepthConvBlock4(2 * N, 2 * N))
        self.y_spatial_prior_reduction = nn.Conv2d(2 * N, N, 1)
        self.y_spatial_prior_adaptor_1 = DepthConvBlock4(2 * N, 2 * N)
        self.y_spatial_prior_adaptor_2 = DepthConvBlock4(2 * N, 2 * N)
        self.y_spatial_prior_adaptor_3 = DepthConvBlock4(2 * N, 2 * N)
        self.y_spatial_prior = nn.Sequential(*[DepthConvBlock4(2 * N, 2 * N) for _ in range(3)])


# ============================================================================ UNet (diffusers 0.32.1 semantics)
class LoraLinear(nn.Module):
    """peft 0.14.0 lora.Linear: y = base(x) + scaling * B(A(x)), scaling = alpha / r (never merged in
    the reference)."""

    def __init__(self, cin, cout, bias=True, r=64, alpha=8.0):
        super().__init__()
        self.base_layer = nn.Linear(cin, cout, bias=bias)
        self.lora_A = nn.ModuleDict({"default": nn.Linear(cin, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, cout, bias=False)})
        self.scaling = alpha / r

    def forward(self, x):
        return self.base_layer(x) + self.scaling * self.lora_B["default"](self.lora_A["default"](x))


class LoraConv2d(nn.Module):
    """peft 0.14.0 lora.Conv2d: A = Conv(in->r, same k/stride/pad, no bias), B = Conv1x1(r->out)."""

    def __init__(self, cin, cout, k, stride=1, padding=0, r=64, alpha=8.0):
        super().__init__()
        self.base_layer = nn.Conv2d(cin, cout, k, stride=stride, padding=padding)
        self.lora_A = nn.ModuleDict({"default": nn.Conv2d(cin, r, k, stride=stride, padding=padding, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Conv2d(r, cout, 1, bias=False)})
        self.scaling = alpha / r

    def forward(self, x):
        return self.base_layer(x) + self.scaling * self.lora_B["default"](self.lora_A["default"](x))


class UNetResnet(nn.Module):                      # diffusers ResnetBlock2D (eps 1e-5, swish, time_embedding_norm default)
    def __init__(self, cin, cout, temb=1280):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-5)
        self.conv1 = LoraConv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = LoraLinear(temb, cout)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-5)
        self.conv2 = LoraConv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = LoraConv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, emb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(emb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h                              # output_scale_factor = 1.0


class Attention(nn.Module):                       # diffusers Attention + AttnProcessor2_0 (SDPA)
    def __init__(self, c, ctx, heads):
        super().__init__()
        self.heads = heads
        self.to_q = LoraLinear(c, c, bias=False)
        self.to_k = LoraLinear(ctx, c, bias=False)
        self.to_v = LoraLinear(ctx, c, bias=False)
        self.to_out = nn.ModuleList([LoraLinear(c, c), nn.Identity()])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, s, c = x.shape
        h = self.heads
        q = self.to_q(x).view(b, s, h, c // h).transpose(1, 2)
        k = self.to_k(ctx).view(b, -1, h, c // h).transpose(1, 2)
        v = self.to_v(ctx).view(b, -1, h, c // h).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)   # scale = 1/sqrt(head_dim)
        return self.to_out[0](o.transpose(1, 2).reshape(b, s, c))


class GEGLU(nn.Module):                           # diffusers GEGLU: hidden, gate = chunk; hidden * gelu(gate) (erf)
    def __init__(self, c, inner):
        super().__init__()
        self.proj = LoraLinear(c, inner * 2)

    def forward(self, x):
        h, g = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(c, 4 * c), nn.Identity(), LoraLinear(4 * c, c)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, c, ctx, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(c)
        self.attn1 = Attention(c, c, heads)
        self.norm2 = nn.LayerNorm(c)
        self.attn2 = Attention(c, ctx, heads)
        self.norm3 = nn.LayerNorm(c)
        self.ff = FeedForward(c)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2D(nn.Module):                   # diffusers Transformer2DModel, use_linear_projection=False
    def __init__(self, c, ctx=768, heads=8):
        super().__init__()
        self.norm = nn.GroupNorm(32, c, eps=1e-6)
        self.proj_in = LoraConv2d(c, c, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(c, ctx, heads)])
        self.proj_out = LoraConv2d(c, c, 1)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        t = self.proj_in(self.norm(x)).permute(0, 2, 3, 1).reshape(b, h * w, c)
        t = self.transformer_blocks[0](t, ctx)
        t = t.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(t) + x


class _Sampler(nn.Module):
    def __init__(self, c, down):
        super().__init__()
        self.down = down
        self.conv = LoraConv2d(c, c, 3, stride=2 if down else 1, padding=1)

    def forward(self, x):
        if not self.down:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x)


class _Block(nn.Module):
    pass


class ReduceResblock(nn.Module):                  # decoder_unet.py:14-29
    def __init__(self, cin, cout):
        super().__init__()
        self.short_cut = nn.Conv2d(cin, cout, 1)
        self.blocks = nn.Sequential(
            nn.GroupNorm(32, cin, eps=1e-6), nn.SiLU(), nn.Conv2d(cin, cin, 3, padding=1),
            nn.GroupNorm(32, cin, eps=1e-6), nn.SiLU(), nn.Conv2d(cin, cout, 3, padding=1))

    def forward(self, x):
        return self.blocks(x) + self.short_cut(x)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, c):
        super().__init__()
        self.linear_1 = nn.Linear(cin, c)
        self.linear_2 = nn.Linear(c, c)

    def forward(self, t):
        return self.linear_2(F.silu(self.linear_1(t)))


def sinusoidal_timestep(t, dim=320):
    """diffusers Timesteps(320, flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    e = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t.float()[:, None] * e[None, :]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


class UNetOracle(nn.Module):
    """SD1.5 UNet2DConditionModel + LoRA(r=64, alpha=8) + replaced conv_in + vae_reduction, with the
    control flow of forward_unet (decoder_unet.py:32-305): returns (eps, reduced_sample)."""

    CH = (320, 640, 1280, 1280)

    def __init__(self, in_ch=320, vae_ch=4):
        super().__init__()
        ch = self.CH
        self.conv_in = nn.Conv2d(in_ch, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], 1280)
        self.down_blocks = nn.ModuleList()
        cin = ch[0]
        for i, c in enumerate(ch):
            b = _Block()
            b.resnets = nn.ModuleList([UNetResnet(cin if j == 0 else c, c) for j in range(2)])
            if i < 3:
                b.attentions = nn.ModuleList([Transformer2D(c) for _ in range(2)])
                b.downsamplers = nn.ModuleList([_Sampler(c, True)])
            self.down_blocks.append(b)
            cin = c
        self.mid_block = _Block()
        self.mid_block.resnets = nn.ModuleList([UNetResnet(1280, 1280), UNetResnet(1280, 1280)])
        self.mid_block.attentions = nn.ModuleList([Transformer2D(1280)])
        skips = [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
        self.up_blocks = nn.ModuleList()
        prev = 1280
        for i, c in enumerate(reversed(ch)):
            b = _Block()
            res = []
            for j in range(3):
                res.append(UNetResnet(prev + skips.pop(), c))
                prev = c
            b.resnets = nn.ModuleList(res)
            if i > 0:
                b.attentions = nn.ModuleList([Transformer2D(c) for _ in range(3)])
            if i < 3:
                b.upsamplers = nn.ModuleList([_Sampler(c, False)])
            self.up_blocks.append(b)
        self.conv_norm_out = nn.GroupNorm(32, 320, eps=1e-5)
        self.conv_out = nn.Conv2d(320, vae_ch, 3, padding=1)
        self.vae_reduction = ReduceResblock(in_ch, vae_ch)

    def forward(self, sample, timestep, ctx):
        reduced = self.vae_reduction(sample)                                   # :100
        emb = self.time_embedding(sinusoidal_timestep(timestep))               # :138-139
        h = self.conv_in(sample)                                               # :165
        stack = [h]
        for b in self.down_blocks:                                             # :205-226
            for j, r in enumerate(b.resnets):
                h = r(h, emb)
                if hasattr(b, "attentions"):
                    h = b.attentions[j](h, ctx)
                stack.append(h)
            if hasattr(b, "downsamplers"):
                h = b.downsamplers[0](h)
                stack.append(h)
        h = self.mid_block.resnets[0](h, emb)                                  # :240-251
        h = self.mid_block.attentions[0](h, ctx)
        h = self.mid_block.resnets[1](h, emb)
        for b in self.up_blocks:                                               # :265-293
            for j, r in enumerate(b.resnets):
                h = r(torch.cat([h, stack.pop()], dim=1), emb)
                if hasattr(b, "attentions"):
                    h = b.attentions[j](h, ctx)
            if hasattr(b, "upsamplers"):
                h = b.upsamplers[0](h)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))                       # :296-299
        return h, reduced


def alphas_cumprod_sd15():
    """DDIMScheduler config of runwayml/stable-diffusion-v1-5: scaled_linear betas in
    [0.00085, 0.012], 1000 steps (model_sd15_with_codec_stage1.py:103-106)."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def x0_from_noise(sample, model_output, alphas_cumprod, timestep):
    """modules/dmd/utils.py:279-284 (called in float64, model_sd15_with_codec_stage1.py:320-322)."""
    a = alphas_cumprod[timestep].reshape(-1, 1, 1, 1)
    return (sample - (1 - a) ** 0.5 * model_output) / a ** 0.5


# ============================================================================ VAE decoder
class VAEResnet(nn.Module):                       # diffusers ResnetBlock2D, temb=None, eps 1e-6
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class VAEAttention(nn.Module):                    # diffusers Attention(heads=1, residual_connection=True, norm_num_groups=32)
    def __init__(self, c):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, c, eps=1e-6)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o).transpose(1, 2).reshape(b, c, h, w)
        return o + x


class _Up(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _VAEDecoder(nn.Module):
    def __init__(self, latent=4, attn_patch=16):
        super().__init__()
        self.attn_patch = attn_patch
        self.conv_in = nn.Conv2d(latent, 512, 3, padding=1)
        self.mid_block = _Block()
        self.mid_block.resnets = nn.ModuleList([VAEResnet(512, 512), VAEResnet(512, 512)])
        self.mid_block.attentions = nn.ModuleList([VAEAttention(512)])
        self.up_blocks = nn.ModuleList()
        prev = 512
        for i, c in enumerate((512, 512, 256, 128)):
            b = _Block()
            res = []
            for j in range(3):
                res.append(VAEResnet(prev, c))
                prev = c
            b.resnets = nn.ModuleList(res)
            if i < 3:
                b.upsamplers = nn.ModuleList([_Up(c)])
            self.up_blocks.append(b)
        self.conv_norm_out = nn.GroupNorm(32, 128, eps=1e-6)
        self.conv_out = nn.Conv2d(128, 3, 3, padding=1)

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid_block.resnets[0](h)
        # eval-mode windowed attention, autoencoders_patch_attn.py:20-29 (windows written back in place;
        # equivalent to block-diagonal attention, edge windows smaller)
        p = self.attn_patch
        out = h.clone()
        for i in range(0, h.shape[2], p):
            for j in range(0, h.shape[3], p):
                out[:, :, i:i + p, j:j + p] = self.mid_block.attentions[0](h[:, :, i:i + p, j:j + p].contiguous())
        h = self.mid_block.resnets[1](out)
        for b in self.up_blocks:
            for r in b.resnets:
                h = r(h)
            if hasattr(b, "upsamplers"):
                h = b.upsamplers[0](h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class VAEOracle(nn.Module):
    """AutoencoderKL_patch_attn.decode (autoencoders_patch_attn.py:65-81) for the SD2.1 VAE;
    vae_decode_image divides by scaling_factor 0.18215 first (model_sd15_with_codec_stage1.py:184-188)."""

    def __init__(self, attn_patch=16):
        super().__init__()
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        self.decoder = _VAEDecoder(4, attn_patch)

    def forward(self, latents):
        return self.decoder(self.post_quant_conv(1 / 0.18215 * latents))     # model...py:186, as written there
