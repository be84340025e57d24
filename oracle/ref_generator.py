"""TEST INFRASTRUCTURE ONLY -- never imported by the product (onedc_b200/).

Executes the reference's OWN generator-side source, unmodified, where it lies under /root/reference/src:

  models/sd15_onedc_codec_stage1/model_sd15_with_codec_stage1.py:296-330  SD15_1step_codec_stage1.decode
  models/sd15_onedc_codec_stage1/model_sd15_with_codec_stage1.py:184-188  vae_decode_image
  models/sd15_onedc_codec_stage1/decoder_unet.py:32-305                   forward_unet
  models/sd15_onedc_codec_stage1/decoder_unet.py:14-29                    reduce_resblock
  modules/vae/autoencoders_patch_attn.py:9-62                             windowed_attn / windowed_attn_forward
  modules/dmd/utils.py:279-284                                            get_x0_from_noise
  models/sd15_onedc_codec_stage1/codec_module.py:357-454                  IntraNoAR.decode / _decompress (via ref_import)

so that the oracle restatement (oracle/nets.py, oracle/decode.py) can be pinned bit-exactly against them
(tests/test_reference_pin_cpu.py) and golden vectors can be generated (tests/golden/gen_golden.py).

What the reference source needs from `diffusers` / `peft` (neither installed nor vendored) is supplied here:
name-only shims for the imports, and *diffusers-shaped* block adapters wrapped around the oracle's own
modules -- each adapter takes the arguments the reference call site passes (decoder_unet.py:205-293,
autoencoders_patch_attn.py:34-62) and restates the published block forward of diffusers 0.32.1
(CrossAttnDownBlock2D, DownBlock2D, UNetMidBlock2DCrossAttn, UpBlock2D, CrossAttnUpBlock2D, Decoder,
AutoencoderKL.decode).  What stays unpinned after this is therefore exactly: the per-module arithmetic of
diffusers (ResnetBlock2D, Transformer2DModel/BasicTransformerBlock/Attention/GEGLU, Up/Downsample2D,
Timesteps/TimestepEmbedding, the DDIM alphas_cumprod) and peft's LoRA layers.  Everything the reference
itself wrote -- control flow, skip stack order, windowing, the x0 formula in float64, the 1/0.18215 scaling,
un-padding -- is executed, not restated.

Only usable where /root/reference exists (the build container).
"""
import sys
import types

import torch
import torch.nn as nn

from . import ref_import
from .nets import UNetOracle, VAEOracle, sinusoidal_timestep


def _shim(name, **kw):
    if name in sys.modules and not getattr(sys.modules[name], "_onedc_shim", False):
        return sys.modules[name]
    m = sys.modules.get(name) or types.ModuleType(name)
    m._onedc_shim = True
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


class _Named(nn.Module):
    """placeholder base class for names the reference only subclasses or type-annotates"""


_cached = {}


def import_reference_generator():
    """Imports the reference generator-side modules unchanged.  Returns a namespace with `model_cls`
    (SD15_1step_codec_stage1), `forward_unet`, `reduce_resblock`, `windowed_attn_forward`, `windowed_attn`,
    `get_x0_from_noise`, `NoOpContext`."""
    if _cached:
        return _cached["ns"]
    ns0 = ref_import.import_reference()                      # codec half + its shims (diffusers name stubs included)

    def missing(name):
        try:
            __import__(name)
            return False
        except Exception:
            return True

    try:                                     # real package; must be imported before the `accelerate` name shim exists
        from transformers import PretrainedConfig  # noqa: F401  (transformers probes accelerate via find_spec)
    except Exception:
        _shim("transformers", PretrainedConfig=object)
    d = sys.modules["diffusers"]
    for n in ("UNet2DConditionModel", "ControlNetModel", "AutoencoderKL", "AutoencoderTiny", "DDIMScheduler"):
        if not hasattr(d, n):
            setattr(d, n, type(n, (_Named,), {}))
    b = sys.modules["diffusers.models.unets.unet_2d_blocks"]
    if not hasattr(b, "Transformer2DModel"):
        b.Transformer2DModel = type("Transformer2DModel", (_Named,), {})
    _shim("diffusers.models.unets.unet_2d_condition", UNet2DConditionOutput=object)
    if missing("peft"):
        _shim("peft", LoraConfig=object)
    if missing("matplotlib"):
        _shim("matplotlib", use=lambda *a, **k: None)
        _shim("matplotlib.backends")
        _shim("matplotlib.backends.backend_agg", FigureCanvasAgg=object)
        _shim("matplotlib.pyplot")
    if missing("imageio"):
        _shim("imageio")
        _shim("imageio.v2")
    if missing("accelerate"):
        _shim("accelerate", Accelerator=object)
    if missing("omegaconf"):
        _shim("omegaconf", OmegaConf=object)
    if missing("wandb"):
        _shim("wandb")
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:
        _shim("torch.utils.tensorboard", SummaryWriter=object)
    import modules.dmd.utils as dmd
    import modules.vae.autoencoders_patch_attn as pa
    import models.sd15_onedc_codec_stage1.decoder_unet as du
    import models.sd15_onedc_codec_stage1.model_sd15_with_codec_stage1 as top
    ns = types.SimpleNamespace(codec=ns0, model_cls=top.SD15_1step_codec_stage1, forward_unet=du.forward_unet,
                               reduce_resblock=du.reduce_resblock, windowed_attn_forward=pa.windowed_attn_forward,
                               windowed_attn=pa.windowed_attn, get_x0_from_noise=dmd.get_x0_from_noise,
                               NoOpContext=dmd.NoOpContext)
    _cached["ns"] = ns
    return ns


# ---------------------------------------------------------------------------------------------------------
# diffusers-shaped adapters around the ORACLE's modules (shared parameters: no copies)
# ---------------------------------------------------------------------------------------------------------
class _Tr(nn.Module):
    """Transformer2DModel call shape: attn(hidden_states, encoder_hidden_states=..., return_dict=False)[0]."""

    def __init__(self, tr):
        super().__init__()
        self.tr = tr

    def forward(self, hidden_states, encoder_hidden_states=None, cross_attention_kwargs=None, attention_mask=None,
                encoder_attention_mask=None, return_dict=False):
        assert attention_mask is None and encoder_attention_mask is None and not cross_attention_kwargs
        return (self.tr(hidden_states, encoder_hidden_states),)


class _DownBlock(nn.Module):
    """diffusers CrossAttnDownBlock2D / DownBlock2D.forward: per layer resnet(+attn), every output pushed;
    then the downsampler, pushed too."""

    def __init__(self, blk):
        super().__init__()
        self.resnets = blk.resnets
        self.has_cross_attention = hasattr(blk, "attentions")
        if self.has_cross_attention:
            self.attentions = nn.ModuleList([_Tr(t) for t in blk.attentions])
        self.downsamplers = blk.downsamplers if hasattr(blk, "downsamplers") else None

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None, encoder_attention_mask=None):
        out = ()
        for j, resnet in enumerate(self.resnets):
            hidden_states = resnet(hidden_states, temb)
            if self.has_cross_attention:
                hidden_states = self.attentions[j](hidden_states, encoder_hidden_states=encoder_hidden_states,
                                                   cross_attention_kwargs=cross_attention_kwargs,
                                                   attention_mask=attention_mask,
                                                   encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
            out = out + (hidden_states,)
        if self.downsamplers is not None:
            for ds in self.downsamplers:
                hidden_states = ds(hidden_states)
            out = out + (hidden_states,)
        return hidden_states, out


class _MidBlock(nn.Module):
    """diffusers UNetMidBlock2DCrossAttn.forward: resnets[0], then (attn, resnet) pairs."""

    has_cross_attention = True

    def __init__(self, blk):
        super().__init__()
        self.resnets = blk.resnets
        self.attentions = nn.ModuleList([_Tr(t) for t in blk.attentions])

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None, encoder_attention_mask=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states, return_dict=False)[0]
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class _UpBlock(nn.Module):
    """diffusers UpBlock2D / CrossAttnUpBlock2D.forward: pop the LAST skip, cat([hidden, skip], 1), resnet(+attn);
    then the upsampler."""

    def __init__(self, blk):
        super().__init__()
        self.resnets = blk.resnets
        self.has_cross_attention = hasattr(blk, "attentions")
        if self.has_cross_attention:
            self.attentions = nn.ModuleList([_Tr(t) for t in blk.attentions])
        self.upsamplers = blk.upsamplers if hasattr(blk, "upsamplers") else None

    def forward(self, hidden_states, temb=None, res_hidden_states_tuple=None, encoder_hidden_states=None,
                cross_attention_kwargs=None, upsample_size=None, attention_mask=None, encoder_attention_mask=None):
        assert upsample_size is None, "sizes here are multiples of 8: forward_unet never forwards an upsample size"
        for j, resnet in enumerate(self.resnets):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb)
            if self.has_cross_attention:
                hidden_states = self.attentions[j](hidden_states, encoder_hidden_states=encoder_hidden_states,
                                                   return_dict=False)[0]
        if self.upsamplers is not None:
            for up in self.upsamplers:
                hidden_states = up(hidden_states)
        return hidden_states


class _TimeEmbedding(nn.Module):
    def __init__(self, te):
        super().__init__()
        self.te = te

    def forward(self, t_emb, timestep_cond=None):
        assert timestep_cond is None
        return self.te(t_emb)


class DiffusersShapedUNet(nn.Module):
    """The attribute surface `forward_unet` reads from a diffusers UNet2DConditionModel, built over a UNetOracle
    (same parameter tensors).  `forward` IS the reference's forward_unet, bound exactly as the reference binds it
    (decoder_unet.py:403)."""

    def __init__(self, oracle_unet: UNetOracle, forward_unet):
        super().__init__()
        o = oracle_unet
        self.config = types.SimpleNamespace(center_input_sample=False, class_embeddings_concat=False,
                                            addition_embed_type=None)
        self.num_upsamplers = 3
        self.vae_reduction = o.vae_reduction
        self.time_embedding = _TimeEmbedding(o.time_embedding)
        self.time_embed_act = None
        self.conv_in = o.conv_in
        self.down_blocks = nn.ModuleList([_DownBlock(b) for b in o.down_blocks])
        self.mid_block = _MidBlock(o.mid_block)
        self.up_blocks = nn.ModuleList([_UpBlock(b) for b in o.up_blocks])
        self.conv_norm_out = o.conv_norm_out
        self.conv_act = nn.SiLU()
        self.conv_out = o.conv_out
        self.forward = forward_unet.__get__(self)

    # diffusers UNet2DConditionModel helpers used by forward_unet (:138-160)
    def get_time_embed(self, sample, timestep):
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.long, device=sample.device)
        timestep = timestep.expand(sample.shape[0])
        return sinusoidal_timestep(timestep).to(dtype=sample.dtype)

    def get_class_embed(self, sample, class_labels):
        return None

    def get_aug_embed(self, emb, encoder_hidden_states, added_cond_kwargs):
        return None

    def process_encoder_hidden_states(self, encoder_hidden_states, added_cond_kwargs):
        return encoder_hidden_states


class _VAEMid(nn.Module):
    """diffusers UNetMidBlock2D attribute surface (resnets, attentions, gradient_checkpointing); its forward is the
    reference's windowed_attn_forward, bound exactly as autoencoders_patch_attn.py:71 binds it."""

    def __init__(self, mid, windowed_attn_forward, attn_patch):
        super().__init__()
        self.resnets = nn.ModuleList([_TembRes(r) for r in mid.resnets])
        self.attentions = mid.attentions
        self.gradient_checkpointing = True                   # model...py:48 enables it; inert under no_grad
        self.attn_patch = attn_patch
        self.forward = windowed_attn_forward.__get__(self, self.__class__)


class _TembRes(nn.Module):
    def __init__(self, r):
        super().__init__()
        self.r = r

    def forward(self, x, temb=None):
        return self.r(x)


class DiffusersShapedVAE(nn.Module):
    """AutoencoderKL surface used by vae_decode_image (model...py:184-188): `.config.scaling_factor`,
    `.decode(latents).sample` = Decoder(post_quant_conv(latents)) with diffusers' Decoder.forward order
    (conv_in, mid_block, up_blocks, conv_norm_out, SiLU, conv_out)."""

    def __init__(self, oracle_vae: VAEOracle, windowed_attn_forward, attn_patch=16):
        super().__init__()
        self.config = types.SimpleNamespace(scaling_factor=0.18215)
        self.post_quant_conv = oracle_vae.post_quant_conv
        self.dec = oracle_vae.decoder
        self.mid_block = _VAEMid(self.dec.mid_block, windowed_attn_forward, attn_patch)

    def decode(self, latents):
        d = self.dec
        h = d.conv_in(self.post_quant_conv(latents))
        h = self.mid_block(h, None)
        for b in d.up_blocks:
            for r in b.resnets:
                h = r(h)
            if hasattr(b, "upsamplers"):
                h = b.upsamplers[0](h)
        h = d.conv_out(torch.nn.functional.silu(d.conv_norm_out(h)))
        return types.SimpleNamespace(sample=h)


def build_reference_model(unet_sd, codec_sd, vae_sd, attn_patch=16, timestep=999, alphas_cumprod=None):
    """An instance of the REFERENCE class SD15_1step_codec_stage1 whose `decode()` runs unmodified on CPU fp32.
    __init__ is bypassed (it downloads checkpoints through from_pretrained); the attributes `decode` reads are set
    from the reference's own IntraNoAR (codec half) and the adapters above."""
    from .nets import alphas_cumprod_sd15
    ns = import_reference_generator()
    codec = ref_import.build_reference_codec()
    codec.load_state_dict(codec_sd, strict=False)            # analysis-side keys are absent from the decode inventory
    ou, ov = UNetOracle().eval(), VAEOracle(attn_patch).eval()
    ou.load_state_dict(unet_sd, strict=True)
    ov.load_state_dict(vae_sd, strict=True)
    m = ns.model_cls.__new__(ns.model_cls)
    nn.Module.__init__(m)
    m.accelerator = types.SimpleNamespace(device=torch.device("cpu"))
    m.network_context_manager = ns.NoOpContext()             # use_fp16=False branch of model...py:114
    m.codec_model = codec
    m.feedforward_model = DiffusersShapedUNet(ou, ns.forward_unet).eval()
    m.alphas_cumprod = alphas_cumprod_sd15() if alphas_cumprod is None else alphas_cumprod
    m.conditioning_timestep = timestep
    m.use_large_vae = True
    m.vae_large = DiffusersShapedVAE(ov, ns.windowed_attn_forward, attn_patch).eval()
    m.vae = None
    return m                 # the reference's own eval() touches the (absent) codeformer: every sub-module is in eval mode already
