"""Parameter inventory of the decode hot path + deterministic random initialisation.

The names are the reference's own state-dict keys, so real checkpoints load unchanged:
  * codec  : `model_1.safetensors` -> IntraNoAR (decode-side modules only)
             /root/reference/src/models/sd15_onedc_codec_stage1/codec_module.py:88-217
  * unet   : `model.safetensors`  -> diffusers SD1.5 UNet2DConditionModel wrapped by peft LoRA
             (`*.base_layer.*`, `*.lora_A.default.weight`, `*.lora_B.default.weight`) with the
             replaced conv_in and the added vae_reduction, decoder_unet.py:14-29,331-400
  * vae    : diffusers AutoencoderKL (stabilityai/stable-diffusion-2-1 vae) decoder half,
             model_sd15_with_codec_stage1.py:43-47

No checkpoints exist offline, so every test/bench uses `random_state_dict(spec, seed)`:
weight/bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (PyTorch's default conv/linear init),
norm gamma ~ 1 + 0.1 U(-1,1), beta ~ 0.1 U(-1,1), LoRA A,B ~ N(0, 0.02) (so that merging
is exercised; peft's default B=0 would make the merge untestable).  Each tensor is seeded
by crc32(name) so the result does not depend on construction order.
"""
import math
import zlib

import torch

LORA_RANK = 64
LORA_ALPHA = 8.0
LORA_SCALE = LORA_ALPHA / LORA_RANK          # peft: scaling = lora_alpha / r = 0.125


# ----------------------------------------------------------------------------- spec helpers
class Spec(list):
    def add(self, name, shape, kind, fan_in=None):
        self.append((name, tuple(shape), kind, fan_in))

    def conv(self, name, cout, cin, k=1, bias=True, groups=1):
        fan = (cin // groups) * k * k
        self.add(name + ".weight", (cout, cin // groups, k, k), "w", fan)
        if bias:
            self.add(name + ".bias", (cout,), "b", fan)

    def linear(self, name, cout, cin, bias=True):
        self.add(name + ".weight", (cout, cin), "w", cin)
        if bias:
            self.add(name + ".bias", (cout,), "b", cin)

    def norm(self, name, c):
        self.add(name + ".weight", (c,), "g")
        self.add(name + ".bias", (c,), "beta")

    # peft-wrapped layers
    def lora_conv(self, name, cout, cin, k=1, bias=True):
        self.conv(name + ".base_layer", cout, cin, k, bias)
        self.add(name + ".lora_A.default.weight", (LORA_RANK, cin, k, k), "lora")
        self.add(name + ".lora_B.default.weight", (cout, LORA_RANK, 1, 1), "lora")

    def lora_linear(self, name, cout, cin, bias=True):
        self.linear(name + ".base_layer", cout, cin, bias)
        self.add(name + ".lora_A.default.weight", (LORA_RANK, cin), "lora")
        self.add(name + ".lora_B.default.weight", (cout, LORA_RANK), "lora")


# ----------------------------------------------------------------------------- codec (IntraNoAR decode side)
def _dcb4(s, p, cin, cout):
    """DepthConvBlock4 = DepthConv + ConvFFN3 (modules/dcvc.py:242-265, 353-367, 424-433)."""
    s.conv(p + ".block.0.conv1.0", cin, cin, 1)
    s.conv(p + ".block.0.depth_conv", cin, cin, 3, groups=cin)
    s.conv(p + ".block.0.conv2", cout, cin, 1)
    if cin != cout:
        s.conv(p + ".block.0.adaptor", cout, cin, 1)
    s.conv(p + ".block.1.conv", cout * 4, cout, 1)
    s.conv(p + ".block.1.conv_out", cout, cout * 2, 1)


def _rbu(s, p, cin, cout):
    """ResidualBlockUpsample (modules/dcvc.py:183-208)."""
    s.conv(p + ".subpel_conv.0", cout * 4, cin, 1)
    s.conv(p + ".conv", cout, cout, 3)
    s.conv(p + ".upsample.0", cout * 4, cin, 1)


def _vq_res(s, p, c):
    """vqgan ResnetBlock, bias-free convs (modules/vqgan/blocks.py:15-52)."""
    s.norm(p + ".norm1", c)
    s.conv(p + ".conv1", c, c, 3, bias=False)
    s.norm(p + ".norm2", c)
    s.conv(p + ".conv2", c, c, 3, bias=False)


def _vq_attn(s, p, c):
    """vqgan AttnBlock (modules/vqgan/blocks.py:55-107)."""
    s.norm(p + ".norm", c)
    for n in ("q", "k", "v", "proj_out"):
        s.conv(p + "." + n, c, c, 1)


def codec_spec(N=128, internal=512, sem=768, ctrl=320, zdim=7):
    s = Spec()
    # Decoder (g_s), codec_module.py:88-116
    _dcb4(s, "dec.trans_coding.0", N, internal)
    _dcb4(s, "dec.trans_coding.1", internal, internal)
    for i in range(3):
        _vq_res(s, f"dec.blocks.{i}", internal)
    s.conv("dec.blocks.3", internal * 4, internal, 1)
    s.conv("dec.blocks.5", internal // 2, internal, 3)
    for i in (6, 7, 8):
        _vq_res(s, f"dec.blocks.{i}", internal // 2)
    _rbu(s, "dec.sem_up.0", sem, internal)
    _dcb4(s, "dec.sem_up.1", internal, internal)
    _rbu(s, "dec.sem_up.2", internal, internal // 2)
    _dcb4(s, "dec.sem_up.3", internal // 2, internal // 2)
    _rbu(s, "dec.sem_up.4", internal // 2, internal // 2)
    _dcb4(s, "dec.conv_out", internal, ctrl)
    # SemanticAdaptor, codec_module.py:169-181
    _dcb4(s, "semantic_adaptor.to_semantic.0", N, sem)
    _vq_res(s, "semantic_adaptor.to_semantic.1", sem)
    _vq_attn(s, "semantic_adaptor.to_semantic.2", sem)
    _vq_attn(s, "semantic_adaptor.to_semantic.3", sem)
    _vq_res(s, "semantic_adaptor.to_semantic.4", sem)
    _vq_attn(s, "semantic_adaptor.to_semantic.5", sem)
    _vq_attn(s, "semantic_adaptor.to_semantic.6", sem)
    _dcb4(s, "semantic_adaptor.to_semantic.7", sem, sem)
    # HyperDecoder, codec_module.py:145-166
    s.conv("hyper_dec.feat_in.0", N, zdim, 1)
    _dcb4(s, "hyper_dec.to_entropy.0", N, N)
    _rbu(s, "hyper_dec.to_entropy.1", N, N)
    _dcb4(s, "hyper_dec.to_entropy.2", N, N)
    _rbu(s, "hyper_dec.to_entropy.3", N, N)
    _dcb4(s, "hyper_dec.to_entropy.4", N, N)
    # prior nets, codec_module.py:205-217
    _dcb4(s, "y_prior_fusion.0", N, 2 * N)
    _dcb4(s, "y_prior_fusion.1", 2 * N, 2 * N)
    s.conv("y_spatial_prior_reduction", N, 2 * N, 1)
    for i in (1, 2, 3):
        _dcb4(s, f"y_spatial_prior_adaptor_{i}", 2 * N, 2 * N)
    for i in range(3):
        _dcb4(s, f"y_spatial_prior.{i}", 2 * N, 2 * N)
    return s


# ----------------------------------------------------------------------------- UNet (SD1.5 + LoRA + codec patches)
UNET_CH = (320, 640, 1280, 1280)
UNET_CTX = 768
UNET_TEMB = 1280
UNET_HEADS = 8


def _unet_res(s, p, cin, cout):
    s.norm(p + ".norm1", cin)
    s.lora_conv(p + ".conv1", cout, cin, 3)
    s.lora_linear(p + ".time_emb_proj", cout, UNET_TEMB)
    s.norm(p + ".norm2", cout)
    s.lora_conv(p + ".conv2", cout, cout, 3)
    if cin != cout:
        s.lora_conv(p + ".conv_shortcut", cout, cin, 1)


def _unet_tr(s, p, c):
    s.norm(p + ".norm", c)
    s.lora_conv(p + ".proj_in", c, c, 1)
    t = p + ".transformer_blocks.0"
    s.norm(t + ".norm1", c)
    for n in ("to_q", "to_k", "to_v"):
        s.lora_linear(t + ".attn1." + n, c, c, bias=False)
    s.lora_linear(t + ".attn1.to_out.0", c, c)
    s.norm(t + ".norm2", c)
    s.lora_linear(t + ".attn2.to_q", c, c, bias=False)
    s.lora_linear(t + ".attn2.to_k", c, UNET_CTX, bias=False)
    s.lora_linear(t + ".attn2.to_v", c, UNET_CTX, bias=False)
    s.lora_linear(t + ".attn2.to_out.0", c, c)
    s.norm(t + ".norm3", c)
    s.lora_linear(t + ".ff.net.0.proj", 8 * c, c)
    s.lora_linear(t + ".ff.net.2", c, 4 * c)
    s.lora_conv(p + ".proj_out", c, c, 1)


def unet_spec(in_ch=320, vae_ch=4):
    s = Spec()
    s.conv("conv_in", UNET_CH[0], in_ch, 3)                       # decoder_unet.py:391-393 (not LoRA'd)
    s.linear("time_embedding.linear_1", UNET_TEMB, UNET_CH[0])
    s.linear("time_embedding.linear_2", UNET_TEMB, UNET_TEMB)
    cin = UNET_CH[0]
    for i, c in enumerate(UNET_CH):
        for j in range(2):
            _unet_res(s, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c)
            if i < 3:
                _unet_tr(s, f"down_blocks.{i}.attentions.{j}", c)
        if i < 3:
            s.lora_conv(f"down_blocks.{i}.downsamplers.0.conv", c, c, 3)
        cin = c
    _unet_res(s, "mid_block.resnets.0", 1280, 1280)
    _unet_tr(s, "mid_block.attentions.0", 1280)
    _unet_res(s, "mid_block.resnets.1", 1280, 1280)
    # skip stack (decoder_unet.py:204-226): channels pushed in order
    skips = [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
    rev = list(reversed(UNET_CH))                                  # 1280,1280,640,320
    prev = 1280
    for i, c in enumerate(rev):
        for j in range(3):
            sk = skips.pop()
            _unet_res(s, f"up_blocks.{i}.resnets.{j}", prev + sk, c)
            prev = c
            if i > 0:
                _unet_tr(s, f"up_blocks.{i}.attentions.{j}", c)
        if i < 3:
            s.lora_conv(f"up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    s.norm("conv_norm_out", 320)
    s.conv("conv_out", vae_ch, 320, 3)
    # reduce_resblock (decoder_unet.py:14-29)
    s.norm("vae_reduction.blocks.0", in_ch)
    s.conv("vae_reduction.blocks.2", in_ch, in_ch, 3)
    s.norm("vae_reduction.blocks.3", in_ch)
    s.conv("vae_reduction.blocks.5", vae_ch, in_ch, 3)
    s.conv("vae_reduction.short_cut", vae_ch, in_ch, 1)
    return s


# ----------------------------------------------------------------------------- VAE decoder (SD2.1 KL-VAE)
VAE_CH = (128, 256, 512, 512)


def _vae_res(s, p, cin, cout):
    s.norm(p + ".norm1", cin)
    s.conv(p + ".conv1", cout, cin, 3)
    s.norm(p + ".norm2", cout)
    s.conv(p + ".conv2", cout, cout, 3)
    if cin != cout:
        s.conv(p + ".conv_shortcut", cout, cin, 1)


def vae_spec(latent=4):
    s = Spec()
    s.conv("post_quant_conv", latent, latent, 1)
    d = "decoder"
    s.conv(d + ".conv_in", 512, latent, 3)
    _vae_res(s, d + ".mid_block.resnets.0", 512, 512)
    a = d + ".mid_block.attentions.0"
    s.norm(a + ".group_norm", 512)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s.linear(a + "." + n, 512, 512)
    _vae_res(s, d + ".mid_block.resnets.1", 512, 512)
    prev = 512
    for i, c in enumerate(reversed(VAE_CH)):                       # 512,512,256,128
        for j in range(3):
            _vae_res(s, f"{d}.up_blocks.{i}.resnets.{j}", prev, c)
            prev = c
        if i < 3:
            s.conv(f"{d}.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    s.norm(d + ".conv_norm_out", 128)
    s.conv(d + ".conv_out", 3, 128, 3)
    return s


# ----------------------------------------------------------------------------- init
def random_state_dict(spec, seed=0):
    sd = {}
    for name, shape, kind, fan_in in spec:
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        if kind in ("w", "b"):
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "g":
            t = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        elif kind == "beta":
            t = 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        elif kind == "lora":
            t = torch.randn(shape, generator=g) * 0.02
        else:
            raise ValueError(kind)
        sd[name] = t
    return sd


def merge_lora(sd, name):
    """W' = W + scaling * B @ A  (peft Linear / Conv2d LoRA; decoder_unet.py:331-368).
    Conv: A is (r, cin, k, k), B is (cout, r, 1, 1)."""
    key = name + ".base_layer.weight"
    if key not in sd:
        return sd[name + ".weight"].float(), sd.get(name + ".bias")
    w = sd[key].float()
    a = sd[name + ".lora_A.default.weight"].float()
    b = sd[name + ".lora_B.default.weight"].float()
    delta = (b.flatten(1) @ a.flatten(1)).reshape(w.shape)
    return w + LORA_SCALE * delta, sd.get(name + ".base_layer.bias")


# ----------------------------------------------------------------------------- nn.Module-shaped loading
class IncompatibleKeys(tuple):
    """What torch's load_state_dict returns (the reference prints it, inference.py:92-93)."""

    def __new__(cls, missing_keys, unexpected_keys):
        self = super().__new__(cls, (missing_keys, unexpected_keys))
        self.missing_keys, self.unexpected_keys = missing_keys, unexpected_keys
        return self

    def __repr__(self):
        if not self.missing_keys and not self.unexpected_keys:
            return "<All keys matched successfully>"
        return f"IncompatibleKeys(missing_keys={self.missing_keys}, unexpected_keys={self.unexpected_keys})"


class LazyNet:
    """Weight handling of a kernel-launch network, shaped like torch.nn.Module where the reference touches it
    (src/inference.py:69-72,87-93): construction without weights, `load_state_dict(sd, strict=True)`, `eval()`.

    The kernel-layout weights (LoRA merged, epilogue permutations, bf16) are packed by `_build(sd)`; packing ~1 G
    parameters takes seconds, so it is deferred until the weights are known: `load_state_dict` builds at once, and a
    network that is used without ever being loaded builds itself from the seed-0 random initialisation (the analogue
    of a freshly constructed nn.Module).  Subclasses define `_spec()`, `_build(sd)` and `IGNORED_PREFIXES`, the
    checkpoint keys that belong to parts of the reference module that are not on the decode path."""

    IGNORED_PREFIXES = ()

    def _lazy_init(self, state_dict=None):
        d = self.__dict__
        d["_pending_sd"], d["_built"], d["_on_load"], d["weights_version"] = state_dict, False, [], 0

    def _ensure(self):
        if not self.__dict__["_built"]:
            sd = self.__dict__["_pending_sd"]
            if sd is None:
                sd = random_state_dict(self._spec(), 0)
            self.__dict__["_pending_sd"] = None
            self._build(sd)
            self.__dict__["_built"] = True
        return self

    def __getattr__(self, name):
        # reached only when normal lookup fails: attributes created by _build() before the first use
        d = self.__dict__
        if not name.startswith("__") and "_built" in d and not d["_built"]:
            self._ensure()
            if name in self.__dict__:
                return self.__dict__[name]
        raise AttributeError(f"{type(self).__name__!s} has no attribute {name!r}")

    def load_state_dict(self, state_dict, strict=True):
        spec = {n: shape for n, shape, _, _ in self._spec()}
        missing = [n for n in spec if n not in state_dict]
        unexpected = [k for k in state_dict if k not in spec and not k.startswith(tuple(self.IGNORED_PREFIXES))]
        errors = [f"size mismatch for {n}: checkpoint {tuple(state_dict[n].shape)}, model {spec[n]}"
                  for n in spec if n in state_dict and tuple(state_dict[n].shape) != spec[n]]
        if strict and missing:
            errors.append("Missing key(s) in state_dict: " + ", ".join(repr(k) for k in missing[:8])
                          + (" ..." if len(missing) > 8 else ""))
        if strict and unexpected:
            errors.append("Unexpected key(s) in state_dict: " + ", ".join(repr(k) for k in unexpected[:8])
                          + (" ..." if len(unexpected) > 8 else ""))
        if errors:
            raise RuntimeError(f"Error(s) in loading state_dict for {type(self).__name__}:\n\t" + "\n\t".join(errors))
        sd = {n: state_dict[n] for n in spec if n in state_dict}
        if missing:                                   # strict=False: absent tensors keep the seed-0 initialisation
            init = random_state_dict([e for e in self._spec() if e[0] in set(missing)], 0)
            sd.update(init)
        self.__dict__["_pending_sd"] = None
        self._build(sd)
        self.__dict__["_built"] = True
        self.__dict__["weights_version"] += 1
        for cb in self.__dict__["_on_load"]:
            cb()
        return IncompatibleKeys(missing, unexpected)

    def eval(self):
        return self

    def train(self, mode=True):
        assert not mode, "onedc_b200 is inference only"
        return self

    def to(self, *a, **k):
        return self

    def requires_grad_(self, flag=False):
        return self


def load_checkpoint_file(path, map_location="cpu"):
    """safetensors or torch file -> state dict (src/utils.py load_safetensor equivalent)."""
    if str(path).endswith(".safetensors"):
        from safetensors import safe_open
        with safe_open(path, framework="pt", device=map_location) as f:
            return {k: f.get_tensor(k) for k in f.keys()}
    import torch as _t
    return _t.load(path, map_location=map_location)


_VAE_LEGACY = {".query.": ".to_q.", ".key.": ".to_k.", ".value.": ".to_v.", ".proj_attn.": ".to_out.0."}


def vae_decoder_state_dict(sd):
    """Decode-side subset of an SD-VAE checkpoint in diffusers naming (older files name the mid attention
    query/key/value/proj_attn with 1x1-conv shaped weights)."""
    out = {}
    for k, v in sd.items():
        if not (k.startswith("decoder.") or k.startswith("post_quant_conv.")):
            continue
        for a, b in _VAE_LEGACY.items():
            k = k.replace(a, b)
        if ".attentions." in k and k.endswith(".weight") and v.dim() == 4:
            v = v[:, :, 0, 0]
        out[k] = v
    return out
