"""TEST INFRASTRUCTURE ONLY -- never imported by the product (onedc_b200/).

Imports the *unmodified* reference codec half (IntraNoAR and the entropy model,
/root/reference/src) in this container so the oracle restatement can be pinned
against it and golden vectors can be generated (tests/golden/gen_golden.py).

Only usable where /root/reference exists (the build container); the GPU box
never has it, so nothing in `-m gpu` tests / smoke() / bench.py calls this.

Recipe:
  * reference pybind modules MLCodec_rans / MLCodec_CXX are compiled by
    oracle/Makefile from the sources where they lie into oracle/_ref/modules/entropy/
  * `modules.entropy` has no __init__.py => namespace package, so putting both
    oracle/_ref and /root/reference/src on sys.path merges the two directories
  * third-party packages that are absent here and unused on the decode path are
    replaced by name-only shims (pytorch_msssim, diffusers, vector_quantize_pytorch.FSQ)
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF_SRC = "/root/reference/src"
OVERLAY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available():
    return os.path.isdir(REF_SRC) and os.path.isdir(os.path.join(OVERLAY, "modules", "entropy"))


class _FSQ(nn.Module):
    """Restatement of vector_quantize_pytorch==1.21.2 FSQ.indices_to_codes for
    levels=[4]*7 (call site codec_module.py:431).  level_i = (idx // 4^i) % 4,
    code = (level - 2) / 2  in {-1, -.5, 0, .5}; output is channel-first."""

    def __init__(self, levels):
        super().__init__()
        self.levels = list(levels)
        self.codebook_size = 1
        for l in self.levels:
            self.codebook_size *= l
        basis = [1]
        for l in self.levels[:-1]:
            basis.append(basis[-1] * l)
        self.register_buffer("_basis", torch.tensor(basis, dtype=torch.int64), persistent=False)
        self.register_buffer("_levels", torch.tensor(self.levels, dtype=torch.int64), persistent=False)

    def indices_to_codes(self, indices):
        lv = (indices.unsqueeze(-1) // self._basis) % self._levels          # b h w d
        half = self._levels // 2
        codes = (lv - half).float() / half.float()
        return codes.permute(0, 3, 1, 2).contiguous()


def _shim(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


_cached = {}


def import_reference():
    """Returns a namespace with the reference modules (codec_module, compression_model,
    entropy_models, entropy utils, dcvc, vqgan blocks)."""
    if _cached:
        return _cached["ns"]
    assert reference_available(), "reference sources / oracle/_ref not present (run `make -C oracle ref`)"
    for p in (REF_SRC, OVERLAY):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "pytorch_msssim" not in sys.modules:
        _shim("pytorch_msssim", MS_SSIM=type("MS_SSIM", (nn.Module,), {
            "__init__": lambda s, *a, **k: nn.Module.__init__(s)}))
    if "vector_quantize_pytorch" not in sys.modules:
        _shim("vector_quantize_pytorch", FSQ=_FSQ)
    if "diffusers" not in sys.modules:
        _shim("diffusers", UNet2DModel=object)
        _shim("diffusers.models")
        _shim("diffusers.models.unets")
        _shim("diffusers.models.unets.unet_2d_blocks", AttnDownBlock2D=object)
        _shim("diffusers.models.unets.unet_2d", UNet2DOutput=object)
        _shim("diffusers.utils", USE_PEFT_BACKEND=False, BaseOutput=object, deprecate=None, logging=None,
              scale_lora_layers=None, unscale_lora_layers=None, is_torch_version=None)
    import models.sd15_onedc_codec_stage1.encoder_unet as eu
    eu.prepare_unet_encoder = lambda *a, **k: nn.Identity()
    import models.sd15_onedc_codec_stage1.codec_module as cm
    cm.prepare_unet_encoder = eu.prepare_unet_encoder
    import modules.entropy.compression_model as comp
    import modules.entropy.entropy_models as em
    import modules.entropy.utils as eutils
    import modules.dcvc as dcvc
    import modules.vqgan.blocks as vq
    ns = types.SimpleNamespace(codec_module=cm, compression_model=comp, entropy_models=em,
                               entropy_utils=eutils, dcvc=dcvc, vqgan_blocks=vq, FSQ=_FSQ)
    _cached["ns"] = ns
    return ns


def build_reference_codec():
    """Reference IntraNoAR with the inference config (config_inference.yaml:25-29), eval, CDFs built."""
    ns = import_reference()
    m = ns.codec_module.IntraNoAR(4, 320, 512, 128, [512, 768, 768], [4] * 7).eval()
    m.update(force=True)
    return m
