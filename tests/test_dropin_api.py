"""The drop-in surface: what /root/reference/src/inference.py does to the model object must work here with only the
import line changed.

CPU part: the nn.Module-shaped weight handling (construction without weights, load_state_dict(strict=True) semantics and
its printed result) and the reference signatures.  GPU part: the reference `Evaluator.__init__ / load / evaluate
(decoder_only)` call sequence (inference.py:58-72, 87-93, 102-108) replayed against onedc_b200.
"""
import inspect
import os
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ CPU
class _Tiny:
    pass


def _tiny_net():
    from onedc_b200.weights import LazyNet, Spec

    class Tiny(LazyNet):
        IGNORED_PREFIXES = ("enc.",)

        def __init__(self, sd=None):
            self.builds = []
            self._lazy_init(sd)

        def _spec(self):
            s = Spec()
            s.conv("a", 4, 2, 3)
            s.norm("n", 4)
            return s

        def _build(self, sd):
            self.builds.append({k: v.clone() for k, v in sd.items()})
            self.a_w = sd["a.weight"]

    return Tiny


def test_lazy_net_load_state_dict_semantics():
    from onedc_b200.weights import random_state_dict
    Tiny = _tiny_net()
    m = Tiny()
    assert m.builds == []                                   # nothing packed at construction
    sd = random_state_dict(m._spec(), 3)
    fired = []
    m._on_load.append(lambda: fired.append(1))
    r = m.load_state_dict({**sd, "enc.whatever": torch.zeros(1)}, strict=True)      # analysis-side keys are tolerated
    assert repr(r) == "<All keys matched successfully>" and r.missing_keys == [] and r.unexpected_keys == []
    assert len(m.builds) == 1 and torch.equal(m.a_w, sd["a.weight"]) and fired == [1] and m.weights_version == 1
    with pytest.raises(RuntimeError, match="Missing key"):
        m.load_state_dict({k: v for k, v in sd.items() if k != "n.bias"}, strict=True)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        m.load_state_dict({**sd, "bogus.weight": torch.zeros(1)}, strict=True)
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict({**sd, "a.weight": torch.zeros(4, 2, 1, 1)}, strict=False)
    r = m.load_state_dict({k: v for k, v in sd.items() if k != "n.bias"}, strict=False)
    assert r.missing_keys == ["n.bias"] and "n.bias" in m.builds[-1]
    assert m.eval() is m and m.requires_grad_(False) is m and m.to("cuda") is m


def test_lazy_net_builds_itself_from_seed0_when_never_loaded():
    from onedc_b200.weights import random_state_dict
    Tiny = _tiny_net()
    m = Tiny()
    w = m.a_w                                               # first attribute miss triggers the build
    assert len(m.builds) == 1 and torch.equal(w, random_state_dict(m._spec(), 0)["a.weight"])
    with pytest.raises(AttributeError):
        m.nope


def test_vae_checkpoint_key_mapping():
    from onedc_b200 import weights as W
    sd = W.random_state_dict(W.vae_spec(), 0)
    legacy = {}
    for k, v in sd.items():
        k2 = (k.replace(".to_q.", ".query.").replace(".to_k.", ".key.").replace(".to_v.", ".value.")
              .replace(".to_out.0.", ".proj_attn."))
        if ".attentions." in k and k.endswith(".weight") and v.dim() == 2:
            v = v[:, :, None, None]
        legacy[k2] = v
    legacy["encoder.conv_in.weight"] = torch.zeros(1)
    out = W.vae_decoder_state_dict(legacy)
    assert set(out) == set(sd) and all(torch.equal(out[k], sd[k]) for k in sd)


def _params(fn):
    return [p for p in inspect.signature(fn).parameters if p != "self"]


def test_signatures_match_the_reference():
    """Argument names and order of the entry points the reference calls (file:line in the asserts' comments)."""
    from onedc_b200.codec_module import IntraNoAR
    from onedc_b200.entropy_models import EntropyCoder, GaussianEncoder
    from onedc_b200.model import SD15_1step_codec_stage1
    # codec_module.py:185-186
    assert _params(IntraNoAR.__init__)[:6] == ["cond_ch", "ctrl_ch", "internal_ch", "bottleneck_ch", "unet_ch_config",
                                               "z_fsq_levels"]
    # compression_model.py:369-373
    assert _params(IntraNoAR.decompress_four_part_prior) == [
        "common_params", "y_spatial_prior_adaptor_1", "y_spatial_prior_adaptor_2", "y_spatial_prior_adaptor_3",
        "y_spatial_prior", "y_spatial_prior_reduction"]
    # compression_model.py:421-425
    assert _params(IntraNoAR.forward_four_part_prior_recon_with_z)[:7] == [
        "y", "common_params", "y_spatial_prior_adaptor_1", "y_spatial_prior_adaptor_2", "y_spatial_prior_adaptor_3",
        "y_spatial_prior", "y_spatial_prior_reduction"]
    # codec_module.py:357, :418 ; z_only/codec_module.py:295
    assert _params(IntraNoAR.decode) == ["fp", "stream"]
    assert _params(IntraNoAR._decompress)[:5] == ["bit_stream_y", "bit_stream_z", "pad_height", "pad_width", "bit_stream_caption"]
    assert _params(IntraNoAR.forward)[:4] == ["x", "cond", "fix_encoder", "fix_codec"]
    # model_sd15_with_codec_stage1.py:18, :296
    assert _params(SD15_1step_codec_stage1.__init__)[:2] == ["args", "accelerator"]
    assert _params(SD15_1step_codec_stage1.decode)[:2] == ["fp", "stream"]
    # entropy_models.py:355, :364, :371, :66, :85
    assert _params(GaussianEncoder.build_indexes) == ["scales", "skip_thres"]
    assert _params(GaussianEncoder.encode) == ["x", "scales", "skip_thres"]
    assert _params(GaussianEncoder.decode_stream) == ["scales", "dtype", "device", "skip_thres"]
    assert _params(EntropyCoder.encode_with_indexes) == ["symbols", "indexes", "cdf_group_index"]
    assert _params(EntropyCoder.decode_stream) == ["indexes", "cdf_group_index"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree only exists in the build container")
def test_signatures_against_live_reference_source():
    """Same check against the reference source itself (parsed, not imported)."""
    import ast
    from onedc_b200.codec_module import IntraNoAR
    from onedc_b200.model import SD15_1step_codec_stage1

    def ref_params(path, cls, fn):
        tree = ast.parse(open(os.path.join("/root/reference/src", path)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.ClassDef) and node.name == cls:
                for f in node.body:
                    if isinstance(f, ast.FunctionDef) and f.name == fn:
                        return [a.arg for a in f.args.args if a.arg != "self"]
        raise KeyError((cls, fn))

    cm = "models/sd15_onedc_codec_stage1/codec_module.py"
    comp = "modules/entropy/compression_model.py"
    top = "models/sd15_onedc_codec_stage1/model_sd15_with_codec_stage1.py"
    for mine, ref in ((IntraNoAR.__init__, ref_params(cm, "IntraNoAR", "__init__")),
                      (IntraNoAR.decode, ref_params(cm, "IntraNoAR", "decode")),
                      (IntraNoAR._decompress, ref_params(cm, "IntraNoAR", "_decompress")),
                      (IntraNoAR.decompress_four_part_prior, ref_params(comp, "CompressionModel", "decompress_four_part_prior")),
                      (IntraNoAR.forward_four_part_prior_recon_with_z,
                       ref_params(comp, "CompressionModel", "forward_four_part_prior_recon_with_z")),
                      (SD15_1step_codec_stage1.__init__, ref_params(top, "SD15_1step_codec_stage1", "__init__")),
                      (SD15_1step_codec_stage1.decode, ref_params(top, "SD15_1step_codec_stage1", "decode"))):
        assert _params(mine)[: len(ref)] == ref, (mine.__qualname__, _params(mine), ref)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_evaluator_call_sequence_decoder_only(cuda, tmp_path):
    """inference.py:58-72 (construct, prepare, update, debug), :87-93 (load, strict=True), :102-108 (decode fp)."""
    from safetensors.torch import save_file
    from onedc_b200 import weights as W
    from onedc_b200.model import SD15_1step_codec_stage1          # <- the only line that differs from the reference
    from onedc_b200.weights import load_checkpoint_file as load_safetensor
    sds = (W.random_state_dict(W.unet_spec(), 7), W.random_state_dict(W.codec_spec(), 7),
           W.random_state_dict(W.vae_spec(), 0))
    args = types.SimpleNamespace(
        use_codeformer=False, unet_ckpt=None, codec_ckpt=None, vae_ckpt=None, control_ckpt=None, unet_ckpt_lora=None,
        codeformer_ckpt=None, guidance_ckpt=None, decoder_only=True, vae_attn_patch=16, conditioning_timestep=999,
        codec=types.SimpleNamespace(internal_ch=512, bottleneck_ch=128, unet_ch_config=[512, 768, 768], z_fsq_levels=[4] * 7))
    accelerator = types.SimpleNamespace(device=cuda, is_main_process=True)
    # Evaluator.__init__
    model = SD15_1step_codec_stage1(args, accelerator)
    model.prepare()
    model.codec_model.update(force=True)
    model.codec_model.debug = False
    # Evaluator.load: model_1.safetensors carries the analysis-side keys too
    codec_file = {**sds[1], "enc.pix_emb.weight": torch.zeros(4), "hyper_enc.feat_in.0.weight": torch.zeros(4)}
    save_file({k: v.contiguous() for k, v in codec_file.items()}, str(tmp_path / "model_1.safetensors"))
    codec_sd = load_safetensor(str(tmp_path / "model_1.safetensors"), map_location="cpu")
    r1 = model.feedforward_model.load_state_dict(sds[0], strict=True)
    r2 = model.codec_model.load_state_dict(codec_sd, strict=True)
    print(r1, r2)
    assert repr(r1) == repr(r2) == "<All keys matched successfully>"
    with pytest.raises(RuntimeError, match="Missing key"):
        model.codec_model.load_state_dict({k: v for k, v in codec_sd.items() if k != "dec.blocks.3.weight"}, strict=True)
    model.codec_model.load_state_dict(codec_sd, strict=True)
    # Evaluator.evaluate(decoder_only)
    model.eval()
    direct = SD15_1step_codec_stage1(state_dicts=sds, device=cuda)
    direct.codec_model.update(force=True)
    stream, _ = direct.codec_model.compress_synthetic(192, 128, seed=9)
    bin_path = tmp_path / "img.bin"
    bin_path.write_bytes(stream)
    recon = model.decode(fp=str(bin_path))
    recon_norm = recon.clamp(-1., 1.) * 0.5 + 0.5
    assert recon.shape == (1, 3, 192, 128) and recon.dtype == torch.float32 and recon.is_cuda
    assert torch.equal(recon, direct.decode(stream=stream)), "loaded-by-state-dict model != model built from the same weights"
    assert float(recon_norm.min()) >= 0 and float(recon_norm.max()) <= 1
    # weights loaded after graphs were captured must invalidate them
    model.codec_model.load_state_dict(W.random_state_dict(W.codec_spec(), 8), strict=True)
    assert len(model._graphed) == 0
    # the codec-level entry points with the reference's own call shapes (codec_module.py:357-369, 418-454)
    model.codec_model.load_state_dict(codec_sd, strict=True)
    x_hat, y_sem, hw, phw, pad = model.codec_model.decode(stream=stream)
    assert x_hat.shape == (1, 320, 24, 16) and y_sem.shape == (1, 768, 3, 2) and hw == (192, 128) and phw == (192, 128)
    xb, ysb, _ = direct.codec_model.decode_batch([stream])
    assert torch.equal(x_hat.permute(0, 2, 3, 1), xb) and torch.equal(y_sem.permute(0, 2, 3, 1), ysb)


@pytest.mark.gpu
def test_z_only_forward_dict(cuda):
    """models/sd15_onedc_codec_z_only/codec_module.py:263-308 result keys, decoder half."""
    from onedc_b200 import weights as W
    from onedc_b200.model import SD15_1step_codec_stage1
    model = SD15_1step_codec_stage1(state_dicts=(None, W.random_state_dict(W.codec_spec(), 0), None), device=cuda)
    model.codec_model.update(force=True)
    z = torch.randint(0, 16384, (1, 2, 3), generator=torch.Generator().manual_seed(1), dtype=torch.int32)
    out = model.codec_model(None, None, fix_codec=True, z_vq_indices=z)
    assert set(out) == {"x_hat", "y_hat", "bit", "bpp", "bpp_y", "bpp_hard_y", "y_semantic", "z_semantic", "z_vq_indices",
                        "params_hat", "y_orig"}
    assert out["x_hat"].shape == (1, 320, 16, 24) and out["y_hat"].shape == (1, 128, 8, 12)
    assert out["y_semantic"].shape == (1, 768, 2, 3) and out["params_hat"].shape == (1, 256, 8, 12)
    assert float(out["bpp_hard_y"]) == 0.0
    x_hat, y_sem = model.codec_model.decode_z_only(z.to(cuda))
    assert torch.equal(out["x_hat"].permute(0, 2, 3, 1), x_hat)
