"""CPU: host-side logic and the C-ABI surface (no compute calls)."""
import ctypes as C
import os
import re

import pytest
import torch

from moge_b200 import capi
from moge_b200.configs import model_config


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(os.path.dirname(capi._HERE), "include", "moge_b200.h")).read()
    declared = set(re.findall(r"\b(moge_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.moge_version()


def test_config_struct_roundtrip():
    c = capi.make_config(model_config("vitl"), capi.F16)
    assert (c.embed_dim, c.depth, c.num_heads) == (1024, 24, 16)
    assert list(c.taps)[:4] == [5, 11, 17, 23]
    assert list(c.neck.dim_in)[:5] == [1026, 2, 2, 2, 2]
    assert list(c.points_head.dim_out)[:5] == [0, 0, 0, 0, 3]
    assert list(c.neck.resamplers)[:4] == [0, 0, 0, 1]
    assert c.scale_head_layers == 3 and c.remap_output == 2
    c2 = capi.make_config(model_config("vitb", with_normal=False), capi.BF16)
    assert c2.normal_head.present == 0 and c2.compute_dtype == capi.BF16


def test_no_cpu_fallback():
    from moge.model.v2 import MoGeModel
    from moge_b200.synthetic import make_state_dict, synthetic_images
    cfg = model_config("vits")
    m = MoGeModel(**cfg)
    with pytest.raises(ValueError):
        MoGeModel(**{**cfg, "remap_output": "bogus"})
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m.load_state_dict({"x": torch.zeros(1)})
    with pytest.raises(capi.MogeError):
        m.infer(synthetic_images(1, 28, 28, 0))
    h = C.c_void_p()
    c = capi.make_config(cfg, capi.F16)
    assert capi.lib().moge_engine_create(C.byref(c), 0, C.byref(h)) != 0
    assert b"no CPU fallback" in capi.lib().moge_last_error()


def test_import_model_class_by_version():
    from moge.model import import_model_class_by_version
    from moge.model.v2 import MoGeModel
    assert import_model_class_by_version("v2") is MoGeModel
    with pytest.raises(NotImplementedError):
        import_model_class_by_version("v1")


def test_make_config_rejects_unsupported_decoders():
    cfg = model_config("vits")
    bad = {**cfg, "neck": {**cfg["neck"], "res_block_in_norm": "layer_norm"}}
    with pytest.raises(capi.MogeError):
        capi.make_config(bad, capi.F16)
    bad = {**cfg, "points_head": {**cfg["points_head"], "resamplers": ["pixel_shuffle"] * 4}}
    with pytest.raises(capi.MogeError):
        capi.make_config(bad, capi.F16)
    with pytest.raises(ValueError):
        capi.make_config({**cfg, "encoder": {**cfg["encoder"], "backbone": "dinov2_vitg14"}}, capi.F16)


def test_model_plumbing_without_gpu():
    """nn.Module-like surface of the drop-in class (v2.py:59-74,109-120): dtype/device bookkeeping, training-only stubs."""
    from moge.model.v2 import MoGeModel
    m = MoGeModel(**model_config("vits", with_normal=False), some_deprecated_kwarg=1) if False else MoGeModel(**model_config("vits", with_normal=False))
    assert not hasattr(m, "normal_head") and hasattr(m, "points_head") and hasattr(m, "scale_head")
    assert m.dtype == torch.float32 and m.half().dtype == torch.float16 and m.bfloat16().dtype == torch.bfloat16
    assert m.eval() is m and m.device.type == "cpu"
    with pytest.raises(NotImplementedError):
        m.init_weights()
    with pytest.raises(NotImplementedError):
        m.enable_gradient_checkpointing()
    with pytest.raises(NotImplementedError):
        m.onnx_compatible_mode = True
    assert m.onnx_compatible_mode is False
    with pytest.warns(UserWarning):
        MoGeModel(**model_config("vits"), deprecated_thing=3)          # v2.py:42-43


def test_synthetic_checkpoint_roundtrip(tmp_path):
    """The synthetic checkpoint is a reference-format file: {'model_config', 'model'} loadable with weights_only=True."""
    from moge.model.v2 import MoGeModel
    from moge_b200.synthetic import save_checkpoint, make_state_dict
    cfg = model_config("vits")
    path = tmp_path / "model.pt"
    save_checkpoint(path, cfg, seed=3)
    ck = torch.load(path, map_location="cpu", weights_only=True)
    assert set(ck) == {"model_config", "model"}
    m = MoGeModel.from_pretrained(path)
    sd = make_state_dict(cfg, 3)
    assert set(m.state_dict()) == set(sd) and all(torch.equal(m.state_dict()[k], sd[k]) for k in sd)
    assert m.num_tokens_range == [1200, 3600] and m.remap_output == "exp"
