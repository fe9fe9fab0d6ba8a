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
