"""Mirror of /root/reference/moge/model/__init__.py:9-18."""
import importlib


def import_model_class_by_version(version: str):
    assert version in ["v1", "v2"], f'Unsupported model version: {version}'
    if version == "v1":
        raise NotImplementedError("moge_b200 implements the MoGe-2 (v2) hot path only; MoGe-1 is listed as 'next' (SURVEY.md 8f N3)")
    module = importlib.import_module(".v2", __package__)
    return getattr(module, "MoGeModel")
