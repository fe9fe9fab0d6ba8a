from moge_b200.model import import_model_class_by_version  # noqa: F401
