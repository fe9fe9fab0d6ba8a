from moge_b200.model.v2 import MoGeModel  # noqa: F401
