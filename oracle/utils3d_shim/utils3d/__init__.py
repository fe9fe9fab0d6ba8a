"""Minimal stand-in for the un-vendored `utils3d` dependency (pinned by the reference at
commit 3fab839f, /root/reference/pyproject.toml:23) so that the UNMODIFIED reference package can be
imported in the build container for golden-vector generation.  TEST INFRASTRUCTURE ONLY.

Only the two functions called on the infer() path are provided
(/root/reference/moge/model/v2.py:266 and :276).  Their semantics are restated from the published
utils3d API (normalized intrinsics, pixel-centre UV grid) -- no reference test pins them, so parity
at this boundary is "unpinned" (see DESIGN.md).
"""
from . import pt  # noqa: F401
from . import pt as torch  # noqa: F401  (utils3d exposes both names)
