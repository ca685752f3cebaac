"""``mu.tl`` -- multimodal tools.  ``mofa`` is on the hot path."""
from __future__ import annotations


def mofa(data, *args, **kwargs):
    """Multi-Omics Factor Analysis -- drop-in for ``muon.tl.mofa`` (muon/_core/tools.py:290-708)."""
    from ._mofa import mofa as _impl
    return _impl(data, *args, **kwargs)
