"""``mu.atac`` namespace: ``pp`` (preprocessing) and ``tl`` (tools), as in muon/_atac/__init__.py:1-4."""
from . import pp, tl  # noqa: F401
