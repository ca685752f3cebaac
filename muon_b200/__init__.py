"""muon_b200 -- B200-native (sm_100a) drop-in for muon's sparse hot path.

    import muon_b200 as mu
    mu.atac.pp.tfidf(adata)      # muon/_atac/preproc.py:16
    mu.atac.tl.lsi(adata)        # muon/_atac/tools.py:29
    mu.tl.mofa(mdata)            # muon/_core/tools.py:290
    mu.pp.neighbors(mdata)       # muon/_core/preproc.py:264  (WNN; first version, exact search)

Same signatures and AnnData/MuData slot semantics as the reference; the arithmetic runs in
hand-written CUDA kernels behind the C ABI in include/muon_b200.h.  There is no CPU fallback.
"""
from . import atac, pp, tl  # noqa: F401
from ._containers import SimpleAnnData, SimpleMuData  # noqa: F401
from ._device import DeviceCSR, release_all_resident, release_resident, trim_host_cache  # noqa: F401

__version__ = "0.1.0"
