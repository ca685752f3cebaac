"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference functions from /root/reference.

``import muon`` is impossible in the build container (anndata, mudata, scanpy, h5py are
absent and there is no network).  The two hot-path functions we need as a live oracle,
``muon._atac.preproc.tfidf`` and ``muon._atac.tools.lsi``, only touch those packages for
``isinstance`` checks, ``view_to_actual`` and logging.  This module registers tiny stub
modules under those names (AnnData/MuData stubs are our duck-typed containers) and then
executes the reference source files *in place* from /root/reference -- nothing is copied.

Only ``tests/golden/make_golden.py`` and CPU tests (when /root/reference exists) use this.
It is never importable from the product path and does not exist on the GPU box.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("MUON_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "muon", "_atac", "preproc.py"))


def _install_stubs():
    from muon_b200._containers import SimpleAnnData, SimpleMuData, view_to_actual

    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__dict__["__stub__"] = True
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    if "anndata" not in sys.modules or getattr(sys.modules["anndata"], "__stub__", False):
        mod("anndata", AnnData=SimpleAnnData)
    if "mudata" not in sys.modules or getattr(sys.modules["mudata"], "__stub__", False):
        mod("mudata", MuData=SimpleMuData)
    if "scanpy" not in sys.modules or getattr(sys.modules["scanpy"], "__stub__", False):
        log = mod("scanpy.logging", info=lambda *a, **k: None, warning=lambda *a, **k: None,
                  hint=lambda *a, **k: None, debug=lambda *a, **k: None, error=lambda *a, **k: None)
        utils = mod("scanpy._utils", view_to_actual=view_to_actual)
        sc = mod("scanpy", logging=log, _utils=utils)
        sc.__path__ = []  # mark as package so "from scanpy import logging" resolves


def _load(relpath: str, fullname: str):
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(fullname, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[fullname] = m
    spec.loader.exec_module(m)
    return m


def load_reference_tfidf():
    """Return the reference's own ``tfidf`` function object (muon/_atac/preproc.py:16)."""
    if not reference_available():
        raise FileNotFoundError(REF_ROOT)
    _install_stubs()
    return _load("muon/_atac/preproc.py", "_refmuon_atac_preproc").tfidf


def load_reference_lsi():
    """Return the reference's own ``lsi`` function object (muon/_atac/tools.py:29)."""
    if not reference_available():
        raise FileNotFoundError(REF_ROOT)
    _install_stubs()
    from muon_b200._containers import SimpleMuData
    # tools.py does ``from . import utils`` and ``from .._rna.utils import ...``: give it a
    # throw-away package context whose sub-modules are inert.
    for name in ("_refmuon", "_refmuon._atac", "_refmuon._rna"):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = []
            sys.modules[name] = pkg
    sys.modules["_refmuon"].MuData = SimpleMuData
    if "_refmuon._atac.utils" not in sys.modules:
        sys.modules["_refmuon._atac.utils"] = types.ModuleType("_refmuon._atac.utils")
        sys.modules["_refmuon._atac"].utils = sys.modules["_refmuon._atac.utils"]
    if "_refmuon._rna.utils" not in sys.modules:
        ru = types.ModuleType("_refmuon._rna.utils")
        ru.get_gene_annotation_from_rna = lambda *a, **k: None
        sys.modules["_refmuon._rna.utils"] = ru
    m = sys.modules.get("_refmuon._atac.tools")
    if m is None:
        m = _load("muon/_atac/tools.py", "_refmuon._atac.tools")
    return m.lsi


def load_reference_neighbors():
    """Return the reference's own ``neighbors`` (WNN, muon/_core/preproc.py:264) with its third-party imports
    (umap, pynndescent, scanpy) replaced by the exact stand-ins of oracle/_third_party.py.  Round-2 oracle for the
    WNN row: the reference's control flow with exact instead of approximate nearest-neighbour search."""
    if not reference_available():
        raise FileNotFoundError(REF_ROOT)
    _install_stubs()
    import importlib.metadata as ilm

    from . import _third_party as tp

    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None or not getattr(m, "__stub__", False):
            m = types.ModuleType(name)
            m.__dict__["__stub__"] = True
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    mod("scanpy.tools")
    mod("scanpy.tools._utils", _choose_representation=tp.choose_representation)
    mod("scanpy.neighbors")
    mod("scanpy.neighbors._connectivity", umap=tp.umap_connectivities)
    mod("pynndescent")
    mod("pynndescent.distances", euclidean=tp.euclidean)
    mod("pynndescent.sparse", sparse_euclidean=tp.sparse_euclidean, sparse_jaccard=tp.sparse_jaccard)
    mod("umap")
    mod("umap.umap_", nearest_neighbors=tp.nearest_neighbors)
    m = sys.modules.get("_refmuon_core_preproc")
    if m is None:
        real_version = ilm.version

        def fake_version(name):                     # preproc.py:29-40 branches on the scanpy version
            return "1.10.4" if name == "scanpy" else real_version(name)

        ilm.version = fake_version
        try:
            m = _load("muon/_core/preproc.py", "_refmuon_core_preproc")
        finally:
            ilm.version = real_version
    return m.neighbors
