"""CPU: the drop-in boundary.  Parameter names, order and defaults of tfidf / binarize / lsi / mofa must be the
reference's (read from the reference sources with ast by tests/golden/make_golden.py -> signatures.json)."""
import inspect
import json
import os

import pytest

import muon_b200 as mu
from conftest import GOLDEN


@pytest.mark.parametrize("name,fn", [("tfidf", mu.atac.pp.tfidf), ("binarize", mu.atac.pp.binarize),
                                     ("lsi", mu.atac.tl.lsi), ("mofa", mu.tl.mofa), ("neighbors", mu.pp.neighbors)])
def test_signature_matches_reference(name, fn):
    ref = json.load(open(os.path.join(GOLDEN, "signatures.json")))[name]
    if name == "mofa":
        from muon_b200._mofa import mofa as fn          # mu.tl.mofa is a thin forwarder
    sig = inspect.signature(fn)
    pos = [p for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert [p.name for p in pos] == ref["params"]
    for p, d, i in zip(pos, ref["defaults"], range(len(pos))):
        if i < ref["required"]:
            assert p.default is inspect.Parameter.empty
        else:
            assert repr(p.default) == d, (name, p.name, p.default, d)
    # anything we add on top must be keyword-only so positional calls mean the same thing
    extra = [p for p in sig.parameters.values() if p.kind not in (p.POSITIONAL_OR_KEYWORD,)]
    assert all(p.kind == p.KEYWORD_ONLY for p in extra)


def test_namespaces_mirror_muon():
    # muon/__init__.py:1-16, muon/atac.py:1, muon/_atac/__init__.py:1-4
    assert callable(mu.atac.pp.tfidf) and callable(mu.atac.tl.lsi) and callable(mu.tl.mofa)
