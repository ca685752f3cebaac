"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's arithmetic for the hot path (and helpers that run the unmodified reference
functions in the build container).  Imported by ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of
``bench.py`` only -- never by ``muon_b200``; the product path has no CPU fallback.

  tfidf_ref.py    muon/_atac/preproc.py:92-119            pinned: reference goldens + unmodified reference run
  lsi_ref.py      muon/_atac/tools.py:42-69 + scipy svds   pinned by us (reference has no lsi test)
  mofa_ref.py     mofapy2 CAVI (third-party, absent)       PARITY UNPINNED vs mofapy2; structural reference test
  _refload.py     runs the reference's tfidf / lsi / neighbors in place with stubbed imports
  _third_party.py exact stand-ins for umap / pynndescent / scanpy pieces used by the WNN driver (round-2 groundwork)
"""
