"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatements of the reference algorithms on the GCATE + doubly-robust LFC
hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import anything
from here, and only as the checker or the timed CPU baseline -- never on the
product path (``causarray_b200`` fails loudly when its CUDA library is missing
and never falls back to this package).

Parity status (see DESIGN.md "Oracle"):

* alternating minimisation, AIPW / LFC estimand, BH, size factors, subsamplers:
  pinned against the real reference (``/root/reference``, imported through
  ``oracle/ref_shim.py`` in the build container) -- golden vectors under
  ``tests/golden`` produced by ``tests/golden/make_golden.py``.
* per-gene GLM (IRLS): the arithmetic lives in statsmodels (un-vendored, no
  version pin; ``environment.yaml`` says ``statsmodels>=0.14``), absent from
  ``/root/reference`` and from this image.  ``oracle/glm.py`` restates the
  published statsmodels 0.14 ``GLM._fit_irls`` algorithm and is anchored on the
  reference's call sites (``causarray/gcate_glm.py:187-236``) and on independent
  MLE cross-checks (scipy / sklearn).  PARITY UNPINNED for that piece.
"""
