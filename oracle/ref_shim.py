"""Import the REAL reference package (``causarray``), unmodified.

Test infrastructure only: golden-vector generation, validation of the restatements in this
package, the full-size parity tools and the ``--impl reference`` arm of ``bench.py``.  The product
package (``causarray_b200``) never imports this module.

Where the reference comes from:

* ``/root/reference`` (or ``$CAUSARRAY_REFERENCE_ROOT``) in the build container;
* ``oracle/_ref/`` on the GPU box: ``tools/stage_ref.sh`` copies the package there (git-ignored, it
  travels with the gpurun snapshot) because ``/root/reference`` does not exist on that box.

The reference imports several packages that are not in the image (statsmodels, matplotlib,
sklearn_ensemble_cv).  ``oracle/shims/`` holds importable stand-ins and is appended to the END of
``sys.path`` and to ``PYTHONPATH`` (so joblib's worker processes find them as well; an installed
package of the same name wins):

* ``statsmodels.api.GLM`` / ``families`` -> ``oracle.glm`` restatement of statsmodels' IRLS (the
  reference's per-gene fit, ``causarray/gcate_glm.py:204``) -- parity unpinned for that piece;
* ``statsmodels.stats.multitest.multipletests`` -> BH restatement (``causarray/DR_inference.py:186``);
* matplotlib / sklearn_ensemble_cv -> inert (plots / random forests are out of scope).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SHIMS = os.path.join(HERE, "shims")


def _candidates():
    env = os.environ.get("CAUSARRAY_REFERENCE_ROOT")
    if env:
        yield env
    yield "/root/reference"
    yield os.path.join(HERE, "_ref")


def reference_root():
    for c in _candidates():
        if os.path.isdir(os.path.join(c, "causarray")):
            return c
    return None


REFERENCE_ROOT = reference_root()


def reference_available():
    return reference_root() is not None


def install_stubs():
    """Make the stand-in packages importable here and in child processes."""
    for path in (ROOT, SHIMS):
        if path not in sys.path:
            sys.path.append(path)
    pp = os.environ.get("PYTHONPATH", "")
    parts = [q for q in pp.split(os.pathsep) if q]
    for path in (ROOT, SHIMS):
        if path not in parts:
            parts.append(path)
    os.environ["PYTHONPATH"] = os.pathsep.join(parts)


def import_reference():
    """Return the reference ``causarray`` package."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference package not found (looked at $CAUSARRAY_REFERENCE_ROOT, /root/reference, "
                           "oracle/_ref -- run tools/stage_ref.sh in the build container)")
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    os.environ.setdefault("KMP_WARNINGS", "0")
    import causarray  # noqa: E402
    return causarray
