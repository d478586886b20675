"""Import the REAL reference (``/root/reference``) in the build container.

Test infrastructure only (golden-vector generation and validation of the
restatements in this package).  ``/root/reference`` does not exist on the GPU
box, so nothing that runs there may import this module.

The reference imports several packages that are not in the image
(matplotlib, statsmodels, sklearn_ensemble_cv).  They are stubbed in
``sys.modules`` *before* ``import causarray``:

* ``statsmodels.api.GLM`` / ``families`` -> ``oracle.glm`` restatement of
  statsmodels' IRLS (the reference's per-gene fit, ``causarray/gcate_glm.py:204``);
* ``statsmodels.stats.multitest.multipletests`` -> BH restatement
  (``causarray/DR_inference.py:186``);
* matplotlib / sklearn_ensemble_cv -> inert stubs (plots / random forests are
  out of scope).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CAUSARRAY_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "causarray"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    if "causarray" in sys.modules:
        return
    from oracle import glm as _glm
    from oracle import inference as _inf

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
            from mpl_toolkits.axes_grid1 import host_subplot  # noqa: F401
        except Exception:
            mpl = _stub("matplotlib")
            mpl.pyplot = _stub("matplotlib.pyplot")
            tk = _stub("mpl_toolkits")
            tk.axes_grid1 = _stub("mpl_toolkits.axes_grid1", host_subplot=lambda *a, **k: None)
    try:
        import statsmodels.api  # noqa: F401
    except Exception:
        sm = _stub("statsmodels")
        api = _stub("statsmodels.api", GLM=_glm.SMGLM, families=_glm.families, OLS=None)
        sm.api = api
        st = _stub("statsmodels.stats")
        mt = _stub("statsmodels.stats.multitest", multipletests=_inf.multipletests,
                   fdrcorrection=_inf.fdrcorrection)
        sm.stats = st
        st.multitest = mt
    try:
        import sklearn_ensemble_cv  # noqa: F401
    except Exception:
        _stub("sklearn_ensemble_cv", reset_random_seeds=lambda seed: None, Ensemble=object, ECV=object)


def import_reference():
    """Return the reference ``causarray`` package (build container only)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    os.environ.setdefault("KMP_WARNINGS", "0")
    import causarray  # noqa: E402
    return causarray
