"""Opt-in wall-clock phase accounting (host side; device work is included because every
phase ends in a synchronising C-ABI call).  ``bench.py`` uses it for the per-phase table."""
import contextlib
import time

ENABLED = False
TOTALS = {}


def enable(on=True):
    global ENABLED
    ENABLED = bool(on)


def reset():
    TOTALS.clear()


_NVTX = None


def _nvtx():
    """(push, pop) of the library's NVTX helpers, or False when the library is not loaded yet."""
    global _NVTX
    if _NVTX is None:
        try:
            from . import _lib as L
            lib = L._lib
            _NVTX = (lib.ca_nvtx_push, lib.ca_nvtx_pop) if lib is not None else None
        except Exception:  # noqa: BLE001
            _NVTX = False
    return _NVTX


@contextlib.contextmanager
def phase(name):
    """Host-side phase: always an NVTX range (visible in nsys / ncu --nvtx), wall-clock accounting
    when enabled."""
    nv = _nvtx()
    if nv:
        nv[0](name.encode())
    t0 = time.perf_counter() if ENABLED else 0.0
    try:
        yield
    finally:
        if ENABLED:
            TOTALS[name] = TOTALS.get(name, 0.0) + (time.perf_counter() - t0)
        if nv:
            nv[1]()
