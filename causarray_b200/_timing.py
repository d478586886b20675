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


@contextlib.contextmanager
def phase(name):
    if not ENABLED:
        yield
        return
    t0 = time.perf_counter()
    try:
        yield
    finally:
        TOTALS[name] = TOTALS.get(name, 0.0) + (time.perf_counter() - t0)
