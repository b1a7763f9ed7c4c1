"""Stub pyplot (never called by the golden generator)."""


def __getattr__(name):
    raise RuntimeError("matplotlib is stubbed out in tests/golden/ref_shims")
