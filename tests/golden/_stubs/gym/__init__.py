"""Minimal stand-in for `gym`, only so the read-only reference imports in the
authoring container when generating golden vectors (tests never import it)."""
from . import spaces


class Env:
    pass


def make(*a, **k):
    raise RuntimeError("gym stub: no environments")
