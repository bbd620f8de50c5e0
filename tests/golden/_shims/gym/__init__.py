"""Minimal stand-in for `gym` so the reference env modules import in this container.

Test infrastructure only (used by tests/golden/make_golden.py in the build
container, where /root/reference exists). Never imported by product code.
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def __init__(self, *a, **k):
        pass
