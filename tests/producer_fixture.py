"""Session-wide cache of the producer-written inputs (reads -> reference packer -> reference bam2snapshot): the
pipeline takes ~30 s per config, so CPU and GPU tests share one copy."""
import os
import shutil
import tempfile

import pytest

import gt16_testlib as T

_cache = {}
_root = None


def producer_inputs(name):
    """(index_prefix, out_prefix) of PRODUCER_CONFIGS[name], made once per test session; skips when the reference
    tools are not built (oracle/_ref comes from /root/reference at build time and travels with the repository)"""
    global _root
    for exe in (T.REF_MKINDEX, T.REF_BAM2SNAPSHOT, T.READSIM):
        if not os.path.exists(exe):
            pytest.skip("%s not built" % os.path.relpath(exe, T.ROOT))
    if name not in _cache:
        if _root is None:
            _root = tempfile.mkdtemp(prefix="gt16_producer_")
        wd = os.path.join(_root, name)
        os.makedirs(wd, exist_ok=True)
        _cache[name] = T.regenerate_producer(name, wd)
    return _cache[name]


def cleanup():
    global _root
    if _root:
        shutil.rmtree(_root, ignore_errors=True)
        _root = None
    _cache.clear()
