"""Harness-side compatibility shims so the UNMODIFIED reference (petastorm 0.13.1, written for
numpy<2 / pyarrow 6) imports on this image (numpy 2.3 / pyarrow 24).  TEST INFRASTRUCTURE ONLY.

Runs at interpreter start-up when `oracle/shims` is on PYTHONPATH (also in spawned ProcessPool children).
Nothing under /root/reference is modified or copied."""
import sys
import types

try:
    import numpy as np
    for _name, _target in (('unicode_', np.str_), ('string_', np.bytes_), ('float', float)):
        if not hasattr(np, _name):
            setattr(np, _name, _target)
    import pyarrow
    for _mod, _attrs in (('pyarrow.filesystem', ['LocalFileSystem']),
                         ('pyarrow.hdfs', ['HadoopFileSystem', 'connect'])):
        if _mod not in sys.modules:
            _m = types.ModuleType(_mod)
            for _a in _attrs:
                setattr(_m, _a, type(_a, (), {}))
            sys.modules[_mod] = _m
            setattr(pyarrow, _mod.split('.')[1], _m)
    if not hasattr(pyarrow, 'localfs'):
        pyarrow.localfs = object()
except Exception:  # pragma: no cover - never break interpreter start-up
    pass
