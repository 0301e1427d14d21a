"""The C-ABI library loads and exports every symbol include/pst_b200.h declares (no GPU needed)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from petastorm_b200 import native
    header = open(os.path.join(ROOT, 'include', 'pst_b200.h')).read()
    declared = set(re.findall(r'\b(pst_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    for sym in sorted(declared):
        assert hasattr(native.lib, sym), 'libpst_b200.so does not export {}'.format(sym)
    assert set(native.EXPORTED) == declared
    assert native.lib.pst_abi_version() == 2
    assert native.lib.pst_has_cuda() == 1


def test_open_errors_are_reported():
    import pytest
    from petastorm_b200 import native
    with pytest.raises(native.NativeLibraryError):
        native.ParquetFile('/nonexistent/file.parquet')
    p = os.path.join(ROOT, 'include', 'pst_b200.h')
    with pytest.raises(native.NativeLibraryError, match='PAR1'):
        native.ParquetFile(p)
