"""The C-ABI library loads without a GPU and exports every symbol include/pcm_b200.h declares;
the ctypes mirrors of the descriptor structs have the C sizes.  CPU only (no compute calls)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pcm_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    from pcm_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pcm_b200.h")).read()
    declared = set(re.findall(r"\b(pcm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pcm_version() >= 1


def test_struct_sizes_match_c(lib):
    from pcm_b200 import _lib
    src = '#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu", sizeof(pcm_gemm_desc), sizeof(pcm_wgrad_desc), sizeof(pcm_asrc), sizeof(pcm_bsrc), sizeof(pcm_kentry));return 0;}\n' % os.path.join(ROOT, "include", "pcm_b200.h")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", c, "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(_lib.GemmDesc), C.sizeof(_lib.WgradDesc), C.sizeof(_lib.ASrc),
                     C.sizeof(_lib.BSrc), C.sizeof(_lib.KEntry)]


def test_missing_library_fails_loudly(monkeypatch):
    from pcm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpcm_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()
