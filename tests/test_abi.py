"""CPU: the C-ABI library loads and exports every symbol include/ttt_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header="ttt_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ttt_b200_\w+)\s*\(", src)))


def test_header_symbols_are_exported():
    import __graft_entry__
    __graft_entry__.build()
    from ttt_video_dit_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 4
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ttt_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == names, "python binding table out of sync with the header"
    assert _lib.lib().ttt_b200_version() >= 100


def test_debug_probes_live_in_their_own_library():
    """include/ttt_b200_debug.h <-> libttt_b200_selftest.so; the production library exports no ttt_b200_debug_* symbol."""
    import subprocess
    import __graft_entry__
    __graft_entry__.build()
    from ttt_video_dit_b200 import _lib
    D = ctypes.CDLL(_lib.DEBUG_LIB_PATH)
    names = _declared("ttt_b200_debug.h")
    assert names == sorted(_lib.debug_exported_symbols())
    for n in names:
        assert hasattr(D, n), f"{n} declared in include/ttt_b200_debug.h but not exported"
    dyn = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\b(ttt_b200_\w+)", dyn)))
    assert exported == _declared(), "production library must export exactly what include/ttt_b200.h declares"


def test_argument_errors_do_not_need_a_gpu():
    from ttt_video_dit_b200 import _lib
    L = _lib.lib()
    code = L.ttt_b200_mlp_forward(*([None] * 19), 1, 1, 1, 1, None)
    assert code == -1 and b"null pointer" in L.ttt_b200_last_error()
