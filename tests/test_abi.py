"""CPU: the C-ABI library loads and exports every symbol include/ttt_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ttt_b200.h")).read()
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


def test_argument_errors_do_not_need_a_gpu():
    from ttt_video_dit_b200 import _lib
    L = _lib.lib()
    code = L.ttt_b200_mlp_forward(*([None] * 19), 1, 1, 1, 1, None)
    assert code == -1 and b"null pointer" in L.ttt_b200_last_error()
