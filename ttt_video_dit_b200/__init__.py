"""Importable alias of the ``ttt-video-dit_b200/`` package directory (a hyphen cannot appear in a module name).

``import ttt_video_dit_b200`` executes ``ttt-video-dit_b200/__init__.py`` with this module's ``__path__`` pointing
at that directory, so ``ttt_video_dit_b200.mlp_tk`` etc. resolve to the files kept there.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ttt-video-dit_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
