"""Drop-in shim: `import obj_exporter` resolves here when this directory precedes the reference's on sys.path
(tomography_3d_reconstruction.py:14-17 imports the class by this module name)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from tomography_3d_reconstructor_b200.obj_exporter import OBJExporter  # noqa: E402,F401
