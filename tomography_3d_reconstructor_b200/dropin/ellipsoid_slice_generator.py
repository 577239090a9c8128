"""Drop-in shim: `import ellipsoid_slice_generator` resolves here when this directory precedes the reference's on sys.path
(simple_generator.py:4 imports the class by this module name)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from tomography_3d_reconstructor_b200.ellipsoid_slice_generator import EllipsoidSliceGenerator  # noqa: E402,F401
