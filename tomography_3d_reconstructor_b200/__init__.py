"""tomography_3d_reconstructor_b200 -- B200 (sm_100a) implementation of the reconstruction hot path of
victorramirez952/tomography_3d_reconstructor: mask stack -> voxel grid -> smoothing -> marching cubes -> volumes.

Public surface = the reference's three hot-path classes with unchanged signatures:
    VoxelProcessor   (voxel_processor.py:27)   SurfaceExtractor (surface_extractor.py:28)
    VolumeCalculator (volume_calculator.py:10)
`tomography_3d_reconstructor_b200/dropin/` holds same-named shim modules: put that directory first on sys.path
and the unmodified tomography_3d_reconstruction.py, visualizer, obj_exporter and glb_exporter run on top.
"""
from .voxel_processor import VoxelProcessor
from .surface_extractor import SurfaceExtractor
from .volume_calculator import VolumeCalculator

__all__ = ["VoxelProcessor", "SurfaceExtractor", "VolumeCalculator"]
__version__ = "0.1.0"
