#!/usr/bin/env python3
"""Drop-in `glb_exporter` module: the reference's GLBExporter (glb_exporter.py:20-91).  `create_layer_colors` runs on the
device; `export_to_glb` still goes through trimesh, exactly like the reference, when it is installed."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import engine
from ._lib import check


class GLBExporter:
    """Handles exporting 3D models to GLB file format (B200)."""

    def __init__(self):
        pass

    def export_to_glb(self, vertices: np.ndarray, faces: np.ndarray, filename: str = "tomography_model.glb",
                      vertex_colors: Optional[np.ndarray] = None) -> bool:
        """glb_exporter.py:26-50: trimesh does the writing; without it the reference prints and returns False."""
        try:
            import trimesh
        except ImportError:
            print("Trimesh not available, install required")
            return False
        try:
            mesh = trimesh.Trimesh(vertices=vertices, faces=faces, vertex_colors=vertex_colors)
            mesh.fix_normals()
            mesh.export(filename, file_type='glb')
            print(f"Model exported: {filename}")
            return True
        except Exception as e:
            print(f"Export failed: {e}")
            return False

    def create_layer_colors(self, vertices: np.ndarray, slice_depths: np.ndarray, first_section1_slice: int,
                            last_section1_slice: int, highlight_thickness_mm: float = 1.0) -> np.ndarray:
        """Vertex colours (N,4) uint8 RGBA: grey, red within `highlight_thickness_mm` above the first Section_1 slice,
        blue above the last one (glb_exporter.py:52-91)."""
        n = len(vertices)
        if n == 0:
            return np.zeros((0, 4), dtype=np.uint8)
        cumulative_depths = np.cumsum(np.concatenate([[0], slice_depths]))
        has_a = first_section1_slice < len(cumulative_depths) - 1
        has_b = last_section1_slice < len(cumulative_depths) - 1
        a0 = float(cumulative_depths[first_section1_slice]) if has_a else 0.0
        b0 = float(cumulative_depths[last_section1_slice]) if has_b else 0.0
        a1 = float(cumulative_depths[first_section1_slice] + highlight_thickness_mm) if has_a else 0.0
        b1 = float(cumulative_depths[last_section1_slice] + highlight_thickness_mm) if has_b else 0.0
        mesh = engine.meshes.lookup(vertices)
        if mesh is not None and int(mesh.verts.shape[0]) == n:
            verts = mesh.verts.contiguous()
        else:
            verts = torch.from_numpy(np.ascontiguousarray(vertices, dtype=np.float32)).to(engine._require_cuda())
        out = torch.empty((n, 4), dtype=torch.uint8, device=verts.device)
        check(engine._L().t3d_layer_colors(engine._p(verts), n, 1 if has_a else 0, a0, a1, 1 if has_b else 0, b0, b1,
                                           engine._p(out), engine._stream()), "t3d_layer_colors")
        return engine.download(out)
