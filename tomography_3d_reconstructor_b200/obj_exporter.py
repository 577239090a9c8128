#!/usr/bin/env python3
"""Drop-in `obj_exporter` module: the reference's OBJExporter (obj_exporter.py:11-41) with the text formatted on the
device.  Same class / method names, parameters, prints and return values; the file is byte-identical
to the one the reference's per-vertex / per-face Python loop writes."""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._lib import check


def obj_body(mesh: "engine.DeviceMesh") -> np.ndarray:
    """uint8 array: vertex lines, one blank line, face lines (everything after the header) for a device mesh."""
    L = engine._L()
    p, st = engine._p, engine._stream
    verts, faces = mesh.verts.contiguous(), mesh.faces.contiguous()
    V, F = int(verts.shape[0]), int(faces.shape[0])
    is64 = 1 if faces.dtype == torch.int64 else 0
    dev = verts.device
    ws = torch.empty(int(L.t3d_obj_workspace_bytes(V, F)) // 8 + 1, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    check(L.t3d_obj_measure(p(verts), V, p(faces), F, is64, p(total), p(ws), st()), "t3d_obj_measure")
    n = int(total.cpu().item()) + 1
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    check(L.t3d_obj_emit(p(verts), V, p(faces), F, is64, p(total), p(ws), p(out), st()), "t3d_obj_emit")
    return engine.download(out)


class OBJExporter:
    """Handles exporting 3D models to OBJ file format (B200)."""

    def __init__(self):
        pass

    def export_to_obj(self, vertices: np.ndarray, faces: np.ndarray, filename: str = "tomography_model.obj") -> bool:
        """Export 3D model to OBJ format (obj_exporter.py:17-41)."""
        try:
            mesh = engine.mesh_from_host(vertices, faces)      # the device mesh it came from, or an upload
            body = obj_body(mesh) if len(vertices) + len(faces) > 0 else np.frombuffer(b"\n", dtype=np.uint8)
            with open(filename, 'wb') as f:
                f.write(b"# Tomography reconstruction model\n")
                f.write(f"# {len(vertices)} vertices, {len(faces)} faces\n\n".encode())
                f.write(body.tobytes() if body.nbytes < (1 << 20) else memoryview(body))

            print(f"Model exported: {filename}")
            return True

        except engine.T3DUnavailable:
            raise
        except Exception as e:
            print(f"Export failed: {e}")
            return False
