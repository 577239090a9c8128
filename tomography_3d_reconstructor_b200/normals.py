"""Area-weighted vertex normals on the device (additive: the reference drops skimage's normals)."""
from __future__ import annotations

import torch

from . import engine
from ._lib import check


def vertex_normals(mesh: engine.DeviceMesh) -> torch.Tensor:
    v, f = mesh.verts.contiguous(), mesh.faces.contiguous()
    out = torch.empty_like(v)
    check(engine._L().t3d_vertex_normals(engine._p(v), int(v.shape[0]), engine._p(f), int(f.shape[0]),
                                         1 if f.dtype == torch.int64 else 0, engine._p(out), engine._stream()),
          "t3d_vertex_normals")
    return out
