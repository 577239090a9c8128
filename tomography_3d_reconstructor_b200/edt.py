"""Exact Euclidean distance transform / signed distance on the device (additive stage).

Oracle: scipy.ndimage.distance_transform_edt; sdf = edt(occ) - edt(~occ), positive inside, float32."""
from __future__ import annotations

import ctypes

import torch

from . import engine
from ._lib import check


def _sampling(sampling):
    s = [float(v) for v in sampling]
    if len(s) != 3 or min(s) <= 0:
        raise ValueError("sampling must be three positive numbers (z, y, x)")
    return (ctypes.c_double * 3)(*s)


def distance(dv: engine.DeviceVolume, sampling=(1.0, 1.0, 1.0), invert: bool = False) -> torch.Tensor:
    """float32 (Z,H,W): distance from every set (unset if invert) voxel to the nearest voxel of the other kind."""
    L = engine._L()
    Z, H, W = dv.shape
    dev = dv.bits.device
    out = torch.empty((Z, H, W), dtype=torch.float32, device=dev)
    ws = torch.empty(int(L.t3d_edt_workspace_bytes(Z, H, W)) // 8 + 1, dtype=torch.int64, device=dev)
    check(L.t3d_edt(engine._p(dv.bits), Z, H, W, 1 if invert else 0, _sampling(sampling), 1.0, 0, engine._p(out),
                    engine._p(ws), engine._stream()), "t3d_edt")
    return out


def signed_distance(dv: engine.DeviceVolume, sampling=(1.0, 1.0, 1.0)) -> torch.Tensor:
    """float32 (Z,H,W): edt(occ) - edt(~occ)."""
    L = engine._L()
    Z, H, W = dv.shape
    dev = dv.bits.device
    out = torch.empty((Z, H, W), dtype=torch.float32, device=dev)
    ws = torch.empty(int(L.t3d_edt_workspace_bytes(Z, H, W)) // 8 + 1, dtype=torch.int64, device=dev)
    s3 = _sampling(sampling)
    check(L.t3d_edt(engine._p(dv.bits), Z, H, W, 0, s3, 1.0, 0, engine._p(out), engine._p(ws), engine._stream()), "t3d_edt")
    check(L.t3d_edt(engine._p(dv.bits), Z, H, W, 1, s3, -1.0, 1, engine._p(out), engine._p(ws), engine._stream()), "t3d_edt")
    return out
