"""GPU parity of the whole-job pipeline (what bench.py times) against the oracle's reference_pipeline."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(24, 64, 96), (40, 100, 130), (32, 128, 128)])
def test_reconstruct_vs_oracle(eng, oracle, shape):
    from tomography_3d_reconstructor_b200 import pipeline
    Z, H, W = shape
    u8 = oracle.ellipsoid_phantom_u8(Z, H, W)
    u8[0, H // 2 - 2:H // 2 + 2, W // 2 - 3:W // 2 + 3] = 0
    sides = (Z // 8, Z - 2 * (Z // 8), Z // 8)
    res = pipeline.reconstruct(torch.from_numpy(u8).cuda(), 200, sides, 6.0, 143.1, 95.03)
    ref = oracle.reference_pipeline(u8, 200, sides, 6.0, 143.1, 95.03)
    mesh = res["mesh"]
    v, f = mesh.verts.cpu().numpy(), mesh.faces.cpu().numpy()
    assert v.shape == ref["vertices"].shape and f.shape == ref["faces"].shape
    assert np.allclose(v, ref["vertices"], rtol=1e-5, atol=1e-5)
    assert np.array_equal(f, ref["faces"])
    assert res["voxel_volume_mm3"] == ref["voxel_volume"]                 # bit-exact (integer counts, same f64 order)
    assert res["processed_voxel_volume_mm3"] == ref["processed_volume"]
    assert abs(res["mesh_volume_mm3"] - ref["mesh_volume"]) <= 1e-6 * ref["mesh_volume"]
    assert abs(res["surface_area_mm2"] - ref["surface_area"]) <= 1e-6 * ref["surface_area"]
    z, y, x = np.where(ref["voxel_data"])
    assert res["bbox_index"] == (z.min(), z.max(), y.min(), y.max(), x.min(), x.max())
    assert res["active_voxels"] == int(ref["voxel_data"].sum())
    assert mesh.n_ambiguous == ref["n_ambiguous"] == 0
