"""The CPU oracle against everything that can pin it: the reference's own code through committed golden fixtures
(tools/make_golden.py), scipy's documented known answers, analytic volumes, mesh invariants and the survey's
identities."""
import os

import numpy as np
import pytest
from scipy import ndimage

from conftest import random_blobs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unpack(bits, shape):
    return np.unpackbits(bits)[:int(np.prod(shape))].reshape(shape).astype(bool)


@pytest.fixture(scope="module")
def vox_gold():
    return np.load(os.path.join(GOLD, "voxel_path.npz"))


@pytest.fixture(scope="module")
def surf_gold():
    return np.load(os.path.join(GOLD, "surface_path.npz"))


def test_voxel_path_matches_reference_fixture(oracle, vox_gold):
    g = vox_gold
    u8, shape = g["masks_u8"], tuple(g["shape"])
    masks = [u8[z] >= 200 for z in range(shape[0])]
    vox = oracle.create_voxel_data(masks, True)
    assert np.array_equal(vox, unpack(g["voxel_data"], shape))
    assert np.array_equal(oracle.create_voxel_data(masks, False), unpack(g["raw"], shape))
    assert np.array_equal(oracle.close_volume_ends_stencil(np.stack(masks)), vox)       # V3: loop == stencil
    depths = oracle.calculate_slice_depths(6.0, *g["sides"])
    assert np.array_equal(depths, g["slice_depths"])
    mm_x, mm_y = g["mm"]
    assert oracle.calculate_voxel_volume_variable_depth(vox, mm_x, mm_y, depths) == g["vol_var"]
    assert oracle.calculate_voxel_volume(vox, mm_x, mm_y, 0.2) == g["vol_uni"]
    bb = oracle.calculate_bounding_box_variable_depth(vox, mm_x, mm_y, depths)
    assert np.array_equal(np.array([*bb["x"], *bb["y"], *bb["z"], *bb["dimensions"]]), g["bb_var"])
    bb = oracle.calculate_bounding_box(vox, mm_x, mm_y, 0.2)
    assert np.array_equal(np.array([*bb["x"], *bb["y"], *bb["z"], *bb["dimensions"]]), g["bb_uni"])
    assert np.array_equal(oracle.generate_point_cloud(vox, mm_x, mm_y, depths, 1), g["pc1"])
    assert np.array_equal(oracle.generate_point_cloud(vox, mm_x, mm_y, depths, 3), g["pc3"])


def test_surface_postprocessing_matches_reference_fixture(oracle, vox_gold, surf_gold):
    g, s = vox_gold, surf_gold
    shape = tuple(g["shape"])
    depths = g["slice_depths"]
    mm_x, mm_y = (float(v) for v in g["mm"])   # python floats, as the orchestrator passes them (NEP 50: a numpy
    #                                            float64 scalar would make `vertices[:, 1] *= mm` multiply in float64)
    sm = oracle.smooth_voxel_data(unpack(g["voxel_data"], shape), 3, True)
    assert np.array_equal(sm, unpack(s["smoothed"], shape))
    for pad, key in ((True, "pad"), (False, "nopad")):
        v, f, namb = oracle.extract_manifold_surface(sm, depths, mm_y, mm_x, True, True, pad, return_diag=True)
        assert namb == s[key + "_n_ambiguous"] == 0
        assert np.array_equal(v, s[key + "_verts"]) and v.dtype == np.float32     # reference loop + np.unique
        assert np.array_equal(f, s[key + "_faces"]) and f.dtype == np.int64
        lit = float(s[key + "_mesh_volume_literal"])                               # float32 accumulation (NumPy 2)
        assert abs(oracle.calculate_mesh_volume_f64(v, f) - lit) <= 2e-4 * lit
        assert abs(float(oracle.calculate_surface_area(v, f)) - float(s[key + "_area_literal"])) <= 1e-6 * lit
        vv = np.zeros((len(s["zmap_in"]), 3), dtype=np.float32)
        vv[:, 0] = s["zmap_in"]
        oracle.apply_variable_slice_depths(vv, depths, pad)                        # V8: closed form == loop
        assert np.array_equal(vv[:, 0], s["zmap_" + key])


def test_config0_slice_depths_known_answer(oracle):
    g = np.load(os.path.join(GOLD, "config0_known.npz"))
    d = oracle.calculate_slice_depths(6.0, 20, 64, 20)
    assert np.array_equal(d, g["slice_depths"])
    assert d[0] == 0.009375 and d[20] == 0.09375 and abs(d.sum() - 6.375) < 1e-12
    assert len(oracle.calculate_slice_depths(6.0, 0, 0, 0)) == 0
    assert np.array_equal(oracle.calculate_slice_depths(6.0, 2, 0, 2), np.full(4, 1.5))


def test_fill_holes_scipy_docstring_known_answer(oracle):
    a = np.zeros((5, 5), dtype=int)
    a[1:4, 1:4] = 1
    a[2, 2] = 0
    expect = np.zeros((5, 5), dtype=bool)
    expect[1:4, 1:4] = True
    vol = np.stack([a.astype(bool)] * 3)
    out = oracle.close_volume_ends(vol)
    assert np.array_equal(out[0], expect) and np.array_equal(out[2], expect)
    assert np.array_equal(out[1], vol[1] | (expect & expect))


def test_morphology_identities(oracle):
    rng = np.random.default_rng(0)
    cross = ndimage.generate_binary_structure(3, 1)
    for _ in range(20):
        shape = tuple(rng.integers(2, 11, 3))
        x = rng.random(shape) < rng.uniform(0.2, 0.8)
        c1 = oracle.binary_closing6(x)
        assert np.array_equal(oracle.binary_closing6(c1), c1)                     # V4: closing idempotent
        o1 = oracle.binary_opening6(x)
        assert np.array_equal(oracle.binary_opening6(o1), o1)
        # 7-point stencils with constant padding == the scipy calls
        p = np.pad(x, 1, constant_values=True)
        er = p[1:-1, 1:-1, 1:-1] & p[:-2, 1:-1, 1:-1] & p[2:, 1:-1, 1:-1] & p[1:-1, :-2, 1:-1] & p[1:-1, 2:, 1:-1] \
            & p[1:-1, 1:-1, :-2] & p[1:-1, 1:-1, 2:]
        assert np.array_equal(er, oracle.binary_erosion6(x))
        assert np.array_equal(ndimage.binary_dilation(x, structure=cross), oracle.binary_dilation6(x))
        assert np.array_equal(oracle.smooth_voxel_data(x, 3, True), oracle.binary_closing6(oracle.binary_opening6(x)))


def test_gaussian_order_and_sign_rule(oracle):
    """V5 (summation order) and the face-neighbour sign rule the CUDA path relies on."""
    from scipy.ndimage._filters import _gaussian_kernel1d
    k = _gaussian_kernel1d(0.5, 0, 2)
    w0, w1, w2 = k[2], k[1], k[0]
    rng = np.random.default_rng(1)
    x = rng.random(4096)
    ref = ndimage.correlate1d(x, k[::-1], mode="reflect")
    xp = np.pad(x, 2, mode="symmetric")
    mine = xp[2:-2] * w0 + (xp[:-4] + xp[4:]) * w2 + (xp[1:-3] + xp[3:-1]) * w1
    assert np.array_equal(mine, ref)
    assert w0 ** 3 + w0 * w0 * w1 > 0.55 and 1.0 - w0 ** 3 - w0 * w0 * w1 < 0.45
    for seed in range(5):
        vol = rng.random((9, 11, 13)) < rng.uniform(0.1, 0.9)
        f = oracle.scalar_field(vol, True, True)
        p = np.pad(vol, 1)
        q = np.pad(p, 1, mode="symmetric")
        nb = [q[:-2, 1:-1, 1:-1], q[2:, 1:-1, 1:-1], q[1:-1, :-2, 1:-1], q[1:-1, 2:, 1:-1], q[1:-1, 1:-1, :-2], q[1:-1, 1:-1, 2:]]
        any1 = np.logical_or.reduce(nb)
        all1 = np.logical_and.reduce(nb)
        sign = f > 0.5
        assert sign[p & any1].all()           # set voxel with a set face neighbour  -> inside
        assert not sign[~p & ~all1].any()     # clear voxel with a clear face neighbour -> outside


def test_marching_cubes_invariants(oracle):
    rng = np.random.default_rng(3)
    for shape, smooth in (((20, 30, 40), 2.0), ((16, 16, 16), 1.5)):
        vol = oracle.smooth_voxel_data(random_blobs(rng, shape, 0.4, smooth), 3, True)
        if not vol.any():
            continue
        v, f, namb = oracle.extract_manifold_surface(vol, np.ones(shape[0]), 1.0, 1.0, True, True, True, return_diag=True)
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        key = e[:, 0] * (len(v) + 1) + e[:, 1]
        rev = e[:, 1] * (len(v) + 1) + e[:, 0]
        if namb == 0:
            assert len(np.unique(key)) == len(key)                 # oriented: each directed edge once
            assert np.array_equal(np.sort(key), np.sort(rev))      # closed: every edge has its twin
        d = np.diff(v.astype(np.float64), axis=0)                  # np.unique contract: strictly increasing rows
        assert ((d[:, 0] > 0) | ((d[:, 0] == 0) & ((d[:, 1] > 0) | ((d[:, 1] == 0) & (d[:, 2] > 0))))).all()


def test_analytic_volumes(oracle):
    Z, H, W = 48, 96, 128
    vol = oracle.ellipsoid_phantom_u8(Z, H, W) >= 200
    analytic = 4 / 3 * np.pi * (0.42 * Z) * (0.33 * H) * (0.45 * W)
    assert abs(vol.sum() - analytic) / analytic < 0.01
    v, f = oracle.extract_manifold_surface(oracle.smooth_voxel_data(vol, 3, True), np.ones(Z), 1.0, 1.0, True, True, False)
    assert abs(oracle.calculate_mesh_volume_f64(v, f) - analytic) / analytic < 0.01
    box = np.zeros((12, 14, 16), bool)
    box[3:9, 4:10, 5:12] = True
    v, f = oracle.extract_manifold_surface(box, np.ones(12), 1.0, 1.0, True, True, True)
    assert 150 < oracle.calculate_mesh_volume_f64(v, f) < 6 * 6 * 7 + 1


def test_failure_modes_return_none(oracle):
    d = np.ones(4)
    assert oracle.extract_manifold_surface(np.zeros((4, 6, 6), bool), d, 1.0, 1.0) is None
    assert oracle.extract_manifold_surface(np.ones((4, 6, 6), bool), d, 1.0, 1.0, True, True, False) is None
    with pytest.raises(ValueError, match="Load masks first"):
        oracle.create_voxel_data([])


def test_edt_oracle(oracle):
    vol = np.zeros((7, 9, 11), bool)
    vol[2:5, 3:7, 4:9] = True
    sdf = oracle.signed_distance(vol)
    assert sdf[3, 4, 6] > 0 and sdf[0, 0, 0] < 0
    assert sdf[2, 3, 4] == 1.0 and sdf[1, 3, 4] == -1.0
    assert np.array_equal(oracle.squared_edt_index(vol)[3, 4:6, 6], [4, 4])


def test_config0_real_generator_stack_known_answer(oracle):
    """BASELINE configs[0] from the reference's OWN generator + loader (tools/make_golden_config0.py):
    the oracle's voxel path must reproduce the reference's grid, slice depths and voxel volume bit for bit."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config0_stack.npz"))
    shape = tuple(int(v) for v in g["shape"])
    n = int(np.prod(shape))
    masks = np.unpackbits(g["masks_bits"])[:n].reshape(shape).astype(bool)
    vox = oracle.create_voxel_data([m for m in masks], True)
    assert np.array_equal(np.packbits(vox), g["voxel_bits"])
    assert int(vox.sum()) == int(g["active"]) == 8030338
    depths = oracle.calculate_slice_depths(6.0, *(int(s) for s in g["sides"]))
    assert np.array_equal(depths, g["slice_depths"])
    vol = oracle.calculate_voxel_volume_variable_depth(vox, 143.1 / 512, 95.03 / 512, depths)
    assert vol == float(g["volume"]) == 28658.498565015263


def _closed_oriented(v, f):
    """Every directed edge exactly once and its twin present: closed, oriented, no edge with more than two triangles."""
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).astype(np.int64)
    key = e[:, 0] * (len(v) + 1) + e[:, 1]
    rev = e[:, 1] * (len(v) + 1) + e[:, 0]
    return len(np.unique(key)) == len(key) and np.array_equal(np.sort(key), np.sort(rev))


def test_face_test_is_the_asymptotic_decider(oracle):
    """Lewiner's face test on case 3 (corners 0 and 2 positive on the z=0 face): positives joined iff p1*p2 - n1*n2 >= 0."""
    tris, J, tube, row = oracle.resolve_cube([1.0, -1.0, 0.1, -1.0, -1, -1, -1, -1])       # 0.1 - 1 < 0: separated
    assert (J, tube) == (0, 0) and tris.tolist() == [[0, 8, 3], [1, 2, 10]]                # = the classic row (Lewiner 3.1)
    tris, J, tube, row = oracle.resolve_cube([3.0, -1.0, 2.0, -1.0, -1, -1, -1, -1])       # 6 - 1 > 0: joined (3.2)
    assert (J, tube) == (1, 0) and len(tris) == 4
    tris, J, tube, row = oracle.resolve_cube([1.0, -1.0, 1.0, -1.0, -1, -1, -1, -1])       # exact tie -> joined
    assert J == 1
    # the complement (corners 0, 2 negative): the same face, positives 1 and 3
    tris, J, tube, row = oracle.resolve_cube([-1.0, 0.1, -1.0, 1.0, 1, 1, 1, 1])
    assert J == 0 and len(tris) == 4                                                       # negatives joined: the 4-triangle tiling
    tris, J, tube, row = oracle.resolve_cube([-1.0, 3.0, -1.0, 2.0, 1, 1, 1, 1])
    assert J == 1 and len(tris) == 2


def test_interior_test_decides_the_case4_tunnel(oracle):
    """Case 4 (corners 0 and 6 positive): strong corners -> the trilinear surface is one tunnel (4.2, 6 triangles);
    weak corners -> two separate caps (4.1, 2 triangles)."""
    strong = [10.0, -1, -1, -1, -1, -1, 10.0, -1]
    tris, J, tube, row = oracle.resolve_cube(strong)
    assert tube == 1 and len(tris) == 6
    # the trilinear interpolant is indeed positive at the cube centre for the strong cube (tunnel), negative for the weak one
    assert sum(strong) / 8 > 0
    weak = [1.0, -5, -5, -5, -5, -5, 1.0, -5]
    tris, J, tube, row = oracle.resolve_cube(weak)
    assert tube == 0 and tris.tolist() == [[0, 8, 3], [5, 10, 6]] and sum(weak) / 8 < 0
    # complement: the negative corners 0 and 6 joined through the interior
    tris, J, tube, row = oracle.resolve_cube([-v for v in strong])
    assert tube == 1 and len(tris) == 6


def test_resolved_meshes_are_watertight_on_noise(oracle):
    """Neighbouring cubes decide a shared ambiguous face from the same four values: no cracks, whatever the field."""
    for seed in range(3):
        rng = np.random.default_rng(seed)
        vol = np.zeros((14, 15, 16), np.float32)
        vol[1:-1, 1:-1, 1:-1] = rng.random((12, 13, 14)).astype(np.float32)
        v, f, namb = oracle.marching_cubes(vol, 0.5)
        amb, changed, tunnels, interior = oracle.last_mc33_stats
        assert namb == amb > 100 and changed > 50 and tunnels > 5 and interior >= tunnels
        assert _closed_oriented(v, f)
        # and through the whole extraction (Gaussian field of a noisy occupancy)
        occ = rng.random((10, 20, 30)) < 0.35
        ev, ef, n2 = oracle.extract_manifold_surface(occ, np.ones(10), 1.0, 1.0, True, True, True, return_diag=True)
        assert n2 > 0 and _closed_oriented(ev, ef)
