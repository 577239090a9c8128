"""GPU parity of the device-side mesh consumers against golden vectors produced by the unmodified
reference exporters (tools/make_golden_export.py) and against Python's own '%.6f' formatting."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "export_path.npz")


def _export(verts, faces, tmp_path, name="m.obj"):
    from tomography_3d_reconstructor_b200.obj_exporter import OBJExporter
    path = os.path.join(str(tmp_path), name)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        assert OBJExporter().export_to_obj(verts, faces, path) is True
    assert sink.getvalue() == "Model exported: %s\n" % path
    return open(path, "rb").read()


def test_obj_file_is_byte_identical_to_the_reference(eng, tmp_path):
    g = np.load(GOLD)
    assert _export(g["verts"], g["faces"], tmp_path) == bytes(g["obj_bytes"])
    assert _export(g["verts"], g["faces"].astype(np.int32), tmp_path, "i32.obj") == bytes(g["obj_bytes"])
    assert _export(g["verts"][:0], g["faces"][:0], tmp_path, "empty.obj") == bytes(g["obj_empty_bytes"])


def test_obj_number_formatting_matches_python(eng, tmp_path):
    """'%.6f' of float32 values over many magnitudes, halfway cases included, line by line against Python."""
    rng = np.random.default_rng(3)
    n = 30000
    mag = 10.0 ** rng.uniform(-9, 9, size=(n, 3))
    v = (rng.standard_normal((n, 3)) * mag).astype(np.float32)
    k = np.arange(3000)
    v[:3000, 0] = ((2 * k + 1) * 0.5e-6).astype(np.float32)          # near rounding ties of the 6th decimal
    v[:3000, 1] = -(k * 0.125 + 0.0000005).astype(np.float32)
    faces = rng.integers(0, n, size=(1000, 3)).astype(np.int64)
    faces[5] = [n - 1, 0, 99999 % n]
    got = _export(v, faces, tmp_path).decode().split("\n")
    assert got[0] == "# Tomography reconstruction model" and got[1] == "# %d vertices, %d faces" % (n, 1000) and got[2] == ""
    for i in range(n):
        assert got[3 + i] == "v %.6f %.6f %.6f" % (float(v[i, 0]), float(v[i, 1]), float(v[i, 2])), i
    assert got[3 + n] == ""
    for i in range(1000):
        assert got[4 + n + i] == "f %d %d %d" % tuple(int(a) + 1 for a in faces[i])
    assert got[4 + n + 1000] == "" and len(got) == 5 + n + 1000


def test_layer_colors_match_the_reference(eng):
    from tomography_3d_reconstructor_b200.glb_exporter import GLBExporter
    g = np.load(GOLD)
    G = GLBExporter()
    for i, (a, b, t) in enumerate(g["color_cases"]):
        c = G.create_layer_colors(g["color_verts"], g["color_depths"], int(a), int(b), float(t))
        assert c.dtype == np.uint8 and c.shape == (len(g["color_verts"]), 4)
        assert np.array_equal(c, g["colors_%d" % i]), i
    assert G.create_layer_colors(g["color_verts"][:0], g["color_depths"], 20, 83).shape == (0, 4)


def test_exporters_through_dropin_modules_use_the_device_mesh(eng, oracle, tmp_path):
    """`import obj_exporter` / `import glb_exporter` resolve to the shims; a mesh that came from
    extract_manifold_surface is exported from its device copy (no upload) and equals the Python-formatted text."""
    import tomography_3d_reconstructor_b200 as pkg
    shim = os.path.join(os.path.dirname(pkg.__file__), "dropin")
    sys.path.insert(0, shim)
    try:
        for m in ("obj_exporter", "glb_exporter", "surface_extractor", "voxel_processor"):
            sys.modules.pop(m, None)
        import glb_exporter
        import obj_exporter
        import surface_extractor
        import voxel_processor
        u8 = oracle.ellipsoid_phantom_u8(24, 40, 56)
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            vp = voxel_processor.VoxelProcessor()
            vox = vp.create_voxel_data([u8[z] >= 200 for z in range(24)], True, 4, 16, 4)
            depths = vp.calculate_slice_depths(6.0)
            v, f = surface_extractor.SurfaceExtractor().extract_manifold_surface(vp.smooth_voxel_data(vox), depths, 2.0, 2.5)
            path = os.path.join(str(tmp_path), "d.obj")
            assert obj_exporter.OBJExporter().export_to_obj(v, f, path)
            colors = glb_exporter.GLBExporter().create_layer_colors(v, depths, 4, 19, 1.0)
        assert not v.flags.writeable and not f.flags.writeable
        lines = open(path).read().split("\n")
        assert lines[3] == "v %.6f %.6f %.6f" % tuple(float(x) for x in v[0])
        assert lines[3 + len(v) + 1 + len(f) - 1] == "f %d %d %d" % tuple(int(a) + 1 for a in f[-1])
        cum = np.cumsum(np.concatenate([[0], depths]))
        red = (v[:, 0] >= cum[4]) & (v[:, 0] <= cum[4] + 1.0)
        blue = (v[:, 0] >= cum[19]) & (v[:, 0] <= cum[19] + 1.0)
        want = np.full((len(v), 4), [200, 200, 200, 255], dtype=np.uint8)
        want[red] = [255, 0, 0, 255]
        want[blue] = [0, 0, 255, 255]
        assert np.array_equal(colors, want)
    finally:
        sys.path.remove(shim)
        for m in ("obj_exporter", "glb_exporter", "surface_extractor", "voxel_processor"):
            sys.modules.pop(m, None)


def test_glb_writer_reparses_to_the_mesh(eng, oracle, tmp_path):
    """glb_exporter.py:26-50 without trimesh: the binary glTF file written from the device mesh re-parses to exactly the
    vertices, faces (wound outwards, as trimesh's fix_normals() leaves them) and colours it was given."""
    import json
    from tomography_3d_reconstructor_b200 import SurfaceExtractor
    from tomography_3d_reconstructor_b200.glb_exporter import GLBExporter, parse_glb
    vol = oracle.smooth_voxel_data(oracle.ellipsoid_phantom_u8(24, 56, 72) >= 200, 3, True)
    depths = oracle.calculate_slice_depths(6.0, 3, 18, 3)
    se, ex = SurfaceExtractor(), GLBExporter()
    v, f = se.extract_manifold_surface(vol, depths, 0.3, 0.25, True, True, True)
    colors = ex.create_layer_colors(v, depths, 3, 20, 1.0)
    path = str(tmp_path / "model.glb")
    assert ex.export_to_glb(v, f, path, colors) is True
    data = open(path, "rb").read()
    pos, idx, col, doc = parse_glb(data)
    assert np.array_equal(pos.view(np.uint32), v.view(np.uint32)) and np.array_equal(col, colors)
    # winding: outward (positive signed volume in the file's own coordinates), every face the same triangle as given
    p64 = pos.astype(np.float64)
    a, b, c = p64[idx[:, 0]], p64[idx[:, 1]], p64[idx[:, 2]]
    assert np.einsum("ij,ij->i", a, np.cross(b, c)).sum() > 0
    same = np.array_equal(idx.astype(np.int64), f)
    flipped = np.array_equal(idx.astype(np.int64), f[:, [0, 2, 1]])
    assert same or flipped
    acc = doc["accessors"][doc["meshes"][0]["primitives"][0]["attributes"]["POSITION"]]
    assert np.array_equal(np.float32(acc["min"]), v.min(axis=0)) and np.array_equal(np.float32(acc["max"]), v.max(axis=0))
    assert doc["asset"]["version"] == "2.0" and len(data) % 4 == 0
    # host arrays that did not come from this package (fresh copies, int32 faces) and no colours
    assert ex.export_to_glb(v.copy(), f.astype(np.int32), path, None) is True
    pos2, idx2, col2, _ = parse_glb(open(path, "rb").read())
    assert col2 is None and np.array_equal(pos2, v) and np.array_equal(idx2, idx)
    # failures keep the reference's convention: message + False
    assert ex.export_to_glb(v[:0], f[:0], path, None) is False
