#!/usr/bin/env python3
"""Golden vectors for the device-side mesh consumers, produced by the UNMODIFIED reference modules
obj_exporter.OBJExporter.export_to_obj and glb_exporter.GLBExporter.create_layer_colors (both import and run here).

    python tools/make_golden_export.py      # writes tests/golden/export_path.npz
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def awkward_vertices(rng, n):
    v = (rng.standard_normal((n, 3)) * np.array([3.0, 40.0, 60.0])).astype(np.float32)
    special = np.array([0.0, -0.0, 1e-7, -1e-7, 5e-7, -5e-7, 4.9999997e-7, 1.5e-6, 2.5e-6, 0.9999995, 0.99999946, 9.9999995,
                        123456.789, -98765.4321, 1e-30, 3.4e12, 16777216.0, 0.1, 0.2, 0.3, 6.375, 1e6 + 0.5, 2.0000005],
                       dtype=np.float32)
    k = min(len(special), n)
    v[:k, 0] = special[:k]
    v[:k, 1] = special[:k][::-1]
    v[:k, 2] = -special[:k]
    return v


def main():
    sys.path.insert(0, REF)
    import obj_exporter as ref_obj
    import glb_exporter as ref_glb
    sys.path.pop(0)
    rng = np.random.default_rng(99)
    sink = io.StringIO()
    out = {}
    V, F = 500, 900
    verts = awkward_vertices(rng, V)
    faces = rng.integers(0, V, size=(F, 3)).astype(np.int64)
    faces[0] = [0, V - 1, 7]
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(sink):
        path = os.path.join(d, "m.obj")
        assert ref_obj.OBJExporter().export_to_obj(verts, faces, path)
        out["obj_bytes"] = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
        assert ref_obj.OBJExporter().export_to_obj(verts[:0], faces[:0], path)
        out["obj_empty_bytes"] = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
    out["verts"], out["faces"] = verts, faces
    # layer colours: z column in mm against cumulative slice depths
    depths = np.array([0.009375] * 20 + [0.09375] * 64 + [0.009375] * 20)
    zc = rng.uniform(-0.5, 7.0, size=4000).astype(np.float32)
    cum = np.cumsum(np.concatenate([[0], depths]))
    zc[:len(cum)] = cum.astype(np.float32)                         # exactly on the thresholds (float32 vs float64 compare)
    zc[len(cum):2 * len(cum)] = (cum + 1.0).astype(np.float32)
    cv = np.zeros((len(zc), 3), dtype=np.float32)
    cv[:, 0] = zc
    G = ref_glb.GLBExporter()
    cases = [(20, 83, 1.0), (0, 103, 0.25), (20, 200, 1.0), (104, 104, 1.0), (50, 52, 0.05)]
    for i, (a, b, t) in enumerate(cases):
        out["colors_%d" % i] = G.create_layer_colors(cv, depths, a, b, t)
    out["color_cases"] = np.array(cases, dtype=np.float64)
    out["color_verts"], out["color_depths"] = cv, depths
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "export_path.npz"), **out)
    print("written", os.path.join(OUT, "export_path.npz"), os.path.getsize(os.path.join(OUT, "export_path.npz")), "bytes")


if __name__ == "__main__":
    main()
