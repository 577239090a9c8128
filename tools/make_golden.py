#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference modules under /root/reference.

The reference ships no tests or fixtures, so these vectors are produced here by importing
its own code (sys.path -> /root/reference) and running every hot-path function that works without scikit-image:
create_voxel_data (with the module flags patched so the scipy hole-fill branch runs),
calculate_slice_depths, generate_point_cloud, _add_volume_padding, _apply_variable_slice_depths,
_ensure_manifold_mesh, calculate_mesh_volume, calculate_surface_area and all of VolumeCalculator.
Marching cubes itself (skimage) cannot run; the mesh fed to the post-processing functions is the oracle's.

Run in the build container only (the GPU box has no /root/reference):  python tools/make_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def reference_modules():
    sys.path.insert(0, REF)
    import scipy.ndimage
    import voxel_processor as vp
    import surface_extractor as se
    import volume_calculator as vc
    vp.SCIPY_AVAILABLE = True        # skimage is absent, so the import fallback cleared this although scipy is present
    vp.ndimage = scipy.ndimage
    sys.path.pop(0)
    return vp, se, vc


def main():
    sys.path.insert(0, ROOT)
    from oracle import cpu_ref
    vp_m, se_m, vc_m = reference_modules()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(2024)
    sink = io.StringIO()

    # ---- fixture 1: mini config-0 (ellipsoid phantom, 5+20+5 slices of 48x64) through the reference voxel path
    Z, H, W = 30, 48, 64
    u8 = cpu_ref.ellipsoid_phantom_u8(Z, H, W)
    u8[0, 20:26, 28:36] = 255
    u8[0, 22:24, 30:34] = 0            # a hole in the first slice
    u8[-1, 18:30, 20:44] = 255
    u8[-1, 21:27, 25:30] = 0
    masks = [u8[z] >= 200 for z in range(Z)]
    sides = (5, 20, 5)
    with contextlib.redirect_stdout(sink):
        P = vp_m.VoxelProcessor()
        vox = P.create_voxel_data(masks, True, *sides)
        depths = P.calculate_slice_depths(6.0)
        raw = vp_m.VoxelProcessor().create_voxel_data(masks, False)
        mm_x, mm_y = 143.1 / W, 95.03 / H
        V = vc_m.VolumeCalculator()
        vol_var = V.calculate_voxel_volume_variable_depth(vox, mm_x, mm_y, depths)
        vol_uni = V.calculate_voxel_volume(vox, mm_x, mm_y, 0.2)
        bb_var = V.calculate_bounding_box_variable_depth(vox, mm_x, mm_y, depths)
        bb_uni = V.calculate_bounding_box(vox, mm_x, mm_y, 0.2)
        pc1 = P.generate_point_cloud(vox, mm_x, mm_y, depths, 1)
        pc3 = P.generate_point_cloud(vox, mm_x, mm_y, depths, 3)
        props = V.analyze_object_properties(vox, 123.0, 120.0, 50.0, mm_x, mm_y, depths, 143.1, 95.03, 6.0)
    np.savez_compressed(
        os.path.join(OUT, "voxel_path.npz"), masks_u8=u8, sides=np.array(sides), voxel_data=np.packbits(vox),
        raw=np.packbits(raw), shape=np.array(vox.shape), slice_depths=depths, vol_var=vol_var, vol_uni=vol_uni,
        bb_var=np.array([*bb_var["x"], *bb_var["y"], *bb_var["z"], *bb_var["dimensions"]], dtype=np.float64),
        bb_uni=np.array([*bb_uni["x"], *bb_uni["y"], *bb_uni["z"], *bb_uni["dimensions"]], dtype=np.float64),
        pc1=pc1, pc3=pc3, density=props["density"], mm=np.array([mm_x, mm_y]))

    # ---- fixture 2: surface post-processing on a mesh from the oracle's marching cubes
    sm = cpu_ref.smooth_voxel_data(vox, 3, True)
    S = se_m.SurfaceExtractor()
    out = {}
    for pad in (True, False):
        vol = S._add_volume_padding(sm) if pad else sm
        f32 = cpu_ref.scalar_field(sm, True, pad)
        assert f32.shape == vol.shape
        verts, faces, namb = cpu_ref.marching_cubes(f32, 0.5)
        v_in = verts.copy()
        v = verts.copy()
        v -= 1
        S._apply_variable_slice_depths(v, depths, pad)      # reference python loop
        v[:, 1] *= mm_y
        v[:, 2] *= mm_x
        uv, uf = S._ensure_manifold_mesh(v, faces)
        key = "pad" if pad else "nopad"
        out.update({key + "_mc_verts": v_in, key + "_mc_faces": faces, key + "_verts": uv, key + "_faces": uf,
                    key + "_mesh_volume_literal": np.float64(S.calculate_mesh_volume(uv, uf)),
                    key + "_area_literal": np.float64(S.calculate_surface_area(uv, uf)),
                    key + "_n_ambiguous": namb})
    out["smoothed"] = np.packbits(sm)
    # random z values through the reference loop, both padding modes
    zr = (rng.random(2000) * (Z + 4) - 2).astype(np.float32)
    for pad in (True, False):
        vv = np.zeros((len(zr), 3), dtype=np.float32)
        vv[:, 0] = zr
        S._apply_variable_slice_depths(vv, depths, pad)
        out["zmap_%s" % ("pad" if pad else "nopad")] = vv[:, 0].copy()
    out["zmap_in"] = zr
    np.savez_compressed(os.path.join(OUT, "surface_path.npz"), **out)

    # ---- fixture 3: the survey's config-0 known answer (V12) needs cv2 + the reference generator; record constants
    np.savez_compressed(os.path.join(OUT, "config0_known.npz"),
                        slice_depths=vp_m_depths(vp_m, sink, 20, 64, 20), expected_sum=np.float64(6.375))
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-24s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


def vp_m_depths(vp_m, sink, s0, s1, s2):
    P = vp_m.VoxelProcessor()
    P.side_0_count, P.side_1_count, P.side_2_count = s0, s1, s2
    with contextlib.redirect_stdout(sink):
        return P.calculate_slice_depths(6.0)


if __name__ == "__main__":
    main()
