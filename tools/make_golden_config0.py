#!/usr/bin/env python3
"""BASELINE configs[0] through the UNMODIFIED reference generator / loader / voxel path:
a cv2.ellipse base mask (512x512, axes 200x140) -> 64 copies in Section_1, simple_generator half-ellipsoid end caps
(20 + 20 slices) -> ImageLoader -> VoxelProcessor.create_voxel_data / calculate_slice_depths -> VolumeCalculator.
Freezes the mask stack (bit-packed) and the reference's outputs into tests/golden/config0_stack.npz.

    python tools/make_golden_config0.py
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    import cv2
    import scipy.ndimage
    # the generator imports matplotlib only for plotting helpers that are not used here
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import simple_generator
    import image_loader
    import voxel_processor as vp
    import volume_calculator as vc
    vp.SCIPY_AVAILABLE = True
    vp.ndimage = scipy.ndimage
    sink = io.StringIO()
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(sink):
        base = np.zeros((512, 512), dtype=np.uint8)
        cv2.ellipse(base, (256, 256), (200, 140), 0, 0, 360, 255, -1)
        s1 = os.path.join(d, "Section_1")
        os.makedirs(s1)
        for k in range(1, 65):
            cv2.imwrite(os.path.join(s1, "Mask_Patient_%d.png" % k), base)
        first = os.path.join(s1, "Mask_Patient_1.png")
        simple_generator.generate_slices_from_mask(first, 20, os.path.join(d, "Section_0"), 1, False)
        simple_generator.generate_slices_from_mask(os.path.join(s1, "Mask_Patient_64.png"), 20, os.path.join(d, "Section_2"), 64, True)
        loader = image_loader.ImageLoader()
        ok = loader.load_mask_images(d, 200)
        masks = getattr(loader, "mask_images", None)
        if masks is None or len(masks) == 0:
            raise SystemExit("ImageLoader API differs: %s" % [m for m in dir(loader) if not m.startswith("_")])
        counts = (loader.side_0_count, loader.side_1_count, loader.side_2_count) if hasattr(loader, "side_0_count") else (20, 64, 20)
        P = vp.VoxelProcessor()
        vox = P.create_voxel_data(masks, True, *counts)
        depths = P.calculate_slice_depths(6.0)
        mm_x, mm_y = 143.1 / 512, 95.03 / 512
        V = vc.VolumeCalculator()
        vol = V.calculate_voxel_volume_variable_depth(vox, mm_x, mm_y, depths)
        bb = V.calculate_bounding_box_variable_depth(vox, mm_x, mm_y, depths)
    stack = np.stack(masks)
    print("slices", stack.shape, "counts", counts, "active", int(vox.sum()), "volume", repr(vol), "bbox z", bb["z"])
    np.savez_compressed(os.path.join(OUT, "config0_stack.npz"), masks_bits=np.packbits(stack), shape=np.array(stack.shape),
                        sides=np.array(counts), voxel_bits=np.packbits(vox), slice_depths=depths, volume=np.float64(vol),
                        active=np.int64(vox.sum()), bbox=np.array([bb["x"], bb["y"], bb["z"]], dtype=np.float64))
    print("written", os.path.getsize(os.path.join(OUT, "config0_stack.npz")), "bytes")


if __name__ == "__main__":
    main()
