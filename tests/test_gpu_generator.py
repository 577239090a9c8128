"""End-cap generator on the device against OpenCV itself: cv2.warpAffine is installed here and on the GPU
box, so the kernel's fixed-point arithmetic is checked bit for bit against the real thing (ellipsoid_slice_generator.py:61-77)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _base(H, W, cx, cy, a, b):
    img = np.zeros((H, W), np.uint8)
    cv2.ellipse(img, (cx, cy), (a, b), 0, 0, 360, 255, -1)
    return img


@pytest.mark.parametrize("shape,centre", [((200, 260), (130.3, 99.6)), ((512, 512), (256.0, 256.0)), ((97, 131), (70.25, 40.5))])
def test_scaled_slices_equal_cv2_warp_affine(eng, shape, centre):
    from tomography_3d_reconstructor_b200.ellipsoid_slice_generator import scaled_slices
    H, W = shape
    base = _base(H, W, int(centre[0]), int(centre[1]), W // 3, H // 4)
    base[H // 2, :] = 255                                   # something that reaches the image border
    factors = [1.0, 0.9987, 0.75, 0.5012, 0.31, 0.1234, 0.0, 1e-3]
    got = scaled_slices(base, centre, factors).cpu().numpy()
    for k, f in enumerate(factors):
        ref = cv2.warpAffine(base, cv2.getRotationMatrix2D(centre, 0, f), (W, H)) if f > 0 else np.zeros_like(base)
        assert np.array_equal(got[k], ref), f


def test_half_ellipsoid_caps_equal_the_reference_recipe(eng, tmp_path):
    """The class: same ellipse fit (cv2 on the host), same z positions, same file names, same pixels as the reference's loop
    (restated here with cv2.warpAffine: ellipsoid_slice_generator.py:107-143)."""
    from tomography_3d_reconstructor_b200.ellipsoid_slice_generator import EllipsoidSliceGenerator
    base = _base(256, 320, 160, 128, 120, 70)
    path = str(tmp_path / "Mask_Patient_1.png")
    cv2.imwrite(path, base)
    gen = EllipsoidSliceGenerator(path)
    c = min(gen.ellipse_params['semi_major_axis'], gen.ellipse_params['semi_minor_axis'])
    centre = gen.ellipse_params['center']
    for increase, num_start in ((False, 1), (True, 64)):
        n = 9
        out_dir = str(tmp_path / ("cap_%d" % increase))
        os.makedirs(out_dir)
        files = gen.generate_slices_half_ellipsoid(n, out_dir, num_start, increase)
        numbers, stack = gen.half_ellipsoid_stack(n, num_start, increase)
        stack = stack.cpu().numpy()
        zs = np.linspace(0, c, n + 2)
        rng = list(range(num_start, num_start + n + 2)) if increase else list(range(num_start - n - 1, num_start + 1))
        assert numbers == rng[1:-1] and len(files) == n + 2
        kept = sorted(f for f in os.listdir(out_dir))
        assert len(kept) == n
        for i, number in enumerate(rng):
            if i in (0, len(rng) - 1):
                assert not os.path.exists(os.path.join(out_dir, "Mask_Patient_%d.png" % number))
                continue
            z = zs[i if increase else len(rng) - 1 - i]
            f = np.sqrt(1 - (z / c) ** 2)
            ref = cv2.warpAffine(gen.middle_slice, cv2.getRotationMatrix2D(centre, 0, f), (320, 256)) if f > 0 else np.zeros_like(base)
            assert np.array_equal(cv2.imread(os.path.join(out_dir, "Mask_Patient_%d.png" % number), cv2.IMREAD_GRAYSCALE), ref)
            assert np.array_equal(stack[i - 1], ref)
