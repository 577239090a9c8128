#!/usr/bin/env python3
"""Drop-in `ellipsoid_slice_generator` module: the reference's EllipsoidSliceGenerator
(ellipsoid_slice_generator.py:8-203) with the per-slice cv2.warpAffine replaced by one device launch for a whole cap.

Loading the base mask, thresholding, contour extraction and the ellipse fit stay with cv2 on the host, as in the reference
(one 2-D image, not a hot path); what moves to the GPU is the generation of the scaled copies, bit-identical to
cv2.warpAffine (csrc/t3d_generator.cu).  `half_ellipsoid_stack` additionally returns a cap as a device stack, so fixtures
for the reconstruction path never take the PNG round trip."""
from __future__ import annotations

import os
from typing import List

import numpy as np
import torch

from . import engine
from ._lib import check


def inverted_affine(center, factor: float) -> np.ndarray:
    """cv2.getRotationMatrix2D(center, 0, factor) inverted exactly as cv::warpAffine inverts its matrix (float64)."""
    cx, cy = float(center[0]), float(center[1])
    alpha, beta = np.cos(0.0) * factor, np.sin(0.0) * factor
    M = np.array([alpha, beta, (1 - alpha) * cx - beta * cy, -beta, alpha, beta * cx + (1 - alpha) * cy], dtype=np.float64)
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def scaled_slices(base_u8, center, factors) -> torch.Tensor:
    """uint8 (n,H,W) on the device: slice k = cv2.warpAffine(base, getRotationMatrix2D(center, 0, factors[k]), (W, H));
    factors[k] <= 0 gives np.zeros_like(base) (ellipsoid_slice_generator.py:63-69)."""
    dev = engine._require_cuda()
    base = base_u8 if isinstance(base_u8, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(base_u8, dtype=np.uint8))
    base = base.to(dev).contiguous()
    H, W = (int(v) for v in base.shape)
    mats = np.zeros((len(factors), 6), dtype=np.float64)
    for k, f in enumerate(factors):
        if f > 0:
            mats[k] = inverted_affine(center, float(f))
    out = torch.empty((len(factors), H, W), dtype=torch.uint8, device=dev)
    if len(factors):
        check(engine._L().t3d_endcap_slices(engine._p(base), H, W, engine._p(torch.from_numpy(mats).to(dev)), len(factors),
                                            engine._p(out), engine._stream()), "t3d_endcap_slices")
    return out


class EllipsoidSliceGenerator:
    def __init__(self, image_path: str):
        """Initialize ellipsoid slice generator with middle slice image."""
        self.image_path = image_path
        self.middle_slice = self._load_and_preprocess_image()
        self.ellipse_params = self._extract_ellipse_parameters()

    def _load_and_preprocess_image(self) -> np.ndarray:
        import cv2
        img = cv2.imread(self.image_path, cv2.IMREAD_GRAYSCALE)
        if img is None:
            raise ValueError(f"Could not load image from {self.image_path}")
        _, binary_img = cv2.threshold(img, 127, 255, cv2.THRESH_BINARY)
        return binary_img

    def _extract_ellipse_parameters(self) -> dict:
        import cv2
        contours, _ = cv2.findContours(self.middle_slice, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        if not contours:
            raise ValueError("No contours found in the image")
        largest_contour = max(contours, key=cv2.contourArea)
        if len(largest_contour) < 5:
            raise ValueError("Could not fit ellipse to the contour")
        center, axes, angle = cv2.fitEllipse(largest_contour)
        return {'center': center, 'semi_major_axis': max(axes) / 2, 'semi_minor_axis': min(axes) / 2, 'angle': angle,
                'area': cv2.contourArea(largest_contour)}

    def _calculate_ellipse_area_at_height(self, z: float, c: float) -> float:
        if abs(z) > c:
            return 0.0
        factor = np.sqrt(1 - (z / c) ** 2)
        return np.pi * self.ellipse_params['semi_major_axis'] * factor * self.ellipse_params['semi_minor_axis'] * factor

    @staticmethod
    def _factor(z: float, c: float) -> float:
        """ellipsoid_slice_generator.py:63-69: 0 stands for 'all-zero slice'."""
        if z < 0 or z > c:
            return 0.0
        factor = np.sqrt(1 - (z / c) ** 2) if c > 0 else 0
        return float(factor) if factor > 0 else 0.0

    def _generate_slice_at_height(self, z: float, c: float) -> np.ndarray:
        """Generate slice at height z (z=0: original mask, z=c: smallest slice)."""
        return scaled_slices(self.middle_slice, self.ellipse_params['center'], [self._factor(z, c)])[0].cpu().numpy()

    def _half_ellipsoid_plan(self, num_slices: int, num_start: int, increase: bool):
        c = min(self.ellipse_params['semi_major_axis'], self.ellipse_params['semi_minor_axis'])
        z_positions = np.linspace(0, c, num_slices + 2)
        if increase:
            num_end = num_start + 1 + num_slices
        else:
            num_end = num_start - num_slices - 1
            num_start, num_end = num_end, num_start
        number_range = list(range(num_start, num_end + 1))
        zs = []
        for i, _number in enumerate(number_range):
            z_index = i if increase else len(number_range) - 1 - i
            zs.append(z_positions[z_index] if z_index < len(z_positions) else c)
        return c, number_range, zs

    def half_ellipsoid_stack(self, num_slices: int, num_start: int = 28, increase: bool = True):
        """The slices generate_slices_half_ellipsoid keeps (the two extreme ones are dropped, :140-142), in file-number
        order, as (numbers, uint8 (num_slices,H,W) device stack): one launch, no files."""
        c, number_range, zs = self._half_ellipsoid_plan(num_slices, num_start, increase)
        stack = scaled_slices(self.middle_slice, self.ellipse_params['center'], [self._factor(z, c) for z in zs])
        return number_range[1:-1], stack[1:-1]

    def generate_slices_half_ellipsoid(self, num_slices: int, output_dir: str = "slices", num_start: int = 28,
                                       increase: bool = True) -> List[str]:
        """Generate half-ellipsoid slices with sequential naming (original mask as base); same files as the reference."""
        import cv2
        c, number_range, zs = self._half_ellipsoid_plan(num_slices, num_start, increase)
        stack = scaled_slices(self.middle_slice, self.ellipse_params['center'], [self._factor(z, c) for z in zs]).cpu().numpy()
        saved_files = []
        for number, img in zip(number_range, stack):
            filepath = os.path.join(output_dir, f"Mask_Patient_{number}.png")
            cv2.imwrite(filepath, img)
            saved_files.append(filepath)
        os.remove(saved_files[0])
        os.remove(saved_files[-1])
        return saved_files

    def generate_slices(self, num_slices: int, output_dir: str = "slices") -> List[str]:
        """Generate n slices sorted by area (smallest to largest)."""
        import cv2
        os.makedirs(output_dir, exist_ok=True)
        c = min(self.ellipse_params['semi_major_axis'], self.ellipse_params['semi_minor_axis'])
        z_positions = np.linspace(-c, c, num_slices)
        stack = scaled_slices(self.middle_slice, self.ellipse_params['center'], [self._factor(z, c) for z in z_positions]).cpu().numpy()
        slice_data = [(i, z, stack[i], np.sum(stack[i] > 0)) for i, z in enumerate(z_positions)]
        slice_data.sort(key=lambda x: x[3])
        saved_files = []
        for mask_number, (_i, _z, slice_img, _area) in enumerate(slice_data, 1):
            filepath = os.path.join(output_dir, f"Mask_{mask_number:03d}.png")
            cv2.imwrite(filepath, slice_img)
            saved_files.append(filepath)
        return saved_files
