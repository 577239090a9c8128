import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu_ref
    cpu_ref.build()
    return cpu_ref


@pytest.fixture(scope="session")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tomography_3d_reconstructor_b200 import engine
    return engine


def random_blobs(rng, shape, density=0.5, smooth=1.5):
    """Random smooth-ish binary volume (thresholded filtered noise)."""
    from scipy import ndimage
    n = rng.standard_normal(shape)
    n = ndimage.gaussian_filter(n, smooth)
    return n > np.quantile(n, 1.0 - density)
