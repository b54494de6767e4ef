import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _unpack(bits, shape):
    return np.unpackbits(bits)[: int(shape[0]) * int(shape[1])].reshape(int(shape[0]), int(shape[1])).astype(np.uint8)


@pytest.fixture(scope="session")
def gt55():
    """test_data/gt_shrinked_img55.png as a {0,1} bitmap (metrics.rs:510-646 input)."""
    z = np.load(os.path.join(GOLDEN, "gt_shrinked_img55.npz"))
    return _unpack(z["bits"], z["shape"])


@pytest.fixture(scope="session")
def gt_others():
    z = np.load(os.path.join(GOLDEN, "gt_shrinked_others.npz"))
    return {n: _unpack(z[n], z[n + "_shape"]) for n in ("img224", "img494", "img545")}


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def image_files():
    z = np.load(os.path.join(GOLDEN, "image_files.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def preprocessed():
    z = np.load(os.path.join(GOLDEN, "preprocessed.npz"))
    return {k: z[k] for k in z.files}


# The reference's golden expectations (metrics.rs:524-569 and :596-641), shared by the
# oracle tests (CPU) and the CUDA parity tests (GPU).
GOLDEN_POLYS_1X = [
    [(562, 75), (559, 108), (532, 108), (427, 105), (435, 68)],
    [(547, 178), (515, 255), (404, 212), (287, 226), (263, 160), (407, 125)],
    [(448, 245), (450, 301), (427, 301), (345, 292), (332, 233)],
    [(400, 322), (534, 303), (550, 361), (401, 385), (263, 319), (278, 271)],
]
GOLDEN_POLYS_2X = [
    [(281, 38), (280, 54), (266, 54), (214, 53), (218, 34)],
    [(274, 89), (258, 128), (202, 106), (144, 113), (132, 80), (204, 63)],
    [(224, 123), (225, 151), (214, 151), (173, 146), (166, 117)],
    [(200, 161), (267, 152), (275, 181), (201, 193), (132, 160), (139, 136)],
]
GOLDEN_SCORES = [0.9819034852546917, 0.9938022931515339, 0.9911894273127754, 0.9923459624952162]

KAT_MAP_5x5 = np.array(
    [[0, 0, 0, 1, 0], [0, 0, 1, 1, 0], [0, 1, 1, 1, 0], [0, 1, 1, 0, 0], [0, 1, 1, 0, 0]], np.float32)
KAT_BOX_SCORES = [  # metrics.rs:426-484
    ([(0, 0), (4, 0), (4, 4), (0, 4)], 10.0 / 25.0),
    ([(1, 0), (4, 0), (4, 3), (1, 3)], 8.0 / 16.0),
    ([(2, 0), (4, 1), (2, 4), (1, 3)], 9.0 / 12.0),
]
KAT_BINARIZE_IN = np.array(  # metrics.rs:486-508
    [[0.01, 0.2, 0.57, 0.58, 0.18], [0.39, 0.01, 0.61, 1.0, 0.42], [0.4, 0.94, 0.835, 0.793, 0.32],
     [0.57, 0.77, 0.62, 0.51, 0.29], [0.11, 0.69, 0.59, 0.21, 0.35]], np.float64)
KAT_BINARIZE_OUT = np.array(
    [[0, 0, 0, 1, 0], [0, 0, 1, 1, 0], [0, 1, 1, 1, 0], [0, 1, 1, 0, 0], [0, 1, 1, 0, 0]], np.uint8)
KAT_MINRECT_IN = [(141, 24), (61, 16), (57, 53), (137, 61)]  # metrics.rs:406-424
KAT_MINRECT_BOX = [(60, 15), (142, 23), (138, 62), (57, 54)]
KAT_MINRECT_SSIDE = 39.11521443121589
