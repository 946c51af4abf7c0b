import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def butterfly_bgra():
    import numpy as np
    from PIL import Image

    im = np.array(Image.open(os.path.join(GOLDEN, "butterfly.png")).convert("RGBA"))
    return np.ascontiguousarray(im[:, :, [2, 1, 0, 3]])


@pytest.fixture(scope="session")
def butterfly_oracle(butterfly_bgra):
    """Oracle run on the reference's fixture image (config[0] of BASELINE.json)."""
    from oracle_lib import Oracle

    h, w = butterfly_bgra.shape[:2]
    o = Oracle(w, h, collect_stats=True)
    kps, counts = o.detect(butterfly_bgra)
    desc, dcounts = o.describe()
    return {"oracle": o, "keypoints": kps, "counts": counts, "descriptors": desc, "dcounts": dcounts}
