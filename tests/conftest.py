import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import __graft_entry__ as graft  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    return graft.load_oracle()


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def nat(pkg):
    return pkg._native


def load_model_npz(name):
    import json

    z = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    t = {k: z[k] for k in z.files if k not in ("metadata_json", "scalars_json", "codec")}
    t.update(json.loads(str(z["scalars_json"])))
    t["metadata"] = json.loads(str(z["metadata_json"]))
    return t


@pytest.fixture(scope="session")
def golden():
    class G:
        mammography = np.load(os.path.join(GOLDEN, "mammography.npz"))
        shuttle = np.load(os.path.join(GOLDEN, "shuttle.npz"))
        scores = np.load(os.path.join(GOLDEN, "mammography_scores.npz"))

        @staticmethod
        def model(name):
            return load_model_npz(name)

    return G


def synth_mixture(n, d, seed):
    """BASELINE.json's synthetic Gaussian mixture: 49% N(0,I), 49% N(3/sqrt(d), I), 2% N(0, 16 I)."""
    rng = np.random.default_rng(seed)
    z = rng.random(n)
    X = rng.standard_normal((n, d), dtype=np.float32)
    X[(z >= 0.49) & (z < 0.98)] += np.float32(3.0 / np.sqrt(d))
    X[z >= 0.98] *= np.float32(4.0)
    return X
