import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def model():
    from rbe550_final_project_b200 import panda_model as pm
    return pm.model_arrays()


@pytest.fixture(scope="session")
def c64(model):
    from oracle.c_oracle import COracle
    return COracle(model, "f64")


@pytest.fixture(scope="session")
def c32(model):
    from oracle.c_oracle import COracle
    return COracle(model, "f32")


@pytest.fixture(scope="session")
def pv():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rbe550_final_project_b200.validity import PandaValidity
    h = PandaValidity(0)
    yield h
    h.close()


def random_configs(n, seed, fingers="open"):
    from rbe550_final_project_b200 import panda_model as pm
    rng = np.random.default_rng(seed)
    q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9))
    if fingers == "open":
        q[:, 7:] = 0.04
    return q.astype(np.float32)
