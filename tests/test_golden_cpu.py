"""The committed golden vectors (tests/golden/validity_golden.npz, made by tools/make_golden.py from the numpy
oracle) against the C oracle and the frozen model -- CPU only."""
import os

import numpy as np

from oracle import panda_oracle as po
from rbe550_final_project_b200 import panda_model as pm
from rbe550_final_project_b200 import scenes as sc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "validity_golden.npz"))
SCENES = ["goal1_scattered", "goal4_task1_pentagon", "goal3_tower"]


def test_golden_matches_model():
    assert str(G["model_fingerprint"]) == pm.model_fingerprint(), "model changed: rerun tools/make_golden.py"
    assert int(G["n_ss_pairs"]) == pm.N_SS_PAIRS and int(G["n_sb_pairs"]) == pm.N_SB_PAIRS


def test_c_oracle_reproduces_golden(c64, c32):
    q = G["q"].astype(np.float64)
    R, p = c64.fk(q)
    assert np.abs(R - G["fk_R"]).max() < 1e-12 and np.abs(p - G["fk_p"]).max() < 1e-12
    for name in SCENES:
        s = sc.FIXTURES[name]().as_oracle_scene()
        assert np.abs(c64.state_margin(q, s) - G[f"{name}/margin"]).max() < 1e-12
        assert np.abs(c64.state_margin(q, s, attached=2) - G[f"{name}/margin_attached2"]).max() < 1e-12
        assert np.abs(c64.state_margin(q, s, flags=0) - G[f"{name}/margin_noself"]).max() < 1e-12
        k = 256
        qb = G["qb"].astype(np.float64)
        assert np.abs(c64.edge_margin(q[:k], qb[:k], s, n_steps=64) - G[f"{name}/edge64"]).max() < 1e-12
        assert np.abs(c64.edge_margin(q[:k], qb[:k], s, n_steps=0) - G[f"{name}/edge_res"]).max() < 1e-12
        # the fp32 port (the CPU baseline) gives the same verdicts away from contact
        m32 = c32.state_margin(G["q"], s)
        far = np.abs(G[f"{name}/margin"]) > 1e-4
        assert ((m32 >= 0) == (G[f"{name}/margin"] >= 0))[far].all()


def test_numpy_oracle_reproduces_golden(model):
    s = sc.goal4_task1_pentagon().as_oracle_scene()
    m = po.state_margin(G["q"][:300].astype(np.float64), s, model)
    assert np.array_equal(m, G["goal4_task1_pentagon/margin"][:300])
    assert np.array_equal(po.sweep_configs(64, 512, 20251212, model, fingers_open=False).view(np.uint32),
                          G["sweep_q"].view(np.uint32))


def test_oracle_vs_genesis_goldens(c64):
    """Pins the oracle to the real reference stack -- IF tools/export_genesis_goldens.py has ever been run where
    Genesis is installed.  It cannot be run in the build container, so this is skipped there and the
    parity of the robot model vs Genesis stays "unpinned"."""
    import glob
    import pytest
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "genesis_*.npz")))
    if not files:
        pytest.skip("no Genesis golden vectors committed (Genesis/OMPL are not installable here): parity unpinned")
    for f in files:
        g = np.load(f)
        s = sc.FIXTURES[str(g["scene"])]().as_oracle_scene()
        R, p = c64.fk(g["q"])
        assert np.abs(p - g["link_pos"]).max() < 1e-5  # the FK chain is exact; only the collision hulls are approximated
        m = c64.state_margin(g["q"], s)
        far = np.abs(m) > 5e-3  # primitives vs convexified meshes: agreement is only claimed away from contact
        agree = ((m >= 0) == g["valid"])[far].mean()
        assert agree > 0.99, (f, agree)
