"""The certified never-collide pruning (tools/certify_never_collide.py) must not change any verdict:
dense random sampling against the UNPRUNED pair list, plus a re-run of the certificate on a few pairs."""
import json
import os
import sys

import numpy as np

from rbe550_final_project_b200 import panda_model as pm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_certificate_is_current():
    assert pm.NEVER_COLLIDE_CERT is not None, "data/never_collide.json missing or stale: rerun tools/certify_never_collide.py"
    cert = pm.NEVER_COLLIDE_CERT
    assert cert["model_fingerprint"] == pm.model_fingerprint()
    assert len(cert["never_ss"]) + pm.N_SS_PAIRS == len(pm.SS_PAIRS_UNPRUNED)
    assert len(cert["never_sb"]) + pm.N_SB_PAIRS == len(pm.SB_PAIRS_UNPRUNED)
    assert cert["margin"] >= 1e-3


def test_pruned_pairs_never_come_close_in_2m_samples():
    from oracle.c_oracle import COracle
    full = pm.model_arrays()
    full["ss_pairs"], full["sb_pairs"] = pm.SS_PAIRS_UNPRUNED.copy(), pm.SB_PAIRS_UNPRUNED.copy()
    dropped = pm.model_arrays()
    keep_ss = {tuple(p) for p in pm.SS_PAIRS.tolist()}
    keep_sb = {tuple(p) for p in pm.SB_PAIRS.tolist()}
    dropped["ss_pairs"] = np.array([p for p in pm.SS_PAIRS_UNPRUNED.tolist() if tuple(p) not in keep_ss], dtype=np.int32)
    dropped["sb_pairs"] = np.array([p for p in pm.SB_PAIRS_UNPRUNED.tolist() if tuple(p) not in keep_sb], dtype=np.int32)
    pruned = pm.model_arrays()
    ora_full, ora_drop, ora_pruned = COracle(full, "f64"), COracle(dropped, "f64"), COracle(pruned, "f64")
    empty = {"obb": np.zeros((0, 16)), "table_z": -10.0}
    rng = np.random.default_rng(2025)
    n = 2_000_000
    q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9))
    # bias a third of the samples towards the joint-limit corners where links fold onto each other
    k = n // 3
    corner = rng.random((k, 9)) < 0.5
    q[:k] = np.where(rng.random((k, 9)) < 0.6, np.where(corner, pm.Q_LOWER, pm.Q_UPPER), q[:k])
    m_drop = ora_drop.state_margin(q, empty)
    assert m_drop.min() >= pm.NEVER_COLLIDE_CERT["margin"] - 1e-9, f"a pruned pair came within {m_drop.min()} m"
    m_full = ora_full.state_margin(q, empty)
    m_pruned = ora_pruned.state_margin(q, empty)
    assert ((m_full >= 0) == (m_pruned >= 0)).all()
    close = m_full < pm.NEVER_COLLIDE_CERT["margin"]
    assert np.array_equal(m_full[close], m_pruned[close])


def test_recertify_a_few_pairs():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import certify_never_collide as cert
    never = pm.NEVER_COLLIDE_CERT["never_ss"]
    for a, b in never[:: max(1, len(never) // 6)][:6]:
        kind, _, _, ok, fmin, used = cert.certify(("ss", a, b))
        assert ok and fmin >= cert.MARGIN
    # and a pair that does collide is refuted
    a, b = pm.SS_PAIRS[0]
    assert cert.certify(("ss", int(a), int(b)))[3] in (False,)
