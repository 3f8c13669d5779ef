"""GPU parity tests: CUDA kernels (through the C-ABI) vs the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): FK poses within 1e-5 m / 1e-5 rad; collision verdicts identical for
every configuration more than 1e-4 m from contact (|oracle margin| > 1e-4); bit packing exact.
"""
import numpy as np
import pytest
import torch

from conftest import random_configs
from oracle import panda_oracle as po
from rbe550_final_project_b200 import panda_model as pm
from rbe550_final_project_b200 import scenes as sc
from rbe550_final_project_b200.validity import PandaValidity, unpack_bits

pytestmark = pytest.mark.gpu

BAND = 1e-4
SCENES = ["goal1_scattered", "goal4_task1_pentagon", "goal3_tower"]


def _dev(q):
    return torch.as_tensor(q, device="cuda")


def _assert_verdicts(gpu_valid, margin, what):
    ora_valid = margin >= 0
    far = np.abs(margin) > BAND
    bad = np.nonzero((gpu_valid != ora_valid) & far)[0]
    assert bad.size == 0, f"{what}: {bad.size} verdict mismatches outside the 1e-4 band, first {bad[:5]}, margins {margin[bad[:5]]}"
    return int((~far).sum()), int(((gpu_valid != ora_valid) & ~far).sum())


def test_fk_parity(pv, c64):
    pv.set_scene(sc.goal1_scattered())
    q = random_configs(50000, 11, fingers="random")
    q[0] = 0
    q[1] = pm.Q_SAFE_HOME
    q[2] = pm.Q_SCENE_INIT
    out = pv.fk(_dev(q)).cpu().numpy()
    R, p = c64.fk(q.astype(np.float64))
    assert np.abs(out[:, :, 0:3] - p).max() < 2e-6  # bar: 1e-5 m
    Rg = out[:, :, 3:12].reshape(-1, 11, 3, 3)
    # rotation angle between the two frames: bar 1e-5 rad
    # (sin of the relative angle from the skew part: arccos of the trace is ill-conditioned near 0)
    rel = np.einsum("nlij,nlik->nljk", R, Rg.astype(np.float64))
    skew = 0.5 * (rel - np.swapaxes(rel, 2, 3))
    ang = np.arcsin(np.clip(np.sqrt((skew ** 2).sum((2, 3)) / 2), 0, 1))
    assert np.abs(Rg - R).max() < 2e-6
    assert ang.max() < 1e-5
    # analytic known answers (base lift 0.01 included)
    assert np.allclose(out[0, 7, 0:3], [0.088, 0, 1.043], atol=2e-6)
    assert np.allclose(out[0, 8, 0:3], [0.088, 0, 0.936], atol=2e-6)
    assert np.allclose(out[1, 8, 0:3], [0.30702, 0, 0.60027], atol=1e-5)


def test_fk_verdict_path(pv, c64):
    """The verdict-bit kernels take their joint sines / cosines from the hardware approximations; the poses they work
    with must still meet the 1e-5 m / 1e-5 rad bar (and stay far inside the 1e-4 m verdict band)."""
    pv.set_scene(sc.goal1_scattered())
    q = random_configs(50000, 12, fingers="random")
    q[0] = pm.Q_UPPER
    q[1] = pm.Q_LOWER
    q[2, 5] = np.float32(3.7525)  # the one joint that exceeds pi inside its limits (reduced by 2 pi before MUFU)
    out = pv.fk(_dev(q), verdict_path=True).cpu().numpy()
    R, p = c64.fk(q.astype(np.float64))
    err_p = np.abs(out[:, :, 0:3] - p).max()
    Rg = out[:, :, 3:12].reshape(-1, 11, 3, 3)
    rel = np.einsum("nlij,nlik->nljk", R, Rg.astype(np.float64))
    skew = 0.5 * (rel - np.swapaxes(rel, 2, 3))
    ang = np.arcsin(np.clip(np.sqrt((skew ** 2).sum((2, 3)) / 2), 0, 1)).max()
    print(f"verdict-path FK: max position error {err_p:.2e} m, max angle error {ang:.2e} rad")
    assert err_p < 2e-6 and ang < 3e-6  # measured 5.5e-7 m, 9.3e-7 rad; bars: 1e-5 m, 1e-5 rad


@pytest.mark.parametrize("scene_name", SCENES)
@pytest.mark.parametrize("attached", [-1, 3])
def test_state_verdicts(pv, c64, scene_name, attached):
    scene = sc.FIXTURES[scene_name]()
    pv.set_scene(scene)
    pv.set_flags(True, False)
    pv.set_attached(attached)
    n = 200_003  # ragged tail on purpose
    q = random_configs(n, 5 + attached, fingers="random" if attached >= 0 else "open")
    bits = pv.check_states(_dev(q))
    torch.cuda.synchronize()
    gpu_valid = unpack_bits(bits, n)
    margin = c64.state_margin(q.astype(np.float64), scene.as_oracle_scene(), attached=attached)
    in_band, flips = _assert_verdicts(gpu_valid, margin, scene_name)
    assert in_band < n * 2e-3
    # tail bits of the last word are zero
    w = bits.cpu().numpy().view(np.uint32)
    assert (w[-1] >> np.uint32(n & 31)) == 0
    pv.set_attached(-1)


def test_state_margins_and_culprits(pv, c64):
    scene = sc.goal4_task1_pentagon()
    pv.set_scene(scene)
    pv.set_flags(True, False)
    q = random_configs(100_000, 21, fingers="random")
    m, cu = pv.state_margins(_dev(q), want_culprit=True)
    m = m.cpu().numpy()
    ref = c64.state_margin(q.astype(np.float64), scene.as_oracle_scene())
    assert np.abs(m - ref).max() < 1e-5
    cu = cu.cpu().numpy()
    assert ((cu != 0) == (m < 0)).all()
    # bits and margins agree with each other away from contact
    bits = unpack_bits(pv.check_states(_dev(q)), len(q))
    far = np.abs(ref) > BAND
    assert (bits[far] == (m[far] >= 0)).all()


def test_flags_self_and_limits(pv, c64):
    """Self-collision can be switched off; joint limits cannot (they are part of the model's validity domain: the
    pruned self-pair lists are certified inside them only), so PV_FLAG_LIMITS changes nothing."""
    from oracle.c_oracle import FLAG_LIMITS, FLAG_SELF
    scene = sc.goal1_scattered()
    pv.set_scene(scene)
    q = random_configs(60_000, 31, fingers="random")
    q[::7, 3] += 3.2  # push joint 4 out of its limits on every 7th config
    q[::11, 8] = 0.0405
    outside = ~((q >= pm.Q_LOWER.astype(np.float32)) & (q <= pm.Q_UPPER.astype(np.float32))).all(axis=1)
    assert outside.sum() > 10_000
    for self_on, lim_on in [(False, False), (True, True), (False, True), (True, False)]:
        pv.set_flags(self_on, lim_on)
        gpu = unpack_bits(pv.check_states(_dev(q)), len(q))
        assert not gpu[outside].any()
        fl = (FLAG_SELF if self_on else 0) | (FLAG_LIMITS if lim_on else 0)
        margin = c64.state_margin(q.astype(np.float64), scene.as_oracle_scene(), flags=fl)
        _assert_verdicts(gpu, margin, f"flags {self_on} {lim_on}")
    pv.set_flags(True, False)


def test_out_of_limit_states_never_pass_through_a_pruned_pair(pv, model):
    """the never-collide certificate that prunes 293 of the 425 self pairs holds inside the joint
    limits only.  Configurations with arm joints beyond the limits that self-collide through a PRUNED pair (found with
    the unpruned lists of the numpy oracle, limits rule bypassed) must not be reported valid -- by any entry point."""
    from oracle import panda_oracle as po
    rng = np.random.default_rng(77)
    n = 40_000
    q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9))
    j = rng.integers(0, 7, n)
    over = rng.uniform(0.02, 1.2, n) * rng.choice([-1.0, 1.0], n)
    q[np.arange(n), j] = np.where(over > 0, pm.Q_UPPER[j] + over, pm.Q_LOWER[j] + over)
    q = q.astype(np.float32)
    empty = sc.SceneSnapshot(obb=np.zeros((0, 16), np.float32))
    pv.set_scene(empty)
    pv.set_flags(True, False)
    # geometry-only margins with the pruned and the unpruned pair lists (limits rule bypassed by widening them)
    wide = dict(model, q_lower=np.full(9, -1e3), q_upper=np.full(9, 1e3))
    unpruned = dict(wide, ss_pairs=pm.SS_PAIRS_UNPRUNED, sb_pairs=pm.SB_PAIRS_UNPRUNED)
    m_pruned = po.state_margin(q.astype(np.float64), empty.as_oracle_scene(), wide)
    m_unpruned = po.state_margin(q.astype(np.float64), empty.as_oracle_scene(), unpruned)
    leak = (m_pruned > 1e-4) & (m_unpruned < -1e-4)  # collides ONLY through pruned pairs
    assert leak.sum() > 20, "the sample must contain states that the pruned lists alone would wave through"
    gpu = unpack_bits(pv.check_states(_dev(q)), n)
    assert not gpu.any()
    assert not unpack_bits(pv.check_states_host(q[leak]), int(leak.sum())).any()
    m = pv.state_margins(_dev(q[leak])).cpu().numpy()
    assert (m < 0).all()
    qa = np.clip(q[leak], pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
    assert not unpack_bits(pv.check_edges_host(qa, q[leak], n_steps=0), int(leak.sum())).any()


def test_state_kernel_variants_are_bit_identical(pv):
    """0 = brute force over every kept pair, 1 = + per-lane bounding-ball culling (default)."""
    for name in SCENES:
        pv.set_scene(sc.FIXTURES[name]())
        for att, fingers in ((-1, "open"), (2, "random")):
            pv.set_attached(att)
            q = _dev(random_configs(300_007, 77 + att, fingers=fingers))
            res = []
            for mode in (0, 1, 2):  # brute force, per-lane culling, sorted by elbow angle + culling (the default)
                pv.set_culling(mode)
                res.append(pv.check_states(q).cpu().numpy())
            pv.set_culling(2)
            assert (res[0] == res[1]).all() and (res[0] == res[2]).all(), (name, att)
            # the sorted kernel with far fewer configurations than one block's share, and a ragged tail
            for m_ in (1, 33, 511, 512, 513, 1025, 70_001, 148 * 512 + 1, 149 * 512, 296 * 512 - 31):
                pv.set_culling(1)
                a_ = pv.check_states(q[:m_]).cpu().numpy()
                pv.set_culling(2)
                assert (pv.check_states(q[:m_]).cpu().numpy() == a_).all(), (name, att, m_)
    pv.set_attached(-1)


@pytest.mark.parametrize("n_steps", [64, 0])
def test_edge_verdicts(pv, c64, n_steps):
    scene = sc.goal4_task1_pentagon()
    pv.set_scene(scene)
    pv.set_flags(True, False)
    n = 20_001
    qa = random_configs(n, 41)
    rng = np.random.default_rng(42)
    qb = np.clip(qa + rng.normal(0, 0.3, qa.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
    qb[:, 7:] = 0.04
    bits = pv.check_edges(_dev(qa), _dev(qb), n_steps=n_steps)
    gpu = unpack_bits(bits, n)
    ref = c64.edge_margin(qa.astype(np.float64), qb.astype(np.float64), scene.as_oracle_scene(), n_steps=n_steps)
    _assert_verdicts(gpu, ref, f"edges n_steps={n_steps}")
    m = pv.edge_margins(_dev(qa), _dev(qb), n_steps=n_steps).cpu().numpy()
    assert np.abs(m - ref).max() < 2e-5
    # property: an edge verdict is the AND of its state verdicts (64-step form)
    if n_steps == 64:
        k = 2000
        t = (np.arange(1, 65, dtype=np.float32) / np.float32(64))[None, :, None]
        states = qa[:k, None, :] + t * (qb[:k, None, :] - qa[:k, None, :])
        states[:, -1, :] = qb[:k]
        sm = pv.state_margins(_dev(states.reshape(-1, 9))).cpu().numpy().reshape(k, 64).min(1)
        far = np.abs(sm) > BAND
        assert (gpu[:k][far] == (sm[far] >= 0)).all()


def test_edge_long_edges_resolution_mode(pv, c64):
    """uniform-pair edges are long (nd up to ~60): exercises the multi-round coarse-to-fine path"""
    scene = sc.goal3_tower()
    pv.set_scene(scene)
    n = 5000
    qa, qb = random_configs(n, 51), random_configs(n, 52)
    gpu = unpack_bits(pv.check_edges(_dev(qa), _dev(qb), n_steps=0), n)
    ref = c64.edge_margin(qa.astype(np.float64), qb.astype(np.float64), scene.as_oracle_scene(), n_steps=0)
    _assert_verdicts(gpu, ref, "long edges")


def test_motion_certificates_do_not_change_verdicts(pv, c64):
    """pv_edge_kernel finishes a motion of more than 32 states after its coarse round when that round's culling tests
    clear by more than the robot can travel between tested states.  The verdict words must be those of
    the exhaustive validator (pv_set_culling(1) switches the certificates off), whatever the scene, the attached box,
    the motion length or the step rule -- including motions that start outside the joint limits."""
    rng = np.random.default_rng(97)
    n = 200_003
    qa = random_configs(n, 95)
    cases = []
    for sigma in (0.05, 0.3, 1.0):
        qb = np.clip(qa + rng.normal(0, sigma, qa.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
        qb[:, 7:] = rng.uniform(0, 0.04, (n, 2)).astype(np.float32)
        cases.append((qa, qb))
    qo = qa.copy()
    qo[::3, 1] = np.float32(pm.Q_LOWER[1]) - rng.uniform(0, 0.05, qo[::3, 1].shape).astype(np.float32)  # start outside
    cases.append((qo, cases[1][1]))
    for name in SCENES:
        pv.set_scene(sc.FIXTURES[name]())
        for att in (-1, 2):
            pv.set_attached(att)
            for a, b in cases:
                for n_steps in (64, 40, 200, 0):
                    A, B = _dev(a), _dev(b)
                    pv.set_culling(1)
                    ref = pv.check_edges(A, B, n_steps=n_steps).cpu().numpy()
                    pv.set_culling(2)
                    got = pv.check_edges(A, B, n_steps=n_steps).cpu().numpy()
                    assert (ref == got).all(), (name, att, n_steps)
    pv.set_attached(-1)
    # against the fp64 oracle on long motions cut finely (every certificate skips 3 of 4 states)
    scene = sc.goal4_task1_pentagon()
    pv.set_scene(scene)
    k = 6000
    a, b = cases[1][0][:k], cases[1][1][:k]
    gpu = unpack_bits(pv.check_edges(_dev(a), _dev(b), n_steps=128), k)
    ref = c64.edge_margin(a.astype(np.float64), b.astype(np.float64), scene.as_oracle_scene(), n_steps=128)
    _assert_verdicts(gpu, ref, "certified motions, 128 steps")


def test_second_tier_certificates_run_and_change_nothing(pv, monkeypatch):
    """pv_edge_cert2_kernel (the self-collision section with slack at the 16 coarse states) sits between the
    scene-only and the self-collision-only list validators: a handle created with PV_EDGE_CERT2=0 leaves it out.  Same
    verdict words either way, one launch more with it."""
    monkeypatch.setenv("PV_EDGE_CERT2", "0")
    pv0 = PandaValidity(0)
    monkeypatch.delenv("PV_EDGE_CERT2")
    rng = np.random.default_rng(5)
    n = 150_017
    qa = random_configs(n, 96)
    for name, sigma in (("goal4_task1_pentagon", 0.3), ("goal3_tower", 0.15)):
        qb = np.clip(qa + rng.normal(0, sigma, qa.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
        qb[:, 7:] = 0.04
        snap = sc.FIXTURES[name]()
        pv.set_scene(snap)
        pv0.set_scene(snap)
        A, B = _dev(qa), _dev(qb)
        for n_steps in (64, 100, 33):
            l0, l1 = pv0.launch_count, pv.launch_count
            w0 = pv0.check_edges(A, B, n_steps=n_steps).cpu().numpy()
            w1 = pv.check_edges(A, B, n_steps=n_steps).cpu().numpy()
            assert (w0 == w1).all(), (name, n_steps)
            assert pv0.launch_count - l0 == 3 and pv.launch_count - l1 == 4
    del pv0


def test_host_entry_points_match_device(pv):
    scene = sc.goal1_scattered()
    pv.set_scene(scene)
    n = 600_011  # > 2 pipeline chunks, ragged
    q = random_configs(n, 61)
    dev = pv.check_states(_dev(q)).cpu().numpy().view(np.uint32)
    host = pv.check_states_host(q)
    assert (dev == host).all()
    pinned = torch.from_numpy(q).pin_memory()
    host2 = pv.check_states_host(pinned.numpy())
    assert (dev == host2).all()
    assert pv.is_state_valid(pm.Q_SAFE_HOME) is True
    ne = 40_000
    qb = np.clip(q[:ne] + 0.2, pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
    d = pv.check_edges(_dev(q[:ne]), _dev(qb), n_steps=16).cpu().numpy().view(np.uint32)
    hh = pv.check_edges_host(q[:ne], qb, n_steps=16)
    assert (d == hh).all()


def test_host_entry_points_around_the_small_batch_threshold(pv):
    """Host batches up to 2 048 rows take the host-mapped path (one launch, no copies, unsorted kernel / one motion per
    warp), larger ones the chunked copy pipeline: same words on both sides of the threshold, ragged sizes included."""
    pv.set_scene(sc.goal3_tower())
    q = random_configs(5000, 91)
    q[7] = np.nan  # a non-finite joint value is an invalid state on every path
    qb = np.clip(q + np.random.default_rng(9).normal(0, 0.3, q.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
    qb[:, 7:] = 0.04
    dev_s = unpack_bits(pv.check_states(_dev(q)).cpu().numpy().view(np.uint32), len(q))
    dev_e = unpack_bits(pv.check_edges(_dev(q), _dev(qb), n_steps=0).cpu().numpy().view(np.uint32), len(q))
    assert not dev_s[7] and not dev_e[7]
    for n in (1, 31, 32, 33, 1000, 2047, 2048, 2049, 4097):
        assert np.array_equal(unpack_bits(pv.check_states_host(q[:n]), n), dev_s[:n]), n
        assert np.array_equal(unpack_bits(pv.check_edges_host(q[:n], qb[:n], n_steps=0), n), dev_e[:n]), n
        assert np.array_equal(unpack_bits(pv.check_states_host_arm(np.ascontiguousarray(q[:n, :7])), n), dev_s[:n]), n


def test_arm_rows_host_entry_point(pv, c64):
    """pv_check_states_host_arm: rows of 7 arm joint values + one gripper opening = the 9-column call on the same
    configurations, bit for bit (ragged sizes, several pipeline chunks, equal and unequal fingers), and the oracle."""
    scene = sc.goal1_scattered()
    pv.set_scene(scene)
    for n, fingers in ((1, (0.04, 0.04)), (33, (0.0, 0.04)), (600_011, (0.04, 0.04)), (300_000, (0.013, 0.027))):
        q = random_configs(n, 62 + n % 7)
        q[:, 7], q[:, 8] = np.float32(fingers[0]), np.float32(fingers[1])
        ref = pv.check_states_host(q)
        got = pv.check_states_host_arm(np.ascontiguousarray(q[:, :7]), fingers)
        assert np.array_equal(ref, got), (n, fingers)
    m = c64.state_margin(q[:20000].astype(np.float64), scene.as_oracle_scene())
    clear = np.abs(m) > 1e-4
    assert np.array_equal(unpack_bits(got, n)[:20000][clear], (m >= 0)[clear])
    with pytest.raises(Exception):
        pv.check_states_host_arm(np.zeros((4, 6), np.float32))


def test_overlapping_launches_keep_stream_order(pv):
    """pv_set_launch_overlap (programmatic dependent launch, default on): back-to-back launches give the words of
    serialised launches; a launch whose input planes come out of a kernel queued just before it sees them; launches that
    reuse ONE output buffer leave the last launch's words in it."""
    pv.set_scene(sc.goal1_scattered())
    n, k = 200_000, 12
    qs = [random_configs(n, 900 + i) for i in range(k)]
    words = (n + 31) // 32
    pv.set_launch_overlap(False)
    ref = [pv.check_states(_dev(q)).cpu().numpy() for q in qs]
    pv.set_launch_overlap(True)
    try:
        for rep in range(3):
            devq = [_dev(q) for q in qs]
            planes = [(d[:, 0:4].contiguous(), d[:, 4:8].contiguous()) for d in devq]
            torch.cuda.synchronize()
            out = torch.zeros(k * words, dtype=torch.int32, device="cuda")
            for i in range(k):  # distinct output slices, nothing between the launches
                pv.check_states(planes[i], out=out[i * words:(i + 1) * words])
            got = out.cpu().numpy().reshape(k, words)
            assert all(np.array_equal(got[i], ref[i]) for i in range(k)), rep
            one = torch.zeros(words, dtype=torch.int32, device="cuda")
            for i in range(k):  # ONE output buffer: the last launch wins, as in a serial stream
                pv.check_states(planes[i], out=one)
            assert np.array_equal(one.cpu().numpy(), ref[k - 1]), rep
            # inputs produced by kernels queued immediately before the launch (the AoS -> planes copies of _planes,
            # here on top of a device-side permutation of the rows)
            perm = torch.randperm(n, device="cuda")
            for i in range(4):
                w = pv.check_states(devq[i][perm])
                exp = unpack_bits(ref[i].view(np.uint32), n)[perm.cpu().numpy()]
                assert np.array_equal(unpack_bits(w.cpu().numpy().view(np.uint32), n), exp), (rep, i)
    finally:
        pv.set_launch_overlap(True)


def test_sweep_matches_oracle_stream(pv, c64, model):
    scene = sc.goal1_scattered()
    pv.set_scene(scene)
    first, n, seed = 64 * 1000, 100_000, 20251212
    for fingers_open in (True, False):
        bits, count, qd = pv.sweep(first, n, seed, fingers_open=fingers_open, want_configs=True)
        qo = po.sweep_configs(first, n, seed, model, fingers_open=fingers_open)
        assert np.array_equal(qd.cpu().numpy().view(np.uint32), qo.view(np.uint32)), "device RNG stream differs"
        gpu = unpack_bits(bits, n)
        assert int(count.item()) == int(gpu.sum())
        margin = c64.state_margin(qo.astype(np.float64), scene.as_oracle_scene())
        _assert_verdicts(gpu, margin, "sweep")


def test_sweep_sorted_identical_to_unsorted(pv):
    """The tile-sorted sweep (culling mode 2: configurations parked in shared memory, visited in key order) returns
    the same verdict words and count as the unsorted kernels, across super-tile and chunk boundaries, in every scene."""
    for name in ("goal1_scattered", "goal3_tower", "goal4_task1_pentagon"):
        pv.set_scene(sc.FIXTURES[name]())
        for n in (31, 4097, 148 * 4096 * 2 + 777):
            ref = None
            for mode in (0, 1, 2):
                pv.set_culling(mode)
                bits, count = pv.sweep(32 * 5, n, 7, fingers_open=(n % 2 == 1))
                got = (bits.cpu().numpy().view(np.uint32).copy(), int(count.item()))
                if ref is None:
                    ref = got
                assert np.array_equal(got[0], ref[0]) and got[1] == ref[1], (name, n, mode)
            pv.set_culling(2)
            assert ref[1] == int(unpack_bits(ref[0], n).sum())


def test_acceptance_poses(pv):
    """App. F: the poses the reference plans from are valid in every scene; a deep table hit is not."""
    for name, f in sc.FIXTURES.items():
        pv.set_scene(f())
        q = np.stack([pm.Q_SAFE_HOME, pm.Q_SAFE_HOME_039, pm.Q_SCENE_INIT]).astype(np.float32)
        v = unpack_bits(pv.check_states(_dev(q)), 3)
        assert v.all(), name
    bad = np.array([[0, 1.7, 0, -0.1, 0, 0.5, 0, 0.04, 0.04]], dtype=np.float32)
    assert not unpack_bits(pv.check_states(_dev(bad)), 1)[0]


def test_error_paths(pv):
    from rbe550_final_project_b200.validity import PandaValidity, PandaValidityError
    h = PandaValidity(0)
    with pytest.raises(PandaValidityError):
        h.check_states(_dev(random_configs(8, 1)))  # no scene yet
    big = sc.goal1_scattered()
    big.obb = np.repeat(big.obb, 6, axis=0)  # 36 boxes > PV_MAX_OBB
    with pytest.raises(PandaValidityError):
        h.set_scene(big)
    h.set_scene(sc.goal1_scattered())
    with pytest.raises(PandaValidityError):
        h.set_attached(17)
    assert h.check_states(_dev(np.zeros((0, 9), np.float32))).numel() == 0
    h.close()


def test_cuda_reproduces_committed_golden_vectors(pv):
    """tests/golden/validity_golden.npz (tools/make_golden.py): FK, state and edge verdicts, sweep stream."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "validity_golden.npz"))
    assert str(G["model_fingerprint"]) == pm.model_fingerprint()
    q, qb = G["q"], G["qb"]
    n = len(q)
    pv.set_scene(sc.goal1_scattered())
    pose = pv.fk(_dev(q)).cpu().numpy()
    assert np.abs(pose[:, :, :3] - G["fk_p"]).max() < 2e-6
    assert np.abs(pose[:, :, 3:].reshape(n, 11, 3, 3) - G["fk_R"]).max() < 2e-6
    for name in SCENES:
        pv.set_scene(sc.FIXTURES[name]())
        for att, flags, key in ((-1, (True, False), "margin"), (2, (True, False), "margin_attached2"),
                                (-1, (False, False), "margin_noself")):
            pv.set_attached(att)
            pv.set_flags(*flags)
            gpu = unpack_bits(pv.check_states(_dev(q)), n)
            _assert_verdicts(gpu, G[f"{name}/{key}"], f"golden {name}/{key}")
        pv.set_attached(-1)
        pv.set_flags(True, False)
        k = 256
        for steps, key in ((64, "edge64"), (0, "edge_res")):
            gpu = unpack_bits(pv.check_edges(_dev(q[:k]), _dev(qb[:k]), n_steps=steps), k)
            _assert_verdicts(gpu, G[f"{name}/{key}"], f"golden {name}/{key}")
            m = pv.edge_margins(_dev(q[:k]), _dev(qb[:k]), n_steps=steps).cpu().numpy()
            ref = G[f"{name}/{key}"]
            inside = ref > -1e29  # states outside the joint limits carry the sentinel -1e30 in both
            assert (m[~inside] < -1e29).all() and np.abs(m[inside] - ref[inside]).max() < 2e-5
    _, _, qd = pv.sweep(64, 512, 20251212, fingers_open=False, want_configs=True)
    assert np.array_equal(qd.cpu().numpy().view(np.uint32), G["sweep_q"].view(np.uint32))


def _random_rotation(rng):
    q = rng.normal(size=4)
    return sc.quat_wxyz_to_mat(q / np.linalg.norm(q))


def test_general_obbs_max_scene(pv, c64):
    """32 boxes (PV_MAX_OBB) of assorted sizes with arbitrary 3-D rotations (toppled blocks): exercises the general
    sphere-vs-OBB path and the 15-axis SAT that the yaw-only fixtures never reach."""
    rng = np.random.default_rng(99)
    recs = []
    for k in range(32):
        center = [rng.uniform(0.25, 0.75), rng.uniform(-0.5, 0.5), rng.uniform(0.03, 0.6)]
        size = rng.uniform(0.03, 0.12, size=3)
        R = _random_rotation(rng) if k % 4 else sc.yaw_mat(rng.uniform(-180, 180))
        recs.append(sc.make_obb(center, size, R))
    snap = sc.SceneSnapshot(obb=np.array(recs, dtype=np.float32), names=[f"x{k}" for k in range(32)],
                            entity_idx=list(range(1, 33)))
    pv.set_scene(snap)
    pv.set_flags(True, False)
    n = 150_000
    q = random_configs(n, 2024, fingers="random")
    for att in (-1, 5):
        pv.set_attached(att)
        margin = c64.state_margin(q.astype(np.float64), snap.as_oracle_scene(), attached=att)
        for mode in (2, 1, 0):
            pv.set_culling(mode)
            gpu = unpack_bits(pv.check_states(_dev(q)), n)
            _assert_verdicts(gpu, margin, f"general obbs att={att} mode={mode}")
        pv.set_culling(2)
        assert 0.05 < (margin >= 0).mean() < 0.9
    pv.set_attached(-1)
    m = pv.state_margins(_dev(q[:20000])).cpu().numpy()
    ref = c64.state_margin(q[:20000].astype(np.float64), snap.as_oracle_scene())
    assert np.abs(m - ref).max() < 2e-5
    qb = np.clip(q[:4000] + 0.15, pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
    ge = unpack_bits(pv.check_edges(_dev(q[:4000]), _dev(qb), n_steps=0), 4000)
    _assert_verdicts(ge, c64.edge_margin(q[:4000].astype(np.float64), qb.astype(np.float64), snap.as_oracle_scene(), n_steps=0),
                     "general obbs edges")


def test_scene_level_cull_random_scenes(pv):
    """The scene-level cull (link-group balls against the padded bounds of all boxes) and the sorted visiting order
    never change a verdict: brute force, per-lane culling and the sorted kernel agree bit for bit on random scenes of
    1..8 boxes -- tiny and isolated, tall, floating above the table, half sunk into it, rotated in 3-D, hugging the
    base -- for states, edges and the device-generated sweep, also with joint values beyond the limits."""
    rng = np.random.default_rng(4711)
    q = random_configs(120_011, 5, fingers="random")
    q[::13, :7] *= 1.5  # some joints outside their limits (legal input when the limit flag is off)
    qd = _dev(q)
    qb = _dev(np.clip(q[:6000] + rng.normal(0, 0.2, (6000, 9)), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32))
    for trial in range(12):
        nb = int(rng.integers(1, 9))
        recs = []
        for k in range(nb):
            rad = rng.uniform(0.12, 0.9)
            ang = rng.uniform(-np.pi, np.pi)
            size = rng.uniform(0.02, 0.05 if trial % 3 == 0 else 0.3, size=3)
            center = [rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-0.05, 1.1)]
            R = _random_rotation(rng) if (k + trial) % 2 else sc.yaw_mat(rng.uniform(-180, 180))
            recs.append(sc.make_obb(center, size, R))
        snap = sc.SceneSnapshot(obb=np.array(recs, dtype=np.float32), names=[f"x{k}" for k in range(nb)],
                                entity_idx=list(range(1, nb + 1)))
        pv.set_scene(snap)
        pv.set_attached(0 if trial % 4 == 1 else -1)
        res, sres = [], []
        for mode in (0, 1, 2):
            pv.set_culling(mode)
            res.append(pv.check_states(qd).cpu().numpy())
            sres.append(pv.sweep(0, 50_000, 99 + trial)[0].cpu().numpy())
        pv.set_culling(2)
        assert (res[0] == res[1]).all() and (res[0] == res[2]).all(), ("states", trial)
        # the edge kernel always culls: its bits against its own brute-force margins, outside the contact band
        eb = unpack_bits(pv.check_edges(qd[:6000], qb, n_steps=0), 6000)
        em = pv.edge_margins(qd[:6000], qb, n_steps=0).cpu().numpy()
        _assert_verdicts(eb, em.astype(np.float64), f"edges trial {trial}")
        assert (sres[0] == sres[1]).all() and (sres[0] == sres[2]).all(), ("sweep", trial)
    pv.set_attached(-1)


def test_order_and_batch_size_invariance(pv):
    """A verdict depends on its configuration only: permuting the batch permutes the bits; splitting the batch at
    arbitrary (unaligned) points changes nothing; empty scene = self / table rule only."""
    pv.set_scene(sc.goal4_task1_pentagon())
    n = 50_001
    q = random_configs(n, 7, fingers="random")
    base = unpack_bits(pv.check_states(_dev(q)), n)
    perm = np.random.default_rng(1).permutation(n)
    assert np.array_equal(unpack_bits(pv.check_states(_dev(q[perm])), n), base[perm])
    cuts = [0, 1, 33, 1000, 31_999, n]
    parts = [unpack_bits(pv.check_states(_dev(q[a:b])), b - a) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts), base)
    assert np.array_equal(unpack_bits(pv.check_states_host(q), n), base)
    # single-state calls agree with the batch
    for i in (0, 17, 4242):
        assert pv.is_state_valid(q[i]) == bool(base[i])
    empty = sc.SceneSnapshot(obb=np.zeros((0, 16), np.float32))
    pv.set_scene(empty)
    e = unpack_bits(pv.check_states(_dev(q)), n)
    assert (e | ~base).all()  # removing obstacles can only turn invalid states valid


def test_contact_lists(pv, c64, model):
    """pv_state_contacts = the pair list detect_collision() gives: empty iff valid, contains the deepest culprit."""
    scene = sc.goal3_tower()
    pv.set_scene(scene)
    pv.set_flags(True, False)
    q = random_configs(4000, 404, fingers="random")
    lists = pv.contacts(_dev(q))
    margin = c64.state_margin(q.astype(np.float64), scene.as_oracle_scene())
    far = np.abs(margin) > BAND
    for i in np.nonzero(far)[0]:
        assert (len(lists[i]) == 0) == (margin[i] >= 0), i
    m, cu = pv.state_margins(_dev(q), want_culprit=True)
    from rbe550_final_project_b200.validity import decode_culprit_pair
    for i in np.nonzero(far & (margin < 0))[0][:300]:
        assert decode_culprit_pair(int(cu[i].item())) in lists[i]
    # a pose deep in the table lists ground contacts of several links
    bad = np.array([[0, 1.7, 0, -0.1, 0, 0.5, 0, 0.04, 0.04]], dtype=np.float32)
    pairs = pv.contacts(_dev(bad))[0]
    assert any(o == "ground" for _, o in pairs) and len(pairs) >= 2


def test_non_finite_joint_values_are_invalid(pv, c64):
    scene = sc.goal1_scattered()
    pv.set_scene(scene)
    q = np.tile(pm.Q_SAFE_HOME.astype(np.float32), (64, 1))
    q[3, 2] = np.nan
    q[10, 0] = np.inf
    q[20, 8] = -np.inf
    q[33, 5] = 1e30
    gpu = unpack_bits(pv.check_states(_dev(q)), 64)
    expect = np.ones(64, bool)
    expect[[3, 10, 20, 33]] = False
    assert np.array_equal(gpu, expect)
    with np.errstate(all="ignore"):
        ref = c64.state_margin(q.astype(np.float64), scene.as_oracle_scene())
    assert np.array_equal(ref >= 0, expect)
    m = pv.state_margins(_dev(q)).cpu().numpy()
    assert (m[[3, 10, 20, 33]] < -1e29).all()
    qb = q.copy()
    qb[:, 0] += 0.1
    ge = unpack_bits(pv.check_edges(_dev(q), _dev(qb), n_steps=8), 64)
    assert not ge[[3, 10, 20, 33]].any()


def test_edges_host_entry_point_multi_chunk(pv):
    pv.set_scene(sc.goal1_scattered())
    n = 600_013  # three pipeline chunks, ragged
    qa = random_configs(n, 71)
    qb = np.clip(qa + 0.1, pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
    qb[:, 7:] = 0.04
    d = pv.check_edges(_dev(qa), _dev(qb), n_steps=4).cpu().numpy().view(np.uint32)
    h = pv.check_edges_host(qa, qb, n_steps=4)
    assert np.array_equal(d, h)


def test_sharded_edge_entry_single_rank(pv):
    """distributed.check_edges_sharded at world size 1 (N GPUs: bench.py --gpus N `edges`, gloo tests on CPU) is the plain
    call, for (n, 9) tensors and for SoA plane tuples."""
    from rbe550_final_project_b200.distributed import check_edges_sharded
    pv.set_scene(sc.goal4_task1_pentagon())
    n = 50_017
    q = random_configs(n, 71)
    qb = np.clip(q + np.random.default_rng(3).normal(0, 0.3, q.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
    qb[:, 7:] = 0.04
    a, b = _dev(q), _dev(qb)
    ref = pv.check_edges(a, b, n_steps=64).cpu().numpy()
    assert np.array_equal(check_edges_sharded(pv, a, b, n_steps=64).cpu().numpy(), ref)
    planes = lambda t: (t[:, 0:4].contiguous(), t[:, 4:8].contiguous())  # noqa: E731
    assert np.array_equal(check_edges_sharded(pv, planes(a), planes(b), n_steps=64).cpu().numpy(), ref)


def test_fused_gather_epilogues_with_a_local_peer(pv):
    """pv_set_gather with ONE peer that is a local buffer (the N-GPU form runs in bench.py --gpus N and
    tools/multi_gpu_sweep.py): the state, sweep and edge kernels store their verdict words at the configured offset of the
    'peer' as well as locally, words beyond the slot's capacity are dropped, and small edge batches never gather."""
    pv.set_scene(sc.goal1_scattered())
    dev = pv.device
    n = 70_000
    words = (n + 31) // 32
    q = random_configs(n, 81)
    qb = np.clip(q + np.random.default_rng(4).normal(0, 0.3, q.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
    qb[:, 7:] = 0.04
    off, cap = 100, words - 7  # the last 7 words fall outside the slot
    peer = torch.full((off + words + 50,), 0x5A5A5A5A, dtype=torch.int32, device=dev)
    table = torch.tensor([peer.data_ptr()], dtype=torch.int64, device=dev)
    try:
        for name, run in (("states", lambda: pv.check_states(_dev(q))),
                          ("edges", lambda: pv.check_edges(_dev(q), _dev(qb), n_steps=8)),
                          # more than 32 states: certificate pass + list kernels, the words forwarded by pv_edge_emit_kernel
                          ("edges64", lambda: pv.check_edges(_dev(q), _dev(qb), n_steps=64)),
                          ("sweep", lambda: pv.sweep(0, n, 5)[0])):
            pv.set_gather(0, 0, 0, 0, 0)
            ref = run().cpu().numpy()
            peer.fill_(0x5A5A5A5A)
            pv.set_gather(table.data_ptr(), 1, 0, off, cap)
            got = run().cpu().numpy()
            torch.cuda.synchronize()
            g = peer.cpu().numpy()
            assert np.array_equal(got, ref), name
            assert np.array_equal(g[off:off + cap], ref[:cap]), name
            assert (g[:off] == 0x5A5A5A5A).all() and (g[off + cap:] == 0x5A5A5A5A).all(), name
        peer.fill_(0x5A5A5A5A)
        pv.check_edges(_dev(q[:5000]), _dev(qb[:5000]), n_steps=8)  # one edge per warp: no whole words, no gather
        # host-buffer calls number their words per chunk: they never forward, whatever layout they stage internally
        pv.check_states_host(q)
        pv.check_states_host_arm(np.ascontiguousarray(q[:, :7]))
        pv.check_edges_host(q, qb, n_steps=8)
        torch.cuda.synchronize()
        assert (peer.cpu().numpy() == 0x5A5A5A5A).all()
    finally:
        pv.set_gather(0, 0, 0, 0, 0)


def test_rejected_scene_keeps_the_previous_one(pv):
    from rbe550_final_project_b200.validity import PandaValidityError
    good = sc.goal1_scattered()
    pv.set_scene(good)
    q = random_configs(20_000, 5)
    before = pv.check_states(_dev(q)).cpu().numpy()
    bad = sc.goal3_tower()
    bad.obb = bad.obb.copy()
    bad.obb[4, 6:15] *= 1.3  # not a rotation
    with pytest.raises(PandaValidityError):
        pv.set_scene(bad)
    bad2 = sc.goal3_tower()
    bad2.obb = bad2.obb.copy()
    bad2.obb[2, 3] = -0.02
    with pytest.raises(PandaValidityError):
        pv.set_scene(bad2)
    assert np.array_equal(pv.check_states(_dev(q)).cpu().numpy(), before)


@pytest.mark.parametrize("fingers", ["open", "random"])
def test_config2_full_size_against_the_oracle(pv, c64, fingers):
    """BASELINE config 2 at its full size: 1 048 576 configurations, joints uniform in the limits from
    numpy default_rng(20251212), fingers open / uniform in [0, 0.04], goal-1 scene: every verdict equals the fp64
    oracle's outside the 1e-4 m band, through the device AND the host entry point."""
    scene = sc.goal1_scattered()
    pv.set_scene(scene)
    pv.set_flags(True, False)
    n = 1 << 20
    rng = np.random.default_rng(20251212)
    q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9))
    if fingers == "open":
        q[:, 7:] = 0.04
    q = q.astype(np.float32)
    margin = c64.state_margin(q.astype(np.float64), scene.as_oracle_scene())
    gpu = unpack_bits(pv.check_states(_dev(q)), n)
    in_band, flips = _assert_verdicts(gpu, margin, f"config 2 fingers={fingers}")
    assert in_band < n * 1e-3
    host = unpack_bits(pv.check_states_host(q), n)
    assert (host == gpu).all()
    assert 0.80 < gpu.mean() < 0.90


def test_config3_full_size_properties(pv, c64):
    """BASELINE config 3 at its full size (10 485 760 edges x 64 states, finished-pentagon scene): too large for the
    oracle, so size-independent properties -- run-to-run identity, agreement with the oracle on a random sample of the
    same batch, and monotonicity (removing scene boxes can only turn invalid edges valid)."""
    scene = sc.goal4_task1_pentagon()
    pv.set_scene(scene)
    pv.set_flags(True, False)
    n = 10 * (1 << 20)
    rng = np.random.default_rng(20251213)
    qa = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32)
    qa[:, 7:] = 0.04
    qb = qa + rng.normal(0.0, 0.3, size=(n, 9)).astype(np.float32)
    np.clip(qb, pm.Q_LOWER.astype(np.float32), pm.Q_UPPER.astype(np.float32), out=qb)
    qb[:, 7:] = 0.04
    da, db = _dev(qa), _dev(qb)
    w1 = pv.check_edges(da, db, n_steps=64).cpu().numpy()
    w2 = pv.check_edges(da, db, n_steps=64).cpu().numpy()
    assert (w1 == w2).all()
    valid = unpack_bits(w1, n)
    assert 0.70 < valid.mean() < 0.85
    pick = rng.choice(n, size=4096, replace=False)
    ref = c64.edge_margin(qa[pick].astype(np.float64), qb[pick].astype(np.float64), scene.as_oracle_scene(), n_steps=64)
    _assert_verdicts(valid[pick], ref, "config 3 sample")
    # fewer obstacles: no valid edge may turn invalid
    fewer = sc.SceneSnapshot(obb=scene.obb[:5].copy(), names=list(scene.names[:5]), table_z=scene.table_z, base=scene.base,
                             entity_idx=list(scene.entity_idx[:5]))
    pv.set_scene(fewer)
    valid_fewer = unpack_bits(pv.check_edges(da, db, n_steps=64), n)
    ref_fewer = c64.edge_margin(qa[pick].astype(np.float64), qb[pick].astype(np.float64), fewer.as_oracle_scene(), n_steps=64)
    _assert_verdicts(valid_fewer[pick], ref_fewer, "config 3 sample, fewer boxes")
    # exact, not just outside the band: the tests against the remaining boxes are the same arithmetic in both runs
    assert not (valid & ~valid_fewer).any()


def test_config5_full_size_properties(pv, c64, model):
    """BASELINE config 5 at its full size on one GPU (104 857 600 device-generated configurations, goal-1 scene):
    count = popcount of the mask; the mask is the concatenation of the masks of arbitrary 32-aligned shards (what the
    ranks of a multi-GPU run compute); a slice re-checked through the state kernel from the exported configurations gives
    the same words; a sample agrees with the fp64 oracle outside the contact band."""
    scene = sc.goal1_scattered()
    pv.set_scene(scene)
    pv.set_flags(True, False)
    n, seed = 104_857_600, 20251212
    bits, count = pv.sweep(0, n, seed)
    words = bits.cpu().numpy().view(np.uint32)
    assert int(count.item()) == int(np.unpackbits(words.view(np.uint8)).sum())
    assert 0.80 < int(count.item()) / n < 0.90
    cuts = [0, 32 * 1_000_003, 32 * 2_222_222, n]
    parts = [pv.sweep(a, b - a, seed)[0].cpu().numpy().view(np.uint32) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts), words)
    first, m = 32 * 1_500_000, 300_000
    sl_bits, _, q_dev = pv.sweep(first, m, seed, want_configs=True)
    again = pv.check_states(q_dev).cpu().numpy().view(np.uint32)
    assert np.array_equal(again, sl_bits.cpu().numpy().view(np.uint32))
    assert np.array_equal(again, words[first // 32: first // 32 + (m + 31) // 32])
    k = 60_000
    q = q_dev[:k].cpu().numpy()
    assert np.array_equal(q.view(np.uint32), po.sweep_configs(first, k, seed, model).view(np.uint32))
    _assert_verdicts(unpack_bits(again, k), c64.state_margin(q.astype(np.float64), scene.as_oracle_scene()), "config 5 slice")
