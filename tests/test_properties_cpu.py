"""Property-based CPU tests (hypothesis) of the host-side logic: path resampling, sharding, bit packing, and the
oracle's own invariants (SURVEY.md App. F suggestions)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import panda_oracle as po
from rbe550_final_project_b200 import panda_model as pm
from rbe550_final_project_b200 import scenes as sc
from rbe550_final_project_b200.distributed import shard_range, words_per_shard
from rbe550_final_project_b200.pathutil import interpolate, path_length
from rbe550_final_project_b200.validity import unpack_bits


@settings(max_examples=60, deadline=None)
@given(n_pts=st.integers(2, 12), count=st.integers(2, 300), seed=st.integers(0, 10_000))
def test_interpolate_properties(n_pts, count, seed):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-2, 2, size=(n_pts, 9))
    out = interpolate(pts, count)
    ref = po.interpolate_path(pts, count)
    assert out.shape == ref.shape and np.allclose(out, ref, atol=1e-14)
    assert np.array_equal(out[0], pts[0]) and np.array_equal(out[-1], pts[-1])
    assert len(out) == (count if count >= n_pts else n_pts)
    # resampling never lengthens the path and keeps it on the original polyline
    assert path_length(out) <= path_length(pts) + 1e-9
    for p in out:
        d = min(_dist_to_segment(p, a, b) for a, b in zip(pts[:-1], pts[1:]))
        assert d < 1e-9


def _dist_to_segment(p, a, b):
    ab = b - a
    t = 0.0 if not ab.any() else float(np.clip(np.dot(p - a, ab) / np.dot(ab, ab), 0, 1))
    return float(np.linalg.norm(p - (a + t * ab)))


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 10_000_000), world=st.sampled_from([1, 2, 3, 4, 8]))
def test_shards_tile_the_range(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert sum(c for _, c in spans) == n
    pos = 0
    for first, count in spans:
        if count:
            assert first == pos and first % 32 == 0
            pos += count
    assert words_per_shard(n, world) * 32 * world >= n


@settings(max_examples=100, deadline=None)
@given(bits=st.lists(st.booleans(), min_size=1, max_size=500))
def test_bit_packing_roundtrip(bits):
    v = np.array(bits, dtype=bool)
    w = po.pack_bits(v)
    assert np.array_equal(unpack_bits(w, len(v)), v) and np.array_equal(po.unpack_bits(w, len(v)), v)
    if len(v) % 32:
        assert (int(w[-1]) >> (len(v) % 32)) == 0


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 100_000), scale=st.floats(1.0, 2.0))
def test_oracle_monotone_under_inflation_and_edge_is_and_of_states(seed, scale):
    model = pm.model_arrays()
    rng = np.random.default_rng(seed)
    q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(40, 9))
    s = sc.goal3_tower().as_oracle_scene()
    m0 = po.state_margin(q, s, model)
    s2 = {"obb": s["obb"].copy(), "table_z": s["table_z"]}
    s2["obb"][:, 3:6] *= scale
    assert (po.state_margin(q, s2, model) <= m0 + 1e-12).all()
    qb = np.clip(q + rng.normal(0, 0.2, q.shape), pm.Q_LOWER, pm.Q_UPPER)
    e = po.edge_margin(q[:6], qb[:6], s, model, n_steps=8)
    t = (np.arange(1, 9) / 8.0)[None, :, None]
    states = q[:6, None, :] + t * (qb[:6, None, :] - q[:6, None, :])
    sm = po.state_margin(states.reshape(-1, 9), s, model).reshape(6, 8).min(1)
    assert np.allclose(e, sm, atol=1e-12)


def test_host_placement_never_raises_and_keeps_the_affinity_without_a_gpu():
    """hostmem.bind_to_gpu_numa is an optimisation for multi-socket hosts: without NVML / a GPU it must change nothing."""
    import os

    from rbe550_final_project_b200.hostmem import bind_to_gpu_numa

    before = os.sched_getaffinity(0)
    info = bind_to_gpu_numa(0)
    assert info["bound"] is False and "why" in info
    assert os.sched_getaffinity(0) == before
