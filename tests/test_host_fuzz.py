"""CPU tier: the random scene graphs of tests/test_gpu_fuzz.py through the host flatten + BVH8 builder and the UNMODIFIED
device traversal header compiled for the host (tests/emul: f32 tests with error bounds, ambiguity detection, the exact
pass, distance refinement), against the oracle's f64 linear scan — on camera rays and on random rays.  What the GPU tier
checks with the production kernels, minus shading."""
import numpy as np
import pytest

import helpers as H
from test_host_bvh import _media_ids


@pytest.fixture(scope="module")
def emul():
    return H.build_emul()


@pytest.mark.parametrize("seed", [1, 3, 12, 63, 95, 202])
def test_random_scene_graph_closest_hits(rtb, orc, emul, seed):
    from ray_tracer_archive_b200 import scenes, scene as S
    from test_gpu_fuzz import _random_scene
    world, lights = _random_scene(seed, S, scenes, seed % 24 >= 12, seed % 2 == 1 or seed >= 200, 1200 if seed >= 200 else None)
    cs = rtb.compile_scene(world, lights)
    hs = rtb.Scene(None, cs)
    osc = orc.OracleScene(cs)
    osc.attach_bvh(hs)  # candidate culling only (bit-identical to the linear scan, test_host_bvh.py)
    assert osc.num_prims() == hs.info()["n_prims"]
    mid = _media_ids(cs) if hs.info()["n_media"] else []
    tri_ids = hs.export_bvh()[1][3][1].reshape(-1, 2)[:, 0]
    cam = rtb.Camera.new((2.0, 7.0, 19.0), (0.0, 2.0, 0.0), (0, 1, 0), 40.0, 1.5, 0.0, 19.0, 0.0, 1.0)
    o, d = H.primary_rays(cam, 160, 100)
    rng = np.random.default_rng(7000 + seed)
    n = 12000
    o2 = rng.uniform([-9, 0.05, -9], [9, 9, 9], (n, 3))
    d2 = rng.normal(0, 1, (n, 3))
    d2 = d2 / np.linalg.norm(d2, axis=1, keepdims=True) * np.exp(rng.uniform(np.log(0.01), np.log(100.0), (n, 1)))
    o32 = np.concatenate([o, o2]).astype(np.float32)
    d32 = np.concatenate([d, d2]).astype(np.float32)
    tm = np.concatenate([np.zeros(len(o)), rng.random(n)]).astype(np.float32)
    for label, tol in (("camera + free-space", 0), ("on-surface", 4e-4)):
        ids, ts, nv, nt = H.emul_trace(emul, hs, o32, d32, tm)
        oid, ot = osc.trace_rays(o32.astype(np.float64), d32.astype(np.float64), tm.astype(np.float64))
        surf = ~(np.isin(oid, mid) | np.isin(ids, mid))
        mism = (ids != oid) & surf
        hit = surf & ~mism & (oid != H.NONE)
        rel = np.abs(ts[hit].astype(np.float64) - ot[hit]) / ot[hit]
        dist = ot[hit] * np.linalg.norm(d32[hit].astype(np.float64), axis=1)
        floor = np.where(np.isin(oid[hit], tri_ids), 3e-6 / np.maximum(dist, 1e-30), 0.0)  # baked triangle transforms
        assert mism.sum() <= tol * len(oid), (label, int(mism.sum()))
        assert np.quantile(np.maximum(rel - floor, 0.0), 0.999) <= 1e-5, label
        p = (o32[hit].astype(np.float64) + ot[hit, None] * d32[hit].astype(np.float64)).astype(np.float32)
        nd = rng.normal(0, 1, p.shape)
        o32, d32 = p, (nd / np.linalg.norm(nd, axis=1, keepdims=True)).astype(np.float32)
        tm = rng.random(len(p)).astype(np.float32)


@pytest.mark.parametrize("seed", [5, 63])
def test_oracle_bvh_culling_equals_its_linear_scan_on_random_graphs(rtb, orc, seed):
    """The fuzz tests above let the oracle cull candidates with the shipped BVH8 (the per-primitive arithmetic stays the
    reference's): on random scene graphs that must return bit-identical (id, t) to the oracle's own linear HittableList scan."""
    from ray_tracer_archive_b200 import scenes, scene as S
    from test_gpu_fuzz import _random_scene
    world, lights = _random_scene(seed, S, scenes, False, True, None)
    cs = rtb.compile_scene(world, lights)
    hs = rtb.Scene(None, cs)
    lin, acc = orc.OracleScene(cs), orc.OracleScene(cs)
    acc.attach_bvh(hs)
    rng = np.random.default_rng(seed)
    n = 6000
    o = rng.uniform([-9, 0.05, -9], [9, 9, 9], (n, 3))
    d = rng.normal(0, 1, (n, 3))
    tm = rng.random(n)
    i1, t1 = lin.trace_rays(o, d, tm)
    i2, t2 = acc.trace_rays(o, d, tm)
    assert np.array_equal(i1, i2) and np.array_equal(t1, t2)
