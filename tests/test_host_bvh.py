"""CPU tier: the host flatten + BVH8 builder, and the UNMODIFIED device traversal header compiled for the host
(tests/emul), against the oracle's linear HittableList scan on the same scenes."""
import numpy as np
import pytest

import helpers as H


@pytest.fixture(scope="module")
def emul():
    return H.build_emul()


def _cases():
    from ray_tracer_archive_b200 import scenes
    return [("cornell", scenes.config_cornell(), 150, 150),
            ("random_spheres", scenes.config_random_spheres(), 240, 135),
            ("final_scene", scenes.config_final_scene(), 128, 128),
            ("mesh", scenes.config_mesh(nx=40, nz=20), 160, 90),
            ("cornell_smoke_surfaces", None, 96, 96)]


@pytest.mark.parametrize("name,cfg,W,Hh", _cases(), ids=[c[0] for c in _cases()])
def test_emulated_device_traversal_matches_oracle(rtb, orc, emul, name, cfg, W, Hh):
    from ray_tracer_archive_b200 import scenes
    if cfg is None:
        cfg = scenes.config_cornell()
        cfg.world = scenes.two_spheres()
    cs = rtb.compile_scene(cfg.world, cfg.lights)
    hs = rtb.Scene(None, cs)
    osc = orc.OracleScene(cs)
    assert osc.num_prims() == hs.info()["n_prims"]
    o, d = H.primary_rays(cfg.camera, W, Hh)
    o32, d32 = o.astype(np.float32), d.astype(np.float32)
    tm = np.full(len(o), cfg.camera.time0, dtype=np.float32)
    # the oracle traces exactly the f32 rays the device code consumes; media are excluded on both sides (surface probe)
    oid, ot = osc.trace_rays(o32.astype(np.float64), d32.astype(np.float64), tm.astype(np.float64))
    ids, ts, nv, nt = H.emul_trace(emul, hs, o32, d32, tm)
    n_media = hs.info()["n_media"]
    if n_media:  # oracle's scan includes the media (xi = 0.5); compare only rays whose oracle hit is a surface
        surf = ~np.isin(oid, _media_ids(cs))
    else:
        surf = np.ones(len(oid), bool)
    mism = (ids != oid) & surf
    # identical rays -> identical ids, knife-edge rays included (exact shared edges, the Cornell box's symmetric corner
    # line): every decision f32 rounding leaves open is re-made with the reference's own f64 arithmetic (exact_hit)
    assert mism.sum() == 0, f"{mism.sum()} mismatches of {len(oid)}"
    ok = surf & ~mism & (oid != H.NONE)
    rel = np.abs(ts[ok] - ot[ok]) / ot[ok]
    assert np.quantile(rel, 0.999) < 1e-5 and rel.max() < 2e-4
    assert nt > 0 and (nv > 0 or hs.info()["n_prims"] <= 16)  # scenes of <= 16 primitives are scanned, not traversed


def _media_ids(cs):
    """ids the flattener gives to ConstantMedium nodes = position in the depth-first leaf order; recomputed here."""
    from ray_tracer_archive_b200 import _ffi as F
    ids, counter = [], [0]

    def walk(i, in_boundary):
        n = cs.nodes[i]
        kids = [int(cs.child_index[n["first_child"] + k]) for k in range(int(n["n_children"]))]
        t = int(n["type"])
        if t == F.NODE_CONSTANT_MEDIUM:
            ids.append(counter[0]); counter[0] += 1
        elif t in (F.NODE_LIST, F.NODE_BVH, F.NODE_TRANSLATE, F.NODE_ROTATE_Y, F.NODE_FLIP_FACE):
            for k in kids:
                walk(k, in_boundary)
        elif t == F.NODE_BOX:
            counter[0] += 6
        elif t == F.NODE_MESH:
            counter[0] += len(cs.meshes[int(n["p"][0])][1])
        else:
            counter[0] += 1
    walk(cs.root, False)
    return np.array(ids, dtype=np.uint32)


def test_smem_split_does_not_change_results(rtb, emul):
    """Nodes below/above the shared-memory staging threshold are fetched through different paths; same answer."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_final_scene()
    hs = rtb.Scene(None, rtb.compile_scene(cfg.world, cfg.lights))
    o, d = H.primary_rays(cfg.camera, 64, 64)
    a = H.emul_trace(emul, hs, o, d, n_snodes=10 ** 6)
    b = H.emul_trace(emul, hs, o, d, n_snodes=7)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_random_interior_rays(rtb, orc, emul):
    """Bounce-like rays: random origins inside the scene, random (un-normalised) directions, random times."""
    from ray_tracer_archive_b200 import scenes
    rng = np.random.default_rng(11)
    for cfg, lo, hi in [(scenes.config_cornell(), (1, 1, 1), (554, 554, 554)),
                        (scenes.config_random_spheres(), (-10, 0.05, -10), (10, 3, 10)),
                        (scenes.config_final_scene(n_small=200, boxes_per_side=8), (-200, 110, -100), (500, 500, 500))]:
        cs = rtb.compile_scene(cfg.world, cfg.lights)
        hs, osc = rtb.Scene(None, cs), orc.OracleScene(cs)
        n = 20000
        o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
        d = (rng.normal(0, 1, (n, 3)) * rng.uniform(0.2, 30, (n, 1))).astype(np.float32)
        tm = rng.random(n).astype(np.float32)
        oid, ot = osc.trace_rays(o.astype(np.float64), d.astype(np.float64), tm.astype(np.float64))
        ids, ts, _, _ = H.emul_trace(emul, hs, o, d, tm)
        surf = ~np.isin(oid, _media_ids(cs)) if hs.info()["n_media"] else np.ones(n, bool)
        # rays whose oracle hit is a medium sample may legitimately see a farther surface in the surface-only probe
        mism = (ids != oid) & surf
        assert mism.sum() <= 3, f"{cfg.name}: {mism.sum()} mismatches"
        ok = surf & ~mism & (oid != H.NONE)
        # hits a few 1e-3 away from an origin with |coordinates| ~ 500 are limited by the f32 ulp of the POSITION, not
        # of t: tolerance = 1e-5 relative + 2 ulp(|o|) of travelled distance
        tol = 1e-5 * ot[ok] + 2.4e-7 * np.abs(o[ok]).max(axis=1) / np.linalg.norm(d[ok].astype(np.float64), axis=1)
        assert (np.abs(ts[ok] - ot[ok]) <= tol).mean() > 0.9995


def test_quantised_boxes_are_conservative(rtb):
    """Every primitive's padded box lies inside the dequantised box of the leaf slot that references it."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_final_scene(n_small=300, boxes_per_side=10)
    hs = rtb.Scene(None, rtb.compile_scene(cfg.world, cfg.lights))
    nodes, prims = hs.export_bvh()
    nd = np.frombuffer(nodes.tobytes(), dtype=np.dtype([("o", "<f4", 3), ("e", "u1", 3), ("imask", "u1"), ("child_base", "<u4"),
                                                        ("prim_base", "<u4"), ("meta", "u1", 8), ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))]))
    assert nd.dtype.itemsize == 80
    spheres = prims[0][0].reshape(-1, 4)
    checked = 0
    for n in nd:
        step = np.ldexp(1.0, n["e"].astype(int) - 127)
        ptype, pbase = int(n["prim_base"]) >> 29, int(n["prim_base"]) & ((1 << 29) - 1)
        for s in range(8):
            cnt, off = int(n["meta"][s]) >> 5, int(n["meta"][s]) & 31
            if cnt == 0 or ptype != 0:
                continue
            lo = n["o"] + n["qlo"][:, s] * step
            hi = n["o"] + n["qhi"][:, s] * step
            for k in range(cnt):
                c, r = spheres[pbase + off + k, :3], spheres[pbase + off + k, 3]
                assert np.all(lo <= c - r) and np.all(hi >= c + r)
                checked += 1
    assert checked == len(spheres)


def test_oracle_bvh_mode_is_bit_identical_to_linear_scan(rtb, orc):
    """The oracle may use the product's exported BVH to CULL candidates (needed for the 1M-triangle config); the
    per-primitive tests and the tie rule stay the reference's.  Closest hits and whole renders must not change by a bit."""
    from ray_tracer_archive_b200 import scenes
    cfgs = [scenes.config_cornell(), scenes.config_random_spheres(), scenes.config_final_scene(n_small=300, boxes_per_side=10),
            scenes.config_mesh(nx=40, nz=20)]
    smoke = scenes.config_cornell()
    smoke.world, smoke.lights = scenes.cornell_smoke(), scenes.cornell_smoke_lights()
    for cfg in cfgs + [smoke]:
        cs = rtb.compile_scene(cfg.world, cfg.lights)
        hs, osc = rtb.Scene(None, cs), orc.OracleScene(cs)
        a_ids, a_t = osc.primary_hits(cfg.camera, 96, 64)
        prm = rtb.make_params(24, 24, 8, cfg.max_depth, cfg.background, seed=4)
        acc1, seg1, _ = osc.render(cfg.camera, prm)
        osc.attach_bvh(hs)
        b_ids, b_t = osc.primary_hits(cfg.camera, 96, 64)
        acc2, seg2, _ = osc.render(cfg.camera, prm)
        assert np.array_equal(a_ids, b_ids) and np.array_equal(a_t, b_t), cfg.name
        assert np.array_equal(acc1, acc2) and seg1 == seg2, cfg.name


def test_parallel_subtree_collapse_large_mesh(rtb, orc, emul, monkeypatch):
    """Trees of >= 50 000 primitives collapse their subtrees on worker threads (bvh_build.cpp): the result must be
    independent of the thread count (byte-identical nodes), give the same closest hits as the serial breadth-first
    layout, and keep its first 1024 nodes in pure breadth-first order (every child index of that prefix lies after its
    parent, levels contiguous) — that prefix is what the extend kernel stages in shared memory."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_mesh(nx=180, nz=180)  # 64 800 triangles + the Cornell walls
    cs = rtb.compile_scene(cfg.world, cfg.lights)

    def build(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        hs = rtb.Scene(None, cs)
        hs.build_bvh()
        nodes, prims = hs.export_bvh()
        for k in env:
            monkeypatch.delenv(k)
        return hs, nodes.tobytes(), prims

    hs_par, n_par, p_par = build({})
    hs_one, n_one, _ = build({"RTB_BVH_THREADS": "1"})
    hs_ser, n_ser, p_ser = build({"RTB_BVH_PAR": "0"})
    assert hs_par.info()["n_prims"] >= 50000
    assert n_par == n_one  # layout does not depend on the number of worker threads
    assert len(n_par) == len(n_ser)  # same tree, different node order
    # same closest hits as the serial layout, and as the oracle
    o, d = H.primary_rays(cfg.camera, 96, 54)
    rng = np.random.default_rng(5)
    o2 = rng.uniform((70, 160, 70), (480, 500, 480), (4000, 3))
    d2 = rng.normal(0, 1, (4000, 3)) * 50
    o32 = np.concatenate([o, o2]).astype(np.float32)
    d32 = np.concatenate([d, d2]).astype(np.float32)
    a = H.emul_trace(emul, hs_par, o32, d32)
    b = H.emul_trace(emul, hs_ser, o32, d32)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    osc = orc.OracleScene(cs)
    osc.attach_bvh(hs_par)
    oid, ot = osc.trace_rays(o32.astype(np.float64), d32.astype(np.float64), np.zeros(len(o32)))
    assert (a[0] != oid).sum() <= 3
    # breadth-first prefix: children of node i (i < 1024) start after every child block of the nodes before it
    nd = np.frombuffer(n_par, dtype=np.dtype([("o", "<f4", 3), ("e", "u1", 3), ("imask", "u1"), ("child_base", "<u4"),
                                              ("prim_base", "<u4"), ("meta", "u1", 8), ("q", "u1", 48)]))
    nxt = 1
    for i in range(min(1024, len(nd))):
        k = bin(int(nd["imask"][i])).count("1")
        if k == 0 or nxt >= 1024:
            continue
        assert int(nd["child_base"][i]) == nxt, i
        nxt += k


def test_build_options_change_the_tree_not_the_hits(rtb, orc, emul):
    """rtb_scene_set_build_options (f2: builder quality knobs): different leaf sizes / collapse thresholds / no global
    primitives give different trees and the same closest hits."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_mesh(nx=60, nz=30)
    cs = rtb.compile_scene(cfg.world, cfg.lights)
    o, d = H.primary_rays(cfg.camera, 120, 68)
    o32, d32 = o.astype(np.float32), d.astype(np.float32)
    ref, nodes = None, set()
    for kw in ({}, {"max_leaf_triangles": 1}, {"max_leaf_triangles": 3, "open_min_extent": 0.0},
               {"keep_huge_primitives_out": False, "open_min_extent": 0.5}):
        hs = rtb.Scene(None)
        hs.set_compiled(cs)
        hs.set_build_options(**kw)
        hs.build_bvh()
        nodes.add(hs.info()["n_bvh_nodes"])
        ids, ts, _, _ = H.emul_trace(emul, hs, o32, d32)
        if ref is None:
            ref = (ids, ts)
        else:
            assert np.array_equal(ids, ref[0]) and np.allclose(ts, ref[1], rtol=1e-6)
    assert len(nodes) >= 3
    with pytest.raises(rtb.RtbError):
        rtb.Scene(None).set_build_options(max_leaf_triangles=7)
