"""GPU tier (-m gpu): the CUDA path, called through the C ABI (ctypes), against the f64 oracle on identical scenes.

P1  primary-ray closest hit: primitive ids bit-exact, t within 1e-5 relative          
P2  converged images at 4096 spp: mean relative luminance error <= 1 %, no pixel beyond 5 sigma of its MC error
P3  white furnace (closed form)      D1  sample split across calls == one call      plus ABI error behaviour.
The oracle finishes in seconds only at reduced sizes; full-size configs are covered by P1 where the linear-scan
oracle is affordable (C1, C2) and by size-independent properties (height-field closest-hit, C4 at 1M triangles).
"""
import ctypes as C
import math

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _scene_pair(rtb, orc, ctx, cfg):
    cs = rtb.compile_scene(cfg.world, cfg.lights)
    return rtb.Scene(ctx, cs), orc.OracleScene(cs), cs


# ------------------------------------------------------------------------------------------------------------- P1
def _p1(rtb, orc, ctx, cfg, W, Hh, oracle_bvh=False):
    """Primary-ray closest hit on IDENTICAL rays: the device's own f32 pixel-centre rays are handed to the f64 oracle.
    Bar (BASELINE.json): primitive ids equal on 100 % of the pixels — no stability mask, no tolerance — and
    |t_gpu - t_ref| <= 1e-5 t_ref.  Only a ConstantMedium's sampled (xi = 0.5) distance is compared in f32 on both sides
    of a knife edge (hit_distance vs distance_inside, constant_medium.rs:58-60): pixels whose id on either side is a medium
    are counted separately."""
    dev, osc, cs = _scene_pair(rtb, orc, ctx, cfg)
    if oracle_bvh:  # candidate culling only; bit-identical to the linear scan (tests/test_host_bvh.py)
        osc.attach_bvh(dev)
    ids, ts, st = dev.primary_hits(cfg.camera, W, Hh)
    o, d, tm = ctx.primary_rays(cfg.camera, W, Hh)
    oid, ot = osc.trace_rays(o.astype(np.float64), d.astype(np.float64), tm.astype(np.float64))
    ids, ts = ids.reshape(-1), ts.reshape(-1)
    media = np.zeros(len(oid), bool)
    if dev.info()["n_media"]:
        from test_host_bvh import _media_ids
        mid = _media_ids(cs)
        media = np.isin(oid, mid) | np.isin(ids, mid)
    mism = ids != oid
    n_surface, n_media = int((mism & ~media).sum()), int((mism & media).sum())
    hit = (oid != H.NONE) & ~mism
    rel = np.abs(ts[hit].astype(np.float64) - ot[hit]) / ot[hit]
    worst = float(rel.max()) if hit.any() else 0.0
    print(f"{cfg.name}: {W}x{Hh} identical rays: {n_surface} id mismatches (+{n_media} on media knife edges) of {oid.size}, max t err {worst:.2e}, "
          f"{st['exact_rays']} rays through the exact pass, {st['nodes_visited'] / oid.size:.2f} nodes/ray {st['prims_tested'] / oid.size:.2f} prims/ray")
    assert n_surface == 0, f"{n_surface} primitive-id mismatches, first at {np.argwhere(mism & ~media)[:5].ravel()}"
    assert n_media <= 3
    assert worst <= 1e-5
    return ids.reshape(Hh, W), oid.reshape(Hh, W)


def test_p1_random_spheres_full_size(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_random_spheres()
    _p1(rtb, orc, ctx, cfg, cfg.width, cfg.height)


def test_p1_cornell_full_size(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_cornell()
    ids, oid = _p1(rtb, orc, ctx, cfg, cfg.width, cfg.height)
    assert set(np.unique(oid).tolist()) >= {0, 1, 2, 3, 4, 5, 12}


def test_p1_final_scene(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_final_scene()
    _p1(rtb, orc, ctx, cfg, 400, 400)


def test_p1_final_scene_full_size(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_final_scene()
    _p1(rtb, orc, ctx, cfg, cfg.width, cfg.height, oracle_bvh=True)


def test_p1_mesh_full_size_1M_triangles(rtb, orc, ctx):
    """C4 at BASELINE.json's full size: 1920x1080 primary rays against 1 000 000 triangles + the Cornell walls."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_mesh()
    _p1(rtb, orc, ctx, cfg, cfg.width, cfg.height, oracle_bvh=True)


def test_p1_mesh_small(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_mesh(nx=100, nz=50)
    _p1(rtb, orc, ctx, cfg, 320, 180)


@pytest.mark.parametrize("which", ["final_scene", "mesh"])
def test_extend_schedulers_agree(rtb, ctx, monkeypatch, which):
    """The three extend schedulers (one ray per thread / dynamic fetch with parked leaf tests / warp queue with the rays
    in shared memory) run the same per-ray sequence of node visits and primitive tests: identical hits, distances, work
    counters and exact-pass sets on identical rays, and the same image from the same seed."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_final_scene() if which == "final_scene" else scenes.config_mesh(nx=200, nz=100)
    cs = rtb.compile_scene(cfg.world, cfg.lights)
    W, Hh = 400, 300
    out = {}
    for mode in ("static", "dynamic", "wq"):
        monkeypatch.setenv("RTB_EXTEND_MODE", mode)  # read when the scene is committed (configure_launch)
        dev = rtb.Scene(ctx, cs)
        ids, ts, st = dev.primary_hits(cfg.camera, W, Hh)
        acc, rst = dev.render(cfg.camera, rtb.make_params(W, Hh, 8, cfg.max_depth, cfg.background, seed=5))
        out[mode] = (ids, ts, st, np.asarray(acc, dtype=np.float64), rst)
        dev.close()
    ids0, ts0, st0, acc0, rst0 = out["static"]
    for mode in ("dynamic", "wq"):
        ids, ts, st, acc, rst = out[mode]
        assert np.array_equal(ids, ids0) and np.array_equal(ts, ts0), mode
        for k in ("nodes_visited", "prims_tested", "exact_rays", "refined_rays"):
            assert st[k] == st0[k], (mode, k, st[k], st0[k])
        assert rst["segments"] == rst0["segments"], mode
        # same paths; only the order of the float atomics into the accumulator differs
        assert np.allclose(acc[..., :3], acc0[..., :3], rtol=1e-4, atol=1e-5), mode


def test_p1_other_scenes(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    base = scenes.config_cornell()
    for name, world, cam in [
        ("two_spheres", scenes.two_spheres(), rtb.Camera.new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.0, 10.0)),
        ("simple_light", scenes.simple_light(), rtb.Camera.new((26, 3, 6), (0, 2, 0), (0, 1, 0), 20.0, 1.5, 0.0, 10.0)),
        ("cornell_smoke", scenes.cornell_smoke(), base.camera),
        ("earth", scenes.earth(), rtb.Camera.new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.0, 10.0)),
        ("two_perlin_spheres", scenes.two_perlin_spheres(), rtb.Camera.new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.0, 10.0)),
    ]:
        base.world, base.lights, base.camera, base.name = world, None, cam, name
        _p1(rtb, orc, ctx, base, 240, 160)


def test_tie_break_and_flat_api(rtb, orc, ctx):
    """Equal t: the later primitive wins (hittable_list.rs:44-47), also through the flat SoA entry points."""
    from ray_tracer_archive_b200 import scene as S
    m = S.Lambertian.construct((0.5, 0.5, 0.5))
    world = S.HittableList([S.XzRect.construct(0, 1, 0, 1, 0.0, m), S.XzRect.construct(0, 1, 0, 1, 0.0, m),
                            S.XzRect.construct(-1, 2, -1, 2, 0.0, m), S.Sphere.construct((5, 0, 5), 1.0, m)])
    cs = rtb.compile_scene(world)
    dev, osc = rtb.Scene(ctx, cs), orc.OracleScene(cs)
    o = [[0.5, 1, 0.5], [1.5, 1, 0.5], [5, 5, 5], [9, 9, 9]]
    d = [[0, -1, 0], [0, -2, 0], [0, -1, 0], [0, 1, 0]]
    ids, ts, _ = dev.trace_rays(o, d)
    oid, ot = osc.trace_rays(o, d)
    assert ids.tolist() == oid.tolist() == [2, 2, 3, H.NONE]
    np.testing.assert_allclose(ts[:3], ot[:3], rtol=1e-6)
    # the same scene through rtb_scene_set_quads / rtb_scene_set_spheres
    lib = rtb._ffi.load()
    flat = rtb.Scene(ctx)
    flat.set_tables(cs)
    q = np.array([[0, 0, 0], [0, 0, 0], [-1, 0, -1]], np.float32)
    u = np.array([[1, 0, 0], [1, 0, 0], [3, 0, 0]], np.float32)
    v = np.array([[0, 0, 1], [0, 0, 1], [0, 0, 3]], np.float32)
    mat = np.zeros(3, np.uint32)
    pid = np.array([0, 1, 2], np.uint32)
    rtb._ffi.check(lib.rtb_scene_set_quads(flat.h, rtb._ffi.ptr(q), rtb._ffi.ptr(u), rtb._ffi.ptr(v), rtb._ffi.ptr(mat), None,
                                           rtb._ffi.ptr(pid), 3))
    sp = np.array([[5, 0, 5, 1]], np.float32)
    rtb._ffi.check(lib.rtb_scene_set_spheres(flat.h, rtb._ffi.ptr(sp), rtb._ffi.ptr(np.zeros(1, np.uint32)), None,
                                             rtb._ffi.ptr(np.array([3], np.uint32)), 1))
    flat.commit()
    ids2, ts2, _ = flat.trace_rays(o, d)
    assert ids2.tolist() == [2, 2, 3, H.NONE]
    np.testing.assert_allclose(ts2[:3], ts[:3], rtol=1e-6)


def test_flat_api_triangles_moving_spheres_media(rtb, orc, ctx):
    """rtb_scene_set_triangles / _moving_spheres / _media (callers that flatten themselves) give the same closest hits
    and the same image as the graph entry point and the oracle."""
    from ray_tracer_archive_b200 import scene as S
    F = rtb._ffi
    lam = S.Lambertian.construct((0.6, 0.5, 0.4))
    boundary = S.Sphere.construct((0.0, 1.0, 0.0), 1.5, S.Dielectric.construct(1.5))
    world = S.HittableList([
        S.Triangle((-3, 0, -3), (3, 0, -3), (0, 0, 3), lam), S.Triangle((-3, 3, -3), (0, 3, 3), (3, 3, -3), lam),
        S.MovingSphere.construct((-1.5, 1, 0), (-1.5, 1.6, 0), 0.0, 1.0, 0.4, lam),
        S.ConstantMedium.construct_color(boundary, 0.8, (0.9, 0.9, 0.9))])
    cs = rtb.compile_scene(world)
    graph, osc = rtb.Scene(ctx, cs), orc.OracleScene(cs)
    lib = rtb._ffi.load()
    flat = rtb.Scene(ctx)
    flat.set_tables(cs)
    v0 = np.array([[-3, 0, -3], [-3, 3, -3]], np.float32)
    v1 = np.array([[3, 0, -3], [0, 3, 3]], np.float32)
    v2 = np.array([[0, 0, 3], [3, 3, -3]], np.float32)
    z2 = np.zeros(2, np.uint32)
    F.check(lib.rtb_scene_set_triangles(flat.h, F.ptr(v0), F.ptr(v1), F.ptr(v2), F.ptr(z2), None, F.ptr(np.array([0, 1], np.uint32)), 2))
    F.check(lib.rtb_scene_set_moving_spheres(flat.h, F.ptr(np.array([[-1.5, 1, 0, 0.4]], np.float32)),
                                             F.ptr(np.array([[-1.5, 1.6, 0]], np.float32)), F.ptr(np.array([[0, 1]], np.float32)),
                                             F.ptr(np.zeros(1, np.uint32)), None, F.ptr(np.array([2], np.uint32)), 1))
    med = np.zeros(1, dtype=np.dtype([("boundary_type", "<u4"), ("material", "<u4"), ("prim_id", "<u4"), ("_pad", "<u4"),
                                      ("density", "<f8"), ("p", "<f8", (6,)), ("rot_y_deg", "<f8"), ("offset", "<f8", (3,))]))
    iso = [i for i in range(len(cs.materials)) if int(cs.materials[i]["type"]) == F.MAT_ISOTROPIC][0]
    med[0] = (0, iso, 3, 0, 0.8, [0.0, 1.0, 0.0, 1.5, 0, 0], 0.0, [0, 0, 0])
    assert med.dtype.itemsize == 104
    F.check(lib.rtb_scene_set_media(flat.h, F.ptr(med), 1))
    flat.commit()
    assert flat.info()["n_prims"] == graph.info()["n_prims"] == 4
    rng = np.random.default_rng(2)
    o = rng.uniform((-3, 0.2, -3), (3, 2.8, 3), (4000, 3)).astype(np.float32)
    d = rng.normal(0, 1, (4000, 3)).astype(np.float32)
    tm = rng.random(4000).astype(np.float32)
    a = graph.trace_rays(o, d, tm)
    b = flat.trace_rays(o, d, tm)
    assert np.array_equal(a[0], b[0])
    np.testing.assert_allclose(a[1], b[1], rtol=1e-5, atol=1e-8)  # (a ray that went through the exact pass sees f32- vs f64-given geometry)
    oid, ot = osc.trace_rays(o.astype(np.float64), d.astype(np.float64), tm.astype(np.float64))
    assert (a[0] != oid).sum() <= 2
    cam = rtb.Camera.new((0, 1.5, 9), (0, 1.2, 0), (0, 1, 0), 40.0, 1.0, 0.0, 9.0)
    prm = rtb.make_params(32, 32, 256, background=(0.7, 0.8, 1.0), seed=5)
    acc_g, _ = graph.render(cam, prm)
    acc_f, _ = flat.render(cam, prm)
    np.testing.assert_allclose(acc_f, acc_g, rtol=2e-4, atol=2e-3)


def test_heightfield_closest_hit_full_1M_triangles(rtb, ctx):
    """C4 at full size (1 000 000 triangles): size-independent properties of a height field.
    (a) vertical rays hit exactly the triangle whose xz-projection contains them, at the interpolated height;
    (b) for slanted rays the reported (id, t) is a true intersection of that triangle and the ray stays above the
        surface before it (so it is the closest hit)."""
    from ray_tracer_archive_b200 import scenes, scene as S
    nx, nz = 1000, 500
    verts, tris = scenes.heightfield_mesh(nx, nz, seed=1)
    m = S.Lambertian.construct((0.73, 0.73, 0.73))
    dev = rtb.Scene(ctx, rtb.compile_scene(S.HittableList([S.TriangleMesh(verts, tris, m)])))
    assert dev.info()["n_triangles"] == 2 * nx * nz
    rng = np.random.default_rng(5)
    n = 200000
    x = rng.uniform(66, 489, n)
    z = rng.uniform(66, 489, n)
    o = np.stack([x, np.full(n, 500.0), z], 1).astype(np.float32)
    d = np.tile(np.array([[0, -1, 0]], np.float32), (n, 1))
    ids, ts, _ = dev.trace_rays(o, d)
    assert (ids != H.NONE).all()
    v = verts.astype(np.float64)
    tri = tris[ids]
    a, b, c = v[tri[:, 0]], v[tri[:, 1]], v[tri[:, 2]]

    def bary_xz(p, a, b, c):
        d00 = (b[:, 0] - a[:, 0]) * (c[:, 2] - a[:, 2]) - (c[:, 0] - a[:, 0]) * (b[:, 2] - a[:, 2])
        w1 = ((p[:, 0] - a[:, 0]) * (c[:, 2] - a[:, 2]) - (c[:, 0] - a[:, 0]) * (p[:, 2] - a[:, 2])) / d00
        w2 = ((b[:, 0] - a[:, 0]) * (p[:, 2] - a[:, 2]) - (p[:, 0] - a[:, 0]) * (b[:, 2] - a[:, 2])) / d00
        return w1, w2
    w1, w2 = bary_xz(o.astype(np.float64), a, b, c)
    eps = 1e-4
    assert ((w1 > -eps) & (w2 > -eps) & (w1 + w2 < 1 + eps)).all()
    y = a[:, 1] + w1 * (b[:, 1] - a[:, 1]) + w2 * (c[:, 1] - a[:, 1])
    np.testing.assert_allclose(500.0 - ts, y, atol=2e-3)
    # (b) slanted rays from above
    o2 = np.stack([rng.uniform(100, 450, n), np.full(n, 420.0), rng.uniform(100, 450, n)], 1).astype(np.float32)
    tgt = np.stack([rng.uniform(70, 485, n), np.full(n, 150.0), rng.uniform(70, 485, n)], 1)
    d2 = (tgt - o2).astype(np.float32)
    ids2, ts2, _ = dev.trace_rays(o2, d2)
    hit = ids2 != H.NONE
    assert hit.mean() > 0.95  # rays aimed near the footprint's border may leave it before reaching the surface
    tri = tris[ids2[hit]]
    a, b, c = v[tri[:, 0]], v[tri[:, 1]], v[tri[:, 2]]
    p = o2[hit].astype(np.float64) + ts2[hit, None].astype(np.float64) * d2[hit].astype(np.float64)
    nrm = np.cross(b - a, c - a)
    dist = np.abs(np.einsum("ij,ij->i", p - a, nrm)) / np.linalg.norm(nrm, axis=1)
    assert dist.max() < 5e-3
    w1, w2 = bary_xz(p, a, b, c)
    assert ((w1 > -1e-3) & (w2 > -1e-3) & (w1 + w2 < 1 + 1e-3)).all()
    # the ray is above the height field at 16 points before the hit: sample heights with vertical probe rays
    fr = rng.uniform(0.05, 0.95, (hit.sum(), 1))
    q = o2[hit].astype(np.float64) + fr * ts2[hit, None] * d2[hit].astype(np.float64)
    inside = (q[:, 0] > 66) & (q[:, 0] < 489) & (q[:, 2] > 66) & (q[:, 2] < 489)
    po = np.stack([q[:, 0], np.full(len(q), 500.0), q[:, 2]], 1).astype(np.float32)
    _, tq, _ = dev.trace_rays(po, np.tile(np.array([[0, -1, 0]], np.float32), (len(q), 1)))
    surf_y = 500.0 - tq
    assert (q[inside, 1] > surf_y[inside] - 5e-3).all()


# ------------------------------------------------------------------------------------------------------------- P2
def _p2(rtb, orc, ctx, cfg, W, Hh, spp=4096, rr=0, max_z=5.0, pool=0, oracle_bvh=False):
    dev, osc, _ = _scene_pair(rtb, orc, ctx, cfg)
    if oracle_bvh:  # candidate culling only; the per-primitive tests stay the reference's own
        osc.attach_bvh(dev)
    prm = rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=3, rr_start_depth=rr, pool_paths=pool)
    acc, st = dev.render(cfg.camera, prm)
    prm_o = rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=1234)  # independent sample set
    oacc, oseg, orej = osc.render(cfg.camera, prm_o)
    mean_rel, z, se_rel = H.compare_images(acc, oacc, spp, spp)
    seg_ratio = st["segments"] / oseg
    print(f"{cfg.name}: {W}x{Hh}x{spp} mean-lum err {100 * mean_rel:.3f}% (MC s.e. of that {100 * se_rel:.3f}%)  max z {z.max():.2f}  99.9% z {np.quantile(z, 0.999):.2f}"
          f"  segments gpu/oracle {seg_ratio:.4f}  rejected {st['rejected']}/{orej}")
    assert st["paths"] == W * Hh * spp
    assert se_rel < 0.004, "test too noisy to resolve the 1 % bar; raise spp or pixels"
    assert mean_rel <= 0.01
    assert z.max() <= max_z
    if rr == 0:
        assert abs(seg_ratio - 1) < 0.01  # same expected path length
    assert st["rejected"] <= 1e-5 * st["paths"] + orej * 2 + 8
    return acc, oacc, st


def test_p2_cornell_mixture_pdf(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    _p2(rtb, orc, ctx, scenes.config_cornell(), 48, 48)


def test_p2_random_spheres(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    _p2(rtb, orc, ctx, scenes.config_random_spheres(), 64, 36)


def test_p2_final_scene_media_textures_motion(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    _p2(rtb, orc, ctx, scenes.config_final_scene(boxes_per_side=8, n_small=120), 40, 40)


def test_p2_mesh_cornell(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    _p2(rtb, orc, ctx, scenes.config_mesh(nx=24, nz=12), 40, 24)


def test_p2_cornell_smoke_box_media(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_cornell()
    cfg.world, cfg.lights, cfg.name = scenes.cornell_smoke(), scenes.cornell_smoke_lights(), "cornell_smoke"
    _p2(rtb, orc, ctx, cfg, 40, 40)


def test_p2_checker_perlin_light(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_cornell()
    cfg.world, cfg.lights, cfg.name = scenes.simple_light(), None, "simple_light"
    cfg.camera = rtb.Camera.new((26, 3, 6), (0, 2, 0), (0, 1, 0), 20.0, 1.5, 0.0, 10.0)
    _p2(rtb, orc, ctx, cfg, 48, 32)
    cfg.world, cfg.name, cfg.background = scenes.two_spheres(), "two_spheres", (0.7, 0.8, 1.0)
    cfg.camera = rtb.Camera.new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.0, 10.0)
    _p2(rtb, orc, ctx, cfg, 48, 32)
    cfg.world, cfg.name = scenes.earth(), "earth"
    _p2(rtb, orc, ctx, cfg, 48, 32, spp=1024)


def test_p2_final_scene_FULL_scene(rtb, orc, ctx):
    """C3's actual scene — 20x20 ground boxes, 1000 rotated + translated spheres, both media, the moving sphere, Perlin and
    the reference's own earthmap.jpg texels — at 200x200, 1024 spp on both sides (independent sample sets)."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_final_scene()
    assert cfg.world is not None and scenes.earthmap().shape == (512, 1024, 3)
    _p2(rtb, orc, ctx, cfg, 200, 200, spp=1024, oracle_bvh=True)


def test_p2_mesh_FULL_1M_triangles(rtb, orc, ctx):
    """C4's actual scene (1 000 000 triangles + the Cornell walls and light) at 208x117, 1024 spp."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_mesh()
    _p2(rtb, orc, ctx, cfg, 208, 117, spp=1024, oracle_bvh=True)


@pytest.mark.parametrize("which", ["C1", "C2", "C3"])
def test_same_seed_paths_coincide(rtb, orc, ctx, which):
    """GPU and oracle share the Philox keying (key = pixel, sample; counter = block, bounce, seed), so with the SAME seed
    they trace the same paths: every draw is identical, closest-hit decisions are exact, and a path only departs from
    its f64 twin where f32 shading arithmetic moves a ray across a silhouette (then it is a different, equally valid
    sample).  Per-pixel means therefore agree far inside the Monte-Carlo error: this pins the device Philox stream and
    the draw order of every sampler against the oracle's (a swapped pair of uniforms anywhere would show up here as a
    full-size statistical difference)."""
    from ray_tracer_archive_b200 import scenes
    cfg = {"C1": scenes.config_random_spheres, "C2": scenes.config_cornell, "C3": scenes.config_final_scene}[which]()
    W, Hh = (cfg.width, cfg.height) if which != "C3" else (400, 400)
    spp = 64 if which != "C3" else 32
    dev, osc, _ = _scene_pair(rtb, orc, ctx, cfg)
    osc.attach_bvh(dev)
    acc, st = dev.render(cfg.camera, rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=21))
    oacc, oseg, _ = osc.render(cfg.camera, rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=21))
    acc2, st2 = dev.render(cfg.camera, rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=22))
    mg, vg = H.image_stats(acc, spp)
    mr, vr = H.image_stats(oacc, spp)
    m2, _ = H.image_stats(acc2, spp)
    sigma = np.sqrt((vg + vr) / spp) + 1e-6          # what two INDEPENDENT sample sets would differ by
    same = np.abs(mg - mr) / sigma
    indep = np.abs(m2 - mr) / sigma
    close = (np.abs(mg - mr) <= 2e-3 * (np.abs(mr) + 1e-3)).mean()
    mean_rel = abs(mg.mean() - mr.mean()) / mr.mean()
    print(f"{cfg.name}: {W}x{Hh}x{spp} same seed: {100 * close:.2f}% of pixels within 0.2%, median |d|/sigma {np.median(same):.4f} "
          f"(independent seeds: {np.median(indep):.3f}), mean-lum err {100 * mean_rel:.4f}%, segments gpu/oracle {st['segments'] / oseg:.5f}, "
          f"exact-pass rays {st['exact_rays']} of {st['segments']}")
    assert np.median(indep) > 0.3                      # sanity: the yardstick itself is of order one
    assert np.median(same) < 0.02                      # same seed: the typical pixel differs by < 2 % of one sigma
    assert close > (0.80 if which == "C1" else 0.60)   # most pixels agree to 0.2 % although each is a 64-sample mean
    assert mean_rel < 1e-3 and abs(st["segments"] / oseg - 1) < 2e-3


def test_rotated_image_textured_sphere(rtb, orc, ctx):
    """H8: uv are object-space (sphere.rs:32-37) — an image-textured sphere under Translate(RotateY) must show the map
    rotated with it (hittable.rs:147-176).  P1 + texel-exact first-hit colours + P2 against the oracle."""
    from ray_tracer_archive_b200 import scenes, scene as S
    img = scenes.earthmap()
    globe = S.Sphere.construct((0.0, 0.0, 0.0), 2.0, S.Lambertian.construct_texture(S.ImageTexture.construct(img, img.shape[1], img.shape[0])))
    world = S.HittableList([S.Translate.construct(S.RotateY.construct(globe, 75.0), (0.5, 0.2, -0.3)),
                            # (wrapped in a Translate like every rotated object of the reference's scenes: a BARE RotateY
                            # orients the normal against the object-space ray, hittable.rs:173 — DESIGN §9)
                            S.Translate.construct(S.RotateY.construct(S.Sphere.construct((4.5, 0.0, 0.0), 1.0, S.Lambertian.construct_texture(
                                S.ImageTexture.construct(scenes.synthetic_earth(64, 32), 64, 32))), -130.0), (0.0, 0.0, 0.0))])
    cfg = scenes.config_cornell()
    cfg.world, cfg.lights, cfg.name, cfg.background = world, None, "rotated earth", (0.7, 0.8, 1.0)
    cfg.camera = rtb.Camera.new((13, 2, 3), (0, 0, 0), (0, 1, 0), 24.0, 1.5, 0.0, 10.0)
    _p1(rtb, orc, ctx, cfg, 240, 160)
    # one segment deep: a pixel's value is background x texel, so the images compare texel by texel (same seed)
    dev, osc, _ = _scene_pair(rtb, orc, ctx, cfg)
    prm = rtb.make_params(96, 64, 16, 2, cfg.background, seed=5)
    acc, _ = dev.render(cfg.camera, prm)
    oacc, _, _ = osc.render(cfg.camera, prm)
    diff = np.abs(acc[..., :3] - oacc[..., :3]).max(axis=2) / 16
    ids, _, _ = dev.primary_hits(cfg.camera, 96, 64)
    for k in (0, 1):
        print(f"rotated sphere {k}: {(diff[ids == k] > 2e-3).mean():.3f} of its pixels differ by > 2e-3 (same seed, 16 spp)")
    # tex_value_slow on that sphere's own leaf: world-space normals -> the oracle's texel at the OBJECT-space uv
    F = rtb._ffi
    lib = orc.load()
    rng = np.random.default_rng(3)
    nrm = rng.normal(0, 1, (2000, 3))
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    nodes, prims = dev.export_bvh()
    leaf_of = {int(prims[0][1][2 * k]): k for k in range(len(prims[0][1]) // 2)}
    types = [int(t["type"]) for t in dev.cs.textures]
    out3 = np.zeros(3)
    for pid, ang in ((0, 75.0), (1, -130.0)):
        ti = [i for i, t in enumerate(types) if t == 3][pid]
        rows = np.concatenate([np.full((2000, 1), ti, np.uint32), np.zeros((2000, 3), np.uint32), nrm.view(np.uint32),
                               np.full((2000, 1), leaf_of[pid], np.uint32)], 1)
        got = ctx.kat(F.KAT_TEXTURE, rows, 3, scene=dev).view(np.float32)
        sn, cs = math.sin(math.radians(ang)), math.cos(math.radians(ang))
        bad = 0
        for i in range(2000):
            x, y, z = (float(v) for v in nrm[i])
            ox, oz = cs * x - sn * z, sn * x + cs * z          # world -> object, hittable.rs:150-156
            theta, phi = math.acos(max(-1.0, min(1.0, -y))), math.atan2(-oz, ox) + math.pi
            lib.orc_kat_texture(osc.h, ti, phi / (2 * math.pi), theta / math.pi, np.zeros(3).ctypes.data_as(C.c_void_p),
                                out3.ctypes.data_as(C.c_void_p))
            bad += not np.allclose(got[i], out3, atol=1e-6)
        print(f"sphere {pid} (RotateY {ang}): {bad} of 2000 texel lookups differ from the oracle's")
        assert bad <= 12   # normals within f32 rounding of a texel boundary
    assert (diff > 2e-3).mean() < 0.06, (diff > 2e-3).mean()   # texel boundaries under jitter (the map has 1-texel detail)
    assert oacc[..., :3].std() > 0.5                             # the map is really visible
    _p2(rtb, orc, ctx, cfg, 48, 32, spp=1024)


def test_front_face_under_rotate_y(rtb, orc, ctx):
    """H8: RotateY::hit calls set_face_normal with the OBJECT-space ray and the WORLD-space normal (hittable.rs:173), so
    under a RotateY `front_face` is q = dot(R^T d, n) < 0 — whatever side was hit — and a BARE RotateY also leaves the
    normal mis-oriented where q is false; a Translate outside re-orients the normal but keeps q as front_face
    (hittable.rs:82-83).  Dielectric (material.rs:127-132) and DiffuseLight (material.rs:183-189) read front_face: a
    strongly rotated emissive panel and rotated glass boxes, in every wrapper arrangement, against the literal oracle."""
    from ray_tracer_archive_b200 import scenes, scene as S
    white = S.Lambertian.construct((0.73, 0.73, 0.73))
    glass = S.Dielectric.construct(1.5)
    lamp = S.DiffuseLight.construct_color((6.0, 5.0, 4.0))
    lamp2 = S.DiffuseLight.construct_color((3.0, 5.0, 7.0))
    world = S.HittableList([
        S.XzRect.construct(-30.0, 30.0, -30.0, 30.0, 0.0, white),                                   # floor
        # emissive panels: Translate(RotateY(FlipFace(rect))), bare RotateY(rect), FlipFace(Translate(RotateY(rect)))
        S.Translate.construct(S.RotateY.construct(S.FlipFace.construct(S.XzRect.construct(-3.0, 3.0, -2.0, 2.0, 0.0, lamp)), 50.0), (-4.0, 7.0, 0.0)),
        S.RotateY.construct(S.XyRect.construct(2.0, 7.0, 1.0, 5.0, -6.0, lamp2), -40.0),
        S.FlipFace.construct(S.Translate.construct(S.RotateY.construct(S.YzRect.construct(1.0, 4.0, -2.0, 2.0, 0.0, lamp), 35.0), (8.0, 0.0, 3.0))),
        # glass boxes: Translate(RotateY(box)) as in the reference's scenes, and a bare RotateY(box)
        S.Translate.construct(S.RotateY.construct(S.Box.construct((0.0, 0.0, 0.0), (2.5, 3.0, 2.5), glass), 30.0), (-2.0, 0.01, 1.0)),
        S.RotateY.construct(S.Box.construct((2.0, 0.01, -1.0), (4.0, 2.5, 1.0), glass), -25.0),
        # a glass sphere under RotateY(Translate(.)) — a Translate INSIDE the rotation
        S.RotateY.construct(S.Translate.construct(S.Sphere.construct((0.0, 0.0, 0.0), 1.2, glass), (-5.0, 1.3, 4.0)), 20.0),
    ])
    cfg = scenes.config_cornell()
    cfg.world, cfg.lights, cfg.name, cfg.background = world, None, "front_face under RotateY", (0.02, 0.02, 0.03)
    cfg.camera = rtb.Camera.new((3.0, 9.0, 22.0), (0.0, 2.5, 0.0), (0, 1, 0), 38.0, 1.5, 0.0, 10.0)
    _p1(rtb, orc, ctx, cfg, 240, 160)
    _p2(rtb, orc, ctx, cfg, 60, 40, spp=4096)


def test_obj_mesh_parity(rtb, orc, ctx, tmp_path):
    """f2: a mesh read by the OBJ loader (io.load_obj -> rtb_scene_set_mesh), transformed, inside the Cornell walls:
    P1 on identical rays + P2 against the oracle."""
    from ray_tracer_archive_b200 import scenes, io, scene as S
    # an OBJ file: icosphere-like blob written as text (quads + triangles + negative indices + v/vt/vn forms)
    rng = np.random.default_rng(8)
    nu, nv = 24, 12
    lines = ["# test mesh"]
    for j in range(nv + 1):
        for i in range(nu):
            th, ph = math.pi * j / nv, 2 * math.pi * i / nu
            r = 1.0 + 0.15 * math.sin(3 * ph) * math.sin(2 * th)
            lines.append(f"v {r * math.sin(th) * math.cos(ph):.6f} {r * math.cos(th):.6f} {r * math.sin(th) * math.sin(ph):.6f}")
    for j in range(nv):
        for i in range(nu):
            a, b = j * nu + i + 1, j * nu + (i + 1) % nu + 1
            c, d = a + nu, b + nu
            lines.append(f"f {a}/1/1 {b}/1/1 {d}/1/1 {c}/1/1" if (i + j) % 2 else f"f {a} {b} {d}\nf {a} {d} {c}")
    path = tmp_path / "blob.obj"
    path.write_text("\n".join(lines))
    mesh = io.load_obj(str(path), S.Lambertian.construct((0.6, 0.7, 0.3)), scale=110.0, offset=(278.0, 180.0, 278.0))
    cfg = scenes.config_cornell()
    walls = scenes.cornell_box()
    world = S.HittableList([o for o in walls.objects[:6]] + [mesh])
    cfg.world, cfg.name = world, "OBJ mesh in the Cornell box"
    cfg.lights = S.HittableList([S.XzRect.construct(213.0, 343.0, 227.0, 332.0, 554.0, S.Lambertian.construct((0, 0, 0)))])
    dev, _, _ = _scene_pair(rtb, orc, ctx, cfg)
    assert dev.info()["n_triangles"] == 2 * nu * nv
    ids, oid = _p1(rtb, orc, ctx, cfg, 300, 300)
    assert (oid >= 6).mean() > 0.05
    _p2(rtb, orc, ctx, cfg, 40, 40, spp=2048)


def _block_stats(acc, spp, bs):
    """Mean luminance and its standard error over bs x bs pixel blocks (accum = sum R,G,B,Y^2 per pixel)."""
    a = np.asarray(acc, dtype=np.float64)
    Hh, W = a.shape[0] // bs * bs, a.shape[1] // bs * bs
    a = a[:Hh, :W].reshape(Hh // bs, bs, W // bs, bs, 4)
    n = bs * bs * spp
    sum_y = H.luminance(a[..., :3]).sum(axis=(1, 3))
    sum_y2 = a[..., 3].sum(axis=(1, 3))
    mean = sum_y / n
    var = np.maximum(sum_y2 / n - mean * mean, 0.0)
    return mean, var / n


@pytest.mark.parametrize("which", ["C1", "C2"])
def test_p2_full_resolution_low_spp(rtb, orc, ctx, which):
    """Full config resolution (1200x675 is not a multiple of the 8x4 pixel tiles; every pixel must receive exactly its
    samples): 16 spp on both sides, compared on 15x15-pixel blocks (3600 samples per block)."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_random_spheres() if which == "C1" else scenes.config_cornell()
    dev, osc, _ = _scene_pair(rtb, orc, ctx, cfg)
    osc.attach_bvh(dev)
    spp = 16
    acc, st = dev.render(cfg.camera, rtb.make_params(cfg.width, cfg.height, spp, cfg.max_depth, cfg.background, seed=11))
    oacc, oseg, _ = osc.render(cfg.camera, rtb.make_params(cfg.width, cfg.height, spp, cfg.max_depth, cfg.background, seed=12))
    assert st["paths"] == cfg.width * cfg.height * spp
    # every pixel got all its samples: background-only pixels (C1's sky) sum to exactly spp * background
    if which == "C1":
        top = acc[0, :, :3] / spp
        np.testing.assert_allclose(top, np.broadcast_to(np.float32(cfg.background), top.shape), rtol=2e-6)
    mg, vg = _block_stats(acc, spp, 15)
    mr, vr = _block_stats(oacc, spp, 15)
    z = np.abs(mg - mr) / np.sqrt(vg + vr + (3e-4 * mr + 1e-6) ** 2)
    mean_rel = abs(mg.mean() - mr.mean()) / mr.mean()
    print(f"{cfg.name}: {cfg.width}x{cfg.height}x{spp} block z max {z.max():.2f}, mean-lum err {100 * mean_rel:.3f}%, "
          f"segments gpu/oracle {st['segments'] / oseg:.4f}")
    assert mean_rel <= 0.01 and z.max() <= 5.0 and abs(st["segments"] / oseg - 1) < 0.01


def test_p2_russian_roulette_is_unbiased(rtb, orc, ctx):
    """RR is not in the reference; with it on (GPU) the image must still match the RR-free oracle."""
    from ray_tracer_archive_b200 import scenes
    _, _, st = _p2(rtb, orc, ctx, scenes.config_cornell(), 48, 48, rr=3)
    assert st["segments"] < 48 * 48 * 4096 * 12


# ------------------------------------------------------------------------------------------------------ P3 / D1 / misc
def test_white_furnace_exact(rtb, ctx):
    from ray_tracer_archive_b200 import scene as S
    world = S.HittableList([S.Sphere((0, 0, 0), 1.0, S.Lambertian.construct((0.6, 0.6, 0.6)))])
    dev = rtb.Scene(ctx, rtb.compile_scene(world))
    cam = rtb.Camera.new((0, 0, 4), (0, 0, 0), (0, 1, 0), 40.0, 1.0, 0.0, 4.0)
    acc, st = dev.render(cam, rtb.make_params(64, 64, 16, background=(1, 1, 1)))
    mean = acc[..., :3] / 16
    assert abs(mean[32, 32, 0] - 0.6) < 1e-5 and abs(mean[0, 0, 0] - 1.0) < 1e-6
    assert np.all((mean > 0.6 - 1e-5) & (mean < 1.0 + 1e-5))
    assert st["rejected"] == 0 and st["segments"] >= 64 * 64 * 16


def test_d1_sample_split_matches_single_call(rtb, ctx):
    """The multi-GPU partition on one GPU: samples [0,n/2) and [n/2,n) accumulated by two calls == one call of n
    samples (same global sample indices -> same sample set; only the f32 summation order differs)."""
    from ray_tracer_archive_b200 import scenes, parallel
    cfg = scenes.config_cornell()
    dev = rtb.Scene(ctx, rtb.compile_scene(cfg.world, cfg.lights))
    W = Hh = 64
    spp = 64
    one, st1 = dev.render(cfg.camera, rtb.make_params(W, Hh, spp, seed=9))
    f0, c0 = parallel.rank_sample_range(spp, 0, 2)
    f1, c1 = parallel.rank_sample_range(spp, 1, 2)
    assert (f0, c0, f1, c1) == (0, 32, 32, 32)
    dev.render(cfg.camera, rtb.make_params(W, Hh, c0, seed=9, sample_offset=f0, total_spp=spp), readback=False)
    two, st2 = dev.render(cfg.camera, rtb.make_params(W, Hh, c1, seed=9, sample_offset=f1, total_spp=spp,
                                                      flags=rtb._ffi.RENDER_ACCUMULATE))
    np.testing.assert_allclose(two, one, rtol=2e-4, atol=1e-4)
    # scheduling independence: a different pool size gives the same sample set
    three, _ = dev.render(cfg.camera, rtb.make_params(W, Hh, spp, seed=9, pool_paths=5000))
    np.testing.assert_allclose(three, one, rtol=2e-4, atol=1e-4)
    assert st1["segments"] > 0


@pytest.mark.parametrize("pool", [1024, 1025, 1279, 4097, 60000, 1 << 20])
def test_slot_stable_pool_any_size_renders_every_path_once(rtb, ctx, pool):
    """The pool is walked in 256-slot chunks and path numbers come from per-chunk cursors: ragged last chunks, pools
    smaller / larger than the path count and pools that are not a multiple of anything must all start every (pixel,
    sample) exactly once — the image equals the default-pool image up to f32 summation order and the segment count is
    identical (paths are deterministic functions of (pixel, sample, seed))."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_random_spheres()
    dev = rtb.Scene(ctx, rtb.compile_scene(cfg.world, cfg.lights))
    W, Hh, spp = 96, 54, 24  # 124 416 paths
    ref, st0 = dev.render(cfg.camera, rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=4))
    acc, st = dev.render(cfg.camera, rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=4, pool_paths=pool))
    assert st["paths"] == W * Hh * spp and st["segments"] == st0["segments"] and st["rejected"] == 0
    np.testing.assert_allclose(acc, ref, rtol=3e-4, atol=1e-4)
    # the sum of the weights deposited per pixel is the sample count: with a constant background and a white
    # furnace-like check we cannot count deposits directly, so check the luminance moments instead
    assert abs(acc[..., :3].sum() / ref[..., :3].sum() - 1.0) < 1e-5


def test_checkpoint_resume_adds_samples(rtb, ctx, tmp_path):
    """Resume = add more spp: 16 spp, checkpoint to disk, load into a NEW device buffer, 16 more spp with
    sample_offset 16 + RTB_RENDER_ACCUMULATE  ==  one 32-spp render (same global sample indices)."""
    import torch
    from ray_tracer_archive_b200 import scenes, io
    cfg = scenes.config_cornell()
    dev = rtb.Scene(ctx, rtb.compile_scene(cfg.world, cfg.lights))
    W = Hh = 48
    ref, _ = dev.render(cfg.camera, rtb.make_params(W, Hh, 32, seed=6))
    a = torch.zeros((Hh, W, 4), dtype=torch.float32, device="cuda")
    dev.render_device(cfg.camera, rtb.make_params(W, Hh, 16, seed=6, total_spp=32), a.data_ptr())
    torch.cuda.synchronize()
    io.save_checkpoint(str(tmp_path / "half"), a.cpu().numpy(), spp_done=16, seed=6)
    acc, done = io.load_checkpoint(str(tmp_path / "half"), width=W, height=Hh, seed=6)
    b = torch.from_numpy(acc).cuda()
    dev.render_device(cfg.camera, rtb.make_params(W, Hh, 32 - done, seed=6, sample_offset=done, total_spp=32,
                                                  flags=rtb._ffi.RENDER_ACCUMULATE), b.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_allclose(b.cpu().numpy(), ref, rtol=2e-4, atol=1e-4)


def test_finalize_rgb8_matches_write_color(rtb, orc, ctx):
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_cornell()
    dev = rtb.Scene(ctx, rtb.compile_scene(cfg.world, cfg.lights))
    acc, _ = dev.render(cfg.camera, rtb.make_params(32, 32, 64))
    rgb = dev.finalize_rgb8(32, 32, 64)
    lib = orc.load()
    out = np.zeros(3, np.uint8)
    worst = 0
    for yx in [(0, 0), (5, 7), (16, 16), (31, 31), (2, 16)]:
        s = np.ascontiguousarray(acc[yx][:3], dtype=np.float64)
        lib.orc_write_color(s.ctypes.data_as(C.c_void_p), 64, out.ctypes.data_as(C.c_void_p))
        worst = max(worst, int(np.abs(out.astype(int) - rgb[yx].astype(int)).max()))
    assert worst <= 1  # f32 sqrt vs f64 sqrt may straddle an integer boundary


def test_abi_state_errors(rtb, ctx):
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_cornell()
    s = rtb.Scene(ctx)
    with pytest.raises(rtb.RtbError) as e:
        s.render(cfg.camera, rtb.make_params(8, 8, 1))
    assert e.value.code == -4
    s.set_compiled(rtb.compile_scene(cfg.world, cfg.lights))
    s.commit()
    with pytest.raises(rtb.RtbError):
        s.render(cfg.camera, rtb.make_params(1, 8, 1))     # W-1 == 0 (main.rs:752)
    with pytest.raises(rtb.RtbError):
        s.render(cfg.camera, rtb.make_params(8, 8, 1, max_depth=0))
    acc, st = s.render(cfg.camera, rtb.make_params(8, 8, 2))
    assert st["paths"] == 128 and np.isfinite(acc).all()
    for bad in (rtb.Camera.new((1, 2, 3), (1, 2, 3), (0, 1, 0), 40.0, 1.0, 0.0, 10.0),          # lookfrom == lookat
                rtb.Camera.new((0, 0, 0), (0, 5, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0),          # vup along the view direction
                rtb.Camera.new((0, 0, 0), (0, 0, -1), (0, 1, 0), float("nan"), 1.0, 0.0, 10.0)):
        with pytest.raises(rtb.RtbError) as e:
            s.render(bad, rtb.make_params(8, 8, 1))
        assert e.value.code == -1
    info = ctx.device_info()
    assert info["sm_count"] > 0
