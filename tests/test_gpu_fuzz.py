"""GPU tier (-m gpu): RANDOM scene graphs — every primitive, material, texture and wrapper the reference's constructors
offer, in arrangements its own scenes never use — through the C ABI against the literal f64 oracle.

The fixed scenes (tests/test_gpu_parity.py) only cover the wrapper chains of `main.rs` (a Translate around a RotateY around a
box).  Here a seeded generator nests Translate / RotateY / FlipFace / ConstantMedium / HittableList / BVHNode at random
around spheres, moving spheres, the three rects and boxes, with Lambertian (solid / checker / noise / image), Metal,
Dielectric and DiffuseLight, and each scene has to pass
  P1  identical primary rays: primitive ids equal on every pixel, |t - t_ref| <= 1e-5 t_ref;
  P2  1024-spp images: mean luminance within 1 %, no pixel beyond 5 sigma, equal expected path length.
(What the flattener documents as not reproduced — two RotateY in one chain — is not generated.)"""
import numpy as np
import pytest

import helpers as H
from test_gpu_parity import _p1, _p2, _scene_pair

pytestmark = pytest.mark.gpu


def _random_scene(seed, S, scenes, dim=False, extended=False, n_objs=None):
    rng = np.random.default_rng(1000 + seed)
    u = lambda a, b: float(rng.uniform(a, b))
    col = lambda lo=0.2, hi=0.9: (u(lo, hi), u(lo, hi), u(lo, hi))

    def material(allow_light=True):
        k = rng.integers(0, 9 if allow_light else 8)
        if k <= 1:
            return S.Lambertian.construct(col())
        if k == 2:
            chk = S.CheckerTexture.construct_color(col(0.1, 0.4), col(0.6, 0.95))
            if extended and rng.random() < 0.5:  # Arc<dyn Texture> children (texture.rs:41-51): noise, image, another checker
                chk = S.CheckerTexture.construct(S.NoiseTexture.construct(u(1, 4), np.random.default_rng(int(rng.integers(1 << 30)))),
                                                 S.CheckerTexture.construct(S.SolidColor(col()), chk))
            return S.Lambertian.construct_texture(chk)
        if k == 3:
            return S.Lambertian.construct_texture(S.NoiseTexture.construct(u(0.5, 4.0), np.random.default_rng(int(rng.integers(1 << 30)))))
        if k == 4:
            img = scenes.synthetic_earth(32, 16)
            return S.Lambertian.construct_texture(S.ImageTexture.construct(img, 32, 16))
        if k == 5:
            return S.Metal.construct(col(0.5, 0.95), u(0.0, 0.6))
        if k in (6, 7):
            return S.Dielectric.construct(u(1.2, 1.8))
        return S.DiffuseLight.construct_color((u(2, 6), u(2, 6), u(2, 6)))

    def primitive():
        k = rng.integers(0, 8 if extended else 6)
        m = material()
        c = (u(-5, 5), u(0.5, 4), u(-5, 5))
        if k == 0:
            return S.Sphere.construct(c, u(0.4, 1.6), m)
        if k == 1:
            return S.MovingSphere.construct(c, (c[0] + u(-0.6, 0.6), c[1] + u(0, 0.8), c[2]), 0.0, 1.0, u(0.4, 1.2), m)
        if k == 2:
            return S.XyRect.construct(c[0], c[0] + u(1, 3), c[1], c[1] + u(1, 3), c[2], m)
        if k == 3:
            return S.XzRect.construct(c[0], c[0] + u(1, 3), c[2], c[2] + u(1, 3), c[1], m)
        if k == 4:
            return S.YzRect.construct(c[1], c[1] + u(1, 3), c[2], c[2] + u(1, 3), c[0], m)
        if k == 5:
            return S.Box.construct(c, (c[0] + u(0.8, 2.5), c[1] + u(0.8, 2.5), c[2] + u(0.8, 2.5)), m)
        # the primitives that have no reference counterpart: general quad, triangle
        e1, e2 = (u(0.8, 2.5), u(-0.5, 0.5), u(-0.5, 0.5)), (u(-0.5, 0.5), u(0.8, 2.5), u(-0.8, 0.8))
        if k == 6:
            return S.Quad(c, e1, e2, m)
        return S.Triangle(c, tuple(a + b for a, b in zip(c, e1)), tuple(a + b for a, b in zip(c, e2)), m)

    def wrapped(obj, rotations_left=1, depth=0):
        """0-3 random wrappers around obj (at most one RotateY per chain)."""
        for _ in range(int(rng.integers(0, 4))):
            k = rng.integers(0, 4)
            if k == 0:
                obj = S.Translate.construct(obj, (u(-2, 2), u(0, 1.5), u(-2, 2)))
            elif k == 1 and rotations_left:
                obj = S.RotateY.construct(obj, u(-80, 80))
                rotations_left -= 1
            elif k == 2:
                obj = S.FlipFace.construct(obj)
            elif k == 3 and depth == 0 and rng.random() < 0.5:
                # a small list (or the reference's BVHNode over it) as one object, wrapped as a whole
                items = [wrapped(primitive(), 0, depth + 1) for _ in range(int(rng.integers(2, 5)))]
                lst = S.HittableList(items + [obj])
                obj = S.BVHNode.construct2(lst, 0.0, 1.0) if rng.random() < 0.5 else lst
        return obj

    objs = [S.XzRect.construct(-40.0, 40.0, -40.0, 40.0, 0.0, S.Lambertian.construct(col(0.4, 0.8)))]  # floor
    import copy
    for _ in range(n_objs if n_objs else int(rng.integers(6, 14))):
        prim = primitive()
        objs.append(wrapped(prim))
        if extended and rng.random() < 0.15 and not isinstance(prim, S.MovingSphere):
            # an exact duplicate with another material, and a coplanar / concentric sibling: equal-t ties go to the LATER
            # object of the list (hittable_list.rs:44-47 with sphere.rs:52 / aarect.rs:33), whatever the traversal order
            twin = copy.copy(prim)
            twin.mat = material()
            objs.append(twin)
            if isinstance(prim, (S.XyRect, S.XzRect, S.YzRect)):
                sib = copy.copy(prim)
                sib.a0, sib.a1, sib.mat = prim.a0 + 0.4 * (prim.a1 - prim.a0), prim.a1 + 0.7, material()
                objs.append(sib)
    if rng.random() < 0.6:  # a participating medium bounded by a (possibly transformed) sphere or box
        b = S.Sphere.construct((u(-3, 3), u(1.5, 3), u(-3, 3)), u(1.0, 2.0), S.Dielectric.construct(1.5)) if rng.random() < 0.5 else \
            S.Box.construct((-1.0, 0.2, -1.0), (1.5, 2.5, 1.2), S.Lambertian.construct(col()))
        if rng.random() < 0.5:
            b = S.Translate.construct(S.RotateY.construct(b, u(-40, 40)), (u(-2, 2), 0.0, u(-2, 2)))
        objs.append(S.ConstantMedium.construct_color(b, u(0.1, 0.6), col()))
    lights = None
    if dim or rng.random() < 0.5:  # light sampling (mixture pdf) on an untransformed panel, as in main.rs:669-686
        lamp = S.DiffuseLight.construct_color((7.0, 7.0, 7.0))
        objs.append(S.FlipFace.construct(S.XzRect.construct(-2.0, 2.0, -2.0, 2.0, 9.0, lamp)))
        lights = S.HittableList.new()
        lights.add(S.XzRect.construct(-2.0, 2.0, -2.0, 2.0, 9.0, lamp))
        if rng.random() < 0.5:  # a second, spherical entry in the light list (sphere.rs:75-90), like main.rs:682-686
            ball = S.Sphere.construct((u(-4, 4), u(4, 6), u(-4, 4)), u(0.5, 1.0), lamp if rng.random() < 0.5 else S.Dielectric.construct(1.5))
            objs.append(ball)
            lights.add(ball)
    rng.shuffle(objs)
    return S.HittableList(list(objs)), lights


import os
_SEEDS = range(int(os.environ.get("RTB_FUZZ_FIRST", "0")), int(os.environ.get("RTB_FUZZ_LAST", "24")))  # (one-off sweeps: more seeds)


@pytest.mark.parametrize("seed", _SEEDS)
def test_random_scene_graph(rtb, orc, ctx, seed):
    from ray_tracer_archive_b200 import scenes, scene as S
    dim = seed % 24 >= 12  # second half: almost all light comes from emitters (front_face-sensitive) instead of the sky
    extended = seed % 2 == 1  # odd seeds: + general quads and triangles, a thin lens, Russian roulette
    world, lights = _random_scene(seed, S, scenes, dim, extended)
    cfg = scenes.config_cornell()
    cfg.world, cfg.lights, cfg.name = world, lights, f"random scene graph {seed}"
    cfg.background = (0.03, 0.04, 0.06) if dim else (0.55, 0.65, 0.85)
    cfg.camera = rtb.Camera.new((2.0, 7.0, 19.0), (0.0, 2.0, 0.0), (0, 1, 0), 40.0, 1.5, 0.3 if extended else 0.0, 19.0, 0.0, 1.0)
    _p1(rtb, orc, ctx, cfg, 192, 128)
    _random_rays(rtb, orc, ctx, cfg, seed)
    _p2(rtb, orc, ctx, cfg, 48, 32, spp=4096 if dim else 1024, rr=4 if extended and seed % 4 == 1 else 0)
    _same_seed(rtb, orc, ctx, cfg)


def _same_seed(rtb, orc, ctx, cfg, W=96, Hh=64, spp=32):
    """GPU and oracle trace the SAME paths from the same seed (shared Philox keying): the per-pixel means of a 32-sample
    render agree far inside the Monte-Carlo error — in every combination of material, texture, medium and light list the
    generator comes up with.  A sampler that draws its uniforms in a different order anywhere shows up here."""
    dev, osc, _ = _scene_pair(rtb, orc, ctx, cfg)
    acc, st = dev.render(cfg.camera, rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=77))
    oacc, oseg, _ = osc.render(cfg.camera, rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=77))
    acc2, _ = dev.render(cfg.camera, rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=78))
    mg, vg = H.image_stats(acc, spp)
    mr, vr = H.image_stats(oacc, spp)
    m2, _ = H.image_stats(acc2, spp)
    sigma = np.sqrt((vg + vr) / spp) + 1e-6
    same, indep = np.abs(mg - mr) / sigma, np.abs(m2 - mr) / sigma
    close = (np.abs(mg - mr) <= 2e-3 * (np.abs(mr) + 1e-3)).mean()
    print(f"{cfg.name}: same seed {W}x{Hh}x{spp}: {100 * close:.1f}% of pixels within 0.2%, mean |d|/sigma {same.mean():.3f} "
          f"(independent seeds: {indep.mean():.3f}), segments gpu/oracle {st['segments'] / oseg:.5f}")
    assert same.mean() < 0.25 * indep.mean()
    assert close > 0.5
    assert abs(st["segments"] / oseg - 1) < 5e-3


def _random_rays(rtb, orc, ctx, cfg, seed, n=20000):
    """Closest hit of rays no camera produces: origins anywhere in the scene's volume (also inside boxes, spheres and
    media), directions uniform on the sphere, un-normalised lengths 0.01-100, random times; and bounce-like rays that
    start ON a surface (the oracle's hit point of the first set).  Ids equal; the second set tolerates the documented
    self-intersection band of rays skimming their own surface (t against t_min is decided in f32)."""
    from test_host_bvh import _media_ids
    dev, osc, cs = _scene_pair(rtb, orc, ctx, cfg)
    rng = np.random.default_rng(7000 + seed)
    o = rng.uniform([-9, 0.05, -9], [9, 9, 9], (n, 3)).astype(np.float32)
    d = rng.normal(0, 1, (n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True) * np.exp(rng.uniform(np.log(0.01), np.log(100.0), (n, 1)))).astype(np.float32)
    tm = rng.random(n).astype(np.float32)
    mid = _media_ids(cs) if dev.info()["n_media"] else []
    tri_ids = dev.export_bvh()[1][3][1].reshape(-1, 2)[:, 0]
    for label, tol in (("free-space", 0), ("on-surface", 4e-4)):
        ids, ts, _ = dev.trace_rays(o, d, tm)
        oid, ot = osc.trace_rays(o.astype(np.float64), d.astype(np.float64), tm.astype(np.float64))
        surf = ~(np.isin(oid, mid) | np.isin(ids, mid))
        mism = (ids != oid) & surf
        hit = surf & ~mism & (oid != H.NONE)
        rel = np.abs(ts[hit].astype(np.float64) - ot[hit]) / ot[hit]
        # a triangle under Translate / RotateY has the transform baked into its f32 vertices: its distances carry
        # the rounding of the world-space vertex, an ABSOLUTE 2^-24 of the coordinate, which is more than 1e-5 of a hit 0.01 away
        dist = ot[hit] * np.linalg.norm(d[hit].astype(np.float64), axis=1)
        floor = np.where(np.isin(oid[hit], tri_ids), 3e-6 / np.maximum(dist, 1e-30), 0.0)
        print(f"{cfg.name}: {label} rays: {int(mism.sum())} id mismatches of {n}, max t err {rel.max() if hit.any() else 0:.2e}")
        assert mism.sum() <= tol * n, (label, np.argwhere(mism)[:5].ravel())
        assert not hit.any() or np.quantile(np.maximum(rel - floor, 0.0), 0.999) <= 1e-5
        # next: from the oracle's hit points, new random directions
        p = (o[hit].astype(np.float64) + ot[hit, None] * d[hit].astype(np.float64)).astype(np.float32)
        nd = rng.normal(0, 1, p.shape)
        o, d = p, (nd / np.linalg.norm(nd, axis=1, keepdims=True)).astype(np.float32)
        tm = rng.random(len(p)).astype(np.float32)
        n = len(p)
        if n == 0:
            break


def test_many_image_and_noise_textures(rtb, orc, ctx):
    """Any number of image / perlin tables per scene (the reference allocates one per texture, texture.rs:79-87,106-116):
    24 spheres, each with its own image or noise texture."""
    from ray_tracer_archive_b200 import scenes, scene as S
    rng = np.random.default_rng(5)
    objs = [S.XzRect.construct(-40.0, 40.0, -40.0, 40.0, 0.0, S.Lambertian.construct((0.5, 0.5, 0.5)))]
    for k in range(24):
        if k % 2:
            img = (rng.integers(0, 256, (8, 16, 3))).astype(np.uint8)
            tex = S.ImageTexture.construct(img, 16, 8)
        else:
            tex = S.NoiseTexture.construct(float(rng.uniform(1, 5)), np.random.default_rng(100 + k))
        objs.append(S.Sphere.construct((-8.0 + 2.8 * (k % 6), 1.0 + 2.2 * (k // 6), float(rng.uniform(-2, 2))), 1.0, S.Lambertian.construct_texture(tex)))
    cfg = scenes.config_cornell()
    cfg.world, cfg.lights, cfg.name, cfg.background = S.HittableList(objs), None, "24 textures", (0.6, 0.7, 0.9)
    cfg.camera = rtb.Camera.new((0.0, 5.0, 24.0), (-1.0, 4.0, 0.0), (0, 1, 0), 40.0, 1.5, 0.0, 10.0)
    _p1(rtb, orc, ctx, cfg, 192, 128)
    acc, oacc, _ = _p2(rtb, orc, ctx, cfg, 48, 32, spp=1024)
    assert oacc[..., :3].std() > 0.05


@pytest.mark.parametrize("seed,n_objs", [(200, 300), (201, 1500), (202, 4000)])
def test_random_scene_graph_large(rtb, orc, ctx, seed, n_objs):
    """The same generator with hundreds to thousands of objects (lists, boxes and wrappers included: up to ~4x as many
    primitives): the BVH8 builder's typed trees, their opened join, the global-primitive rules and the three extend
    schedulers' thresholds on scenes nobody tuned them for.  P1 on camera rays and on random rays."""
    from ray_tracer_archive_b200 import scenes, scene as S
    world, lights = _random_scene(seed, S, scenes, False, True, n_objs)
    cfg = scenes.config_cornell()
    cfg.world, cfg.lights, cfg.name, cfg.background = world, lights, f"random scene graph {seed} ({n_objs} objects)", (0.55, 0.65, 0.85)
    cfg.camera = rtb.Camera.new((2.0, 7.0, 19.0), (0.0, 2.0, 0.0), (0, 1, 0), 40.0, 1.5, 0.0, 19.0, 0.0, 1.0)
    _p1(rtb, orc, ctx, cfg, 192, 128)
    _random_rays(rtb, orc, ctx, cfg, seed, n=8000)
    _same_seed(rtb, orc, ctx, cfg, 48, 32, 16)
