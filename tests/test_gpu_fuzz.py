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


def _random_scene(seed, S, scenes, dim=False, extended=False):
    rng = np.random.default_rng(1000 + seed)
    u = lambda a, b: float(rng.uniform(a, b))
    col = lambda lo=0.2, hi=0.9: (u(lo, hi), u(lo, hi), u(lo, hi))

    def material(allow_light=True):
        k = rng.integers(0, 9 if allow_light else 8)
        if k <= 1:
            return S.Lambertian.construct(col())
        if k == 2:
            return S.Lambertian.construct_texture(S.CheckerTexture.construct_color(col(0.1, 0.4), col(0.6, 0.95)))
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
    for _ in range(int(rng.integers(6, 14))):
        objs.append(wrapped(primitive()))
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
    for label, tol in (("free-space", 0), ("on-surface", 4e-4)):
        ids, ts, _ = dev.trace_rays(o, d, tm)
        oid, ot = osc.trace_rays(o.astype(np.float64), d.astype(np.float64), tm.astype(np.float64))
        surf = ~(np.isin(oid, mid) | np.isin(ids, mid))
        mism = (ids != oid) & surf
        hit = surf & ~mism & (oid != H.NONE)
        rel = np.abs(ts[hit].astype(np.float64) - ot[hit]) / ot[hit]
        print(f"{cfg.name}: {label} rays: {int(mism.sum())} id mismatches of {n}, max t err {rel.max() if hit.any() else 0:.2e}")
        assert mism.sum() <= tol * n, (label, np.argwhere(mism)[:5].ravel())
        assert not hit.any() or np.quantile(rel, 0.999) <= 1e-5
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
