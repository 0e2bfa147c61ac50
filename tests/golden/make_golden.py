"""Generates tests/golden/kat_vectors.json: known-answer vectors for the reference's hot-path formulas, computed by
an independent numpy/python restatement (NOT by the C++ oracle) of the cited reference lines.  The reference itself
ships no vectors (cargo test runs zero tests) and cannot be run here (no rustc), so these pin the oracle against a
second, separately written restatement plus closed-form identities.   Run:  python tests/golden/make_golden.py
"""
import json
import math
import os

import numpy as np

rng = np.random.default_rng(20261018)
out = {}


def sphere_hit(c, r, o, d, tmin, tmax):  # sphere.rs:41-65, 32-37 ; hittable.rs:41-48
    oc = o - c
    a = d @ d
    hb = oc @ d
    cc = oc @ oc - r * r
    det = hb * hb - a * cc
    if det < 0:
        return None
    sq = math.sqrt(det)
    root = (-hb - sq) / a
    if root < tmin or tmax < root:
        root = (-hb + sq) / a
        if root < tmin or tmax < root:
            return None
    p = o + root * d
    n = (p - c) / r
    front = (d @ n) < 0
    theta, phi = math.acos(-n[1]), math.atan2(-n[2], n[0]) + math.pi
    nn = n if front else -n
    return [root, *nn, phi / (2 * math.pi), theta / math.pi, 1.0 if front else 0.0]


cases = []
for _ in range(40):
    c, r = rng.uniform(-5, 5, 3), rng.uniform(0.2, 3)
    o = rng.uniform(-10, 10, 3)
    d = (c + rng.normal(0, r * 0.7, 3)) - o
    d *= rng.uniform(0.3, 3)
    tmin, tmax = 0.001, float("inf") if rng.random() < 0.7 else rng.uniform(0.2, 2.0)
    cases.append(dict(c=[*c, r], o=list(o), d=list(d), tmin=tmin, tmax=tmax, expect=sphere_hit(c, r, o, d, tmin, tmax)))
# inside-the-sphere and negative-root cases
cases.append(dict(c=[0, 0, 0, 2], o=[0.5, 0, 0], d=[1, 0.2, 0], tmin=0.001, tmax=float("inf"),
                  expect=sphere_hit(np.zeros(3), 2.0, np.array([0.5, 0, 0]), np.array([1, 0.2, 0]), 0.001, float("inf"))))
out["sphere_hit"] = cases


def rect_hit(axis, a0, a1, b0, b1, k, o, d, tmin, tmax):  # aarect.rs:31-48,81-98,150-167
    ia, ib = (1 if axis == 0 else 0), (1 if axis == 2 else 2)
    t = (k - o[axis]) / d[axis]
    if t < tmin or t > tmax:
        return None
    a, b = o[ia] + t * d[ia], o[ib] + t * d[ib]
    if a < a0 or a > a1 or b < b0 or b > b1:
        return None
    n = np.zeros(3)
    n[axis] = 1.0
    front = (d @ n) < 0
    nn = n if front else -n
    return [t, *nn, (a - a0) / (a1 - a0), (b - b0) / (b1 - b0), 1.0 if front else 0.0]


cases = []
for _ in range(45):
    axis = int(rng.integers(0, 3))
    a0, b0 = rng.uniform(-3, 0, 2)
    a1, b1 = a0 + rng.uniform(0.5, 4), b0 + rng.uniform(0.5, 4)
    k = rng.uniform(-2, 2)
    o, d = rng.uniform(-5, 5, 3), rng.normal(0, 1, 3)
    if rng.random() < 0.7:  # aim at (or just outside) the rect so that most cases are hits / near misses
        ia, ib = (1 if axis == 0 else 0), (1 if axis == 2 else 2)
        tgt = np.zeros(3)
        tgt[axis], tgt[ia], tgt[ib] = k, rng.uniform(a0 - 0.3, a1 + 0.3), rng.uniform(b0 - 0.3, b1 + 0.3)
        d = (tgt - o) * rng.uniform(0.3, 2.0)
    cases.append(dict(axis=axis, abk=[a0, a1, b0, b1, k], o=list(o), d=list(d), tmin=0.001, tmax=float("inf"),
                      expect=rect_hit(axis, a0, a1, b0, b1, k, o, d, 0.001, float("inf"))))
# closed-interval edge: a ray that lands exactly on the rect's corner is a hit (strict < / > rejects, aarect.rs:38)
cases.append(dict(axis=1, abk=[0, 1, 0, 1, 0], o=[1, 1, 1], d=[0, -1, 0], tmin=0.001, tmax=float("inf"),
                  expect=rect_hit(1, 0, 1, 0, 1, 0, np.array([1., 1, 1]), np.array([0., -1, 0]), 0.001, float("inf"))))
out["rect_hit"] = cases

cases = []
for _ in range(20):  # vec3.rs:115-117, 246-251
    n = rng.normal(0, 1, 3)
    n /= np.linalg.norm(n)
    v = rng.normal(0, 1, 3)
    v /= np.linalg.norm(v)
    if v @ n > 0:
        v = -v
    eta = float(rng.choice([1.5, 1 / 1.5, 1.0, 2.4]))
    refl = v - 2 * (v @ n) * n
    cos_t = min(-v @ n, 1.0)
    perp = eta * (v + cos_t * n)
    par = -math.sqrt(abs(1.0 - perp @ perp)) * n
    cases.append(dict(v=list(v), n=list(n), eta=eta, reflect=list(refl), refract=list(perp + par)))
out["reflect_refract"] = cases

cases = []
for _ in range(20):  # onb.rs:19-42
    n = rng.normal(0, 1, 3) * rng.uniform(0.1, 5)
    w = n / np.linalg.norm(n)
    a = np.array([0., 1, 0]) if abs(w[0]) > 0.9 else np.array([1., 0, 0])
    v = np.cross(w, a)
    v /= np.linalg.norm(v)
    u = np.cross(w, v)
    cases.append(dict(n=list(n), uvw=[*u, *v, *w]))
cases.append(dict(n=[1, 0, 0], uvw=[0, 1, 0, 0, 0, 1, 1, 0, 0]))  # |w.x| > 0.9 branch: a = (0,1,0), v = w x a = (0,0,1), u = w x v = (0,-1,0)?
w = np.array([1., 0, 0]); v = np.cross(w, [0., 1, 0]); u = np.cross(w, v)
cases[-1]["uvw"] = [*u, *v, *w]
out["onb"] = cases


def philox(pixel, sample, block, bounce, seed):
    M = 0xFFFFFFFF
    c = [block, bounce, seed, 0x52544232]
    k = [pixel, sample]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & M, p1 & M, ((p0 >> 32) ^ c[3] ^ k[1]) & M, p0 & M]
        k = [(k[0] + 0x9E3779B9) & M, (k[1] + 0xBB67AE85) & M]
    return c


out["philox"] = [dict(args=list(map(int, a)), expect=philox(*map(int, a)))
                 for a in [(0, 0, 0, 0, 0), (1, 2, 3, 4, 5), (809999, 499, 2, 50, 1), (0xFFFFFFFF, 0xFFFFFF, 9, 255, 0xDEADBEEF)]
                 + [tuple(rng.integers(0, 2 ** 32, 5)) for _ in range(8)]]

# light pdfs: closed forms.  Sphere (sphere.rs:75-84): pdf * solid angle of the cap = 1.  XzRect (aarect.rs:107-117)
cases = []
for _ in range(10):
    c, r = rng.uniform(-3, 3, 3), rng.uniform(0.3, 1.5)
    o = c + rng.normal(0, 1, 3) * 4
    if np.linalg.norm(o - c) < 1.2 * r:
        o = c + (o - c) / np.linalg.norm(o - c) * 3 * r
    cos_max = math.sqrt(1 - r * r / ((c - o) @ (c - o)))
    cases.append(dict(c=[*c, r], o=list(o), d=list((c - o) * rng.uniform(0.5, 2)), pdf=1 / (2 * math.pi * (1 - cos_max))))
out["sphere_pdf"] = cases
cases = []
for _ in range(10):
    x0, z0 = rng.uniform(-2, 0, 2)
    x1, z1 = x0 + rng.uniform(1, 3), z0 + rng.uniform(1, 3)
    k = rng.uniform(2, 5)
    o = np.array([rng.uniform(x0, x1), rng.uniform(-1, 1), rng.uniform(z0, z1)])
    target = np.array([rng.uniform(x0, x1), k, rng.uniform(z0, z1)])
    d = (target - o) * rng.uniform(0.3, 2.5)
    dist2 = (target - o) @ (target - o)
    cosine = abs(d[1]) / np.linalg.norm(d)
    cases.append(dict(abk=[x0, x1, z0, z1, k], o=list(o), d=list(d), pdf=dist2 / (cosine * (x1 - x0) * (z1 - z0))))
out["xzrect_pdf"] = cases

# write_color main.rs:141-169
cases = []
for s, spp in [([50.0, 200.0, 1e9], 100), ([float("nan"), 1.0, 0.25], 1), ([0.0, 1080.0, 540.0], 1080), ([3.7, 0.001, 99.0], 16)]:
    exp = []
    for v in s:
        v = 0.0 if v != v else v
        v = math.sqrt(v / spp)
        v = min(max(v, 0.0), 0.999)
        exp.append(int(256.0 * v))
    cases.append(dict(sum=[("nan" if v != v else v) for v in s], spp=spp, expect=exp))
out["write_color"] = cases

# Perlin::noise / turb with the reference's double smoothing, perlin.rs:26-52,67-98
prng = np.random.default_rng(5)
ranvec = prng.uniform(-1, 1, (256, 3))
ranvec /= np.linalg.norm(ranvec, axis=1, keepdims=True)
perms = [prng.permutation(256) for _ in range(3)]


def noise(p):
    fl = np.floor(p)
    u, v, w = p - fl
    u, v, w = (x * x * (3 - 2 * x) for x in (u, v, w))
    i, j, k = (int(x) for x in fl)
    uu, vv, ww = (x * x * (3 - 2 * x) for x in (u, v, w))
    acc = 0.0
    for a in range(2):
        for b in range(2):
            for c in range(2):
                g = ranvec[perms[0][(i + a) & 255] ^ perms[1][(j + b) & 255] ^ perms[2][(k + c) & 255]]
                acc += ((a * uu + (1 - a) * (1 - uu)) * (b * vv + (1 - b) * (1 - vv)) * (c * ww + (1 - c) * (1 - ww))
                        * (g @ np.array([u - a, v - b, w - c])))
    return acc


def turb(p):
    acc, wgt, tp = 0.0, 1.0, np.array(p, dtype=float)
    for _ in range(7):
        acc += wgt * noise(tp)
        wgt *= 0.5
        tp = tp * 2
    return abs(acc)


pts = [list(prng.uniform(-300, 300, 3)) for _ in range(12)] + [[0.0, 0.0, 0.0], [-0.5, 2.25, 7.75]]
out["perlin"] = dict(ranvec=ranvec.tolist(), perm=[p.tolist() for p in perms],
                     cases=[dict(p=p, noise=noise(np.array(p)), turb=turb(p)) for p in pts])

# camera.rs:21-70
cases = []
for lookfrom, lookat, vfov, aspect, ap, fd in [((13, 2, 3), (0, 0, 0), 20, 16 / 9, 0.1, 10), ((278, 278, -800), (278, 278, 0), 40, 1, 0, 10)]:
    lf, la, vup = np.array(lookfrom, float), np.array(lookat, float), np.array([0., 1, 0])
    h = math.tan(math.radians(vfov) / 2)
    vh, vw = 2 * h, aspect * 2 * h
    w = (lf - la) / np.linalg.norm(lf - la)
    u = np.cross(vup, w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    hor, ver = fd * vw * u, fd * vh * v
    llc = lf - hor / 2 - ver / 2 - fd * w
    for _ in range(3):
        s, t, dx, dy = rng.random(), rng.random(), rng.uniform(-.7, .7), rng.uniform(-.7, .7)
        off = u * (ap / 2 * dx) + v * (ap / 2 * dy)
        cases.append(dict(cam=[*lookfrom, *lookat, 0, 1, 0, vfov, aspect, ap, fd, 0, 1], s=s, t=t, dx=dx, dy=dy,
                          o=list(lf + off), d=list(llc + s * hor + t * ver - lf - off)))
out["camera"] = cases

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_vectors.json")
with open(path, "w") as f:
    json.dump(out, f)
print("wrote", path, os.path.getsize(path), "bytes")
