"""Shared helpers for the parity tests (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NONE = 0xFFFFFFFF


def luminance(rgb):
    return 0.2126 * rgb[..., 0] + 0.7152 * rgb[..., 1] + 0.0722 * rgb[..., 2]


def image_stats(accum, spp):
    """accum (H,W,4) = sum R,G,B,Y^2  ->  (mean luminance, variance of the per-sample luminance) per pixel."""
    a = np.asarray(accum, dtype=np.float64)
    mean = luminance(a) / spp
    var = np.maximum(a[..., 3] / spp - mean * mean, 0.0)
    return mean, var


def compare_images(acc_gpu, acc_ref, spp_gpu, spp_ref):
    """P2 criteria: mean relative luminance error of the image, its own Monte-Carlo standard error, and
    the per-pixel deviation in units of the combined standard error.  The standard error gets a floor of 3e-4 relative
    (f32 accumulation of thousands of samples) so zero-variance pixels (pure background) compare by rounding."""
    mg, vg = image_stats(acc_gpu, spp_gpu)
    mr, vr = image_stats(acc_ref, spp_ref)
    mean_rel = abs(mg.mean() - mr.mean()) / max(mr.mean(), 1e-12)
    se2 = vg / spp_gpu + vr / spp_ref
    se_rel = np.sqrt(se2.sum()) / mg.size / max(mr.mean(), 1e-12)
    z = np.abs(mg - mr) / np.sqrt(se2 + (3e-4 * np.abs(mr) + 1e-6) ** 2)
    return mean_rel, z, se_rel


def build_emul():
    """tests/emul/libemul.so: the device header compiled for the host (see emul_traverse.cpp)."""
    src = os.path.join(HERE, "emul", "emul_traverse.cpp")
    out = os.path.join(HERE, "emul", "libemul.so")
    csrc = os.path.join(ROOT, "ray_tracer_archive_b200", "csrc")
    deps = [src, os.path.join(ROOT, "include", "rtb200.h")] + [os.path.join(csrc, f) for f in ("rtb_device.cuh", "rtb_internal.hpp")]
    if not os.path.exists(out) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(out):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wno-attributes", "-D__noinline__=",
                               "-I/usr/local/cuda/include", "-o", out, src])
    lib = C.CDLL(out)
    VP = C.c_void_p
    lib.emul_trace.argtypes = ([VP, C.c_uint32] + [VP] * 8 + [VP] * 3 + [C.c_float, C.c_float, C.c_uint32] +
                               [VP, C.c_uint32, C.c_uint32, C.c_uint32, VP, VP, VP, C.c_uint32, VP, VP, VP, VP])
    return lib


def _p(a):
    return None if a is None or len(a) == 0 else a.ctypes.data_as(C.c_void_p)


def emul_trace(lib, host_scene, origin, direction, time=None, n_snodes=10 ** 6):
    info = host_scene.info()
    nodes, prims = host_scene.export_bvh()
    o = np.ascontiguousarray(origin, dtype=np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(direction, dtype=np.float32).reshape(-1, 3)
    tm = None if time is None else np.ascontiguousarray(time, dtype=np.float32)
    n = len(o)
    ids = np.empty(n, dtype=np.uint32)
    ts = np.empty(n, dtype=np.float32)
    nv, nt = C.c_uint64(), C.c_uint64()
    args = [_p(nodes), info["n_bvh_nodes"]]
    for g, inf in prims:
        args += [_p(g), _p(inf)]
    exact, coord_max, eps_ab = host_scene.export_exact()
    args += [_p(x) for x in exact] + [coord_max, eps_ab, info["global_f64_mask"]]
    glob = host_scene.export_globals()
    args += [_p(glob), len(glob), 1 if len(glob) == info["n_spheres"] + info["n_moving"] + info["n_quads"] + info["n_triangles"] else 0]
    lib.emul_trace(*args, n_snodes, _p(o), _p(d), _p(tm), n, _p(ids), _p(ts), C.byref(nv), C.byref(nt))
    return ids, ts, nv.value, nt.value


def primary_rays(cam, W, H):
    """Pixel-centre primary rays in f64 exactly as camera.rs:21-70 builds them (row 0 = top)."""
    import math
    lf, la, vup = np.array(cam.lookfrom[:]), np.array(cam.lookat[:]), np.array(cam.vup[:])
    h = math.tan(cam.vfov_deg * math.pi / 180.0 / 2.0)
    vh = 2.0 * h
    vw = cam.aspect_ratio * vh
    w = (lf - la) / np.linalg.norm(lf - la)
    u = np.cross(vup, w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    hor, ver = cam.focus_dist * vw * u, cam.focus_dist * vh * v
    llc = lf - hor / 2 - ver / 2 - cam.focus_dist * w
    j = H - 1 - np.arange(H)
    s, t = (np.arange(W) + 0.5) / (W - 1), (j + 0.5) / (H - 1)
    d = llc[None, None, :] + s[None, :, None] * hor[None, None, :] + t[:, None, None] * ver[None, None, :] - lf[None, None, :]
    o = np.broadcast_to(lf, (H, W, 3))
    return o.reshape(-1, 3).copy(), d.reshape(-1, 3).copy()


def check_primary_parity(ids_dev, t_dev, oid, ot, stable, spread, max_unstable_frac=0.005):
    """P1: primitive ids bit-exact and t within 1e-5 relative (plus the oracle's own f32-input uncertainty of t) on
    every pixel whose oracle answer is well-defined at f32 ray resolution."""
    ids_dev, oid = ids_dev.reshape(oid.shape), oid
    unstable = (~stable).sum()
    assert unstable <= max_unstable_frac * oid.size, f"{unstable} unstable pixels"
    bad = (ids_dev != oid) & stable
    assert bad.sum() == 0, f"{bad.sum()} primitive-id mismatches on stable pixels, first at {np.argwhere(bad)[:5]}"
    hit = (oid != NONE) & stable
    err = np.abs(t_dev.reshape(oid.shape)[hit].astype(np.float64) - ot[hit])
    tol = 1e-5 * ot[hit] + spread[hit]
    assert (err <= tol).all(), f"t error {np.max(err / ot[hit]):.3e} relative exceeds 1e-5 (+ input spread)"
    return int(unstable), float(np.max(err / ot[hit])) if hit.any() else 0.0
