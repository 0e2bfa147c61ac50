"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see rt_oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
PARITY UNPINNED: the reference has no golden vectors and cannot be built here; see rt_oracle.hpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None
_VP, _U32 = C.c_void_p, C.c_uint32


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("rt_oracle.cpp", "rt_oracle.hpp")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.orc_scene_create.restype = _VP
        lib.orc_scene_create.argtypes = [_VP, _U32, _VP, _U32, _U32, _VP, _U32, _VP, _U32, _VP, _U32]
        lib.orc_scene_destroy.argtypes = [_VP]
        lib.orc_scene_error.restype = C.c_char_p
        lib.orc_scene_error.argtypes = [_VP]
        lib.orc_scene_set_image.argtypes = [_VP, _U32, _VP, _U32, _U32]
        lib.orc_scene_set_perlin.argtypes = [_VP, _U32, _VP, _VP, _VP, _VP]
        lib.orc_scene_set_mesh.argtypes = [_VP, _U32, _VP, _U32, _VP, _U32]
        lib.orc_scene_build.argtypes = [_VP]
        lib.orc_scene_attach_bvh.argtypes = [_VP, _VP, _U32, _VP, _U32, _VP, _U32, _VP, _U32, _VP, _U32, _VP, _U32]
        lib.orc_scene_use_bvh.argtypes = [_VP, C.c_int]
        lib.orc_scene_num_prims.restype = _U32
        lib.orc_scene_num_prims.argtypes = [_VP]
        lib.orc_primary_hits.argtypes = [_VP, _VP, _U32, _U32, _VP, _VP, _VP, _VP, C.c_double, C.c_int]
        lib.orc_trace_rays.argtypes = [_VP, _VP, _VP, _VP, _U32, _VP, _VP]
        lib.orc_render.argtypes = [_VP, _VP, _VP, _VP, _VP, _VP, C.c_int]
        lib.orc_write_color.argtypes = [_VP, _U32, _VP]
        lib.orc_kat_sphere_hit.argtypes = [_VP, _VP, _VP, C.c_double, C.c_double, _VP]
        lib.orc_kat_rect_hit.argtypes = [C.c_int, _VP, _VP, _VP, C.c_double, C.c_double, _VP]
        lib.orc_kat_reflect.argtypes = [_VP, _VP, _VP]
        lib.orc_kat_refract.argtypes = [_VP, _VP, C.c_double, _VP]
        lib.orc_kat_onb.argtypes = [_VP, _VP]
        lib.orc_kat_sphere_pdf.restype = C.c_double
        lib.orc_kat_sphere_pdf.argtypes = [_VP, _VP, _VP]
        lib.orc_kat_sphere_random.argtypes = [_VP, _VP, C.c_double, C.c_double, _VP]
        lib.orc_kat_xzrect_pdf.restype = C.c_double
        lib.orc_kat_xzrect_pdf.argtypes = [_VP, _VP, _VP]
        lib.orc_kat_philox.argtypes = [_U32, _U32, _U32, _U32, _U32, _VP]
        lib.orc_kat_perlin_noise.restype = C.c_double
        lib.orc_kat_perlin_noise.argtypes = [_VP, _U32, _VP]
        lib.orc_kat_perlin_turb.restype = C.c_double
        lib.orc_kat_perlin_turb.argtypes = [_VP, _U32, _VP]
        lib.orc_kat_texture.argtypes = [_VP, _U32, C.c_double, C.c_double, _VP, _VP]
        lib.orc_kat_camera_ray.argtypes = [_VP, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _VP]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d(x):
    return np.ascontiguousarray(x, dtype=np.float64)


class OracleScene:
    """The reference's trait-object world rebuilt from the same rtb_node records the product library consumes."""

    def __init__(self, cs):
        lib = load()
        self.lib = lib
        self._keep = cs
        self.h = lib.orc_scene_create(_p(cs.nodes), len(cs.nodes), _p(cs.child_index), len(cs.child_index), cs.root,
                                      _p(cs.materials), len(cs.materials), _p(cs.textures), len(cs.textures),
                                      _p(cs.lights) if len(cs.lights) else None, len(cs.lights))
        for i, img in enumerate(cs.images):
            a = np.ascontiguousarray(img, dtype=np.uint8)
            lib.orc_scene_set_image(self.h, i, _p(a), a.shape[1], a.shape[0])
        for i, pt in enumerate(cs.perlins):
            rv = _d(pt.ranvec)
            px, py, pz = (np.ascontiguousarray(x, dtype=np.uint32) for x in (pt.perm_x, pt.perm_y, pt.perm_z))
            lib.orc_scene_set_perlin(self.h, i, _p(rv), _p(px), _p(py), _p(pz))
        for i, (v, idx) in enumerate(cs.meshes):
            lib.orc_scene_set_mesh(self.h, i, _p(v), len(v), _p(idx), len(idx))
        if lib.orc_scene_build(self.h) != 0:
            raise RuntimeError("oracle scene: " + lib.orc_scene_error(self.h).decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_scene_destroy(self.h)
            self.h = None

    def attach_bvh(self, host_or_device_scene):
        """Use the product's exported BVH to cull candidates (the per-primitive tests stay the reference's own)."""
        nodes, prims = host_or_device_scene.export_bvh()
        self._bvh_keep = (nodes, prims)
        args = [_p(nodes), len(nodes) // 80]
        for g, inf in prims:
            args += [_p(inf) if len(inf) else None, len(inf) // 2]
        # primitives the product keeps out of the tree ("globals"): always candidates
        refs = host_or_device_scene.export_globals()
        gids = np.array([prims[int(r) >> 29][1][2 * (int(r) & ((1 << 29) - 1))] for r in refs], dtype=np.uint32)
        args += [_p(gids) if len(gids) else None, len(gids)]
        rc = self.lib.orc_scene_attach_bvh(self.h, *args)
        if rc != 0:
            raise RuntimeError("attach_bvh: " + self.lib.orc_scene_error(self.h).decode())

    def use_bvh(self, on=True):
        self.lib.orc_scene_use_bvh(self.h, 1 if on else 0)

    def num_prims(self):
        return self.lib.orc_scene_num_prims(self.h)

    def primary_hits(self, cam, width, height, threads=0, stability_eps=None):
        """Closest hit of the pixel-centre rays.  With stability_eps also returns (mask of pixels whose primitive id is
        well-defined at that relative ray perturbation, spread of t under it) — see rt_oracle.cpp."""
        ids = np.empty((height, width), dtype=np.uint32)
        ts = np.empty((height, width), dtype=np.float64)
        stable = np.ones((height, width), dtype=np.uint8) if stability_eps is not None else None
        spread = np.zeros((height, width), dtype=np.float64) if stability_eps is not None else None
        rc = self.lib.orc_primary_hits(self.h, C.byref(cam), width, height, _p(ids), _p(ts), _p(stable), _p(spread),
                                       float(stability_eps or 0.0), threads)
        assert rc == 0
        if stability_eps is None:
            return ids, ts
        return ids, ts, stable.astype(bool), spread

    def trace_rays(self, origin, direction, time=None):
        o, d = _d(origin).reshape(-1, 3), _d(direction).reshape(-1, 3)
        tm = None if time is None else _d(time)
        ids = np.empty(len(o), dtype=np.uint32)
        ts = np.empty(len(o), dtype=np.float64)
        rc = self.lib.orc_trace_rays(self.h, _p(o), _p(d), _p(tm), len(o), _p(ids), _p(ts))
        assert rc == 0
        return ids, ts

    def render(self, cam, params, threads=0):
        """Returns (accum (H,W,4) float64 = sum R,G,B,Y^2 ; segments ; rejected)."""
        acc = np.zeros((params.height, params.width, 4), dtype=np.float64)
        seg, rej = C.c_uint64(), C.c_uint64()
        rc = self.lib.orc_render(self.h, C.byref(cam), C.byref(params), _p(acc), C.byref(seg), C.byref(rej), threads)
        assert rc == 0
        return acc, seg.value, rej.value
