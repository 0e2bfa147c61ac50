"""Memory-safety tier S1 (-m gpu).  compute-sanitizer is not allowed on the GPU pool, so the library has a bounds-asserting
build (`make -C csrc checked` -> librtb200_checked.so, -DRTB_CHECKED): every index the kernels form — traversal stack, node,
primitive, pool slot, fix-up queue, chunk list — is checked on the device and failures are counted by kind.  This test renders
every scene family with that build (ragged pools, media, meshes, the exact pass) and requires all counters to be zero."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
sys.path.insert(0, os.environ["RTB_ROOT"])
import numpy as np
import ray_tracer_archive_b200 as rtb
from ray_tracer_archive_b200 import scenes
ctx = rtb.Context(0)
assert (ctx.check_failures() == 0).all(), "not a checked build"
smoke = scenes.config_cornell(); smoke.world, smoke.lights, smoke.name = scenes.cornell_smoke(), scenes.cornell_smoke_lights(), "cornell_smoke"
cfgs = [scenes.config_cornell(), scenes.config_random_spheres(), scenes.config_final_scene(), scenes.config_mesh(nx=200, nz=100), smoke]
for cfg in cfgs:
    sc = rtb.Scene(ctx, rtb.compile_scene(cfg.world, cfg.lights))
    sc.primary_hits(cfg.camera, 160, 90)
    for pool in (0, 1025, 60000):
        acc, st = sc.render(cfg.camera, rtb.make_params(96, 54, 12, cfg.max_depth, cfg.background, pool_paths=pool))
        assert st["paths"] == 96 * 54 * 12 and np.isfinite(acc).all()
    sc.close()
f = ctx.check_failures()
print("check failures", f.tolist())
assert (f == 0).all(), f
ctx.close()
print("checked ok")
'''


def test_checked_build_finds_no_out_of_bounds_index():
    lib = os.path.join(ROOT, "ray_tracer_archive_b200", "librtb200_checked.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "ray_tracer_archive_b200", "csrc"), "checked"])
    env = dict(os.environ, RTB200_LIB=lib, RTB_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "checked ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
