#!/usr/bin/env python
"""Tiny render of every scene family (for compute-sanitizer): python tools/tiny_render.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ray_tracer_archive_b200 as rtb
from ray_tracer_archive_b200 import scenes
ctx = rtb.Context(0)
cfgs = [scenes.config_cornell(), scenes.config_random_spheres(), scenes.config_final_scene(boxes_per_side=6, n_small=80),
        scenes.config_mesh(nx=40, nz=20)]
smoke = scenes.config_cornell(); smoke.world, smoke.lights, smoke.name = scenes.cornell_smoke(), scenes.cornell_smoke_lights(), "cornell_smoke"
for cfg in cfgs + [smoke]:
    sc = rtb.Scene(ctx, rtb.compile_scene(cfg.world, cfg.lights))
    ids, ts, _ = sc.primary_hits(cfg.camera, 48, 32)
    acc, st = sc.render(cfg.camera, rtb.make_params(48, 32, 8, cfg.max_depth, cfg.background, pool_paths=4096))
    rgb = sc.finalize_rgb8(48, 32, 8)
    print(cfg.name, "paths", st["paths"], "segments", st["segments"], "finite", bool(np.isfinite(acc).all()), "hits", int((ids != 0xFFFFFFFF).sum()))
    sc.close()
ctx.close()
print("tiny_render ok")
