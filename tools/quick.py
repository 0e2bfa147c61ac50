#!/usr/bin/env python
"""Lean A/B timer (GPU box): python tools/quick.py C1:100 C2:200 ...  -> Mrays/s per workload (device time of rtb_render)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ray_tracer_archive_b200 as rtb
from bench import get_config
F = rtb._ffi
ctx = rtb.Context(0)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("RTB_"))
for arg in sys.argv[1:]:
    name, spp = arg.split(":")
    cfg = get_config(name, int(spp))
    sc = rtb.Scene(ctx, rtb.compile_scene(cfg.world, cfg.lights))
    prm = lambda fl=0: rtb.make_params(cfg.width, cfg.height, cfg.spp, cfg.max_depth, cfg.background, seed=1, flags=fl, pool_paths=int(os.environ.get('RTB_POOL', '0')))
    sc.render(cfg.camera, prm(), readback=False)
    best = None
    for _ in range(2):
        _, st = sc.render(cfg.camera, prm(), readback=False)
        v = st["segments"] / st["ms_total"] / 1e3
        best = max(best or 0, v)
    _, stt = sc.render(cfg.camera, prm(F.RENDER_TIME_EXTEND), readback=False)
    _, stc = sc.render(cfg.camera, prm(F.RENDER_COUNT), readback=False)
    print(f"[{tag}] {name} spp={spp}: {best:8.1f} Mrays/s  ext_share={stt['ms_extend'] / stt['ms_total']:.3f} "
          f"ext_ms={stt['ms_extend']:.1f} total_ms={stt['ms_total']:.1f} nodes/seg={stc['nodes_visited'] / stc['segments']:.2f} "
          f"prims/seg={stc["prims_tested"] / stc["segments"]:.2f} iters={st["iterations"]} exact_rays={st["exact_rays"] / st["segments"]:.5f} refined={st["refined_rays"] / st["segments"]:.5f}", flush=True)
    sc.close()
