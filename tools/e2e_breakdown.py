#!/usr/bin/env python
"""Where does an e2e step go?  python tools/e2e_breakdown.py C3:500"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ray_tracer_archive_b200 as rtb
from bench import get_config
name, spp = sys.argv[1].split(":")
cfg = get_config(name, int(spp))
t = time.perf_counter(); cs = rtb.compile_scene(cfg.world, cfg.lights); print("compile_scene (python)", time.perf_counter() - t)
ctx = rtb.Context(0)
accum = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.float32, device="cuda")
host = torch.empty((cfg.height, cfg.width, 4), dtype=torch.float32).pin_memory()
prm = rtb.make_params(cfg.width, cfg.height, cfg.spp, cfg.max_depth, cfg.background)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sc = rtb.Scene(ctx); t1 = time.perf_counter()
    sc.set_compiled(cs); t2 = time.perf_counter()
    sc.build_bvh(); t3 = time.perf_counter()
    sc.commit(); t4 = time.perf_counter()
    st = sc.render_device(cfg.camera, prm, accum.data_ptr(), torch.cuda.current_stream().cuda_stream); t5 = time.perf_counter()
    host.copy_(accum, non_blocking=True); torch.cuda.synchronize(); t6 = time.perf_counter()
    sc.close(); t7 = time.perf_counter()
    print(f"iter {it}: create {t1-t0:.4f} set_compiled {t2-t1:.4f} build_bvh {t3-t2:.4f} commit {t4-t3:.4f} render {t5-t4:.4f} (device {st['ms_total']/1e3:.4f}) d2h {t6-t5:.4f} close {t7-t6:.4f}")
