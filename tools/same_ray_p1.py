"""Same-ray P1 on CPU: the device header compiled for the host (tests/emul) vs the oracle on IDENTICAL f32 rays."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import helpers as H
import orc
import ray_tracer_archive_b200 as rtb
from ray_tracer_archive_b200 import scenes

emul = H.build_emul()
which = sys.argv[1:] or ["C1", "C2", "C3", "C4"]
mk = {"C1": scenes.config_random_spheres, "C2": scenes.config_cornell, "C3": scenes.config_final_scene, "C4": scenes.config_mesh}
for w in which:
    cfg = mk[w]()
    cs = rtb.compile_scene(cfg.world, cfg.lights)
    hs = rtb.Scene(None, cs)
    osc = orc.OracleScene(cs)
    osc.attach_bvh(hs)
    o, d = H.primary_rays(cfg.camera, cfg.width, cfg.height)
    o32, d32 = o.astype(np.float32), d.astype(np.float32)
    tm = np.full(len(o), cfg.camera.time0, dtype=np.float32)
    t0 = time.time()
    oid, ot = osc.trace_rays(o32.astype(np.float64), d32.astype(np.float64), tm.astype(np.float64))
    t1 = time.time()
    ids, ts, nv, nt = H.emul_trace(emul, hs, o32, d32, tm)
    t2 = time.time()
    mism = ids != oid
    ok = ~mism & (oid != H.NONE)
    rel = np.abs(ts[ok] - ot[ok]) / ot[ok]
    import ctypes as C
    emul.emul_exact_rays.restype = C.c_ulonglong
    emul.emul_refined_rays.restype = C.c_ulonglong
    print("   rays through the exact pass:", emul.emul_exact_rays(), "refined:", emul.emul_refined_rays(), "of", len(oid))
    print(w, len(oid), "mismatch", int(mism.sum()), "max rel t", rel.max(), f"oracle {t1-t0:.1f}s emul {t2-t1:.1f}s")
    idx = np.argwhere(mism)[:10, 0]
    for i in idx:
        print("   ", i, divmod(int(i), cfg.width), "dev", ids[i], ts[i], "orc", oid[i], ot[i])
