"""Same-ray closest hit for BOUNCE-like rays (origins on surfaces, random directions): emulated device header vs oracle."""
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
mk = {"C1": scenes.config_random_spheres, "C2": scenes.config_cornell, "C3": scenes.config_final_scene, "C4": scenes.config_mesh}
for w in sys.argv[1:] or ["C1", "C2", "C3", "C4"]:
    cfg = mk[w]()
    cs = rtb.compile_scene(cfg.world, cfg.lights)
    hs = rtb.Scene(None, cs)
    osc = orc.OracleScene(cs)
    osc.attach_bvh(hs)
    W, Hh = cfg.width // 2, cfg.height // 2
    o, d = H.primary_rays(cfg.camera, W, Hh)
    o32, d32 = o.astype(np.float32), d.astype(np.float32)
    tm = np.zeros(len(o), np.float32)
    rng = np.random.default_rng(7)
    tot = mism_tot = 0
    for bounce in range(4):
        oid, ot = osc.trace_rays(o32.astype(np.float64), d32.astype(np.float64), tm.astype(np.float64))
        ids, ts, nv, nt = H.emul_trace(emul, hs, o32, d32, tm)
        if hs.info()["n_media"]:
            from test_host_bvh import _media_ids
            surf = ~np.isin(oid, _media_ids(cs))
        else:
            surf = np.ones(len(oid), bool)
        mism = (ids != oid) & surf
        ok = surf & ~mism & (oid != H.NONE)
        rel = np.abs(ts[ok] - ot[ok]) / ot[ok]
        print(w, "bounce", bounce, len(oid), "mismatch", int(mism.sum()), "max rel t", rel.max() if ok.any() else 0)
        for i in np.argwhere(mism)[:6, 0]:
            print("    ", i, "dev", ids[i], ts[i], "orc", oid[i], ot[i], "o", o32[i], "d", d32[i])
        # next rays: from the device's f32 hit point, random direction (mix of unit-length and long un-normalised)
        hit = ok
        p = (o32[hit] + ts[hit, None] * d32[hit]).astype(np.float32)
        nd = rng.normal(0, 1, p.shape).astype(np.float32)
        scale = np.where(rng.random(len(p)) < 0.3, rng.uniform(50, 400, len(p)), 1.0).astype(np.float32)
        o32, d32 = p, (nd * scale[:, None]).astype(np.float32)
        tm = rng.random(len(p)).astype(np.float32)
