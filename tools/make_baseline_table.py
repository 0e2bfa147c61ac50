#!/usr/bin/env python
"""Regenerates the result tables of BASELINE.md §3 from the bench.py JSON lines under profiles/<round>/.

    python tools/make_baseline_table.py profiles/r3            # prints markdown

One row per config (bench_C1..C5.json: one GPU, full spp) and, if present, the scaling lines bench_n{1,2,4,8}.json
(weak C1 `value` + the C5 `strong_scaling` sub-record)."""
import json
import os
import sys


def load(path):
    try:
        with open(path) as f:
            lines = [l for l in f.read().strip().splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else None
    except OSError:
        return None


def main(d):
    rows = []
    print("| Config | Prims | Res × spp | CPU-faithful Mrays/s (16 cores) | CPU-fair | 1×B200 Mrays/s | e2e (host buffers) | GPU / CPU faithful / fair | `k_extend` roofline | FP32 frac | nodes / prim tests per segment (by type) | extend share | exact pass: re-traced / refined per segment |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for w in ("C1", "C2", "C3", "C4", "C5"):
        j = load(os.path.join(d, f"bench_{w}.json"))
        if not j:
            continue
        c, r = j["config"], j["roofline"]
        cb, cf = j.get("cpu_baseline"), j.get("cpu_baseline_fair")
        a = r["algorithmic"]
        by = ", ".join(f"{k} {v:.2f}" for k, v in a["prim_tests_per_segment_by_type"].items() if v > 0)
        cpu = f"{cb['value']:.3g}" if cb else "—"
        fair = f"{cf['value']:.3g}" if cf else "—"
        ratio = f"{j['e2e']['value'] / cb['value']:.3g}× / {j['e2e']['value'] / cf['value']:.3g}×" if cb and cf else "—"
        print(f"| {c['workload']} | {c['prims']} | {c['width']}×{c['height']} × {c['total_spp']} | {cpu} | {fair} | **{j['value']:.0f}** | {j['e2e']['value']:.0f} | {ratio} | "
              f"{100 * r['frac']:.0f} % of {r['bound']} ({r['roofline_mrays_s']:.0f} Mrays/s) | {100 * r['fp32']['frac']:.1f} % | "
              f"{a['nodes_per_segment']:.2f} / {by} | {100 * r['extend_share_of_step']:.0f} % | "
              f"{r['exact_pass']['rays_retraced_per_segment']:.5f} / {r['exact_pass']['hits_refined_per_segment']:.4f} |")
    j1 = load(os.path.join(d, "bench_C1.json"))
    if j1:
        bw = j1["roofline"]["bandwidths"]
        print(f"\nMeasured in the C1 run by `rtb_measure_bandwidth` (read-only 128-bit streams): L2 {bw['l2_read_gbs']:.0f} GB/s, "
              f"HBM {bw['hbm_read_gbs']:.0f} GB/s, shared memory {bw['shared_read_gbs']:.0f} GB/s; clocks {j1['clocks']['sm_mhz']:.0f} MHz, reasons {j1['clocks']['reasons']}.")
    sc = [(n, load(os.path.join(d, f"bench_n{n}.json"))) for n in (1, 2, 4, 8)]
    sc = [(n, j) for n, j in sc if j]
    if sc:
        base = sc[0][1]
        print("\n| GPUs | C1 WEAK (500 spp per GPU) Mrays/s | e2e | ms/step | vs 1 GPU | reduce 13 MB | C5 STRONG (3840×2160, total spp split) Mrays/s | ms/step | speed-up | efficiency | reduce 133 MB | slowest / fastest rank render ms |")
        print("|---|---|---|---|---|---|---|---|---|---|---|---|")
        for n, j in sc:
            s, s0 = j.get("strong_scaling"), base.get("strong_scaling")
            strong = (f"{s['value']:.0f} ({s['total_spp']} spp, {s['spp_per_gpu']} per GPU) | {s['ms_per_step']:.0f} | {s['value'] / s0['value']:.2f}× | "
                      f"{100 * s['value'] / s0['value'] / n * base['n_gpus']:.1f} % | {s['reduce_ms_rank0']:.2f} ms = {100 * s['reduce_frac_of_step']:.2f} % | "
                      f"{s['ms_render_slowest_rank']:.0f} / {s['ms_render_fastest_rank']:.0f}") if s and s0 else "— | — | — | — | — | —"
            print(f"| {n} | {j['value']:.0f} | {j['e2e']['value']:.0f} | {j['ms_per_step']:.1f} | {j['value'] / base['value']:.2f}× | "
                  f"{j['reduce']['ms_per_step_rank0']:.3f} ms = {100 * j['reduce']['frac_of_step']:.3f} % | {strong} |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/r3")
