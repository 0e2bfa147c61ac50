"""CPU tier: the C++ mirror of the reference's constructors (include/rtb200_scene.hpp) builds the live cornell_box()
scene (main.rs:337-433) through the C ABI, and serialises exactly the records the Python mirror does."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_scene_mirror")


def _build():
    src = EXE + ".cpp"
    lib_dir = os.path.join(ROOT, "ray_tracer_archive_b200")
    deps = [src, os.path.join(ROOT, "include", "rtb200_scene.hpp"), os.path.join(ROOT, "include", "rtb200.h")]
    if not os.path.exists(EXE) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(EXE):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", EXE, src, "-L" + lib_dir, "-lrtb200",
                               "-Wl,-rpath," + lib_dir])


def _run(*args):
    _build()
    return subprocess.run([EXE, *args], capture_output=True, text=True, timeout=300)


def test_cpp_mirror_builds_cornell_and_matches_python_records(rtb):
    out = _run()
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0].startswith("quads 12 spheres 1 prims 13 lights 2 materials 5 textures 4 nodes ")
    assert "commit_without_context -4" in lines[-1]  # RTB_ERR_STATE
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_cornell()
    cs = rtb.compile_scene(cfg.world, cfg.lights)
    recs = [l.split()[1:] for l in lines if l.startswith("node ")]
    assert len(recs) == len(cs.nodes)
    for r, n in zip(recs, cs.nodes):
        assert [int(x) for x in r[:4]] == [int(n["type"]), int(n["material"]), int(n["first_child"]), int(n["n_children"])]
        np.testing.assert_array_equal(np.array([float(x) for x in r[4:]]), n["p"])


def _final_scene_data(rtb, path):
    """The random numbers and texels of the Python mirror's final_scene(), for the C++ mirror to build the same scene."""
    from ray_tracer_archive_b200 import scenes, _ffi as F
    cfg = scenes.config_final_scene()
    cs = rtb.compile_scene(cfg.world, cfg.lights)
    boxes = np.array([n["p"][:6] for n in cs.nodes if int(n["type"]) == F.NODE_BOX], dtype=np.float64)
    small = np.array([n["p"][:3] for n in cs.nodes if int(n["type"]) == F.NODE_SPHERE and n["p"][3] == 10.0], dtype=np.float64)
    assert boxes.shape == (400, 6) and small.shape == (1000, 3)
    pt = cs.perlins[0]
    with open(path, "wb") as f:
        f.write(boxes.tobytes())
        f.write(small.tobytes())
        f.write(np.ascontiguousarray(pt.ranvec, dtype=np.float64).tobytes())
        for perm in (pt.perm_x, pt.perm_y, pt.perm_z):
            f.write(np.ascontiguousarray(perm, dtype=np.uint32).tobytes())
        f.write(np.ascontiguousarray(cs.images[0], dtype=np.uint8).tobytes())
    return cfg, cs


def _check_final_records(out, cs):
    lines = out.stdout.strip().splitlines()
    recs = [l.split()[1:] for l in lines if l.startswith("node ")]
    assert len(recs) == len(cs.nodes)
    for r, n in zip(recs, cs.nodes):
        assert [int(x) for x in r[:4]] == [int(n["type"]), int(n["material"]), int(n["first_child"]), int(n["n_children"])]
        np.testing.assert_array_equal(np.array([float(x) for x in r[4:]]), n["p"])
    mats = [l.split()[1:] for l in lines if l.startswith("material ")]
    assert len(mats) == len(cs.materials)
    for r, m in zip(mats, cs.materials):
        assert (int(r[0]), int(r[1]), float(r[2])) == (int(m["type"]), int(m["texture"]), float(m["param"]))
    texs = [l.split()[1:] for l in lines if l.startswith("texture ")]
    assert len(texs) == len(cs.textures)
    for r, t in zip(texs, cs.textures):
        assert [int(x) for x in r[:4]] == [int(t["type"]), int(t["even"]), int(t["odd"]), int(t["table"])]
        np.testing.assert_array_equal(np.array([float(x) for x in r[4:7]]), t["rgb"])
    return lines


def test_cpp_mirror_final_scene_records_match_python(rtb, tmp_path):
    """The book-2 final scene (main.rs:521-649: 400 boxes under a BVHNode, moving sphere, both ConstantMedium, the earthmap
    and Perlin textures, 1000 spheres under Translate(RotateY(BVHNode))) written with the C++ mirror emits exactly the
    node / material / texture records the Python mirror emits — the records the Rust shim's flatten() methods emit
    (integration/rust/src/flatten_impls.rs follows the same post-order, share-by-address rules)."""
    data = str(tmp_path / "final_scene.bin")
    cfg, cs = _final_scene_data(rtb, data)
    out = _run("final", data)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr
    lines = _check_final_records(out, cs)
    assert lines[-1] == "final quads 2401 spheres 1005 moving 1 media 2 prims 3409 lights 1"


@pytest.mark.gpu
def test_cpp_mirror_final_scene_renders_like_python(rtb, ctx, tmp_path):
    """... and rendering those records through the C ABI from C++ gives the image the Python-driven render gives (same
    seed: same paths; sums equal to f32 summation order)."""
    data = str(tmp_path / "final_scene.bin")
    cfg, cs = _final_scene_data(rtb, data)
    out = _run("final", data, "render")
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr
    last = out.stdout.strip().splitlines()[-1].split()
    acc, st = rtb.Scene(ctx, cs).render(cfg.camera, rtb.make_params(96, 96, 8, cfg.max_depth, cfg.background, seed=1))
    assert last[0] == "rendered" and int(last[2]) == st["paths"] == 96 * 96 * 8 and int(last[4]) == st["segments"]
    assert abs(float(last[6]) / float(acc.astype(np.float64).sum()) - 1.0) < 1e-5


@pytest.mark.gpu
def test_cpp_mirror_renders():
    out = _run("render")
    assert out.returncode == 0, out.stdout + out.stderr
    last = out.stdout.strip().splitlines()[-1].split()
    assert last[0] == "rendered" and int(last[2]) == 64 * 64 * 16 and int(last[4]) > 64 * 64 * 16


@pytest.mark.gpu
def test_cpp_multi_gpu_context_reduces_inside_the_library():
    """SURVEY §8b/§8e through the C ABI from C++: rtb_context_create_multi over every GPU of the box (ncclCommInitAll),
    rtb_render splits the samples and issues one ncclReduce; against the same render on one GPU (tier D1: <= 1e-6
    relative on the sums).  Needs >= 2 GPUs (gpurun --gpus N); skipped on a single-GPU box."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    out = _run("multi", str(n))
    assert out.returncode == 0, out.stdout + out.stderr
    f = out.stdout.strip().splitlines()[-1].split()
    print(out.stdout.strip().splitlines()[-1])
    assert f[0] == "multi" and int(f[2]) == n and int(f[4].rstrip(")")) == n
    assert f[6] == f[8] and f[10] == f[12]                # same paths, same segments: the same sample set
    assert float(f[14]) <= 1e-6 and float(f[16]) <= 2e-5  # sums equal to f32 summation order
    assert float(f[18]) > 0.0
