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
    return subprocess.run([EXE, *args], capture_output=True, text=True, timeout=120)


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
