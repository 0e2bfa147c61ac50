"""CPU tier: the C-ABI library loads, exports every symbol include/rtb200.h declares, and its host-side error
behaviour.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(rtb):
    lib = rtb._ffi.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in rtb200.h but not exported by librtb200.so"
        assert n in rtb._ffi.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.rtb_abi_version() == 2


def test_struct_layouts_match_header(rtb):
    F = rtb._ffi
    assert C.sizeof(F.Camera) == 15 * 8
    assert C.sizeof(F.Params) == 13 * 4
    assert C.sizeof(F.Stats) == 20 * 8
    assert F.NODE_DTYPE.itemsize == 112 and F.LIGHT_DTYPE.itemsize == 48


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device error path")
def test_no_cpu_fallback(rtb):
    with pytest.raises(rtb.RtbError) as e:
        rtb.Context(0)
    assert e.value.code == -2  # RTB_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_host_only_scene_errors(rtb):
    from ray_tracer_archive_b200 import scene as S
    lib = rtb._ffi.load()
    m = S.Lambertian.construct((0.5, 0.5, 0.5))
    cs = rtb.compile_scene(S.HittableList([S.Sphere((0, 0, 0), 1.0, m)]))
    s = rtb.Scene(None)
    # build before materials are set
    with pytest.raises(rtb.RtbError) as e:
        s.build_bvh()
    assert e.value.code == -4
    s.set_compiled(cs)
    s.build_bvh()
    assert s.info()["n_prims"] == 1 and s.info()["n_bvh_nodes"] == 1
    with pytest.raises(rtb.RtbError) as e:  # host-only scenes cannot be committed
        s.commit()
    assert e.value.code == -4
    # malformed graph: child index out of range
    bad = cs.nodes.copy()
    bad[-1]["first_child"] = 99
    rc = lib.rtb_scene_set_graph(s.h, rtb._ffi.ptr(bad), len(bad), rtb._ffi.ptr(cs.child_index), len(cs.child_index), cs.root)
    assert rc == -1 and b"child" in lib.rtb_last_error()
    # unsupported light type
    lights = np.zeros(1, dtype=rtb._ffi.LIGHT_DTYPE)
    lights[0]["type"] = 7
    assert lib.rtb_scene_set_lights(s.h, rtb._ffi.ptr(lights), 1) == -5


def test_flatten_ids_and_face_modes(rtb):
    """Primitive ids follow list order with Box -> 6 sides (boxes.rs:19-68); the Cornell scene has ids 0..12."""
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_cornell()
    s = rtb.Scene(None, rtb.compile_scene(cfg.world, cfg.lights))
    info = s.info()
    assert (info["n_quads"], info["n_spheres"], info["n_prims"], info["n_lights"]) == (12, 1, 13, 2)
    _, prims = s.export_bvh()
    quad_info = prims[2][1].reshape(-1, 2)
    ids = sorted(quad_info[:, 0].tolist())
    assert ids == list(range(12))
    mode = {int(i): int(m >> 24) & 15 for i, m in quad_info}  # (bit 31: exact axis-aligned plane)
    assert mode[2] == 1                       # FlipFace(light): front_face toggled (hittable.rs:195-201)
    # Translate(RotateY(Box)): front_face = q, RotateY's object-ray / world-normal test; the Translate re-orients the normal
    # (hittable.rs:173 then 82-83) -> FaceMode Q = 4
    assert all(mode[i] == 4 for i in range(6, 12))
    assert all(mode[i] == 0 for i in (0, 1, 3, 4, 5))
    assert int(prims[0][1][0]) == 12          # the glass sphere is the last object


def test_face_modes_of_wrapper_chains(rtb):
    """flatten.cpp eval_face: the reference's per-hit front_face / normal rewriting (hittable.rs:82-83,173,197-201)
    evaluated symbolically per wrapper chain.  FaceMode: 0 natural, 1 flipped, 2 true, 3 false, 4 q, 5 !q, +8 bare."""
    from ray_tracer_archive_b200 import scene as S
    m = S.Lambertian.construct((0.5, 0.5, 0.5))
    r = lambda: S.XzRect.construct(0.0, 1.0, 0.0, 1.0, 0.0, m)
    T, R, Fl = (lambda x: S.Translate.construct(x, (1.0, 0.0, 0.0))), (lambda x: S.RotateY.construct(x, 30.0)), S.FlipFace.construct
    chains = [(r(), 0), (Fl(r()), 1), (T(r()), 2), (Fl(T(r())), 3), (T(Fl(r())), 2),
              (T(R(r())), 4), (T(R(Fl(r()))), 4), (Fl(T(R(r()))), 5), (R(r()), 12), (Fl(R(r())), 13), (R(T(r())), 12),
              (T(T(R(r()))), 2),          # the second Translate sees a correctly oriented normal: front = true
              (T(R(R(r()))), 2)]          # two rotations: documented fall-back
    world = S.HittableList([c for c, _ in chains])
    s = rtb.Scene(None, rtb.compile_scene(world, None))
    _, prims = s.export_bvh()
    mode = {int(i): int(w >> 24) & 15 for i, w in prims[2][1].reshape(-1, 2)}
    assert [mode[k] for k in range(len(chains))] == [want for _, want in chains]


def test_obj_loader_and_ppm_writer(rtb, tmp_path):
    """OBJ-style mesh -> TriangleMesh -> host flatten."""
    from ray_tracer_archive_b200 import io, scene as S
    obj = ["# unit quad + a triangle", "v 0 0 0", "v 1 0 0", "v 1 1 0", "v 0 1 0", "v 0.5 2 0", "f 1 2 3 4", "f 4/1/1 3/2/2 -1"]
    mesh = io.load_obj(obj, S.Lambertian.construct((0.5, 0.5, 0.5)), scale=2.0, offset=(1, 0, 0))
    assert mesh.indices.tolist() == [[0, 1, 2], [0, 2, 3], [3, 2, 4]]
    np.testing.assert_allclose(mesh.vertices[2], [3, 2, 0])
    s = rtb.Scene(None, rtb.compile_scene(S.HittableList([mesh])))
    assert s.info()["n_triangles"] == 3 and s.info()["n_prims"] == 3
    img = (np.arange(4 * 5 * 3) % 256).astype(np.uint8).reshape(4, 5, 3)
    out = tmp_path / "x.ppm"
    io.save_image(str(out), img)
    raw = out.read_bytes()
    assert raw.startswith(b"P6\n5 4\n255\n") and raw[-60:] == img.tobytes()
    with pytest.raises(ValueError):
        io.load_obj(["v 0 0 0", "f 1 2 3"], mesh.mat)


def test_checkpoint_roundtrip(tmp_path):
    """io.save_checkpoint / load_checkpoint: sums, sample count and the guards against resuming the wrong render."""
    import numpy as np
    import pytest
    from ray_tracer_archive_b200 import io
    acc = np.random.default_rng(3).random((5, 7, 4)).astype(np.float32)
    p = str(tmp_path / "ck")
    io.save_checkpoint(p, acc, spp_done=48, seed=9, scene="cornell")
    got, done = io.load_checkpoint(p, width=7, height=5, seed=9)
    assert done == 48 and got.dtype == np.float32 and np.array_equal(got, acc)
    with pytest.raises(ValueError):
        io.load_checkpoint(p, width=8, height=5)
    with pytest.raises(ValueError):
        io.load_checkpoint(p, seed=10)
    with pytest.raises(ValueError):
        io.save_checkpoint(p, acc[..., :3], 1, 1)
