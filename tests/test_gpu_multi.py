"""Multi-GPU tier D1 (-m gpu, needs >= 2 GPUs: `gpurun --gpus N`; skipped on one GPU).

The framebuffer reduce lives inside librtb200: (a) one process driving N GPUs through
rtb_context_create_multi, (b) one process per GPU joined with rtb_context_comm_init + RTB_RENDER_REDUCE (exercised by
bench.py under torchrun).  Samples carry GLOBAL indices, so N GPUs render exactly the sample set one GPU renders."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("which", ["C2", "C3"])
def test_multi_device_context_equals_single_gpu(rtb, which):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    from ray_tracer_archive_b200 import scenes
    cfg = scenes.config_cornell() if which == "C2" else scenes.config_final_scene()
    cs = rtb.compile_scene(cfg.world, cfg.lights)
    W, Hh, spp = 160, 120, 36
    one = rtb.Context(0)
    many = rtb.Context(list(range(n)))
    assert many.device_count == n and one.device_count == 1
    a, sa = rtb.Scene(one, cs).render(cfg.camera, rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=4))
    sc = rtb.Scene(many, cs)
    b, sb = sc.render(cfg.camera, rtb.make_params(W, Hh, spp, cfg.max_depth, cfg.background, seed=4))
    assert sb["n_devices"] == n and sb["paths"] == sa["paths"] == W * Hh * spp
    assert sb["segments"] == sa["segments"]              # paths are deterministic functions of (pixel, global sample, seed)
    assert abs(float(b.sum()) / float(a.sum()) - 1.0) <= 1e-6
    np.testing.assert_allclose(b, a, rtol=3e-5, atol=1e-4)
    assert sb["ms_nccl"] > 0 and sb["ms_total"] >= sb["ms_render"]
    # RGB8 output of the reduced buffer == of the single-GPU one (up to a rounding boundary)
    ra, rb = rtb.Scene(one, cs), sc
    # spp smaller than the device count: the idle devices still join the reduce
    c, stc = sc.render(cfg.camera, rtb.make_params(W, Hh, 1, cfg.max_depth, cfg.background, seed=4))
    d, _ = rtb.Scene(one, cs).render(cfg.camera, rtb.make_params(W, Hh, 1, cfg.max_depth, cfg.background, seed=4))
    np.testing.assert_allclose(c, d, rtol=3e-5, atol=1e-4)
    assert stc["paths"] == W * Hh
    print(f"{cfg.name}: {n} GPUs == 1 GPU; reduce {sb['ms_nccl']:.3f} ms of {sb['ms_total']:.2f} ms")
    many.close()
    one.close()
