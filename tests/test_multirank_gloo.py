"""CPU tier: the N>1 host logic with torch.distributed (gloo, world_size 2): per-rank sample ranges, global sample
indices and the accumulation-buffer reduce.  The per-rank renderer here is the oracle (no GPU in this tier); on GPUs
the same code path reduces the CUDA accumulation buffers with NCCL (bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import torch
    import torch.distributed as dist
    import orc
    import ray_tracer_archive_b200 as rtb
    from ray_tracer_archive_b200 import parallel, scenes
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cfg = scenes.config_cornell()
    osc = orc.OracleScene(rtb.compile_scene(cfg.world, cfg.lights))
    spp = 9
    first, count = parallel.rank_sample_range(spp, rank, world)
    prm = rtb.make_params(16, 16, count, seed=5, sample_offset=first, total_spp=spp)
    acc, seg, _ = osc.render(cfg.camera, prm, threads=2)
    t = torch.from_numpy(acc)
    segs = torch.tensor([seg], dtype=torch.int64)
    parallel.reduce_accum(t, dst=0)
    dist.reduce(segs, dst=0)
    if rank == 0:
        np.save(out_path, np.concatenate([t.numpy().ravel(), [float(segs.item())]]))
    dist.destroy_process_group()


def test_rank_sample_range_partitions():
    from ray_tracer_archive_b200.parallel import rank_sample_range
    for spp in (1, 7, 16, 1000, 16384):
        for world in (1, 2, 3, 4, 8):
            ranges = [rank_sample_range(spp, r, world) for r in range(world)]
            assert sum(c for _, c in ranges) == spp
            pos = 0
            for f, c in ranges:
                assert f == pos
                pos += c
    with pytest.raises(ValueError):
        rank_sample_range(8, 2, 2)


def test_world_size_2_reduce_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import orc
    import ray_tracer_archive_b200 as rtb
    from ray_tracer_archive_b200 import scenes
    orc.build()
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    cfg = scenes.config_cornell()
    osc = orc.OracleScene(rtb.compile_scene(cfg.world, cfg.lights))
    ref, seg, _ = osc.render(cfg.camera, rtb.make_params(16, 16, 9, seed=5), threads=2)
    np.testing.assert_allclose(got[:-1].reshape(16, 16, 4), ref, rtol=1e-12, atol=1e-12)
    assert int(got[-1]) == seg
