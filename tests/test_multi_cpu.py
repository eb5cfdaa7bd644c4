"""The N>1 path on CPU (gloo, world_size 2): the host-side sharding logic of bench.py / grt_render_multi —
rank r renders the strata s = r (mod N), one reduce(sum) to rank 0 — with the oracle standing in for the GPU
renderer.  Checks that the union of the shards is exactly the reference's stratified set (camera.go:97-99,
277-282) and that the result does not depend on N."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import go_raytracer_b200 as g
    from oracle import oracle_py as O
    s, cfg = g.builtin_scene(6, width=20, spp=16)
    ow = O.OracleWorld(s)
    # the same call shape bench.py makes on a GPU: sample_first=rank, sample_stride=world
    part, _, _, _ = ow.render(cfg, seed=0xC0FFEE, sample_first=rank, sample_stride=world, nthreads=2)
    acc = torch.from_numpy(part.astype(np.float32).reshape(-1))
    dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)            # "max over ranks" timing reduction
    assert t.item() == world
    if rank == 0:
        np.save(out_path, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_strata_shards_reduce_to_the_full_render(tmp_path):
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out).reshape(20, 20, 3)
    import go_raytracer_b200 as g
    from oracle import oracle_py as O
    s, cfg = g.builtin_scene(6, width=20, spp=16)
    full, _, _, _ = O.OracleWorld(s).render(cfg, seed=0xC0FFEE)
    assert np.allclose(got, full, rtol=1e-6, atol=1e-6)


def test_shard_sizes_cover_every_stratum_once():
    # the arithmetic grt_render_device uses: n_my = ceil((S2 - first) / stride)
    for S2 in (1, 9, 16, 100, 4096):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(S2, dtype=int)
            for rank in range(world):
                if rank >= S2:
                    continue
                n_my = (S2 - rank + world - 1) // world
                idx = rank + world * np.arange(n_my)
                assert idx.max() < S2
                seen[idx] += 1
            assert (seen == 1).all()
