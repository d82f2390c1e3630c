"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path (sharding + the single all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dirt_b200 import distributed as dd


def test_shard_ranges_partition_the_batch():
    for batch in (0, 1, 7, 64, 255, 256):
        for world in (1, 2, 3, 8):
            ranges = [dd.shard_range(batch, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == batch
            for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
                assert e0 == b1
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, V, C, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)                       # identical on every rank
        gv = torch.randn(batch, V, 4, generator=g)
        gc = torch.randn(batch, V, C, generator=g)
        my_gv, my_gc = dd.shard_batch([gv, gc])
        begin, end = dd.shard_range(batch, rank, world)
        assert my_gv.shape[0] == end - begin
        total = dd.reduce_shared_vertex_grads(my_gv, my_gc)
        expected = torch.cat([gv.sum(0), gc.sum(0)], dim=1)
        # the accumulated form the backward kernel fills directly (DIRT_BWD_SHARED_GEOMETRY): one flat buffer, one all-reduce
        shared = dd.SharedVertexGrads(V, C)
        shared.grad_vertices.copy_(my_gv.sum(0))
        shared.grad_vertex_colors.copy_(my_gc.sum(0))
        assert shared.all_reduce() is None
        err2 = max(float((shared.grad_vertices - gv.sum(0)).abs().max()), float((shared.grad_vertex_colors - gc.sum(0)).abs().max()))
        np.save(os.path.join(out_dir, 'err_%d.npy' % rank), np.maximum((total - expected).abs().max().numpy(), err2))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reduction_matches_single_process(tmp_path):
    world, batch, V, C = 2, 7, 50, 4
    mp.spawn(_worker, args=(world, _free_port(), batch, V, C, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert float(np.load(tmp_path / ('err_%d.npy' % rank))) < 1e-5


def test_reduce_without_process_group():
    gv, gc = torch.ones(3, 5, 4), torch.ones(3, 5, 2)
    out = dd.reduce_shared_vertex_grads(gv, gc)
    assert out.shape == (5, 6) and float(out.min()) == 3.0 and float(out.max()) == 3.0


def test_shared_buffer_views_alias_one_allocation():
    shared = dd.SharedVertexGrads(7, 3)
    assert shared.flat.numel() == 52 and shared.flat.numel() % 4 == 0   # 7 * (4 + 3) = 49 floats, padded to whole 16-byte words
    shared.grad_vertices.fill_(1.0)
    shared.grad_vertex_colors.fill_(2.0)
    assert float(shared.flat[:28].min()) == 1.0 and float(shared.flat[28:49].min()) == 2.0 and float(shared.flat[49:].max()) == 0.0
    assert shared.grad_vertex_colors.data_ptr() % 16 == shared.grad_vertices.data_ptr() % 16
    assert shared.all_reduce() is None   # no process group: nothing to do
