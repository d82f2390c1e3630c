"""Two-GPU test of the peer-memory exchange (dirt_peer_exchange through dirt_b200.distributed.PeerExchange) against NCCL's
all-reduce.  Needs two visible GPUs with peer access; skipped elsewhere (the single-GPU box of the round-end run)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, count, steps, result_queue):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    device = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=device)
    try:
        from dirt_b200.distributed import PeerExchange
        ex = PeerExchange(count, device)
        gen = torch.Generator(device=device).manual_seed(100 + rank)
        side = torch.cuda.Stream(device)
        locals_, outs = [], []
        # back-to-back exchanges with no host synchronisation in between: slot reuse across step parities is exercised
        for k in range(steps):
            local = torch.randn(count, generator=gen, device=device) * (k + 1)
            out = torch.empty_like(local)
            side.wait_stream(torch.cuda.current_stream(device))
            ex.exchange(local, out, side)
            locals_.append(local); outs.append(out)
        torch.cuda.synchronize(device)
        worst, identical = 0.0, True
        for local, out in zip(locals_, outs):
            want = local.clone()
            dist.all_reduce(want)
            worst = max(worst, float((out - want).abs().max() / want.abs().max()))
            ref = out.clone()
            dist.broadcast(ref, src=0)
            identical = identical and bool(torch.equal(ref, out))
        result_queue.put((rank, worst, identical, None))
    except Exception as e:   # surfaced by the parent
        result_queue.put((rank, None, None, '%s: %s' % (type(e).__name__, e)))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_matches_nccl_all_reduce(cuda_lib):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    world, count, steps = 2, 4 * 5124, 9
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, count, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, worst, identical, err in results:
        assert err is None, 'rank %d: %s' % (rank, err)
        assert worst < 1e-6, worst          # two addends: the same sum up to the order of one addition
        assert identical                     # every rank holds the same bits
