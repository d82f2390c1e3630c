"""Batch sharding across GPUs and the one collective the path needs.

The reference has no multi-GPU machinery at all: the op simply runs on whichever device it is placed
(one GL thread per CUDA context, csrc/gl_dispatcher.h:27,101-108; tests/multi_gpu_test.py is a crash test).
Every batch item is an independent render (per-item draw loops, csrc/rasterise_egl.cpp:362-380), so the
batch shards over ranks with no exchange inside the op.  When the geometry / colours are parameters shared
by the whole batch, their gradient is the sum over the batch of the per-item gradients: the backward kernel
accumulates the local shard straight into one flat [V*4 | V*C] buffer (DIRT_BWD_SHARED_GEOMETRY), then ONE
all-reduce(sum) of that buffer (82 KB for the 5k-triangle mesh at C=4) -- on a side stream, so that it overlaps whatever
the caller enqueues next (bench.py pipelines it under the following step's forward pass).  `PeerExchange` is that sum as
the library's own kernel over NVLink peer memory (dirt_peer_exchange); `SharedVertexGrads.all_reduce` is the NCCL / gloo form.
"""
import ctypes

import torch
import torch.distributed as dist


def shard_range(batch, rank, world_size):
    """Half-open range [begin, end) of the batch items rank `rank` renders: contiguous shards whose sizes
    differ by at most one item (the first `batch % world_size` ranks take the extra item)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError('need 0 <= rank < world_size')
    base, extra = divmod(int(batch), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank=None, world_size=None):
    """Slices dim 0 of every tensor (all with the same batch size) to this rank's shard."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    batch = tensors[0].shape[0]
    for t in tensors:
        if t.shape[0] != batch:
            raise ValueError('all tensors must share the leading (batch) dimension')
    begin, end = shard_range(batch, rank, world_size)
    return [t[begin:end] for t in tensors]


def reduce_shared_vertex_grads(grad_vertices, grad_vertex_colors, out=None, group=None):
    """Gradient of batch-shared geometry: sum_b grad_vertices[b] | sum_b grad_vertex_colors[b] -> [V, 4+C],
    summed over the local shard and then over all ranks with one all-reduce.  Works without an initialised
    process group (single process: the local sum)."""
    V, C = grad_vertices.shape[1], grad_vertex_colors.shape[2]
    if out is None:
        out = torch.empty((V, 4 + C), dtype=grad_vertices.dtype, device=grad_vertices.device)
    torch.sum(grad_vertices, dim=0, out=out[:, :4])
    torch.sum(grad_vertex_colors, dim=0, out=out[:, 4:])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


class SharedVertexGrads:
    """The gradient of batch-shared geometry as ONE flat buffer: `grad_vertices` [V,4] and `grad_vertex_colors` [V,C] are
    views into it (both 16-byte aligned), which is what `rasterise_backward_raw(..., shared_geometry=True)`-style calls
    fill and what a single all-reduce moves."""

    def __init__(self, V, C, device=None, dtype=torch.float32):
        self.flat = torch.zeros(4 * ((V * (4 + C) + 3) // 4), dtype=dtype, device=device)   # whole 16-byte words
        self.grad_vertices = self.flat[:V * 4].view(V, 4)
        self.grad_vertex_colors = self.flat[V * 4:V * (4 + C)].view(V, C)

    def all_reduce(self, group=None, stream=None):
        """Sums the buffer over all ranks.  With `stream` (CUDA) the collective is enqueued there after everything
        already on the current stream, and the event marking its completion is returned (the caller's stream is not
        blocked); without, it runs on the current stream / synchronously (gloo).  No-op without a process group."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return None
        if stream is None or not self.flat.is_cuda:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            return None
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.flat.device))
        with torch.cuda.stream(stream):
            stream.wait_event(ready)
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            done = torch.cuda.Event()
            done.record(stream)
        return done


class PeerExchange:
    """Sum over the ranks of one node of a flat fp32 buffer through peer memory: dirt_peer_exchange (csrc/exchange.cu), one
    kernel of `world` CTAs per rank and step -- push into every peer's slot, flag, wait for the own flags, add the slots in
    rank order (bit-identical result on every rank).  torch's symmetric memory supplies the peer mapping of the exchange
    area (allocation + handle exchange: plumbing); the data path is the library's kernel.

        ex = PeerExchange(count, device)            # collective: every rank of the group, once
        ex.exchange(local, out, stream)             # every rank, once per step, same order; local != out

    Raises if the peer mapping cannot be set up (no P2P between the devices, symmetric memory unavailable): the caller
    decides whether to fall back to `SharedVertexGrads.all_reduce` (NCCL) and says so.
    """

    def __init__(self, count, device, group=None):
        from . import _lib
        import torch.distributed._symmetric_memory as symm
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('PeerExchange needs an initialised process group')
        self.lib = _lib.lib()
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.count = int(count)
        if self.count <= 0 or self.count % 4:
            raise ValueError('count must be a positive multiple of 4 floats')
        self.device = torch.device(device)
        self.slot_bytes = int(self.lib.dirt_peer_exchange_bytes(self.world, self.count))
        if self.slot_bytes == 0:
            raise ValueError('unsupported world size %d' % self.world)
        flag_bytes = 256 * ((2 * self.world * 4 + 255) // 256)
        self.area = symm.empty((self.slot_bytes + flag_bytes) // 4, dtype=torch.float32, device=self.device)
        self.area.zero_()
        self.handle = symm.rendezvous(self.area, self.group)
        bases = [int(p) for p in self.handle.buffer_ptrs]
        if len(bases) != self.world or bases[self.rank] != self.area.data_ptr():
            raise RuntimeError('symmetric memory returned an unexpected peer table')
        vp = ctypes.c_void_p
        self.slots = (vp * self.world)(*[vp(b) for b in bases])
        self.flags = (vp * self.world)(*[vp(b + self.slot_bytes) for b in bases])
        self.sequence = 0
        torch.cuda.synchronize(self.device)   # the zeroed flags are in memory ...
        dist.barrier(group=self.group)        # ... on every rank before anyone pushes

    def exchange(self, local, out, stream=None):
        """out = sum over ranks of `local` (both flat fp32 CUDA tensors of `count` floats, different storage).  Enqueued on
        `stream` (default: the current stream)."""
        from . import _lib
        for t in (local, out):
            if not t.is_cuda or t.dtype != torch.float32 or t.numel() != self.count or not t.is_contiguous():
                raise ValueError('local / out must be contiguous fp32 CUDA tensors of %d floats' % self.count)
        stream = torch.cuda.current_stream(self.device) if stream is None else stream
        self.sequence += 1
        rc = self.lib.dirt_peer_exchange(ctypes.c_void_p(local.data_ptr()), ctypes.c_void_p(out.data_ptr()), self.slots, self.flags,
                                         self.world, self.rank, self.count, self.sequence, ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, 'PeerExchange')
        return out
