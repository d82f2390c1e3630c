"""Batch sharding across GPUs and the one collective the path needs.

The reference has no multi-GPU machinery at all: the op simply runs on whichever device it is placed
(one GL thread per CUDA context, csrc/gl_dispatcher.h:27,101-108; tests/multi_gpu_test.py is a crash test).
Every batch item is an independent render (per-item draw loops, csrc/rasterise_egl.cpp:362-380), so the
batch shards over ranks with no exchange inside the op.  When the geometry / colours are parameters shared
by the whole batch, their gradient is the sum over the batch of the per-item gradients: the backward kernel
accumulates the local shard straight into one flat [V*4 | V*C] buffer (DIRT_BWD_SHARED_GEOMETRY), then ONE
all-reduce(sum) of that buffer (72 KB for the 5k-triangle mesh) -- on a side stream, so that it overlaps whatever
the caller enqueues next (bench.py pipelines it under the following step's forward pass).
"""
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
        self.flat = torch.zeros(V * (4 + C), dtype=dtype, device=device)
        self.grad_vertices = self.flat[:V * 4].view(V, 4)
        self.grad_vertex_colors = self.flat[V * 4:].view(V, C)

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
