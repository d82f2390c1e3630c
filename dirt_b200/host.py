"""Host-buffer entry point: rasterise (+ gradient) with inputs and outputs in pinned HOST memory.

The reference op only accepts device tensors; a caller whose scene data lives on the host pays the PCIe
transfers around it.  `HostRasteriser` hides them: the batch is cut into chunks and three CUDA streams
(copy-in, compute, copy-out) are chained with events, so the host->device copy of chunk i+1, the kernels of
chunk i and the device->host copy of chunk i-1 overlap (PCIe is full duplex; the kernels are ~1 % of the
transfer time).  Each chunk is one dirt_rasterise_forward + one dirt_rasterise_backward call on slices of
preallocated device buffers; the forward call waits for the chunk's background only, the backward call for its
grad_pixels, and each result is sent down as soon as its kernel has finished.  Geometry (vertices, colours, faces) is small next to the images and goes up once per step;
the chunks carry the two image tensors each way.  Measured at BASELINE cfg3 on a B200 with 12 chunks: 12.4 ms per step against 11.2 ms
for the same transfers with no kernels at all (bench.py, `e2e`; 13.2 ms before the two calls of a chunk waited for their
own upload only, profiles/r02_e2e_chunks_split_waits.txt).
"""
import ctypes

import torch

from . import _lib


class HostRasteriser:
    """fwd+bwd for a fixed problem shape with host-resident tensors.

    step(background, vertices, vertex_colors, faces, grad_pixels) -> dict of pinned host tensors
    (pixels, grad_background, grad_vertices, grad_vertex_colors); all arguments are pinned host tensors
    of the shapes given at construction.  Results are valid after the returned event / `synchronize()`.
    """

    def __init__(self, B, H, W, C, V, F, device=None, chunks=12):
        self.lib = _lib.lib()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (B, H, W, C, V, F)
        self.chunks = max(1, min(int(chunks), B))
        base, extra = divmod(B, self.chunks)
        self.bounds, begin = [], 0
        for i in range(self.chunks):
            n = base + (1 if i < extra else 0)
            self.bounds.append((begin, begin + n))
            begin += n
        dev, f32, i32 = self.device, torch.float32, torch.int32
        self.d = {
            'background': torch.empty((B, H, W, C), dtype=f32, device=dev),
            'vertices': torch.empty((B, V, 4), dtype=f32, device=dev),
            'vertex_colors': torch.empty((B, V, C), dtype=f32, device=dev),
            'faces': torch.empty((B, F, 3), dtype=i32, device=dev),
            'grad_pixels': torch.empty((B, H, W, C), dtype=f32, device=dev),
            'pixels': torch.empty((B, H, W, C), dtype=f32, device=dev),
            'face_ids': torch.empty((B, H, W), dtype=i32, device=dev),
            'grad_background': torch.empty((B, H, W, C), dtype=f32, device=dev),
            'grad_vertices': torch.empty((B, V, 4), dtype=f32, device=dev),
            'grad_vertex_colors': torch.empty((B, V, C), dtype=f32, device=dev),
        }
        self.h_out = {k: torch.empty(self.d[k].shape, dtype=self.d[k].dtype).pin_memory()
                      for k in ('pixels', 'grad_background', 'grad_vertices', 'grad_vertex_colors')}
        nmax = max(e - b for b, e in self.bounds)
        self.ws_bytes = int(self.lib.dirt_workspace_bytes_min(nmax, H, W, C, V, F))   # the backward calls are handed the face ids
        # one workspace per chunk in flight (the backward call reuses the forward's setup records)
        self.ws = [torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev) for _ in range(self.chunks)]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.h2d_bytes = sum(self.d[k].numel() * self.d[k].element_size()
                             for k in ('background', 'vertices', 'vertex_colors', 'faces', 'grad_pixels'))
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.h_out.values())

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    def step(self, background, vertices, vertex_colors, faces, grad_pixels):
        B, H, W, C, V, F = self.shape
        host_in = {'background': background, 'vertices': vertices, 'vertex_colors': vertex_colors, 'faces': faces,
                   'grad_pixels': grad_pixels}
        for k, t in host_in.items():
            if t.is_cuda or tuple(t.shape) != tuple(self.d[k].shape) or t.dtype != self.d[k].dtype:
                raise ValueError('%s must be a host tensor of shape %s and dtype %s' % (k, tuple(self.d[k].shape), self.d[k].dtype))
        current = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(current)
        d = self.d
        # geometry is small next to the images: it goes up in one piece, the per-chunk copies are the two image tensors
        with torch.cuda.stream(self.s_in):
            for k in ('vertices', 'vertex_colors', 'faces'):
                d[k].copy_(host_in[k], non_blocking=True)
        for i, (b0, b1) in enumerate(self.bounds):
            n = b1 - b0
            stream = ctypes.c_void_p(self.s_run.cuda_stream)
            # The forward call needs the background only and the backward call grad_pixels only: each kernel starts as soon
            # as ITS upload has landed and each result goes down as soon as ITS kernel is done, so what cannot overlap
            # anything at the two ends of the pipeline is half a chunk, not a whole one.
            with torch.cuda.stream(self.s_in):
                d['background'][b0:b1].copy_(host_in['background'][b0:b1], non_blocking=True)
                ev_bg = torch.cuda.Event()
                ev_bg.record(self.s_in)
                d['grad_pixels'][b0:b1].copy_(host_in['grad_pixels'][b0:b1], non_blocking=True)
                ev_gp = torch.cuda.Event()
                ev_gp.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_bg)
                rc = self.lib.dirt_rasterise_forward(
                    self._p(d['background'][b0:b1]), self._p(d['vertices'][b0:b1]), self._p(d['vertex_colors'][b0:b1]),
                    self._p(d['faces'][b0:b1]), self._p(d['pixels'][b0:b1]), self._p(d['face_ids'][b0:b1]),
                    n, H, W, C, V, F, self._p(self.ws[i]), self.ws_bytes, stream)
                _lib.check(rc, 'Rasterise')
                ev_fwd = torch.cuda.Event()
                ev_fwd.record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_fwd)
                self.h_out['pixels'][b0:b1].copy_(d['pixels'][b0:b1], non_blocking=True)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_gp)
                rc = self.lib.dirt_rasterise_backward(
                    self._p(d['vertices'][b0:b1]), self._p(d['faces'][b0:b1]), self._p(d['pixels'][b0:b1]),
                    self._p(d['grad_pixels'][b0:b1]), self._p(d['face_ids'][b0:b1]), self._p(d['grad_background'][b0:b1]),
                    self._p(d['grad_vertices'][b0:b1]), self._p(d['grad_vertex_colors'][b0:b1]),
                    n, H, W, C, V, F, None, 0, 1, self._p(self.ws[i]), self.ws_bytes, stream)
                _lib.check(rc, 'RasteriseGrad')
                ev_bwd = torch.cuda.Event()
                ev_bwd.record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_bwd)
                self.h_out['grad_background'][b0:b1].copy_(d['grad_background'][b0:b1], non_blocking=True)
        with torch.cuda.stream(self.s_out):   # after the last chunk's kernels (already awaited on this stream)
            for k in ('grad_vertices', 'grad_vertex_colors'):
                self.h_out[k].copy_(d[k], non_blocking=True)
        current.wait_stream(self.s_out)
        current.wait_stream(self.s_in)
        return self.h_out

    def synchronize(self):
        torch.cuda.synchronize(self.device)
