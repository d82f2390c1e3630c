"""Times the hot kernels of one or more prebuilt variants of libdirt_b200.so on a BASELINE workload.

    python profiles/kbench.py [--workload cfg3] [--steps 20] lib_a.so lib_b.so ...

Each library is loaded in its own process (DIRT_B200_LIB) and run through bench.PreparedStep: forward call, backward
call (shared-geometry accumulation, as the bench step), and each hot kernel alone through the library's timer hooks.
Variants are built here with `python profiles/build_variant.py TAG -DFLAG=...` and travel to the GPU box as .so files.
One JSON line per library.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(args):
    sys.path.insert(0, ROOT)
    import torch
    import bench
    from dirt_b200 import scenes
    gen, kwargs, _ = bench.WORKLOADS[args.workload]
    kwargs = dict(kwargs)
    if args.batch:
        kwargs['batch'] = args.batch
    if args.workload != 'cfg2':
        kwargs['seed'] = 1
    scene = getattr(scenes, gen)(**kwargs)
    device = torch.device('cuda', 0)
    prep = bench.PreparedStep(scene, device)
    for _ in range(3):
        prep.local_step(0)
    torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def kernel(which, fn, n):
        prep.lib.dirt_kernel_timer_enable(which)
        total = 0.0
        for _ in range(n):
            fn()
            total += float(prep.lib.dirt_kernel_timer_elapsed_ms())
        prep.lib.dirt_kernel_timer_enable(0)
        return total / n

    n = args.steps
    out = {'lib': os.path.basename(os.environ.get('DIRT_B200_LIB', 'default')), 'workload': args.workload,
           'forward_call_ms': timed(prep.forward, n), 'backward_call_ms': timed(prep.backward, n),
           'step_ms': timed(lambda: prep.local_step(0), n),
           'forward_kernel_ms': kernel(1, prep.forward, n), 'backward_kernel_ms': kernel(2, prep.backward, n)}
    # the same backward call with per-item vertex gradients ([B,V,.] instead of the batch-accumulated [V,.])
    B, H, W, C, V, F = prep.dims
    gv = torch.empty((B, V, 4), device=device); gc = torch.empty((B, V, C), device=device)

    def backward_per_item():
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        prep.lib.dirt_rasterise_backward(prep._p(prep.vertices), prep._p(prep.faces), prep._p(prep.pixels), prep._p(prep.grad_pixels),
                                         prep._p(prep.face_ids), prep._p(prep.grad_background), prep._p(gv), prep._p(gc),
                                         B, H, W, C, V, F, None, 0, 1, prep._p(prep.workspace), prep.ws_bytes, stream)
    out['backward_kernel_per_item_ms'] = kernel(2, backward_per_item, n)
    prep.capture()
    out['graph_step_ms'] = timed(lambda: prep.graphs[0].replay(), n)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('libs', nargs='*')
    ap.add_argument('--workload', default='cfg3')
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--one', action='store_true')
    args = ap.parse_args()
    if args.one:
        return one(args)
    libs = args.libs or [os.path.join(ROOT, 'dirt_b200', 'libdirt_b200.so')]
    for lib in libs:
        env = dict(os.environ, DIRT_B200_LIB=os.path.abspath(lib))
        cmd = [sys.executable, os.path.abspath(__file__), '--one', '--workload', args.workload, '--steps', str(args.steps)]
        if args.batch:
            cmd += ['--batch', str(args.batch)]
        proc = subprocess.run(cmd, env=env, capture_output=True, text=True)
        line = [l for l in proc.stdout.splitlines() if l.startswith('{')]
        print(line[-1] if line else json.dumps({'lib': lib, 'error': (proc.stderr or proc.stdout)[-400:]}), flush=True)


if __name__ == '__main__':
    main()
