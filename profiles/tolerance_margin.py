#!/usr/bin/env python
"""Worst error/tolerance ratio of the CUDA gradients against the oracle over the randomised parity scenes, repeated
(float atomics reorder the sums from run to run).  usage: python profiles/tolerance_margin.py [repeats]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from conftest import rel_close
from dirt_b200 import scenes, rasterise_ops as ops
from oracle import oracle

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
worst = {}
cases = [('soup c%d s%d' % (c, s), scenes.random_soup(channels=c, seed=s)) for c in (1, 3, 4, 5) for s in (0, 1, 2)]
cases += [('behind s%d' % s, scenes.random_soup(seed=s, behind_camera=True)) for s in (0, 1)]
cases += [('cfg3 2x128', scenes.config3(batch=2, width=128, height=96)), ('cfg5 1x256', scenes.config5(batch=1, width=256, height=256, n_long=96, n_lat=48))]
for name, s in cases:
    pixels_o = oracle.forward(**s)
    gp = np.random.default_rng(0).standard_normal(pixels_o.shape).astype(np.float32)
    gb_o, gv_o, gc_o = oracle.backward(s['vertices'], s['faces'], pixels_o, gp)
    t = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    w = 0.0
    for r in range(reps):
        gb, gv, gc = ops.rasterise_backward_raw(t['vertices'], t['faces'], torch.from_numpy(pixels_o).cuda(), torch.from_numpy(gp).cuda(), None, None)
        w = max(w, rel_close(gv.cpu().numpy(), gv_o)[1], rel_close(gc.cpu().numpy(), gc_o)[1])
    worst[name] = w
    print('%-14s worst error / tolerance = %.3f' % (name, w))
print('max over all: %.3f' % max(worst.values()))
