"""Vectorised numpy rasteriser + gradient: the reference's own CPU style, generalised from a square to meshes.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/dirt_oracle.c).  The one CPU implementation the reference holds
for this path is `get_non_dirt_pixels` of tests/square_test.py:11-17: coverage evaluated analytically on a meshgrid of
pixel centres.  This module is that idea for arbitrary meshes -- per-face bounding box, vectorised integer edge
functions and a z-buffer update (rules S1-S7 of the specification in dirt_oracle.c) -- followed by assemble_grads
(csrc/rasterise_grad_egl.cu:93-236) vectorised over pixels.  It exists so that bench.py can time "numpy on one core"
and "numpy x multiprocessing" next to the C/OpenMP port, and as a third, independently written
check of the C oracle (tests/test_oracle_golden.py): face ids agree, values agree to rounding (numpy has no fused
multiply-add, so a depth key or a barycentric may differ in the last bit).
"""
import numpy as np

KEY_EMPTY = 1 << 23
GUARD_BAND = np.float32(8388608.0)
f32 = np.float32


def _pixel_plane(a, b, c, W, H):
    """NDC plane (a,b,c) -> plane over pixel indices (col,row): S6."""
    return a * (2.0 / W), -b * (2.0 / H), a * (1.0 / W - 1.0) + b * (1.0 - 1.0 / H) + c


def setup(vertices, faces, H, W):
    """Per-face setup of one image, vectorised over faces.  vertices [V,4] f32, faces [F,3] i32."""
    vertices = np.asarray(vertices, np.float32)
    faces = np.asarray(faces, np.int64)
    V, F = vertices.shape[0], faces.shape[0]
    ok = ((faces >= 0) & (faces < V)).all(axis=1)
    p = vertices[np.clip(faces, 0, max(V - 1, 0))] if V > 0 else np.zeros((F, 3, 4), np.float32)   # [F,3,4]
    ok &= np.isfinite(p).all(axis=(1, 2))
    behind = ~(p[:, :, 3] > 0)
    ok &= ~behind.all(axis=1)
    hard = behind.any(axis=1)
    with np.errstate(all='ignore'):
        xn = p[:, :, 0] / p[:, :, 3]
        yn = p[:, :, 1] / p[:, :, 3]
        fx = ((xn + f32(1)) * f32(0.5 * W)) * f32(256)
        fy = ((f32(1) - yn) * f32(0.5 * H)) * f32(256)
        hard |= ~((np.abs(fx) <= GUARD_BAND) & (np.abs(fy) <= GUARD_BAND)).all(axis=1)
        xi = np.where(hard[:, None], 0, np.rint(fx)).astype(np.int64)
        yi = np.where(hard[:, None], 0, np.rint(fy)).astype(np.int64)
        # S6: inverse of [x y w] in double
        x, y, w, z = (p[:, :, k].astype(np.float64) for k in (0, 1, 3, 2))
        c0 = np.stack([y[:, 1] * w[:, 2] - y[:, 2] * w[:, 1], y[:, 2] * w[:, 0] - y[:, 0] * w[:, 2], y[:, 0] * w[:, 1] - y[:, 1] * w[:, 0]], 1)
        c1 = np.stack([w[:, 1] * x[:, 2] - w[:, 2] * x[:, 1], w[:, 2] * x[:, 0] - w[:, 0] * x[:, 2], w[:, 0] * x[:, 1] - w[:, 1] * x[:, 0]], 1)
        c2 = np.stack([x[:, 1] * y[:, 2] - x[:, 2] * y[:, 1], x[:, 2] * y[:, 0] - x[:, 0] * y[:, 2], x[:, 0] * y[:, 1] - x[:, 1] * y[:, 0]], 1)
        det = (x[:, 0] * c0[:, 0] + y[:, 0] * c1[:, 0]) + w[:, 0] * c2[:, 0]
        ok &= (det != 0) & np.isfinite(det)
        rdet = 1.0 / np.where(det != 0, det, 1.0)
        inv = np.stack([c0, c1, c2], 1) * rdet[:, None, None]        # inv[f, row, k]
        gq = [np.stack(_pixel_plane(inv[:, 0, k], inv[:, 1, k], inv[:, 2, k], W, H), 1) for k in range(3)]   # each [F,3]
        gs = np.stack(_pixel_plane((inv[:, 0, 0] + inv[:, 0, 1]) + inv[:, 0, 2], (inv[:, 1, 0] + inv[:, 1, 1]) + inv[:, 1, 2],
                                   (inv[:, 2, 0] + inv[:, 2, 1]) + inv[:, 2, 2], W, H), 1)
        a = (inv[:, 0, 0] * z[:, 0] + inv[:, 0, 1] * z[:, 1]) + inv[:, 0, 2] * z[:, 2]
        b = (inv[:, 1, 0] * z[:, 0] + inv[:, 1, 1] * z[:, 1]) + inv[:, 1, 2] * z[:, 2]
        c = (inv[:, 2, 0] * z[:, 0] + inv[:, 2, 1] * z[:, 1]) + inv[:, 2, 2] * z[:, 2]
        g = np.stack(_pixel_plane(a, b, c, W, H), 1)
        gz = np.stack([0.5 * g[:, 0], 0.5 * g[:, 1], 0.5 * g[:, 2] + 0.5], 1)
    ok &= np.isfinite(gz).all(1) & np.isfinite(gs).all(1) & np.isfinite(np.stack(gq, 1)).all(axis=(1, 2))
    # S5: edge functions of the normal faces
    area2 = (xi[:, 1] - xi[:, 0]) * (yi[:, 2] - yi[:, 0]) - (xi[:, 2] - xi[:, 0]) * (yi[:, 1] - yi[:, 0])
    ok &= hard | (area2 != 0)
    swap = area2 < 0
    px = np.where(swap[:, None], xi[:, [0, 2, 1]], xi)
    py = np.where(swap[:, None], yi[:, [0, 2, 1]], yi)
    A = np.empty((F, 3), np.int64); B = np.empty((F, 3), np.int64); q = np.empty((F, 3), np.int64)
    for k in range(3):
        ia, ib = (k + 1) % 3, (k + 2) % 3
        A[:, k] = py[:, ia] - py[:, ib]
        B[:, k] = px[:, ib] - px[:, ia]
        C = -(A[:, k] * px[:, ia] + B[:, k] * py[:, ia])
        tl = (A[:, k] > 0) | ((A[:, k] == 0) & (B[:, k] > 0))
        q[:, k] = (128 * (A[:, k] + B[:, k]) + C - np.where(tl, 0, 1)) >> 8
    cmin = np.maximum((px.min(1) + 127) >> 8, 0); cmax = np.minimum((px.max(1) - 128) >> 8, W - 1)
    rmin = np.maximum((py.min(1) + 127) >> 8, 0); rmax = np.minimum((py.max(1) - 128) >> 8, H - 1)
    cmin = np.where(hard, 0, cmin); cmax = np.where(hard, W - 1, cmax)
    rmin = np.where(hard, 0, rmin); rmax = np.where(hard, H - 1, rmax)
    ok &= (cmin <= cmax) & (rmin <= rmax)
    cr, rr = cmin.astype(np.float64), rmin.astype(np.float64)
    rebase = lambda pl: np.stack([pl[:, 0], pl[:, 1], (pl[:, 0] * cr + pl[:, 1] * rr) + pl[:, 2]], 1).astype(np.float32)
    return dict(kind=np.where(ok, np.where(hard, 2, 1), 0), A=A, B=B, q=q, cmin=cmin, cmax=cmax, rmin=rmin, rmax=rmax,
                z=gz.astype(np.float32), q0=rebase(gq[0]), q1=rebase(gq[1]), s=rebase(gs), hq=np.stack(gq, 1), hz=gz,
                v=faces.astype(np.int64))


def _depth_key(zf32):
    """S7: rint(z * 2^23) as an integer key; fragments exist iff 0 <= key < 2^23 (anything else -> KEY_EMPTY)."""
    with np.errstate(all='ignore'):
        k = np.rint(zf32.astype(np.float64) * 8388608.0)
    return np.where((k >= 0) & (k < KEY_EMPTY), k, KEY_EMPTY).astype(np.int64)


def visibility(tri, H, W):
    """z-buffer over the faces in ascending order: ids [H,W] (-1 background)."""
    ids = np.full((H, W), -1, np.int32)
    keys = np.full((H, W), KEY_EMPTY, np.int64)
    for f in np.nonzero(tri['kind'])[0]:
        r0, r1, c0, c1 = tri['rmin'][f], tri['rmax'][f] + 1, tri['cmin'][f], tri['cmax'][f] + 1
        rows = np.arange(r0, r1)[:, None]
        cols = np.arange(c0, c1)[None, :]
        if tri['kind'][f] == 1:
            inside = np.ones((r1 - r0, c1 - c0), bool)
            for k in range(3):
                inside &= (tri['A'][f, k] * cols + tri['B'][f, k] * rows + tri['q'][f, k]) >= 0
            zA, zB, zC = tri['z'][f]
            zval = zA * cols.astype(np.float32) + (zB * rows.astype(np.float32) + zC)   # fp32, mul+add (no fma in numpy)
        else:
            hq, hz = tri['hq'][f], tri['hz'][f]
            inside = np.ones((r1 - r0, c1 - c0), bool)
            total = 0.0
            for k in range(3):
                v = hq[k, 0] * cols + (hq[k, 1] * rows + hq[k, 2])
                own = (hq[k, 0] > 0) or (hq[k, 0] == 0 and hq[k, 1] > 0)
                inside &= (v > 0) | ((v == 0) & own)
                total = total + v
            inside &= total > 0
            zval = (hz[0] * cols + (hz[1] * rows + hz[2])).astype(np.float32)
        if not inside.any():
            continue
        key = _depth_key(zval)
        sub_keys = keys[r0:r1, c0:c1]
        win = inside & (key < sub_keys)
        sub_keys[win] = key[win]
        ids[r0:r1, c0:c1][win] = f
    return ids


def gbuffer(tri, ids):
    """(bary0, bary1, bary2, clip_w) per pixel, (-1,-1,-1,inf) where uncovered: rule G, fp32."""
    H, W = ids.shape
    cov = ids >= 0
    f = np.where(cov, ids, 0)
    rows, cols = np.mgrid[0:H, 0:W]
    dc = (cols - tri['cmin'][f]).astype(np.float32)
    dr = (rows - tri['rmin'][f]).astype(np.float32)
    plane = lambda pl: pl[f, 0] * dc + (pl[f, 1] * dr + pl[f, 2])
    with np.errstate(all='ignore'):
        cw = f32(1) / plane(tri['s'])
        b0 = plane(tri['q0']) * cw
        b1 = plane(tri['q1']) * cw
    g = np.stack([b0, b1, (f32(1) - b0) - b1, cw], -1).astype(np.float32)
    g[~cov] = (-1, -1, -1, np.inf)
    return g


def forward(background, vertices, vertex_colors, faces):
    """One image: background [H,W,C], vertices [V,4], vertex_colors [V,C], faces [F,3] -> pixels, ids."""
    H, W, C = background.shape
    tri = setup(vertices, faces, H, W)
    ids = visibility(tri, H, W)
    g = gbuffer(tri, ids)
    cov = ids >= 0
    v = tri['v'][np.where(cov, ids, 0)]
    cols = np.asarray(vertex_colors, np.float64)
    c0, c1, c2 = cols[v[..., 0]], cols[v[..., 1]], cols[v[..., 2]]
    shaded = c2 + (g[..., 0:1].astype(np.float64) * (c0 - c2) + g[..., 1:2].astype(np.float64) * (c1 - c2))
    pixels = np.where(cov[..., None], shaded, background).astype(np.float32)
    return pixels, ids, tri, g


def _group_taps(pixels, c0, n):
    """at(): three components of the group at every pixel, edge-clamped 3x3 neighbourhoods -> taps[dy+1][dx+1] of [H,W,3]."""
    H, W, C = pixels.shape
    if n == 3:
        comp = pixels[..., c0:c0 + 3]
    else:   # flat-order reads of a 1-wide group (0 past the end of the IMAGE here: images are processed one at a time)
        flat = np.concatenate([pixels[..., c0].reshape(-1), np.zeros(2, np.float32)])
        idx = np.arange(H * W).reshape(H, W)
        comp = np.stack([flat[idx], flat[idx + 1], flat[idx + 2]], -1)
    pad = np.pad(comp, ((1, 1), (1, 1), (0, 0)), mode='edge')
    return [[pad[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] for dx in (-1, 0, 1)] for dy in (-1, 0, 1)]


def backward(vertices, faces, pixels, grad_pixels, channel_groups=None, next_image_head=None):
    """assemble_grads for one image, vectorised over pixels.  -> grad_background [H,W,C], grad_vertices [V,4],
    grad_vertex_colors [V,C].  `next_image_head`: the first two values of the following image's `pixels` per channel
    ([2,C]) -- what the flat-order reads of 1-wide groups see past the end of this image (zeros for the last one)."""
    pixels = np.asarray(pixels, np.float32)
    grad_pixels = np.asarray(grad_pixels, np.float32)
    vertices = np.asarray(vertices, np.float32)
    H, W, C = pixels.shape
    V = vertices.shape[0]
    if channel_groups is None:
        channel_groups = [C] if C in (1, 3) else [3] * (C // 3) + [1] * (C % 3)
    tri = setup(vertices, faces, H, W)
    ids = visibility(tri, H, W)
    g = gbuffer(tri, ids)
    cov = ids >= 0
    vid = np.where(cov[..., None], tri['v'][np.where(cov, ids, 0)], -1)    # [H,W,3]
    gv = np.zeros((V, 4), np.float64)
    gc = np.zeros((V, C), np.float64)
    grad_background = np.where(cov[..., None], f32(0), grad_pixels)
    # colour gradients: undilated barycentrics (:135-142)
    for k in range(3):
        idx = vid[..., k][cov]
        for ch in range(C):
            gc[:, ch] += np.bincount(idx, weights=(grad_pixels[..., ch][cov].astype(np.float64) * g[..., k][cov]), minlength=V)
    rows, cols = np.mgrid[0:H, 0:W]
    interior = (cols > 0) & (rows > 0) & (cols < W - 1) & (rows < H - 1)
    odd = ((cols + rows) & 1) == 1
    c0 = 0
    for n in channel_groups:
        t = _group_taps(pixels, c0, n)
        if n == 1 and next_image_head is not None:
            # the last two pixels of the image read into the next image
            flat_extra = np.asarray(next_image_head, np.float32)[:, c0]
            comp = np.concatenate([pixels[..., c0].reshape(-1), flat_extra])
            idx = np.arange(H * W).reshape(H, W)
            full = np.stack([comp[idx], comp[idx + 1], comp[idx + 2]], -1)
            pad = np.pad(full, ((1, 1), (1, 1), (0, 0)), mode='edge')
            t = [[pad[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] for dx in (-1, 0, 1)] for dy in (-1, 0, 1)]
        # at(ox,oy) is image (row - oy, col + ox): t[1 - oy][1 + ox]
        at = lambda ox, oy: t[1 - oy][1 + ox]
        sch = lambda a, b, c, d, e, ff: ((((a + b) - c) - d).astype(np.float64) * 0.09375 + ((e - ff) * f32(0.3125)).astype(np.float64)).astype(np.float32)
        sx = sch(at(-1, -1), at(-1, +1), at(+1, -1), at(+1, +1), at(-1, 0), at(+1, 0))
        sy = sch(at(-1, -1), at(+1, -1), at(-1, +1), at(+1, +1), at(0, -1), at(0, +1))
        l1x = (np.abs(sx[..., 0]) + np.abs(sx[..., 1])) + np.abs(sx[..., 2])
        l1y = (np.abs(sy[..., 0]) + np.abs(sy[..., 1])) + np.abs(sy[..., 2])
        horiz = l1x > l1y
        sign = np.where(odd, -1, 1)
        dx = np.where(horiz, sign, 0); dy = np.where(horiz, 0, sign)          # buffer (y-up) orientation
        # candidate neighbours in image coordinates: (row - dy, col + dx), then the opposite one
        bary = g[..., :3].copy(); clip_w = g[..., 3].copy(); index = vid.copy()
        dilated = np.zeros((H, W), bool)
        for s in (1, -1):
            nr = np.clip(rows - s * dy, 0, H - 1); nc = np.clip(cols + s * dx, 0, W - 1)
            n_vid, n_g = vid[nr, nc], g[nr, nc]
            take = interior & ~dilated & (n_vid[..., 0] != -1) & (n_vid != vid).any(-1) & (g[..., 3] > n_g[..., 3])
            bary[take] = n_g[..., :3][take]; clip_w[take] = n_g[..., 3][take]; index[take] = n_vid[take]
            dilated |= take
        has = index[..., 0] != -1
        gp = grad_pixels[..., c0:c0 + n].astype(np.float64)
        dLdx = (gp * sx[..., :n]).sum(-1); dLdy = (gp * sy[..., :n]).sum(-1)
        vx, vy = vertices[:, 0].astype(np.float64), vertices[:, 1].astype(np.float64)
        safe = np.where(has[..., None], index, 0)
        clip_x = (bary * vx[safe]).sum(-1); clip_y = (bary * vy[safe]).sum(-1)
        with np.errstate(all='ignore'):
            cw = clip_w.astype(np.float64)
            dxv = 0.5 * W / cw; dyv = 0.5 * H / cw
            dxw = -0.5 * W * clip_x / (cw * cw); dyw = -0.5 * H * clip_y / (cw * cw)
        for k in range(3):
            idx = safe[..., k][has]
            bk = bary[..., k].astype(np.float64)[has]
            gv[:, 0] += np.bincount(idx, weights=dLdx[has] * bk * dxv[has], minlength=V)
            gv[:, 1] += np.bincount(idx, weights=dLdy[has] * bk * dyv[has], minlength=V)
            gv[:, 3] += np.bincount(idx, weights=dLdx[has] * bk * dxw[has] + dLdy[has] * bk * dyw[has], minlength=V)
        c0 += n
    return grad_background.astype(np.float32), gv.astype(np.float32), gc.astype(np.float32)


def forward_backward_image(args):
    """One image of a batch, forward then backward (picklable: the unit of work of the multiprocessing baseline)."""
    background, vertices, vertex_colors, faces, grad_pixels, next_head = args
    pixels = forward(background, vertices, vertex_colors, faces)[0]
    return (pixels,) + backward(vertices, faces, pixels, grad_pixels, None, next_head)


def forward_backward_batch(scene, grad_pixels, processes=1):
    """fwd+bwd of a batch (dict of [B,...] arrays as dirt_b200.scenes makes them).  processes > 1: a multiprocessing
    pool over the images.  Note: the flat-order reads of 1-wide groups into the NEXT image use that image's background-free
    forward result, which this per-image recipe does not have; they are taken as zeros (affects <= 2 pixels per image)."""
    B = scene['background'].shape[0]
    jobs = [(scene['background'][b], scene['vertices'][b], scene['vertex_colors'][b], scene['faces'][b], grad_pixels[b], None)
            for b in range(B)]
    if processes <= 1:
        results = [forward_backward_image(j) for j in jobs]
    else:
        import multiprocessing as mp
        with mp.get_context('fork').Pool(processes) as pool:
            results = pool.map(forward_backward_image, jobs, chunksize=1)
    return [np.stack([r[i] for r in results]) for i in range(4)]
