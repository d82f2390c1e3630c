"""Second, independent restatement of the reference in numpy (small cases only).

TEST INFRASTRUCTURE ONLY (see oracle/dirt_oracle.c).  Its purpose is to cross-check the C
oracle with a differently-shaped implementation:

* `square_reference_pixels` is the reference's own CPU path, tests/square_test.py:11-17
  (`get_non_dirt_pixels`), restated in numpy -- the golden vector generator.
* `assemble_grads` follows csrc/rasterise_grad_egl.cu:93-236 literally, in the reference's
  own *buffer* coordinates (x right, y UP, `y_in_frame = H-1 - buffer_y`), one pixel at a
  time, taking the G-buffer (barycentrics, clip_w, vertex indices) as input exactly as the
  CUDA kernel takes its two GL textures.  The C oracle works in image coordinates (y down),
  so agreement between the two checks every sign/orientation translation.
* `coverage_exact` restates the S3-S5 coverage rule with Python integers and Fractions.
"""
from fractions import Fraction

import numpy as np


def square_reference_pixels(canvas_width=128, canvas_height=128, centre_x=32, centre_y=64, square_size=16):
    """tests/square_test.py:11-17 get_non_dirt_pixels()."""
    xs, ys = np.meshgrid(np.arange(canvas_width), np.arange(canvas_height))
    xs = xs.astype(np.float32) + 0.5
    ys = ys.astype(np.float32) + 0.5
    x_in_range = np.abs(xs - centre_x) <= square_size / 2
    y_in_range = np.abs(ys - centre_y) <= square_size / 2
    return np.logical_and(x_in_range, y_in_range).astype(np.float32)


def square_scene(canvas_width=128, canvas_height=128, centre_x=32, centre_y=64, square_size=16):
    """tests/square_test.py:20-36 get_dirt_pixels() inputs: (background, vertices, vertex_colors, faces)."""
    sq = np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32) * square_size - square_size / 2.
    sq = sq + np.array([centre_x, centre_y], np.float32)
    sq = sq * 2. / np.array([canvas_width, canvas_height], np.float32) - 1.
    vertices = np.concatenate([sq, np.zeros([4, 1], np.float32), np.ones([4, 1], np.float32)], axis=1)
    faces = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    vertex_colors = np.ones([4, 1], np.float32)
    background = np.zeros([canvas_height, canvas_width, 1], np.float32)
    return background, vertices.astype(np.float32), vertex_colors, faces


def coverage_exact(vertices, faces, height, width):
    """Coverage + snapped-vertex depth-free face map for w>0, in-guard-band faces, restating S3-S5
    with exact rational arithmetic on the snapped vertices.  Returns a list (per face) of boolean
    masks [H,W].  (No depth test: use on non-overlapping scenes, or combine with your own order.)"""
    f32 = np.float32
    masks = []
    for face in faces:
        p = vertices[face].astype(np.float32)
        pts = []
        for k in range(3):
            xn = f32(p[k, 0] / p[k, 3])
            yn = f32(p[k, 1] / p[k, 3])
            X = f32(f32(xn + f32(1)) * f32(0.5 * width))
            Y = f32(f32(f32(1) - yn) * f32(0.5 * height))
            xi = int(np.rint(f32(X * f32(256))))
            yi = int(np.rint(f32(Y * f32(256))))
            pts.append((xi, yi))
        (ax, ay), (bx, by), (cx, cy) = pts
        area2 = (bx - ax) * (cy - ay) - (cx - ax) * (by - ay)
        mask = np.zeros((height, width), bool)
        if area2 != 0:
            if area2 < 0:
                pts[1], pts[2] = pts[2], pts[1]
            for r in range(height):
                for c in range(width):
                    px, py = 256 * c + 128, 256 * r + 128
                    inside = True
                    for k in range(3):
                        (x0, y0), (x1, y1) = pts[(k + 1) % 3], pts[(k + 2) % 3]
                        A, B = y0 - y1, x1 - x0
                        E = Fraction(A) * (px - x0) + Fraction(B) * (py - y0)
                        topleft = A > 0 or (A == 0 and B > 0)
                        if not (E > 0 or (E == 0 and topleft)):
                            inside = False
                            break
                    mask[r, c] = inside
        masks.append(mask)
    return masks


def assemble_grads(vertices, face_vertex_ids, gbuffer, pixels, grad_pixels):
    """Literal per-pixel restatement of assemble_grads for ONE image and ONE channel group.

    vertices [V,4]; face_vertex_ids [H,W,3] float (vertex indices of the visible face, -1 where
    uncovered) and gbuffer [H,W,4] (bary xyz, clip_w; (-1,-1,-1,inf) uncovered) are indexed in IMAGE
    orientation (row 0 top) and converted to the reference's buffer orientation below.
    pixels / grad_pixels [H,W,Cg] with Cg in {1,3}; for Cg == 1 the out-of-range channel reads of
    `at()` see the next pixels in flat order, so `pixels_flat_tail` semantics are emulated on the
    single image (reads past the end of the image give 0 -- only exact for the last image of a batch).
    Returns grad_vertices [V,4], grad_vertex_colors [V,Cg], grad_background [H,W,Cg] (float64).
    """
    H, W, C = grad_pixels.shape
    V = vertices.shape[0]
    grad_vertices = np.zeros((V, 4), np.float64)
    grad_vertex_colors = np.zeros((V, C), np.float64)
    grad_background = np.zeros((H, W, C), np.float64)
    flat = pixels.reshape(-1).astype(np.float32)
    f32 = np.float32

    def pix3(y_in_frame, x_in_frame):
        if C == 3:
            return pixels[y_in_frame, x_in_frame, :].astype(np.float32)
        base = y_in_frame * W + x_in_frame
        out = np.zeros(3, np.float32)
        for k in range(3):
            if base + k < flat.size:
                out[k] = flat[base + k]
        return out

    # the GL textures are bottom-row-first: texel (buffer_x, buffer_y) shows image row H-1-buffer_y
    def bary_tex(bx, by):
        return gbuffer[H - 1 - by, bx]

    def idx_tex(bx, by):
        return face_vertex_ids[H - 1 - by, bx]

    for buffer_x in range(W):
        for buffer_y in range(H):
            x_in_frame = buffer_x
            y_in_frame = H - 1 - buffer_y

            def at(offset_x, offset_y):
                ux = x_in_frame + offset_x
                uy = y_in_frame - offset_y
                cx_ = max(0, min(W - 1, ux))
                cy_ = max(0, min(H - 1, uy))
                return pix3(cy_, cx_)

            def scharr(a, b, c, d, e, f):
                X = f32(f32(f32(a + b) - c) - d)
                Y = f32(e - f)
                # fmaf(X, 3/32, Y*10/32) -- the contraction nvcc applies to the reference file (oracle/_ref SASS);
                # the single rounding is emulated in float64 (the products are exact there)
                return np.array([f32(np.float64(X[i]) * 0.09375 + np.float64(f32(Y[i] * f32(0.3125))))
                                 for i in range(3)], np.float32)

            scharr_x = scharr(at(-1, -1), at(-1, +1), at(+1, -1), at(+1, +1), at(-1, 0), at(+1, 0))
            scharr_y = scharr(at(-1, -1), at(+1, -1), at(-1, +1), at(+1, +1), at(0, -1), at(0, +1))

            bd = bary_tex(buffer_x, buffer_y)
            barycentric = np.array(bd[:3], np.float32)
            clip_w = f32(bd[3])
            index_f = np.array(idx_tex(buffer_x, buffer_y), np.float32)

            if barycentric[0] != -1.:
                for iip in range(3):
                    vi = int(index_f[iip])
                    for ch in range(C):
                        grad_vertex_colors[vi, ch] += float(grad_pixels[y_in_frame, x_in_frame, ch]) * float(barycentric[iip])
            else:
                for ch in range(C):
                    grad_background[y_in_frame, x_in_frame, ch] = grad_pixels[y_in_frame, x_in_frame, ch]

            if 0 < x_in_frame < W - 1 and 0 < y_in_frame < H - 1:
                l1x = f32(f32(abs(scharr_x[0]) + abs(scharr_x[1])) + abs(scharr_x[2]))
                l1y = f32(f32(abs(scharr_y[0]) + abs(scharr_y[1])) + abs(scharr_y[2]))
                off = (1, 0) if l1x > l1y else (0, 1)
                if (x_in_frame + y_in_frame) % 2 == 1:
                    off = (-off[0], -off[1])
                dilated = False
                for o in (off, (-off[0], -off[1])):
                    if dilated:
                        break
                    idx_o = np.array(idx_tex(buffer_x + o[0], buffer_y + o[1]), np.float32)
                    bd_o = bary_tex(buffer_x + o[0], buffer_y + o[1])
                    cw_o = f32(bd_o[3])
                    if idx_o[0] != -1. and np.any(idx_o != index_f) and clip_w > cw_o:
                        barycentric = np.array(bd_o[:3], np.float32)
                        index_f = idx_o
                        clip_w = cw_o
                        dilated = True

            if barycentric[0] != -1.:
                dL_dx = 0.
                dL_dy = 0.
                for ch in range(C):
                    g = float(grad_pixels[y_in_frame, x_in_frame, ch])
                    dL_dx += g * float(scharr_x[ch])
                    dL_dy += g * float(scharr_y[ch])
                clip_x = 0.
                clip_y = 0.
                for iip in range(3):
                    vi = int(index_f[iip])
                    clip_x += float(barycentric[iip]) * float(vertices[vi, 0])
                    clip_y += float(barycentric[iip]) * float(vertices[vi, 1])
                cw = float(clip_w)
                for iip in range(3):
                    d_xview_by_xclip = .5 * W / cw
                    d_yview_by_yclip = .5 * H / cw
                    d_xview_by_wclip = -.5 * W * clip_x / (cw * cw)
                    d_yview_by_wclip = -.5 * H * clip_y / (cw * cw)
                    tx = dL_dx * float(barycentric[iip])
                    ty = dL_dy * float(barycentric[iip])
                    vi = int(index_f[iip])
                    grad_vertices[vi, 0] += tx * d_xview_by_xclip
                    grad_vertices[vi, 1] += ty * d_yview_by_yclip
                    grad_vertices[vi, 3] += tx * d_xview_by_wclip + ty * d_yview_by_wclip
    return grad_vertices, grad_vertex_colors, grad_background
