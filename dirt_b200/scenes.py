"""Synthetic scenes (numpy, float32, seeded) for tests, smoke and bench.

The five BASELINE.json configurations plus the scenes the reference's own tests
draw (tests/square_test.py, tests/rasterise_tests.py:11-89, tests/deferred_grad_test.py:19-55,
samples/simple.py:15-74), restated here with numpy so that the CPU oracle and the CUDA path see
bit-identical inputs.  Every generator returns a dict with
    background [B,H,W,C], vertices [B,V,4] (clip space), vertex_colors [B,V,C], faces [B,F,3] (int32).
"""
import math

import numpy as np


# ---- meshes ---------------------------------------------------------------------------------------

def icosphere(level):
    """Unit icosphere: level 3 -> V=642, F=1280; level 4 -> V=2562, F=5120."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
             (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [tuple(np.array(v, np.float64) / np.linalg.norm(v)) for v in verts]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
             (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(level):
        cache = {}

        def midpoint(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = (np.array(verts[a]) + np.array(verts[b])) / 2.0
                verts.append(tuple(m / np.linalg.norm(m)))
                cache[key] = len(verts) - 1
            return cache[key]

        new_faces = []
        for a, b, c in faces:
            ab, bc, ca = midpoint(a, b), midpoint(b, c), midpoint(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    return np.array(verts, np.float32), np.array(faces, np.int32)


def uv_sphere(n_long, n_lat, displacement=0.0):
    """UV sphere with pole vertices: V = n_long*(n_lat-1)+2, F = 2*n_long*(n_lat-1).
    224 x 112 gives V=24866, F=49728 (BASELINE cfg5).  `displacement` scales the radial bump
    1 + d*sin(5*theta)*sin(7*phi) that creates self-occlusion."""
    verts = [(0.0, 1.0, 0.0)]
    for i in range(1, n_lat):
        phi = math.pi * i / n_lat
        for j in range(n_long):
            theta = 2.0 * math.pi * j / n_long
            r = 1.0 + displacement * math.sin(5.0 * theta) * math.sin(7.0 * phi)
            verts.append((r * math.sin(phi) * math.cos(theta), r * math.cos(phi), r * math.sin(phi) * math.sin(theta)))
    verts.append((0.0, -1.0, 0.0))
    south = len(verts) - 1
    faces = []
    for j in range(n_long):
        faces.append((0, 1 + (j + 1) % n_long, 1 + j))
    for i in range(n_lat - 2):
        a0, b0 = 1 + i * n_long, 1 + (i + 1) * n_long
        for j in range(n_long):
            j1 = (j + 1) % n_long
            faces.append((a0 + j, a0 + j1, b0 + j))
            faces.append((b0 + j, a0 + j1, b0 + j1))
    last = 1 + (n_lat - 2) * n_long
    for j in range(n_long):
        faces.append((south, last + j, last + (j + 1) % n_long))
    return np.array(verts, np.float32), np.array(faces, np.int32)


def cube():
    """samples/simple.py:15-23 build_cube(): 8 vertices, 12 triangles."""
    vertices = [[x, y, z] for z in [-1, 1] for y in [-1, 1] for x in [-1, 1]]
    quads = [[0, 1, 3, 2], [4, 5, 7, 6], [1, 5, 4, 0], [2, 6, 7, 3], [4, 6, 2, 0], [3, 7, 5, 1]]
    triangles = sum([[[a, b, c], [c, d, a]] for [a, b, c, d] in quads], [])
    return np.array(vertices, np.float32), np.array(triangles, np.int32)


def cylinder(radius=0.2, height=0.75, end_offset=0.1, bevel=0.0, segments=10):
    """tests/rasterise_tests.py:11-47 make_cylinder(): y-axis cylinder with bevelled conical ends."""
    angles = np.linspace(0., 2 * math.pi, segments, endpoint=False, dtype=np.float32)
    xz = np.stack([np.cos(angles), np.sin(angles)], axis=1) * radius
    ones = np.ones(segments)
    rings = [
        np.stack([xz[:, 0] * (1. - bevel), ones * -height / 2. - radius * bevel, xz[:, 1] * (1. - bevel)], axis=1),
        np.stack([xz[:, 0], ones * -height / 2., xz[:, 1]], axis=1),
        np.stack([xz[:, 0], ones * height / 2., xz[:, 1]], axis=1),
        np.stack([xz[:, 0] * (1. - bevel), ones * height / 2. + radius * bevel, xz[:, 1] * (1. - bevel)], axis=1),
    ]
    ends = [[0., -height / 2. - end_offset, 0.], [0., height / 2. + end_offset, 0.]]
    vertices = np.concatenate(rings + [ends], axis=0)
    faces = []
    for start in (0, segments, 2 * segments):
        for q in range(segments):
            u0, u1 = start + q, start + (q + 1) % segments
            l0, l1 = u0 + segments, u1 + segments
            faces += [[u0, u1, l0], [l0, u1, l1]]
    for t0 in range(segments):
        t1 = (t0 + 1) % segments
        b0 = t0 + segments * 3
        b1 = (b0 + 1) % segments  # (sic) the reference wraps the bottom cap index like this
        faces += [[segments * 4, t0, t1], [segments * 4 + 1, b0, b1]]
    return vertices.astype(np.float32), np.array(faces, np.int32)


def split_vertices_by_face(vertices, faces):
    """numpy twin of lighting.split_vertices_by_face."""
    return vertices[faces.reshape(-1)], np.arange(faces.shape[0] * 3, dtype=np.int32).reshape(-1, 3)


def vertex_normals(vertices, faces):
    v = vertices[:, :3].astype(np.float64)
    n = np.cross(v[faces[:, 1]] - v[faces[:, 0]], v[faces[:, 2]] - v[faces[:, 0]])
    n /= (np.linalg.norm(n, axis=1, keepdims=True) + 1e-12)
    out = np.zeros_like(v)
    for k in range(3):
        np.add.at(out, faces[:, k], n)
    out /= (np.linalg.norm(out, axis=1, keepdims=True) + 1e-12)
    return out.astype(np.float32)


# ---- transforms (row vectors, right-multiplied; dirt/matrices.py conventions) ------------------------

def rodrigues(v):
    v = np.asarray(v, np.float64) + 1e-12
    angle = np.linalg.norm(v)
    x, y, z = v / angle
    K = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    axis = np.array([x, y, z])
    R = math.cos(angle) * np.eye(3) + (1 - math.cos(angle)) * np.outer(axis, axis) + math.sin(angle) * K
    out = np.eye(4)
    out[:3, :3] = R
    return out


def translation(t):
    out = np.eye(4)
    out[3, :3] = t
    return out


def perspective_projection(near, far, right, aspect):
    top = right * aspect
    out = np.zeros((4, 4))
    out[0, 0] = near / right
    out[1, 1] = near / top
    out[2, 2] = -(far + near) / (far - near)
    out[3, 2] = -2. * far * near / (far - near)
    out[2, 3] = -1.
    return out


def _homogeneous(v):
    return np.concatenate([v[:, :3].astype(np.float64), np.ones((v.shape[0], 1))], axis=1)


def _pack(background, vertices, vertex_colors, faces):
    return dict(background=np.ascontiguousarray(background, np.float32),
                vertices=np.ascontiguousarray(vertices, np.float32),
                vertex_colors=np.ascontiguousarray(vertex_colors, np.float32),
                faces=np.ascontiguousarray(faces, np.int32))


# ---- the reference's own scenes ------------------------------------------------------------------------

def square_scene(width=128, height=128, centre_x=32, centre_y=64, size=16):
    """BASELINE cfg1 = tests/square_test.py:20-36."""
    sq = np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32) * size - size / 2.
    sq = sq + np.array([centre_x, centre_y], np.float32)
    sq = sq * 2. / np.array([width, height], np.float32) - 1.
    vertices = np.concatenate([sq, np.zeros([4, 1], np.float32), np.ones([4, 1], np.float32)], axis=1)
    return _pack(np.zeros([1, height, width, 1]), vertices[None], np.ones([1, 4, 1]),
                 np.array([[[0, 1, 2], [0, 2, 3]]], np.int32))


def cylinder_scene(width=48, height=36, rotation_xy=0.5, translate=(0., 0., -0.25), seed=0, batch=1):
    """tests/rasterise_tests.py:50-89: 80-triangle cylinder split to 240 vertices, perspective view."""
    rng = np.random.default_rng(seed)
    v, f = cylinder(0.2, 0.75, 0.1, 0., 10)
    v, f = split_vertices_by_face(_homogeneous(v), f)
    view1 = np.array([[0.5 * math.cos(rotation_xy), -0.5 * math.sin(rotation_xy), 0, 0],
                      [0.5 * math.sin(rotation_xy), 0.5 * math.cos(rotation_xy), 0, 0],
                      [0, 0, 0.5, 0], [0, 0, 0, 1]])
    clip = v @ view1 @ translation(translate) @ perspective_projection(0.1, 20., 0.2, float(height) / width)
    colours = rng.uniform(size=[batch, v.shape[0], 3])
    bg = np.concatenate([np.tile(rng.uniform(size=[batch, 1, 1, 3]), [1, height // 2, width, 1]),
                         np.ones([batch, height - height // 2, width, 3])], axis=1)
    return _pack(bg, np.tile(clip[None], [batch, 1, 1]), colours, np.tile(f[None], [batch, 1, 1]))


def bent_square_scene(width=32, height=32, rotation=0.2, scale=1.0, translate=(0.1, -0.2, 0.0), channels=3, seed=0):
    """tests/deferred_grad_test.py:19-55: bent two-triangle square, perspective view; `channels` of random
    per-vertex attributes (7 = the deferred test's G-buffer: groups 3+3+1)."""
    rng = np.random.default_rng(seed)
    square_size = 4.0  # tests/deferred_grad_test.py:9
    v = np.array([[-1, -1, 0.], [-1, 1, 0], [1, 1, 0], [1, -1, -1.3]], np.float64) * square_size / 2
    f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    v, f = split_vertices_by_face(v, f)
    world = _homogeneous(v) @ rodrigues([0., 0., rotation]) * scale + np.array(list(translate) + [0.])
    clip = world @ translation([-0.5, 0., -3.5]) @ perspective_projection(0.1, 20., 0.1, float(height) / width)
    colours = rng.uniform(size=[1, v.shape[0], channels])
    bg = rng.uniform(size=[1, height, width, channels])
    return _pack(bg, clip[None], colours, f[None])


def cube_scene(width=640, height=480):
    """samples/simple.py:28-74: lit cube."""
    v, f = cube()
    v, f = split_vertices_by_face(v, f)
    world = _homogeneous(v) @ rodrigues([0., 0.5, 0.])
    normals = np.zeros((v.shape[0], 3))
    fn = np.cross(world[f[:, 1], :3] - world[f[:, 0], :3], world[f[:, 2], :3] - world[f[:, 0], :3])
    fn /= (np.linalg.norm(fn, axis=1, keepdims=True) + 1e-12)
    for k in range(3):
        normals[f[:, k]] = fn
    view = translation([0., -1.5, -3.5]) @ rodrigues([-0.3, 0., 0.])
    clip = world @ view @ perspective_projection(0.1, 20., 0.1, float(height) / width)
    cosines = np.abs(normals @ -np.array([1., 0., 0.]))[:, None]
    colours = np.ones((v.shape[0], 3)) * cosines * 0.8 + 0.2
    return _pack(np.zeros([1, height, width, 3]), clip[None], colours[None], f[None])


def cube_batch(batch=64, width=640, height=480, seed=5):
    """The lit cube of samples/simple.py at a different orientation per batch item: twelve triangles of tens of thousands
    of pixels each -- the large-face path of the rasteriser (faces beyond SMALL_TILE_LIMIT tiles go to the per-image list)."""
    rng = np.random.default_rng(seed)
    v, f = cube()
    v, f = split_vertices_by_face(v, f)
    proj = perspective_projection(0.1, 20., 0.1, float(height) / width)
    view = translation([0., -1.5, -3.5]) @ rodrigues([-0.3, 0., 0.])
    clips, colours = [], []
    for _ in range(batch):
        world = _homogeneous(v) @ rodrigues([rng.uniform(-0.4, 0.4), rng.uniform(0., 2. * np.pi), 0.])
        fn = np.cross(world[f[:, 1], :3] - world[f[:, 0], :3], world[f[:, 2], :3] - world[f[:, 0], :3])
        fn /= (np.linalg.norm(fn, axis=1, keepdims=True) + 1e-12)
        normals = np.zeros((v.shape[0], 3))
        for k in range(3):
            normals[f[:, k]] = fn
        cosines = np.abs(normals @ -np.array([1., 0., 0.]))[:, None]
        clips.append(world @ view @ proj)
        colours.append(np.ones((v.shape[0], 3)) * cosines * 0.8 + 0.2)
    return _pack(np.zeros([batch, height, width, 3]), np.stack(clips), np.stack(colours), np.repeat(f[None], batch, axis=0))


# ---- BASELINE configurations -------------------------------------------------------------------------------

def _posed_sphere_batch(verts, faces, batch, width, height, seed, ndc_radius=0.75, jitter=0.1):
    """Per-item pose: random rotation, distance chosen so the unit sphere has NDC radius ~ndc_radius,
    xy jitter in NDC; projection as samples/simple.py (near .1, far 20, right .1)."""
    rng = np.random.default_rng(seed)
    aspect = float(height) / width
    proj = perspective_projection(0.1, 20., 0.1, aspect)
    hv = _homogeneous(verts)
    # near/right = 1: the silhouette of a unit sphere at distance d has NDC radius 1/sqrt(d^2-1)
    dist = math.sqrt(1.0 + 1.0 / (ndc_radius * ndc_radius))
    out = np.empty((batch, verts.shape[0], 4), np.float32)
    rots = []
    for b in range(batch):
        r = rng.standard_normal(3)
        j = rng.uniform(-jitter, jitter, size=2)
        R = rodrigues(r)
        T = translation([j[0] * dist, j[1] * dist * aspect, -dist])
        out[b] = (hv @ R @ T @ proj).astype(np.float32)
        rots.append(R)
    return out, rots


def config2(seed=0):
    """cfg2: B=1, 256x256, C=3, icosphere-3 (V=642, F=1280), samples/simple.py camera and lighting."""
    W = H = 256
    v, f = icosphere(3)
    world = _homogeneous(v) @ rodrigues([0., 0.5, 0.])
    normals = vertex_normals(world.astype(np.float32), f)
    view = translation([0., -1.5, -3.5]) @ rodrigues([-0.3, 0., 0.])
    clip = world @ view @ perspective_projection(0.1, 20., 0.1, float(H) / W)
    cosines = np.abs(normals.astype(np.float64) @ -np.array([1., 0., 0.]))[:, None]
    colours = np.ones((v.shape[0], 3)) * cosines * 0.8 + 0.2
    return _pack(np.zeros([1, H, W, 3]), clip[None], colours[None], f[None])


def config3(batch=64, width=512, height=512, seed=1, level=4, background='zeros'):
    """cfg3 (the north-star workload): B=64, 512x512, C=4 G-buffer [mask, nx, ny, nz], icosphere-4
    (V=2562, F=5120), per-item pose."""
    v, f = icosphere(level)
    clip, rots = _posed_sphere_batch(v, f, batch, width, height, seed)
    attrs = np.empty((batch, v.shape[0], 4), np.float32)
    for b in range(batch):
        attrs[b, :, 0] = 1.0
        attrs[b, :, 1:] = (v.astype(np.float64) @ rots[b][:3, :3]).astype(np.float32)  # unit normals of the posed sphere
    if background == 'zeros':
        bg = np.zeros([batch, height, width, 4], np.float32)
    else:
        bg = np.random.default_rng(seed + 100).uniform(size=[batch, height, width, 4]).astype(np.float32)
    return _pack(bg, clip, attrs, np.tile(f[None], [batch, 1, 1]))


def config4(batch=32, width=512, height=512, seed=4, level=4):
    """cfg4 (per-GPU shard): C=3 lit colours, icosphere-4, per-item camera on shared geometry."""
    v, f = icosphere(level)
    clip, rots = _posed_sphere_batch(v, f, batch, width, height, seed)
    cols = np.empty((batch, v.shape[0], 3), np.float32)
    for b in range(batch):
        n = v.astype(np.float64) @ rots[b][:3, :3]
        cols[b] = (np.abs(n @ -np.array([1., 0., 0.]))[:, None] * 0.8 + 0.2).astype(np.float32)
    return _pack(np.zeros([batch, height, width, 3]), clip, cols, np.tile(f[None], [batch, 1, 1]))


def config5(batch=64, width=1024, height=1024, seed=3, n_long=224, n_lat=112):
    """cfg5: B=64, 1024x1024, C=3, displaced UV sphere (V=24866, F=49728)."""
    v, f = uv_sphere(n_long, n_lat, displacement=0.1)
    clip, rots = _posed_sphere_batch(v, f, batch, width, height, seed, ndc_radius=0.68)
    base = vertex_normals(v, f)
    cols = np.empty((batch, v.shape[0], 3), np.float32)
    for b in range(batch):
        n = base.astype(np.float64) @ rots[b][:3, :3]
        cols[b] = (np.abs(n @ -np.array([1., 0., 0.]))[:, None] * 0.8 + 0.2).astype(np.float32)
    return _pack(np.zeros([batch, height, width, 3]), clip, cols, np.tile(f[None], [batch, 1, 1]))


def random_soup(batch=2, width=64, height=48, n_faces=60, channels=3, seed=0, behind_camera=False, shared_vertices=True):
    """Random triangle soup in clip space with perspective w, overlapping in depth; optionally with some
    vertices behind the camera (w <= 0) to exercise the homogeneous path."""
    rng = np.random.default_rng(seed)
    V = n_faces + 2 if shared_vertices else 3 * n_faces
    w = rng.uniform(0.5, 3.0, size=(batch, V, 1))
    if behind_camera:
        flip = rng.uniform(size=(batch, V, 1)) < 0.15
        w = np.where(flip, -rng.uniform(0.05, 1.0, size=w.shape), w)
    xy = rng.uniform(-1.4, 1.4, size=(batch, V, 2)) * np.abs(w)
    z = rng.uniform(-1.2, 1.2, size=(batch, V, 1)) * np.abs(w)
    vertices = np.concatenate([xy, z, w], axis=2)
    if shared_vertices:
        faces = np.stack([rng.permutation(V)[:3] for _ in range(batch * n_faces)]).reshape(batch, n_faces, 3)
    else:
        faces = np.tile(np.arange(3 * n_faces).reshape(1, n_faces, 3), [batch, 1, 1])
    colours = rng.uniform(size=(batch, V, channels))
    bg = rng.uniform(size=(batch, height, width, channels))
    return _pack(bg, vertices, colours, faces)
