"""CPU: the oracle against the reference's only golden vector and against its own second restatement."""
import os

import numpy as np
import pytest

from dirt_b200 import scenes
from oracle import numpy_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def test_square_golden_file_matches_reference_cpu_path():
    # tests/golden/square_test_expected.npy was produced by tests/golden/make_golden.py from the reference's
    # get_non_dirt_pixels() (tests/square_test.py:11-17)
    expected = np.load(os.path.join(GOLDEN, 'square_test_expected.npy'))
    assert expected.shape == (128, 128)
    assert expected.sum() == 256
    np.testing.assert_array_equal(expected, numpy_oracle.square_reference_pixels())


def test_oracle_square_exact(oracle):
    # the reference's acceptance check: "successful: all pixels agree" (tests/square_test.py:54-57)
    s = scenes.square_scene()
    pixels, ids = oracle.forward(**s, return_face_ids=True)
    expected = np.load(os.path.join(GOLDEN, 'square_test_expected.npy'))
    np.testing.assert_array_equal(pixels[0, :, :, 0], expected)
    assert set(np.unique(ids)) == {-1, 0, 1}


@pytest.mark.parametrize('cx,cy,size', [(32, 64, 16), (64, 64, 32), (17, 90, 6), (100, 30, 40)])
def test_oracle_square_variants_exact(oracle, cx, cy, size):
    # the scene is built in GL window space (y up); the image is top-row-first (csrc/rasterise_egl.cu:23,80),
    # so a square centred at window y = cy appears centred at row 128 - cy
    s = scenes.square_scene(128, 128, cx, cy, size)
    pixels = oracle.forward(**s)
    np.testing.assert_array_equal(pixels[0, :, :, 0], numpy_oracle.square_reference_pixels(128, 128, cx, 128 - cy, size))


def test_oracle_vertical_orientation(oracle):
    # row 0 is the top of the image = clip-space y = +1 (csrc/rasterise_egl.cu:23,80): a square placed at
    # clip y > 0 must land in the upper half
    s = scenes.square_scene(64, 64, 32, 48, 8)   # pixel-space y = 48 of 64 -> clip y = +0.5
    pixels = oracle.forward(**s)[0, :, :, 0]
    rows = np.nonzero(pixels.sum(axis=1))[0]
    assert rows.min() == 12 and rows.max() == 19   # rows 64-52 .. 64-44-1


def test_coverage_is_watertight_and_matches_exact_rational_rule(oracle):
    rng = np.random.default_rng(5)
    H, W = 40, 56
    # a fan of triangles sharing edges, vertices off-grid
    centre = np.array([0.07, -0.03])
    ring = [centre + 0.8 * np.array([np.cos(a), np.sin(a)]) for a in np.linspace(0, 2 * np.pi, 9, endpoint=False)]
    verts = np.array([list(centre) + [0.0, 1.0]] + [list(p) + [0.0, 1.0] for p in ring], np.float32)
    verts[:, :2] += rng.uniform(-0.01, 0.01, size=verts[:, :2].shape).astype(np.float32)
    faces = np.array([[0, 1 + i, 1 + (i + 1) % 9] for i in range(9)], np.int32)
    masks = numpy_oracle.coverage_exact(verts, faces, H, W)
    count = np.sum(masks, axis=0)
    assert count.max() == 1, 'a pixel is covered by two triangles of a planar fan'
    ids, _ = oracle.visibility(verts[None], faces[None], H, W)
    np.testing.assert_array_equal(ids[0] >= 0, count == 1)
    for f, m in enumerate(masks):
        np.testing.assert_array_equal(ids[0] == f, m)
    # interior of the fan has no holes: every pixel whose centre is well inside the polygon is covered
    assert (ids[0, 15:25, 20:36] >= 0).all()


def test_depth_order_and_first_drawn_wins_ties(oracle):
    H = W = 32
    def quad(z, x0, x1):
        return [[x0, -0.5, z, 1], [x1, -0.5, z, 1], [x1, 0.5, z, 1], [x0, 0.5, z, 1]]
    verts = np.array(quad(0.5, -0.6, 0.2) + quad(-0.2, -0.2, 0.6) + quad(-0.2, -0.2, 0.6), np.float32)
    faces = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [8, 9, 10], [8, 10, 11]], np.int32)
    ids, gbuf = oracle.visibility(verts[None], faces[None], H, W)
    # nearer quad (z=-0.2) hides the farther one where they overlap; the coincident copy (faces 4,5) never shows
    assert not np.isin(ids, [4, 5]).any()
    overlap = ids[0, 16, 14:18]
    assert np.isin(overlap, [2, 3]).all()
    assert np.isin(ids[0, 16, 8], [0, 1])
    # near / far clipping is per pixel: a quad outside [-1,1] in z is invisible
    verts_far = verts.copy(); verts_far[:, 2] = 1.5
    ids_far, _ = oracle.visibility(verts_far[None], faces[None], H, W)
    assert (ids_far == -1).all()
    verts_far[:, 2] = 1.0   # z_win == 1.0 fails LESS against the cleared depth
    ids_far, _ = oracle.visibility(verts_far[None], faces[None], H, W)
    assert (ids_far == -1).all()


def test_gbuffer_properties(oracle):
    s = scenes.config3(batch=1, width=96, height=96, level=2)
    ids, gbuf = oracle.visibility(s['vertices'], s['faces'], 96, 96)
    cov = ids >= 0
    assert 0.3 < cov.mean() < 0.6
    bary = gbuf[..., :3][cov]
    np.testing.assert_allclose(bary.sum(-1), 1.0, atol=1e-6)
    assert bary.min() > -2e-2   # slightly negative only through 1/256-px snapping
    assert np.isinf(gbuf[..., 3][~cov]).all() and (gbuf[..., :3][~cov] == -1).all()
    # clip_w equals the interpolated clip-space w of the fragment
    v = s['vertices'][0]; f = s['faces'][0]
    w_interp = (gbuf[0][..., :3] * v[f[np.maximum(ids[0], 0)]][..., 3]).sum(-1)
    # perspective-correct interpolation of w itself: sum_k lambda_k w_k == clip_w
    np.testing.assert_allclose(w_interp[cov[0]], gbuf[0][..., 3][cov[0]], rtol=2e-5)


def test_pixels_are_linear_in_colours_and_gradients_exact(oracle):
    # pixels are linear in vertex_colors and background, so the colour/background gradients are exact adjoints
    s = scenes.cylinder_scene()
    rng = np.random.default_rng(3)
    pixels = oracle.forward(**s)
    gp = rng.standard_normal(pixels.shape).astype(np.float32)
    gb, gv, gc = oracle.backward(s['vertices'], s['faces'], pixels, gp)
    dcol = rng.standard_normal(s['vertex_colors'].shape).astype(np.float32)
    dbg = rng.standard_normal(s['background'].shape).astype(np.float32)
    s2 = dict(s, vertex_colors=s['vertex_colors'] + dcol, background=s['background'] + dbg)
    dpix = oracle.forward(**s2).astype(np.float64) - pixels
    lhs = (dpix * gp).sum()
    rhs = (gc.astype(np.float64) * dcol).sum() + (gb.astype(np.float64) * dbg).sum()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs))
    assert (gv[..., 2] == 0).all()   # z gradient is never written (csrc/rasterise_grad_egl.cu:228-230)
    ids, _ = oracle.visibility(s['vertices'], s['faces'], pixels.shape[1], pixels.shape[2])
    np.testing.assert_array_equal(gb[ids < 0], gp[ids < 0])
    assert (gb[ids >= 0] == 0).all()


@pytest.mark.parametrize('channels', [1, 3])
def test_c_oracle_backward_matches_literal_numpy_restatement(oracle, channels):
    # the numpy version follows assemble_grads in the reference's buffer (y-up) coordinates, the C oracle works in
    # image coordinates: agreement checks every flip
    s = scenes.bent_square_scene(24, 20, channels=channels)
    pixels, ids = oracle.forward(**s, return_face_ids=True)
    gp = np.random.default_rng(1).standard_normal(pixels.shape).astype(np.float32)
    gb, gv, gc = oracle.backward(s['vertices'], s['faces'], pixels, gp)
    _, gbuf = oracle.visibility(s['vertices'], s['faces'], 20, 24)
    fv = np.where(ids[0][..., None] >= 0, s['faces'][0][np.maximum(ids[0], 0)], -1).astype(np.float32)
    gv2, gc2, gb2 = numpy_oracle.assemble_grads(s['vertices'][0], fv, gbuf[0], pixels[0], gp[0])
    np.testing.assert_allclose(gv[0], gv2, rtol=1e-4, atol=1e-4 * np.abs(gv2).max())
    np.testing.assert_allclose(gc[0], gc2, rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(gb[0], gb2.astype(np.float32))


def test_channel_groups_change_vertex_gradient_only(oracle):
    # C=4 -> groups {3,1}: grad_vertices is the sum of the per-group passes the reference would run
    # (dirt/rasterise_ops.py:86-108,163); colours/background are per channel and do not care
    s = scenes.bent_square_scene(24, 24, channels=4)
    pixels = oracle.forward(**s)
    gp = np.random.default_rng(2).standard_normal(pixels.shape).astype(np.float32)
    gb, gv, gc = oracle.backward(s['vertices'], s['faces'], pixels, gp)
    assert oracle.default_groups(4) == [3, 1] and oracle.default_groups(7) == [3, 3, 1]
    assert oracle.default_groups(2) == [1, 1] and oracle.default_groups(5) == [3, 1, 1]
    gb3, gv3, gc3 = oracle.backward(s['vertices'], s['faces'], pixels[..., :3].copy(), gp[..., :3].copy())
    gb1, gv1, gc1 = oracle.backward(s['vertices'], s['faces'], pixels[..., 3:].copy(), gp[..., 3:].copy())
    np.testing.assert_allclose(gv, gv3 + gv1, rtol=1e-5, atol=1e-5 * np.abs(gv).max())
    np.testing.assert_allclose(gc, np.concatenate([gc3, gc1], -1), rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(gb, np.concatenate([gb3, gb1], -1))


def test_degenerate_and_invalid_faces_are_ignored(oracle):
    s = scenes.square_scene(32, 32, 16, 16, 8)
    faces = np.concatenate([s['faces'], np.array([[[0, 0, 1], [0, 1, 99], [-1, 1, 2]]], np.int32)], axis=1)
    verts = s['vertices'].copy()
    pixels_ref = oracle.forward(**s)
    pixels = oracle.forward(s['background'], verts, s['vertex_colors'], faces)
    np.testing.assert_array_equal(pixels, pixels_ref)
    verts_nan = np.concatenate([verts, np.full((1, 1, 4), np.nan, np.float32)], axis=1)
    cols = np.concatenate([s['vertex_colors'], np.ones((1, 1, 1), np.float32)], axis=1)
    faces_nan = np.concatenate([s['faces'], np.array([[[0, 1, 4]]], np.int32)], axis=1)
    np.testing.assert_array_equal(oracle.forward(s['background'], verts_nan, cols, faces_nan), pixels_ref)
    empty = oracle.forward(s['background'], verts, s['vertex_colors'], np.zeros((1, 0, 3), np.int32))
    np.testing.assert_array_equal(empty, s['background'])


def test_behind_camera_faces_use_homogeneous_path(oracle):
    # a large ground-plane triangle with one vertex behind the camera (w < 0) must still cover the part of the
    # screen in front of the camera, with depth and perspective-correct barycentrics that stay finite
    H = W = 64
    verts = np.array([[-1.0, -0.5, 0.0, 1.0], [1.0, -0.5, 0.0, 1.0], [0.0, 2.0, -0.5, -0.5]], np.float32)
    faces = np.array([[0, 1, 2]], np.int32)
    ids, gbuf = oracle.visibility(verts[None], faces[None], H, W)
    cov = ids[0] == 0
    assert cov.any() and not cov.all()
    assert np.isfinite(gbuf[0][cov]).all()
    assert (gbuf[0][..., 3][cov] > 0).all()
    np.testing.assert_allclose(gbuf[0][..., :3][cov].sum(-1), 1.0, atol=1e-5)


def test_random_planar_triangulations_are_covered_exactly_once(oracle):
    # property test (hypothesis): a Delaunay-like fan triangulation of random points in the plane z=0, w=1 has
    # no overlaps, so every pixel is claimed by at most one face and the claimed set equals the exact rational rule
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=15, deadline=None)
    @given(st.integers(0, 10 ** 6))
    def check(seed):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(5, 9))
        angles = np.sort(rng.uniform(0, 2 * np.pi, n))
        radius = rng.uniform(0.3, 0.9, n)
        centre = rng.uniform(-0.1, 0.1, 2)
        ring = centre + np.stack([np.cos(angles), np.sin(angles)], 1) * radius[:, None]
        verts = np.concatenate([np.concatenate([[centre], ring]), np.zeros((n + 1, 1)), np.ones((n + 1, 1))], axis=1).astype(np.float32)
        faces = np.array([[0, 1 + i, 1 + (i + 1) % n] for i in range(n)], np.int32)
        # keep only fans whose consecutive angles are < pi apart (star-shaped, non-overlapping)
        gaps = np.diff(np.concatenate([angles, [angles[0] + 2 * np.pi]]))
        if gaps.max() >= np.pi * 0.95:
            return
        H, W = 24, 28
        ids, gbuf = oracle.visibility(verts[None], faces[None], H, W)
        masks = numpy_oracle.coverage_exact(verts, faces, H, W)
        count = np.sum(masks, axis=0)
        assert count.max() <= 1
        for f, m in enumerate(masks):
            np.testing.assert_array_equal(ids[0] == f, m)
        cov = ids[0] >= 0
        if cov.any():
            np.testing.assert_allclose(gbuf[0][..., :3][cov].sum(-1), 1.0, atol=1e-5)
            np.testing.assert_allclose(gbuf[0][..., 3][cov], 1.0, rtol=1e-6)

    check()


@pytest.mark.parametrize('threads', [1, 3, 64, 500])
def test_results_do_not_depend_on_the_thread_count(oracle, threads):
    """The oracle splits its work over images and, with fewer images than threads, over bands of rows (so that the CPU
    baseline uses every host thread).  Visibility, pixels and grad_background must be identical for every split; the
    vertex gradients are double-precision sums added up in a different grouping, rounded to fp32 once."""
    from dirt_b200 import scenes
    s = scenes.random_soup(batch=2, channels=4, seed=11)
    gp = None
    results = []
    before = oracle.threads()
    try:
        for t in (1, threads):
            oracle.set_threads(t)
            px, ids = oracle.forward(**s, return_face_ids=True)
            if gp is None:
                gp = np.random.default_rng(5).standard_normal(px.shape).astype(np.float32)
            results.append((px, ids) + tuple(oracle.backward(s['vertices'], s['faces'], px, gp)))
    finally:
        oracle.set_threads(before)
    (px0, ids0, gb0, gv0, gc0), (px1, ids1, gb1, gv1, gc1) = results
    np.testing.assert_array_equal(ids1, ids0)
    np.testing.assert_array_equal(px1, px0)
    np.testing.assert_array_equal(gb1, gb0)
    np.testing.assert_allclose(gv1, gv0, rtol=1e-6, atol=1e-6 * np.abs(gv0).max())
    np.testing.assert_allclose(gc1, gc0, rtol=1e-6, atol=1e-6 * np.abs(gc0).max())


def test_position_gradient_has_the_sign_and_size_of_a_moving_square(oracle):
    """The vertex-position gradient is an approximation by construction (Scharr edge filter + dilation, Appendix A), but
    for a rigid shift of a bright square over a dark background under a linear loss it must come out with the right sign
    and roughly the right size in BOTH axes -- this pins the y-up clip / row-down image convention of the gradient, which
    no reference test asserts.  dL/dpixel = column index: shifting the 16x16 square one pixel to the right raises L by
    16*16 (checked against the forward pass itself), i.e. dL/dx_clip = 256 * W/2 summed over the four vertices."""
    H = W = 64
    size, cx, cy = 16, 30.0, 32.0

    def verts(cx_, cy_):
        v = np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float64) * size - size / 2 + [cx_, cy_]
        v = v * 2 / np.array([W, H]) - 1
        return np.concatenate([v, np.zeros((4, 1)), np.ones((4, 1))], 1).astype(np.float32)[None]

    faces = np.array([[[0, 1, 2], [0, 2, 3]]], np.int32)
    cols = np.ones((1, 4, 1), np.float32)
    bg = np.zeros((1, H, W, 1), np.float32)
    ramp_x = np.broadcast_to(np.arange(W, dtype=np.float32)[None, None, :, None], (1, H, W, 1)).copy()
    ramp_y = np.broadcast_to(np.arange(H, dtype=np.float32)[None, :, None, None], (1, H, W, 1)).copy()

    def loss(cx_, cy_, ramp):
        return float((oracle.forward(bg, verts(cx_, cy_), cols, faces) * ramp).sum())

    px = oracle.forward(bg, verts(cx, cy), cols, faces)
    # the forward pass itself: one pixel to the right adds 256 to L; one pixel up in clip space (cy is measured in clip-y-up
    # pixels here) moves the square to SMALLER rows and takes 256 off a row-index loss
    assert (loss(cx + 1, cy, ramp_x) - loss(cx - 1, cy, ramp_x)) / 2 == 256.
    assert (loss(cx, cy + 1, ramp_y) - loss(cx, cy - 1, ramp_y)) / 2 == -256.
    _, gv, _ = oracle.backward(verts(cx, cy), faces, px, ramp_x)
    want = 256. * W / 2
    assert abs(gv[0, :, 0].sum() - want) < 0.1 * want and abs(gv[0, :, 1].sum()) < 0.15 * want
    _, gv, _ = oracle.backward(verts(cx, cy), faces, px, ramp_y)
    want = -256. * H / 2
    assert abs(gv[0, :, 1].sum() - want) < 0.1 * abs(want) and abs(gv[0, :, 0].sum()) < 0.15 * abs(want)
    assert (gv[..., 2] == 0).all()


# ---- the vectorised numpy rasteriser (oracle/numpy_raster.py): third implementation, CPU-baseline variant ------------

@pytest.mark.parametrize('name,kwargs', [
    ('square_scene', dict()),
    ('cylinder_scene', dict()),
    ('bent_square_scene', dict(channels=4)),
    ('bent_square_scene', dict(channels=1)),
    ('random_soup', dict(batch=2, width=61, height=45, n_faces=70, channels=3, seed=1, behind_camera=True)),
    ('config3', dict(batch=1, width=96, height=80, level=2, background='uniform')),
])
def test_numpy_rasteriser_agrees_with_the_c_oracle(oracle, name, kwargs):
    from oracle import numpy_raster as npr
    from conftest import rel_close
    s = getattr(scenes, name)(**kwargs)
    B = s['background'].shape[0]
    pixels_o, ids_o = oracle.forward(**s, return_face_ids=True)
    gp = np.random.default_rng(0).standard_normal(pixels_o.shape).astype(np.float32)
    gb_o, gv_o, gc_o = oracle.backward(s['vertices'], s['faces'], pixels_o, gp)
    for b in range(B):
        pixels, ids = npr.forward(s['background'][b], s['vertices'][b], s['vertex_colors'][b], s['faces'][b])[:2]
        np.testing.assert_array_equal(ids, ids_o[b])
        assert rel_close(pixels, pixels_o[b])[0]
        head = pixels_o[b + 1].reshape(-1, pixels_o.shape[-1])[:2] if b + 1 < B else None
        gb, gv, gc = npr.backward(s['vertices'][b], s['faces'][b], pixels_o[b], gp[b], None, head)
        np.testing.assert_array_equal(gb, gb_o[b])
        assert rel_close(gv, gv_o[b])[0] and rel_close(gc, gc_o[b])[0]


def test_numpy_rasteriser_square_golden():
    # BASELINE cfg1 through the numpy path: the reference's own acceptance check (tests/square_test.py:54-57)
    from oracle import numpy_raster as npr
    s = scenes.square_scene()
    pixels = npr.forward(s['background'][0], s['vertices'][0], s['vertex_colors'][0], s['faces'][0])[0]
    np.testing.assert_array_equal(pixels[:, :, 0], np.load(os.path.join(GOLDEN, 'square_test_expected.npy')))


def test_numpy_batch_driver_with_processes():
    from oracle import numpy_raster as npr
    s = scenes.config3(batch=3, width=64, height=48, level=1)
    gp = np.random.default_rng(1).standard_normal(s['background'].shape).astype(np.float32)
    one = npr.forward_backward_batch(s, gp, processes=1)
    two = npr.forward_backward_batch(s, gp, processes=2)
    for a, b in zip(one, two):
        np.testing.assert_array_equal(a, b)
