"""CPU tests of the helpers on either side of the hot path: dirt_b200.matrices / lighting /
projection against independent restatements of the reference formulas (plain numpy loops on small cases, scipy for the
rotation), each citing the reference lines it follows.  The reference package itself needs TensorFlow 1.x and cannot be
imported here."""
import numpy as np
import pytest
import torch

from dirt_b200 import lighting, matrices, projection, scenes


def _rng(seed):
    return np.random.default_rng(seed)


# ---- matrices (dirt/matrices.py) ------------------------------------------------------------------------------------

def test_rodrigues_matches_the_reference_layout():
    """dirt/matrices.py:15-61: Rodrigues' formula with K laid out exactly as the reference writes it (:43-48, "follows the
    OpenCV docs' definition"): the result equals the usual column-vector rotation matrix element for element, so used on
    row vectors (`v @ R`, the module's convention) it turns them by MINUS the angle.  Batched, 4x4 by default."""
    from scipy.spatial.transform import Rotation
    vecs = _rng(0).standard_normal((5, 7, 3)).astype(np.float32)
    ours = matrices.rodrigues(torch.from_numpy(vecs)).numpy()
    assert ours.shape == (5, 7, 4, 4)
    want = Rotation.from_rotvec(vecs.reshape(-1, 3).astype(np.float64)).as_matrix().reshape(5, 7, 3, 3)
    np.testing.assert_allclose(ours[..., :3, :3], want, atol=2e-6)
    np.testing.assert_array_equal(ours[..., 3, :], np.broadcast_to([0., 0., 0., 1.], (5, 7, 4)))
    np.testing.assert_array_equal(ours[..., :3, 3], np.zeros((5, 7, 3)))
    np.testing.assert_allclose(matrices.rodrigues(torch.from_numpy(vecs), three_by_three=True).numpy(), ours[..., :3, :3])
    # the literal K of the reference for one vector (:43-47), indexed [in][out]
    x, y, z = 0.3, -0.5, 0.8
    n = np.sqrt(x * x + y * y + z * z)
    a = np.array([x, y, z]) / n
    K = np.array([[0., -a[2], a[1]], [a[2], 0., -a[0]], [-a[1], a[0], 0.]])
    lit = np.cos(n) * np.eye(3) + (1 - np.cos(n)) * np.outer(a, a) + np.sin(n) * K
    np.testing.assert_allclose(matrices.rodrigues([x, y, z], three_by_three=True).numpy(), lit, atol=1e-6)
    # the zero vector is the identity and differentiable (the reference adds 1e-12 for that, :36-38)
    z = torch.zeros(3, requires_grad=True)
    r = matrices.rodrigues(z)
    np.testing.assert_allclose(r.detach().numpy(), np.eye(4), atol=1e-6)
    r.sum().backward()
    assert torch.isfinite(z.grad).all()


def test_translation_scale_compose_act_on_row_vectors():
    """dirt/matrices.py:64-107,183-207: points are rows, `compose(a, b)` applies a first."""
    p = np.array([[1., 2., 3., 1.]], np.float32)
    t = matrices.translation([10., 20., 30.]).numpy()
    s = matrices.scale([2., 3., 4.]).numpy()
    np.testing.assert_allclose(p @ t, [[11., 22., 33., 1.]])
    np.testing.assert_allclose(p @ s, [[2., 6., 12., 1.]])
    np.testing.assert_allclose(p @ matrices.compose(t, s).numpy(), (p @ t) @ s)
    np.testing.assert_allclose(p @ matrices.compose(s, t).numpy(), (p @ s) @ t)
    batch = matrices.translation(torch.from_numpy(_rng(1).standard_normal((4, 3)).astype(np.float32)))
    assert batch.shape == (4, 4, 4)
    np.testing.assert_allclose(matrices.pad_3x3_to_4x4(np.eye(3, dtype=np.float32) * 2.).numpy(), np.diag([2., 2., 2., 1.]))


def test_perspective_projection_maps_the_frustum_to_the_ndc_cube():
    """dirt/matrices.py:110-153: OpenGL convention, camera looks along -z, `aspect` = height / width."""
    near, far, right, aspect = 0.5, 40., 0.25, 0.75
    m = matrices.perspective_projection(near, far, right, aspect).numpy()
    top = right * aspect

    def ndc(p):
        c = np.array(list(p) + [1.], np.float64) @ m.astype(np.float64)
        return c[:3] / c[3]

    np.testing.assert_allclose(ndc([right, top, -near]), [1., 1., -1.], atol=1e-6)
    np.testing.assert_allclose(ndc([-right, -top, -near]), [-1., -1., -1.], atol=1e-6)
    np.testing.assert_allclose(ndc([0., 0., -far]), [0., 0., 1.], atol=1e-5)
    s = far / near
    np.testing.assert_allclose(ndc([right * s, top * s, -far]), [1., 1., 1.], atol=1e-5)
    # clip-space w is the view-space distance along the viewing direction
    assert (np.array([0.3, -0.2, -7., 1.]) @ m)[3] == pytest.approx(7.)
    np.testing.assert_allclose(m, scenes.perspective_projection(near, far, right, aspect), atol=1e-7)
    assert matrices.perspective_projection(torch.tensor([0.1, 0.2]), 20., 0.1, 1.).shape == (2, 4, 4)


# ---- lighting (dirt/lighting.py) ------------------------------------------------------------------------------------

def _face_normal(v, f):
    n = np.cross(v[f[1]] - v[f[0]], v[f[2]] - v[f[0]])
    return n / (np.linalg.norm(n) + 1.e-12)


def _vertex_normals_loops(v, faces):
    """dirt/lighting.py:24-31,34-93: every face adds its UNIT normal to its three vertices; the sum is renormalised."""
    out = np.zeros((v.shape[0], 3))
    for f in faces:
        n = _face_normal(v[:, :3].astype(np.float64), f)
        for k in range(3):
            out[f[k]] += n
    return out / (np.linalg.norm(out, axis=-1, keepdims=True) + 1.e-12)


@pytest.mark.parametrize('with_w', [False, True])
def test_vertex_normals_match_the_loop_restatement(with_w):
    v, f = scenes.icosphere(1)
    r = _rng(2)
    v = (v * (1. + 0.3 * r.standard_normal((v.shape[0], 1)))).astype(np.float32)   # not a sphere any more
    if with_w:
        v = np.concatenate([v, np.ones((v.shape[0], 1), np.float32)], axis=1)      # the w coordinate is dropped (:59)
    got = lighting.vertex_normals(torch.from_numpy(v), torch.from_numpy(f)).numpy()
    np.testing.assert_allclose(got, _vertex_normals_loops(v, f), atol=2e-6)
    np.testing.assert_allclose(np.linalg.norm(got, axis=-1), 1., atol=1e-5)
    # batched meshes with common topology (:37-38)
    vb = np.stack([v, v[:, ::-1].copy() if not with_w else v * 1.5])
    gb = lighting.vertex_normals(torch.from_numpy(vb), torch.from_numpy(f)).numpy()
    for i in range(2):
        np.testing.assert_allclose(gb[i], _vertex_normals_loops(vb[i], f), atol=2e-6)


def test_split_vertices_and_pre_split_normals():
    """dirt/lighting.py:101-179: after splitting, every vertex belongs to one face and carries that face's unit normal."""
    v, f = scenes.cube()
    sv, sf = lighting.split_vertices_by_face(torch.from_numpy(v), torch.from_numpy(f))
    assert sv.shape == (f.shape[0] * 3, 3) and sf.dtype == torch.int32
    np.testing.assert_array_equal(sf.numpy(), np.arange(f.shape[0] * 3).reshape(-1, 3))
    np.testing.assert_array_equal(sv.numpy(), v[f.reshape(-1)])
    n = lighting.vertex_normals_pre_split(sv, sf).numpy()
    for i, face in enumerate(f):
        want = _face_normal(v.astype(np.float64), face)
        for k in range(3):
            np.testing.assert_allclose(n[3 * i + k], want, atol=1e-6)
    # the general routine gives the same answer on a split mesh
    np.testing.assert_allclose(lighting.vertex_normals(sv, sf).numpy(), n, atol=1e-6)


@pytest.mark.parametrize('double_sided', [True, False])
def test_reflectance_models_match_the_loop_restatements(double_sided):
    """dirt/lighting.py:182-225 (Lambert, directional), :228-288 (Phong, directional; the epsilon is added AFTER the
    normalising division, :279), :291-343 (Lambert, point light)."""
    r = _rng(3)
    V, C = 9, 3
    pos = r.standard_normal((2, V, 3)).astype(np.float32)
    nrm = r.standard_normal((2, V, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    col = r.uniform(size=(2, V, C)).astype(np.float32)
    ldir = r.standard_normal((2, 3)).astype(np.float32)
    ldir /= np.linalg.norm(ldir, axis=-1, keepdims=True)
    lcol = r.uniform(size=(2, C)).astype(np.float32)
    lpos = r.standard_normal((2, 3)).astype(np.float32) * 3.
    cam = r.standard_normal((2, 3)).astype(np.float32) * 4.
    shin = np.array([3., 8.], np.float32)
    side = (lambda c: abs(c)) if double_sided else (lambda c: max(c, 0.))

    want_dd = np.zeros((2, V, C)); want_sp = np.zeros((2, V, C)); want_dp = np.zeros((2, V, C))
    for b in range(2):
        for i in range(V):
            n, p = nrm[b, i].astype(np.float64), pos[b, i].astype(np.float64)
            want_dd[b, i] = lcol[b] * col[b, i] * side(float(n @ -ldir[b]))
            to_light = -ldir[b].astype(np.float64)
            reflected = -to_light + 2. * (n @ to_light) * n
            to_cam = cam[b] - p
            want_sp[b, i] = lcol[b] * col[b, i] * side(float((to_cam / np.linalg.norm(to_cam) + 1.e-12) @ reflected)) ** shin[b]
            rel = p - lpos[b]
            want_dp[b, i] = lcol[b] * col[b, i] * side(float(n @ (rel / (np.linalg.norm(rel) + 1.e-12))))
    t = torch.from_numpy
    np.testing.assert_allclose(lighting.diffuse_directional(t(nrm), t(col), t(ldir), t(lcol), double_sided).numpy(), want_dd, atol=2e-6)
    np.testing.assert_allclose(lighting.specular_directional(t(pos), t(nrm), t(col), t(ldir), t(lcol), t(cam), t(shin), double_sided).numpy(),
                               want_sp, rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(lighting.diffuse_point(t(pos), t(nrm), t(col), t(lpos), t(lcol), double_sided).numpy(), want_dp, atol=2e-6)


# ---- projection (dirt/projection.py) --------------------------------------------------------------------------------

def test_unprojected_rays_pass_through_the_points_that_project_to_the_pixels():
    """dirt/projection.py:6-70: pixel (x, y) with y down -> NDC with y up; ray start on the near plane, delta towards the
    z_ndc = 0 surface.  A world point must lie on the ray of the pixel it projects to."""
    W, H = 64, 48
    view = matrices.compose(matrices.rodrigues([0.1, -0.4, 0.05]), matrices.translation([0.2, -0.1, -5.]))
    proj = matrices.perspective_projection(0.1, 30., 0.08, H / W)
    world_to_clip = matrices.compose(view, proj).numpy().astype(np.float64)
    clip_to_world = np.linalg.inv(world_to_clip)
    pts = _rng(4).uniform(-1., 1., size=(11, 3))
    clip = np.concatenate([pts, np.ones((11, 1))], axis=1) @ world_to_clip
    ndc = clip[:, :3] / clip[:, 3:]
    assert (np.abs(ndc) < 1.).all()
    pix = np.stack([(ndc[:, 0] + 1.) * 0.5 * W, (1. - ndc[:, 1]) * 0.5 * H], axis=1)   # y down, as rasterise lays rows out
    starts, deltas = projection.unproject_pixels_to_rays(pix.astype(np.float32), clip_to_world.astype(np.float32), [W, H])
    starts, deltas = starts.numpy().astype(np.float64), deltas.numpy().astype(np.float64)
    for i in range(11):
        d = deltas[i] / np.linalg.norm(deltas[i])
        off = pts[i] - starts[i]
        assert np.linalg.norm(off - (off @ d) * d) < 2e-3          # the point is on the ray
        assert off @ d > 0                                           # in front of the near plane
    # ray starts are on the near plane: they project to z_ndc = -1
    sc = np.concatenate([starts, np.ones((11, 1))], axis=1) @ world_to_clip
    np.testing.assert_allclose(sc[:, 2] / sc[:, 3], -1., atol=2e-3)
    # leading image dimensions (A*) with their own matrices and sizes, trailing pixel dimensions (B*)
    s2, d2 = projection.unproject_pixels_to_rays(np.tile(pix.astype(np.float32)[None, None], (2, 3, 1, 1)).reshape(2, 3, 11, 2),
                                                 np.tile(clip_to_world.astype(np.float32)[None], (2, 1, 1)), [[W, H], [W, H]])
    assert s2.shape == (2, 3, 11, 3) and d2.shape == (2, 3, 11, 3)
    np.testing.assert_allclose(s2[1, 2].numpy(), starts, atol=1e-4)


# ---- texture lookup of examples/textured.py (samples/textured.py:15-60) ----------------------------------------------

def _load_textured_example():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples', 'textured.py')
    spec = importlib.util.spec_from_file_location('textured_example', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_texture_lookup_of_the_textured_example():
    ex = _load_textured_example()
    r = _rng(6)
    tex = r.uniform(size=(5, 7, 3)).astype(np.float32)
    uv = r.uniform(-1.5, 2.5, size=(4, 6, 2)).astype(np.float32)
    idx = ex.uvs_to_pixel_indices(torch.from_numpy(uv), tex.shape[:2]).numpy()
    # (u, v) -> (row, column) = (v mod 1 * H, u mod 1 * W): u = v = 0 is the top-left corner
    np.testing.assert_allclose(idx[..., 0], (uv[..., 1] % 1.) * 5, atol=1e-5)
    np.testing.assert_allclose(idx[..., 1], (uv[..., 0] % 1.) * 7, atol=1e-5)
    clamped = ex.uvs_to_pixel_indices(torch.from_numpy(uv), tex.shape[:2], mode='clamp').numpy()
    np.testing.assert_allclose(clamped[..., 0], np.clip(uv[..., 1], 0, 1) * 5, atol=1e-5)
    got = ex.sample_texture(torch.from_numpy(tex), torch.from_numpy(idx)).numpy()
    want = np.zeros(idx.shape[:-1] + (3,))
    for i in np.ndindex(idx.shape[:-1]):
        rr, cc = idx[i]
        r0, c0 = int(np.floor(rr)), int(np.floor(cc))
        fr, fc = rr - r0, cc - c0
        t = lambda a, b: tex[a % 5, b % 7].astype(np.float64)
        want[i] = t(r0, c0) * (1 - fc) * (1 - fr) + t(r0, c0 + 1) * fc * (1 - fr) + t(r0 + 1, c0) * (1 - fc) * fr + t(r0 + 1, c0 + 1) * fc * fr
    np.testing.assert_allclose(got, want, atol=1e-5)
    # integer indices return the texel itself; the lookup is differentiable w.r.t. the texture
    np.testing.assert_allclose(ex.sample_texture(torch.from_numpy(tex), torch.tensor([[2., 3.]])).numpy()[0], tex[2, 3], atol=1e-6)
    t_param = torch.from_numpy(tex).requires_grad_(True)
    ex.sample_texture(t_param, torch.from_numpy(idx)).sum().backward()
    assert float(t_param.grad.sum()) == pytest.approx(idx[..., 0].size * 3, rel=1e-4)   # bilinear weights sum to one


def test_shader_of_the_textured_example_on_a_synthetic_gbuffer():
    ex = _load_textured_example()
    tex = ex.checker_texture(size=32, squares=4)
    assert tex.shape == (32, 32, 3) and float(tex.min()) >= 0. and float(tex.max()) <= 1.
    g = torch.zeros(4, 5, 6)
    g[1:3, 1:4, 0] = 1.                                    # covered block
    g[..., 1:3] = torch.rand(4, 5, 2, generator=torch.Generator().manual_seed(0))
    g[..., 3:] = torch.tensor([0., 0., 1.])
    out = ex.shader_fn(g, tex, torch.tensor([0., 0., -1.]))
    assert out.shape == (4, 5, 3)
    np.testing.assert_allclose(out[0, 0].numpy(), [0., 0., 0.3], atol=1e-6)              # background colour where mask = 0
    unlit = ex.sample_texture(tex, ex.uvs_to_pixel_indices(g[..., 1:3], tex.shape[:2]))
    np.testing.assert_allclose(out[1, 1].numpy(), (unlit[1, 1] * (0.4 + 0.6)).numpy(), atol=1e-5)   # normal faces the light
    verts, uvs, faces = ex.build_cube()
    assert len(verts) == 24 and len(uvs) == 24 and len(faces) == 12


def test_deferred_example_shades_an_oracle_gbuffer():
    """examples/deferred.py (samples/deferred.py): the scene and the per-pixel shader, with the CPU oracle standing in
    for the rasteriser; the shaded image must equal shading evaluated pixel by pixel from the formulas."""
    import importlib.util
    import os
    from oracle import oracle
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples', 'deferred.py')
    spec = importlib.util.spec_from_file_location('deferred_example', path)
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    ex.frame_width, ex.frame_height = 80, 60
    vo, vc, attrs, faces, view = ex.scene(torch.device('cpu'))
    assert vc.shape == (36, 4) and attrs.shape == (36, 10) and faces.shape == (12, 3)
    g = oracle.forward(np.zeros((1, 60, 80, 10), np.float32), vc.detach().numpy()[None], attrs.detach().numpy()[None], faces.numpy()[None])[0]
    light = torch.nn.functional.normalize(torch.tensor([1., -0.3, -0.5]), dim=0)
    img = ex.shader_fn(torch.from_numpy(g), view, light).numpy()
    assert img.shape == (60, 80, 3) and img.min() >= 0. and img.max() <= 1.
    covered = g[..., 0] > 0.5
    assert 0.05 < covered.mean() < 0.6
    np.testing.assert_allclose(img[~covered], np.broadcast_to([0., 0., 0.3], img[~covered].shape), atol=1e-6)
    cam = np.linalg.inv(view.numpy().astype(np.float64))[3, :3]
    l = light.numpy().astype(np.float64)
    rows, cols = np.nonzero(covered)
    for r, c in list(zip(rows, cols))[::37]:
        m, p, a, n = g[r, c, 0], g[r, c, 1:4].astype(np.float64), g[r, c, 4:7].astype(np.float64), g[r, c, 7:].astype(np.float64)
        diffuse = np.array([1., 0., 0.]) * a * max(float(n @ -l), 0.)
        refl = l + 2. * (n @ -l) * n
        to_cam = cam - p
        spec = a * max(float((to_cam / np.linalg.norm(to_cam) + 1e-12) @ refl), 0.) ** 6.
        want = np.clip((diffuse + spec + 0.2 * a) * m + np.array([0., 0., 0.3]) * (1. - m), 0., 1.)
        np.testing.assert_allclose(img[r, c], want, atol=2e-5)


def test_cube_batch_scene_is_the_sample_cube_at_other_orientations():
    # the bench's large-face workload: same mesh, camera and shading as cube_scene (samples/simple.py:28-74)
    from dirt_b200 import scenes
    one, many = scenes.cube_scene(64, 48), scenes.cube_batch(batch=3, width=64, height=48, seed=2)
    assert many['vertices'].shape == (3,) + one['vertices'].shape[1:] and many['faces'].shape == (3, 12, 3)
    assert (many['faces'][0] == one['faces'][0]).all() and many['background'].shape == (3, 48, 64, 3)
    assert not np.allclose(many['vertices'][0], many['vertices'][1])          # a pose per item
    w = many['vertices'][..., 3]
    assert (w > 0).all()                                                      # the whole cube in front of the camera
    ndc = many['vertices'][..., :2] / w[..., None]
    assert np.abs(ndc).max() < 1.0                                            # and inside the frame
