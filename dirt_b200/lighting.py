"""Mesh normals and simple reflectance models (torch); mirrors the API of dirt/lighting.py."""
import torch


def _prepare(vertices, faces):
    vertices = vertices if isinstance(vertices, torch.Tensor) else torch.as_tensor(vertices, dtype=torch.float32)
    faces = faces if isinstance(faces, torch.Tensor) else torch.as_tensor(faces)
    return vertices, faces.to(device=vertices.device, dtype=torch.long)


def _face_normals(vertices, faces):
    corners = vertices[..., faces, :]  # [*, F, 3, 3]
    n = torch.linalg.cross(corners[..., 1, :] - corners[..., 0, :], corners[..., 2, :] - corners[..., 0, :], dim=-1)
    return n / (torch.linalg.norm(n, dim=-1, keepdim=True) + 1.e-12)


def vertex_normals(vertices, faces, name=None):
    """Per-vertex normals: normalised sum of the unit normals of the faces using each vertex
    (dirt/lighting.py:34-93).  vertices [*,V,3|4], faces [F,3] -> [*,V,3]."""
    vertices, faces = _prepare(vertices, faces)
    vertices = vertices[..., :3]
    normals_by_face = _face_normals(vertices, faces)  # [*, F, 3]
    summed = torch.zeros_like(vertices)
    for corner in range(3):
        summed = summed.index_add(-2, faces[:, corner], normals_by_face)
    return summed / (torch.linalg.norm(summed, dim=-1, keepdim=True) + 1.e-12)


def vertex_normals_pre_split(vertices, faces, name=None, static=False):
    """As `vertex_normals` for meshes where every vertex belongs to exactly one face
    (dirt/lighting.py:101-133)."""
    vertices, faces = _prepare(vertices, faces)
    vertices = vertices[..., :3]
    normals_by_face = _face_normals(vertices, faces)
    out = torch.zeros_like(vertices)
    for corner in range(3):
        out = out.index_copy(-2, faces[:, corner], normals_by_face)
    return out


def split_vertices_by_face(vertices, faces, name=None):
    """Duplicates vertices so that each is used by one face: returns (new_vertices [*,3F,D], new_faces [F,3])
    (dirt/lighting.py:136-179)."""
    vertices, faces = _prepare(vertices, faces)
    new_vertices = vertices[..., faces.reshape(-1), :]
    new_faces = torch.arange(faces.shape[0] * 3, device=vertices.device, dtype=torch.int32).reshape(-1, 3)
    return new_vertices, new_faces


def _cosine_term(cosines, double_sided):
    return cosines.abs() if double_sided else cosines.clamp(min=0.)


def diffuse_directional(vertex_normals, vertex_colors, light_direction, light_color, double_sided=True, name=None):
    """Lambertian reflectance under one directional light (dirt/lighting.py:182-225)."""
    vertex_normals = torch.as_tensor(vertex_normals, dtype=torch.float32)
    dev = vertex_normals.device
    vertex_colors = torch.as_tensor(vertex_colors, dtype=torch.float32, device=dev)
    light_direction = torch.as_tensor(light_direction, dtype=torch.float32, device=dev)
    light_color = torch.as_tensor(light_color, dtype=torch.float32, device=dev)
    cosines = torch.matmul(vertex_normals, -light_direction[..., None])  # [*, V, 1]
    return light_color[..., None, :] * vertex_colors * _cosine_term(cosines, double_sided)


def specular_directional(vertex_positions, vertex_normals, vertex_reflectivities, light_direction, light_color,
                         camera_position, shininess, double_sided=True, name=None):
    """Phong reflectance under one directional light (dirt/lighting.py:228-288)."""
    vertex_positions = torch.as_tensor(vertex_positions, dtype=torch.float32)
    dev = vertex_positions.device
    vertex_normals = torch.as_tensor(vertex_normals, dtype=torch.float32, device=dev)
    vertex_reflectivities = torch.as_tensor(vertex_reflectivities, dtype=torch.float32, device=dev)
    light_direction = torch.as_tensor(light_direction, dtype=torch.float32, device=dev)
    light_color = torch.as_tensor(light_color, dtype=torch.float32, device=dev)
    camera_position = torch.as_tensor(camera_position, dtype=torch.float32, device=dev)
    shininess = torch.as_tensor(shininess, dtype=torch.float32, device=dev)
    to_light = -light_direction
    # (the reference adds `-to_light` of shape [*,3] to a [*,V,3] tensor, which only broadcasts without batch dimensions,
    # dirt/lighting.py:272; the vertex axis is inserted here so that batches work as its docstring promises)
    reflected = -to_light[..., None, :] + 2. * torch.matmul(vertex_normals, to_light[..., None]) * vertex_normals
    to_camera = camera_position[..., None, :] - vertex_positions
    # the reference adds its epsilon after the division (dirt/lighting.py:279); kept for parity
    cosines = ((to_camera / torch.linalg.norm(to_camera, dim=-1, keepdim=True) + 1.e-12) * reflected).sum(-1, keepdim=True)
    return light_color[..., None, :] * vertex_reflectivities * torch.pow(_cosine_term(cosines, double_sided),
                                                                         shininess[..., None, None])


def diffuse_point(vertex_positions, vertex_normals, vertex_colors, light_position, light_color, double_sided=True, name=None):
    """Lambertian reflectance under one point light (dirt/lighting.py:291-343)."""
    vertex_positions = torch.as_tensor(vertex_positions, dtype=torch.float32)
    dev = vertex_positions.device
    vertex_normals = torch.as_tensor(vertex_normals, dtype=torch.float32, device=dev)
    vertex_colors = torch.as_tensor(vertex_colors, dtype=torch.float32, device=dev)
    light_position = torch.as_tensor(light_position, dtype=torch.float32, device=dev)
    light_color = torch.as_tensor(light_color, dtype=torch.float32, device=dev)
    relative = vertex_positions - light_position[..., None, :]
    incident = relative / (torch.linalg.norm(relative, dim=-1, keepdim=True) + 1.e-12)
    cosines = (vertex_normals * incident).sum(-1)
    return light_color[..., None, :] * vertex_colors * _cosine_term(cosines, double_sided)[..., None]
