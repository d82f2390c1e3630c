"""Deferred shading with per-pixel lighting -- the scene of the reference's samples/deferred.py (cube with per-face
normals; G-buffer of [mask, world position, albedo, normal] = 10 channels; ambient + Lambertian + Phong terms evaluated
per pixel in the shader function), with torch in place of TensorFlow.

    python examples/deferred.py [out.png]

`shader_fn` is plain torch (tests/test_helpers_cpu.py runs it on the CPU on a G-buffer rendered by the oracle); only
`main()` needs the GPU.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

frame_width, frame_height = 640, 480


def build_cube():
    vertices = [[x, y, z] for z in [-1, 1] for y in [-1, 1] for x in [-1, 1]]
    quads = [[0, 1, 3, 2], [4, 5, 7, 6], [1, 5, 4, 0], [2, 6, 7, 3], [4, 6, 2, 0], [3, 7, 5, 1]]
    triangles = sum([[[a, b, c], [c, d, a]] for [a, b, c, d] in quads], [])
    return vertices, triangles


def shader_fn(gbuffer, view_matrix, light_direction):
    """G-buffer [H, W, 10] = (mask, position xyz, albedo rgb, normal xyz) -> shaded pixels [H, W, 3]
    (samples/deferred.py:61-104)."""
    from dirt_b200 import lighting
    dev = gbuffer.device
    mask, positions, unlit, normals = gbuffer[..., :1], gbuffer[..., 1:4], gbuffer[..., 4:7], gbuffer[..., 7:]
    ambient = unlit * 0.2
    diffuse = lighting.diffuse_directional(normals.reshape(-1, 3), unlit.reshape(-1, 3), light_direction,
                                           light_color=torch.tensor([1., 0., 0.], device=dev), double_sided=False)
    camera_position_world = torch.linalg.inv(view_matrix)[3, :3]
    specular = lighting.specular_directional(positions.reshape(-1, 3), normals.reshape(-1, 3), unlit.reshape(-1, 3),
                                             light_direction, light_color=torch.ones(3, device=dev),
                                             camera_position=camera_position_world, shininess=6., double_sided=False)
    lit = diffuse.reshape(unlit.shape) + specular.reshape(unlit.shape) + ambient
    background = torch.tensor([0., 0., 0.3], device=dev)
    # clipped: the specular term saturates some pixels
    return (lit * mask + background * (1. - mask)).clamp(0., 1.)


def scene(device):
    """Clip-space vertices, G-buffer attributes, faces, view matrix of the sample (samples/deferred.py:25-58,112-126)."""
    from dirt_b200 import lighting, matrices
    vertices_object, faces = build_cube()
    vertices_object = torch.tensor(vertices_object, dtype=torch.float32, device=device)
    faces = torch.tensor(faces, dtype=torch.int32, device=device)
    vertices_object, faces = lighting.split_vertices_by_face(vertices_object, faces)   # per-face normals
    colors = torch.ones_like(vertices_object)
    vertices_object = torch.cat([vertices_object, torch.ones_like(vertices_object[:, -1:])], dim=1).requires_grad_(True)
    vertices_world = vertices_object @ matrices.rodrigues([0., 0.5, 0.]).to(device)
    normals_world = lighting.vertex_normals_pre_split(vertices_world, faces)
    view_matrix = matrices.compose(matrices.translation([0., -1.5, -3.5]), matrices.rodrigues([-0.3, 0., 0.])).to(device)
    projection_matrix = matrices.perspective_projection(near=0.1, far=20., right=0.1, aspect=float(frame_height) / frame_width).to(device)
    vertices_clip = vertices_world @ view_matrix @ projection_matrix
    attributes = torch.cat([torch.ones_like(vertices_object[:, :1]), vertices_world[:, :3], colors, normals_world], dim=1)
    return vertices_object, vertices_clip, attributes, faces, view_matrix


def main():
    import dirt_b200 as dirt
    device = torch.device('cuda')
    vertices_object, vertices_clip, attributes, faces, view_matrix = scene(device)
    light_direction = torch.nn.functional.normalize(torch.tensor([1., -0.3, -0.5], device=device), dim=0).requires_grad_(True)
    # anything the shader needs gradients for goes through shader_additional_inputs (here: the light direction)
    pixels = dirt.rasterise_deferred(
        background_attributes=torch.zeros([frame_height, frame_width, 10], device=device),
        vertices=vertices_clip, vertex_attributes=attributes, faces=faces,
        shader_fn=shader_fn, shader_additional_inputs=[view_matrix, light_direction])

    loss = -pixels.mean()
    loss.backward()
    print('rendered %dx%d; |d loss / d vertices| max = %.3e, d loss / d light = %s'
          % (frame_width, frame_height, float(vertices_object.grad.abs().max()), [round(float(x), 5) for x in light_direction.grad]))

    out = sys.argv[1] if len(sys.argv) > 1 else None
    if out:
        try:
            import cv2
            cv2.imwrite(out, (pixels.detach() * 255).byte().cpu().numpy()[:, :, ::-1])
            print('wrote', out)
        except ImportError:
            print('cv2 not available; image not written')


if __name__ == '__main__':
    main()
