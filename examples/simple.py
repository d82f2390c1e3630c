"""Renders a lit cube with dirt_b200 and takes one gradient step's worth of gradients -- the same scene as the
reference's samples/simple.py (cube, Lambertian lighting, perspective camera), with torch in place of TensorFlow.

    python examples/simple.py [out.png]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import dirt_b200 as dirt  # noqa: E402
from dirt_b200 import lighting, matrices  # noqa: E402

frame_width, frame_height = 640, 480


def build_cube():
    vertices = [[x, y, z] for z in [-1, 1] for y in [-1, 1] for x in [-1, 1]]
    quads = [[0, 1, 3, 2], [4, 5, 7, 6], [1, 5, 4, 0], [2, 6, 7, 3], [4, 6, 2, 0], [3, 7, 5, 1]]
    triangles = sum([[[a, b, c], [c, d, a]] for [a, b, c, d] in quads], [])
    return vertices, triangles


def main():
    device = torch.device('cuda')
    cube_vertices_object, cube_faces = build_cube()
    cube_vertices_object = torch.tensor(cube_vertices_object, dtype=torch.float32, device=device)
    cube_faces = torch.tensor(cube_faces, dtype=torch.int32, device=device)
    # replicate shared vertices so that normals are per face
    cube_vertices_object, cube_faces = lighting.split_vertices_by_face(cube_vertices_object, cube_faces)
    cube_vertex_colors = torch.ones_like(cube_vertices_object)
    cube_vertices_object = torch.cat([cube_vertices_object, torch.ones_like(cube_vertices_object[:, -1:])], dim=1)
    cube_vertices_object.requires_grad_(True)

    cube_vertices_world = cube_vertices_object @ matrices.rodrigues([0., 0.5, 0.]).to(device)
    cube_normals_world = lighting.vertex_normals_pre_split(cube_vertices_world, cube_faces)
    view_matrix = matrices.compose(matrices.translation([0., -1.5, -3.5]), matrices.rodrigues([-0.3, 0., 0.])).to(device)
    cube_vertices_camera = cube_vertices_world @ view_matrix
    projection_matrix = matrices.perspective_projection(near=0.1, far=20., right=0.1, aspect=float(frame_height) / frame_width).to(device)
    cube_vertices_clip = cube_vertices_camera @ projection_matrix

    vertex_colors_lit = lighting.diffuse_directional(
        cube_normals_world, cube_vertex_colors,
        light_direction=torch.tensor([1., 0., 0.], device=device), light_color=torch.tensor([1., 1., 1.], device=device)
    ) * 0.8 + cube_vertex_colors * 0.2

    pixels = dirt.rasterise(
        vertices=cube_vertices_clip, faces=cube_faces, vertex_colors=vertex_colors_lit,
        background=torch.zeros([frame_height, frame_width, 3], device=device),
        width=frame_width, height=frame_height, channels=3)

    # a toy loss: make the image brighter; gradients flow back to the object-space vertices
    loss = -pixels.mean()
    loss.backward()
    print('rendered %dx%d, %.1f%% of the pixels covered, |d loss / d vertices| max = %.3e'
          % (frame_width, frame_height, 100. * float((pixels.sum(-1) > 0).float().mean()), float(cube_vertices_object.grad.abs().max())))

    out = sys.argv[1] if len(sys.argv) > 1 else None
    if out:
        try:
            import cv2
            cv2.imwrite(out, (pixels.detach().clamp(0, 1) * 255).byte().cpu().numpy()[:, :, ::-1])
            print('wrote', out)
        except ImportError:
            print('cv2 not available; image not written')


if __name__ == '__main__':
    main()
