"""Pixel -> world-space ray unprojection (torch); mirrors dirt/projection.py:22-70."""
import torch


def _pixel_to_ndc(pixel_locations, image_size):
    flip = torch.tensor([1., -1.], dtype=pixel_locations.dtype, device=pixel_locations.device)
    return (-1. + 2. * pixel_locations / image_size) * flip


def _unproject_ndc_to_world(x_ndc, clip_to_world_matrix):
    homogeneous = torch.cat([x_ndc, torch.ones_like(x_ndc[..., :1])], dim=-1)
    x_world_scaled = torch.matmul(homogeneous[..., None, :], clip_to_world_matrix)[..., 0, :]
    return x_world_scaled[..., :3] / x_world_scaled[..., 3:]


def unproject_pixels_to_rays(pixel_locations, clip_to_world_matrix, image_size, name=None):
    """pixel_locations [A*, B*, 2] (x,y in pixels), clip_to_world_matrix [A*,4,4], image_size [A*,2] (width,height)
    -> (ray starts on the near plane [A*,B*,3], unnormalised ray directions [A*,B*,3])."""
    pixel_locations = torch.as_tensor(pixel_locations, dtype=torch.float32)
    dev = pixel_locations.device
    clip_to_world_matrix = torch.as_tensor(clip_to_world_matrix, dtype=torch.float32, device=dev)
    image_size = torch.as_tensor(image_size, device=dev).to(torch.float32)
    per_iib_dims = pixel_locations.dim() - image_size.dim()
    image_size = image_size.reshape(image_size.shape[:-1] + (1,) * per_iib_dims + (2,))
    clip_to_world_matrix = clip_to_world_matrix.reshape(clip_to_world_matrix.shape[:-2] + (1,) * per_iib_dims + (4, 4))
    ndc = _pixel_to_ndc(pixel_locations, image_size)
    starts = _unproject_ndc_to_world(torch.cat([ndc, -torch.ones_like(ndc[..., :1])], dim=-1), clip_to_world_matrix)
    deltas = _unproject_ndc_to_world(torch.cat([ndc, torch.zeros_like(ndc[..., :1])], dim=-1), clip_to_world_matrix) - starts
    return starts, deltas
