"""Drop-in for the reference's ``camera`` module (camera.py:5-76)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib

vp, i32 = C.c_void_p, C.c_int
_lib.register_sigs({"rnr_view_dir_map": [vp, vp, vp, vp, i32, i32, i32, vp]})


def get_view_dir_map(img_size, proj_inv, R_inv):
    """img_size (H, W), proj_inv [N,3,3], R_inv [N,3,3] -> (view_dir_map [N,H,W,3] world, view_dir_map_cam) (camera.py:5-32):
    ray through the pixel centre, -K^-1 [u+.5, v+.5, 1], normalised; world = normalise(R^-1 cam).  One kernel for the batch."""
    if not proj_inv.is_cuda:
        raise TypeError('get_view_dir_map needs CUDA tensors (librnr_b200 has no CPU path)')
    H, W = int(img_size[0]), int(img_size[1])
    N = proj_inv.shape[0]
    Ki = proj_inv.float().contiguous()
    Ri = R_inv.to(Ki.device).float().contiguous()
    world = torch.empty((N, H, W, 3), dtype=torch.float32, device=Ki.device)
    cam = torch.empty_like(world)
    _lib.check(_lib.lib().rnr_view_dir_map(Ki.data_ptr(), Ri.data_ptr(), world.data_ptr(), cam.data_ptr(), N, H, W,
                                           torch.cuda.current_stream().cuda_stream), 'rnr_view_dir_map')
    return world, cam


def get_reflect_dir(orig_dir, pivot_dir, dim=-1):
    """Mirror ``orig_dir`` about ``pivot_dir`` along ``dim`` and normalise (camera.py:35-45).  Generic-shape helper; the hot
    path (network.RaySampler) has this fused into its kernel."""
    d = (pivot_dir * orig_dir).sum(dim=dim, keepdim=True)
    return torch.nn.functional.normalize(d * 2.0 * pivot_dir - orig_dir, dim=dim)


def RT_from_pos_lookat(cam_pos, cam_lookat=np.array([0., 0., 0.]), cam_up=np.array([0., 1., 0.])):
    """4x4 world->camera matrix of a camera at ``cam_pos`` looking at ``cam_lookat`` (x right, y down, z forward; camera.py:48-69)."""
    fwd = cam_lookat - cam_pos
    fwd = fwd / np.linalg.norm(fwd)
    right = np.cross(fwd, cam_up)
    right = right / np.linalg.norm(right)
    up = np.cross(right, fwd)
    RT = np.eye(4)
    RT[:3, :3] = np.stack((right, -up, fwd)).astype(cam_pos.dtype)
    RT[:3, 3] = -RT[:3, :3].dot(cam_pos)
    return RT


def get_spiral(step_azi=-2, step_ele=90.0 / 720):
    n = int(np.floor(90.0 / step_ele))
    return np.arange(0, step_azi * n, step=step_azi), np.arange(0, step_ele * n, step=step_ele)
