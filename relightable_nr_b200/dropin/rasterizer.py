"""``network.Rasterizer`` (network.py:100-216) on the tile-culled B200 rasterizer.

One projection kernel, one per-face set-up kernel and ONE tile kernel produce the whole 14-tuple of the reference's
``forward``: coverage / z-buffer, vertical flip, perspective-correct weights and the uv / normal / position maps are fused
(the reference materialises three ``[N,H,W,3,C]`` gathers and loops over the batch in Python, network.py:176-214)."""
import torch
import torch.nn as nn

from .. import ops
from . import neural_renderer as nr


class Rasterizer(nn.Module):
    def __init__(self, obj_fp, img_size, global_RT=None):
        super().__init__()
        v_attr, f_attr = nr.load_obj(obj_fp, normalization=False)
        vertices, faces = v_attr['v'], f_attr['f_v_idx']
        vertices_texcoords, faces_vt_idx = v_attr['vt'], f_attr['f_vt_idx']
        vertices_normals, faces_vn_idx = v_attr['vn'], f_attr['f_vn_idx']
        self.num_vertex = vertices.shape[0]
        self.num_face = faces.shape[0]
        print('vertices shape:', vertices.shape)
        print('faces shape:', faces.shape)
        print('vertices_texcoords shape:', vertices_texcoords.shape)
        print('faces_vt_idx shape:', faces_vt_idx.shape)
        print('vertices_normals shape:', vertices_normals.shape)
        print('faces_vn_idx shape:', faces_vn_idx.shape)
        self.img_size = img_size
        if global_RT is not None:            # network.py:124-126
            g = global_RT.to(vertices.device).to(vertices.dtype)
            vertices = vertices @ g[:3, :3].t() + g[:3, 3]
            vertices_normals = torch.nn.functional.normalize(vertices_normals @ g[:3, :3].t(), dim=1)
        self.register_buffer('vertices', vertices[None].contiguous())
        self.register_buffer('faces', faces[None].contiguous())
        self.register_buffer('vertices_texcoords', vertices_texcoords[None].contiguous())
        self.register_buffer('faces_vt_idx', faces_vt_idx[None].contiguous())
        self.register_buffer('vertices_normals', vertices_normals[None].contiguous())
        self.register_buffer('faces_vn_idx', faces_vn_idx[None].contiguous())
        self.mesh_span = (self.vertices[0].max(dim=0)[0] - self.vertices[0].min(dim=0)[0]).max()
        # per-face 4^3 rgb texture of the reference (network.py:138-142): zero, only ever rendered into a discarded image
        self.textures = nn.Parameter(torch.zeros(1, self.faces.shape[1], 4, 4, 4, 3, dtype=torch.float32))
        renderer = nr.Renderer(image_size=img_size, camera_mode='projection', orig_size=img_size, near=0.0, far=1e5)
        renderer.light_intensity_directional = 0.0
        renderer.light_intensity_ambient = 1.0
        renderer.anti_aliasing = False
        renderer.fill_back = False
        self.renderer = renderer
        self._static = None

    def _static_faces(self):
        """faces_v / faces_vt of network.py:187,208 depend on the mesh only: gathered once."""
        if self._static is None or self._static[0].device != self.vertices.device:
            self._static = (nr.vertex_attrs_to_faces(self.vertices, self.faces),
                            nr.vertex_attrs_to_faces(self.vertices_texcoords, self.faces_vt_idx))
        return self._static

    def forward(self, proj, pose, dist_coeffs, offset, scale):
        r = self.renderer
        if r.fill_back or r.anti_aliasing:
            raise NotImplementedError('Rasterizer: fill_back / anti_aliasing are fixed to False on the relighting path (network.py:149-152)')
        N = proj.shape[0]
        dev = self.vertices.device
        R = pose[:, :3, :3].contiguous()
        t = pose[:, :3, -1].contiguous()
        dist = dist_coeffs if dist_coeffs is not None else torch.zeros((1, 5), dtype=torch.float32, device=dev)
        v_uvz = nr.projection(self.vertices, proj.to(dev), R.to(dev), t.to(dev)[:, None, :], dist.to(dev), r.orig_size,
                              offset=offset, scale=scale)
        g = nr.raster_gbuffer(r.image_size, r.near, r.far, uvz=v_uvz, faces_idx=self.faces, flip_y=True,
                              attrs=dict(v=self.vertices[0], f_v_idx=self.faces[0], vt=self.vertices_texcoords[0],
                                         f_vt_idx=self.faces_vt_idx[0], vn=self.vertices_normals[0], f_vn_idx=self.faces_vn_idx[0],
                                         pose_R=R.to(dev), pose_t=t.to(dev)))
        depth, alpha, face_index_map = g['depth'], g['alpha'], g['face_index_map']
        # per-vertex visibility (network.py:170-173); like the reference it uses the depth map of batch item 0
        v_uvz[..., 0] = (v_uvz[..., 0] * 0.5 + 0.5) * depth.shape[2]
        v_uvz[..., 1] = (1 - (v_uvz[..., 1] * 0.5 + 0.5)) * depth.shape[1]
        v_depth = ops.interpolate_bilinear(depth[0, :, :, None], v_uvz[..., 0], v_uvz[..., 1])
        v_front_mask = ((v_uvz[0, :, 2] - v_depth[0, :, 0]) < self.mesh_span * 5e-3)[None, :]
        faces_v, faces_vt = self._static_faces()
        return (g['uv_map'], alpha, face_index_map, g['weight_pc'][..., None], self.faces, g['normal_map'], g['normal_map_cam'],
                faces_v, faces_vt, g['position_map'], g['position_map_cam'], depth[..., None], v_uvz, v_front_mask)
