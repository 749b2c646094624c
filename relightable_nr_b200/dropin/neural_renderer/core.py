"""Host side of the ``neural_renderer`` drop-in: OBJ loading, camera projection, face gathers and the rasterizer front end
(neural_renderer/neural_renderer/{load_obj,projection,vertices_to_faces,lighting,rasterize,renderer}.py), on librnr_b200.so."""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from ... import _lib

vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
_lib.register_sigs({
    "rnr_project_vertices": [vp, i32, i32, vp, vp, vp, vp, vp, vp, f32, f32, vp, i32, vp],
    "rnr_raster_face_setup": [vp, i32, vp, i32, vp, i32, i32, vp, vp, vp, i32, vp],
    "rnr_raster_tiles": [vp, vp, vp, i32, i32, f32, f32, i32, vp, vp, vp, vp, vp, vp, i32, vp],
})

DEFAULT_IMAGE_SIZE = 256
DEFAULT_ANTI_ALIASING = True
DEFAULT_NEAR = 0.1
DEFAULT_FAR = 100
DEFAULT_EPS = 1e-4
DEFAULT_BACKGROUND_COLOR = (0, 0, 0)


class RasterAttrs(C.Structure):
    _fields_ = [(k, vp) for k in ('v', 'f_v_idx', 'vt', 'f_vt_idx', 'vn', 'f_vn_idx', 'pose_R', 'pose_t', 'weight_pc', 'uv_map',
                                  'normal_map', 'normal_map_cam', 'position_map', 'position_map_cam')]


def _s():
    return torch.cuda.current_stream().cuda_stream


def _cf(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError('%s must be a CUDA tensor (librnr_b200 has no CPU path)' % name)
    return t.float().contiguous()


def _p(t):
    return t.data_ptr() if t is not None else None


# ----------------------------------------------------------------------------------------------------------------------
# load_obj.py:108-215
# ----------------------------------------------------------------------------------------------------------------------
def load_obj(filename_obj, normalization=True, texture_size=4, load_texture=False, texture_wrapping='REPEAT', use_bilinear=True,
             use_cuda=True):
    """Wavefront OBJ -> (v_attr {'v','vn','vt'}, f_attr {'f_v_idx','f_vn_idx','f_vt_idx'}) with 0-based int32 faces.
    One pass over the file (the reference re-reads the line list four times); triangles only, 'f v/vt/vn' records."""
    if load_texture:
        raise NotImplementedError('load_obj(load_texture=True) is outside the relighting hot path (network.py:106 never asks for it)')
    v, vn, vt, fv, fvt, fvn = [], [], [], [], [], []
    with open(filename_obj) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == 'v':
                v.append([float(x) for x in tok[1:4]])
            elif tok[0] == 'vn':
                vn.append([float(x) for x in tok[1:4]])
            elif tok[0] == 'vt':
                vt.append([float(x) for x in tok[1:3]])
            elif tok[0] == 'f':
                parts = [t.split('/') for t in tok[1:]]
                fv.append([int(p[0]) for p in parts])
                if len(parts[0]) > 1 and parts[0][1] != '':
                    fvt.append([int(p[1]) for p in parts])
                if len(parts[0]) > 1:
                    fvn.append([int(p[-1]) for p in parts])
    dev = 'cuda' if use_cuda else 'cpu'
    as_f = lambda a, w: torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(-1, w)).to(dev)
    as_i = lambda a: (torch.from_numpy(np.asarray(a, dtype=np.int32).reshape(-1, 3)) - 1).to(dev)
    vertices = as_f(v, 3)
    if normalization:
        vertices -= vertices.min(0)[0][None, :]
        vertices /= torch.abs(vertices).max()
        vertices *= 2
        vertices -= vertices.max(0)[0][None, :] / 2
    v_attr = {'v': vertices, 'vn': as_f(vn, 3) if vn else [], 'vt': as_f(vt, 2) if vt else []}
    f_attr = {'f_v_idx': as_i(fv), 'f_vn_idx': as_i(fvn if vn else []), 'f_vt_idx': as_i(fvt if vt else [])}
    return v_attr, f_attr


# ----------------------------------------------------------------------------------------------------------------------
# projection.py:6-53, vertices_to_faces.py:4-46
# ----------------------------------------------------------------------------------------------------------------------
def projection(vertices, K, R, t, dist_coeffs, orig_size, offset=None, scale=None, eps=1e-9):
    """vertices [1|N,nv,3], K/R [N,3,3], t [N,1,3], dist_coeffs [1|N,5] -> [N,nv,3] = (u, v in [-1,1] with y up, z)."""
    v = _cf(vertices, 'vertices')
    K, R, t = _cf(K, 'K'), _cf(R, 'R'), _cf(t, 't').reshape(-1, 3)
    N = K.shape[0]
    if R.shape[0] != N or t.shape[0] != N or v.shape[0] not in (1, N):
        raise ValueError('projection: inconsistent batch sizes')
    dist = None
    if dist_coeffs is not None:
        dist = _cf(dist_coeffs, 'dist_coeffs').reshape(-1, 5)
        if dist.shape[0] != N:
            dist = dist.expand(N, 5).contiguous()
    off = _cf(offset, 'offset').reshape(N, 2) if offset is not None else None
    sc = _cf(scale, 'scale').reshape(N, 2) if scale is not None else None
    out = torch.empty((N, v.shape[1], 3), dtype=torch.float32, device=v.device)
    _lib.check(_lib.lib().rnr_project_vertices(v.data_ptr(), v.shape[0], v.shape[1], K.data_ptr(), R.data_ptr(), t.data_ptr(), _p(dist),
                                               _p(off), _p(sc), float(orig_size), float(eps), out.data_ptr(), N, _s()),
               'rnr_project_vertices')
    return out


def vertices_to_faces(vertices, faces):
    """vertices [N,nv,3], faces [1|N,nf,3] -> [N,nf,3,3]."""
    if faces.shape[0] == 1 and vertices.shape[0] != 1:
        faces = faces.expand(vertices.shape[0], -1, -1)
    assert vertices.ndimension() == 3 and faces.ndimension() == 3
    assert vertices.shape[0] == faces.shape[0] and vertices.shape[2] == 3 and faces.shape[2] == 3
    bs, nv = vertices.shape[:2]
    idx = faces.long() + (torch.arange(bs, device=vertices.device) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[idx]


def vertex_attrs_to_faces(vertex_attrs, faces):
    """vertex_attrs [N,nv,A], faces [N,nf,3] -> [N,nf,3,A]."""
    assert vertex_attrs.ndimension() == 3 and faces.ndimension() == 3
    assert vertex_attrs.shape[0] == faces.shape[0] and faces.shape[2] == 3
    bs, nv, na = vertex_attrs.shape
    idx = faces.long() + (torch.arange(bs, device=vertex_attrs.device) * nv)[:, None, None]
    return vertex_attrs.reshape(bs * nv, na)[idx]


def lighting(faces, textures, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
             color_directional=(1, 1, 1), direction=(0, 1, 0)):
    """Per-face ambient + directional shading of the face textures (lighting.py:5-57), in place like the reference."""
    bs, nf = faces.shape[:2]
    dev = faces.device
    tt = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a, dtype=torch.float32, device=dev).reshape(-1, 3)
    light = torch.zeros(bs, nf, 3, dtype=torch.float32, device=dev)
    if intensity_ambient != 0:
        light += intensity_ambient * tt(color_ambient)[:, None, :]
    if intensity_directional != 0:
        f = faces.reshape(bs * nf, 3, 3)
        nrm = torch.nn.functional.normalize(torch.cross(f[:, 0] - f[:, 1], f[:, 2] - f[:, 1], dim=1), eps=1e-5).reshape(bs, nf, 3)
        cos = torch.relu((nrm * tt(direction)[:, None, :]).sum(2))
        light += intensity_directional * (tt(color_directional)[:, None, :] * cos[:, :, None])
    textures *= light[:, :, None, None, None, :]
    return textures


def _vec(a, dev, bs):
    a = torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a, dtype=torch.float32, device=dev)
    return a[None].repeat(bs, 1) if a.dim() == 1 else a


def look_at(vertices, eye, at=[0, 0, 0], up=[0, 1, 0]):
    """look_at.py: rotate/translate vertices [N,nv,3] into the frame of a camera at ``eye`` looking at ``at``."""
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    bs, dev = vertices.shape[0], vertices.device
    eye, at, up = _vec(eye, dev, bs), _vec(at, dev, bs), _vec(up, dev, bs)
    nz = torch.nn.functional.normalize
    z = nz(at - eye, eps=1e-5)
    x = nz(torch.cross(up, z, dim=1), eps=1e-5)
    y = nz(torch.cross(z, x, dim=1), eps=1e-5)
    r = torch.stack((x, y, z), dim=1)
    return torch.matmul(vertices - eye[:, None, :], r.transpose(1, 2))


def look(vertices, eye, direction=[0, 1, 0], up=None):
    """look.py: like look_at with an explicit viewing direction."""
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    bs, dev = vertices.shape[0], vertices.device
    eye, direction = _vec(eye, dev, 1), _vec(direction, dev, 1)
    up = _vec([0, 1, 0] if up is None else up, dev, 1)
    nz = torch.nn.functional.normalize
    z = nz(direction, eps=1e-5)
    x = nz(torch.cross(up, z, dim=1), eps=1e-5)
    y = nz(torch.cross(z, x, dim=1), eps=1e-5)
    r = torch.stack((x, y, z), dim=1)
    return torch.matmul(vertices - eye[:, None, :], r.transpose(1, 2))


def perspective(vertices, angle=30.):
    """perspective.py: x, y divided by z tan(angle)."""
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    width = math.tan(angle / 180 * math.pi)
    z = vertices[:, :, 2]
    return torch.stack((vertices[:, :, 0] / z / width, vertices[:, :, 1] / z / width, z), dim=2)


def get_points_from_angles(distance, elevation, azimuth, degrees=True):
    if isinstance(distance, (float, int)):
        if degrees:
            elevation, azimuth = math.radians(elevation), math.radians(azimuth)
        return (distance * math.cos(elevation) * math.sin(azimuth), distance * math.sin(elevation),
                -distance * math.cos(elevation) * math.cos(azimuth))
    if degrees:
        elevation, azimuth = math.pi / 180. * elevation, math.pi / 180. * azimuth
    return torch.stack([distance * torch.cos(elevation) * torch.sin(azimuth), distance * torch.sin(elevation),
                        -distance * torch.cos(elevation) * torch.cos(azimuth)]).transpose(1, 0)


# ----------------------------------------------------------------------------------------------------------------------
# rasterizer front end
# ----------------------------------------------------------------------------------------------------------------------
def raster_gbuffer(image_size, near, far, faces=None, uvz=None, faces_idx=None, flip_y=True, attrs=None, want_alpha=True):
    """Runs face set-up + the tile kernel.  Either ``faces`` [N,nf,3,3] or (``uvz`` [N,nv,3], ``faces_idx`` [1|N,nf,3] int32).
    ``attrs``: optional dict with v, f_v_idx, vt, f_vt_idx, vn, f_vn_idx, pose_R [N,3,3], pose_t [N,3] -> the fused maps of
    network.Rasterizer.forward are produced too.  Returns a dict of tensors."""
    L = _lib.lib()
    s = _s()
    is_ = int(image_size)
    if faces is not None:
        faces = _cf(faces, 'faces')
        N, nf = faces.shape[0], faces.shape[1]
        dev = faces.device
        faces_out = faces
    else:
        uvz = _cf(uvz, 'uvz')
        N, nv = uvz.shape[0], uvz.shape[1]
        fi = faces_idx.to(torch.int32).contiguous()
        nf = fi.shape[1]
        dev = uvz.device
        faces_out = torch.empty((N, nf, 3, 3), dtype=torch.float32, device=dev)
    faces_inv = torch.empty((N, nf, 3, 3), dtype=torch.float32, device=dev)
    bbox = torch.empty((N, nf, 2), dtype=torch.int32, device=dev)
    if faces is not None:
        _lib.check(L.rnr_raster_face_setup(None, 0, None, 1, faces.data_ptr(), nf, is_, None, faces_inv.data_ptr(), bbox.data_ptr(), N, s),
                   'rnr_raster_face_setup')
    else:
        _lib.check(L.rnr_raster_face_setup(uvz.data_ptr(), nv, fi.data_ptr(), fi.shape[0], None, nf, is_, faces_out.data_ptr(),
                                           faces_inv.data_ptr(), bbox.data_ptr(), N, s), 'rnr_raster_face_setup')
    out = {
        'faces': faces_out,
        'face_index_map': torch.empty((N, is_, is_), dtype=torch.int32, device=dev),
        'weight_map': torch.empty((N, is_, is_, 3), dtype=torch.float32, device=dev),
        'depth': torch.empty((N, is_, is_), dtype=torch.float32, device=dev),
        'alpha': torch.empty((N, is_, is_), dtype=torch.float32, device=dev) if want_alpha else None,
    }
    a_ptr, keep = None, []
    if attrs is not None:
        st = RasterAttrs()
        for k in ('v', 'vt', 'vn', 'pose_R', 'pose_t'):
            t = _cf(attrs[k], k)
            keep.append(t)
            setattr(st, k, t.data_ptr())
        for k in ('f_v_idx', 'f_vt_idx', 'f_vn_idx'):
            t = attrs[k].to(torch.int32).contiguous()
            if t.shape[-2] != nf:
                raise ValueError('%s must have one row per face' % k)
            keep.append(t)
            setattr(st, k, t.data_ptr())
        for k, c in (('weight_pc', 3), ('uv_map', 2), ('normal_map', 3), ('normal_map_cam', 3), ('position_map', 3), ('position_map_cam', 3)):
            out[k] = torch.empty((N, is_, is_, c), dtype=torch.float32, device=dev)
            setattr(st, k, out[k].data_ptr())
        a_ptr = C.cast(C.pointer(st), vp)
        keep.append(st)
    _lib.check(L.rnr_raster_tiles(faces_out.data_ptr(), faces_inv.data_ptr(), bbox.data_ptr(), nf, is_, float(near), float(far),
                                  1 if flip_y else 0, out['face_index_map'].data_ptr(), out['weight_map'].data_ptr(),
                                  out['depth'].data_ptr(), _p(out['alpha']), None, a_ptr, N, s), 'rnr_raster_tiles')
    return out


def rasterize_rgbad(faces, textures=None, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING, near=DEFAULT_NEAR,
                    far=DEFAULT_FAR, eps=DEFAULT_EPS, background_color=DEFAULT_BACKGROUND_COLOR, return_rgb=True, return_alpha=True,
                    return_depth=True):
    """rasterize.py:255-340.  'rgb' is the shaded face-texture image; the relighting path feeds an all-zero texture
    (tanh(0), network.py:140-159) and discards it, so it is returned as the background-composited constant image without
    sampling.  Maps come back vertically flipped exactly like the reference's index-list gather (:313-321)."""
    is_ = int(image_size) * (2 if anti_aliasing else 1)
    g = raster_gbuffer(is_, near, far, faces=faces, flip_y=True, want_alpha=True)
    alpha, depth = g['alpha'], g['depth']
    rgb = None
    if return_rgb:
        if textures is not None and bool((textures != 0).any()):
            raise NotImplementedError('rasterize_rgbad: sampling a non-zero face texture is outside the relighting hot path')
        bgc = torch.tensor(background_color, dtype=torch.float32, device=alpha.device)
        rgb = (bgc[None, :, None, None] * (1 - alpha[:, None])).contiguous()
    if anti_aliasing:
        pool = torch.nn.functional.avg_pool2d
        rgb = pool(rgb, kernel_size=(2, 2)) if rgb is not None else None
        alpha = pool(alpha[:, None], kernel_size=(2, 2))[:, 0]
        depth = pool(depth[:, None], kernel_size=(2, 2))[:, 0]
    return {'rgb': rgb if return_rgb else None, 'alpha': alpha if return_alpha else None, 'depth': depth if return_depth else None,
            'face_index_map': g['face_index_map'], 'weight_map': g['weight_map']}


def rasterize(faces, textures, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING, near=DEFAULT_NEAR, far=DEFAULT_FAR,
              eps=DEFAULT_EPS, background_color=DEFAULT_BACKGROUND_COLOR):
    return rasterize_rgbad(faces, textures, image_size, anti_aliasing, near, far, eps, background_color, True, False, False)['rgb']


def rasterize_silhouettes(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING, near=DEFAULT_NEAR, far=DEFAULT_FAR,
                          eps=DEFAULT_EPS):
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None, False, True, False)['alpha']


def rasterize_depth(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING, near=DEFAULT_NEAR, far=DEFAULT_FAR,
                    eps=DEFAULT_EPS):
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None, False, False, True)['depth']


class Rasterize(nn.Module):
    """rasterize.py:224-252: module wrapper returning (rgb, alpha, depth, face_index_map, weight_map), UNFLIPPED."""

    def __init__(self, image_size, near, far, eps, background_color, return_rgb=False, return_alpha=False, return_depth=False):
        super().__init__()
        self.image_size, self.near, self.far, self.eps = image_size, near, far, eps
        self.background_color = background_color
        self.return_rgb, self.return_alpha, self.return_depth = return_rgb, return_alpha, return_depth

    def forward(self, faces, textures=None):
        g = raster_gbuffer(self.image_size, self.near, self.far, faces=faces, flip_y=False)
        e = torch.tensor([])
        rgb = e
        if self.return_rgb:
            bgc = torch.tensor(self.background_color, dtype=torch.float32, device=faces.device)
            rgb = (bgc[None, None, None, :] * (1 - g['alpha'][..., None])).contiguous()
        return rgb, (g['alpha'] if self.return_alpha else e), (g['depth'] if self.return_depth else e), g['face_index_map'], g['weight_map']


class Renderer(nn.Module):
    """renderer.py:11-257.  ``forward(vertices, faces, textures, K=..., R=..., t=...)`` (mode None) returns
    (rgb, depth, alpha, face_index_map, weight_map, vertices_uvz, faces_uvz, faces) like renderer.py:207-257."""

    def __init__(self, image_size=256, anti_aliasing=True, background_color=[0, 0, 0], fill_back=True, camera_mode='projection',
                 K=None, R=None, t=None, dist_coeffs=None, orig_size=1024, offset=None, scale=None, perspective=True,
                 viewing_angle=30, camera_direction=[0, 0, 1], near=0.1, far=100, light_intensity_ambient=0.5,
                 light_intensity_directional=0.5, light_color_ambient=[1, 1, 1], light_color_directional=[1, 1, 1],
                 light_direction=[0, 1, 0]):
        super().__init__()
        self.image_size = image_size
        self.anti_aliasing = anti_aliasing
        self.background_color = background_color
        self.fill_back = fill_back
        self.camera_mode = camera_mode
        if camera_mode == 'projection':
            cu = lambda a: torch.as_tensor(a, dtype=torch.float32).cuda() if isinstance(a, np.ndarray) else a
            self.K, self.R, self.t = cu(K), cu(R), cu(t)
            self.dist_coeffs = dist_coeffs if dist_coeffs is not None else torch.zeros((1, 5), dtype=torch.float32, device='cuda')
            self.orig_size, self.offset, self.scale = orig_size, offset, scale
        elif camera_mode in ('look', 'look_at'):
            self.perspective = perspective
            self.viewing_angle = viewing_angle
            self.eye = [0, 0, -(1. / math.tan(math.radians(viewing_angle)) + 1)]
            self.camera_direction = [0, 0, 1]
        else:
            raise ValueError('Camera mode has to be one of projection, look or look_at')
        self.near, self.far = near, far
        self.light_intensity_ambient = light_intensity_ambient
        self.light_intensity_directional = light_intensity_directional
        self.light_color_ambient = light_color_ambient
        self.light_color_directional = light_color_directional
        self.light_direction = light_direction
        self.rasterizer_eps = 1e-3

    def _camera(self, vertices, K, R, t, dist_coeffs, orig_size, offset, scale):
        if self.camera_mode == 'look_at':
            vertices = look_at(vertices, self.eye)
            return perspective(vertices, angle=self.viewing_angle) if self.perspective else vertices
        if self.camera_mode == 'look':
            vertices = look(vertices, self.eye, self.camera_direction)
            return perspective(vertices, angle=self.viewing_angle) if self.perspective else vertices
        K = self.K if K is None else K
        R = self.R if R is None else R
        t = self.t if t is None else t
        dist_coeffs = self.dist_coeffs if dist_coeffs is None else dist_coeffs
        orig_size = self.orig_size if orig_size is None else orig_size
        offset = self.offset if offset is None else offset
        scale = self.scale if scale is None else scale
        return projection(vertices, K, R, t, dist_coeffs, orig_size, offset=offset, scale=scale)

    def forward(self, vertices, faces, textures=None, mode=None, K=None, R=None, t=None, dist_coeffs=None, orig_size=None,
                offset=None, scale=None):
        if mode is None:
            return self.render(vertices, faces, textures, K, R, t, dist_coeffs, orig_size, offset=offset, scale=scale)
        if mode == 'silhouettes':
            return self.render_silhouettes(vertices, faces, K, R, t, dist_coeffs, orig_size)
        if mode == 'depth':
            return self.render_depth(vertices, faces, K, R, t, dist_coeffs, orig_size)
        if mode == 'rgb':
            return self.render_rgb(vertices, faces, textures, K, R, t, dist_coeffs, orig_size)
        raise ValueError("mode should be one of None, 'silhouettes' or 'depth'")

    def _faces(self, faces):
        if self.fill_back:
            faces = torch.cat((faces, faces.flip(-1)), dim=1).detach()
        return faces

    def render_silhouettes(self, vertices, faces, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        faces = self._faces(faces)
        vertices = self._camera(vertices, K, R, t, dist_coeffs, orig_size, None, None)
        return rasterize_silhouettes(vertices_to_faces(vertices, faces), self.image_size, self.anti_aliasing, self.near, self.far,
                                     self.rasterizer_eps)

    def render_depth(self, vertices, faces, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        faces = self._faces(faces)
        vertices = self._camera(vertices, K, R, t, dist_coeffs, orig_size, None, None)
        return rasterize_depth(vertices_to_faces(vertices, faces), self.image_size, self.anti_aliasing, self.near, self.far,
                               self.rasterizer_eps)

    def render_rgb(self, vertices, faces, textures, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        return self.render(vertices, faces, textures, K, R, t, dist_coeffs, orig_size)[0]

    def render(self, vertices, faces, textures, K=None, R=None, t=None, dist_coeffs=None, orig_size=None, offset=None, scale=None):
        faces = self._faces(faces)
        if textures is not None and self.fill_back:
            textures = torch.cat((textures, textures.permute((0, 1, 4, 3, 2, 5))), dim=1)
        vertices = self._camera(vertices, K, R, t, dist_coeffs, orig_size, offset, scale)
        faces_v = vertices_to_faces(vertices, faces)
        out = rasterize_rgbad(faces_v, textures, self.image_size, self.anti_aliasing, self.near, self.far, self.rasterizer_eps,
                              self.background_color)
        return out['rgb'], out['depth'], out['alpha'], out['face_index_map'], out['weight_map'], vertices, faces_v, faces
