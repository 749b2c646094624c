"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's proxy-mesh rasterization path.

Follows (all paths under /root/reference):
  neural_renderer/neural_renderer/projection.py:6-53                      -> projection()
  neural_renderer/neural_renderer/vertices_to_faces.py:4-46               -> vertices_to_faces() / vertex_attrs_to_faces()
  neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu:24-68     -> face_inv()            (kernel 1)
  neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu:70-169    -> face_index_map()      (kernel 2)
  neural_renderer/neural_renderer/rasterize.py:313-321                    -> the vertical flip in rasterize_rgbad()
  network.py:156-216                                                      -> rasterizer_forward()

Pinning: projection / vertices_to_faces are checked against the real reference's Python (tests/golden/raster.npz, made by
tests/golden/make_golden.py); the two CUDA kernels are checked against the reference's own kernel bodies compiled for the
CPU (oracle/_ref/libref_raster.so, built by oracle/build_oracle.py from the sources where they lie under /root/reference).
The kernels' arithmetic is restated operation by operation in numpy float32 (no FMA contraction), with the
double-promoted sub-expressions of the CUDA source evaluated in float64, so integer outputs are bit-reproducible.
The product path never imports this module.
"""
import numpy as np
import torch

f32 = np.float32


# ------------------------------------------------------------------------------------------------
# projection.py:6-53
# ------------------------------------------------------------------------------------------------
def projection(vertices, K, R, t, dist_coeffs, orig_size, offset=None, scale=None, eps=1e-9):
    """vertices [1|N,nv,3], K/R [N,3,3], t [N,1,3], dist_coeffs [N,5] -> [N,nv,3] (u, v in [-1,1], z)."""
    vertices = torch.matmul(vertices, R.transpose(2, 1)) + t
    x, y, z = vertices[:, :, 0], vertices[:, :, 1], vertices[:, :, 2]
    x_ = x / (z + eps)
    y_ = y / (z + eps)
    k1, k2, p1, p2, k3 = [dist_coeffs[:, None, i] for i in range(5)]
    r = torch.sqrt(x_ ** 2 + y_ ** 2)
    x__ = x_ * (1 + k1 * (r ** 2) + k2 * (r ** 4) + k3 * (r ** 6)) + 2 * p1 * x_ * y_ + p2 * (r ** 2 + 2 * x_ ** 2)
    y__ = y_ * (1 + k1 * (r ** 2) + k2 * (r ** 4) + k3 * (r ** 6)) + p1 * (r ** 2 + 2 * y_ ** 2) + 2 * p2 * x_ * y_
    vertices = torch.stack([x__, y__, torch.ones_like(z)], dim=-1)
    vertices = torch.matmul(vertices, K.transpose(1, 2))
    u, v = vertices[:, :, 0], vertices[:, :, 1]
    if offset is not None and scale is not None:
        u = (u + offset[:, None, 1]) * scale[:, None, 1]
        v = (v + offset[:, None, 0]) * scale[:, None, 0]
    v = orig_size - v
    u = 2 * (u - orig_size / 2.) / orig_size
    v = 2 * (v - orig_size / 2.) / orig_size
    return torch.stack([u, v, z], dim=-1)


def vertices_to_faces(vertices, faces):
    """vertices [N,nv,A], faces [1|N,nf,3] int -> [N,nf,3,A]   (vertices_to_faces.py:4-46; same gather for attributes)."""
    if faces.shape[0] == 1 and vertices.shape[0] != 1:
        faces = faces.repeat(vertices.shape[0], 1, 1)
    bs, nv = vertices.shape[:2]
    faces = faces.long() + (torch.arange(bs) * nv)[:, None, None]
    return vertices.reshape(bs * nv, -1)[faces]


vertex_attrs_to_faces = vertices_to_faces


# ------------------------------------------------------------------------------------------------
# rasterize_cuda_kernel.cu:24-68 (kernel 1)
# ------------------------------------------------------------------------------------------------
def _backside(f):
    return (f[..., 7] - f[..., 1]) * (f[..., 3] - f[..., 0]) < (f[..., 4] - f[..., 1]) * (f[..., 6] - f[..., 0])


def face_inv(faces, image_size):
    """faces [B,nf,9] float32 -> (faces_inv [B,nf,9] (zeros for back faces), p [B,nf,3,2] pixel-space vertices)."""
    f = np.ascontiguousarray(faces, dtype=f32).reshape(faces.shape[0], -1, 9)
    is_ = f32(image_size)
    p = np.zeros(f.shape[:2] + (3, 2), dtype=f32)
    for num in range(3):
        for d in range(2):
            p[..., num, d] = f32(0.5) * (f[..., 3 * num + d] * is_ + is_ - f32(1))
    inv = np.stack([
        p[..., 1, 1] - p[..., 2, 1], p[..., 2, 0] - p[..., 1, 0], p[..., 1, 0] * p[..., 2, 1] - p[..., 2, 0] * p[..., 1, 1],
        p[..., 2, 1] - p[..., 0, 1], p[..., 0, 0] - p[..., 2, 0], p[..., 2, 0] * p[..., 0, 1] - p[..., 0, 0] * p[..., 2, 1],
        p[..., 0, 1] - p[..., 1, 1], p[..., 1, 0] - p[..., 0, 0], p[..., 0, 0] * p[..., 1, 1] - p[..., 1, 0] * p[..., 0, 1]], -1).astype(f32)
    den = (p[..., 2, 0] * (p[..., 0, 1] - p[..., 1, 1]) + p[..., 0, 0] * (p[..., 1, 1] - p[..., 2, 1])
           + p[..., 1, 0] * (p[..., 2, 1] - p[..., 0, 1])).astype(f32)
    with np.errstate(divide='ignore', invalid='ignore'):
        inv = (inv / den[..., None]).astype(f32)
    inv[_backside(f)] = 0
    return inv, p


# ------------------------------------------------------------------------------------------------
# rasterize_cuda_kernel.cu:70-169 (kernel 2)
# ------------------------------------------------------------------------------------------------
def face_index_map(faces, image_size, near=0.0, far=1e5, brute_force=False, return_face_inv=False):
    """faces [B,nf,3,3] (u,v in [-1,1], z) -> face_index_map [B,is,is] int32 (-1 = background), weight_map [B,is,is,3],
    depth_map [B,is,is] (init far), [face_inv_map [B,is,is,9]] -- UNFLIPPED, exactly what the CUDA extension returns.
    Faces are visited in ascending index with a strict ``zp < depth_min`` update, so ties keep the lowest index (:142).
    ``brute_force`` tests every pixel against every face; otherwise only the pixels of a conservative bounding box."""
    B, nf = faces.shape[:2]
    is_ = int(image_size)
    f = np.ascontiguousarray(np.asarray(faces, dtype=f32)).reshape(B, nf, 9)
    inv, p = face_inv(f, is_)
    back = _backside(f)
    fim = np.full((B, is_, is_), -1, dtype=np.int32)
    wm = np.zeros((B, is_, is_, 3), dtype=f32)
    dm = np.full((B, is_, is_), f32(far), dtype=f32)
    fiv = np.zeros((B, is_, is_, 9), dtype=f32) if return_face_inv else None
    idx = np.arange(is_)
    # (2. * i + 1 - is) / is in double, then narrowed (:96-97)
    ndc = ((2. * idx + 1 - is_) / is_).astype(f32)
    near, far = f32(near), f32(far)
    for b in range(B):
        for fn in range(nf):
            if back[b, fn]:
                continue
            fc = f[b, fn]
            if not np.isfinite(fc).all():
                continue
            if brute_force:
                x0, x1, y0, y1 = 0, is_ - 1, 0, is_ - 1
            else:
                x0 = max(int(np.floor(max(p[b, fn, :, 0].min(), -8.0))) - 2, 0)
                x1 = min(int(np.ceil(min(p[b, fn, :, 0].max(), is_ + 8.0))) + 2, is_ - 1)
                y0 = max(int(np.floor(max(p[b, fn, :, 1].min(), -8.0))) - 2, 0)
                y1 = min(int(np.ceil(min(p[b, fn, :, 1].max(), is_ + 8.0))) + 2, is_ - 1)
                if x0 > x1 or y0 > y1:
                    continue
            yp = ndc[y0:y1 + 1][:, None]
            xp = ndc[x0:x1 + 1][None, :]
            inside = ~(((yp - fc[1]) * (fc[3] - fc[0]) < (xp - fc[0]) * (fc[4] - fc[1])) |
                       ((yp - fc[4]) * (fc[6] - fc[3]) < (xp - fc[3]) * (fc[7] - fc[4])) |
                       ((yp - fc[7]) * (fc[0] - fc[6]) < (xp - fc[6]) * (fc[1] - fc[7])))
            if not inside.any():
                continue
            yy, xx = np.nonzero(inside)
            yi = (yy + y0).astype(f32)
            xi = (xx + x0).astype(f32)
            fi = inv[b, fn]
            w = np.stack([fi[3 * k] * xi + fi[3 * k + 1] * yi + fi[3 * k + 2] for k in range(3)], -1).astype(f32)
            w = np.minimum(np.maximum(w, f32(0)), f32(1))
            wsum = (f32(0) + w[:, 0]) + w[:, 1]
            wsum = (wsum + w[:, 2]).astype(f32)
            with np.errstate(divide='ignore', invalid='ignore'):
                w = (w / wsum[:, None]).astype(f32)
                s = ((w[:, 0] / fc[2] + w[:, 1] / fc[5]) + w[:, 2] / fc[8]).astype(f32)
                zp = (1. / s.astype(np.float64)).astype(f32)
            ok = ~((zp <= near) | (far <= zp))
            py, px = yy + y0, xx + x0
            upd = ok & (zp < dm[b, py, px])
            if not upd.any():
                continue
            py, px = py[upd], px[upd]
            dm[b, py, px] = zp[upd]
            fim[b, py, px] = fn
            wm[b, py, px] = w[upd]
            if fiv is not None:
                fiv[b, py, px] = fi
    return (fim, wm, dm, fiv) if return_face_inv else (fim, wm, dm)


def rasterize_rgbad(faces, image_size, near=0.0, far=1e5):
    """rasterize.py:255-340 with the fixed configuration of network.py:145-153 (no anti-aliasing, the rgb of a zero texture is
    skipped): returns dict(alpha, depth, face_index_map, weight_map) as torch tensors, vertically flipped (:313-321)."""
    fim, wm, dm = face_index_map(faces.detach().cpu().numpy(), image_size, near, far)
    alpha = (fim >= 0).astype(f32)
    flip = lambda a: torch.from_numpy(np.ascontiguousarray(a[:, ::-1]))
    return dict(alpha=flip(alpha), depth=flip(dm), face_index_map=flip(fim), weight_map=flip(wm))


# ------------------------------------------------------------------------------------------------
# network.Rasterizer.forward  (network.py:156-216)
# ------------------------------------------------------------------------------------------------
def rasterizer_forward(mesh, img_size, proj, pose, dist_coeffs=None, offset=None, scale=None, v_uvz=None):
    """``mesh`` dict: vertices [1,nv,3], faces [1,nf,3] int, vertices_texcoords [1,nvt,2], faces_vt_idx, vertices_normals,
    faces_vn_idx (the buffers of network.Rasterizer, network.py:129-134).  Returns the reference's 14-tuple (network.py:216)."""
    from .pixel_ops import interpolate_bilinear
    N = proj.shape[0]
    R = pose[:, :3, :3]
    t = pose[:, :3, -1, None].permute(0, 2, 1)
    if dist_coeffs is None:
        dist_coeffs = torch.zeros(N, 5)
    if v_uvz is None:            # (tests may inject the projected vertices to compare everything downstream of them exactly)
        v_uvz = projection(mesh['vertices'], proj, R, t, dist_coeffs, img_size, offset, scale)
    faces_v_idx = mesh['faces']
    faces_v_uvz = vertices_to_faces(v_uvz, faces_v_idx)
    out = rasterize_rgbad(faces_v_uvz, img_size, 0.0, 1e5)
    depth, alpha, fim, weight_map = out['depth'], out['alpha'], out['face_index_map'], out['weight_map']
    v_uvz = v_uvz.clone()
    v_uvz[..., 0] = (v_uvz[..., 0] * 0.5 + 0.5) * depth.shape[2]
    v_uvz[..., 1] = (1 - (v_uvz[..., 1] * 0.5 + 0.5)) * depth.shape[1]
    v_depth = interpolate_bilinear(depth[0, :, :, None], v_uvz[..., 0], v_uvz[..., 1])
    mesh_span = (mesh['vertices'][0].max(dim=0)[0] - mesh['vertices'][0].min(dim=0)[0]).max()
    v_front_mask = ((v_uvz[0, :, 2] - v_depth[0, :, 0]) < mesh_span * 5e-3)[None, :]
    z_inv = torch.stack([1 / faces_v_uvz[i, fim[i].long()][..., -1] for i in range(N)])
    depth = depth.unsqueeze(-1)
    weight_map = ((z_inv * weight_map) * depth).unsqueeze(-1)
    faces_vt = vertex_attrs_to_faces(mesh['vertices_texcoords'], mesh['faces_vt_idx'])
    uv_map = (faces_vt[0, fim.long()] * weight_map).sum(-2)
    uv_map = uv_map - uv_map.floor()
    faces_vn = vertex_attrs_to_faces(mesh['vertices_normals'], mesh['faces_vn_idx'])
    normal_map = torch.nn.functional.normalize((faces_vn[0, fim.long()] * weight_map).sum(-2), dim=-1)
    nflat = normal_map.flatten(1, 2).permute(0, 2, 1)
    normal_map_cam = torch.nn.functional.normalize(R.matmul(nflat).permute(0, 2, 1).reshape(normal_map.shape), dim=-1)
    faces_v = vertex_attrs_to_faces(mesh['vertices'], faces_v_idx)
    position_map = (faces_v[0, fim.long()] * weight_map).sum(-2)
    pflat = position_map.flatten(1, 2).permute(0, 2, 1)
    position_map_cam = R.matmul(pflat).permute(0, 2, 1).reshape(position_map.shape) + pose[:, :3, -1][:, None, None, :]
    return (uv_map, alpha, fim, weight_map, faces_v_idx, normal_map, normal_map_cam, faces_v, faces_vt, position_map,
            position_map_cam, depth, v_uvz, v_front_mask)


# ------------------------------------------------------------------------------------------------
# synthetic proxy meshes: UV sphere with seam-duplicated vertices, 'f v/vt/vn' faces
# ------------------------------------------------------------------------------------------------
def uv_sphere(n_lat=16, n_lon=32, radius=1.0):
    """-> dict of numpy arrays v [nv,3], vt [nv,2], vn [nv,3], f [nf,3] (0-based; the same index addresses v, vt and vn)."""
    lat = np.linspace(0, np.pi, n_lat + 1)
    lon = np.linspace(0, 2 * np.pi, n_lon + 1)
    la, lo = np.meshgrid(lat, lon, indexing='ij')
    v = np.stack([np.sin(la) * np.cos(lo), np.cos(la), np.sin(la) * np.sin(lo)], -1).reshape(-1, 3)
    vt = np.stack([lo / (2 * np.pi), 1 - la / np.pi], -1).reshape(-1, 2)
    f = []
    for i in range(n_lat):
        for j in range(n_lon):
            a, b = i * (n_lon + 1) + j, i * (n_lon + 1) + j + 1
            c, d = (i + 1) * (n_lon + 1) + j, (i + 1) * (n_lon + 1) + j + 1
            f.append([a, b, c])
            f.append([b, d, c])
    return dict(v=(v * radius).astype(f32), vt=vt.astype(f32), vn=v.astype(f32), f=np.asarray(f, dtype=np.int32))


def write_obj(path, m):
    with open(path, 'w') as fh:
        for p in m['v']:
            fh.write('v %.8f %.8f %.8f\n' % tuple(p))
        for p in m['vt']:
            fh.write('vt %.8f %.8f\n' % tuple(p))
        for p in m['vn']:
            fh.write('vn %.8f %.8f %.8f\n' % tuple(p))
        for t in m['f'] + 1:
            fh.write('f %d/%d/%d %d/%d/%d %d/%d/%d\n' % (t[0], t[0], t[0], t[1], t[1], t[1], t[2], t[2], t[2]))
