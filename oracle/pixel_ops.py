"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch fp32 restatement of the reference's per-pixel operators.

Each function cites the reference lines it follows (/root/reference/...).  Pinned against the real
reference by tests/golden/make_golden.py -> tests/golden/pixel_ops.npz (tests/test_oracle_golden.py).
The product path never imports this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# misc.interpolate_bilinear  (misc.py:5-42)
# ------------------------------------------------------------------------------------------------
def interpolate_bilinear(data, sx, sy):
    """data [H,W,C]; sx, sy [...] float pixel coordinates -> [..., C].
    Hard validity mask (misc.py:14), clamped neighbours (:22-25), edge weight fix-up (:33-35)."""
    Hd, Wd = data.shape[0], data.shape[1]
    valid = ((sx >= 0) & (sx <= Wd - 1) & (sy >= 0) & (sy <= Hd - 1)).to(data.dtype)
    x0 = torch.floor(sx).long()
    y0 = torch.floor(sy).long()
    x1, y1 = x0 + 1, y0 + 1
    x0, x1 = x0.clamp(0, Wd - 1), x1.clamp(0, Wd - 1)
    y0, y1 = y0.clamp(0, Hd - 1), y1.clamp(0, Hd - 1)
    taps = [data[y0, x0], data[y1, x0], data[y0, x1], data[y1, x1]]
    x0w = x0 - (x0 == x1).long()
    y0w = y0 - (y0 == y1).long()
    ax, bx = x1.to(data.dtype) - sx, sx - x0w.to(data.dtype)
    ay, by = y1.to(data.dtype) - sy, sy - y0w.to(data.dtype)
    wts = [ax * ay * valid, ax * by * valid, bx * ay * valid, bx * by * valid]
    out = taps[0] * wts[0][..., None]
    for t, w in zip(taps[1:], wts[1:]):
        out = out + t * w[..., None]
    return out


# ------------------------------------------------------------------------------------------------
# network.TextureMapper  (network.py:67-99)
# ------------------------------------------------------------------------------------------------
def texture_mapper_forward(textures, uv_map, sh_basis_map=None, sh_start_ch=3, apply_sh=True):
    """textures: list of [1,S_i,S_i,C]; uv_map [N,H,W,2]; returns [N,C,H,W] (network.py:73-91)."""
    out = None
    for tex in textures:
        S = tex.shape[1]
        tx = uv_map[..., 0] * (S - 1)
        ty = (S - 1) - uv_map[..., 1] * (S - 1)
        s = interpolate_bilinear(tex[0], tx, ty).permute(0, 3, 1, 2)
        out = s if out is None else out + s
    if apply_sh and sh_basis_map is not None:
        sh = sh_basis_map.permute(0, 3, 1, 2)
        out = torch.cat([out[:, :sh_start_ch], out[:, sh_start_ch:sh_start_ch + 9] * sh, out[:, sh_start_ch + 9:]], 1)
    return out


def flatten_mipmap(textures, start_ch, end_ch):
    """network.py:93-99: level 0 + bilinear up-sampling (align_corners=False) of the coarser levels."""
    S0 = textures[0].shape[1]
    out = textures[0][..., start_ch:end_ch]
    for tex in textures[1:]:
        up = F.interpolate(tex[..., start_ch:end_ch].permute(0, 3, 1, 2), size=(S0, S0), mode='bilinear')
        out = out + up.permute(0, 2, 3, 1)
    return out


# ------------------------------------------------------------------------------------------------
# render.spherical_mapping*  (render.py:87-121)
# ------------------------------------------------------------------------------------------------
def spherical_uv(d, dim):
    """direction -> equirect (u,v): u = atan2(z,x)/(2pi)+0.5, v = acos(y)/pi along ``dim`` (render.py:96-102)."""
    x, y, z = d.select(dim, 0), d.select(dim, 1), d.select(dim, 2)
    u = torch.atan2(z, x) * 0.5 / np.pi + 0.5
    v = torch.acos(y) * 1.0 / np.pi
    return torch.stack((u, v), dim=dim)


def spherical_mapping_inv(uv):
    """render.py:108-121, uv [2,M] -> dir [3,M]."""
    y = torch.cos(uv[1] * np.pi)
    rxz = (1 - y ** 2).sqrt()
    t = uv[0] * 2 - 1
    x = rxz * torch.cos(t * np.pi)
    z = rxz * torch.sin(t * np.pi)
    z = z * ((~(t == 1.0)).to(rxz.dtype) * 2 - 1)
    z = z * ((~(t == -1.0)).to(rxz.dtype) * 2 - 1)
    return F.normalize(torch.stack((x, y, z), 0), dim=0)


# ------------------------------------------------------------------------------------------------
# network.RaySampler  (network.py:417-472), data_util.euler_to_rot (data_util.py:176-191)
# ------------------------------------------------------------------------------------------------
def ray_sampler_constants(num_azi, num_polar, interval_polar=5):
    """Rs [R,3,3] and pivots_dir [3,R] (network.py:426-443): R = Rz(azi) Ry(polar) Rx(0), pivot = R e_z."""
    num_azi, num_polar, interval_polar = int(num_azi), int(num_polar), float(interval_polar)
    pol = np.arange(1, num_polar + 1) * interval_polar * np.pi / 180.0
    azi = np.arange(num_azi) * 2 * np.pi / num_azi
    Rs = [np.eye(3)]
    for a in azi:                      # np.meshgrid(pol, azi) flattened: azimuth outer, polar inner
        for p in pol:
            cy, sy_, cz, sz = math.cos(p), math.sin(p), math.cos(a), math.sin(a)
            Ry = np.array([[cy, 0, sy_], [0, 1, 0], [-sy_, 0, cy]])
            Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
            Rs.append(Rz @ Ry)
    Rs = torch.from_numpy(np.stack(Rs).astype(np.float32))
    pivots = torch.matmul(Rs, torch.tensor([0.0, 0.0, 1.0])[:, None])[..., 0].permute(1, 0).contiguous()
    return Rs, pivots


def ray_sampler_forward(pivots, TBN, vdt, alpha, mode='reflect'):
    """TBN [N,H,W,3,3], vdt [N,H,W,3], alpha [N,H,W,1] -> rays_dir [N,H,W,3,R], rays_uv [N,H,W,2,R], tangent dirs."""
    R = pivots.shape[1]
    if mode == 'reflect':
        v = vdt[..., None]                                               # [N,H,W,3,1]
        refl = (pivots * v).sum(-2, keepdim=True) * 2.0 * pivots - v    # camera.py:44
        tan = F.normalize(refl, dim=-2) * alpha[..., None]
        world = torch.matmul(TBN.reshape(-1, 3, 3), tan.reshape(-1, 3, R)).reshape(*TBN.shape[:-1], R)
    else:
        tan = pivots
        world = torch.matmul(TBN.reshape(-1, 3, 3), pivots).reshape(*TBN.shape[:-1], R)
    world = F.normalize(world, dim=-2)
    uv = spherical_uv(world, dim=-2)
    uv = uv * alpha[..., None] - (alpha[..., None] == 0).to(world.dtype)
    return world, uv, tan


# ------------------------------------------------------------------------------------------------
# network.RayRenderer  (network.py:481-527)
# ------------------------------------------------------------------------------------------------
def ray_renderer_forward(albedo_specular, rays_uv, rays_lt, lp, albedo_diffuse=None, num_ray_diffuse=0,
                         no_albedo=False, seperate_albedo=False):
    """lp [1 or N, Hl, Wl, 3] (already scaled).  Returns the reference's 7-tuple minus lp."""
    Hl, Wl = lp.shape[1], lp.shape[2]
    n_spec = rays_uv.shape[-1] - num_ray_diffuse
    sx = (rays_uv[..., 0, :] * float(Wl)).clamp(max=Wl - 1)
    sy = (rays_uv[..., 1, :] * float(Hl)).clamp(max=Hl - 1)
    if lp.shape[0] == 1:
        col = interpolate_bilinear(lp[0], sx, sy)
    else:
        col = torch.stack([interpolate_bilinear(lp[i], sx[i], sy[i]) for i in range(lp.shape[0])])
    col = col.permute(0, 3, 4, 1, 2)                                     # [N,R,3,H,W]
    ltt_s = (rays_lt[:, :n_spec] * col[:, :n_spec]).sum(1) / n_spec
    out_s = ltt_s if no_albedo else albedo_specular * ltt_s
    if num_ray_diffuse > 0:
        ltt_d = (rays_lt[:, n_spec:] * col[:, n_spec:]).sum(1) / num_ray_diffuse
        if no_albedo:
            out_d = ltt_d
        else:
            out_d = (albedo_diffuse if seperate_albedo else albedo_specular) * ltt_d
    else:
        ltt_d = torch.zeros_like(ltt_s)
        out_d = torch.zeros_like(out_s)
    return out_s + out_d, out_s, out_d, ltt_s, ltt_d, col


# ------------------------------------------------------------------------------------------------
# network.RaysLTChromLoss  (network.py:395-411)
# ------------------------------------------------------------------------------------------------
def rays_lt_chrom_loss(rays_lt, alpha_map, img=None):
    chrom = F.normalize(rays_lt, dim=2)
    mean = F.normalize(chrom.mean(dim=1, keepdim=True), dim=2)
    diff = (1 - (chrom * mean).sum(2)) * alpha_map
    if img is not None:
        diff = diff * (img.norm(dim=1, keepdim=True) * 20).clamp(max=1.0)
    return diff.sum() / alpha_map.sum() / diff.shape[1], chrom, mean, diff


# ------------------------------------------------------------------------------------------------
# sph_harm  (sph_harm.py:41-102).  evaluate_sh_basis calls pyshtools==4.5 (environment.yml:143), which is
# NOT in /root/reference: restated from its published definition -- real, 4pi-orthonormal ('ortho'),
# csphase=1 (no Condon-Shortley phase), order (l, m=-l..l), m<0 -> sin(|m| phi).  PARITY UNPINNED for this
# function (no reference fixture exists); pinned only by closed-form known answers and Gram-matrix
# orthonormality (tests/test_sh.py).
# ------------------------------------------------------------------------------------------------
def evaluate_sh_basis(lmax, directions):
    """directions [M,3] (numpy, any norm) -> [M,(lmax+1)^2] float64, following sph_harm.py:54-69."""
    d = np.asarray(directions)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    azi = np.arctan2(y, x)                                   # same dtype as the input (float32 in the callers)
    ele = np.arctan2(z, np.sqrt(x ** 2 + y ** 2))
    pol = (np.pi / 2.0 - ele).astype(d.dtype)
    azi = (azi * (180 / np.pi)).astype(d.dtype)
    pol = (pol * (180 / np.pi)).astype(d.dtype)
    phi = np.deg2rad(azi.astype(np.float64))
    theta = np.deg2rad(pol.astype(np.float64))
    ct, st = np.cos(theta), np.sin(theta)
    M = d.shape[0]
    # associated Legendre P_l^m(cos theta) without Condon-Shortley phase, stable upward recurrences
    P = np.zeros((lmax + 1, lmax + 1, M))
    P[0, 0] = 1.0
    for m in range(1, lmax + 1):
        P[m, m] = P[m - 1, m - 1] * (2 * m - 1) * st
    for m in range(0, lmax):
        P[m + 1, m] = (2 * m + 1) * ct * P[m, m]
    for m in range(0, lmax + 1):
        for l in range(m + 2, lmax + 1):
            P[l, m] = ((2 * l - 1) * ct * P[l - 1, m] - (l + m - 1) * P[l - 2, m]) / (l - m)
    out = np.zeros((M, (lmax + 1) ** 2))
    k = 0
    for l in range(lmax + 1):
        for m in range(-l, l + 1):
            am = abs(m)
            norm = math.sqrt((2 - (am == 0)) * (2 * l + 1) / (4 * math.pi) * math.factorial(l - am) / math.factorial(l + am))
            ang = np.cos(am * phi) if m >= 0 else np.sin(am * phi)
            out[:, k] = norm * P[l, am] * ang
            k += 1
    return out


def fit_sh_coeff(samples, basis):
    """sph_harm.py:74-88: coeff = 4pi/N * sum_s samples[s] * basis[s]."""
    w = 4.0 * np.pi / samples.shape[-2]
    if samples.dim() == 2:
        return (samples[:, None, :] * basis[:, :, None]).sum(-3) * w
    return (samples[:, :, None, :] * basis[None, :, :, None]).sum(-3) * w


def reconstruct_sh(coeff, basis):
    """sph_harm.py:91-102."""
    if coeff.dim() == 2:
        return (basis[..., None] * coeff[None]).sum(-2)
    return (basis[None, :, :, None] * coeff[:, None]).sum(-2)
