"""The per-view precomputation of the reference as ONE device pass, and the map cache behind the training loader.

Reference: precompute.py:140-253 rasterises every training view, derives eight maps from the G-buffer and round-trips each of
them through numpy, pyshtools (CPU) and one ``.mat`` file per map per view; training re-loads them through scipy.io.loadmat for
every view (dataio.py:219-245).  Here:

* ``ViewPrecompute``  -- rasterise + TBN / view-direction / tangent-space view direction / SH basis / reflection direction for a batch
  of cameras entirely on the device (the kernels of csrc/raster.cu, gbuffer.cu, sh.cu); ``maps()`` returns exactly the tensors a
  ``dataio.ViewDataset`` item holds after ``load_precompute`` (same names, shapes, dtypes, the ``uv - floor(uv)`` of dataio.py:228
  included), ready for ``RNRPipeline.train_step``; ``write_reference_layout()`` emits the reference's on-disk layout so that the
  unchanged training scripts can read it;
* ``PackedViewCache`` -- all maps of all views in ONE binary file (a JSON header + raw little-endian arrays, views back to back),
  memory-mapped and staged through pinned buffers on a copy stream: the loader that replaces eight ``loadmat`` calls per view.

Both are outside the timed step (the callers on the other side of the hot path).
"""
import json
import os

import numpy as np
import torch

from .dropin import camera, network, render, sph_harm

#: (name, trailing shape, numpy dtype) of the per-view maps of dataio.py:219-245 the RNR / DNR steps consume
VIEW_MAPS = (('uv_map', (2,), 'float32'), ('sh_basis_map', (9,), 'float32'), ('normal_map', (3,), 'float32'),
             ('view_dir_map', (3,), 'float32'), ('view_dir_map_tangent', (3,), 'float32'), ('TBN_map', (3, 3), 'float32'),
             ('reflect_dir_map', (3,), 'float32'), ('alpha_map', (), 'float32'))


class ViewPrecompute:
    """precompute.py:58-64,140-253 on the device.  ``obj_fp``: the proxy mesh; ``img_size``: render size."""

    def __init__(self, obj_fp, img_size, global_RT=None, device='cuda'):
        self.device = torch.device(device)
        self.img_size = int(img_size)
        self.rasterizer = network.Rasterizer(obj_fp=obj_fp, img_size=self.img_size, global_RT=global_RT).to(self.device).eval()

    @torch.no_grad()
    def maps(self, proj, pose, proj_inv=None, R_inv=None, img_gt=None, with_raster=False):
        """proj [N,3,3], pose [N,4,4] (world->camera) -> dict of per-view maps on the device (batch-first, fp32).

        Follows precompute.py:152-247: rasterise; TBN from the per-face tangents; world-space view direction per pixel; its
        tangent-space version normalised; degree-2 SH basis of the view direction; reflection of the view direction about the
        normal, masked by alpha.  ``img_gt`` [N,3,H,W] in [0,1] (optional) reproduces the "removed padded regions" mask of
        precompute.py:191 (alpha *= img_gt[0] <= 2)."""
        dev = self.device
        proj, pose = proj.to(dev).float(), pose.to(dev).float()
        if proj_inv is None:
            proj_inv = torch.inverse(proj)
        if R_inv is None:
            R_inv = pose[:, :3, :3].transpose(1, 2).contiguous()
        S = self.img_size
        r = self.rasterizer(proj=proj, pose=pose, dist_coeffs=None, offset=None, scale=None)
        uv_map, alpha_map, fim, weight_map, faces_v_idx, normal_map, _, faces_v, faces_vt = r[:9]
        TBN = render.get_TBN_map(normal_map, fim, faces_v=faces_v[0], faces_texcoord=faces_vt[0], tangent=None)
        if img_gt is not None:
            alpha_map = alpha_map * (img_gt[:, 0].to(dev) * 255.0 <= 2.0 * 255).to(alpha_map.dtype)
        view_dir, _ = camera.get_view_dir_map((S, S), proj_inv.to(dev).float(), R_inv.to(dev).float())
        vdt = torch.matmul(TBN.reshape(-1, 3, 3).transpose(-2, -1), view_dir.reshape(-1, 3, 1))[..., 0].reshape(view_dir.shape)
        vdt = torch.nn.functional.normalize(vdt, dim=-1)
        sh = sph_harm.evaluate_sh_basis_l2(view_dir.contiguous())
        refl = camera.get_reflect_dir(view_dir, normal_map) * alpha_map[..., None]
        out = {'uv_map': uv_map - torch.floor(uv_map), 'sh_basis_map': sh, 'normal_map': normal_map, 'view_dir_map': view_dir,
               'view_dir_map_tangent': vdt, 'TBN_map': TBN, 'reflect_dir_map': refl, 'alpha_map': alpha_map}
        if with_raster:
            out['raster'] = {'face_index_map': fim, 'weight_map': weight_map, 'faces_v_idx': faces_v_idx, 'v_uvz': r[12], 'v_front_mask': r[13]}
        return {k: (v.contiguous() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}

    def write_reference_layout(self, precomp_dir, names, maps):
        """Write ``maps`` (as returned by ``maps(..., with_raster=True)`` for the views ``names``) in the layout of
        precompute.py:86-138 that dataio.py:219-245 reads: <precomp_dir>/resol_<S>/<map>/<name>.mat (+ alpha_map/<name>.png)."""
        import cv2
        import scipy.io
        root = os.path.join(precomp_dir, 'resol_%d' % self.img_size)
        for i, name in enumerate(names):
            for key in ('TBN_map', 'uv_map', 'normal_map', 'view_dir_map', 'view_dir_map_tangent', 'sh_basis_map', 'reflect_dir_map'):
                d = os.path.join(root, key)
                os.makedirs(d, exist_ok=True)
                scipy.io.savemat(os.path.join(d, name + '.mat'), {key: maps[key][i].cpu().numpy()})
            d = os.path.join(root, 'alpha_map')
            os.makedirs(d, exist_ok=True)
            cv2.imwrite(os.path.join(d, name + '.png'), maps['alpha_map'][i].cpu().numpy() * 255)
            if 'raster' in maps:
                d = os.path.join(root, 'raster')
                os.makedirs(d, exist_ok=True)
                rs = maps['raster']
                scipy.io.savemat(os.path.join(d, name + '.mat'), {
                    'face_index_map': rs['face_index_map'][i].cpu().numpy(), 'weight_map': rs['weight_map'][i].cpu().numpy(),
                    'faces_v_idx': rs['faces_v_idx'][0].cpu().numpy(), 'v_uvz': rs['v_uvz'][i].cpu().numpy(),
                    'v_front_mask': rs['v_front_mask'][min(i, rs['v_front_mask'].shape[0] - 1)].cpu().numpy()})


class PackedViewCache:
    """All per-view maps of a scene in one file: ``b'RNRCACHE'``, a little-endian uint64 header length, a JSON header
    {img_size, n_views, names, maps: [(name, shape, dtype, offset within a view record)], record_bytes}, then the view records back to
    back (every array 64-byte aligned inside its record).  Reading is a memory map: a view is one contiguous slice."""

    MAGIC = b'RNRCACHE'

    def __init__(self, path):
        self.path = path
        with open(path, 'rb') as fh:
            if fh.read(8) != self.MAGIC:
                raise ValueError('%s is not a packed view cache' % path)
            n = int(np.frombuffer(fh.read(8), dtype='<u8')[0])
            self.header = json.loads(fh.read(n).decode())
            self.data_offset = (16 + n + 63) // 64 * 64
        self.n_views = int(self.header['n_views'])
        self.record_bytes = int(self.header['record_bytes'])
        self._mm = np.memmap(path, dtype=np.uint8, mode='r', offset=self.data_offset, shape=(self.n_views, self.record_bytes))
        self._pinned = None

    @classmethod
    def write(cls, path, views, names=None):
        """``views``: list of dicts of per-view tensors / arrays WITHOUT the batch dimension (every view the same keys and shapes)."""
        first = views[0]
        entries, off = [], 0
        for k in sorted(first.keys()):
            a = first[k].detach().cpu().numpy() if isinstance(first[k], torch.Tensor) else np.asarray(first[k])
            entries.append((k, list(a.shape), str(a.dtype), off))
            off = (off + a.nbytes + 63) // 64 * 64
        header = {'n_views': len(views), 'names': list(names) if names is not None else ['%05d' % i for i in range(len(views))],
                  'maps': entries, 'record_bytes': off}
        hb = json.dumps(header).encode()
        with open(path, 'wb') as fh:
            fh.write(cls.MAGIC)
            fh.write(np.array([len(hb)], dtype='<u8').tobytes())
            fh.write(hb)
            fh.write(b'\0' * ((16 + len(hb) + 63) // 64 * 64 - 16 - len(hb)))
            rec = np.zeros(off, dtype=np.uint8)
            for v in views:
                rec[:] = 0
                for k, shape, dt, o in entries:
                    a = v[k].detach().cpu().numpy() if isinstance(v[k], torch.Tensor) else np.asarray(v[k])
                    a = np.asarray(a.astype(dt, copy=False), order='C')     # (np.ascontiguousarray would turn a 0-d array into [1])
                    if list(a.shape) != shape:
                        raise ValueError('view map %s has shape %s, expected %s' % (k, list(a.shape), shape))
                    rec[o:o + a.nbytes] = np.frombuffer(a.tobytes(), dtype=np.uint8)
                fh.write(rec.tobytes())
        return cls(path)

    def __len__(self):
        return self.n_views

    def view_numpy(self, i):
        rec = self._mm[i]
        return {k: rec[o:o + int(np.prod(shape, dtype=np.int64)) * np.dtype(dt).itemsize].view(dt).reshape(shape)
                for k, shape, dt, o in self.header['maps']}

    def load(self, i, device='cuda', stream=None, slot=0):
        """View ``i`` as a dict of device tensors with batch dimension 1: ONE host->device copy of the whole record from a pinned
        staging buffer (two slots for double buffering), then zero-copy views into it.  With ``stream`` the copy is enqueued there
        (the caller orders its consumer with an event); without it the copy is synchronous."""
        if self._pinned is None:
            self._pinned = [torch.empty(self.record_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
            self._dev = {}
        pin = self._pinned[slot % 2]
        pin.numpy()[:] = self._mm[i]
        key = (str(device), slot % 2)
        if key not in self._dev:
            self._dev[key] = torch.empty(self.record_bytes, dtype=torch.uint8, device=device)
        buf = self._dev[key]
        if stream is not None:
            with torch.cuda.stream(stream):
                buf.copy_(pin, non_blocking=True)
        else:
            buf.copy_(pin)
        out = {}
        for k, shape, dt, o in self.header['maps']:
            n = int(np.prod(shape, dtype=np.int64)) * np.dtype(dt).itemsize
            out[k] = buf[o:o + n].view(getattr(torch, dt)).view(*shape)[None]
        return out
