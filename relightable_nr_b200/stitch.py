"""Light-probe stitching on the device: the start-up step that turns the BACKGROUND of the training views
into the equirectangular probe ``LightingLP`` / ``LightingSH`` are initialised from (train_rnr.py:293-305 reads its output).

Reference: stitch_lp.py:95-159.  Per view it (a) rasterises the proxy mesh's silhouette with one ``cv2.fillPoly`` call per face,
dilates it at 512^2 and keeps the pixels left uncovered (:121-134), (b) turns every such pixel into a world-space ray
(camera2ray :27-35), the ray into probe coordinates (spherical_mapping :22-24) and (c) scatters the pixel colours into the
1600 x 3200 probe with numpy fancy indexing (:136-147), then divides by the hit count (:149-150).

Here (a) stays on the host with the SAME OpenCV calls -- it is the definition of the mask (fillPoly's edge rule, INTER_LINEAR
resizes) and costs a few milliseconds once the Python-level loop over faces is gone from the hot part --, (b) and (c) are two
kernels per view (csrc/stitch.cu) in fp64 that keep numpy's last-write-wins scatter rule, and the division is a third.
``python -m relightable_nr_b200.stitch`` takes the arguments of stitch_lp.py and writes the same four files.
"""
import argparse
import ctypes as C
import os

import numpy as np
import torch

from . import _lib

vp, i32 = C.c_void_p, C.c_int
_lib.register_sigs({"rnr_stitch_view": [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp],
                    "rnr_stitch_finish": [vp, vp, vp, i32, i32, vp]})


def selected_views(num_view, sampling_pattern):
    """View indices stitch_lp.py:104-120 keeps: 'all', 'skip_k' (every k-th), 'skipinv_k' (all but every k-th), 'first_k'."""
    out = []
    for i in range(num_view):
        if sampling_pattern[:5] == 'skip_' and i % int(sampling_pattern.split('_')[-1]) != 0:
            continue
        if sampling_pattern[:8] == 'skipinv_' and i % int(sampling_pattern.split('_')[-1]) == 0:
            continue
        if sampling_pattern[:6] == 'first_' and i >= int(sampling_pattern.split('_')[-1]):
            continue
        out.append(i)
    return out


def background_mask(vertices_h, faces, pose, proj, img_h, img_w):
    """stitch_lp.py:121-134: silhouette of the proxy (vertices_h [4, nv] homogeneous world positions) filled face by face at
    integer pixel positions, dilated by a 17 x 17 box at 512^2, resized back; True where nothing of it is left."""
    import cv2
    v = proj.dot(pose.dot(vertices_h)[:3])
    v[0] /= v[2]
    v[1] /= v[2]
    v = v.astype('int32')
    v[v < 0] = 0
    v[0, v[0] > img_w - 1] = img_w - 1
    v[1, v[1] > img_h - 1] = img_h - 1
    tri = np.ascontiguousarray(v[:2].T[faces])              # [nf, 3, 2]
    m = np.zeros((img_h, img_w))
    for t in tri:
        cv2.fillPoly(m, [t], 255)
    m = cv2.resize(cv2.dilate(cv2.resize(m, (512, 512)), np.ones((17, 17), np.uint8)), (img_w, img_h))
    return m == 0


class ProbeStitcher:
    """Accumulates views into one probe.  ``add_view`` takes the image as the reference reads it ([h, w, >=3] float, 0..1)."""

    def __init__(self, lp_h=1600, lp_w=3200, device='cuda'):
        self.lp_h, self.lp_w, self.device = int(lp_h), int(lp_w), torch.device(device)
        if self.device.type != 'cuda':
            raise TypeError('ProbeStitcher: a CUDA device is required (librnr_b200 has no CPU path)')
        self.env = torch.zeros((self.lp_h, self.lp_w, 3), dtype=torch.float64, device=self.device)
        self.count = torch.zeros((self.lp_h, self.lp_w, 3), dtype=torch.float32, device=self.device)
        self.winner = torch.full((self.lp_h * self.lp_w,), -1, dtype=torch.int32, device=self.device)
        self._texel = None

    def add_view(self, img, bg_mask, pose, proj):
        """img [h, w, 3+] float32, bg_mask [h, w] bool, pose [4, 4] world->camera (global_RT already folded in), proj [3, 3]."""
        h, w = bg_mask.shape
        img_d = torch.as_tensor(np.ascontiguousarray(img[:, :, :3], dtype=np.float32)).to(self.device)
        bg_d = torch.as_tensor(np.ascontiguousarray(bg_mask, dtype=np.uint8)).to(self.device)
        if self._texel is None or self._texel.numel() < h * w:
            self._texel = torch.empty(h * w, dtype=torch.int32, device=self.device)
        kinv = np.ascontiguousarray(np.linalg.inv(np.asarray(proj, dtype=np.float64)))
        rinv = np.ascontiguousarray(np.linalg.inv(np.asarray(pose, dtype=np.float64)[:3, :3]))
        _lib.check(_lib.lib().rnr_stitch_view(img_d.data_ptr(), bg_d.data_ptr(), kinv.ctypes.data, rinv.ctypes.data, h, w, self.lp_h,
                                              self.lp_w, self._texel.data_ptr(), self.winner.data_ptr(), self.env.data_ptr(),
                                              self.count.data_ptr(), torch.cuda.current_stream().cuda_stream), 'rnr_stitch_view')

    def finish(self):
        """-> (env [lp_h, lp_w, 3] float64, mask [lp_h, lp_w] bool, count [lp_h, lp_w, 3] float32) as numpy arrays."""
        mask = torch.empty((self.lp_h, self.lp_w), dtype=torch.uint8, device=self.device)
        _lib.check(_lib.lib().rnr_stitch_finish(self.env.data_ptr(), self.count.data_ptr(), mask.data_ptr(), self.lp_h, self.lp_w,
                                                torch.cuda.current_stream().cuda_stream), 'rnr_stitch_finish')
        return self.env.cpu().numpy(), mask.cpu().numpy() > 0, self.count.cpu().numpy()


def read_obj_geometry(path):
    """(vertices [nv, 3] float64, faces [nf, 3] int) of a triangle OBJ: `v` records and the position index of every `f` corner --
    what ``trimesh.load(path, process=False)`` exposes as .vertices / .faces up to vertex duplication, which a filled silhouette
    does not see."""
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            p = line.split()
            if not p:
                continue
            if p[0] == 'v':
                vs.append([float(x) for x in p[1:4]])
            elif p[0] == 'f':
                idx = [int(tok.split('/')[0]) for tok in p[1:]]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    fs.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(vs, dtype=np.float64), np.asarray(fs, dtype=np.int64)


def stitch_scene(calib, vertices, faces, read_image, sampling_pattern='skipinv_10', lp_h=1600, lp_w=3200, device='cuda'):
    """stitch_lp.py:86-150 for one lighting: ``calib`` = the loaded calib.mat dict, ``read_image(i)`` -> [h, w, 3+] float image of
    view i.  Returns (env, mask, count, num_view)."""
    poses, projs, img_hws = calib['poses'], calib['projs'], calib['img_hws']
    global_RT = calib['global_RT']
    global_RT_inv = np.linalg.inv(global_RT)
    vh = global_RT.dot(np.hstack((vertices, np.ones((vertices.shape[0], 1)))).T)
    st = ProbeStitcher(lp_h, lp_w, device)
    for i in selected_views(poses.shape[0], sampling_pattern):
        img_w, img_h = int(img_hws[i, 1]), int(img_hws[i, 0])
        pose = poses[i].dot(global_RT_inv)
        bg = background_mask(vh, faces, pose, projs[i], img_h, img_w)
        st.add_view(read_image(i), bg, pose, projs[i])
    env, mask, count = st.finish()
    return env, mask, count, int(poses.shape[0])


def main(argv=None):
    import cv2
    import scipy.io
    ap = argparse.ArgumentParser(description='device version of stitch_lp.py (same arguments, same output files)')
    ap.add_argument('--data_root', type=str, default='./data/material_sphere')
    ap.add_argument('--calib_fp', type=str, default='_/calib.mat')
    ap.add_argument('--obj_fp', type=str, default='_/mesh.obj')
    ap.add_argument('--lighting_idx', default=0, type=int)
    ap.add_argument('--sampling_pattern', type=str, default='skipinv_10')
    ap.add_argument('--img_suffix', type=str, default='.exr')
    ap.add_argument('--lp_h', type=int, default=1600)
    ap.add_argument('--lp_w', type=int, default=3200)
    opt = ap.parse_args(argv)
    os.environ.setdefault('OPENCV_IO_ENABLE_OPENEXR', '1')
    if opt.calib_fp[:2] == '_/':
        opt.calib_fp = os.path.join(opt.data_root, opt.calib_fp[2:])
    if opt.obj_fp[:2] == '_/':
        opt.obj_fp = os.path.join(opt.data_root, opt.obj_fp[2:])
    img_dir = os.path.join(opt.data_root, 'rgb' + str(opt.lighting_idx))
    out = os.path.join(opt.data_root, 'light_probe_stitch_' + opt.sampling_pattern)
    for d in (out, os.path.join(out, 'mask'), os.path.join(out, 'count')):
        os.makedirs(d, exist_ok=True)

    def read_image(i):
        if opt.img_suffix == '.exr':
            return cv2.imread(img_dir + ('/%03d' % i) + opt.img_suffix, cv2.IMREAD_ANYCOLOR | cv2.IMREAD_ANYDEPTH)
        return cv2.imread(img_dir + ('/%06d' % i) + opt.img_suffix, cv2.IMREAD_UNCHANGED).astype(np.float32)[:, :, :3] / 255.

    v, f = read_obj_geometry(opt.obj_fp)
    env, mask, count, num_view = stitch_scene(scipy.io.loadmat(opt.calib_fp), v, f, read_image, opt.sampling_pattern, opt.lp_h, opt.lp_w)
    k = str(opt.lighting_idx)
    cv2.imwrite(os.path.join(out, k + '.png'), (env * 255).astype('uint8'))
    cv2.imwrite(os.path.join(out, k + '.exr'), env.astype(np.float32))
    cv2.imwrite(os.path.join(out, 'mask', k + '.png'), (mask * 255).astype('uint8'))
    cv2.imwrite(os.path.join(out, 'count', k + '.png'), (count / float(num_view) * 255.0).astype('uint8'))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
