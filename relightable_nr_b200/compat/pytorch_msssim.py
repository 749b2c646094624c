"""``pytorch_msssim.ssim`` as metric.py:78-84 calls it (``ssim(X, Y, data_range=255, size_average=False)``): the standard
Gaussian-window SSIM of Wang et al. -- 11-tap window, sigma 1.5, K = (0.01, 0.03), separable 'valid' filtering, mean over
channels and positions per image.  Registered by the launcher only when the real package is absent; runs on whatever device
the tensors live on (the validation code of train_rnr.py:707-887 hands it CPU tensors)."""
import torch
import torch.nn.functional as F


def _gauss_1d(size, sigma, dtype, device):
    x = torch.arange(size, dtype=dtype, device=device) - size // 2
    g = torch.exp(-(x ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x, w):
    """Separable 'valid' Gaussian blur of [N,C,H,W] (dimensions shorter than the window are left unfiltered)."""
    C = x.shape[1]
    if x.shape[2] >= w.numel():
        x = F.conv2d(x, w.view(1, 1, -1, 1).expand(C, 1, -1, 1), groups=C)
    if x.shape[3] >= w.numel():
        x = F.conv2d(x, w.view(1, 1, 1, -1).expand(C, 1, 1, -1), groups=C)
    return x


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, K=(0.01, 0.03), nonnegative_ssim=False):
    if X.shape != Y.shape or X.dim() != 4:
        raise ValueError('ssim expects two [N,C,H,W] tensors of equal shape, got %s and %s' % (tuple(X.shape), tuple(Y.shape)))
    X, Y = X.float(), Y.float()
    w = win.flatten().to(X) if win is not None else _gauss_1d(win_size, win_sigma, X.dtype, X.device)
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _filter(X, w), _filter(Y, w)
    s11 = _filter(X * X, w) - mu1 * mu1
    s22 = _filter(Y * Y, w) - mu2 * mu2
    s12 = _filter(X * Y, w) - mu1 * mu2
    cs = (2 * s12 + C2) / (s11 + s22 + C2)
    val = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs
    per_channel = val.flatten(2).mean(-1)
    if nonnegative_ssim:
        per_channel = torch.relu(per_channel)
    return per_channel.mean() if size_average else per_channel.mean(1)


def ms_ssim(*a, **k):
    raise NotImplementedError('ms_ssim is not used by the reference (metric.py calls ssim only)')
