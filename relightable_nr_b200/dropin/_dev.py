"""Device staging for host tensors handed to the drop-in operators.

The reference scripts call a few operators with CPU tensors while they set things up (before
``module.to(device)``).  librnr_b200 has no CPU implementation, so such inputs are copied to the
current CUDA device, the kernel runs there, and results are copied back to where the first argument
lived -- the caller sees the reference's device semantics, the arithmetic is always the CUDA kernel."""
import torch


def on_cuda(*tensors):
    first = next(t for t in tensors if isinstance(t, torch.Tensor))
    home = first.device
    if not torch.cuda.is_available():
        raise RuntimeError('librnr_b200 needs a CUDA device; there is no CPU fallback')
    dev = home if home.type == 'cuda' else torch.device('cuda', torch.cuda.current_device())
    moved = tuple(t.to(dev) if isinstance(t, torch.Tensor) else t for t in tensors)

    def back(out):
        if home.type == 'cuda':
            return out
        if isinstance(out, (tuple, list)):
            return type(out)(o.to(home) if isinstance(o, torch.Tensor) else o for o in out)
        return out.to(home)

    return moved, back
