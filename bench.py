#!/usr/bin/env python
"""Benchmarks of the per-view deferred-relighting path (BASELINE.json: views/sec at 512^2, fwd+bwd, 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config rnr_train|rnr_infer|rnr_relight|dnr_train]

Default workload (config.workload): BASELINE.json configs[2] restricted to one box -- the body of train_rnr.py:490-623 (texture
sample -> 26 rays/pixel -> 108->78-channel U-Net -> SH envmap ray render -> 4 losses -> backward -> Adam) on synthetic 512x512
views of the material-sphere proxy, 1 view per GPU per step (the reference's own constraint), views sharded over
ranks (weak scaling), one NCCL sum of the gradients per step.  Other configs: rnr_infer (configs[1]: test_rnr.py:265-393 incl.
rasterisation), rnr_relight (configs[4]: one U-Net pass + 4 envmap renders per view), dnr_train (configs[3]: train_dnr.py:240-275).

Prints ONE JSON line (rank 0).  ``value`` = views/s with the per-view maps resident in HBM; ``e2e`` = the same step fed from pinned
host buffers (H2D of the per-view maps + D2H of the loss inside the timed region); ``roofline`` = live conv FLOPs of the tcgen05
implicit-GEMM kernel (conv_halo_kernel) / its CUDA-event time / measured bf16 burst peak; ``cpu_baseline`` = the reference's own
PyTorch modules (staged under baseline/_ref by tools/stage_reference.py) on the host cores -- kind "reference".
``--impl reference`` times that CPU path alone, on the same synthetic views the GPU arm sees.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'views/sec at 512^2 (fwd+bwd)'
UNIT = 'views/s'


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d.get('bf16_tflops_sustained', d['bf16_tflops']), src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.1)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def _workload(args, world):
    size = args.size
    return {
        'rnr_train': 'RNR train step (train_rnr.py:490-623), %dx%d material-sphere proxy views, 1 view/GPU/step, texture 512^2x24ch x4 mips, '
                     'U-Net 108->78 nf0=64, 26 rays, SH lmax 10 envmap 256x512' % (size, size),
        'rnr_infer': 'RNR inference (test_rnr.py:265-393): rasterise the 65 536-face proxy + TBN / view-dir / SH maps + texture + 26 rays + U-Net '
                     'forward + SH envmap render, %dx%d, 1 view per step per GPU' % (size, size),
        'rnr_relight': 'RNR relighting (test_rnr.py:265-393 over spiral_step720 x 4 envmaps): per view rasterise + maps + ONE U-Net forward + 4 '
                       'envmap renders, %dx%d, views sharded over GPUs, no collective' % (size, size),
        'dnr_train': 'DNR train step (train_dnr.py:240-275): 16-ch 512^2 x4-mip neural texture -> U-Net 16->3 nf0=80 -> masked L1 -> backward -> '
                     'Adam, %dx%d, 1 view/GPU/step' % (size, size),
    }[args.config]


def _config(args, world, impl='ours'):
    cfg = {'workload': _workload(args, world), 'views_per_step': max(world, 1)}
    if impl == 'reference':
        cfg.update({'parallelism': 'host cores only (the reference\'s own PyTorch modules from baseline/_ref, CPU, fp32)',
                    'launch': 'python, eager', 'step': 'reference modules called in the order of the script body',
                    'l2': 'n/a (CPU)'})
        return cfg
    train = args.config in ('rnr_train', 'dnr_train')
    cfg.update({
        'parallelism': ('dp%d (views sharded, NCCL gradient sum, 1/world inside the fused Adam)' if train else 'dp%d (views sharded, no collective)') % max(world, 1),
        'launch': 'eager' if (args.no_graph or args.config != 'rnr_train') else 'one CUDA graph per step',
        'step': ('module-by-module (drop-in operator API)' if (args.no_fused or args.config == 'dnr_train') else
                 'fused head/tail kernels around the U-Net, optimiser fused into the backward pass (relightable_nr_b200/fused.py)'),
        'l2': 'per-step working set (~3 GB of activations/gradients at 512^2) >> 126 MB L2; 4 distinct views cycled'})
    return cfg


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules (baseline/_ref), driven like the script body
# ----------------------------------------------------------------------------------------------------------------------
def _sphere_samples():
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'sphere_samples_4096.npz'))['sphere_samples']
    return torch.from_numpy(z.astype(np.float32)).t().contiguous()            # [3, 4096]  (train_rnr.py:167-168)


def _crop_view(view, c):
    """Centre c x c window of a per-view map dict (bounded CPU sample)."""
    H = view['alpha_map'].shape[1]
    a = (H - c) // 2
    out = {}
    for k, v in view.items():
        if k == 'img_gt':
            out[k] = v[:, :, a:a + c, a:a + c].contiguous()
        else:
            out[k] = v[:, a:a + c, a:a + c].contiguous()
    return out


class ReferenceRNR:
    """train_rnr.py:245-376 (module set) and :490-623 (iteration) with the REAL reference classes on the CPU."""

    def __init__(self, with_gcn=True):
        import types
        import torch
        from tests.golden import ref_import
        if not ref_import.available():
            raise RuntimeError('reference not staged: run tools/stage_reference.py in the build container')
        ref = ref_import.import_reference()
        self.ref, self.torch = ref, torch
        torch.manual_seed(0)
        l_dir = _sphere_samples()
        N = ref.network
        self.tm = N.TextureMapper(texture_size=512, texture_num_ch=24, mipmap_level=4, apply_sh=True)
        init = torch.randn((1, 121, 3), generator=torch.Generator().manual_seed(11)) * 0.1
        init[:, 0] = 1.0
        self.lm = N.LightingSH(l_dir, lmax=10, num_lighting=1, num_channel=3, init_coeff=init, lp_recon_h=256, lp_recon_w=512)
        self.rs = N.RaySampler(num_azi=6, num_polar=2, interval_polar=5)
        self.rsd = N.RaySampler(num_azi=6, num_polar=2, interval_polar=10, mode='diffuse')
        self.R = self.rs.num_ray + self.rsd.num_ray
        self.net = N.RenderingNet(nf0=64, in_channels=self.R * 3 + 6 + 24, out_channels=3 * self.R, num_down_unet=5, out_channels_gcn=512)
        self.rr = N.RayRenderer(self.lm, N.Interpolater())
        self.chrom = N.RaysLTChromLoss()
        self.gcn = None
        params = list(self.tm.parameters()) + list(self.lm.parameters()) + list(self.net.parameters())
        if with_gcn:
            sys.path.insert(0, os.path.join(ROOT, 'tools'))
            import make_scene
            v = torch.tensor(make_scene.grid_sphere(75, 100)[0], dtype=torch.float32)
            opt = types.SimpleNamespace(n_filters=64, kernel_size=16, act_type='relu', norm_type='batch', bias=True, epsilon=0.2,
                                        stochastic=True, conv_type='edge', n_blocks=20, num_v_gcn=7500, out_channels_gcn=512, in_channels=6,
                                        block_type='res')
            self.gcn = N.DenseDeepGCN(opt)
            self.gcn_input = types.SimpleNamespace(pos=v, x=v)
            params = list(self.gcn.parameters()) + params
        for m in (self.tm, self.lm, self.rs, self.rsd, self.net, self.rr) + ((self.gcn,) if self.gcn is not None else ()):
            m.train()
        self.opt = torch.optim.Adam(params, lr=1e-3)
        self.opt.zero_grad()
        with torch.no_grad():
            self.l_init = ref.sph_harm.reconstruct_sh(self.lm.coeff.data[0], self.lm.basis_val).clone()
            self.l_init += 0.05 * torch.randn(self.l_init.shape, generator=torch.Generator().manual_seed(31))
            self.l_mask = torch.ones(4096, dtype=torch.bool)
            self.l_mask[::7] = False

    def step(self, view):
        """One iteration; returns (seconds of the pixel-proportional part, seconds of the per-iteration fixed part)."""
        torch, ref = self.torch, self.ref
        t0 = time.time()
        v_feature = self.gcn(self.gcn_input) if self.gcn is not None else None          # train_rnr.py:490
        coeff = self.lm.get_lighting_params(0)
        l_est = ref.sph_harm.reconstruct_sh(coeff, self.lm.basis_val)
        m = self.l_mask
        loss_small = (self.l_init[m] - l_est[m]).abs().sum() / m.float().sum() + (self.l_init[~m] - l_est[~m]).abs().sum() / (~m).float().sum() * 0.1
        for c0 in (3, 0):
            tex = self.tm.flatten_mipmap(start_ch=c0, end_ch=c0 + 3)
            valid = (tex != self.tm.tex_flatten_mipmap_init[..., c0:c0 + 3]).any(dim=-1, keepdim=True).to(tex.dtype)
            if valid.sum() > 0:
                loss_small = loss_small + ((tex * valid).sum(dim=(0, 1, 2)) / valid.sum(dim=(0, 1, 2)) - 0.5).abs().sum() / 3
        loss_small.backward()
        t_fix = time.time() - t0
        t0 = time.time()
        alpha = view['alpha_map'][:, None]
        N, _, H, W = alpha.shape
        neural = self.tm(view['uv_map'], view['sh_basis_map'], sh_start_ch=6)
        a_last = alpha.permute(0, 2, 3, 1)
        d0, uv0, _ = self.rs(view['TBN_map'], view['view_dir_map_tangent'], a_last)
        d1, uv1, _ = self.rsd(view['TBN_map'], view['view_dir_map_tangent'], a_last)
        rays_dir, rays_uv = torch.cat((d0, d1), -1), torch.cat((uv0, uv1), -1)
        x = torch.cat((rays_dir.permute(0, -1, -2, 1, 2).reshape(N, -1, H, W), view['normal_map'].permute(0, 3, 1, 2),
                       view['view_dir_map'].permute(0, 3, 1, 2), neural), 1)
        lt = (self.net(x, v_feature).reshape(N, self.R, -1, H, W) * 0.5 + 0.5) * 2.0
        out = self.rr(neural[:, 3:6], rays_uv, lt, lighting_idx=0, albedo_diffuse=neural[:, :3], num_ray_diffuse=d1.shape[-1], seperate_albedo=True)[0]
        a = alpha[:, :, 5:-5, 5:-5]
        loss = torch.nn.functional.l1_loss((out[:, :, 5:-5, 5:-5] * a).reshape(-1), (view['img_gt'][:, :, 5:-5, 5:-5] * a).reshape(-1))
        loss = loss + self.chrom(lt, alpha, view['img_gt'])[0]
        loss.backward()
        t_px = time.time() - t0
        t0 = time.time()
        self.opt.step()
        self.opt.zero_grad()
        t_fix += time.time() - t0
        return t_px, t_fix


class ReferenceDNR:
    """train_dnr.py:138-193, 240-275 with the real reference classes on the CPU."""

    def __init__(self):
        import torch
        from tests.golden import ref_import
        ref = ref_import.import_reference()
        self.torch = torch
        torch.manual_seed(0)
        self.tm = ref.network.TextureMapper(texture_size=512, texture_num_ch=16, mipmap_level=4, apply_sh=True)
        self.net = ref.network.RenderingNet(nf0=80, in_channels=16, out_channels=3, num_down_unet=5, use_gcn=False)
        self.tm.train(); self.net.train()
        self.opt = torch.optim.Adam(list(self.tm.parameters()) + list(self.net.parameters()), lr=1e-3)

    def step(self, view):
        torch = self.torch
        t0 = time.time()
        out = (self.net(self.tm(view['uv_map'], view['sh_basis_map']), None) * 0.5 + 0.5) * 2.0
        a = view['alpha_map'][:, None, 5:-5, 5:-5]
        loss = torch.nn.functional.l1_loss((out[:, :, 5:-5, 5:-5] * a).reshape(-1), (view['img_gt'][:, :, 5:-5, 5:-5] * a).reshape(-1))
        self.opt.zero_grad()
        loss.backward()
        t_px = time.time() - t0
        t0 = time.time()
        self.opt.step()
        return t_px, time.time() - t0


def cpu_reference_rate(args, steps, warmup, crop=128):
    """views/s of the reference's own modules on all host cores.  Bounded sample: every step processes the centre ``crop``^2 window of a
    512^2 view (the synthetic_view tensors the GPU arm sees) through the pixel-proportional part of the iteration, and the
    per-iteration fixed part (GCN forward over the 7500-vertex mesh, lighting / albedo losses, Adam over all parameters) in full;
    per-view time = fixed + pixel part x (512/crop)^2."""
    import torch
    from relightable_nr_b200.pipeline import synthetic_view
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    size = args.size
    crop = min(crop, size)
    model = ReferenceDNR() if args.config == 'dnr_train' else ReferenceRNR(with_gcn=True)
    views = [_crop_view(synthetic_view(size, view_idx=7 * i, device='cpu'), crop) for i in range(2)]
    frac = (crop * crop) / float(size * size)
    for i in range(max(warmup, 1)):
        model.step(views[i % 2])
    tp = tf = 0.0
    n = 0
    t_begin = time.time()
    while n < max(steps, 1):
        a, b = model.step(views[n % 2])
        tp += a; tf += b
        n += 1
        if time.time() - t_begin > args.cpu_budget and n >= 1:
            break
    per_view = tf / n + (tp / n) / frac
    what = 'train_dnr.py:240-275' if args.config == 'dnr_train' else 'train_rnr.py:490-623 incl. the GCN forward of :490'
    sample = ('reference modules from baseline/_ref (%s); per step: centre %dx%d window of a %dx%d synthetic_view through texture -> U-Net -> '
              'render -> pixel losses -> backward (%.2f s, scaled by 1/%.4f) + the per-iteration fixed part in full (%.2f s: %s); %d timed steps'
              % (what, crop, crop, size, size, tp / n, frac, tf / n,
                 'Adam' if args.config == 'dnr_train' else 'GCN forward over 7500 vertices, lighting + albedo-mean losses, Adam', n))
    return 1.0 / per_view, dict(cores=cores, kind='reference', sample=sample, ms_per_step=per_view * 1e3, torch_threads=torch.get_num_threads())


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if args.config in ('rnr_infer', 'rnr_relight'):
        print(json.dumps({'impl': 'reference', 'unavailable': 'the reference rasterizer (neural_renderer.cuda) has no CPU path; '
                          'see the cpu_baseline of --config rnr_train for the CPU arm of the network path'}))
        return
    v, info = cpu_reference_rate(args, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': info['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': _config(args, 0, 'reference'),
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': info['cores'], 'kind': info['kind'], 'sample': info['sample']},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
class _Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py needs a CUDA device: librnr_b200 has no CPU path (use --impl reference for the CPU arm)')
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K calls of fn(i) between CUDA events, barrier + synchronize on both sides, max over ranks (ms)."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item()

    def teardown(self):
        # destroy_process_group() blocks forever while a captured CUDA graph still references the communicator's work
        # (measured: both ranks parked in it after a complete run), so the ranks leave through os._exit once rank 0 has
        # printed its line; the process exit tears NCCL down.
        if self.world > 1:
            sys.stdout.flush()
            sys.stderr.flush()
            self.torch.cuda.synchronize()
            self.dist.barrier()
            os._exit(0)


def _e2e_runner(D, host_views, step):
    """Double-buffered end-to-end step: pinned host maps -> device (copy stream, prefetching the next view) -> step -> loss.item()."""
    torch = D.torch
    dev = D.dev
    nviews = len(host_views)
    dev_bufs = [{k: torch.empty_like(v, device=dev) for k, v in host_views[0].items()} for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    last = [None]

    def upload(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[b])
            for k, v in host_views[i % nviews].items():
                dev_bufs[b][k].copy_(v, non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_step(i):
        b = i % 2
        if i == 0:
            upload(0)
        upload(i + 1)                       # prefetch the next view while this one computes (both inside the timed region)
        torch.cuda.current_stream().wait_event(ready[b])
        out = step(dev_bufs[b])
        done[b].record()
        last[0] = float(out.reshape(-1)[0].item())          # D2H of the step's result, every step

    def reset():
        torch.cuda.synchronize()
        for b in range(2):
            done[b].record()

    return e2e_step, reset, last


def run_rnr_train(args, D):
    torch, dist = D.torch, D.dist
    world, rank, dev = D.world, D.rank, D.dev
    from relightable_nr_b200 import _lib
    from relightable_nr_b200.pipeline import RNRPipeline, synthetic_view
    L = _lib.lib()
    pipe = RNRPipeline(device=dev, img_size=args.size, seed=0, capturable=not args.no_graph, l_dir=_sphere_samples())
    nviews = 4
    views = [synthetic_view(args.size, view_idx=7 * (rank * nviews + i), device=dev) for i in range(nviews)]
    host_views = [{k: v.cpu().pin_memory() for k, v in vw.items()} for vw in views]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host_views[0].values())
    params = [p for grp in pipe.optimizer.param_groups for p in grp['params']]

    def sync_grads():
        if world == 1:
            return
        gs = [p.grad for p in params if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in gs])
        dist.all_reduce(flat)
        flat.mul_(1.0 / world)
        o = 0
        for g in gs:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()

    if not args.no_fused and world > 1:
        # fused step: the weight gradients are summed over the ranks in GEMM order while the backward pass is still running,
        # the early optimiser group follows right behind them (relightable_nr_b200/fused.py); 1/world is applied inside Adam
        pipe.fused.allreduce_sum = lambda t: dist.all_reduce(t)
        pipe.fused.world = world

    def eager_step(view):
        if not args.no_fused:
            return pipe.fused.train_step(view)[0]
        final, rays_lt, alpha_map = pipe.forward(view)
        loss, _ = pipe.losses(view, final, rays_lt, alpha_map)
        loss.backward()
        sync_grads()
        pipe.optimizer.step()
        pipe.optimizer.zero_grad()
        return loss

    if args.no_graph:
        step = eager_step
    elif args.no_fused:
        step, _static = pipe.make_graphed_step(views[0], grad_hook=(lambda ps: sync_grads()) if world > 1 else None)
    else:
        step, _static = pipe.make_graphed_step(views[0], fused=True)

    if args.profile_steps:
        # ncu --profile-from-start off: only these eager steps are captured (tools/gpu_profile.sh)
        for i in range(3):
            eager_step(views[i % nviews])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        for i in range(args.profile_steps):
            eager_step(views[i % nviews])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return None

    # ---- device-resident arm -----------------------------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step(views[i % nviews])
    clocks = ClockSampler(D.local).start() if rank == 0 else None
    n0 = L.rnr_launch_count()
    ms = D.timed(lambda i: step(views[i % nviews]), args.steps)
    launches = L.rnr_launch_count() - n0
    if not args.no_graph:
        launches = pipe.graph_launches * args.steps      # kernels of librnr_b200.so recorded in the replayed graph
    clk = clocks.stop() if rank == 0 else None
    value = world * args.steps / (ms / 1e3)

    # ---- end-to-end arm: per-view maps from pinned host memory, loss read back -----------------------------------------
    e2e_step, reset, last_loss = _e2e_runner(D, host_views, step)
    reset()
    for i in range(3):
        e2e_step(i)
    reset()
    ms_e2e = D.timed(e2e_step, args.steps)
    e2e_value = world * args.steps / (ms_e2e / 1e3)

    # ---- sustained leg: >= args.sustain seconds of back-to-back steps with their own clock record ------------------------
    sustained = None
    if args.sustain > 0:
        n_s = max(int(args.sustain * value / max(world, 1)), args.steps)
        clocks2 = ClockSampler(D.local).start() if rank == 0 else None
        ms_s = D.timed(lambda i: step(views[i % nviews]), n_s)
        clk2 = clocks2.stop() if rank == 0 else None
        sustained = {'value': world * n_s / (ms_s / 1e3), 'unit': UNIT, 'steps': n_s, 'seconds': ms_s / 1e3, 'clocks': clk2}

    # ---- roofline leg: CUDA events around every conv launch (same stream), outside the timed regions -----------------------
    roof = None
    # (every rank runs the eager steps: at N > 1 they contain the gradient all-reduce, a collective)
    eng = [e for k, e in pipe.render_net.net._runner._engines.items() if e.need_backward][0]
    ROOF_ITERS = 10
    for i in range(3):                      # untimed: the eager path (Python launches) warm again after the graph replays
        eager_step(views[i % nviews])
    eng.timing = []
    for i in range(ROOF_ITERS):
        eager_step(views[i % nviews])
    torch.cuda.synchronize()
    rec, eng.timing = eng.timing, None
    if rank == 0:
        t = {'fwd': 0.0, 'dgrad': 0.0, 'wgrad': 0.0}
        nl = {'fwd': 0, 'dgrad': 0, 'wgrad': 0}
        per_layer = {}
        for kind, name, a, b in rec:
            dt = a.elapsed_time(b) / ROOF_ITERS
            t[kind] += dt
            per_layer.setdefault(name, {'fwd': 0.0, 'dgrad': 0.0, 'wgrad': 0.0})[kind] += dt * 1e3
        fl = {'fwd': 0.0, 'dgrad': 0.0, 'wgrad': 0.0}
        for sp in eng.specs:
            f = eng.layer_flops(sp, eng.N)
            st = eng.layers[sp.name]
            fl['fwd'] += f
            fl['wgrad'] += f
            nl['fwd'] += len(st.fwd_plans); nl['wgrad'] += 1; nl['dgrad'] += len(st.dgrad_plans)
            if st.dgrad_plans:
                r0, r1 = (eng.input_grad_range if sp.name == 'in' else (0, sum(sp.cin)))
                fl['dgrad'] += f * (r1 - r0) / sum(sp.cin)
        pk = _peaks()
        traffic, traffic_src = None, None
        for tp in ('r02_conv_halo_traffic.json', 'r01_conv_halo_traffic.json'):     # dram bytes per launch from the committed ncu capture
            tp = os.path.join(ROOT, 'profiles', tp)
            if os.path.exists(tp) and args.size == 512:
                traffic = json.load(open(tp)).get('traffic_bytes_per_launch')
                traffic_src = os.path.relpath(tp, ROOT)
                break
        conv_flops = fl['fwd'] + fl['dgrad']
        conv_ms = t['fwd'] + t['dgrad']
        ach = conv_flops / (conv_ms * 1e-3) / 1e12
        roof = {'kernel': 'conv_halo_kernel (tcgen05 implicit GEMM with shared-memory halo reuse: forward + data-gradient launches of the 22 U-Net layers)',
                'bound': 'tensor', 'achieved': ach, 'peak': pk['tf_burst'], 'unit': 'TFLOP/s', 'frac': ach / pk['tf_burst'], 'traffic': traffic,
                'traffic_source': (traffic_src + ': dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu)') if traffic else None,
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst: the timed region is sub-second) (%s)' % pk['src'],
                'frac_of_sustained_peak': ach / pk['tf_sust'],
                'flops_per_step': conv_flops, 'launches_per_step': nl['fwd'] + nl['dgrad'], 'avg_launch_us': conv_ms * 1e3 / (nl['fwd'] + nl['dgrad']),
                'kernel_ms_per_step': conv_ms, 'share_of_step': conv_ms / (ms / args.steps),
                'wgrad_tc_kernel': {'achieved': fl['wgrad'] / (t['wgrad'] * 1e-3) / 1e12, 'kernel_ms_per_step': t['wgrad'],
                                    'launches_per_step': nl['wgrad']},
                'whole_step_tflops': (fl['fwd'] + fl['dgrad'] + fl['wgrad']) * args.steps / (ms * 1e-3) / 1e12,
                # in-step duration of every layer's launches (us: forward, data gradient, weight gradient), same CUDA events
                'per_layer_us': {k: [round(v['fwd'], 1), round(v['dgrad'], 1), round(v['wgrad'], 1)] for k, v in per_layer.items()}}

    # ---- extras: the paths the unchanged scripts take (module-by-module, graph and eager), and the step with the GCN of :490 ----
    extras = {}
    if args.extras and world == 1:
        try:
            pipe2 = RNRPipeline(device=dev, img_size=args.size, seed=0, capturable=True, l_dir=_sphere_samples())

            def mod_step(view):
                final, rays_lt, alpha_map = pipe2.forward(view)
                loss, _ = pipe2.losses(view, final, rays_lt, alpha_map)
                loss.backward()
                pipe2.optimizer.step()
                pipe2.optimizer.zero_grad()
                return loss
            for i in range(3):
                mod_step(views[i % nviews])
            ms_m = D.timed(lambda i: mod_step(views[i % nviews]), 10)
            extras['module_path_eager_views_per_s'] = 10 / (ms_m / 1e3)
            gstep, _ = pipe2.make_graphed_step(views[0])
            for i in range(3):
                gstep(views[i % nviews])
            ms_g = D.timed(lambda i: gstep(views[i % nviews]), 10)
            extras['module_path_graph_views_per_s'] = 10 / (ms_g / 1e3)
            del gstep, pipe2
            torch.cuda.empty_cache()
        except Exception as e:       # extras must never cost the headline
            extras['module_path_error'] = repr(e)[:200]
        try:
            extras.update(_gcn_extras(args, D, pipe, views))
        except Exception as e:
            extras['gcn_error'] = repr(e)[:200]

    if rank != 0:
        D.teardown()
        return None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, info = cpu_reference_rate(args, 2, 1)
            cpu = {'value': v, 'unit': UNIT, 'cores': info['cores'], 'kind': info['kind'], 'sample': info['sample']}
        except Exception as e:
            cpu = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'reference', 'sample': 'failed: ' + repr(e)[:200]}
    return {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f16 operands / bf16 gradients, fp32 accumulate (tcgen05 kind::f16); fp32 per-pixel ops, fp32 master weights + Adam state',
        'data': 'synthetic', 'config': _config(args, world), 'clocks': clk,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu, 'sustained': sustained, 'extras': extras or None,
        'last_loss': last_loss[0],
    }


def _gcn_extras(args, D, pipe, views):
    """`v_feature = gcn(gcn_input)` of train_rnr.py:490 (DenseDeepGCN, 7500 vertices, 20 blocks): its result only feeds the dead
    branch of UnetSkipConnectionBlock, so the headline step does not run it; this measures what it costs when it is run
    every iteration, alone and next to the step (second stream)."""
    import types
    torch = D.torch
    from relightable_nr_b200.dropin import network
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import make_scene
    opt = types.SimpleNamespace(n_filters=64, kernel_size=16, act_type='relu', norm_type='batch', bias=True, epsilon=0.2, stochastic=True,
                                conv_type='edge', n_blocks=20, num_v_gcn=7500, out_channels_gcn=512, in_channels=6, block_type='res')
    gcn = network.DenseDeepGCN(opt).to(D.dev).train()
    v = torch.tensor(make_scene.grid_sphere(75, 100)[0], dtype=torch.float32, device=D.dev)
    inp = types.SimpleNamespace(pos=v, x=v)
    side = torch.cuda.Stream(device=D.dev)
    with torch.no_grad():
        for _ in range(2):
            gcn(inp)
        ms_gcn = D.timed(lambda i: gcn(inp), 5) / 5

        def both(i):
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                gcn(inp)
            pipe.fused.train_step(views[i % len(views)])
            torch.cuda.current_stream().wait_stream(side)
        both(0)
        ms_both = D.timed(both, 5) / 5
    return {'gcn_forward_ms': ms_gcn, 'step_with_gcn_every_iteration_views_per_s': 1e3 / ms_both,
            'gcn_note': 'DenseDeepGCN forward of train_rnr.py:490 (V=7500, 20 blocks); output feeds only the dead branch, not part of `value`'}


def _proxy_obj(D):
    """mesh.obj of the material-sphere proxy (128 x 256 UV sphere, 65 536 faces) in a temp directory."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import make_scene
    d = tempfile.mkdtemp(prefix='rnr_bench_')
    fp = os.path.join(d, 'mesh.obj')
    make_scene.write_obj(fp, *make_scene.uv_sphere(128, 256))
    return fp


def run_rnr_infer(args, D, relight=False):
    """test_rnr.py:265-393 per view, all on the device: Rasterizer -> get_TBN_map / get_view_dir_map / tangent-space view dir / SH basis ->
    fused head + U-Net forward + tail (x4 envmaps when ``relight``)."""
    torch = D.torch
    world, rank, dev = D.world, D.rank, D.dev
    from relightable_nr_b200 import _lib
    from relightable_nr_b200.dropin import camera, network, render, sph_harm
    from relightable_nr_b200.pipeline import RNRPipeline
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import make_scene
    L = _lib.lib()
    size = args.size
    pipe = RNRPipeline(device=dev, img_size=size, seed=0, dropout=False, l_dir=_sphere_samples())
    rast = network.Rasterizer(_proxy_obj(D), size).to(dev).eval()
    K = torch.tensor([[1.2 * size, 0, size / 2.0], [0, 1.2 * size, size / 2.0], [0, 0, 1.0]], dtype=torch.float32)
    n_cam = 720
    poses = [torch.tensor(make_scene.spiral_pose(i)[0], dtype=torch.float32) for i in range(rank, n_cam, world)]
    host = [{'pose': p[None].pin_memory(), 'proj': K[None].clone().pin_memory()} for p in poses]
    coeffs = None
    if relight:
        g = torch.Generator().manual_seed(5)
        coeffs = []
        for _ in range(4):
            c = torch.randn(121, 3, generator=g) * 0.1
            c[0] = 1.0
            coeffs.append(c.to(dev).contiguous())
    img_gt = torch.zeros((1, 3, size, size), device=dev)

    def view_maps(cam):
        pose, proj = cam['pose'].to(dev, non_blocking=True), cam['proj'].to(dev, non_blocking=True)
        r = rast(proj=proj, pose=pose, dist_coeffs=None, offset=None, scale=None)
        uv_map, alpha_map, fim, normal_map, faces_v, faces_vt = r[0], r[1], r[2], r[5], r[7], r[8]
        TBN = render.get_TBN_map(normal_map, fim, faces_v=faces_v[0], faces_texcoord=faces_vt[0], tangent=None)
        view_dir, _ = camera.get_view_dir_map((size, size), torch.inverse(proj), pose[:, :3, :3].transpose(1, 2).contiguous())
        vdt = torch.nn.functional.normalize(torch.matmul(TBN.reshape(-1, 3, 3).transpose(-2, -1), view_dir.reshape(-1, 3, 1))[..., 0]
                                            .reshape(view_dir.shape), dim=-1)
        return {'uv_map': uv_map, 'alpha_map': alpha_map, 'normal_map': normal_map, 'TBN_map': TBN, 'view_dir_map': view_dir,
                'view_dir_map_tangent': vdt, 'sh_basis_map': sph_harm.evaluate_sh_basis_l2(view_dir), 'img_gt': img_gt}

    out_host = torch.empty((4 if relight else 1, 3, size, size), dtype=torch.float32).pin_memory()

    @torch.no_grad()
    def step(i, readback=False):
        vw = view_maps(host[i % len(host)])
        outs = pipe.fused.render_relight(vw, coeffs) if relight else [pipe.fused.render(vw)]
        if readback:
            for k, o in enumerate(outs):
                out_host[k].copy_(o[0], non_blocking=True)
            torch.cuda.synchronize()
        return outs[0]

    for i in range(max(args.warmup, 3)):
        step(i)
    clocks = ClockSampler(D.local).start() if rank == 0 else None
    n0 = L.rnr_launch_count()
    ms = D.timed(lambda i: step(i), args.steps)
    launches = L.rnr_launch_count() - n0
    clk = clocks.stop() if rank == 0 else None
    ms_e2e = D.timed(lambda i: step(i, readback=True), args.steps)
    # the rasterizer alone, and "the kernel it must beat": the reference's own CUDA rasterizer recompiled for sm_100a
    raster = {}
    cam = {k: v.to(dev) for k, v in host[0].items()}
    with torch.no_grad():
        ms_r = D.timed(lambda i: rast(proj=cam['proj'], pose=cam['pose'], dist_coeffs=None, offset=None, scale=None), 20) / 20
    raster['ours_rasterizer_forward_ms'] = ms_r
    try:
        raster.update(_reference_raster_ms(D, rast, cam, size))
    except Exception as e:
        raster['reference_kernel'] = 'unavailable: ' + repr(e)[:160]
    if rank != 0:
        D.teardown()
        return None
    per_view_imgs = 4 if relight else 1
    return {
        'metric': 'views/sec at 512^2 (inference%s)' % (', 4 envmaps per view' if relight else ''), 'value': world * args.steps / (ms / 1e3), 'unit': UNIT,
        'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16 operands, fp32 accumulate (tcgen05 kind::f16); fp32 per-pixel ops',
        'data': 'synthetic', 'config': _config(args, world), 'clocks': clk,
        'e2e': {'value': world * args.steps / (ms_e2e / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': 100,
                'd2h_bytes_per_step': per_view_imgs * 3 * size * size * 4, 'ms_per_step': ms_e2e / args.steps},
        'images_per_s': per_view_imgs * world * args.steps / (ms / 1e3), 'gpu_launches': int(launches), 'rasterizer': raster,
        'roofline': None, 'cpu_baseline': None,
    }


def _reference_raster_ms(D, rast, cam, size):
    """forward_face_index_map of the reference's own extension (baseline/_ref/nr_ext, built by tools/build_ref_ext.py) on the same faces."""
    torch = D.torch
    so = os.path.join(ROOT, 'baseline', '_ref', 'nr_ext', 'ref_rasterize', 'ref_rasterize.so')
    if not os.path.exists(so):
        raise RuntimeError('baseline/_ref/nr_ext not built')
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_rasterize', so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from relightable_nr_b200.dropin import neural_renderer as nr
    with torch.no_grad():
        R, t = cam['pose'][:, :3, :3], cam['pose'][:, :3, 3][:, None, :]
        v = nr.projection(rast.vertices, cam['proj'], R, t, torch.zeros(1, 5, device=D.dev), size)
        faces = nr.vertices_to_faces(v, rast.faces).contiguous()
        nf = faces.shape[1]

        def run(i):
            fim = torch.full((1, size, size), -1, dtype=torch.int32, device=D.dev)
            wm = torch.zeros((1, size, size, 3), device=D.dev)
            dm = torch.full((1, size, size), 1e5, device=D.dev)
            inv_map = torch.zeros(1, device=D.dev)
            finv = torch.zeros((1, nf, 3, 3), device=D.dev)
            mod.forward_face_index_map(faces, fim, wm, dm, inv_map, finv, size, 0.0, 1e5, 0, 1, 1)
        run(0)
        ms = D.timed(run, 5) / 5
    return {'reference_forward_face_index_map_ms': ms,
            'reference_kernel': 'neural_renderer/cuda/rasterize_cuda_kernel.cu:24-169 recompiled -gencode arch=compute_100a,code=sm_100a (host glue patched, kernels untouched)'}


def run_dnr_train(args, D):
    torch, dist = D.torch, D.dist
    world, rank, dev = D.world, D.rank, D.dev
    from relightable_nr_b200 import _lib
    from relightable_nr_b200.pipeline import DNRPipeline, synthetic_view
    L = _lib.lib()
    pipe = DNRPipeline(device=dev, img_size=args.size, texture_size=512, texture_num_ch=16, mipmap_level=4, nf0=80)
    nviews = 4
    views = [synthetic_view(args.size, view_idx=7 * (rank * nviews + i), device=dev) for i in range(nviews)]
    host_views = [{k: v.cpu().pin_memory() for k, v in vw.items()} for vw in views]
    h2d = sum(v.numel() * v.element_size() for v in host_views[0].values())
    params = [p for grp in pipe.optimizer.param_groups for p in grp['params']]

    def step(view):
        out = pipe.forward(view)
        a = view['alpha_map'][:, None, 5:-5, 5:-5]
        loss = torch.nn.functional.l1_loss((out[:, :, 5:-5, 5:-5] * a).reshape(-1), (view['img_gt'][:, :, 5:-5, 5:-5] * a).reshape(-1))
        pipe.optimizer.zero_grad()
        loss.backward()
        if world > 1:
            gs = [p.grad for p in params if p.grad is not None]
            flat = torch.cat([g.reshape(-1) for g in gs])
            dist.all_reduce(flat)
            flat.mul_(1.0 / world)
            o = 0
            for g in gs:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()
        pipe.optimizer.step()
        return loss.detach()

    for i in range(max(args.warmup, 3)):
        step(views[i % nviews])
    clocks = ClockSampler(D.local).start() if rank == 0 else None
    n0 = L.rnr_launch_count()
    ms = D.timed(lambda i: step(views[i % nviews]), args.steps)
    launches = L.rnr_launch_count() - n0
    clk = clocks.stop() if rank == 0 else None
    e2e_step, reset, last = _e2e_runner(D, host_views, step)
    reset()
    for i in range(3):
        e2e_step(i)
    reset()
    ms_e2e = D.timed(e2e_step, args.steps)
    if rank != 0:
        D.teardown()
        return None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, info = cpu_reference_rate(args, 2, 1)
        cpu = {'value': v, 'unit': UNIT, 'cores': info['cores'], 'kind': info['kind'], 'sample': info['sample']}
    return {
        'metric': 'views/sec at 512^2 (fwd+bwd), DNR', 'value': world * args.steps / (ms / 1e3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f16 operands / bf16 gradients, fp32 accumulate (tcgen05 kind::f16); fp32 master weights', 'data': 'synthetic',
        'config': _config(args, world), 'clocks': clk,
        'e2e': {'value': world * args.steps / (ms_e2e / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches), 'roofline': None, 'cpu_baseline': cpu, 'last_loss': last[0],
    }


def run_ours(args):
    D = _Dist()
    if args.config == 'rnr_train':
        line = run_rnr_train(args, D)
    elif args.config in ('rnr_infer', 'rnr_relight'):
        line = run_rnr_infer(args, D, relight=args.config == 'rnr_relight')
    else:
        line = run_dnr_train(args, D)
    if line is not None:
        print(json.dumps(line))
    D.teardown()


def main():
    if os.environ.get('RNR_BENCH_WATCHDOG'):
        # debugging aid: dump every thread's Python stack and exit if the run is still going after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['RNR_BENCH_WATCHDOG']), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='rnr_train', choices=['rnr_train', 'rnr_infer', 'rnr_relight', 'dnr_train'])
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-budget', type=float, default=150.0, help='wall-clock bound (s) of the timed CPU steps of the reference arm')
    ap.add_argument('--sustain', type=float, default=5.0, help='seconds of back-to-back steps for the `sustained` key (0 = skip)')
    ap.add_argument('--extras', action='store_true', help='also time the module-by-module path (graph / eager) and the step with the GCN of '
                    'train_rnr.py:490 run every iteration (extra keys; N=1 only)')
    ap.add_argument('--profile-steps', type=int, default=0, help='run K eager steps between cudaProfilerStart/Stop and exit (for ncu)')
    ap.add_argument('--no-fused', action='store_true', help='drive the step operator by operator through the drop-in modules '
                    '(the reference script\'s call sequence) instead of the fused head/tail kernels')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying one CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
