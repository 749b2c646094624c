#!/usr/bin/env python
"""Benchmark of the RNR per-view training step (BASELINE.json: views/sec at 512^2, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 512]

Workload (config.workload): configs[2] of BASELINE.json restricted to one box -- the body of train_rnr.py:490-623
(texture sample -> 26 rays/pixel -> 108->78-channel U-Net -> SH envmap ray render -> 4 losses -> backward -> Adam) on
synthetic 512x512 views of the material-sphere proxy, 1 view per GPU per step (the reference's own constraint),
views sharded over ranks (weak scaling), one NCCL all-reduce of the gradients per step.

Prints ONE JSON line (rank 0).  ``value`` = views/s with the per-view maps resident in HBM; ``e2e`` = the same step fed
from pinned host buffers (H2D of the 8 per-view maps + D2H of the loss inside the timed region); ``roofline`` = live
conv FLOPs of the tcgen05 implicit-GEMM kernel (conv_halo_kernel) / its CUDA-event time / measured bf16 peak; ``cpu_baseline`` = the oracle
port of the same step on the host cores.  ``--impl reference`` times that CPU path alone (the reference is PyTorch-CPU
Python: it cannot travel to the GPU box, so the arm is the golden-pinned oracle port of it -- kind "port").
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
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
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


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference step
# ----------------------------------------------------------------------------------------------------------------------
def _cpu_state(size, nf0, seed=0):
    """Random-init oracle state of the RNR module set at the benchmark config (no GPU involved)."""
    import math
    import torch
    from oracle import pixel_ops as P
    from oracle.unet import make_unet_state_dict
    g = torch.Generator().manual_seed(seed)
    C, R = 24, 26
    base = [torch.full((1, s, s, C), 1.0 if i == 0 else 0.01) for i, s in enumerate((512, 256, 128, 64))]
    tex_init = torch.relu(P.flatten_mipmap(base, 0, 6))
    tex = [b + 0.05 * torch.randn(b.shape, generator=g) for b in base]
    sd = {'net.' + k: v for k, v in make_unet_state_dict(R * 3 + 6 + C, 3 * R, nf0, num_down=5, seed=seed).items()}
    n = 4096
    i = torch.arange(n, dtype=torch.float64) + 0.5
    z = 1 - 2 * i / n
    r = torch.sqrt(1 - z * z)
    phi = math.pi * (1 + 5 ** 0.5) * i
    l_dir = torch.stack((r * torch.cos(phi), r * torch.sin(phi), z), 1).float()
    basis_val = torch.from_numpy(P.evaluate_sh_basis(10, l_dir.numpy())).float()
    lh, lw = 256, 512
    vv, uu = torch.meshgrid(torch.arange(lh, dtype=torch.float32) / (lh - 1), torch.arange(lw, dtype=torch.float32) / (lw - 1), indexing='ij')
    grid_dir = P.spherical_mapping_inv(torch.stack((uu, vv)).flatten(1)).t().contiguous()
    basis_recon = torch.from_numpy(P.evaluate_sh_basis(10, grid_dir.numpy())).float()
    coeff = torch.randn((121, 3), generator=g) * 0.1
    coeff[0] = 1.0
    mask = torch.ones(n, dtype=torch.bool)
    mask[::7] = False
    return dict(textures=tex, tex_init=tex_init, unet_sd=sd, coeff=coeff, basis_val=basis_val,
                basis_val_recon=basis_recon, lp_hw=(lh, lw), pivots_s=P.ray_sampler_constants(6, 2, 5)[1],
                pivots_d=P.ray_sampler_constants(6, 2, 10)[1], l_init=P.reconstruct_sh(coeff, basis_val), l_mask=mask,
                w=dict(lighting=1.0, lighting_uncovered=0.1, rays_lt_chrom=1.0, alb=1.0))


def _cpu_view(size, seed=0):
    """Synthetic per-view maps on the CPU (random but well-formed: unit normals/TBN, alpha disc)."""
    import torch
    g = torch.Generator().manual_seed(seed + 100)
    H = W = size
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing='ij')
    alpha = (((xx - W / 2) ** 2 + (yy - H / 2) ** 2) < (0.4 * W) ** 2).float()[None]
    TBN = torch.linalg.qr(torch.randn(1, H, W, 3, 3, generator=g))[0] * alpha[..., None, None]
    nrm = torch.nn.functional.normalize(torch.randn(1, H, W, 3, generator=g), dim=-1)
    return dict(uv_map=torch.rand(1, H, W, 2, generator=g) * alpha[..., None], sh_basis_map=torch.randn(1, H, W, 9, generator=g),
                normal_map=nrm * alpha[..., None], view_dir_map=nrm.flip(-1), view_dir_map_tangent=nrm.roll(1, -1), TBN_map=TBN,
                alpha_map=alpha, img_gt=torch.rand(1, 3, H, W, generator=g) * alpha[:, None])


def cpu_step_rate(size, steps, warmup, budget_s=150.0, nf0=64):
    """views/s of the oracle port (fwd + losses + bwd) on all host cores.  Each step is one view; if the projected run
    exceeds ``budget_s`` the view is cropped to a centred (size/2)^2 window and the rate scaled by the pixel fraction."""
    import torch
    from oracle.rnr_step import rnr_step
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state = _cpu_state(size, nf0)
    cur = size
    view = _cpu_view(cur)
    t0 = time.time()
    rnr_step(state, view)
    t_first = time.time() - t0
    sample = '1 view %dx%d per step, fwd+4 losses+bwd' % (cur, cur)
    done_warm = 1
    if t_first * (steps + max(warmup - 1, 0)) > budget_s and size >= 256:
        cur = size // 2
        view = _cpu_view(cur)
        sample = 'centre %dx%d window of a %dx%d view per step (rate scaled by pixel fraction 1/4), fwd+4 losses+bwd' % (cur, cur, size, size)
        done_warm = 0
    for _ in range(max(warmup - done_warm, 0)):
        rnr_step(state, view)
    n = 0
    t0 = time.time()
    while n < steps:
        rnr_step(state, view)
        n += 1
        if time.time() - t0 > budget_s and n >= 1:
            break
    dt = (time.time() - t0) / n
    frac = (cur * cur) / float(size * size)
    return frac / dt, dict(cores=cores, kind='port', sample=sample + '; %d timed steps' % n, ms_per_step=dt * 1e3 / frac)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    v, info = cpu_step_rate(args.size, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': info['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': _config(args, 0),
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': info['cores'], 'kind': info['kind'], 'sample': info['sample']},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0,
    }
    print(json.dumps(line))


def _config(args, world):
    return {'workload': 'RNR train step (train_rnr.py:490-623), %dx%d material-sphere proxy views, 1 view/GPU/step, texture 512^2x24ch x4 mips, '
                        'U-Net 108->78 nf0=64, 26 rays, SH lmax 10 envmap 256x512' % (args.size, args.size),
            'views_per_step': max(world, 1), 'parallelism': 'dp%d (views sharded, NCCL grad all-reduce)' % max(world, 1), 'launch': 'eager' if args.no_graph else 'one CUDA graph per step',
            'step': 'module-by-module (drop-in operator API)' if args.no_fused else 'fused head/tail kernels around the U-Net (relightable_nr_b200/fused.py)',
            'l2': 'per-step working set (~3 GB of activations/gradients) >> 126 MB L2; 4 distinct views cycled'}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: librnr_b200 has no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from relightable_nr_b200 import _lib
    from relightable_nr_b200.pipeline import RNRPipeline, synthetic_view
    L = _lib.lib()
    pipe = RNRPipeline(device=dev, img_size=args.size, seed=0, capturable=not args.no_graph)
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

    def sync_grad_buffers(bufs):
        # fused step: the U-Net gradients already live in ONE flat buffer -> one large all-reduce + 5 small ones, no packing copies
        if world == 1:
            return
        for g in bufs:
            dist.all_reduce(g)
            g.mul_(1.0 / world)

    def allreduce_mean(t):
        dist.all_reduce(t)
        t.mul_(1.0 / world)

    if not args.no_fused and world > 1:
        # fused step: the weight gradients are all-reduced in GEMM order while the backward pass is still running
        # (relightable_nr_b200/fused.py); RNR_AR_OVERLAP=0 falls back to one bucket list after the backward pass
        if os.environ.get('RNR_AR_OVERLAP', '1') != '0':
            pipe.fused.allreduce = allreduce_mean
        else:
            pipe.fused.grad_hook = sync_grad_buffers

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
    else:
        # the whole iteration (incl. the gradient all-reduce) as one CUDA graph; per-view maps are copied into its static inputs
        if args.no_fused:
            step, _static = pipe.make_graphed_step(views[0], grad_hook=(lambda ps: sync_grads()) if world > 1 else None)
        else:
            step, _static = pipe.make_graphed_step(views[0], fused=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

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
        return

    # ---- device-resident arm -----------------------------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step(views[i % nviews])
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = L.rnr_launch_count()
    ms = timed(lambda i: step(views[i % nviews]), args.steps)
    launches = L.rnr_launch_count() - n0
    if not args.no_graph:
        launches = pipe.graph_launches * args.steps      # kernels of librnr_b200.so recorded in the replayed graph
    clk = clocks.stop() if rank == 0 else None
    value = world * args.steps / (ms / 1e3)

    # ---- end-to-end arm: per-view maps from pinned host memory, loss read back -----------------------------------------
    dev_bufs = [{k: torch.empty_like(v, device=dev) for k, v in host_views[0].items()} for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    last_loss = [None]

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
        loss = step(dev_bufs[b])
        done[b].record()
        last_loss[0] = loss.item()          # D2H of the step's result, every step

    for b in range(2):
        done[b].record()
    for i in range(3):
        e2e_step(i)
    torch.cuda.synchronize()
    for b in range(2):
        done[b].record()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = world * args.steps / (ms_e2e / 1e3)

    # ---- roofline leg: CUDA events around every conv launch (same stream), outside the timed regions -----------------------
    roof = None
    # (every rank runs the eager steps: at N > 1 they contain the gradient all-reduce, a collective)
    eng = [e for k, e in pipe.render_net.net._runner._engines.items() if e.need_backward][0]
    eng.timing = []
    for i in range(3):
        eager_step(views[i % nviews])
    torch.cuda.synchronize()
    rec, eng.timing = eng.timing, None
    if rank == 0:
        t = {'fwd': 0.0, 'dgrad': 0.0, 'wgrad': 0.0}
        nl = {'fwd': 0, 'dgrad': 0, 'wgrad': 0}
        for kind, name, a, b in rec:
            t[kind] += a.elapsed_time(b) / 3
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
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'r01_conv_halo_traffic.json')     # dram bytes per launch from the committed ncu capture
        if os.path.exists(tp) and args.size == 512:
            traffic = json.load(open(tp)).get('traffic_bytes_per_launch')
        conv_flops = fl['fwd'] + fl['dgrad']
        conv_ms = t['fwd'] + t['dgrad']
        ach = conv_flops / (conv_ms * 1e-3) / 1e12
        roof = {'kernel': 'conv_halo_kernel (tcgen05 implicit GEMM with shared-memory halo reuse: forward + data-gradient launches of the 22 U-Net layers)', 'bound': 'tensor',
                'achieved': ach, 'peak': pk['tf_sust'], 'unit': 'TFLOP/s', 'frac': ach / pk['tf_sust'], 'traffic': traffic,
                'traffic_source': 'profiles/r01_conv_halo_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu, 44 launches of one step)',
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (%s)' % pk['src'],
                'flops_per_step': conv_flops, 'launches_per_step': nl['fwd'] + nl['dgrad'], 'avg_launch_us': conv_ms * 1e3 / (nl['fwd'] + nl['dgrad']),
                'kernel_ms_per_step': conv_ms, 'share_of_step': conv_ms / (ms / args.steps),
                'wgrad_tc_kernel': {'achieved': fl['wgrad'] / (t['wgrad'] * 1e-3) / 1e12, 'kernel_ms_per_step': t['wgrad'],
                                    'launches_per_step': nl['wgrad']},
                'whole_step_tflops': (fl['fwd'] + fl['dgrad'] + fl['wgrad']) * args.steps / (ms * 1e-3) / 1e12}

    def teardown():
        # destroy_process_group() blocks forever while a captured CUDA graph still references the communicator's work
        # (measured: both ranks parked in it after a complete run), so the ranks leave through os._exit once rank 0 has
        # printed its line; the process exit tears NCCL down.
        if world > 1:
            sys.stdout.flush()
            sys.stderr.flush()
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        teardown()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, info = cpu_step_rate(args.size, 1, 1, budget_s=40.0)
        cpu = {'value': v, 'unit': UNIT, 'cores': info['cores'], 'kind': info['kind'], 'sample': info['sample']}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f16 operands / bf16 gradients, fp32 accumulate (tcgen05 kind::f16); fp32 per-pixel ops',
        'data': 'synthetic', 'config': _config(args, world), 'clocks': clk,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu, 'last_loss': last_loss[0],
    }
    print(json.dumps(line))
    teardown()


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
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
