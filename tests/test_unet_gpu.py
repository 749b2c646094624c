"""GPU parity of the U-Net engine (a15) against the CPU oracle (oracle/unet.py).

Tolerances (stated, none exist in the reference): forward PSNR >= 50 dB on the tanh output
(BASELINE.json north_star) and max-abs <= 2e-2 (fp16 operands, fp32 accumulation);
parameter gradients: relative L2 <= 3e-2 / cosine >= 0.999 against the oracle run with the engine's
fp16 storage emulated and the engine's own ReLU/LeakyReLU gate decisions (oracle.unet q16 + gates mode,
so only bf16 gradient rounding is left), and cosine >= 0.98 against the pure-fp32 oracle (there ~0.1% of the ReLU/LeakyReLU gates flip
because the forward pre-activations differ by ~1e-3, which alone is worth a few % relative L2)."""
import pytest
import torch

from tests.util import cosine, psnr, rel_l2

pytestmark = pytest.mark.gpu


def _setup(in_ch, out_ch, nf0, H, N, num_down, seed=0):
    from oracle.unet import make_unet_state_dict
    sd = make_unet_state_dict(in_ch, out_ch, nf0, num_down=num_down, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(N, in_ch, H, H, generator=g)
    return sd, x


def _engine(sd, x, out_ch, nf0, num_down, impl, grad_range, wgrad_impl=None):
    from relightable_nr_b200.engine.unet import UNetEngine, unet_layer_specs
    N, in_ch, H, W = x.shape
    dev = torch.device('cuda:0')
    params = {k: v.to(dev).contiguous() for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
    buffers = {k: v.to(dev).clone() for k, v in sd.items() if 'running' in k}
    specs = unet_layer_specs(in_ch, out_ch, nf0, num_down, 8 * nf0, H, W)
    eng = UNetEngine(specs, params, buffers, N, in_ch, dev, impl=impl, input_grad_range=grad_range, wgrad_impl=wgrad_impl)
    return eng, params


def _oracle_fwd_bwd(sd, x, num_down, R, grad_range, q16=False, gates=None):
    from oracle.unet import unet_forward
    sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k else v) for k, v in sd.items()}
    xg = x.clone().requires_grad_(True)
    out = torch.tanh(unet_forward(sdg, xg, num_down=num_down, q16=q16, gates=gates))
    (out * R).sum().backward()
    grads = {k: v.grad for k, v in sdg.items() if isinstance(v, torch.Tensor) and v.requires_grad and v.grad is not None}
    return out.detach(), grads, xg.grad[:, grad_range[0]:grad_range[1]]


@pytest.mark.parametrize("impl,wgrad_impl", [("simt", "simt"), ("tc", "simt"), ("tc", "tc")])
@pytest.mark.parametrize("cfg", [
    dict(in_ch=20, out_ch=6, nf0=16, H=64, N=1, num_down=5, grad_range=(4, 20)),
    dict(in_ch=108, out_ch=78, nf0=64, H=64, N=2, num_down=5, grad_range=(84, 108)),
])
def test_unet_forward_backward(cfg, impl, wgrad_impl):
    """("tc", "tc") is the product configuration: tcgen05 forward / data gradient AND tcgen05 weight gradient (wgrad_tc /
    wgrad_halo kernels, GEMM-order scratch + un-transpose) against the gate-matched oracle; the SIMT rows isolate the
    validation kernels."""
    sd, x = _setup(cfg['in_ch'], cfg['out_ch'], cfg['nf0'], cfg['H'], cfg['N'], cfg['num_down'])
    g = torch.Generator().manual_seed(7)
    R = torch.randn(cfg['N'], cfg['out_ch'], cfg['H'], cfg['H'], generator=g) / (cfg['H'] * cfg['H'])
    ref_out, ref_grads32, _ = _oracle_fwd_bwd(sd, x, cfg['num_down'], R, cfg['grad_range'])

    eng, params = _engine(sd, x, cfg['out_ch'], cfg['nf0'], cfg['num_down'], impl, cfg['grad_range'],
                          wgrad_impl=wgrad_impl)
    eng.set_input_nchw(x.cuda())
    eng.forward(training=True, drop_masks=None)
    out = eng.output_nchw().cpu()
    p = psnr(out * 0.5 + 0.5, ref_out * 0.5 + 0.5)
    err = (out - ref_out).abs().max().item()
    print(f"[{impl}] forward psnr {p:.1f} dB  max-abs {err:.2e}")
    assert p >= 50.0 and err <= 2e-2

    gates = {k: v.cpu() for k, v in eng.gate_masks().items()}
    _, ref_grads, ref_gx = _oracle_fwd_bwd(sd, x, cfg['num_down'], R, cfg['grad_range'], q16=True, gates=gates)
    gx = eng.backward_from_nchw(R.cuda())
    torch.cuda.synchronize()
    worst = 0.0
    for k, gref in ref_grads.items():
        gm = eng.grad_view(k).cpu()
        r, c = rel_l2(gm, gref), cosine(gm, gref)
        worst = max(worst, r)
        assert c >= 0.999 and r <= 3e-2, f"{k}: rel_l2 {r:.3e} cosine {c:.6f}"
        c32 = cosine(gm, ref_grads32[k])
        assert c32 >= 0.98, f"{k}: cosine vs fp32 oracle {c32:.5f}"
    r = rel_l2(gx.cpu(), ref_gx)
    print(f"[{impl}/{wgrad_impl}] worst param-grad rel_l2 {worst:.2e}; input-grad rel_l2 {r:.2e}")
    assert r <= 3e-2


@pytest.mark.parametrize("cfg", [
    dict(in_ch=108, out_ch=78, nf0=64, H=64, N=2, num_down=5, grad_range=(84, 108)),
    dict(in_ch=108, out_ch=78, nf0=64, H=128, N=1, num_down=5, grad_range=(84, 108)),
    dict(in_ch=20, out_ch=6, nf0=16, H=64, N=1, num_down=5, grad_range=(4, 20)),
])
def test_batchnorm_backward_sums_from_the_data_gradient_epilogue(cfg, monkeypatch):
    """BatchNorm backward with its per-channel sums accumulated by the consumers' data-gradient launches (conv_halo epilogue,
    rnr_conv_plan_set_gstats + rnr_bn_bwd_apply_src) against the separate reduction pass (RNR_BN_BWD_FUSED=0), same weights, same
    input, same Dropout2d masks: every parameter gradient within rel-L2 4e-2 / cosine 0.9995, the input gradient within 2e-2 (the
    sums now see the fp32 gradient before its bf16 rounding -- the BatchNorm weight gradients, sums of strongly cancelling terms,
    move by up to ~1 % --; nothing else differs).  Both variants are held to the oracle by test_unet_forward_backward."""
    sd, x = _setup(cfg['in_ch'], cfg['out_ch'], cfg['nf0'], cfg['H'], cfg['N'], cfg['num_down'])
    g = torch.Generator().manual_seed(7)
    R = (torch.randn(cfg['N'], cfg['out_ch'], cfg['H'], cfg['H'], generator=g) / (cfg['H'] * cfg['H'])).cuda()
    res = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('RNR_BN_BWD_FUSED', mode)
        eng, params = _engine(sd, x, cfg['out_ch'], cfg['nf0'], cfg['num_down'], 'tc', cfg['grad_range'], wgrad_impl='tc')
        gm = torch.Generator(device='cuda').manual_seed(3)
        masks = {sp.name: (torch.rand(cfg['N'], sp.cout, device='cuda', generator=gm) > 0.5).float() * 2.0
                 for sp in eng.specs if sp.drop and sp.dst != 'out'}
        assert mode == '1' or not eng.gstat_layers
        assert mode == '0' or cfg['nf0'] < 64 or len(eng.gstat_layers) >= 8, sorted(eng.gstat_layers)
        if mode == '1':
            print('BatchNorm layers with fused backward sums: %d of %d: %s' % (
                len(eng.gstat_layers), sum(1 for sp in eng.specs if sp.bn_key), sorted(eng.gstat_layers)))
        eng.set_input_nchw(x.cuda())
        for rep in range(2):                  # twice: the second run proves that the accumulators were re-armed
            eng.forward(training=True, drop_masks=masks)
            gi = eng.backward_from_nchw(R)
        torch.cuda.synchronize()
        res[mode] = ({k: eng.grad_view(k).clone() for k in eng.grad_slices}, gi.clone())
    errs = []
    for k, a in res['0'][0].items():
        b = res['1'][0][k]
        if a.abs().max().item() == 0.0:
            assert b.abs().max().item() == 0.0, k
            continue
        errs.append((rel_l2(b.cpu(), a.cpu()), cosine(b.cpu(), a.cpu()), k))
    errs.sort(reverse=True)
    for e, c, k in errs[:6]:
        print('  rel-L2 %.2e cosine %.6f  %s' % (e, c, k))
    e_in = rel_l2(res['1'][1].cpu(), res['0'][1].cpu())
    print('input gradient rel-L2 %.2e' % e_in)
    for e, c, k in errs:
        assert e <= 4e-2 and c >= 0.9995, (k, e, c)
    assert e_in <= 2e-2


def test_conv_kernel_variants_agree(monkeypatch):
    """The halo conv kernel's variants -- CTA pair (tcgen05.mma.cta_group::2, M = 256) vs single CTA, register-direct vs staged
    epilogue -- run the same MMAs over the same K order: the first layer's raw output (no BatchNorm statistics upstream) must be
    BIT-IDENTICAL across them, the network output and every parameter gradient agree to rounding of the statistics' summation
    order through 22 BatchNorm layers (PSNR >= 70 dB, gradient cosine >= 0.99; pair vs single CTA: rel-L2 <= 1e-4)."""
    cfg = dict(in_ch=108, out_ch=78, nf0=64, H=128, N=1, num_down=5, grad_range=(84, 108))
    sd, x = _setup(cfg['in_ch'], cfg['out_ch'], cfg['nf0'], cfg['H'], cfg['N'], cfg['num_down'])
    g = torch.Generator().manual_seed(7)
    R = (torch.randn(cfg['N'], cfg['out_ch'], cfg['H'], cfg['H'], generator=g) / (cfg['H'] * cfg['H'])).cuda()
    res = {}
    for name, env in (('default', {}), ('single_cta', {'RNR_CONV_PAIR': '0'}), ('staged', {'RNR_CONV_DIRECT': '0'}),
                      ('pair_everywhere_direct_everywhere', {'RNR_CONV_PAIR_MINBN': '16', 'RNR_CONV_DIRECT': '2'})):
        for k in ('RNR_CONV_PAIR', 'RNR_CONV_DIRECT', 'RNR_CONV_PAIR_MINBN'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng, _ = _engine(sd, x, cfg['out_ch'], cfg['nf0'], cfg['num_down'], 'tc', cfg['grad_range'], wgrad_impl='tc')
        eng.set_input_nchw(x.cuda())
        eng.forward(training=True, drop_masks=None)
        out = eng.output_nchw().clone()
        gi = eng.backward_from_nchw(R)
        torch.cuda.synchronize()
        res[name] = (eng.layers['in'].raw.clone(), out, {k: eng.grad_view(k).clone() for k in eng.grad_slices}, gi.clone())
    ref = res['single_cta']
    for name, (raw0, out, grads, gi) in res.items():
        assert torch.equal(raw0.view(torch.int16), ref[0].view(torch.int16)), name
        p = psnr(out.cpu() * 0.5 + 0.5, ref[1].cpu() * 0.5 + 0.5)
        errs = sorted(((rel_l2(grads[k].cpu(), ref[2][k].cpu()), cosine(grads[k].cpu(), ref[2][k].cpu()), k)
                       for k in grads if ref[2][k].abs().max().item() > 0), reverse=True)
        worst, wcos, wkey = errs[0]
        print('%-36s output PSNR vs the single-CTA run %.1f dB; worst gradient rel-L2 %.2e (cosine %.5f, %s)' % (name, p, worst, wcos, wkey))
        assert p >= 70.0, name
        if name in ('default', 'single_cta'):
            assert worst <= 1e-4, name          # same epilogue, same per-tile arithmetic: only the CTA -> tile assignment differs
        else:
            # other summation order of the fp32 batch sums (E[x^2] - mean^2 cancels): the perturbation passes through 22 BatchNorm
            # layers and a noise-like upstream gradient; direction and size of every gradient must still agree
            assert min(c for _, c, _ in errs) >= 0.99 and worst <= 0.2, (name, errs[:3])


def test_eval_mode_batchnorm_uses_running_statistics():
    """module.eval() WITHOUT the scripts' set_bn_train (test_rnr.py:220-233 re-enables train mode): nn.BatchNorm2d then
    normalises with running_mean / running_var.  Engine vs oracle (F.batch_norm(training=False)) on non-trivial running
    statistics: PSNR >= 50 dB; and the statistics must be left untouched."""
    from oracle.unet import unet_forward
    sd, x = _setup(20, 6, 16, 64, 1, 5)
    g = torch.Generator().manual_seed(11)
    for k in list(sd):
        if k.endswith('running_mean'):
            sd[k] = 0.3 * torch.randn(sd[k].shape, generator=g)
        if k.endswith('running_var'):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g)
    eng, _ = _engine(sd, x, 6, 16, 5, 'tc', None)
    before = {k: v.clone() for k, v in eng.buffers.items()}
    eng.set_input_nchw(x.cuda())
    eng.forward(training=False, drop_masks=None)
    out = eng.output_nchw().cpu()
    ref = torch.tanh(unet_forward(sd, x, num_down=5, bn_eval=True))
    ref_train = torch.tanh(unet_forward(sd, x, num_down=5))
    p = psnr(out * 0.5 + 0.5, ref * 0.5 + 0.5)
    print('eval-mode BN: psnr %.1f dB (vs the batch-statistic output: %.1f dB)' % (p, psnr(out * 0.5 + 0.5, ref_train * 0.5 + 0.5)))
    assert p >= 50.0
    assert psnr(ref * 0.5 + 0.5, ref_train * 0.5 + 0.5) < 40.0, 'the two BatchNorm modes must differ for this test to mean anything'
    for k, v in eng.buffers.items():
        assert torch.equal(v, before[k]), k


def test_batched_weight_prep_is_bit_identical_to_per_layer():
    """rnr_wprep_run (one launch, smem-staged) against rnr_weight_prep (one launch per matrix): identical 16-bit matrices for
    every forward / data-gradient matrix of the RNR U-Net (incl. the input-gradient channel range and ConvTranspose parities)."""
    sd, x = _setup(108, 78, 64, 64, 1, 5)
    eng, _ = _engine(sd, x, 78, 64, 5, 'tc', (84, 108))
    items = [w for sp in eng.specs for w in eng.layers[sp.name].wprep_fwd + eng.layers[sp.name].wprep_dgrad]
    assert len(items) >= 70
    eng.prepare_weights_per_layer(backward=True)
    torch.cuda.synchronize()
    ref = [w.dst.clone() for w in items]
    for w in items:
        w.dst.fill_(float('nan'))
    eng.prepare_weights(backward=True)
    torch.cuda.synchronize()
    for w, r in zip(items, ref):
        assert torch.equal(w.dst.view(torch.int16), r.view(torch.int16)), (w.src_key, w.ntaps, w.s_r, w.s_c)


@pytest.mark.parametrize('use_gcn', [True, False])
def test_dropout_reference_rng_order_matches_the_real_module(use_gcn):
    """Dropout ON (the training configuration, train_rnr.py:398-405): with ``rng_order = 'reference'`` the drop-in draws one
    bernoulli per Dropout2d call of the reference's forward in the reference's order -- including the 22 draws of the dead GCN pass
     -- so that under the same torch.manual_seed its output equals the REAL reference module's (staged copy,
    run on the same GPU, TF32 off): PSNR >= 50 dB.  A single wrong draw would put the two ~25 dB apart (checked: the 'fast' order is)."""
    from tests.golden import ref_import
    if not ref_import.available():
        pytest.skip('reference not staged')
    ref = ref_import.import_reference()
    from relightable_nr_b200.dropin import network
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    kw = dict(nf0=16, in_channels=20, out_channels=6, num_down_unet=5, use_gcn=use_gcn)
    ref_net = ref.network.RenderingNet(**kw).cuda().train()
    ours = network.RenderingNet(**kw).cuda().train()
    ours.load_state_dict(ref_net.state_dict(), strict=True)
    x = torch.randn(2, 20, 64, 64, device='cuda')
    v_fea = torch.zeros(2, 512, device='cuda')
    with torch.no_grad():
        torch.manual_seed(1234)
        want = ref_net(x, v_fea if use_gcn else None)
        ours.net._runner.rng_order = 'reference'
        torch.manual_seed(1234)
        got = ours(x, None)
        ours.net._runner.rng_order = 'fast'
        torch.manual_seed(1234)
        other = ours(x, None)
    p, p_fast = psnr(got * 0.5 + 0.5, want * 0.5 + 0.5), psnr(other * 0.5 + 0.5, want * 0.5 + 0.5)
    print('dropout on, use_gcn=%s: PSNR vs the real module %.1f dB with the reference draw order, %.1f dB with the fast order' % (use_gcn, p, p_fast))
    assert p >= 50.0
    assert p_fast < 45.0, 'the two draw orders must differ for this test to mean anything'
