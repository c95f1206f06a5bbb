"""Whole-network autograd nodes for the DCGAN generator / discriminator over the library's kernels.

``generator_forward(module, z)`` and ``discriminator_forward(module, x)`` are what the drop-in
``networks.ConvGenerator`` / ``networks.SNDiscriminator`` call on CUDA.  The module tree only holds
the parameters and buffers (reference ``state_dict`` format); the arithmetic runs here:

  generator      Linear+ReLU -> 3 x [ConvT k4s2 -> BatchNorm(batch stats) -> ReLU] -> ConvT k3s1 -> Tanh
                 (networks/conv_generator.py:13-27)
  discriminator  7 x [spectral-norm Conv + LeakyReLU(0.1)] -> spectral-norm Linear(8192 -> 1)
                 (networks/sn_discriminator.py:8-25)

Activations are NHWC bf16 between layers, fp32 NCHW at the module boundary; master weights fp32.
Every dense contraction (forward, data gradient, weight gradient) is a tcgen05 tap GEMM
(``dense.Plan`` / ``dense.WGradPlan``); normalisation, patch gathering and the final GEMV are the
kernels of csrc/nn_misc.cu.  Nothing here synchronises or allocates outside PyTorch's caching
allocator, so a whole training step can be captured in one CUDA graph.
"""
import ctypes

import torch
import torch.nn as nn

from . import dense
from ._lib import check, lib

BN_EPS_DEFAULT = 1e-5


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


# ----------------------------------------------------------------------------------------- kernel wrappers
def im2col3(x, tanh_out=None):
    B, C, H, W = x.shape
    assert C == 3 and x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty(B, H, W, 64, device=x.device, dtype=torch.bfloat16)
    check(lib().ipr_im2col3_bf16(_p(x), _p(tanh_out), _p(out), B, H, W, _st()), "ipr_im2col3_bf16")
    return out


def bn_finalize(stats, count, bn, update_running):
    C = bn.num_features
    dev = stats.device
    scale = torch.empty(C, device=dev)
    shift = torch.empty(C, device=dev)
    mean = torch.empty(C, device=dev)
    rstd = torch.empty(C, device=dev)
    momentum = bn.momentum if bn.momentum is not None else 0.1
    rm = bn.running_mean if update_running else None
    rv = bn.running_var if update_running else None
    nbt = bn.num_batches_tracked if update_running else None
    check(lib().ipr_bn_finalize_f32(_p(stats), stats.shape[0], C, float(count), float(bn.eps), float(momentum),
                                    _p(bn.weight.detach()), _p(bn.bias.detach()), _p(rm), _p(rv), _p(nbt),
                                    _p(scale), _p(shift), _p(mean), _p(rstd), _st()), "ipr_bn_finalize_f32")
    return scale, shift, mean, rstd


def bn_apply_relu(x, scale, shift):
    y = torch.empty_like(x)
    C = x.shape[-1]
    check(lib().ipr_bn_apply_relu_bf16(_p(x), _p(y), _p(scale), _p(shift), x.numel() // C, C, _st()),
          "ipr_bn_apply_relu_bf16")
    return y


def bn_relu_bwd(dy, xraw, act, gamma, mean, rstd, dgamma, dbeta, accumulate, sign=None, gamma0=0.0, sign_scale=0.0):
    C = dy.shape[-1]
    nbytes = lib().ipr_bn_bwd_workspace_bytes(C)
    ws = torch.empty(nbytes // 4, device=dy.device, dtype=torch.float32)
    dx = torch.empty_like(dy)
    check(lib().ipr_bn_relu_bwd_bf16(_p(dy), _p(xraw), _p(act), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dgamma),
                                     _p(dbeta), int(bool(accumulate)), _p(sign), float(gamma0), float(sign_scale),
                                     _p(ws), nbytes, dy.numel() // C, C, _st()), "ipr_bn_relu_bwd_bf16")
    return dx


def dfc_fwd(a, w, sigma, bias):
    B, K = a.shape
    logits = torch.empty(B, device=a.device, dtype=torch.float32)
    check(lib().ipr_dfc_fwd_bf16(_p(a), _p(w), _p(sigma), _p(bias), _p(logits), B, K, _st()), "ipr_dfc_fwd_bf16")
    return logits


def dfc_bwd(a, w, sigma, dlogit, want_dw, slope):
    B, K = a.shape
    da = torch.empty_like(a)
    dw = torch.empty(K, device=a.device, dtype=torch.float32) if want_dw else None
    check(lib().ipr_dfc_bwd_bf16(_p(a), _p(w), _p(sigma), _p(dlogit), _p(da), _p(dw), 0, float(slope), B, K, _st()),
          "ipr_dfc_bwd_bf16")
    return da, dw


# ----------------------------------------------------------------------------------------- plan caches
class _Packs(object):
    """bf16 packed weight matrices, refreshed when the fp32 master changes (``_version``)."""

    def __init__(self):
        self.store = {}

    def get(self, key, param, fn):
        ver = (param._version, param.data_ptr())
        hit = self.store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        with torch.no_grad():
            val = fn(param.detach())
        self.store[key] = (ver, val)
        return val

    def clear(self):
        self.store.clear()


_ALL_PACKS = []


def reset_caches():
    """Drop every cached weight pack (call before CUDA-graph capture so packing is part of the graph)."""
    for p in _ALL_PACKS:
        p.clear()


def _col_off_patch27(n_is_first):
    """Column table for the 3-channel im2col layers: k = (kh*3+kw)*3 + c  ->  offset inside a (., 3, 3, 3) weight row."""
    off = torch.full((1, 64), -1, dtype=torch.int32)
    for kh in range(3):
        for kw in range(3):
            for c in range(3):
                off[0, (kh * 3 + kw) * 3 + c] = c * 9 + kh * 3 + kw
    return off


class GenPlans(object):
    def __init__(self, module):
        mg = module.mg
        self.mg = mg
        feat = 512 * mg * mg
        # NHWC position n' = hw*512 + c  <->  reference feature f = c*mg*mg + hw
        hw = torch.arange(mg * mg).view(-1, 1)
        c = torch.arange(512).view(1, -1)
        self.perm = (c * (mg * mg) + hw).reshape(-1)                    # perm[n'] = f
        self.fc = dense.Plan("linear", module.z_dim, feat)
        self.fc_wg = dense.WGradPlan(self.fc, (feat, module.z_dim), row_perm=self.perm)
        chans = (512, 256, 128, 64)
        self.ct = [dense.Plan("convT4s2", ci, co) for ci, co in zip(chans[:-1], chans[1:])]
        self.ct_dg = [dense.Plan("convT4s2_dgrad", co, ci) for ci, co in zip(chans[:-1], chans[1:])]
        self.ct_wg = [dense.WGradPlan(p, (p.cin, p.cout, 4, 4)) for p in self.ct]
        self.last = dense.Plan("convT3", 64, 3, n_pad=16)
        self.last_dg = dense.Plan("linear", 64, 64)                    # d(a3) = patches(dY) x W'
        self.last_wg = dense.WGradPlan(dense.Plan("linear", 64, 64), (64, 3, 3, 3), col_off=_col_off_patch27(True),
                                       s_n=27)
        self.last_dg.k_valid_override = 27            # flop accounting: 27 of the 64 patch columns carry data
        self.last_wg.k_valid_override = 27
        self.packs = _Packs()
        self._perm_dev = {}
        _ALL_PACKS.append(self.packs)

    def perm_on(self, device):
        key = str(device)
        if key not in self._perm_dev:
            self._perm_dev[key] = self.perm.to(device)
        return self._perm_dev[key]

    def pack_last_dgrad(self, w):
        # B[n = ci][k = (kh*3+kw)*3 + co] = W[ci, co, kh, kw]
        m = torch.zeros(1, 64, 64, device=w.device, dtype=torch.bfloat16)
        m[0, :, :27] = w.permute(0, 2, 3, 1).reshape(64, 27).to(torch.bfloat16)
        return m


class DisPlans(object):
    def __init__(self, module):
        md = module.md
        self.md = md
        specs = [("conv4s2", 64, 64), ("conv3", 64, 128), ("conv4s2", 128, 128), ("conv3", 128, 256),
                 ("conv4s2", 256, 256), ("conv3", 256, 512)]
        self.first = dense.Plan("linear", 64, 64)                      # patches(x) x W1
        self.first_wg = dense.WGradPlan(dense.Plan("linear", 64, 64), (64, 3, 3, 3), col_off=_col_off_patch27(False),
                                        s_n=27)
        self.first_dg = dense.Plan("conv3_dgrad", 64, 3, n_pad=16)
        self.conv = [dense.Plan(k, ci, co) for k, ci, co in specs]
        self.conv_dg = [dense.Plan(k + "_dgrad", co, ci) for k, ci, co in specs]
        self.conv_wg = [dense.WGradPlan(p, (p.cout, p.cin, 3 if p.kind == "conv3" else 4, 3 if p.kind == "conv3" else 4))
                        for p in self.conv]
        hw = torch.arange(md * md).view(-1, 1)
        c = torch.arange(512).view(1, -1)
        self.perm = (c * (md * md) + hw).reshape(-1)                    # NHWC feature n' -> reference feature
        self.first.k_valid_override = 27
        self.first_wg.k_valid_override = 27
        self.packs = _Packs()
        self._perm_dev = {}
        _ALL_PACKS.append(self.packs)

    def perm_on(self, device):
        key = str(device)
        if key not in self._perm_dev:
            self._perm_dev[key] = self.perm.to(device)
        return self._perm_dev[key]

    def pack_first(self, w):
        # B[n = co][k = (kh*3+kw)*3 + ci] = W[co, ci, kh, kw]
        m = torch.zeros(1, 64, 64, device=w.device, dtype=torch.bfloat16)
        m[0, :, :27] = w.permute(0, 2, 3, 1).reshape(64, 27).to(torch.bfloat16)
        return m


def _plans(module, cls):
    p = getattr(module, "_ipr_plans", None)
    if p is None:
        p = cls(module)
        object.__setattr__(module, "_ipr_plans", p)
    return p


# ----------------------------------------------------------------------------------------- generator
class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, z, fc_w, fc_b, w1, g1, b1, w2, g2, b2, w3, g3, b3, w4):
        P = _plans(module, GenPlans)
        B = z.shape[0]
        mg = P.mg
        bns = [module.convs[i][1] for i in range(3)]
        cts = (w1, w2, w3)
        a0 = z.detach().to(torch.bfloat16).contiguous().view(B, 1, 1, -1)
        perm = P.perm_on(z.device)
        fcp = P.packs.get("fc", fc_w, lambda w: P.fc.pack(w, perm))
        fcb = P.packs.get("fcb", fc_b, lambda b: b[perm].contiguous())
        h, _ = P.fc.run(a0, fcp, epi=dense.EPI_BIAS_LRELU, slope=0.0, bias=fcb)
        acts = [h.view(B, mg, mg, 512)]
        raws, means, rstds = [], [], []
        ctx.eval_stats = False
        for i in range(3):
            wp = P.packs.get("ct%d" % i, cts[i], P.ct[i].pack)
            bn = bns[i]
            batch_stats = bn.training or not bn.track_running_stats or bn.running_mean is None
            raw, stats = P.ct[i].run(acts[-1], wp, want_stats=batch_stats)
            if batch_stats:
                count = raw.numel() // raw.shape[-1]
                update = bn.training and bn.track_running_stats
                scale, shift, mean, rstd = bn_finalize(stats, count, bn, update)
            else:  # eval: running statistics
                rstd = torch.rsqrt(bn.running_var + bn.eps)
                mean = bn.running_mean
                scale = bn.weight.detach() * rstd
                shift = bn.bias.detach() - mean * scale
            acts.append(bn_apply_relu(raw, scale, shift))
            raws.append(raw)
            means.append(mean)
            rstds.append(rstd)
            ctx.eval_stats = not batch_stats
        w4p = P.packs.get("ct3", w4, P.last.pack)
        out, _ = P.last.run(acts[-1], w4p, epi=dense.EPI_TANH_NCHW, n_valid=3)
        ctx.module = module
        ctx.save_for_backward(a0, out, fc_w, w1, w2, w3, w4, g1, g2, g3, *acts, *raws, *means, *rstds)
        return out

    @staticmethod
    def backward(ctx, dout):
        module = ctx.module
        P = _plans(module, GenPlans)
        sv = ctx.saved_tensors
        a0, out, fc_w, w1, w2, w3, w4, g1, g2, g3 = sv[:10]
        acts, raws, means, rstds = sv[10:14], sv[14:17], sv[17:20], sv[20:23]
        cts, gammas = (w1, w2, w3), (g1, g2, g3)
        if ctx.eval_stats:
            raise RuntimeError("generator backward in eval mode (running statistics) is not on the training path")
        B = a0.shape[0]
        dev = a0.device
        # last layer: Tanh' fused into the patch gather, then one GEMM for d(a3) and one for dW4
        col = im2col3(dout.contiguous(), out)
        dw4 = torch.empty_like(w4)
        P.last_wg.run(acts[3], col, dw4)
        w4d = P.packs.get("ct3_dg", w4, P.pack_last_dgrad)
        d_act, _ = P.last_dg.run(col, w4d)
        dws, dgs, dbs = [None] * 3, [None] * 3, [None] * 3
        sign_hook = getattr(module, "_ipr_sign_hook", None)
        for i in (2, 1, 0):
            dgs[i] = torch.empty_like(gammas[i])
            dbs[i] = torch.empty_like(gammas[i])
            sg, g0, sc = (None, 0.0, 0.0)
            if sign_hook is not None:
                sg, g0, sc = sign_hook(i)
            dx = bn_relu_bwd(d_act, raws[i], acts[i + 1], gammas[i], means[i], rstds[i], dgs[i], dbs[i], False, sg, g0, sc)
            dws[i] = torch.empty_like(cts[i])
            P.ct_wg[i].run(dx, acts[i], dws[i])
            wd = P.packs.get("ct%d_dg" % i, cts[i], P.ct_dg[i].pack)
            if i > 0:
                d_act, _ = P.ct_dg[i].run(dx, wd)
            else:  # into the Linear's ReLU
                d_act, _ = P.ct_dg[i].run(dx, wd, epi=dense.EPI_MASK, slope=0.0, mask=acts[0])
        dh = d_act.view(B, 1, 1, -1)
        dfc_w = torch.empty_like(fc_w)
        P.fc_wg.run(dh, a0, dfc_w)
        perm = P.perm_on(dev)
        dfc_b = torch.empty(fc_w.shape[0], device=dev, dtype=torch.float32)
        dfc_b[perm] = dh.view(B, -1).float().sum(0)
        return (None, None, dfc_w, dfc_b, dws[0], dgs[0], dbs[0], dws[1], dgs[1], dbs[1], dws[2], dgs[2], dbs[2], dw4)


def generator_forward(module, z):
    z = z.to(device=module.fc[0].weight.device, dtype=torch.float32)
    c = module.convs
    return _GeneratorFn.apply(module, z, module.fc[0].weight, module.fc[0].bias,
                              c[0][0].weight, c[0][1].weight, c[0][1].bias,
                              c[1][0].weight, c[1][1].weight, c[1][1].bias,
                              c[2][0].weight, c[2][1].weight, c[2][1].bias, c[3].weight)


# ----------------------------------------------------------------------------------------- discriminator
def _sn_layers(module):
    net = module.net
    return [net[0][0], net[0][2], net[1][0], net[1][2], net[2][0], net[2][2], net[3], net[6]]


def _power_iteration(layer, training, eps=1e-12):
    """torch.nn.utils.spectral_norm (legacy hook) semantics: one iteration per training forward, in place on the
    u / v buffers; sigma = u . (W v) with the updated vectors.  Returns (sigma 0-dim, u, v) detached."""
    w = layer.weight_orig.detach()
    mat = w.reshape(w.shape[0], -1)
    u, v = layer.weight_u, layer.weight_v
    with torch.no_grad():
        if training:
            nv = torch.mv(mat.t(), u)
            v.copy_(nv / nv.norm().clamp_min(eps))
            nu = torch.mv(mat, v)
            u.copy_(nu / nu.norm().clamp_min(eps))
        sigma = torch.dot(u, torch.mv(mat, v))
    return sigma, u.clone(), v.clone()


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, param_grads, x, *params):
        P = _plans(module, DisPlans)
        layers = _sn_layers(module)
        ws, bs = params[0::2], params[1::2]
        B = x.shape[0]
        sig = [_power_iteration(l, module.training) for l in layers]
        xin = x.detach().contiguous()
        col = im2col3(xin)
        w1p = P.packs.get("c0", ws[0], P.pack_first)
        a, _ = P.first.run(col, w1p, epi=dense.EPI_BIAS_LRELU, slope=0.1, sigma=sig[0][0], bias=bs[0].detach())
        acts = [a]
        for i, plan in enumerate(P.conv):
            wp = P.packs.get("c%d" % (i + 1), ws[i + 1], plan.pack)
            a, _ = plan.run(acts[-1], wp, epi=dense.EPI_BIAS_LRELU, slope=0.1, sigma=sig[i + 1][0], bias=bs[i + 1].detach())
            acts.append(a)
        perm = P.perm_on(x.device)
        w8 = P.packs.get("fcw", ws[7], lambda w: w.reshape(-1)[perm].contiguous())
        logits = dfc_fwd(acts[-1].view(B, -1), w8, sig[7][0], bs[7].detach())
        ctx.module, ctx.param_grads = module, param_grads
        ctx.x_needs_grad = x.requires_grad
        flat = [t for s in sig for t in s]
        ctx.save_for_backward(col, w8, *ws, *acts, *flat)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        module = ctx.module
        P = _plans(module, DisPlans)
        sv = ctx.saved_tensors
        col, w8 = sv[0], sv[1]
        ws, acts = sv[2:10], sv[10:17]
        flat = sv[17:]
        sig = [flat[3 * i:3 * i + 3] for i in range(8)]
        want = ctx.param_grads
        B = col.shape[0]
        dev = col.device
        dlogits = dlogits.contiguous().float()
        gW, gB = [None] * 8, [None] * 8

        def sn_correct(G, i):
            # d/dW_orig of W_orig / sigma(W_orig), sigma = u^T W v  (u, v constants):  (G - <G, W_sn> u v^T) / sigma
            sigma, u, v = sig[i]
            W = ws[i]
            Gm = G.reshape(G.shape[0], -1)
            inner = (Gm * W.reshape(W.shape[0], -1)).sum() / sigma
            return ((Gm - inner * torch.outer(u, v)) / sigma).reshape(W.shape)

        a7 = acts[-1].view(B, -1)
        dy, dw8 = dfc_bwd(a7, w8, sig[7][0], dlogits, want, 0.1)
        if want:
            perm = P.perm_on(dev)
            g8 = torch.empty_like(dw8)
            g8[perm] = dw8
            gW[7] = sn_correct(g8.view(1, -1), 7)
            gB[7] = dlogits.sum().view(1)
        dy = dy.view(acts[-1].shape)
        for i in range(5, -1, -1):                 # conv layers 7..2 (index i+1 in the layer list)
            li = i + 1
            if want:
                gB[li] = dy.float().sum((0, 1, 2))
                G = torch.empty_like(ws[li])
                P.conv_wg[i].run(dy, acts[i], G)
                gW[li] = sn_correct(G, li)
            wd = P.packs.get("c%d_dg" % li, ws[li], P.conv_dg[i].pack)
            dy, _ = P.conv_dg[i].run(dy, wd, epi=dense.EPI_MASK, slope=0.1, mask=acts[i], sigma=sig[li][0])
        if want:
            gB[0] = dy.float().sum((0, 1, 2))
            G = torch.empty_like(ws[0])
            P.first_wg.run(dy, col, G)
            gW[0] = sn_correct(G, 0)
        dx = None
        if ctx.x_needs_grad:
            wd = P.packs.get("c0_dg", ws[0], P.first_dg.pack)
            dx, _ = P.first_dg.run(dy, wd, epi=dense.EPI_LINEAR_NCHW, sigma=sig[0][0], n_valid=3)
        grads = []
        for i in range(8):
            grads += [gW[i], gB[i]]
        return (None, None, dx, *grads)


def discriminator_forward(module, x):
    layers = _sn_layers(module)
    x = x.to(device=layers[0].weight_orig.device, dtype=torch.float32)
    params = []
    for l in layers:
        params += [l.weight_orig, l.bias]
    want = not getattr(module, "_ipr_skip_param_grads", False)
    return _DiscriminatorFn.apply(module, want, x, *params)


# ----------------------------------------------------------------------------------------- smoke
def smoke_step(orc):
    """One tiny generator/discriminator forward+backward on cuda:0 against the oracle's fp32 networks
    (bf16 tolerance 2e-2 of the tensor scale)."""
    import networks
    torch.manual_seed(7)
    G = networks.ConvGenerator32().cuda()
    D = networks.SNDiscriminator32().cuda()
    Go, Do = orc.make_generator(), orc.make_discriminator()
    Go.load_state_dict(G.state_dict())
    Do.load_state_dict(D.state_dict())
    z = torch.randn(8, 128)
    fake = generator_forward(G, z.cuda())
    logits = discriminator_forward(D, fake)
    (-logits.mean()).backward()
    fo = Go(z)
    lo = Do(fo)
    (-lo.mean()).backward()

    def rel(a, b):
        a, b = a.detach().cpu().float(), b.detach().float()
        return float((a - b).norm() / (b.norm() + 1e-12))

    assert rel(fake, fo) < 2e-2, rel(fake, fo)
    assert rel(logits, lo) < 6e-2, rel(logits, lo)      # two chained bf16 networks
