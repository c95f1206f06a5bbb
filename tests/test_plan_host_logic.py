"""CPU: the host-side planning of the tcgen05 layers -- tap tables, parity sub-grids, output-parity phases, weight
packing (`dense.Plan`) and the weight-gradient column tables (`dense.WGradPlan`) -- executed by a plain PyTorch
emulation of what the kernels do with them (include/ipr_b200.h: ipr_tapgemm_t / ipr_wgrad_t) and compared with
torch's own convolutions.  No GPU, no library call: this pins the descriptors the CUDA kernels are driven with."""
import pytest
import torch
import torch.nn.functional as F

from ipr_gan_b200 import dense, engine


def shift(src, dh, dw):
    """out[n, y, x] = src[n, y + dh, x + dw], zero outside (TMA out-of-bounds fill)."""
    N, H, W, C = src.shape
    out = torch.zeros_like(src)
    ys, xs = slice(max(0, -dh), min(H, H - dh)), slice(max(0, -dw), min(W, W - dw))
    yd, xd = slice(max(0, dh), min(H, H + dh)), slice(max(0, dw), min(W, W + dw))
    out[:, ys, xs] = src[:, yd, xd]
    return out


def grid_of(t, parity, mp):
    return t[:, (mp // 2)::2, (mp % 2)::2] if parity else t


def emulate_tapgemm(plan, a, b):
    """a: (N, H, W, C) fp32 NHWC, b: [phases][n_total][taps*C] -> (N, oh, ow, n_total)."""
    N, H, W, C = a.shape
    oh, ow = plan.out_hw(H, W)
    out = torch.zeros(N, oh, ow, plan.n_total, dtype=a.dtype)
    for ph, taps in enumerate(plan.taps):
        acc = 0
        for t, (mp, dh, dw, _, _) in enumerate(taps):
            acc = acc + shift(grid_of(a, plan.a_parity, mp), dh, dw) @ b[ph][:, t * C:(t + 1) * C].t()
        o = plan.out_o[ph]
        out[:, o[0]::plan.out_s, o[1]::plan.out_s] = acc
    return out


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


CASES = [("conv3", 8, 12, 8), ("conv4s2", 8, 16, 8), ("convT4s2", 16, 8, 4), ("convT3", 8, 3, 8),
         ("conv3_dgrad", 12, 8, 8), ("conv4s2_dgrad", 16, 8, 4), ("convT4s2_dgrad", 8, 16, 8)]


@pytest.mark.parametrize("kind,C,O,H", CASES)
def test_tap_tables_and_packing(kind, C, O, H):
    torch.manual_seed(len(kind) + C)
    x = torch.randn(2, C, H, H, dtype=torch.float64)
    plan = dense.Plan(kind, C, O, n_pad=16)
    if kind == "conv3":
        w = torch.randn(O, C, 3, 3, dtype=torch.float64); want = F.conv2d(x, w, padding=1)
    elif kind == "conv4s2":
        w = torch.randn(O, C, 4, 4, dtype=torch.float64); want = F.conv2d(x, w, stride=2, padding=1)
    elif kind == "convT4s2":
        w = torch.randn(C, O, 4, 4, dtype=torch.float64); want = F.conv_transpose2d(x, w, stride=2, padding=1)
    elif kind == "convT3":
        w = torch.randn(C, O, 3, 3, dtype=torch.float64); want = F.conv_transpose2d(x, w, stride=1, padding=1)
    elif kind == "conv3_dgrad":          # x plays dY (C = conv out channels), result = dX of Conv2d(O, C, 3, 1, 1)
        w = torch.randn(C, O, 3, 3, dtype=torch.float64); want = F.conv_transpose2d(x, w, stride=1, padding=1)
    elif kind == "conv4s2_dgrad":        # dX of Conv2d(O -> C, 4, 2, 1): weight (C, O, 4, 4)
        w = torch.randn(C, O, 4, 4, dtype=torch.float64); want = F.conv_transpose2d(x, w, stride=2, padding=1)
    else:                                # convT4s2_dgrad: dX of ConvTranspose2d(O -> C, 4, 2, 1): weight (O, C, 4, 4)
        w = torch.randn(O, C, 4, 4, dtype=torch.float64); want = F.conv2d(x, w, stride=2, padding=1)
    got = emulate_tapgemm(plan, nhwc(x), plan.pack_layout(w))
    assert got.shape[-1] == 16 and torch.all(got[..., O:] == 0)          # padded columns stay zero
    torch.testing.assert_close(got[..., :O].permute(0, 3, 1, 2), want, rtol=1e-10, atol=1e-10)


def test_linear_packing_with_row_permutation():
    torch.manual_seed(3)
    perm = torch.randperm(32)
    plan = dense.Plan("linear", 8, 32)
    x, w = torch.randn(5, 8, dtype=torch.float64), torch.randn(32, 8, dtype=torch.float64)
    got = emulate_tapgemm(plan, x.view(5, 1, 1, 8), plan.pack_layout(w, perm)).view(5, 32)
    torch.testing.assert_close(got, (x @ w.t())[:, perm])


def emulate_wgrad(wg, y, x, weight_shape):
    """ws[p][n][t*Cx + c] = sum_m Y[pix_y(m)][n] * X[pix(m) + tap_t][c], then the dst_off scatter of ipr_wgrad_reduce."""
    fwd = wg.fwd
    Cx = wg.x_c
    grad = torch.zeros(weight_shape, dtype=y.dtype).reshape(-1)
    off = wg.dst_off_host.view(wg.n_phases, wg.k_total)
    for ph, taps in enumerate(fwd.taps):
        yq = grid_of(y, wg.y_parity, 2 * fwd.out_o[ph][0] + fwd.out_o[ph][1])
        for t, (mp, dh, dw, _, _) in enumerate(taps):
            xs = shift(grid_of(x, wg.x_parity, mp), dh, dw)
            part = torch.einsum("nhwo,nhwc->oc", yq, xs)                 # [n][c]
            for c in range(Cx):
                o = int(off[ph, t * Cx + c])
                if o >= 0:
                    idx = torch.arange(wg.rows) * wg.s_n + o
                    grad[idx] += part[:, c]
    return grad.view(weight_shape)


@pytest.mark.parametrize("kind,C,O,H", [("conv3", 8, 16, 8), ("conv4s2", 8, 16, 8), ("convT4s2", 8, 16, 4)])
def test_weight_gradient_tables(kind, C, O, H):
    torch.manual_seed(C + O)
    x = torch.randn(2, C, H, H, dtype=torch.float64)
    ksz = 3 if kind == "conv3" else 4
    shape = (C, O, ksz, ksz) if kind == "convT4s2" else (O, C, ksz, ksz)
    w = torch.randn(shape, dtype=torch.float64, requires_grad=True)
    if kind == "conv3":
        out = F.conv2d(x, w, padding=1)
    elif kind == "conv4s2":
        out = F.conv2d(x, w, stride=2, padding=1)
    else:
        out = F.conv_transpose2d(x, w, stride=2, padding=1)
    dy = torch.randn_like(out)
    out.backward(dy)
    plan = dense.Plan(kind, C, O)
    wg = dense.WGradPlan(plan, shape)
    got = emulate_wgrad(wg, nhwc(dy), nhwc(x), shape)
    torch.testing.assert_close(got, w.grad, rtol=1e-10, atol=1e-10)
    # the destination-major reduction (large weights) addresses grad[n*s_n + c*s_c + j] with tap_of[j] = phase*n_taps + tap:
    # every destination tap must be produced by exactly one (phase, tap)
    kk = ksz * ksz
    tap_of = [-1] * kk
    for ph, taps in enumerate(plan.taps):
        for t, (_, _, _, kh, kw) in enumerate(taps):
            tap_of[kh * ksz + kw] = ph * plan.n_taps + t
    assert sorted(tap_of) == list(range(kk))
    if wg.tap_of is not None:
        assert list(wg.tap_of) == tap_of


def test_three_channel_tables():
    """27-column patch tables of the 3-channel layers (stored 32 wide) and the tap-expanded (27 rows) packing."""
    w = torch.arange(64 * 27, dtype=torch.float64).view(64, 3, 3, 3)
    m = engine._patch27_layout(w)
    assert m.shape == (1, 64, 32) and torch.all(m[0, :, 27:] == 0)
    for kh in range(3):
        for kw in range(3):
            for c in range(3):
                assert torch.equal(m[0, :, (kh * 3 + kw) * 3 + c], w[:, c, kh, kw])
    off = engine._col_off_patch27(True)
    assert off.shape == (1, 32) and int(off[0, 27]) == -1
    assert int(off[0, (2 * 3 + 1) * 3 + 2]) == 2 * 9 + 2 * 3 + 1            # (kh=2, kw=1, c=2) -> c*9 + kh*3 + kw
    r = engine._tap27_rows_layout(w)
    assert r.shape == (1, 32, 64) and torch.all(r[0, 27:] == 0)
    assert torch.equal(r[0, (1 * 3 + 2) * 3 + 1], w[:, 1, 1, 2])            # row (kh=1, kw=2, c=1) holds W[:, 1, 1, 2]


def _emulate_gather_pack(arena_values, index):
    """What csrc/optim.cu: gather_pack_kernel does: out[i] = bf16(arena[index[i]]), 0 where index < 0."""
    out = torch.zeros(index.numel(), dtype=torch.float32)
    ok = index >= 0
    out[ok] = arena_values[index[ok].long()]
    return out


def test_one_launch_packing_tables_cover_every_layer():
    """engine.PackSet derives ONE gather table per network from the layers' layout functions applied to parameter
    indices.  Executed on the CPU, the table must reproduce every layer's packed bf16 operand from the flat parameter
    arena (the kernel itself is a plain gather, csrc/optim.cu)."""
    import networks
    torch.manual_seed(5)
    G, D = networks.ConvGenerator32(), networks.SNDiscriminator32()
    for module, plans_cls in ((G, engine.GenPlans), (D, engine.DisPlans)):
        plans = plans_cls(module)
        packs = plans.packs
        flat_params = packs.arena.param.detach().reshape(-1)
        packed = _emulate_gather_pack(flat_params, packs.index.cpu())
        checked = 0
        for key, (off, n, shape) in packs.slices.items():
            got = packed[off:off + n].view(shape)
            # recompute the layout directly from the parameter values with the same layout function
            want = None
            if plans_cls is engine.GenPlans:
                cv = module.convs
                table = {"fc": (module.fc[0].weight, lambda w: plans.fc.pack_layout(w, plans.perm)),
                         "ct3": (cv[3].weight, engine._tap27_rows_layout), "ct3_dg": (cv[3].weight, engine._patch27_layout)}
                for i in range(3):
                    table["ct%d" % i] = (cv[i][0].weight, plans.ct[i].pack_layout)
                    table["ct%d_dg" % i] = (cv[i][0].weight, plans.ct_dg[i].pack_layout)
                param, fn = table[key]
                want = fn(param.detach())
            else:
                layers = engine._sn_layers(module)
                if key == "c0":
                    want = engine._patch27_layout(layers[0].weight_orig.detach())
                elif key == "c0_dg":
                    want = engine._tap27_rows_layout(layers[0].weight_orig.detach())
                else:
                    i = int(key[1:].split("_")[0])
                    plan = plans.conv_dg[i - 1] if key.endswith("_dg") else plans.conv[i - 1]
                    want = plan.pack_layout(layers[i].weight_orig.detach())
            assert got.shape == want.shape, key
            assert torch.equal(got, want.float()), key
            checked += 1
        assert checked == len(packs.slices) and checked >= 9
