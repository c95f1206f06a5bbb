"""NumPy restatement of the operation order of torch's CPU bicubic kernel (align_corners=False,
A=-0.75) that csrc/pdq.cu follows; established empirically against F.interpolate."""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def _cc1(x):
    p = f32(_fma(f32(1.25), x, f32(-2.25)))
    return f32(f32(p * x) * x + f32(1))


def _cc2(x):
    t = f32(_fma(f32(-0.75), x, f32(3.75)))
    t = f32(_fma(t, x, f32(-6)))
    return f32(t * x + f32(3))


def taps(insz, outsz):
    scale = f32(insz) / f32(outsz)
    idx = np.zeros((outsz, 4), np.int64)
    w = np.zeros((outsz, 4), f32)
    for i in range(outsz):
        real = f32(_fma(scale, f32(i) + f32(0.5), -f32(0.5)))
        i0 = min(int(np.floor(real)), insz - 1)
        t = f32(min(max(f32(real - f32(i0)), f32(0)), f32(1)))
        u = f32(f32(1) - t)
        w[i] = [_cc2(f32(t + f32(1))), _cc1(t), _cc1(u), _cc2(f32(u + f32(1)))]
        idx[i] = [min(max(i0 - 1 + j, 0), insz - 1) for j in range(4)]
    return idx, w


def _mix(t, w):
    r = _fma(t[0], w[0], t[1] * w[1])
    r = _fma(t[2], w[2], r)
    return _fma(t[3], w[3], r)


def bicubic_numpy(x, hout, wout):
    ih, wh = taps(x.shape[2], hout)
    iw, ww = taps(x.shape[3], wout)
    rows = []
    for a in range(4):
        xr = x[:, :, ih[:, a], :]
        t = [xr[..., iw[:, j]] for j in range(4)]
        wt = [np.broadcast_to(ww[:, j], t[0].shape) for j in range(4)]
        rows.append(_mix(t, wt))
    wcol = [np.broadcast_to(wh[:, a][None, None, :, None], rows[0].shape) for a in range(4)]
    return _mix(rows, wcol)
