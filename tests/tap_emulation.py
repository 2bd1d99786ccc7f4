"""NumPy emulation of the tap-table kernels' index arithmetic (dpl_pad_plane_f32, dpl_tap_conv_tf32,
dpl_tap_wgrad_tf32) — test infrastructure: lets the host-side geometry (kernels.ReconConvPlan: tap shifts,
parity classes, staging pitches) be checked against torch's convolution gradients without a GPU."""
import numpy as np


def pad_plane(x, stride, origin, hp, wp, planes):
    n, c, h, w = x.shape
    out = np.zeros((planes * n * hp * wp, c), dtype=x.dtype)
    for pl in range(planes):
        a, b = pl // stride, pl % stride
        for i in range(hp):
            for j in range(wp):
                hh, ww = stride * (i - origin) + a, stride * (j - origin) + b
                if i >= origin and j >= origin and hh < h and ww < w:
                    rows = (pl * n + np.arange(n)) * hp * wp + i * wp + j
                    out[rows] = x[:, :, hh, ww]
    return out


def _rows(mat, start, count):
    """mat[start:start+count] with zero fill outside (TMA out-of-bounds semantics)."""
    out = np.zeros((count, mat.shape[1]), dtype=mat.dtype)
    lo, hi = max(start, 0), min(start + count, mat.shape[0])
    if hi > lo:
        out[lo - start:hi - start] = mat[lo:hi]
    return out


def tap_conv(xp, wt, n, cn, H, W, hp, wp, origin, os_, oa, ob, shifts, taps, y):
    q_total = n * hp * wp
    acc = np.zeros((q_total, cn), dtype=np.float64)
    for sh, t in zip(shifts, taps):
        acc += _rows(xp, sh, q_total).astype(np.float64) @ wt[t].astype(np.float64).T
    for q in range(q_total):
        img, r = divmod(q, hp * wp)
        i, j = divmod(r, wp)
        hq, wq = i - origin, j - origin
        ho, wo = hq * os_ + oa, wq * os_ + ob
        if hq >= 0 and wq >= 0 and ho < H and wo < W:
            y[img, :, ho, wo] = acc[q]
    return y


def tap_wgrad(gp, xp, co, ci, t_full, shifts, cols):
    q_total = gp.shape[0]
    dw = np.zeros((co, ci, t_full), dtype=np.float64)
    for sh, col in zip(shifts, cols):
        dw[:, :, col] = gp.astype(np.float64).T @ _rows(xp, sh, q_total).astype(np.float64)
    return dw
