"""NumPy stand-ins for dipoorlet_b200.kernels, used ONLY to exercise the host-side logic
(batching, per-image bookkeeping, registry plumbing, multi-rank reductions) in the CPU test
suite. They are built from the oracle and never shipped."""
import numpy as np
import torch

from oracle import stats as O


class BlobBatch:
    def __init__(self, tensors, stat_index=None):
        self.tensors = list(tensors)
        self.n_blobs = len(self.tensors)
        self.n_segments = sum(t.shape[0] for t in self.tensors)
        self.device = self.tensors[0].device
        self.elements = int(sum(t.numel() for t in self.tensors))
        self.max_seg_len = max(t.numel() // t.shape[0] for t in self.tensors)


class Workspace:
    def __init__(self, device):
        pass


_n = 0


def launches():
    return _n


def _segs(batch):
    for b, t in enumerate(batch.tensors):
        for i in range(t.shape[0]):
            yield b, t[i].numpy()


def segstats(batch, seg_min, seg_max, seg_abssum=None, seg_nnz=None, blob_min=None, blob_max=None,
             workspace=None, ctas_per_sm=0):
    for k, (b, x) in enumerate(_segs(batch)):
        seg_min[k], seg_max[k] = float(x.min()), float(x.max())
        if seg_abssum is not None:
            seg_abssum[k] = float(np.abs(x).astype(np.float64).sum())
        if seg_nnz is not None:
            seg_nnz[k] = int((np.abs(x) > 0).sum())
        if blob_min is not None:
            blob_min[b] = min(float(blob_min[b]), float(x.min()))
            blob_max[b] = max(float(blob_max[b]), float(x.max()))


def absmax(blob_min, blob_max, data_max):
    data_max.copy_(torch.maximum(blob_max, -blob_min))


def hist_abs(batch, data_max, counts, bins, variant=0):
    for b, x in _segs(batch):
        h = np.histogram(np.abs(x), bins, (0, np.float32(data_max[b].item())))[0]
        counts[b] += torch.from_numpy(h)


def hist_percentile(counts, bins, threshold, data_max, blob_min, blob_max, clip, sel_bin=None):
    for b in range(counts.shape[0]):
        mm = {"x": {"min": [np.float32(blob_min[b].item())], "max": [np.float32(blob_max[b].item())]}}
        c, s = O.clip_hist(mm, {"x": [counts[b].numpy()]}, bins, threshold, return_bins=True)
        clip[b, 0], clip[b, 1] = float(c["x"][0]), float(c["x"][1])
        if sel_bin is not None:
            sel_bin[b] = s["x"]


def octav(batch, seg_abssum, seg_nnz, k_const, out_s, out_iters=None, max_iter=20, workspace=None):
    unsigned = int(round((1 / 4 ** 8 / 3) / k_const))       # k = 1 / 4**8 / 3 / unsigned (forward_net.py:319-328)
    for k, (b, x) in enumerate(_segs(batch)):
        out_s[k] = float(O.octav_stats({"x": [x]}, unsigned_of=lambda m: unsigned)["x"]["optimal_s"][0])


def fakequant(x, scale, zero_point=None, qlo=-128, qhi=127, axis=None, drop_prob=1.0, seed=0, out=None):
    """K5 stand-in: QuantizeLinear + DequantizeLinear of the ONNX operator spec (round half even, saturate),
    no QDrop (the host-logic tests never enable it)."""
    assert drop_prob >= 1.0
    shape = [1] * x.dim()
    if scale.numel() > 1:
        shape[axis] = -1
    s = scale.reshape(shape)
    zp = 0.0 if zero_point is None else zero_point.to(torch.float32).reshape(shape)
    y = (torch.clamp(torch.round(x / s) + zp, qlo, qhi) - zp) * s
    global _n
    _n += 1
    if out is not None:
        out.copy_(y)
        return out
    return y


def cosine3(a, b, out):
    """K7b stand-in: per segment (first axis) sum(a b), sum(a a), sum(b b) in float64, accumulated into out."""
    x, y = a.reshape(a.shape[0], -1).double(), b.reshape(b.shape[0], -1).double()
    out[:, 0] += (x * y).sum(1)
    out[:, 1] += (x * x).sum(1)
    out[:, 2] += (y * y).sum(1)
    global _n
    _n += 1
