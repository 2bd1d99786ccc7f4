"""The activation calibrators behind `tensor_cali_dispatcher` — same registry keys,
signatures and return values as dipoorlet/tensor_cali/basic_algorithm.py:8-91 — computed
from device-resident statistics (forward_net.CalibrationSession) instead of per-image
NumPy lists. With several ranks the statistics (not the clip values) are combined, see
dist_helper.py."""
import numpy as np

from .. import dist_helper
from .. import forward_net as fwd
from ..platform_settings import LAYER_HAS_WEIGHT
from ..utils import dispatch_functool, logger


@dispatch_functool
def tensor_cali_dispatcher(*args, **kwargs):
    logger.info("Calibration Algorithm Not Found!")


def _global_images(sess, t):
    return dist_helper.allgather_images(t).cpu().numpy()


def _release(sess):
    """The calibrator is done with the blobs: hand the slab back before the weight transforms and the
    profiling allocate their activation caches (the statistics stay on the session)."""
    sess.resident = None
    sess.arena = None


@tensor_cali_dispatcher.register('minmax')
def find_clip_val_minmax(onnx_graph, args, **kwargs):
    """[min over images, max over images] per blob (basic_algorithm.py:13-22)."""
    sess = fwd._session(onnx_graph, args, fresh=True)
    sess.run_minmax(per_image=False, keep_for_hist=False)   # only the range over all images is used (:20-21)
    dist_helper.allreduce_minmax(sess.blob_min, sess.blob_max)
    lo, hi = sess.blob_min.cpu().numpy(), sess.blob_max.cpu().numpy()
    _release(sess)
    return {name: [lo[i], hi[i]] for i, name in enumerate(sess.names)}


@tensor_cali_dispatcher.register('hist')
def find_clip_val_hist(onnx_graph, args, store_stats=None, **kwargs):
    """Percentile of the |x| histogram (basic_algorithm.py:25-54): pass 1 range, pass 2
    histogram with the global range, K3 search on the device."""
    sess = fwd._session(onnx_graph, args, fresh=True)
    bins = int(args.bins)
    if store_stats:
        import torch
        mm, hist = store_stats['minmax'], store_stats['hist']
        sess.blob_min.copy_(torch.tensor([np.min(mm[n]['min']) for n in sess.names], dtype=torch.float32))
        sess.blob_max.copy_(torch.tensor([np.max(mm[n]['max']) for n in sess.names], dtype=torch.float32))
        from .. import kernels as K
        sess.data_max = torch.empty(sess.n_stats, dtype=torch.float32, device=sess.device)
        K.absmax(sess.blob_min, sess.blob_max, sess.data_max)
        sess.counts = torch.from_numpy(np.stack([np.asarray(hist[n], dtype=np.int64) for n in sess.names])
                                       ).to(sess.device).contiguous()
    else:
        sess.run_minmax(per_image=False)
        sess.run_hist(bins)
    clip, sel = sess.percentile_clip(bins, args.threshold)
    clip = clip.cpu().numpy()
    _release(sess)
    return {name: [clip[i, 0], clip[i, 1]] for i, name in enumerate(sess.names)}


@tensor_cali_dispatcher.register('mse')
def find_clip_val_octav(onnx_graph, args, **kwargs):
    """OCTAV (basic_algorithm.py:57-69): per-image fixed point on the device (K4), then the
    reference's own float32 mean over images on the host (123 x N values)."""
    fwd.forward_net_octav(onnx_graph, args)
    sess = fwd._session(onnx_graph, args)
    s = _global_images(sess, sess.seg_s)
    mx = _global_images(sess, sess.seg_max)
    mn = _global_images(sess, sess.seg_min)
    _release(sess)
    clip_val = {}
    for i, name in enumerate(sess.names):
        data_max, data_min = mx[i].max(), mn[i].min()
        m = s[i].mean()
        clip_val[name] = [max(data_min, -m), min(data_max, m)]
    return clip_val


def find_clip_val_minmax_weight(onnx_graph, args):
    """Per-output-channel min / max of every weight-like initializer (basic_algorithm.py:72-91).
    On the GPU this is ONE K1 launch over the weights the engine already holds in HBM (a channel
    is a segment); the host NumPy path remains for ConvTranspose weights (transposed first) and
    for the CPU unit tests."""
    names, transposed = [], set()
    for node in onnx_graph.graph.node:
        if node.op_type in LAYER_HAS_WEIGHT:
            for name in node.input[1:]:
                if name not in names:
                    names.append(name)
            if node.op_type == 'ConvTranspose':
                transposed.add(node.input[1])
    names = [n for n in names if onnx_graph.get_initializer(n).ndim >= 1]
    out = {}
    dev = fwd._device_of(args)
    on_gpu = [n for n in names if n not in transposed and onnx_graph.get_initializer(n).dtype == np.float32
              and onnx_graph.get_initializer(n).size > 0]
    import torch
    if dev.type != "cuda" or not torch.cuda.is_available():
        on_gpu = []                      # host NumPy, as the reference (weights are a8-small)
    if on_gpu:
        from .. import kernels as K
        eng = fwd._engine_for(onnx_graph, dev)
        tensors = []
        for n in on_gpu:
            t = eng.params[n]
            tensors.append(t.reshape(t.shape[0], -1) if t.dim() > 1 else t.reshape(-1, 1))
        batch = K.BlobBatch(tensors)
        lo = torch.empty(batch.n_segments, dtype=torch.float32, device=dev)
        hi = torch.empty_like(lo)
        K.segstats(batch, lo, hi)
        lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
        for n, (off, c) in zip(on_gpu, batch.seg_slices()):
            out[n] = [lo[off:off + c].copy(), hi[off:off + c].copy()]
    res = {}
    for n in names:                      # the reference's insertion order
        if n in out:
            res[n] = out[n]
            continue
        t = onnx_graph.get_initializer(n)
        if n in transposed:
            t = t.transpose([1, 0, 2, 3])
        flat = t.reshape((t.shape[0], -1))
        res[n] = [flat.min(-1), flat.max(-1)]
    return res
