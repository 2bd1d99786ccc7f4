"""The reference's calibration job end to end on the CPU (tensor_calibration ->
save/load clip values -> trt deploy), assembled from the restated pieces. Used as the
checker in tests/ and as the timed CPU baseline of bench.py (`cpu_baseline`,
`--impl reference`): per image one fp32 forward + the reference's single-threaded NumPy
statistics (dipoorlet/forward_net.py:192-342), sharded over worker processes the way the
reference shards over ranks (forward_net.py:207-209)."""
import multiprocessing as mp
import os
import time

import numpy as np

from . import forward as OF
from . import stats as OS


def calibrate(model, images, algo, bins=2048, threshold=0.99999, threads=None):
    """images: float32 [n, 1, C, H, W] for the single network input. -> act_clip_val."""
    g = model.graph
    in_name = [vi.name for vi in g.inputs if vi.name not in g.initializers][0]
    n = images.shape[0]
    if algo == "mse":
        # one pass: forward + OCTAV per image
        blobs = OF.blobs_for_images(model, {in_name: images}, n, threads)
        return OS.clip_octav(OS.octav_stats(blobs))
    blobs = OF.blobs_for_images(model, {in_name: images}, n, threads)
    mm = OS.minmax_stats(blobs)
    if algo == "minmax":
        return OS.clip_minmax(mm)
    # the reference runs the network a second time for the histogram pass
    blobs = OF.blobs_for_images(model, {in_name: images}, n, threads)
    return OS.clip_hist(mm, OS.hist_stats(blobs, mm, bins), bins, threshold)


_WORK = {}


def _worker(a):
    shard, algo, bins, threshold = a
    import torch
    torch.set_num_threads(1)
    t0 = time.perf_counter()
    calibrate(_WORK["model"], _WORK["images"][shard[0]:shard[1]], algo, bins, threshold, threads=1)
    return time.perf_counter() - t0


def timed_parallel(model, images, algo, procs, bins=2048, threshold=0.99999):
    """Wall time of `procs` single-threaded workers, each calibrating a contiguous shard
    (the reference's own data parallelism, one process per rank). Returns seconds."""
    n = images.shape[0]
    procs = max(1, min(procs, n))
    per = n // procs
    shards = [(i * per, (i + 1) * per) for i in range(procs)]
    _WORK["model"], _WORK["images"] = model, images
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        pool.map(_worker, [(s, algo, bins, threshold) for s in shards])
    return time.perf_counter() - t0, per * procs
