"""Calibration forward passes on the GPU — the B200 replacement of
dipoorlet/forward_net.py.

The reference runs ONNXRuntime one image at a time, copies every node output (106 MB per
ResNet-50 image) back to the host and reduces it with single-threaded NumPy. Here a
`CalibrationSession` pushes a batch of images through the graph (`engine.Engine`), keeps
all blobs of the batch in HBM and reduces them in place with ONE launch per statistic over
the whole blob table (libdpl_b200.so: K1 segment stats, K2 histogram, K3 percentile, K4
OCTAV). Statistics stay on the device between passes; only the final per-blob numbers
(a few KB) travel to the host.

The public functions keep the reference names, arguments and return shapes
(forward_net.py:192-342): forward_get_minmax / forward_get_hist / forward_net_octav,
ActivationCache, input_data_generator, forward_get_tensor.
"""
import os

import numpy as np
import torch

from . import dist_helper
from . import kernels as K
from .engine import Engine, RangeSink
from .platform_settings import platform_setting_table
from .utils import logger


# ------------------------------------------------------------------------ inputs
def input_data_generator(input_dir, input_name_list, data_st_idx, data_ed_idx):
    """Same files as the reference: {input_dir}/{name}/{idx}.bin raw float32
    (forward_net.py:459-464)."""
    for idx in range(data_st_idx, data_ed_idx):
        yield {name: np.fromfile(f"{input_dir}/{name}/{idx}.bin", "float32")
               for name in input_name_list}


class ArrayInput:
    """In-memory calibration set: name -> float32 array [N, ...] (host, ideally pinned).
    Accepted wherever the reference takes `args.input_dir`; lets bench.py time the whole
    job from host buffers without going through the filesystem."""

    def __init__(self, arrays, pin=True, start=0):
        self.start = start  # global index of the first image held (a rank's shard)
        self.tensors = {}
        for name, arr in arrays.items():
            t = arr if torch.is_tensor(arr) else torch.from_numpy(np.ascontiguousarray(arr))
            if pin and torch.cuda.is_available() and not t.is_pinned():
                t = t.pin_memory()
            self.tensors[name] = t

    def fetch(self, name, st, ed, shape):
        t = self.tensors[name][st - self.start:ed - self.start]
        return t.reshape((ed - st,) + tuple(shape))


class FileInput:
    def __init__(self, input_dir):
        self.input_dir = input_dir

    def fetch(self, name, st, ed, shape):
        n = int(np.prod(shape))
        buf = torch.empty((ed - st, n), dtype=torch.float32,
                          pin_memory=torch.cuda.is_available())
        view = buf.numpy()
        for i, idx in enumerate(range(st, ed)):
            view[i] = np.fromfile(f"{self.input_dir}/{name}/{idx}.bin", "float32")
        return buf.reshape((ed - st,) + tuple(shape))


def as_input_source(input_dir):
    return input_dir if hasattr(input_dir, "fetch") else FileInput(input_dir)


def _device_of(args):
    """cuda:<local_rank>. `args._test_device` exists only so that tests/ can drive the host
    logic with mocked kernels on a box without a GPU; the CLI never sets it."""
    forced = getattr(args, "_test_device", None)
    return torch.device(forced) if forced else torch.device("cuda", getattr(args, "local_rank", 0) or 0)


def _engine_for(onnx_graph, dev):
    """One Engine (= one upload of the weights, 100 MB for ResNet-50) per graph object and
    device, reused by successive calibration calls on the same graph."""
    cache = onnx_graph.__dict__.setdefault("_dpl_engines", {})
    eng = cache.get(str(dev))
    version = getattr(onnx_graph, "_init_version", 0)
    if eng is None or eng.nodes != list(onnx_graph.model.graph.nodes):
        eng = cache[str(dev)] = Engine(onnx_graph, dev, _unit_test_cpu=dev.type != "cuda")
        eng._init_version = version
    elif getattr(eng, "_init_version", 0) != version:     # weights were edited since the upload
        eng.refresh_initializers()
        eng._init_version = version
    return eng


def shard_range(args):
    """forward_net.py:207-209: contiguous shard, floor division (the tail is dropped)."""
    rank_num = args.data_num // args.world_size
    st = args.rank * rank_num
    return st, min((args.rank + 1) * rank_num, args.data_num)


def _per_image_shape(onnx_graph, name):
    shape = list(onnx_graph.get_tensor_shape(name))
    if not shape or shape[0] not in (0, 1):
        raise NotImplementedError(
            f"network input {name!r} has leading dimension {shape[:1]}; the batched engine "
            "needs models exported with batch 1 (as every BASELINE.json config is)")
    return shape[1:]


def default_batch(onnx_graph, engine, budget_bytes=28 << 30, cap=256):
    per_img = 4 * sum(int(np.prod(onnx_graph.get_tensor_shape(n)[1:]))
                      for n in engine.blob_names() if n in onnx_graph.tensor_name_shape_map)
    return int(max(1, min(cap, budget_bytes // max(per_img, 1))))


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


# ------------------------------------------------------------------------ session
class CalibrationSession:
    """Device-resident calibration statistics of one rank's image shard."""

    def __init__(self, onnx_graph, args, engine=None):
        self.g = onnx_graph
        self.args = args
        dev = _device_of(args)
        self.device = dev
        self.engine = engine or _engine_for(onnx_graph, dev)
        self.names = self.engine.blob_names()
        self.n_stats = len(self.names)
        self.st, self.ed = shard_range(args)
        self.n_local = self.ed - self.st
        self.source = as_input_source(args.input_dir)
        self.batch_size = int(getattr(args, "calib_bs", 0) or default_batch(onnx_graph, self.engine))
        self.ws = K.Workspace(dev)
        self.in_shapes = {n: _per_image_shape(onnx_graph, n) for n in onnx_graph.network_inputs}
        self.h2d_bytes = 0
        self._copy_stream = None
        self._stats_stream = None
        self._slot_busy = [None, None]
        self._slot = 0
        self._keep_now = False      # this pass keeps its blobs resident (set by run_minmax)
        self.arena = None
        f32 = dict(dtype=torch.float32, device=dev)
        self.blob_min = torch.full((self.n_stats,), float("inf"), **f32)
        self.blob_max = torch.full((self.n_stats,), float("-inf"), **f32)
        self.seg_min = self.seg_max = self.seg_abssum = self.seg_nnz = self.seg_s = None
        self.range_ready = False    # blob_min / blob_max hold this shard's range
        self.data_max = None
        self.counts = None
        # Resident mode: when the shard's blobs fit in HBM (1024 ResNet-50 images = 109 GB of
        # the B200's 180 GB) pass 1 keeps every batch's blobs and pass 2 histograms them in
        # place instead of running the network a second time as the reference does.
        self.resident = None
        self.keep_resident = self._fits_resident() if getattr(args, "resident", None) is None \
            else bool(args.resident)

    def _fits_resident(self):
        """The slab of a resident pass (every batch's blobs, exactly what _attach_arena will take) plus a
        margin for the engine's weights / staging scratch and the statistics must fit in what is free now."""
        if self.device.type != "cuda":
            return False
        need = sum(self._batch_bytes(b1 - b0) for b0, b1 in self._ranges()) + (6 << 30)
        free, _ = torch.cuda.mem_get_info(self.device)
        reusable = torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        return need < 0.92 * (free + reusable)

    def _ranges(self):
        return [(b0, min(b0 + self.batch_size, self.ed)) for b0 in range(self.st, self.ed, self.batch_size)]

    def _batch_bytes(self, b):
        """Arena bytes of one forward batch of b images: every blob, 256-byte aligned."""
        al = K.BlobArena.ALIGN
        total = 0
        for n in self.names:
            if n in self.g.tensor_name_shape_map:
                nb = 4 * b * int(np.prod(self.g.get_tensor_shape(n)[1:]))
                total += (nb + al - 1) // al * al
        return total + (total >> 6) + (8 << 20)   # slack for blobs whose shape the graph does not declare

    def _attach_arena(self, ranges):
        """One slab for the blobs of this pass: all batches when they stay resident for pass 2,
        else one batch, recycled (stream order makes the reuse safe)."""
        if self.device.type != "cuda":
            return
        if self._keep_now:
            need = sum(self._batch_bytes(b1 - b0) for b0, b1 in ranges)
        else:
            # two batch-sized halves, used alternately: the statistics pass of batch i runs on a side
            # stream underneath the forward of batch i + 1
            self._half = max(self._batch_bytes(b1 - b0) for b0, b1 in ranges)
            self._half = (self._half + 255) // 256 * 256
            need = 2 * self._half
        if self.arena is None or self.arena.buf.numel() < need:
            self.arena = None               # release before taking the larger slab
            self.arena = K.BlobArena(need, self.device)
        self.arena.reset()
        self._slot_busy = [None, None]      # event after which a recompute-mode half may be overwritten

    def _upload(self, b0, b1):
        feeds = {}
        for name, shape in self.in_shapes.items():
            host = self.source.fetch(name, b0, b1, shape)
            self.h2d_bytes += host.numel() * 4
            feeds[name] = host.to(self.device, non_blocking=True)
        return feeds

    def batches(self, sink=None):
        """Forward batches with the NEXT batch's host->device copy issued on a side stream
        while the current batch computes (the images come from pinned host memory).
        `sink` (engine.RangeSink): fused range statistics, see run_minmax."""
        ranges = self._ranges()
        if self.device.type != "cuda" or not ranges:
            for b0, b1 in ranges:
                blobs = self.engine.run(self._upload(b0, b1), want="all")
                yield b0 - self.st, b1 - self.st, K.BlobBatch([blobs[n] for n in self.names])
            return
        self._attach_arena(ranges)
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            # two persistent device buffers per input, allocated on the main stream once: side-stream
            # allocations would live in their own allocator pool and force cudaMalloc / cache flushes
            # when the resident blobs already fill most of HBM
            self._in_bufs = [{n: torch.empty((self.batch_size,) + tuple(shp), dtype=torch.float32, device=self.device)
                              for n, shp in self.in_shapes.items()} for _ in range(2)]
        copy = self._copy_stream

        def prefetch(slot, rng):
            b0, b1 = rng
            copy.wait_stream(main)          # the buffer's previous reader (two batches back) is done
            feeds = {}
            with torch.cuda.stream(copy):
                for name, shape in self.in_shapes.items():
                    host = self.source.fetch(name, b0, b1, shape)
                    self.h2d_bytes += host.numel() * 4
                    dst = self._in_bufs[slot][name][:b1 - b0]
                    dst.copy_(host, non_blocking=True)
                    feeds[name] = dst
            ev = torch.cuda.Event()
            ev.record(copy)
            return feeds, ev

        nxt = prefetch(0, ranges[0])
        for i, (b0, b1) in enumerate(ranges):
            feeds, ev = nxt
            main.wait_event(ev)
            if not self._keep_now:
                self.arena.off = (i & 1) * self._half
                if self._slot_busy[i & 1] is not None:
                    main.wait_event(self._slot_busy[i & 1])
                    self._slot_busy[i & 1] = None
                self._slot = i & 1
            else:   # the input blob is kept for pass 2: detach it from the staging buffer
                kept = {}
                for k, v in feeds.items():
                    kept[k] = self.arena.alloc(v.shape)
                    kept[k].copy_(v)
                feeds = kept
            if i + 1 < len(ranges):
                nxt = prefetch((i + 1) & 1, ranges[i + 1])
            self.engine.arena = self.arena
            try:
                blobs = self.engine.run(feeds, want="all", stats=sink)
            finally:
                self.engine.arena = None
            yield b0 - self.st, b1 - self.st, K.BlobBatch([blobs[n] for n in self.names])

    # -- pass 1: min / max (+ moments for OCTAV) ------------------------------------
    def run_minmax(self, moments=False, octav_k=None, per_image=True, keep_for_hist=None, octav_k_zero_min=None):
        """Pass 1. per_image=False (the minmax / hist calibrators, which only use the range over all
        images, basic_algorithm.py:20-21,33-36): the forward's streaming kernels fold min / max of the
        blobs they write into blob_min / blob_max themselves, and K1 reads only the remaining blobs
        (convolution outputs, the input); seg_min / seg_max are then not produced. With DPL_STATS_OVERLAP=1 the statistics kernels of batch i are enqueued on a side
        stream and run underneath the forward of batch i + 1 (K1 on a reduced grid so that it fits beside
        the tensor-core CTAs). Measured on B200 (ResNet-50, batch 128): 6084 images/s with the overlap,
        6201 without - the forward's kernels lose more to the contention than the hidden K1 pass saves -
        so the default keeps everything on one stream."""
        n, dev = self.n_local, self.device
        # the blobs stay resident only when a histogram pass will read them again (keep_for_hist=False: the
        # minmax / mse calibrators, which must not take a 110 GB slab for nothing)
        self._keep_now = bool(self.keep_resident and octav_k is None and keep_for_hist is not False)
        fused = (not per_image and not moments and octav_k is None and dev.type == "cuda"
                 and os.environ.get("DPL_FUSED_RANGE", "1") != "0")
        if fused:
            return self._run_minmax_fused()
        self.seg_min = torch.empty((self.n_stats, n), dtype=torch.float32, device=dev)
        self.seg_max = torch.empty((self.n_stats, n), dtype=torch.float32, device=dev)
        if moments:
            self.seg_abssum = torch.empty((self.n_stats, n), dtype=torch.float64, device=dev)
            self.seg_nnz = torch.empty((self.n_stats, n), dtype=torch.int64, device=dev)
        if octav_k is not None:
            self.seg_s = torch.empty((self.n_stats, n), dtype=torch.float32, device=dev)
        keep = [] if self._keep_now else None
        overlap = dev.type == "cuda" and os.environ.get("DPL_STATS_OVERLAP", "0") == "1"
        main = torch.cuda.current_stream(dev) if dev.type == "cuda" else None
        if overlap and self._stats_stream is None:
            self._stats_stream = torch.cuda.Stream(device=dev)
        side = self._stats_stream if overlap else None
        parts = []          # per-batch outputs, alive until the side stream has been joined
        for lo, hi, batch in self.batches():
            if keep is not None:
                keep.append((lo, hi, batch))
            b = hi - lo
            smin = torch.empty(self.n_stats * b, dtype=torch.float32, device=dev)
            smax = torch.empty_like(smin)
            ssum = torch.empty(self.n_stats * b, dtype=torch.float64, device=dev) if moments else None
            snnz = torch.empty(self.n_stats * b, dtype=torch.int64, device=dev) if moments else None
            s = torch.empty(self.n_stats * b, dtype=torch.float32, device=dev) if octav_k is not None else None
            if side is not None:
                ready = torch.cuda.Event()
                ready.record(main)
                side.wait_event(ready)
            with torch.cuda.stream(side) if side is not None else _nullcontext():
                K.segstats(batch, smin, smax, ssum, snnz, self.blob_min, self.blob_max, self.ws,
                           ctas_per_sm=2 if side is not None else 0)
                if octav_k is not None:
                    K.octav(batch, ssum, snnz, octav_k, s, workspace=self.ws)
                    if octav_k_zero_min is not None:
                        # 'dynamic_sym' platforms (forward_net.py:318-322): a segment whose minimum is 0 gains a
                        # bit (unsigned = 4). The constant is per launch, so the fixed point is evaluated with both
                        # and each segment takes the one its own minimum selects.
                        s_alt = torch.empty_like(s)
                        K.octav(batch, ssum, snnz, octav_k_zero_min, s_alt, workspace=self.ws)
                        s = torch.where(smin.abs() < 1e-6, s_alt, s)
                if side is not None and not self._keep_now:
                    done = torch.cuda.Event()
                    done.record(side)
                    self._slot_busy[self._slot] = done
            parts.append((lo, hi, b, smin, smax, ssum, snnz, s, batch))
            del batch
        if side is not None:
            main.wait_stream(side)
        for lo, hi, b, smin, smax, ssum, snnz, s, _ in parts:
            self.seg_min[:, lo:hi] = smin.view(self.n_stats, b)
            self.seg_max[:, lo:hi] = smax.view(self.n_stats, b)
            if moments:
                self.seg_abssum[:, lo:hi] = ssum.view(self.n_stats, b)
                self.seg_nnz[:, lo:hi] = snnz.view(self.n_stats, b)
            if octav_k is not None:
                self.seg_s[:, lo:hi] = s.view(self.n_stats, b)
        del parts
        self.resident = keep
        self.range_ready = True

    def _run_minmax_fused(self):
        dev = self.device
        self.seg_min = self.seg_max = None
        sink = RangeSink(self.blob_min, self.blob_max, self.names)
        keep = [] if self._keep_now else None
        for lo, hi, batch in self.batches(sink=sink):
            if keep is not None:
                keep.append((lo, hi, batch))
            rest = [i for i, nm in enumerate(self.names) if nm not in sink.covered]
            if rest:
                sub = K.BlobBatch([batch.tensors[i] for i in rest], stat_index=rest)
                smin = torch.empty(sub.n_segments, dtype=torch.float32, device=dev)
                smax = torch.empty_like(smin)
                K.segstats(sub, smin, smax, None, None, self.blob_min, self.blob_max, self.ws)
            del batch
        self.resident = keep
        self.range_ready = True

    # -- pass 2: histogram ------------------------------------------------------------
    def run_hist(self, bins, variant=0):
        dist_helper.allreduce_minmax(self.blob_min, self.blob_max)   # global range
        self.data_max = torch.empty(self.n_stats, dtype=torch.float32, device=self.device)
        K.absmax(self.blob_min, self.blob_max, self.data_max)
        self.counts = torch.zeros((self.n_stats, int(bins)), dtype=torch.int64, device=self.device)
        events = getattr(self, "hist_events", None)  # bench.py: per-launch CUDA-event timing
        source = self.resident if self.resident is not None else self.batches()
        for lo, hi, batch in source:
            if events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            K.hist_abs(batch, self.data_max, self.counts, int(bins), variant)
            if events is not None:
                e1.record()
                events.append((e0, e1, batch.elements * 4))
            del batch
        self.resident = None  # release the blobs
        self._keep_now = False
        if self.keep_resident:
            self.arena = None
        dist_helper.allreduce_sum(self.counts)

    def percentile_clip(self, bins, threshold):
        clip = torch.empty((self.n_stats, 2), dtype=torch.float32, device=self.device)
        sel = torch.empty(self.n_stats, dtype=torch.int32, device=self.device)
        K.hist_percentile(self.counts, int(bins), float(threshold), self.data_max, self.blob_min,
                          self.blob_max, clip, sel)
        return clip, sel

    # -- host views in the reference's return shapes ------------------------------------
    def minmax_dict(self):
        mx, mn = self.seg_max.cpu().numpy(), self.seg_min.cpu().numpy()
        return {name: {"max": mx[i], "min": mn[i]} for i, name in enumerate(self.names)}


_SESSIONS = {}


def _session(onnx_graph, args, fresh=False):
    key = id(onnx_graph)
    if fresh or key not in _SESSIONS:
        _SESSIONS.clear()
        _SESSIONS[key] = CalibrationSession(onnx_graph, args)
    return _SESSIONS[key]


def forward_get_minmax(onnx_graph, args):
    """-> {name: {'max': float32[n_local], 'min': float32[n_local]}} (forward_net.py:192-237);
    per-image values in image order, as arrays instead of lists of scalars."""
    sess = _session(onnx_graph, args, fresh=True)
    # the blobs are kept for forward_get_hist (the reference's two-call sequence) when they fit
    sess.run_minmax()
    return sess.minmax_dict()


def forward_get_hist(onnx_graph, stats_min_max, args):
    """-> {name: [int64[bins]]}: the histogram of |x| over this job's images, already
    summed over images (and ranks), in a one-element list so that the reference's
    `np.stack(hist).sum(0)` (basic_algorithm.py:38) is the identity (forward_net.py:240-281)."""
    sess = _session(onnx_graph, args)
    if not sess.range_ready:   # stats came from elsewhere (store_stats): seed the range
        mn = np.array([np.min(stats_min_max[n]["min"]) for n in sess.names], np.float32)
        mx = np.array([np.max(stats_min_max[n]["max"]) for n in sess.names], np.float32)
        sess.blob_min.copy_(torch.from_numpy(mn))
        sess.blob_max.copy_(torch.from_numpy(mx))
    sess.run_hist(int(args.bins))
    counts = sess.counts.cpu().numpy()
    return {name: [counts[i]] for i, name in enumerate(sess.names)}


def forward_net_octav(onnx_graph, args):
    """-> {name: {'optimal_s': float32[n_local], 'min': ..., 'max': ...}} (forward_net.py:284-342)."""
    sess = _session(onnx_graph, args, fresh=True)
    dyn = "dynamic_sym" in platform_setting_table[args.deploy]["qi_params"]
    sess.run_minmax(moments=True, octav_k=1 / (4 ** 8) / 3 / 1,
                    octav_k_zero_min=(1 / (4 ** 8) / 3 / 4) if dyn else None)
    s, mx, mn = sess.seg_s.cpu().numpy(), sess.seg_max.cpu().numpy(), sess.seg_min.cpu().numpy()
    return {name: {"optimal_s": s[i], "min": mn[i], "max": mx[i]} for i, name in enumerate(sess.names)}


# ------------------------------------------------------------------------ activation cache
class ActivationCache:
    """Any tensor of the graph for all images of a shard, resident in HBM
    (forward_net.py:23-189 keeps NumPy lists on the host and builds one ORT session per
    node). `cache[name]` -> float32 CUDA tensor [n, ...]; initializers return the array."""

    def __init__(self, graph, args, st=None, ed=None, engine=None):
        self.graph = graph
        self.args = args
        self.st = 0 if st is None else st
        self.ed = args.data_num if st is None else ed
        self.device = _device_of(args)
        self.engine = engine or Engine(graph, self.device, _unit_test_cpu=self.device.type != "cuda")
        self.source = as_input_source(args.input_dir)
        self.batch_size = int(getattr(args, "calib_bs", 0) or 64)
        self.activation_cache = {}
        self._pos_graph, self._pos_len = None, -1
        self.in_shapes = {n: _per_image_shape(graph, n) for n in graph.network_inputs}

    def update_graph(self, graph):
        """New weights, same topology (forward_net.py:182-189): every cached activation is
        dropped. Prefer `update_initializers`, which keeps what is still valid."""
        self.graph = graph
        self.engine.g = graph
        self.engine.refresh_initializers()
        self.activation_cache.clear()

    def update_initializers(self, names, node):
        """The initializers `names` of `node` changed (a rounded weight, a corrected bias):
        re-upload them and drop exactly the cached activations downstream of `node` — what
        the reference's incremental `prev_act_cache` bookkeeping achieves (adaround.py:41-54)."""
        self.engine.refresh_initializers(names)
        dead = self._downstream(node)
        for t in [t for t in self.activation_cache if t in dead]:
            del self.activation_cache[t]

    def _downstream(self, node):
        seen, stack = set(), list(node.output)
        while stack:
            t = stack.pop()
            if t in seen:
                continue
            seen.add(t)
            for consumer in self.graph.input_map.get(t, []):
                stack.extend(consumer.output)
        return seen

    def reset(self):
        self.activation_cache.clear()

    def __getitem__(self, name):
        return self.get([name])[name]

    def _positions(self):
        """producer index and last-consumer index of every tensor (network inputs: -1)."""
        if self._pos_graph is not self.graph.model.graph.nodes or self._pos_len != len(self.graph.model.graph.nodes):
            prod, last = {n: -1 for n in self.graph.network_inputs}, {}
            for i, node in enumerate(self.graph.model.graph.nodes):
                for o in node.output:
                    prod[o] = i
                for t in node.input:
                    last[t] = i
            for t in self.graph.network_outputs:
                last[t] = len(self.graph.model.graph.nodes)
            self._prod, self._last = prod, last
            self._pos_graph, self._pos_len = self.graph.model.graph.nodes, len(self.graph.model.graph.nodes)
        return self._prod, self._last

    def get(self, names, keep=True):
        """Tensors for all images of the shard. Besides the requested tensors the cache keeps
        the FRONTIER at the furthest requested node — every activation produced at or before
        it and still consumed after it (for a ResNet: the main path and the identity branch) —
        so that the next request further down the network resumes there instead of at the
        images. Walking the layers in order therefore costs one forward in total, which is
        what the reference's ref-counted NumPy cache achieves on the host (forward_net.py:81-136)."""
        names = list(names)
        out = {}
        missing = []
        for n in names:
            if n in self.graph.initializer:
                out[n] = self.graph.get_initializer(n)
            elif n in self.activation_cache:
                out[n] = self.activation_cache[n]
            else:
                missing.append(n)
        if missing:
            prod, last = self._positions()
            cut = max(prod.get(n, -1) for n in missing)
            frontier = [t for t, pi in prod.items()
                        if pi <= cut < last.get(t, -1) and t not in self.graph.initializer]
            extra = [t for t in frontier if t not in self.activation_cache and t not in missing] if keep else []
            want = missing + extra
            parts = {n: [] for n in want}
            fed = self.engine.inputs_required(want, self.activation_cache.keys())
            for b0 in range(self.st, self.ed, self.batch_size):
                b1 = min(b0 + self.batch_size, self.ed)
                cache = {k: v[b0 - self.st:b1 - self.st] for k, v in self.activation_cache.items()}
                feeds = {nm: self.source.fetch(nm, b0, b1, self.in_shapes[nm]).to(self.device, non_blocking=True)
                         for nm in fed}
                res = self.engine.run(feeds, want=want, cache=cache)
                for n in want:
                    parts[n].append(res[n])
            for n in want:
                t = torch.cat(parts[n], 0) if len(parts[n]) > 1 else parts[n][0]
                if n in missing:
                    out[n] = t
                if keep:
                    self.activation_cache[n] = t
            if keep:   # drop what no node after the cut reads any more
                for t in [t for t in self.activation_cache
                          if last.get(t, -1) <= cut and t not in names and prod.get(t, -1) < cut]:
                    del self.activation_cache[t]
        return out

    def drop(self, names):
        for n in names:
            self.activation_cache.pop(n, None)


def forward_get_tensor(graph, net, index, args, engine=None):
    """All non-Q/DQ node outputs of one image (forward_net.py:467-485). `net` is kept for
    signature compatibility."""
    dev = _device_of(args)
    engine = engine or Engine(graph, dev, _unit_test_cpu=dev.type != "cuda")
    src = as_input_source(args.input_dir)
    feeds = {n: src.fetch(n, index, index + 1, _per_image_shape(graph, n)).to(dev)
             for n in graph.network_inputs}
    return engine.run(feeds, want="all")
