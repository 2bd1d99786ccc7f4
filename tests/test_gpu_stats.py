"""GPU parity of K1 (segment stats), K2 (histogram), K3 (percentile) and K4 (OCTAV)
against the NumPy oracle on identical blobs, through the C-ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _blobs(rng, n_img, shapes, relu_every=2, scale_growth=1.7):
    """Synthetic activation blobs: N(0, sigma_b) with ReLU on alternating blobs."""
    blobs = {}
    for b, shp in enumerate(shapes):
        sigma = scale_growth ** (b % 5)
        imgs = []
        for _ in range(n_img):
            x = (rng.standard_normal(shp) * sigma).astype(np.float32)
            if b % relu_every == 1:
                x = np.maximum(x, 0)
            imgs.append(x)
        blobs[f"blob{b}"] = imgs
    return blobs


SHAPES = [(3, 32, 32), (64, 28, 28), (1000,), (7,), (256, 14, 14), (2048, 1, 1), (33, 5, 5),
          (64, 112, 112)]


def _to_batch(blobs, torch, K):
    tensors = [torch.from_numpy(np.stack(v)).cuda().contiguous() for v in blobs.values()]
    return K.BlobBatch(tensors)


def _run_segstats(batch, torch, K):
    n = batch.n_segments
    dev = batch.device
    smin = torch.empty(n, dtype=torch.float32, device=dev)
    smax = torch.empty(n, dtype=torch.float32, device=dev)
    ssum = torch.empty(n, dtype=torch.float64, device=dev)
    snnz = torch.empty(n, dtype=torch.int64, device=dev)
    bmin = torch.full((batch.n_blobs,), float("inf"), dtype=torch.float32, device=dev)
    bmax = torch.full((batch.n_blobs,), float("-inf"), dtype=torch.float32, device=dev)
    K.segstats(batch, smin, smax, ssum, snnz, bmin, bmax)
    return smin, smax, ssum, snnz, bmin, bmax


def test_segstats_matches_oracle(dpl_built):
    import torch
    from dipoorlet_b200 import kernels as K
    from oracle import stats as O
    rng = np.random.default_rng(1)
    blobs = _blobs(rng, 5, SHAPES)
    batch = _to_batch(blobs, torch, K)
    smin, smax, ssum, snnz, bmin, bmax = _run_segstats(batch, torch, K)
    torch.cuda.synchronize()
    mm = O.minmax_stats(blobs)
    for (name, imgs), (off, nseg) in zip(blobs.items(), batch.seg_slices()):
        got_min = smin[off:off + nseg].cpu().numpy()
        got_max = smax[off:off + nseg].cpu().numpy()
        assert np.array_equal(got_min, np.array(mm[name]["min"], dtype=np.float32)), name
        assert np.array_equal(got_max, np.array(mm[name]["max"], dtype=np.float32)), name
        ref_sum = np.array([np.abs(x).astype(np.float64).sum() for x in imgs])
        ref_nnz = np.array([(np.abs(x) > 0).sum() for x in imgs])
        assert np.allclose(ssum[off:off + nseg].cpu().numpy(), ref_sum, rtol=2e-6)
        assert np.array_equal(snnz[off:off + nseg].cpu().numpy(), ref_nnz)
    clip = O.clip_minmax(mm)
    for i, name in enumerate(blobs):
        assert bmin[i].item() == clip[name][0] and bmax[i].item() == clip[name][1]


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 7])
@pytest.mark.parametrize("bins", [2048, 128, 1000])
def test_hist_bit_exact(dpl_built, variant, bins):
    import torch
    from dipoorlet_b200 import kernels as K
    from oracle import stats as O
    if variant == 3 and bins > 2048:
        pytest.skip("TMA variant supports <= 2048 bins")
    rng = np.random.default_rng(2)
    blobs = _blobs(rng, 4, SHAPES)
    blobs["zeros"] = [np.zeros((8, 8), np.float32) for _ in range(4)]          # data_max == 0
    blobs["const"] = [np.full((300,), 2.5, np.float32) for _ in range(4)]      # all mass on the right edge
    blobs["odd"] = [rng.standard_normal(1001).astype(np.float32) for _ in range(4)]  # misaligned segments
    batch = _to_batch(blobs, torch, K)
    smin, smax, ssum, snnz, bmin, bmax = _run_segstats(batch, torch, K)
    dm = torch.empty(batch.n_blobs, dtype=torch.float32, device=batch.device)
    K.absmax(bmin, bmax, dm)
    counts = torch.zeros((batch.n_blobs, bins), dtype=torch.int64, device=batch.device)
    K.hist_abs(batch, dm, counts, bins, variant=variant)
    K.hist_abs(batch, dm, counts, bins, variant=variant)  # accumulates: second "batch"
    torch.cuda.synchronize()
    mm = O.minmax_stats(blobs)
    ref = O.hist_stats(blobs, mm, bins)
    got = counts.cpu().numpy()
    for i, name in enumerate(blobs):
        want = 2 * np.stack(ref[name]).sum(0)
        assert dm[i].item() == O.data_max_of(mm[name]), name
        assert np.array_equal(got[i], want), (name, np.nonzero(got[i] != want)[0][:8])


def test_hist_edge_values_bit_exact(dpl_built):
    """Elements sitting exactly on (and one ulp around) float32 bin edges."""
    import torch
    from dipoorlet_b200 import kernels as K
    from oracle import stats as O
    rng = np.random.default_rng(3)
    bins = 2048
    for dm in [1.0, 3.3721, 1e-3, 6.25e4, 0.7]:
        dm = np.float32(dm)
        edges = np.linspace(0, dm, bins + 1, dtype=np.float32)
        pts = np.concatenate([edges, np.nextafter(edges, np.float32(0)), np.nextafter(edges, dm),
                              rng.uniform(0, dm, 20000).astype(np.float32)])
        pts = np.clip(pts, 0, dm).astype(np.float32)
        pts[0] = dm
        sign = np.where(rng.random(pts.size) < 0.5, -1, 1).astype(np.float32)
        blobs = {"e": [pts * sign]}
        batch = _to_batch(blobs, torch, K)
        smin, smax, ssum, snnz, bmin, bmax = _run_segstats(batch, torch, K)
        dmt = torch.empty(1, dtype=torch.float32, device=batch.device)
        K.absmax(bmin, bmax, dmt)
        for variant in (1, 2, 3, 4, 5, 7):
            counts = torch.zeros((1, bins), dtype=torch.int64, device=batch.device)
            K.hist_abs(batch, dmt, counts, bins, variant=variant)
            want = np.histogram(np.abs(blobs["e"][0]), bins, (0, dm))[0]
            assert np.array_equal(counts.cpu().numpy()[0], want), (float(dm), variant)


def test_percentile_matches_oracle(dpl_built):
    import torch
    from dipoorlet_b200 import kernels as K
    from oracle import stats as O
    rng = np.random.default_rng(4)
    bins = 2048
    blobs = _blobs(rng, 6, SHAPES)
    blobs["zeros"] = [np.zeros((8, 8), np.float32) for _ in range(6)]
    batch = _to_batch(blobs, torch, K)
    smin, smax, ssum, snnz, bmin, bmax = _run_segstats(batch, torch, K)
    dm = torch.empty(batch.n_blobs, dtype=torch.float32, device=batch.device)
    K.absmax(bmin, bmax, dm)
    counts = torch.zeros((batch.n_blobs, bins), dtype=torch.int64, device=batch.device)
    K.hist_abs(batch, dm, counts, bins)
    mm = O.minmax_stats(blobs)
    ref_h = O.hist_stats(blobs, mm, bins)
    for thr in (0.99999, 0.999, 0.5, 1.0, 1.5):
        clip = torch.empty((batch.n_blobs, 2), dtype=torch.float32, device=batch.device)
        sel = torch.empty(batch.n_blobs, dtype=torch.int32, device=batch.device)
        K.hist_percentile(counts, bins, thr, dm, bmin, bmax, clip, sel)
        ref_c, ref_sel = O.clip_hist(mm, ref_h, bins, thr, return_bins=True)
        got_c, got_sel = clip.cpu().numpy(), sel.cpu().numpy()
        for i, name in enumerate(blobs):
            assert got_sel[i] == ref_sel[name], (name, thr)
            assert got_c[i, 0] == np.float32(ref_c[name][0]), (name, thr)
            assert got_c[i, 1] == np.float32(ref_c[name][1]), (name, thr)


def test_octav_matches_oracle(dpl_built):
    import torch
    from dipoorlet_b200 import kernels as K
    from oracle import stats as O
    rng = np.random.default_rng(5)
    blobs = _blobs(rng, 3, SHAPES)
    blobs["odd"] = [rng.standard_normal(1001).astype(np.float32) for _ in range(3)]
    batch = _to_batch(blobs, torch, K)
    smin, smax, ssum, snnz, bmin, bmax = _run_segstats(batch, torch, K)
    s = torch.empty(batch.n_segments, dtype=torch.float32, device=batch.device)
    iters = torch.empty(batch.n_segments, dtype=torch.int32, device=batch.device)
    K.octav(batch, ssum, snnz, 1 / (4 ** 8) / 3 / 1, s, iters)
    torch.cuda.synchronize()
    ref = O.octav_stats(blobs)
    got = s.cpu().numpy()
    for (name, _), (off, nseg) in zip(blobs.items(), batch.seg_slices()):
        want = np.array(ref[name]["optimal_s"], dtype=np.float32)
        assert np.allclose(got[off:off + nseg], want, rtol=1e-5, atol=0), (name, got[off:off + nseg], want)
