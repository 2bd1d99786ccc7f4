"""Activation statistics and clip search, restated from the reference with NumPy.

Blobs are passed explicitly (name -> list of per-image float32 arrays) because the
reference's ONNXRuntime forward (forward_net.py:200,216) is third-party and is restated
separately in oracle/forward.py.
"""
import numpy as np


def minmax_stats(blobs):
    """forward_get_minmax, dipoorlet/forward_net.py:220-235: per image per blob
    x.max() / x.min() appended to lists."""
    stats = {}
    for name, imgs in blobs.items():
        stats[name] = {"max": [x.max() for x in imgs], "min": [x.min() for x in imgs]}
    return stats


def clip_minmax(stats):
    """find_clip_val_minmax, dipoorlet/tensor_cali/basic_algorithm.py:19-22."""
    return {name: [np.min(t["min"]), np.max(t["max"])] for name, t in stats.items()}


def data_max_of(stats_entry):
    """forward_net.py:266-267: max(np.max(maxs), -np.min(mins)) (np.float32)."""
    return max(np.max(stats_entry["max"]), -np.min(stats_entry["min"]))


def hist_stats(blobs, stats_min_max, bins):
    """forward_get_hist, dipoorlet/forward_net.py:265-280: per image
    np.histogram(np.abs(x), int(bins), (0, data_max))."""
    out = {}
    for name, imgs in blobs.items():
        dm = data_max_of(stats_min_max[name])
        out[name] = [np.histogram(np.abs(x), int(bins), (0, dm))[0] for x in imgs]
    return out


def clip_hist(stats_min_max, act_stats_hist, bins, threshold, return_bins=False):
    """find_clip_val_hist, dipoorlet/tensor_cali/basic_algorithm.py:37-53.
    act_stats_hist: name -> list of per-image histograms (summed here as at :38)."""
    clip_val, sel = {}, {}
    for name, hist in act_stats_hist.items():
        hist = np.stack(hist).sum(0)
        hist = hist.astype(np.float32) / hist.sum()
        data_max = max(-np.min(stats_min_max[name]["min"]), np.max(stats_min_max[name]["max"]))
        accum = 0
        sel[name] = -1
        for i in range(len(hist)):
            accum += hist[i]
            if accum >= threshold:
                clip_value = (i + 0.5) * (data_max / bins)
                clip_val[name] = [max(-clip_value, np.min(stats_min_max[name]["min"])),
                                  min(clip_value, np.max(stats_min_max[name]["max"]))]
                sel[name] = i
                break
        if name not in clip_val:
            clip_val[name] = [np.min(stats_min_max[name]["min"]),
                              np.max(stats_min_max[name]["max"])]
    return (clip_val, sel) if return_bins else clip_val


def octav_stats(blobs, unsigned_of=None):
    """forward_net_octav, dipoorlet/forward_net.py:315-340. `unsigned` is 1 for every
    platform without 'dynamic_sym' in qi_params (trt: platform_settings.py:14-18)."""
    stats = {}
    for name, imgs in blobs.items():
        entry = {"optimal_s": [], "min": [], "max": []}
        for x in imgs:
            data_max = x.max()
            data_min = x.min()
            unsigned = unsigned_of(data_min) if unsigned_of else 1
            abs_x = np.abs(x)
            with np.errstate(invalid="ignore", divide="ignore"):
                s_n = abs_x.sum() / abs_x[abs_x > 0].size
                for _ in range(20):
                    s_n_plus_1 = abs_x[abs_x > s_n].sum() / \
                        (1 / (4 ** 8) / 3 / unsigned * abs_x[abs_x <= s_n].size + abs_x[abs_x > s_n].size)
                    if np.abs(s_n_plus_1 - s_n) < 1e-6:
                        break
                    s_n = s_n_plus_1
            entry["optimal_s"].append(s_n)
            entry["min"].append(data_min)
            entry["max"].append(data_max)
        stats[name] = entry
    return stats


def clip_octav(optimal_s):
    """find_clip_val_octav, dipoorlet/tensor_cali/basic_algorithm.py:63-69."""
    clip_val = {}
    for k, v in optimal_s.items():
        data_max = np.array(v["max"]).max()
        data_min = np.array(v["min"]).min()
        clip_val[k] = [max(data_min, -np.array(v["optimal_s"]).mean()),
                       min(data_max, np.array(v["optimal_s"]).mean())]
    return clip_val


def weight_minmax(weights, transpose_names=()):
    """find_clip_val_minmax_weight, dipoorlet/tensor_cali/basic_algorithm.py:81-91.
    weights: name -> ndarray (inputs[1:] of every LAYER_HAS_WEIGHT node)."""
    out = {}
    for name, tensor in weights.items():
        if len(tensor.shape) < 1:
            continue
        if name in transpose_names:
            tensor = tensor.transpose([1, 0, 2, 3])
        c_num = tensor.shape[0]
        out[name] = [np.min(tensor.reshape((c_num, -1)), -1), np.max(tensor.reshape((c_num, -1)), -1)]
    return out


def trt_blob_range(act_clip_val):
    """gen_trt_range, dipoorlet/deploy/deploy_trt.py:9-11 after the JSON round trip of
    save_clip_val/load_clip_val (dipoorlet/utils.py:314-316,353-356):
    np.float32 -> .tolist() -> json -> np.float64; value = max(-lo, hi) as float."""
    out = {}
    for k, (lo, hi) in act_clip_val.items():
        lo = np.float64(np.asarray(lo).tolist())
        hi = np.float64(np.asarray(hi).tolist())
        out[k] = max(-lo.astype(float), hi.astype(float))
    return out
