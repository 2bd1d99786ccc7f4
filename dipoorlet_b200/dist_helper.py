"""One process per GPU over torch.distributed (NCCL on NVLink 5 / NVSwitch).

The reference shards calibration images by rank and then combines per-rank results
through JSON files on a shared filesystem (dipoorlet/utils.py:326-345), averaging the
per-rank CLIP VALUES for hist / mse. Here the per-rank STATISTICS are combined on the
device — MIN/MAX of the ranges, integer SUM of the histograms, an all-gather of the
per-image OCTAV values — so the result does not depend on the world size and equals the
reference at world_size = 1 over the same images. Payloads are tiny
(<= 2 MB), so each phase is a single latency-bound collective.

The helpers accept CPU tensors too (gloo), which is how tests/ cover the N > 1 path
without GPUs. Cluster launchers (slurm / mpirun parsing, dipoorlet/dist_helper.py:8-49)
are out of scope: torchrun's environment variables are the only bootstrap.
"""
import os

import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def init_from_env(backend=None):
    """torchrun-style bootstrap (dipoorlet/__main__.py:62-64). Single-process runs do not
    need a process group at all."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank % max(torch.cuda.device_count(), 1))
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank % max(torch.cuda.device_count(), 1))
    return rank, local_rank, world


def get_rank():
    return dist.get_rank() if is_dist() else 0


def get_world_size():
    return dist.get_world_size() if is_dist() else 1


def barrier():
    if is_dist():
        dist.barrier()


def allreduce_minmax(blob_min, blob_max):
    """Phase 1: global activation range, in place. MIN on the minima, MAX on the maxima
    (identical to the reference's own combine for minmax, utils.py:342-344)."""
    if is_dist():
        dist.all_reduce(blob_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(blob_max, op=dist.ReduceOp.MAX)


def allreduce_sum(t):
    """Phase 2: integer SUM of the per-rank histograms (exact, order independent)."""
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def allgather_images(t):
    """Per-image statistics [n_stats, n_local] of every rank, concatenated in rank (=
    image) order -> [n_stats, n_local * world]. Every rank has the same n_local because
    the shard rule is floor division (forward_net.py:207-209)."""
    if not is_dist():
        return t
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, dim=1)


def allreduce_mean_(t):
    """DDP semantics for the rounding parameters' gradients (adaround.py:121)."""
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= dist.get_world_size()
