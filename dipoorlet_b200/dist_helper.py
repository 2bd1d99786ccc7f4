"""One process per GPU over torch.distributed (NCCL on NVLink 5 / NVSwitch).

The reference shards calibration images by rank and then combines per-rank results
through JSON files on a shared filesystem (dipoorlet/utils.py:326-345), averaging the
per-rank CLIP VALUES for hist / mse. Here the per-rank STATISTICS are combined on the
device — MIN/MAX of the ranges, integer SUM of the histograms, an all-gather of the
per-image OCTAV values — so the result does not depend on the world size and equals the
reference at world_size = 1 over the same images. Payloads are tiny
(<= 2 MB), so each phase is a single latency-bound collective.

The helpers accept CPU tensors too (gloo), which is how tests/ cover the N > 1 path
without GPUs. The cluster launchers of the reference (dipoorlet/dist_helper.py:8-49) are reduced to what
they are — translations of the launcher's environment into torchrun's variables (`env_from_slurm`,
`env_from_mpi`), after which `init_from_env` is the only bootstrap.
"""
import os
import re

import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def env_from_mpi(environ):
    """`--mpirun` (dist_helper.py:8-23): Open MPI's rank / size, the master from the HNP URI unless given."""
    out = {"RANK": environ["OMPI_COMM_WORLD_RANK"], "WORLD_SIZE": environ["OMPI_COMM_WORLD_SIZE"],
           "MASTER_PORT": environ.get("MASTER_PORT", "29500")}
    if "MASTER_ADDR" in environ:
        out["MASTER_ADDR"] = environ["MASTER_ADDR"]
    else:
        out["MASTER_ADDR"] = re.search(r'.*tcp://((\d{1,3}\.){3}\d{1,3})[:,].*',
                                       environ["OMPI_MCA_orte_hnp_uri"]).group(1)
    return out


def env_from_slurm(environ):
    """`--slurm` (dist_helper.py:26-49): rank / size from SLURM, port from the job id, master = the first node of
    SLURM_NODELIST, whose name encodes its address after an 8-character prefix ("SH-IDC1-10-5-30-[12,14]" ->
    10.5.30.12) — the reference's cluster convention, kept as is."""
    nodes = environ["SLURM_NODELIST"]
    if "[" in nodes:
        beg = nodes.find("[")
        ends = [p for p in (nodes.find("-", beg), nodes.find(",", beg)) if p >= 0]
        nodes = nodes[:min(ends + [1000])].replace("[", "")
    return {"RANK": str(int(environ["SLURM_PROCID"])), "WORLD_SIZE": str(int(environ["SLURM_NTASKS"])),
            "MASTER_PORT": str(24553 + int(environ["SLURM_JOB_ID"]) % 10000),
            "MASTER_ADDR": nodes[8:].replace("-", ".")}


def init_from_launcher(kind):
    """kind: 'slurm' | 'mpirun'. Ranks map to GPUs round robin (rank % device_count), as in the reference."""
    os.environ.update((env_from_slurm if kind == "slurm" else env_from_mpi)(os.environ))
    n_dev = max(torch.cuda.device_count(), 1) if torch.cuda.is_available() else 1
    os.environ["LOCAL_RANK"] = str(int(os.environ["RANK"]) % n_dev)
    return init_from_env()


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
