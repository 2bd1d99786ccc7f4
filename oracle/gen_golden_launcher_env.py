"""Golden fixture for the cluster bootstraps: the REFERENCE's own init_from_slurm / init_from_mpi
(/root/reference/dipoorlet/dist_helper.py, unmodified; torch.distributed.init_process_group and torch.cuda
replaced by no-ops) on a few launcher environments -> tests/golden/launcher_env.json, the RANK / WORLD_SIZE /
MASTER_ADDR / MASTER_PORT they export.

    python oracle/gen_golden_launcher_env.py        # build container only; the fixture is committed
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

CASES = {
    "slurm": [
        {"SLURM_JOB_ID": "123456", "SLURM_PROCID": "3", "SLURM_NTASKS": "8", "SLURM_NODELIST": "SH-IDC1-10-5-30-[12,14]"},
        {"SLURM_JOB_ID": "7", "SLURM_PROCID": "0", "SLURM_NTASKS": "1", "SLURM_NODELIST": "SH-IDC1-10-5-30-12"},
        {"SLURM_JOB_ID": "99999", "SLURM_PROCID": "15", "SLURM_NTASKS": "16", "SLURM_NODELIST": "SH-IDC1-10-5-30-[12-15]"},
        {"SLURM_JOB_ID": "20000", "SLURM_PROCID": "1", "SLURM_NTASKS": "2", "SLURM_NODELIST": "BJ-IDC2-10-198-4-[7-9,11]"},
    ],
    "mpi": [
        {"OMPI_COMM_WORLD_RANK": "1", "OMPI_COMM_WORLD_SIZE": "4", "OMPI_MCA_orte_hnp_uri": "12345.0;tcp://10.1.2.3,192.168.0.1:4567"},
        {"OMPI_COMM_WORLD_RANK": "0", "OMPI_COMM_WORLD_SIZE": "8", "OMPI_MCA_orte_hnp_uri": "777.0;tcp://172.16.5.9:40000"},
        {"OMPI_COMM_WORLD_RANK": "2", "OMPI_COMM_WORLD_SIZE": "3", "MASTER_ADDR": "10.0.0.1", "MASTER_PORT": "31000"},
    ],
}
KEYS = ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")


def main():
    dist.init_process_group = lambda **kw: None
    dist.get_rank = lambda: int(os.environ["RANK"])
    torch.cuda.device_count = lambda: 8
    torch.cuda.set_device = lambda d: None
    from dipoorlet import dist_helper as R
    out = {}
    for kind, fn in (("slurm", R.init_from_slurm), ("mpi", R.init_from_mpi)):
        out[kind] = []
        for env in CASES[kind]:
            for k in list(os.environ):
                if k in KEYS or k.startswith("SLURM_") or k.startswith("OMPI_"):
                    del os.environ[k]
            os.environ.update(env)
            fn()
            out[kind].append({"env": env, "exports": {k: os.environ[k] for k in KEYS}})
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "launcher_env.json"), "w"), indent=1)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
