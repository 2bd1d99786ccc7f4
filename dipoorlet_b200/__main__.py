"""`python -m dipoorlet_b200 -M model.onnx -I calib_dir -N 1024 -A hist -D trt ...` —
the reference CLI (dipoorlet/__main__.py:23-161) on the B200 path: same flags, same files in
the output directory. Launch one process per GPU with torchrun for multi-GPU runs."""
import copy
import os
import sys
import time

from . import dist_helper
from .cli_args import build_parser
from .deploy import to_deploy
from .graph import load_graph
from .profiling import (quantize_profiling_multipass, show_model_profiling_res, show_model_ranges,
                        weight_need_perchannel)
from .tensor_cali import tensor_calibration
from .utils import load_clip_val, logger, save_clip_val, save_profiling_res, setup_logger
from .weight_transform import weight_calibration


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.optim_transformer or args.model_type is not None or args.quant_format == "QOP":
        sys.exit("transformer / QOperator paths are outside the B200 hot path (CNN QDQ calibration only)")
    if args.slurm or args.mpirun:
        rank, local_rank, world = dist_helper.init_from_launcher("slurm" if args.slurm else "mpirun")
    else:
        rank, local_rank, world = dist_helper.init_from_env()
    if args.output_dir is None:
        args.output_dir = os.path.join(os.path.abspath(os.path.dirname(args.model)), 'results')
    if rank == 0:
        os.makedirs(args.output_dir, exist_ok=True)
        setup_logger(args)
    dist_helper.barrier()
    logger.parent = None
    start = time.time()
    onnx_graph = load_graph(args.model, args.output_dir, args.deploy, args.model_type)
    args.rank, args.local_rank, args.world_size = rank, local_rank, world
    args.acti_quant = False

    if rank == 0:
        logger.info("Do tensor calibration...")
    act_clip_val, weight_clip_val = tensor_calibration(onnx_graph, args)
    # statistics were combined on the device, so every rank holds the final values: rank 0
    # writes the files the reference writes (per-rank file kept for tools that look for it)
    if rank == 0:
        save_clip_val(copy.deepcopy(act_clip_val), copy.deepcopy(weight_clip_val), args,
                      act_fname='act_clip_val.json.rank0', weight_fname='weight_clip_val.json.rank0')
        save_clip_val(act_clip_val, weight_clip_val, args)
    dist_helper.barrier()
    act_clip_val, weight_clip_val = load_clip_val(args)

    if rank == 0:
        logger.info("Weight transform...")
    graph, graph_ori, act_clip_val, weight_clip_val = weight_calibration(onnx_graph, act_clip_val,
                                                                         weight_clip_val, args)
    dist_helper.barrier()

    if rank == 0:
        logger.info("Profiling...")
    layer_cos, model_cos, quant_node_list = quantize_profiling_multipass(
        graph, graph_ori, copy.deepcopy(act_clip_val), copy.deepcopy(weight_clip_val), args)
    if rank == 0:
        save_profiling_res(dict(layer_cos), {k: list(v) for k, v in model_cos.items()}, args, rank=0)
        show_model_profiling_res(graph, layer_cos, model_cos, quant_node_list, args)
        show_model_ranges(graph, act_clip_val, weight_clip_val, args)
        weight_need_perchannel(graph, args)
        logger.info("Deploy to " + args.deploy + '...')
        to_deploy(graph, act_clip_val, weight_clip_val, args)
        logger.info("Total time cost: {} seconds.".format(int(time.time() - start)))
    dist_helper.barrier()


if __name__ == "__main__":
    main()
