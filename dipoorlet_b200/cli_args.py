"""The reference CLI's flags (dipoorlet/__main__.py:23-55), as a parser that can also be
used programmatically: `make_args(model=..., input_dir=..., data_num=..., ...)`.

Superset fixes: `--bins` is parsed as int; `--calib_bs` sets the
images per forward batch of the GPU engine (0 = sized from the blob inventory)."""
import argparse


def build_parser():
    p = argparse.ArgumentParser(prog="dipoorlet_b200")
    p.add_argument("-M", "--model", help="onnx model")
    p.add_argument("-I", "--input_dir", help="calibration data", required=True)
    p.add_argument("-O", "--output_dir", help="output data path")
    p.add_argument("-N", "--data_num", help="num of calibration pics", type=int, required=True)
    p.add_argument("--we", help="weight euqalization", action="store_true")
    p.add_argument("--bc", help="bias correction", action="store_true")
    p.add_argument("--update_bn", help="update BN", action="store_true")
    p.add_argument("--adaround", help="Adaround", action="store_true")
    p.add_argument("--brecq", help="BrecQ", action="store_true")
    p.add_argument("--drop", help="QDrop", action="store_true")
    p.add_argument("-A", "--act_quant", help="algorithm of activation quantization",
                   choices=["minmax", "hist", "mse"], default="mse")
    p.add_argument("-D", "--deploy", help="deploy platform",
                   choices=["trt", "stpu", "magicmind", "rv", "atlas", "snpe", "ti", "imx"],
                   required=True)
    p.add_argument("--bins", help="bins for histogram and kl", default=2048, type=int)
    p.add_argument("--threshold", help="threshold for histogram", default=0.99999, type=float)
    p.add_argument("--savefp", help="Save FP output of model.", action="store_true")
    p.add_argument("--ada_bs", help="Batch size for adaround.", type=int, default=64)
    p.add_argument("--ada_epoch", help="Epoch for adaround.", type=int, default=5000)
    p.add_argument("--skip_layers", help="Skip layer name", default=[], type=str, nargs="+")
    p.add_argument("--stpu_wg", help="Enable winograd for stpu.", action="store_true")
    p.add_argument("--skip_prof_layer", help="Skip profiling by layer.", default=False, action="store_true")
    p.add_argument("--slurm", help="Launch task from slurm", default=False, action="store_true")
    p.add_argument("--mpirun", help="Launch task from mpirun", default=False, action="store_true")
    p.add_argument("--sparse", help="Sparse on/off", default=False, action="store_true")
    p.add_argument("--sparse_rate", help="Sparse rate", type=float, default=0.5)
    p.add_argument("--pattern", help="Sparse pattern", choices=["unstruction", "nv24"], default="unstruction")
    p.add_argument("--optim_transformer", help="Transformer model optimization", default=False, action="store_true")
    p.add_argument("--model_type", help="Transformer model type", choices=["unet"], default=None)
    p.add_argument("--quant_format", default="QDQ", type=str, choices=["QOP", "QDQ"])
    p.add_argument("--calib_bs", help="images per GPU forward batch (0 = auto)", type=int, default=0)
    p.add_argument("--resident", help="keep pass-1 blobs in HBM for the histogram pass (default: auto, "
                   "when they fit)", type=int, default=None, choices=[0, 1])
    return p


def make_args(**kw):
    """Namespace with the CLI defaults, overridden by keyword (input_dir may be a path or
    a forward_net.ArrayInput)."""
    p = build_parser()
    ns = argparse.Namespace(**{a.dest: a.default for a in p._actions if a.dest != "help"})
    ns.rank, ns.local_rank, ns.world_size = 0, 0, 1
    ns.acti_quant = False
    for k, v in kw.items():
        setattr(ns, k, v)
    return ns
