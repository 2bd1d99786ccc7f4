"""The learnable-rounding layer on the GPU (dipoorlet/weight_transform/ada_quant_layer.py).

The reference builds a torch.nn module per layer and lets autograd + torch.optim.Adam drive
dozens of small kernels per iteration. Here a layer is a bundle of device buffers and each
iteration is a fixed sequence of launches with no autograd graph:

    K6 weight   W_soft = clamp(floor(W/s) + h(alpha), qmin, qmax) * s        (libdpl_b200)
    conv / gemm forward                                                      (cuDNN/cuBLAS stand-in)
    K6 epilogue relu [+ drop-fakequant], or fused L2 loss + dL/do at the block end (libdpl_b200)
    conv / gemm weight-gradient (+ data-gradient inside a block)            (cuDNN/cuBLAS stand-in)
    K6 step     dL/dalpha (+ regulariser) and Adam on alpha, fused            (libdpl_b200)

STAND-IN NOTICE: the dense contraction is issued through torch.ops.aten.convolution /
convolution_backward (true fp32); a tcgen05 implicit-GEMM tile is the planned replacement
("what comes next").
"""
import os

import numpy as np
import torch

from .. import kernels as K

__all__ = ["adaround_reg", "AdaQLayer", "TempDecay", "ZETA", "GAMMA"]

ZETA, GAMMA = 1.1, -0.1


class TempDecay:
    """beta(t): 0 during the first 20 % (regulariser off), then cosine 20 -> 2
    (ada_quant_layer.py:117-130)."""

    def __init__(self, t_max, rel_start_decay=0.2, start_b=20, end_b=2):
        self.t_max = t_max
        self.start_decay = rel_start_decay * t_max
        self.start_b = start_b
        self.end_b = end_b

    def __call__(self, t):
        if t < self.start_decay:
            return 0.0
        rel_t = (t - self.start_decay) / (self.t_max - self.start_decay)
        return self.end_b + 0.5 * (self.start_b - self.end_b) * (1 + np.cos(rel_t * np.pi))


class adaround_reg:
    """Holds the regulariser's schedule; the value and its gradient are computed inside the
    fused step kernel (ada_quant_layer.py:96-110)."""

    def __init__(self, max_iter=10000, zeta=ZETA, gamma=GAMMA, alpha=0.01, beta=20):
        self.zeta, self.gamma, self.alpha, self.beta = zeta, gamma, alpha, beta
        self.temp_anneal = TempDecay(max_iter)

    def rectified_sigmoid(self, round_mask):
        return ((self.zeta - self.gamma) * torch.sigmoid(round_mask) + self.gamma).clamp(0, 1)

    def update(self, it):
        self.beta = self.temp_anneal(it)
        return self.beta


class AdaQLayer:
    """One Conv / Gemm / ConvTranspose with learnable rounding of its weight."""

    def __init__(self, node, weight, bias, qw_scale, q_min, q_max, relu_flag, qi=None,
                 acti_quant=False, drop_ratio=0.5, device=None):
        dev = device or torch.device("cuda")
        self.node = node
        self.type = node.op_type
        a = node.attrs
        self.weight = torch.as_tensor(weight, dtype=torch.float32, device=dev).contiguous()
        if self.type == 'ConvTranspose':   # per-channel axis is dim 1 of an [in, out/g, k, k] kernel
            self.weight = self.weight.transpose(0, 1).contiguous()
        self.bias = None if bias is None else torch.as_tensor(bias, dtype=torch.float32, device=dev)
        self.scale = qw_scale.reshape(-1).contiguous()
        self.q_min, self.q_max = float(q_min), float(q_max)
        self.relu_flag = bool(relu_flag)
        self.qi = qi                       # (scale, qmin, qmax) of the output activation or None
        self.acti_quant = bool(acti_quant) and qi is not None
        self.drop_ratio = drop_ratio
        nd = self.weight.dim() - 2
        self.stride = list(a.get("strides", [1] * nd))
        self.padding = list(a.get("pads", [0] * (2 * nd))[:nd])
        self.dilation = list(a.get("dilations", [1] * nd))
        self.groups = int(a.get("group", 1))
        self.output_padding = list(a.get("output_padding", [0] * nd))
        # alpha0 = -log((zeta - gamma) / (rest - gamma) - 1)  =>  h(alpha0) = rest
        self.round_mask, self.wfloor = K.adaround_init(self.weight, self.scale)
        self.m = torch.zeros_like(self.round_mask)
        self.v = torch.zeros_like(self.round_mask)
        self.w_soft = torch.empty_like(self.weight)
        self.saved = None

    # ---- weights ----------------------------------------------------------------
    def quant_weight(self, soft=True):
        K.adaround_weight(self.wfloor, self.round_mask, self.scale, self.q_min, self.q_max, soft,
                          out=self.w_soft)
        return self.w_soft

    def hard_weight(self):
        """floor(W/s) + (alpha >= 0), clamped, * s — in the graph's own layout."""
        w = K.adaround_weight(self.wfloor, self.round_mask, self.scale, self.q_min, self.q_max, False)
        if self.type == 'ConvTranspose':
            w = w.transpose(0, 1).contiguous()
        return w

    def _w_for_op(self, w):
        return w.transpose(0, 1) if self.type == 'ConvTranspose' else w

    # ---- forward / backward of the dense op (stand-in: cuDNN / cuBLAS, true fp32) --------
    def _tensor_core_ok(self, x):
        """tcgen05 TF32 tile (libdpl_b200 dpl_gemm_tf32) for the shapes TMA can address: 1x1
        stride-1 ungrouped convolutions with C_in and H*W multiples of 4, and Gemm layers with
        K a multiple of 4. TF32 is what the reference's torch conv uses by default
        (cudnn.allow_tf32 = True). DPL_TCGEN05=0 keeps everything on the fp32 library path."""
        if os.environ.get("DPL_TCGEN05", "1") == "0":
            return False
        if getattr(self, "_tc_disabled", False):
            return False
        if self.type == 'Gemm':   # K and the output width are leading dimensions of TMA operands
            return (x.dim() == 2 and x.shape[1] % 4 == 0 and self.weight.shape[0] % 4 == 0
                    and x.is_contiguous())
        if self.type != 'Conv' or x.dim() != 4 or not x.is_contiguous():
            return False
        k = self.weight.shape[2:]
        return (list(k) == [1, 1] and self.stride == [1, 1] and self.padding == [0, 0]
                and self.dilation == [1, 1] and self.groups == 1 and x.shape[1] % 4 == 0
                and (x.shape[2] * x.shape[3]) % 4 == 0)

    def dense_forward(self, x, w):
        self._tc = self._tensor_core_ok(x)
        if self._tc:
            try:
                if self.type == 'Gemm':
                    return K.linear_forward(x, w, self.bias)
                return K.conv1x1_forward(x, w.view(w.shape[0], w.shape[1]), self.bias)
            except K.GemmUnsupported:     # e.g. a mis-aligned slice: keep this layer on the library path
                self._tc = False
                self._tc_disabled = True
        if self.type == 'Gemm':
            return torch.nn.functional.linear(x, w, self.bias)
        return torch.ops.aten.convolution(x, self._w_for_op(w), self.bias, self.stride, self.padding,
                                          self.dilation, self.type == 'ConvTranspose',
                                          self.output_padding, self.groups)

    def dense_backward(self, x, w, go, need_dx):
        if getattr(self, "_tc", False):
            try:
                if self.type == 'Gemm':
                    return (K.linear_dgrad(go, w) if need_dx else None), K.linear_wgrad(go, x)
                w2 = w.view(w.shape[0], w.shape[1])
                gw = K.conv1x1_wgrad(go, x).view_as(w)
                return (K.conv1x1_dgrad(go, w2) if need_dx else None), gw
            except K.GemmUnsupported:
                self._tc = False
                self._tc_disabled = True
        if self.type == 'Gemm':
            gw = go.t() @ x
            gx = go @ w if need_dx else None
            return gx, gw
        wt = self._w_for_op(w)
        gx, gw, _ = torch.ops.aten.convolution_backward(
            go, x, wt, None, self.stride, self.padding, self.dilation, self.type == 'ConvTranspose',
            self.output_padding, self.groups, [bool(need_dx), True, False])
        if self.type == 'ConvTranspose':
            gw = gw.transpose(0, 1)
        return gx, gw.contiguous()

    def act_cfg(self, seed):
        quant = self.qi if self.acti_quant else None
        prob = self.drop_ratio if self.acti_quant else 1.0
        return dict(relu=self.relu_flag, quant=quant, prob=prob, seed=seed)
