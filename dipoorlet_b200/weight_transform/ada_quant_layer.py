"""The learnable-rounding layer on the GPU (dipoorlet/weight_transform/ada_quant_layer.py).

The reference builds a torch.nn module per layer and lets autograd + torch.optim.Adam drive
dozens of small kernels per iteration. Here a layer is a bundle of device buffers and each
iteration is a fixed sequence of launches with no autograd graph:

    K6 weight   W_soft = clamp(floor(W/s) + h(alpha), qmin, qmax) * s        (libdpl_b200)
    conv / gemm forward                                      (libdpl_b200: tcgen05 TF32 tiles, see _classify)
    K6 epilogue relu [+ drop-fakequant], or fused L2 loss + dL/do at the block end (libdpl_b200)
    conv / gemm weight-gradient (+ data-gradient inside a block)                  (libdpl_b200)
    K6 step     dL/dalpha (+ regulariser) and Adam on alpha, fused            (libdpl_b200)

Every layer of both model families of BASELINE.json (ResNet-50, MobileNetV2) runs its contraction and both
gradients on libdpl_b200 kernels; torch / cuDNN remains only for layer types neither family contains
(ConvTranspose, grouped or dilated convolutions), announced by a warning.
"""
import os

import numpy as np
import torch

from .. import kernels as K

__all__ = ["adaround_reg", "AdaQLayer", "TempDecay", "ZETA", "GAMMA", "classify_layer"]

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


def classify_layer(op_type, wshape, attrs, xshape):
    """Which libdpl_b200 contraction serves this layer (ada_quant_layer.py:224-244 is F.conv2d / F.linear /
    F.conv_transpose2d on cuDNN / cuBLAS):
      'gemm'  Gemm on the tcgen05 TF32 tile                         (dpl_gemm_tf32)
      'c1x1'  1x1 stride-1 convolution straight from NCHW           (dpl_gemm_tf32)
      'taps'  3x3 (stride 1 / 2) and the other 1x1 convolutions on the tap-table TF32 kernels
              (dpl_tap_conv_tf32 / dpl_tap_wgrad_tf32) over channel-last staging copies
      'dw'    depthwise 3x3 / 5x5, exact fp32 on the FMA pipe        (dpl_dwconv2d_*)
      'stem'  few input channels (3-channel stem): direct fp32 forward, im2col + tcgen05 weight gradient
      'lib'   anything else (ConvTranspose, grouped, dilated, ... - in neither model family of
              BASELINE.json): torch / cuDNN, announced once; DPL_STRICT_NATIVE=1 raises instead.
    TF32 is what the reference's torch conv uses by default (cudnn.allow_tf32 = True).
    DPL_TCGEN05=0 keeps everything on the library path (debugging only)."""
    if os.environ.get("DPL_TCGEN05", "1") == "0":
        return 'lib'
    if op_type == 'Gemm':
        ok = len(xshape) == 2 and xshape[1] % 4 == 0 and wshape[0] % 4 == 0
        return 'gemm' if ok else 'lib'
    nd = len(wshape) - 2
    dilation = list(attrs.get("dilations", [1] * nd))
    if op_type != 'Conv' or len(xshape) != 4 or dilation != [1, 1]:
        return 'lib'
    co, cig, kh, kw = wshape
    ci = xshape[1]
    stride = list(attrs.get("strides", [1, 1]))
    pads = list(attrs.get("pads", [0, 0, 0, 0]))
    groups = int(attrs.get("group", 1))
    if kh != kw or stride[0] != stride[1] or pads[0] != pads[1] or pads[:2] != pads[2:]:
        return 'lib'
    k, st, pd = kh, stride[0], pads[0]
    if groups == 1:
        if k == 1 and st == 1 and pd == 0 and ci % 4 == 0 and (xshape[2] * xshape[3]) % 4 == 0:
            return 'c1x1'
        if ((k, pd) in ((3, 1), (1, 0))) and st in (1, 2) and ci % 4 == 0 and co % 4 == 0:
            return 'taps'
        if ci * k * k <= 256 and (k, st) in ((7, 2), (3, 2), (3, 1), (5, 2), (5, 1)):
            return 'stem'
        return 'lib'
    if groups == ci == co and cig == 1 and k in (3, 5):
        return 'dw'
    return 'lib'


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

    # ---- forward / backward of the dense op: libdpl_b200 only for both model families ---------------
    def _classify(self, x):
        """Which libdpl_b200 contraction serves this layer for an input of x's shape (see classify_layer)."""
        if not x.is_contiguous():
            return 'lib'
        return classify_layer(self.type, tuple(self.weight.shape), self.node.attrs, tuple(x.shape))

    def _lib_notice(self):
        if os.environ.get("DPL_STRICT_NATIVE", "0") == "1":
            raise RuntimeError("layer %s (%s) has no libdpl_b200 contraction" % (self.node.name, self.type))
        if not getattr(AdaQLayer, "_lib_warned", False):
            AdaQLayer._lib_warned = True
            from ..utils import logger
            logger.warning("%s: layer shape outside the libdpl_b200 contraction kernels, using torch / cuDNN"
                           % self.node.name)

    def dense_forward(self, x, w):
        kind = self._classify(x) if not getattr(self, "_tc_disabled", False) else 'lib'
        try:
            if kind == 'gemm':
                y = K.linear_forward(x, w, self.bias)
            elif kind == 'c1x1':
                y = K.conv1x1_forward(x, w.view(w.shape[0], w.shape[1]), self.bias)
            elif kind == 'taps':
                plan = self._plan(x)
                self._wf, self._wd = K.taps_layout(w, True, True, getattr(self, "_wf", None),
                                                   getattr(self, "_wd", None))
                self._xp = K.recon_stage_input(x, plan, getattr(self, "_xp", None))
                y = K.recon_conv_forward(self._xp, plan, self._wf, self.bias)
            elif kind == 'dw':
                y = K.dwconv2d_forward(x, w, self.bias, self.stride[0], self.padding[0])
            elif kind == 'stem':
                y = K.conv_direct_forward(x, w, self.bias, self.stride[0], self.padding[0])
        except K.GemmUnsupported:     # e.g. a mis-aligned slice: keep this layer on the library path
            kind = 'lib'
            self._tc_disabled = True
        self._kind = kind
        if kind != 'lib':
            return y
        self._lib_notice()
        if self.type == 'Gemm':
            return torch.nn.functional.linear(x, w, self.bias)
        return torch.ops.aten.convolution(x, self._w_for_op(w), self.bias, self.stride, self.padding,
                                          self.dilation, self.type == 'ConvTranspose',
                                          self.output_padding, self.groups)

    def _plan(self, x):
        key = tuple(x.shape)
        plan = getattr(self, "_plan_cache", {}).get(key)
        if plan is None:
            plan = K.ReconConvPlan(x.shape[0], x.shape[2], x.shape[3], self.weight.shape[2], self.stride[0],
                                   self.padding[0])
            self._plan_cache = {key: plan}
        return plan

    def dense_backward(self, x, w, go, need_dx):
        """-> (dL/dx or None, dL/dw) for the forward that dense_forward just ran on the same x."""
        kind = getattr(self, "_kind", 'lib')
        if kind == 'gemm':
            return (K.linear_dgrad(go, w) if need_dx else None), K.linear_wgrad(go, x)
        if kind == 'c1x1':
            w2 = w.view(w.shape[0], w.shape[1])
            gw = K.conv1x1_wgrad(go, x).view_as(w)
            return (K.conv1x1_dgrad(go, w2) if need_dx else None), gw
        if kind == 'taps':
            plan = self._plan(x)
            self._gp = K.recon_stage_grad(go, plan, getattr(self, "_gp", None))
            gw = K.recon_conv_wgrad(self._gp, self._xp, plan, w.shape[0], w.shape[1])
            gx = K.recon_conv_dgrad(self._gp, plan, self._wd) if need_dx else None
            return gx, gw
        if kind == 'dw':
            k, st, pd = w.shape[2], self.stride[0], self.padding[0]
            gw = K.dwconv2d_wgrad(x, go, k, st, pd)
            gx = K.dwconv2d_dgrad(go, w, (x.shape[2], x.shape[3]), st, pd) if need_dx else None
            return gx, gw
        if kind == 'stem' and not need_dx:
            try:
                return None, K.conv_im2col_wgrad(x, go, (w.shape[2], w.shape[3]), self.stride[0], self.padding[0])
            except K.GemmUnsupported:      # output plane not a multiple of 4 pixels (TMA stride rule)
                pass
        if self.type == 'Gemm':
            gw = go.t() @ x
            gx = go @ w if need_dx else None
            return gx, gw
        if kind != 'lib':
            self._lib_notice()
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
