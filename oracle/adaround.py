"""Learned rounding (AdaRound / BRECQ / QDrop) restated with torch autograd, op for op as
the reference evaluates it, so that a CPU run reproduces the reference's rounded weights
exactly (pinned by tests/golden/*/wt_adaround.npz, wt_brecq.npz) and a CUDA run is the
"plain PyTorch fp32 reference" the fused kernels are checked against.

Follows dipoorlet/weight_transform/ada_quant_layer.py:28-50,96-130,133-252,
adaround.py:62-103,119-144, brecq.py:62-112,158-200.
"""
import numpy as np
import torch
import torch.nn.functional as F

ZETA, GAMMA = 1.1, -0.1


def rectified_sigmoid(round_mask):
    """ada_quant_layer.py:105-106."""
    return ((ZETA - GAMMA) * torch.sigmoid(round_mask) + GAMMA).clamp(0, 1)


def temp_decay(t, t_max, rel_start_decay=0.2, start_b=20, end_b=2):
    """ada_quant_layer.py:117-130."""
    start_decay = rel_start_decay * t_max
    if t < start_decay:
        return 0.0
    rel_t = (t - start_decay) / (t_max - start_decay)
    return end_b + 0.5 * (start_b - end_b) * (1 + np.cos(rel_t * np.pi))


def reg_loss(round_mask, beta, alpha=0.01):
    """ada_quant_layer.py:108-110."""
    return alpha * (1 - torch.pow((rectified_sigmoid(round_mask) - 0.5).abs() * 2, beta)).sum()


def quant_weight(weight, round_mask, scale, q_min, q_max, soft=True):
    """ada_quant_layer.py:39-50, per-channel branch (trt)."""
    if soft:
        weight = (weight / scale).floor() + rectified_sigmoid(round_mask)
    else:
        weight = (weight / scale).floor() + (round_mask >= 0).float()
    weight = torch.max(weight, q_min)
    weight = torch.min(weight, q_max)
    return weight * scale


def quant_acti(x, scale, q_min, q_max, prob, generator=None):
    """ada_quant_layer.py:28-36 (round has no straight-through gradient)."""
    x_ori = x
    x = (x / scale).round()
    x = torch.max(x, q_min)
    x = torch.min(x, q_max)
    x = x * scale
    if prob < 1.0:
        x = torch.where(torch.rand(x.shape, generator=generator, device=x.device) < prob, x, x_ori)
    return x


def l2_norm(pred, tgt):
    """ada_quant_layer.py:113-114."""
    return (pred - tgt).pow(2.0).sum(1).mean()


def alpha_init(weight, scale):
    """adaround.py:71 + ada_quant_layer.py:153."""
    rest = (weight / scale) - (weight / scale).floor()
    return -torch.log((ZETA - GAMMA) / (rest - GAMMA) - 1)


class Layer:
    """One AdaQLayer (Conv or Gemm): frozen weight / bias, learnable round_mask."""

    def __init__(self, op_type, attrs, weight, bias, scale, q_min, q_max, relu_flag, qi=None,
                 acti_quant=False):
        self.op_type = op_type
        self.attrs = attrs
        self.weight = weight
        self.bias = bias
        self.scale, self.q_min, self.q_max = scale, q_min, q_max
        self.relu_flag = relu_flag
        self.qi = qi                      # (scale, q_min, q_max) 0-d tensors
        self.acti_quant = acti_quant
        self.round_mask = alpha_init(weight, scale).clone().requires_grad_(True)

    def forward(self, x, generator=None):
        """AdaQLayer.forward, ada_quant_layer.py:224-251."""
        w = quant_weight(self.weight, self.round_mask, self.scale, self.q_min, self.q_max)
        if self.op_type == "Conv":
            a = self.attrs
            x = F.conv2d(x, w, self.bias, a["strides"], a["pads"][:2], a["dilations"], a["group"])
        else:
            x = F.linear(x, w, self.bias)
        if self.relu_flag:
            x = F.relu(x)
        if self.acti_quant and self.qi is not None:
            x = quant_acti(x, self.qi[0], self.qi[1], self.qi[2], 0.5, generator)
        return x

    def hard_weight(self):
        return quant_weight(self.weight, self.round_mask.detach(), self.scale, self.q_min, self.q_max,
                            soft=False)


def learn(layers, q_in, tgt, total_iter, batch_size, max_epoch, fp_in=None, drop=False,
          generator=None, record=None):
    """learning_round_mask of adaround.py:119-144 (one layer) / brecq.py:158-200 (block).
    `record(cur_iter, loss, layers)` is called after every optimiser step when given."""
    opt = torch.optim.Adam([layer.round_mask for layer in layers])
    cur_iter = 0
    ratio = 0.5 if drop else 1.0
    loss = None
    for epoch in range(max_epoch):
        if ratio < 1.0:
            in_tensor = torch.where(torch.rand(q_in.shape, generator=generator, device=q_in.device) < ratio,
                                    q_in, fp_in)
        else:
            in_tensor = q_in
        for idx in range(int(np.ceil(len(in_tensor) / batch_size))):
            st = idx * batch_size
            ed = st + batch_size
            out = in_tensor[st:ed]
            for layer in layers:
                out = layer.forward(out, generator)
            beta = temp_decay(cur_iter, total_iter)
            loss = l2_norm(out, tgt[st:ed])
            for layer in layers:
                loss = loss + reg_loss(layer.round_mask, beta)
            cur_iter += 1
            opt.zero_grad()
            loss.backward()
            opt.step()
            if record is not None:
                record(cur_iter, float(loss), layers)
    return None if loss is None else float(loss.detach())
