"""Synthetic workloads named by BASELINE.json: blob inventories and (see build_*) the
BN-folded ResNet-50 / MobileNetV2 graphs with random-init weights. There is no network
for checkpoints or datasets, so weights are seeded random and images are N(0, 1)."""
import numpy as np


def resnet50_blob_shapes():
    """Per-image shapes of the 123 activation blobs of BN-folded ResNet-50 at 3x224x224
    (network input + every node output, in node order)."""
    shapes = [(3, 224, 224), (64, 112, 112), (64, 112, 112), (64, 56, 56)]
    res = 56
    inplanes = 64
    for stage, (planes, blocks) in enumerate([(64, 3), (128, 4), (256, 6), (512, 3)]):
        for blk in range(blocks):
            stride = 2 if (blk == 0 and stage > 0) else 1
            out_res = res // stride
            shapes += [(planes, res, res)] * 2            # conv1, relu
            shapes += [(planes, out_res, out_res)] * 2    # conv2 (stride here, torchvision v1.5), relu
            shapes += [(planes * 4, out_res, out_res)]    # conv3
            if blk == 0:
                shapes += [(planes * 4, out_res, out_res)]  # downsample conv
            shapes += [(planes * 4, out_res, out_res)] * 2  # add, relu
            res = out_res
            inplanes = planes * 4
    shapes += [(2048, 1, 1), (2048,), (1000,)]
    return shapes


def blob_elements(shapes):
    return int(sum(int(np.prod(s)) for s in shapes))


# --------------------------------------------------------------------------- graphs
def _fold_bn(conv, bn):
    """Conv + eval-mode BatchNorm -> (weight, bias) float32, as onnxsim / the TorchScript
    exporter leave a torchvision model."""
    import torch
    with torch.no_grad():
        w = conv.weight.double()
        b = conv.bias.double() if conv.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64)
        f = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
        w = w * f.reshape(-1, 1, 1, 1)
        b = (b - bn.running_mean.double()) * f + bn.bias.double()
    return w.float().numpy(), b.float().numpy()


def _randomise_bn(model, gen):
    """Non-trivial BN statistics (default-initialised BN folds to identical biases, which
    the exporter dedupes into Identity nodes — SURVEY.md §8d)."""
    import torch
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) + 0.5)
                m.weight.copy_(torch.rand(m.weight.shape, generator=gen) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)


class _Builder:
    def __init__(self, name, input_shape):
        from . import onnx_lite as ol
        self.ol = ol
        self.g = ol.Graph(name)
        self.g.inputs.append(ol.ValueInfo("input", ol.FLOAT, [1] + list(input_shape)))
        self.k = 0

    def _out(self):
        self.k += 1
        return f"t{self.k}"

    def add(self, op, inputs, attrs=None, inits=None):
        out = self._out()
        names = list(inputs)
        for suffix, arr in (inits or []):
            nm = f"{op.lower()}{self.k}_{suffix}"
            self.g.initializers[nm] = arr
            names.append(nm)
        self.g.nodes.append(self.ol.Node(op, names, [out], f"{op}_{len(self.g.nodes)}", attrs))
        return out

    def conv(self, x, conv, bn=None):
        import torch
        if bn is not None:
            w, b = _fold_bn(conv, bn)
        else:
            w = conv.weight.detach().float().numpy()
            b = conv.bias.detach().float().numpy() if conv.bias is not None else None
        attrs = {"dilations": list(conv.dilation), "group": int(conv.groups),
                 "kernel_shape": list(conv.kernel_size),
                 "pads": list(conv.padding) * 2, "strides": list(conv.stride)}
        inits = [("weight", w)] + ([("bias", b)] if b is not None else [])
        return self.add("Conv", [x], attrs, inits)

    def finish(self, out, out_shape):
        self.g.outputs.append(self.ol.ValueInfo(out, self.ol.FLOAT, [1] + list(out_shape)))
        return self.ol.Model(self.g, ir_version=7, opsets={"": 13})


class _MiniResNet:
    """Bottleneck ResNet with free stage widths (torchvision.models.ResNet hard-codes 64..512);
    same attribute names, same initialisation, used for small test fixtures."""

    def __init__(self, blocks, planes, stem, num_classes):
        import torch.nn as nn
        from torchvision.models.resnet import Bottleneck
        self.conv1 = nn.Conv2d(3, stem, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(stem)
        inplanes = stem
        self.layers = []
        for i, (p, n) in enumerate(zip(planes, blocks)):
            stage = []
            for j in range(n):
                stride = 2 if (j == 0 and i > 0) else 1
                down = None
                if stride != 1 or inplanes != p * 4:
                    down = nn.Sequential(nn.Conv2d(inplanes, p * 4, 1, stride, bias=False),
                                         nn.BatchNorm2d(p * 4))
                stage.append(Bottleneck(inplanes, p, stride, down))
                inplanes = p * 4
            self.layers.append(stage)
        self.fc = nn.Linear(inplanes, num_classes)
        self._mods = [self.conv1, self.bn1, self.fc] + [b for st in self.layers for b in st]
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def modules(self):
        for top in self._mods:
            yield from top.modules()

    def eval(self):
        for m in self._mods:
            m.eval()


def build_resnet50(seed=0, blocks=None, width=64, num_classes=1000, image=224, planes=None, stem=64):
    """BN-folded torchvision ResNet-50 (53 Conv, 49 Relu, 16 Add, MaxPool,
    GlobalAveragePool, Flatten, Gemm = 122 nodes, 123 blobs) with seeded random weights.
    `blocks` / `width` / `planes` / `stem` shrink it for tests and fixtures."""
    import torch
    import torchvision
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    if planes is not None:
        net = _MiniResNet(blocks or [1, 1, 1, 1], planes, stem, num_classes)
        stages = net.layers
    else:
        net = torchvision.models.ResNet(torchvision.models.resnet.Bottleneck, blocks or [3, 4, 6, 3],
                                        num_classes=num_classes, width_per_group=width)
        stages = (net.layer1, net.layer2, net.layer3, net.layer4)
    _randomise_bn(net, gen)
    net.eval()
    b = _Builder("resnet50", (3, image, image))
    x = b.conv("input", net.conv1, net.bn1)
    x = b.add("Relu", [x])
    x = b.add("MaxPool", [x], {"kernel_shape": [3, 3], "pads": [1, 1, 1, 1], "strides": [2, 2],
                               "ceil_mode": 0})
    for layer in stages:
        for blk in layer:
            idt = x
            y = b.add("Relu", [b.conv(x, blk.conv1, blk.bn1)])
            y = b.add("Relu", [b.conv(y, blk.conv2, blk.bn2)])
            y = b.conv(y, blk.conv3, blk.bn3)
            if blk.downsample is not None:
                idt = b.conv(x, blk.downsample[0], blk.downsample[1])
            x = b.add("Relu", [b.add("Add", [y, idt])])
    x = b.add("GlobalAveragePool", [x])
    x = b.add("Flatten", [x], {"axis": 1})
    x = b.add("Gemm", [x], {"alpha": 1.0, "beta": 1.0, "transB": 1},
              [("weight", net.fc.weight.detach().float().numpy()),
               ("bias", net.fc.bias.detach().float().numpy())])
    return b.finish(x, (num_classes,))


def build_mobilenetv2(seed=0, width_mult=1.0, num_classes=1000, image=224):
    """BN-folded torchvision MobileNetV2 (52 Conv of which 17 depthwise, 35 Clip, 10 Add,
    GlobalAveragePool, Flatten, Gemm; 101 blobs)."""
    import torch
    import torchvision
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    net = torchvision.models.MobileNetV2(num_classes=num_classes, width_mult=width_mult)
    _randomise_bn(net, gen)
    net.eval()
    b = _Builder("mobilenetv2", (3, image, image))
    lo, hi = np.array(0., np.float32), np.array(6., np.float32)

    def conv_bn_relu6(x, seq):
        y = b.conv(x, seq[0], seq[1])
        return b.add("Clip", [y], None, [("min", lo), ("max", hi)])

    x = "input"
    feats = list(net.features)
    x = conv_bn_relu6(x, feats[0])
    for blk in feats[1:-1]:
        layers = list(blk.conv)
        y = x
        for sub in layers[:-2]:          # expand (optional) and depthwise, each Conv-BN-ReLU6
            y = conv_bn_relu6(y, sub)
        y = b.conv(y, layers[-2], layers[-1])  # linear projection
        x = b.add("Add", [x, y]) if blk.use_res_connect else y
    x = conv_bn_relu6(x, feats[-1])
    x = b.add("GlobalAveragePool", [x])
    x = b.add("Flatten", [x], {"axis": 1})
    fc = net.classifier[1]
    x = b.add("Gemm", [x], {"alpha": 1.0, "beta": 1.0, "transB": 1},
              [("weight", fc.weight.detach().float().numpy()),
               ("bias", fc.bias.detach().float().numpy())])
    return b.finish(x, (num_classes,))


def build_preact_net(seed=0, width=8, blocks=2, num_classes=10, image=32):
    """A small pre-activation residual net whose BatchNormalization nodes follow a Relu or an Add, so that no
    simplifier can fold them into a Conv: the model family `--update_bn` exists for (update_bn.py:13-25)."""
    gen = np.random.default_rng(seed)

    def w(*shape):
        fan_in = int(np.prod(shape[1:]))
        return (gen.standard_normal(shape) * (2.0 / fan_in) ** .5).astype(np.float32)

    def bn(b, x, c):
        return b.add("BatchNormalization", [x], {"epsilon": 1e-5, "momentum": 0.9},
                     [("scale", gen.uniform(0.5, 1.5, c).astype(np.float32)),
                      ("bias", (gen.standard_normal(c) * 0.1).astype(np.float32)),
                      ("mean", (gen.standard_normal(c) * 0.2).astype(np.float32)),
                      ("var", gen.uniform(0.5, 1.5, c).astype(np.float32))])

    def conv(b, x, ci, co, k, stride=1):
        return b.add("Conv", [x], {"dilations": [1, 1], "group": 1, "kernel_shape": [k, k],
                                   "pads": [k // 2] * 4, "strides": [stride, stride]},
                     [("weight", w(co, ci, k, k)), ("bias", (gen.standard_normal(co) * 0.05).astype(np.float32))])

    b = _Builder("preact", (3, image, image))
    x = b.add("Relu", [conv(b, "input", 3, width, 3)])
    for _ in range(blocks):
        y = conv(b, b.add("Relu", [bn(b, x, width)]), width, width, 3)
        y = conv(b, b.add("Relu", [y]), width, width, 1)
        x = b.add("Add", [x, y])
    x = b.add("Relu", [bn(b, x, width)])
    x = b.add("GlobalAveragePool", [x])
    x = b.add("Flatten", [x], {"axis": 1})
    x = b.add("Gemm", [x], {"alpha": 1.0, "beta": 1.0, "transB": 1},
              [("weight", w(num_classes, width)), ("bias", np.zeros(num_classes, np.float32))])
    return b.finish(x, (num_classes,))


def synthetic_images(n, shape=(3, 224, 224), seed=0, start=0):
    """Image `idx` is a function of (seed, idx) only, so shards and the CPU baseline see
    the same data whatever the batch size. float32 N(0, 1), [n, 1, C, H, W]."""
    out = np.empty((n, 1) + tuple(shape), dtype=np.float32)
    for i in range(n):
        out[i, 0] = np.random.default_rng([seed, start + i]).standard_normal(shape, dtype=np.float32)
    return out


def write_input_dir(images, input_dir, input_name="input", start=0):
    """The reference's on-disk layout: {input_dir}/{input_name}/{idx}.bin raw float32
    (dipoorlet/forward_net.py:459-464)."""
    import os
    d = os.path.join(input_dir, input_name)
    os.makedirs(d, exist_ok=True)
    for i in range(images.shape[0]):
        images[i].astype(np.float32).tofile(os.path.join(d, f"{start + i}.bin"))
