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
