"""AmoebaNet-D, "D2" (fused halo) spatial variant (reference src/models/amoebanet_d2.py): a normal
cell that runs on tiles does TWO wide exchanges up front -- s3 = halo_exchange_layer(3)(s1) for the
1x7/7x1 branches, s4 = halo_exchange_layer(2)(s2) for the chained 3x3 pools -- and then runs every
op of the cell as a valid (padding=0) conv / pool on the widened tensors, recomputing the overlap
instead of exchanging per layer (Cell_D2, :569-676).  Reduction cells and everything past pipeline
stage 0 are the ordinary cells of models/amoebanet.py.

Same module tree / state-dict keys as the reference.  The valid convs and pools are
torchgems.spatial.local_conv2d / local_pool2d (libspconv kernels) where the reference calls plain
nn.Conv2d / nn.AvgPool2d.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from .amoebanet import (NORMAL_CONCAT, REDUCTION_CONCAT, REDUCTION_OPERATIONS, Cell, Classify, FactorizedReduce,  # noqa: F401
                        Operation, Stem, _make_op, _plan, _relu_conv_bn_chain, amoebanetd, get_start_end_layer_index,
                        relu_conv_bn)

# (input state, op): states are [s1, s2, s3 = s1 + halo 3, s4 = s2 + halo 2, s5 = s2 + halo 1, sums...]
NORMAL_OPERATIONS_D2 = [(4, "conv_1x1"), (3, "max_pool_3x3_d2"), (1, "none"), (2, "conv_1x7_7x1_d2"), (0, "conv_1x1"),
                        (2, "conv_1x7_7x1_d2"), (5, "max_pool_3x3_d2"), (5, "none"), (4, "avg_pool_3x3_d2"), (8, "conv_1x1")]
NORMAL_CONCAT_D2 = [0, 6, 7, 9]


def _local_conv(cin, cout, k=1, stride=1):
    from ..torchgems.spatial import local_conv2d
    return local_conv2d(cin, cout, k, stride=stride, padding=0, bias=False)


def _make_op_d2(name, c):
    q = c // 4
    if name in ("max_pool_3x3_d2", "avg_pool_3x3_d2"):       # both are 3x3 average pools (:88-117)
        from ..torchgems.spatial import local_pool2d
        return local_pool2d("AvgPool2d", 3, stride=1, padding=0)
    if name == "conv_1x7_7x1_d2":
        return _relu_conv_bn_chain([_local_conv(c, q), _local_conv(q, q, (1, 7)), _local_conv(q, q, (7, 1)),
                                    _local_conv(q, c)])
    return _make_op(name, None, c, 1)                         # conv_1x1 / none: the ordinary modules


class Cell_D2(nn.Module):
    def __init__(self, sp, channels_prev_prev, channels_prev, channels, reduction_prev):
        super().__init__()
        from ..torchgems.spatial import halo_exchange_layer
        self.reduce1 = relu_conv_bn(sp, channels_prev, channels)
        if reduction_prev:
            self.reduce2 = FactorizedReduce(channels_prev_prev, channels)
        elif channels_prev_prev != channels:
            self.reduce2 = relu_conv_bn(sp, channels_prev_prev, channels)
        else:
            self.reduce2 = nn.Identity()
        self.concat = NORMAL_CONCAT_D2
        self.indices = tuple(i for i, _ in NORMAL_OPERATIONS_D2)
        self.s3_layer = halo_exchange_layer(halo_len=3, **sp)
        self.s4_layer = halo_exchange_layer(halo_len=2, **sp)
        self.operations = nn.ModuleList(Operation(name.replace("_d2", ""), _make_op_d2(name, channels))
                                        for _, name in NORMAL_OPERATIONS_D2)

    def extra_repr(self):
        return "indices: %s" % (self.indices,)

    def forward(self, input_or_states):
        s1, s2 = input_or_states if isinstance(input_or_states, tuple) else (input_or_states, input_or_states)
        skip = s1
        s1, s2 = self.reduce1(s1), self.reduce2(s2)
        s3, s4 = self.s3_layer(s1), self.s4_layer(s2)
        states = [s1, s2, s3, s4, s4[:, :, 1:-1, 1:-1]]
        for j in range(0, len(self.operations), 2):
            a = self.operations[j](states[self.indices[j]])
            b = self.operations[j + 1](states[self.indices[j + 1]])
            if j == 6:                                        # pooled (H) + identity of the halo-1 sum (H+2)
                b = b[:, :, 1:-1, 1:-1]
            states.append(a + b)
        return torch.cat([states[i] for i in self.concat], dim=1), skip


def amoebanetd_spatial(local_rank, spatial_size, num_spatial_parts, mp_size, balance=None, slice_method="square",
                       num_classes=10, num_layers=4, num_filters=512):
    sp = dict(local_rank=local_rank, spatial_size=spatial_size, num_spatial_parts=num_spatial_parts, slice_method=slice_method)
    assert num_layers % 3 == 0
    _, end_layer = get_start_end_layer_index((num_layers // 3) * 3 + 6, balance, mp_size, local_rank=0)
    assert end_layer > 3, "There should be atleast 3 layers in "
    layers = OrderedDict()
    channels = num_filters // 4
    c_pp = c_p = channels
    reduction_prev = False
    layers["stem1"] = Stem(sp, channels)
    counter = 1
    for name, reduction, advance in _plan(num_layers):
        if sp is not None and counter >= end_layer:
            sp = None
        counter += advance
        if reduction:
            channels *= 2
        if not reduction and sp is not None:
            cell = Cell_D2(sp, c_pp, c_p, channels, reduction_prev)
        else:
            cell = Cell(sp, c_pp, c_p, channels, reduction, reduction_prev)
        c_pp, c_p = c_p, channels * len(cell.concat)
        reduction_prev = reduction
        layers[name] = cell
    layers["classify"] = Classify(c_p, num_classes)
    return nn.Sequential(layers)
