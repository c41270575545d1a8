"""ResNet v1 / v2 (CIFAR-style, He et al.) in the layout the reference's SP+LP scripts expect: a flat
nn.Sequential of cells that `torchgems.mp_pipeline.model_generator` cuts into stages, with the
first stages' convolutions replaced by `torchgems.spatial.conv_spatial` when they run on tiles.

One builder serves the three reference modules:
    models/resnet.py          get_resnet_v1 / get_resnet_v2              (:145-178, :270-323)
    models/resnet_spatial.py  get_resnet_v1 / get_resnet_v2 (spatial)    (:304-389, :545-633)
State-dict keys ("<cell>.r1.conv1.weight", "<cell>.r1.batch_first.*", "<n>.fc1.*", cells numbered
from "1" in the sequential module and "0" in the spatial one) match the reference's, so its
checkpoints load.  As in the reference, BatchNorm inside a spatial stage normalises over the local
tile only.
"""
from collections import OrderedDict

import torch.nn as nn
import torch.nn.functional as F


class _SpatialCtx:
    """Which conv class a cell uses.  `None` -> nn.Conv2d; else conv_spatial bound to a tile."""

    def __init__(self, local_rank, spatial_size, num_spatial_parts, slice_method):
        self.kw = dict(local_rank=local_rank, spatial_size=spatial_size, num_spatial_parts=num_spatial_parts,
                       slice_method=slice_method)

    def conv(self, cin, cout, k, stride):
        from ..torchgems.spatial import conv_spatial
        return conv_spatial(in_channels=cin, out_channels=cout, kernel_size=k, stride=stride, padding=(k - 1) // 2,
                            **self.kw)


def _conv(ctx, cin, cout, k, stride):
    if ctx is None:
        return nn.Conv2d(cin, cout, k, stride=stride, padding=(k - 1) // 2)
    return ctx.conv(cin, cout, k, stride)


class resnet_layer(nn.Module):
    """conv-BN-ReLU (v1 order) or BN-ReLU-conv (v2 pre-activation order).  Both BatchNorms exist
    whichever is used, as in the reference (resnet.py:24-78), to keep the state-dict identical."""

    def __init__(self, in_num_filters, num_filters=16, kernel_size=3, strides=1, activation="relu",
                 batch_normalization=True, conv_first=True, ctx=None):
        super().__init__()
        self.conv_first = conv_first
        self.activation = activation
        self.batch_normalization = batch_normalization
        self.conv1 = _conv(ctx, in_num_filters, num_filters, kernel_size, strides)
        self.batch_first = nn.BatchNorm2d(in_num_filters)
        self.batch_last = nn.BatchNorm2d(num_filters)
        self.act = nn.ReLU()

    def _norm_act(self, x, bn):
        if self.batch_normalization:
            x = bn(x)
        return x if self.activation is None else self.act(x)

    def forward(self, x):
        if self.conv_first:
            return self._norm_act(self.conv1(x), self.batch_last)
        return self.conv1(self._norm_act(x, self.batch_first))


class make_cell_v1(nn.Module):
    """Basic block: two 3x3 units + identity / 1x1-projection shortcut, ReLU after the add."""

    def __init__(self, stack, resblock, strides, in_filters, out_filters, ctx=None):
        super().__init__()
        self.r1 = resnet_layer(in_filters, out_filters, strides=strides, ctx=ctx)
        self.r2 = resnet_layer(out_filters, out_filters, activation=None, ctx=ctx)
        self.project = resblock == 0 and stack > 0
        if self.project:
            self.r3 = resnet_layer(in_filters, out_filters, kernel_size=1, strides=strides, activation=None,
                                   batch_normalization=False, ctx=ctx)

    def forward(self, x):
        y = self.r2(self.r1(x))
        if self.project:
            x = self.r3(x)
        return F.relu(x + y)


class make_cell_v2(nn.Module):
    """Pre-activation bottleneck: 3x3, 3x3, 1x1 units (the reference's layout, resnet.py:181-231),
    1x1 projection on the first block of a stage, no ReLU after the add."""

    def __init__(self, resblock, strides, in_filters, out_filters1, out_filters2, activation, batch_normalization,
                 ctx=None):
        super().__init__()
        self.r1 = resnet_layer(in_filters, out_filters1, strides=strides, activation=activation,
                               batch_normalization=batch_normalization, conv_first=False, ctx=ctx)
        self.r2 = resnet_layer(out_filters1, out_filters1, conv_first=False, ctx=ctx)
        self.r3 = resnet_layer(out_filters1, out_filters2, kernel_size=1, conv_first=False, ctx=ctx)
        self.project = resblock == 0
        if self.project:
            self.r4 = resnet_layer(in_filters, out_filters2, kernel_size=1, strides=strides, activation=None,
                                   batch_normalization=False, ctx=ctx)

    def forward(self, x):
        y = self.r3(self.r2(self.r1(x)))
        if self.project:
            x = self.r4(x)
        return x + y


class _Head(nn.Module):
    """[BN-ReLU (v2)] - AvgPool(8) - flatten - Linear - softmax  (resnet.py:117-142, :234-267)."""

    def __init__(self, kernel_size, num_filters, image_size, num_classes, pre_norm):
        super().__init__()
        if pre_norm:
            self.batch_last = nn.BatchNorm2d(num_filters)
        self.pre_norm = pre_norm
        self.pool = nn.AvgPool2d(kernel_size)
        side = int(image_size / (4 * kernel_size))
        self.flatten_size = num_filters * side * side
        self.fc1 = nn.Linear(self.flatten_size, num_classes)

    def forward(self, x):
        if self.pre_norm:
            x = F.relu(self.batch_last(x))
        x = self.pool(x).view(-1, self.flatten_size)
        return F.softmax(self.fc1(x), dim=1)


def get_start_end_layer_index(num_layers, balance, mp_size, local_rank=0):
    """Cells [start, end) of pipeline stage `local_rank` (resnet_spatial.py:270-301)."""
    if balance is None:
        per = int(num_layers / mp_size)
        start = local_rank * per
        return start, (start + per if local_rank != mp_size - 1 else num_layers)
    assert sum(balance) == num_layers, "balance and number of layers differs"
    start = sum(balance[:local_rank])
    return start, start + balance[local_rank]


def _build(version, input_shape, depth, num_classes, first_name, spatial=None, n_spatial_cells=0):
    """`spatial` is a _SpatialCtx; cells with index < n_spatial_cells are built on it."""
    per_block = 6 if version == 1 else 9
    if (depth - 2) % per_block != 0:
        raise ValueError("depth should be 6n+2 (eg 20, 32, 44 in [a])" if version == 1 else
                         "depth should be 9n+2 (eg 56 or 110 in [b])")
    blocks = (depth - 2) // per_block
    cells = []

    def ctx():
        return spatial if len(cells) < n_spatial_cells else None

    cells.append(resnet_layer(3, ctx=ctx()))
    cin, width = 16, 16
    for stage in range(3):
        for b in range(blocks):
            stride = 2 if (stage > 0 and b == 0) else 1
            if version == 1:
                cells.append(make_cell_v1(stage, b, stride, cin, width, ctx=ctx()))
                cin = width
            else:
                cout = width * (4 if stage == 0 else 2)
                plain_first = stage == 0 and b == 0          # the stem already normalised + activated
                cells.append(make_cell_v2(b, stride, cin, width, cout, None if plain_first else "relu",
                                          not plain_first, ctx=ctx()))
                cin = cout
        width = width * 2 if version == 1 else cout
    head_filters = width // 2 if version == 1 else width
    cells.append(_Head(8, int(head_filters), input_shape[2], num_classes, pre_norm=(version == 2)))
    return nn.Sequential(OrderedDict((str(first_name + i), c) for i, c in enumerate(cells)))


def get_resnet_v1(input_shape, depth, num_classes=10):
    return _build(1, input_shape, depth, num_classes, first_name=1)


def get_resnet_v2(input_shape, depth, num_classes=10):
    return _build(2, input_shape, depth, num_classes, first_name=1)


def num_cells(version, depth):
    """Length of the Sequential (what `balance` must sum to): depth/2+1 for v1, 3n+2 for v2."""
    return int(depth / 2 + 1) if version == 1 else ((depth - 2) // 9) * 3 + 2
