"""models.resnet_spatial_d2 -- the "D2" (fused halo) ResNet of the reference (src/models/resnet_spatial_d2.py):
instead of one halo exchange per 3x3 convolution, ONE wide `halo_exchange_layer` in front of a group of
`fused_layers` bottleneck cells, whose convolutions then run as VALID convolutions (conv_spatial(halo_len=0,
padding=0): no exchange, no padding, the tile shrinks) and whose residual branch is cropped to match
(resnet_spatial_d2.py:462-480).

get_resnet_v2 (:578-727): cells below `end_layer` (the last cell of pipeline stage spatial_size-1) are spatial; every
`fused_layers`-th of them is preceded by a halo_exchange_layer named "<cell>_halo" whose width follows the reference's
formulae (:651-698) -- a stride-2 cell needs twice the halo of what follows it -- and `balance[0]` grows by one per
inserted layer (the builder RETURNS the adjusted balance, :727).  get_resnet_v1 has no D2 behaviour in the reference
(:322-394 equals resnet_spatial.get_resnet_v1).  State-dict keys and module order are the reference's
(tests/golden/model_d2_golden.json, generated from the unmodified reference by tools/gen_model_d2_golden.py).

Deviation: the reference builds the cells' convolutions with the default slice_method ("square") whatever was asked
for (:118-170 never forwards it); valid convolutions do not depend on the grid, so nothing is lost by passing it on.
"""
from collections import OrderedDict

import torch.nn as nn

from . import resnet_spatial
from .resnet import _Head, _SpatialCtx, get_start_end_layer_index, make_cell_v2, resnet_layer


def get_balance(num_layers, mp_size):
    """Even split, remainder to the last stage (resnet_spatial_d2.py:287-292)."""
    per = int(num_layers / mp_size)
    balance = [per] * mp_size
    balance[mp_size - 1] += num_layers - sum(balance)
    return balance


class _ValidCtx(_SpatialCtx):
    """Cells of a fused group: conv_spatial(halo_len=0, padding=0) -- a valid convolution of the widened tile."""

    def conv(self, cin, cout, k, stride):
        from ..torchgems.spatial import conv_spatial
        return conv_spatial(in_channels=cin, out_channels=cout, kernel_size=k, stride=stride, padding=0, halo_len=0, **self.kw)


class make_cell_v2_spatial(make_cell_v2):
    """Bottleneck cell on a widened tile (resnet_spatial_d2.py:396-480): the two valid 3x3 convolutions eat 2 pixels per
    side (3 when the first one has stride 2, counted on the input grid), so the shortcut is cropped by as much."""

    def __init__(self, resblock, strides, in_filters, out_filters1, out_filters2, activation, batch_normalization, halo_len, ctx):
        super().__init__(resblock, strides, in_filters, out_filters1, out_filters2, activation, batch_normalization, ctx=ctx)
        self.halo_len = halo_len
        self.strides = strides

    def forward(self, x):
        if self.halo_len > 0:
            temp = x[:, :, 3:-3, 3:-3] if self.strides == 2 else x[:, :, 2:-2, 2:-2]
        else:
            temp = x
        y = self.r3(self.r2(self.r1(x)))
        if self.project:
            temp = self.r4(temp)
        return temp + y


def get_resnet_v1(input_shape, depth, local_rank, mp_size, spatial_size=1, num_spatial_parts=4, balance=None, num_classes=10,
                  slice_method="square"):
    return resnet_spatial.get_resnet_v1(input_shape, depth, local_rank, mp_size, spatial_size=spatial_size,
                                        num_spatial_parts=num_spatial_parts, balance=balance, num_classes=num_classes,
                                        slice_method=slice_method)


def halo_len_for(name, res_block, stage, num_res_blocks, fused_layers, end_layer):
    """Width of the halo_exchange_layer in front of cell `name` (resnet_spatial_d2.py:651-698)."""
    without_stride = num_res_blocks - res_block
    if name + fused_layers - 1 < end_layer:           # a full group of fused_layers cells follows
        if res_block == 0 and stage != 0:             # stride 2 needs double halo len
            return 2 * (2 * fused_layers - 1) + 1
        if fused_layers > without_stride:             # the group runs into the next stage's stride-2 cell
            return 2 * (fused_layers - without_stride) + 2 * (2 * without_stride - 1) + 1
        return 2 * fused_layers
    if res_block == 0 and stage != 0:                  # the last, shorter group
        return 2 * (2 * (end_layer - name) - 1) + 1
    if end_layer - name + 1 > without_stride:
        return 2 * (end_layer - name - without_stride) + 2 * (2 * without_stride - 1) + 1
    return 2 * (end_layer - name)


def get_resnet_v2(input_shape, depth, local_rank, mp_size, spatial_size=1, num_spatial_parts=4, balance=None, num_classes=10,
                  fused_layers=1, slice_method="square"):
    """Returns (model, balance) like the reference (:727)."""
    if (depth - 2) % 9 != 0:
        raise ValueError("depth should be 9n+2 (eg 56 or 110 in [b])")
    from ..torchgems.spatial import halo_exchange_layer

    num_res_blocks = int((depth - 2) / 9)
    num_layers = num_res_blocks * 3 + 2
    balance = list(balance) if balance is not None else get_balance(num_layers, mp_size)
    _, end_layer = get_start_end_layer_index(num_layers, balance, mp_size, local_rank=spatial_size - 1)
    assert fused_layers <= num_res_blocks, "number of fused layers greater than num_res_blocks is not supported"
    halo_kw = dict(local_rank=local_rank, spatial_size=spatial_size, num_spatial_parts=num_spatial_parts, slice_method=slice_method)
    stem_ctx = _SpatialCtx(local_rank, spatial_size, num_spatial_parts, slice_method)
    valid_ctx = _ValidCtx(local_rank, spatial_size, num_spatial_parts, slice_method)

    layers = OrderedDict()
    layers["0"] = resnet_layer(3, ctx=stem_ctx)
    name, in_filters, width, halo_temp = 1, 16, 16, 0
    for stage in range(3):
        for res_block in range(num_res_blocks):
            strides = 2 if (stage > 0 and res_block == 0) else 1
            cout = width * (4 if stage == 0 else 2)
            plain_first = stage == 0 and res_block == 0
            act, bn = (None, False) if plain_first else ("relu", True)
            if name >= end_layer:
                layers[str(name)] = make_cell_v2(res_block, strides, in_filters, width, cout, act, bn)
            else:
                cell_halo = 1
                if halo_temp == 0:
                    hl = halo_len_for(name, res_block, stage, num_res_blocks, fused_layers, end_layer)
                    layers[str(name) + "_halo"] = halo_exchange_layer(halo_len=hl, **halo_kw)
                    cell_halo = 2
                    balance[0] += 1
                layers[str(name)] = make_cell_v2_spatial(res_block, strides, in_filters, width, cout, act, bn, cell_halo, valid_ctx)
                halo_temp = (halo_temp + 1) % fused_layers
            name += 1
            in_filters = cout
        width = cout
    layers[str(name)] = _Head(8, int(width), input_shape[2], num_classes, pre_norm=True)
    return nn.Sequential(layers), balance
