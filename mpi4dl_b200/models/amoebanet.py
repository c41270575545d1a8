"""AmoebaNet-D as the reference's SP+LP scripts use it (src/models/amoebanet.py): a flat
nn.Sequential  stem1, stem2, stem3, cell1_normal*, cell2_reduction, cell3_normal*, cell4_reduction,
cell5_normal*, classify  whose cells pass (state, skip) tuples, cut into pipeline stages by
torchgems.mp_pipeline.model_generator.  `amoebanetd_spatial` builds the cells of pipeline stage 0
on torchgems.spatial layers (conv_spatial / Pool) bound to this rank's tile -- that spatial stage
at 8192x8192 is the workload bench.py measures (BASELINE.json configs[1]).

Written table-first: a cell is a list of (input state, op name) pairs and each op name maps to a
small layer recipe, instead of one constructor function per op.  The module tree (and so every
state-dict key: "stem2.reduce2.conv1.weight", "cell1_normal1.operations.3.module.4.weight", ...)
is the reference's, so its checkpoints load.  Reference behaviours kept on purpose:
  * the op called max_pool_3x3 is a 3x3 *average* pool (amoebanet.py:108-125);
  * FactorizedReduce uses two identical-offset 1x1 stride-2 convs (:56-76); they need no exchange (a 1x1
    stride-2 conv is tile-local for even tiles), so the reference leaves them plain nn.Conv2d (cuDNN) in
    spatial cells.  Here they are `torchgems.spatial.local_conv2d` there: same parameters and state-dict
    keys, but on the libspconv kernels -- no cuDNN convolution is left in a spatial stage;
  * likewise conv_1x1 and the outer 1x1s of conv_3x3 (:241-276); the 1x7/7x1 op is conv_spatial
    throughout (:147-238);
  * only pipeline stage 0 is spatial, and the cut-over cell is found with the reference's layer
    counter, which advances twice for stem2 and stem3 (:651-699).
"""
from collections import OrderedDict

import torch
import torch.nn as nn

# (input state, op) pairs; consecutive pairs are summed into a new state (amoebanet.py:296-347)
NORMAL_OPERATIONS = [(1, "conv_1x1"), (1, "max_pool_3x3"), (1, "none"), (0, "conv_1x7_7x1"), (0, "conv_1x1"),
                     (0, "conv_1x7_7x1"), (2, "max_pool_3x3"), (2, "none"), (1, "avg_pool_3x3"), (5, "conv_1x1")]
NORMAL_CONCAT = [0, 3, 4, 6]
REDUCTION_OPERATIONS = [(0, "max_pool_2x2"), (0, "max_pool_3x3"), (2, "none"), (1, "conv_3x3"), (2, "conv_1x7_7x1"),
                        (2, "max_pool_3x3"), (3, "none"), (1, "max_pool_2x2"), (2, "avg_pool_3x3"), (3, "conv_1x1")]
REDUCTION_CONCAT = [4, 5, 6]


def _pair(v):
    return v if isinstance(v, tuple) else (v, v)


def _conv(sp, cin, cout, k=1, stride=1, padding=0, local=False):
    """bias-free conv; `sp` (dict of tile-binding kwargs) selects conv_spatial.  `local`: a tile-local conv
    INSIDE a spatial cell (the reference's plain nn.Conv2d there, amoebanet.py:241-276): same parameters and
    state-dict keys, but run by libspconv (torchgems.spatial.local_conv2d) so that no cuDNN convolution is
    left in a spatial stage."""
    if sp is None:
        if local:
            from ..torchgems.spatial import local_conv2d
            return local_conv2d(cin, cout, k, stride=stride, padding=padding, bias=False)
        return nn.Conv2d(cin, cout, k, stride, padding, bias=False)
    from ..torchgems.spatial import conv_spatial
    return conv_spatial(in_channels=cin, out_channels=cout, kernel_size=k, stride=stride, padding=padding, bias=False, **sp)


def _pool(sp, kind, k, stride, padding):
    if sp is None:
        if kind == "AvgPool2d":
            return nn.AvgPool2d(k, stride=stride, padding=padding, count_include_pad=False)
        return nn.MaxPool2d(k, stride=stride, padding=padding)
    from ..torchgems.spatial import Pool
    return Pool(operation=kind, kernel_size=k, stride=stride, padding=padding, count_include_pad=False, **sp)


def _relu_conv_bn_chain(convs, fused=False):
    """[ReLU, conv, BN] per conv, flattened into one Sequential (indices 0,1,2, 3,4,5, ...).  `fused` (cells of a
    spatial stage): the same children and state-dict keys, but every BatchNorm2d runs fused with the ReLU that
    follows it on libspconv (torchgems.fused.relu_conv_bn_chain)."""
    mods = []
    for c in convs:
        mods += [nn.ReLU(inplace=False), c, nn.BatchNorm2d(c.out_channels)]
    if fused:
        from ..torchgems.fused import relu_conv_bn_chain
        return relu_conv_bn_chain(*mods)
    return nn.Sequential(*mods)


def _bn(sp, bn, x):
    """BatchNorm2d of a cell: fused statistics + apply on libspconv inside a spatial stage."""
    if sp:
        from ..torchgems.fused import bn_relu
        return bn_relu(x, bn, relu=False)
    return bn(x)


def relu_conv_bn(sp, in_channels, out_channels, kernel_size=1, stride=1, padding=0):
    return _relu_conv_bn_chain([_conv(sp, in_channels, out_channels, kernel_size, stride, padding)], fused=sp is not None)


class FactorizedReduce(nn.Module):
    def __init__(self, in_channels, out_channels, local=False):
        super().__init__()
        self.relu = nn.ReLU(inplace=False)
        self.pad = nn.ZeroPad2d((0, 1, 0, 1))          # unused, kept for the module tree
        self.conv1 = _conv(None, in_channels, out_channels // 2, 1, 2, local=local)
        self.conv2 = _conv(None, in_channels, out_channels // 2, 1, 2, local=local)
        self.bn = nn.BatchNorm2d(out_channels)
        self._fused = local

    def forward(self, x):
        x = self.relu(x)
        return _bn(self._fused, self.bn, torch.cat([self.conv1(x), self.conv2(x)], dim=1))


def _make_op(name, sp, c, stride):
    q = c // 4
    loc = sp is not None          # tile-local convs of a spatial cell run on libspconv too
    if name == "none":
        return nn.Identity() if stride == 1 else FactorizedReduce(c, c, local=loc)
    if name in ("avg_pool_3x3", "max_pool_3x3"):
        return _pool(sp, "AvgPool2d", 3, stride, 1)
    if name == "max_pool_2x2":
        return _pool(sp, "MaxPool2d", 2, stride, 0)
    if name == "conv_1x1":
        return _relu_conv_bn_chain([_conv(None, c, c, 1, stride, local=loc)], fused=loc)
    if name == "conv_3x3":
        return _relu_conv_bn_chain([_conv(None, c, q, local=loc), _conv(sp, q, q, 3, stride, 1), _conv(None, q, c, local=loc)],
                                   fused=loc)
    if name == "conv_1x7_7x1":
        return _relu_conv_bn_chain([_conv(sp, c, q), _conv(sp, q, q, (1, 7), (1, stride), (0, 3)),
                                    _conv(sp, q, q, (7, 1), (stride, 1), (3, 0)), _conv(sp, q, c)], fused=loc)
    raise KeyError(name)


class Operation(nn.Module):
    """Named wrapper (the reference's, amoebanet.py:39-53): keeps the `.module` level in
    state-dict keys and the op name in repr."""

    def __init__(self, name, module):
        super().__init__()
        self.name = name
        self.module = module

    def __repr__(self):
        return "%s[%s]" % (self.__class__.__name__, self.name)

    def forward(self, *args):
        return self.module(*args)


class Stem(nn.Module):
    def __init__(self, sp, channels):
        super().__init__()
        self.conv = _conv(sp, 3, channels, 3, stride=2, padding=1)
        self.relu = nn.ReLU(inplace=False)
        self.bn = nn.BatchNorm2d(channels)
        self._fused = sp is not None

    def forward(self, x):
        return _bn(self._fused, self.bn, self.conv(self.relu(x)))


class Cell(nn.Module):
    def __init__(self, sp, channels_prev_prev, channels_prev, channels, reduction, reduction_prev):
        super().__init__()
        self.reduce1 = relu_conv_bn(sp, channels_prev, channels)
        if reduction_prev:
            self.reduce2 = FactorizedReduce(channels_prev_prev, channels, local=sp is not None)
        elif channels_prev_prev != channels:
            self.reduce2 = relu_conv_bn(sp, channels_prev_prev, channels)
        else:
            self.reduce2 = nn.Identity()
        table = REDUCTION_OPERATIONS if reduction else NORMAL_OPERATIONS
        self.concat = REDUCTION_CONCAT if reduction else NORMAL_CONCAT
        self.indices = tuple(i for i, _ in table)
        self.operations = nn.ModuleList(
            Operation(name, _make_op(name, sp, channels, 2 if (reduction and i < 2) else 1)) for i, name in table)

    def extra_repr(self):
        return "indices: %s" % (self.indices,)

    def forward(self, input_or_states):
        s1, s2 = input_or_states if isinstance(input_or_states, tuple) else (input_or_states, input_or_states)
        skip = s1
        states = [self.reduce1(s1), self.reduce2(s2)]
        for j in range(0, len(self.operations), 2):
            a = self.operations[j](states[self.indices[j]])
            b = self.operations[j + 1](states[self.indices[j + 1]])
            states.append(a + b)
        return torch.cat([states[i] for i in self.concat], dim=1), skip


class Classify(nn.Module):
    def __init__(self, channels_prev, num_classes):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool2d((1, 1))
        self.flat = nn.Flatten()
        self.fc = nn.Linear(channels_prev, num_classes)

    def forward(self, states):
        x, _ = states
        return self.fc(self.flat(self.pool(x)))


def get_start_end_layer_index(num_layers, balance, mp_size, local_rank=0):
    """Cells [start, end) of pipeline stage `local_rank` (amoebanet.py:713-737)."""
    if balance is None:
        per = int(num_layers / mp_size)
        start = local_rank * per
        return start, (start + per if local_rank != mp_size - 1 else num_layers)
    assert sum(balance) == num_layers, "balance and number of layers differs"
    start = sum(balance[:local_rank])
    return start, start + balance[local_rank]


def _plan(num_layers):
    """(name, reduction?, counter advance) for every cell after stem1, in order.  The advance
    column reproduces the reference's `layers_processed` bookkeeping (stem2/stem3 count twice)."""
    n = num_layers // 3
    plan = [("stem2", True, 2), ("stem3", True, 2)]
    plan += [("cell1_normal%d" % (i + 1), False, 1) for i in range(n)]
    plan += [("cell2_reduction", True, 1)]
    plan += [("cell3_normal%d" % (i + 1), False, 1) for i in range(n)]
    plan += [("cell4_reduction", True, 1)]
    plan += [("cell5_normal%d" % (i + 1), False, 1) for i in range(n)]
    return plan


def _assemble(num_classes, num_layers, num_filters, sp, end_layer):
    """`sp` None -> ordinary model.  Else cells whose counter value is < end_layer are spatial;
    once one is not, all later ones are not either."""
    assert num_layers % 3 == 0
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
        cell = Cell(sp, c_pp, c_p, channels, reduction, reduction_prev)
        c_pp, c_p = c_p, channels * len(cell.concat)
        reduction_prev = reduction
        layers[name] = cell
    layers["classify"] = Classify(c_p, num_classes)
    return nn.Sequential(layers)


def amoebanetd(num_classes=10, num_layers=4, num_filters=512):
    """AmoebaNet-D (num_layers = 3 x normal cells per group, num_filters = 4 x stem width)."""
    return _assemble(num_classes, num_layers, num_filters, None, 0)


def amoebanetd_spatial(local_rank, spatial_size, num_spatial_parts, mp_size, balance=None, slice_method="square",
                       num_classes=10, num_layers=4, num_filters=512):
    sp = dict(local_rank=local_rank, spatial_size=spatial_size, num_spatial_parts=num_spatial_parts, slice_method=slice_method)
    assert num_layers % 3 == 0
    _, end_layer = get_start_end_layer_index((num_layers // 3) * 3 + 6, balance, mp_size, local_rank=0)
    assert end_layer > 3, "There should be atleast 3 layers in "
    return _assemble(num_classes, num_layers, num_filters, sp, end_layer)
