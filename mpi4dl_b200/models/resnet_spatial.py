"""models.resnet_spatial -- ResNet whose first pipeline stages run on image tiles
(reference src/models/resnet_spatial.py: get_resnet_v1 :304-389, get_resnet_v2 :545-633).

Cells 0 .. end-1, where `end` is the last cell of pipeline stage `spatial_size - 1` under
`balance` / `mp_size`, use conv_spatial bound to this rank's tile; the rest are ordinary cells that
run on the stitched feature map after the join rank.  `fused_layers` is accepted and ignored, as
in the reference (it only matters for the D2 builder)."""
from .resnet import _SpatialCtx, _build, get_start_end_layer_index, num_cells  # noqa: F401


def _spatial(version, input_shape, depth, local_rank, mp_size, spatial_size, num_spatial_parts, balance, num_classes,
             slice_method):
    _, end = get_start_end_layer_index(num_cells(version, depth), balance, mp_size, local_rank=spatial_size - 1)
    ctx = _SpatialCtx(local_rank, spatial_size, num_spatial_parts, slice_method)
    return _build(version, input_shape, depth, num_classes, first_name=0, spatial=ctx, n_spatial_cells=max(end, 1))


def get_resnet_v1(input_shape, depth, local_rank, mp_size, spatial_size=1, num_spatial_parts=4, balance=None,
                  num_classes=10, slice_method="square"):
    return _spatial(1, input_shape, depth, local_rank, mp_size, spatial_size, num_spatial_parts, balance, num_classes,
                    slice_method)


def get_resnet_v2(input_shape, depth, local_rank, mp_size, spatial_size=1, num_spatial_parts=4, balance=None,
                  num_classes=10, fused_layers=1, slice_method="square"):
    return _spatial(2, input_shape, depth, local_rank, mp_size, spatial_size, num_spatial_parts, balance, num_classes,
                    slice_method)
