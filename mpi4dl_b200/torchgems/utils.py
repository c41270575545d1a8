"""torchgems.utils mirror (reference src/torchgems/utils.py:20-30)."""

_LAYERS_PER_BLOCK = {1: 6, 2: 9}      # ResNet v1: basic blocks (6n+2 layers); v2: bottlenecks (9n+2)


def isPowerTwo(num):
    """True for 1, 2, 4, ... (a single set bit)."""
    return num > 0 and bin(num).count("1") == 1


def get_depth(version, n):
    """Depth of the Keras-style CIFAR ResNet with n blocks per stage."""
    try:
        return _LAYERS_PER_BLOCK[version] * n + 2
    except KeyError:
        raise ValueError("ResNet version must be 1 or 2") from None
