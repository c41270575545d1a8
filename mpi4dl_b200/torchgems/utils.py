"""torchgems.utils mirror (reference src/torchgems/utils.py:20-30)."""


def isPowerTwo(num):
    return num > 0 and (num & (num - 1)) == 0


def get_depth(version, n):
    """Depth of the Keras-style ResNet: v1 = 6n+2, v2 = 9n+2."""
    if version == 1:
        return n * 6 + 2
    if version == 2:
        return n * 9 + 2
    raise ValueError("ResNet version must be 1 or 2")
