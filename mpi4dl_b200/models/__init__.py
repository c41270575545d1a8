"""Model builders the reference's SP/LP benchmark scripts import (src/models/): the callers of the
spatial conv / pool path.  Plain PyTorch module graphs; the spatial stages instantiate
torchgems.spatial layers."""
