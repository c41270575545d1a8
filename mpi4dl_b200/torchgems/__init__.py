"""Host-side mirror of the reference's `torchgems` package (hot-path modules only)."""
