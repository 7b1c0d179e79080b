"""Prototypes of the C-ABI groups beyond the leaf kernels (filled in as those groups land)."""
from __future__ import annotations


def declare(L) -> None:
    pass
