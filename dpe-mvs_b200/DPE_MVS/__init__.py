"""DPE_MVS — drop-in Python surface of the reference package (src/DPE_MVS/__init__.py:6-18):
`dpe_mvs(dense_folder, gpu_index, verbose, fusion, viz, depth, normal, weak, edge) -> int`.

The native module `_dpe` is built in-tree by `dpe-mvs_b200/build.py`; importing this package
without it raises — there is no Python or CPU fallback.  `gpu_index < 0` (or the environment
variable DPE_GPUS=0,1,...) spreads the reference views over several GPUs of the box.
"""
from __future__ import annotations

from ._dpe import dpe_mvs as _native  # type: ignore

__all__ = ["dpe_mvs"]


def dpe_mvs(
    dense_folder: str,
    gpu_index: int = 0,
    verbose: bool = True,
    fusion: bool = False,
    viz: bool = False,
    depth: bool = True,
    normal: bool = False,
    weak: bool = False,
    edge: bool = False,
) -> int:
    """Run the DPE-MVS pipeline on a colmap2mvsnet-layout folder; returns 0."""
    return _native(str(dense_folder), gpu_index, verbose, fusion, viz, depth, normal, weak, edge)
