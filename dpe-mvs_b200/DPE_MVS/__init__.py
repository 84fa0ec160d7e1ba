"""DPE_MVS — drop-in Python surface of the reference package (src/DPE_MVS/__init__.py:6-18):
`dpe_mvs(dense_folder, gpu_index, verbose, fusion, viz, depth, normal, weak, edge) -> int`.

The native module `_dpe` is built in-tree by `dpe-mvs_b200/build.py`; importing this package
without it raises — there is no Python or CPU fallback.  `gpu_index < 0` (or the environment
variable DPE_GPUS=0,1,...) spreads the reference views over several GPUs of the box.
"""
from __future__ import annotations

from ._dpe import dpe_mvs as _native  # type: ignore

__all__ = ["dpe_mvs"]


def _nccl_hint():
    """Multi-GPU runs bind NCCL at run time (csrc/dpe_capi.cu: NcclApi).  When this interpreter has a pip-installed
    NCCL (PyTorch's, newer than the system's, same soname), point the library at that file so that a later
    `import torch` in the same process finds the NCCL it was built against already loaded."""
    import importlib.util
    import os
    if os.environ.get("DPE_NCCL_LIB"):
        return
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["DPE_NCCL_LIB"] = cand
                return
    except Exception:
        pass


_nccl_hint()


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
