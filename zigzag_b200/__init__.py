"""zigzag_b200 — B200 (sm_100a) implementation of zigzag's per-gene MCMC sweep behind the reference's move API.

The compute path is the CUDA library zigzag_b200/libzigzag_gpu.so (C ABI: include/zigzag_gpu.h); there is no CPU
fallback.  `zigzag` mirrors the reference's `zigzag$new()/burnin()/mcmc()` object (R/zigzag.R, R/zigzagBurnin.R,
R/zigzagMCMC.R) with the per-gene moves running on the GPU.
"""
from . import _capi
from ._capi import Handle, Hyper, ZigzagGpuError

__all__ = ["Handle", "Hyper", "ZigzagGpuError", "_capi"]
