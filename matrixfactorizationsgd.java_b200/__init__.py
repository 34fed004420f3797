"""matrixfactorizationsgd.java_b200 -- B200-native drop-in for the factorization path of
vbarbosadev/MatrixFactorizationSGD.java.

Layout: csrc/ (sm_100a CUDA kernels + the C ABI of include/mfsgd.h), lib/libmfsgd.so (built in-tree),
host.py (mirror of the reference's entry point), engine.py (handle-level mirror of the C ABI),
ring.py (torch.distributed bootstrap for one-process-per-GPU DSGD), workloads.py (BASELINE configs).
Importing fails loudly when libmfsgd.so is missing -- there is no CPU fallback.
"""
from . import _capi as capi  # noqa: F401  (loads libmfsgd.so or raises)
from ._capi import MfsgdError  # noqa: F401
from .engine import (Engine, apply_updates_forced, device_count, generate_to_host, make_config, measure_ceilings,  # noqa: F401
                     nccl_unique_id, synth_params, synth_params_of)
from .host import EarlyStopResult, Factors, MatrixFactorizationSGD, Model, RatingsFile, read_ratings  # noqa: F401
from .workloads import SEED, WORKLOADS, bytes_per_update  # noqa: F401

__all__ = ["MatrixFactorizationSGD", "Factors", "Engine", "MfsgdError", "make_config", "synth_params", "synth_params_of",
           "device_count", "nccl_unique_id", "apply_updates_forced", "generate_to_host", "WORKLOADS", "SEED",
           "bytes_per_update", "capi", "read_ratings", "RatingsFile"]
