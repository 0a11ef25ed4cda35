"""legion_b200 — B200-native implementation of Legion's mini-batch data path.

Layout: csrc/ (CUDA kernels + the C ABI of include/legion_b200.h), capi.py (ctypes binding),
runner.py (host mirror of the reference's GPURunner / UnifiedCache for one GPU), synth.py
(synthetic datasets in Legion's layout).  The C++ server and the trainer-side `ipc_service`
extension live in sampling_server/ and training_backend/.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
