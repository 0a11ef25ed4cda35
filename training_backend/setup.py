"""Builds the `ipc_service` torch extension in place (reference: training_backend/setup.py)."""
import os

from setuptools import setup
from torch.utils.cpp_extension import BuildExtension, CUDAExtension

os.environ.setdefault("CUDA_HOME", "/usr/local/cuda")
os.environ["CC"] = "/usr/bin/gcc"
os.environ["CXX"] = "/usr/bin/g++"
setup(
    name="ipcservice",
    ext_modules=[CUDAExtension("ipc_service", ["ipc_service.cpp"], extra_compile_args={"cxx": ["-O2"]},
                               libraries=["rt"])],
    cmdclass={"build_ext": BuildExtension.with_options(use_ninja=False)},
)
