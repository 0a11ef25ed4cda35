"""Builds the `ipc_service` torch extension in place (reference: training_backend/setup.py)."""
import os

from setuptools import setup
from torch.utils.cpp_extension import BuildExtension, CUDAExtension

os.environ.setdefault("CUDA_HOME", "/usr/local/cuda")
os.environ["CC"] = "/usr/bin/gcc"
os.environ["CXX"] = "/usr/bin/g++"
setup(
    name="ipcservice",
    # lg_block_csc (get_next_csc) lives in liblegion_b200.so, found next to this directory at run time
    ext_modules=[CUDAExtension("ipc_service", ["ipc_service.cpp"], extra_compile_args={"cxx": ["-O2"]},
                               libraries=["rt", "legion_b200"],
                               library_dirs=[os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "legion_b200")],
                               extra_link_args=["-Wl,-rpath,$ORIGIN/../legion_b200"])],
    cmdclass={"build_ext": BuildExtension.with_options(use_ninja=False)},
)
