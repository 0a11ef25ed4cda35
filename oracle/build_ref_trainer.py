"""Builds the REFERENCE's trainer-side extension `ipc_service` from the sources where they lie under
/root/reference/training_backend (ipc_service.cpp, helper_multiprocess.cpp, ipc_cuda_kernel.cu — nothing is copied
into this repository) into oracle/_ref/ref_trainer/.  TEST INFRASTRUCTURE ONLY: tests/test_server_gpu.py runs the
repo's sampling_server against THIS module, i.e. against the unmodified consumer side of the wire
(training_backend/ipc_cuda_kernel.cu:35-235).  Same recipe as the reference's own training_backend/setup.py, with
absolute source paths and the build directory outside the (read-only) reference tree.
usage: python oracle/build_ref_trainer.py [reference_root]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "/root/reference"
SRC = os.path.join(REF, "training_backend")
OUT = os.path.join(HERE, "_ref", "ref_trainer")


def main():
    srcs = [os.path.join(SRC, f) for f in ("ipc_service.cpp", "helper_multiprocess.cpp", "ipc_cuda_kernel.cu")]
    if not all(os.path.exists(s) for s in srcs):
        print("reference tree absent: using prebuilt oracle/_ref/ref_trainer")
        return
    os.makedirs(OUT, exist_ok=True)
    have = [f for f in os.listdir(OUT) if f.startswith("ipc_service") and f.endswith(".so")]
    if have and all(os.path.getmtime(os.path.join(OUT, have[0])) >= os.path.getmtime(s) for s in srcs):
        return
    os.environ.setdefault("CUDA_HOME", "/usr/local/cuda")
    os.environ["CC"], os.environ["CXX"] = "/usr/bin/gcc", "/usr/bin/g++"
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from setuptools import setup
    from torch.utils.cpp_extension import BuildExtension, CUDAExtension
    sys.argv = [sys.argv[0], "-q", "build_ext", "--build-lib", OUT, "--build-temp", "/tmp/legion_ref_trainer_build"]
    setup(name="ipcservice_reference",
          ext_modules=[CUDAExtension("ipc_service", srcs, include_dirs=[SRC], libraries=["rt"])],
          cmdclass={"build_ext": BuildExtension.with_options(use_ninja=False)})


if __name__ == "__main__":
    main()
