"""oracle/build_ref.py -- TEST INFRASTRUCTURE.  Compile the REFERENCE's own DCN CUDA
extension, unmodified, from where its sources lie under /root/reference, into
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).

    python oracle/build_ref.py

The sources are the reference's codes/models/archs/dcn/src/deform_conv_cuda.cpp and
deform_conv_cuda_kernel.cu; nothing is copied into the repo.  The reference's own
setup.py is not run (it would write into the read-only tree); this is the same two-file
CUDAExtension (setup.py:12-29, same -D__CUDA_NO_HALF_* flags) built through
torch.utils.cpp_extension with an explicit sm_100a target.

The resulting oracle/_ref/deform_conv_cuda.so is (a) a second, GPU-side oracle for
the product DCN kernel and (b) the "existing kernel to beat" in profiles/.  It can only
be *run* on the GPU box.  On the GPU box /root/reference does not exist: there
``load_ref()`` just imports the prebuilt .so.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = os.path.join(os.environ.get("RVSR_REFERENCE", "/root/reference"),
                       "codes", "models", "archs", "dcn", "src")


def build(verbose=False):
    from torch.utils import cpp_extension
    srcs = [os.path.join(REF_SRC, "deform_conv_cuda.cpp"),
            os.path.join(REF_SRC, "deform_conv_cuda_kernel.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return None
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "deform_conv_cuda.so")
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    cpp_extension.load(
        name="deform_conv_cuda", sources=srcs, build_directory=OUT, verbose=verbose,
        extra_cuda_cflags=["-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                           "-D__CUDA_NO_HALF2_OPERATORS__", "-lineinfo"],
        is_python_module=False)
    return so if os.path.exists(so) else None


def load_ref():
    """Import the prebuilt reference extension (GPU box or here). None if absent."""
    so = os.path.join(OUT, "deform_conv_cuda.so")
    if not os.path.exists(so):
        return None
    import torch  # noqa: F401  (the .so links against libtorch)
    spec = importlib.util.spec_from_file_location("deform_conv_cuda", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("built:", p)
