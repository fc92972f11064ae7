"""Build the reference's own torch_hash CUDA op (unmodified sources, read where they lie under
/root/reference) into oracle/_ref/ so it can serve as the GPU oracle / kernel-to-beat on the B200 box.

TEST INFRASTRUCTURE ONLY: nothing under pcseqlearning_b200/ may import this.

Sources compiled (never copied):
  /root/reference/pcdet/ops/torch_hash/src/torch_hash_api.cpp
  /root/reference/pcdet/ops/torch_hash/src/torch_hash_kernel.cu
Output: oracle/_ref/torch_hash_cuda_ref.so  (git-ignored; travels to the GPU box with the snapshot)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pcdet/ops/torch_hash/src"
OUT = os.path.join(HERE, "_ref")
NAME = "torch_hash_cuda_ref"


def ref_so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    """Compile the reference op for sm_100a. Returns the .so path, or None when the sources are absent."""
    if not os.path.isdir(REF_SRC):
        return ref_so_path() if os.path.exists(ref_so_path()) else None
    if os.path.exists(ref_so_path()):
        return ref_so_path()
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    load(
        name=NAME,
        sources=[os.path.join(REF_SRC, "torch_hash_api.cpp"), os.path.join(REF_SRC, "torch_hash_kernel.cu")],
        extra_include_paths=[REF_SRC],
        extra_cflags=["-O2", "-w"],
        extra_cuda_cflags=["-O2", "-w"],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=True,
    )
    return ref_so_path()


def load_ref():
    """Import the prebuilt reference op (GPU box: only the prebuilt .so is used). Returns module or None."""
    path = ref_so_path()
    if not os.path.exists(path):
        return None
    import importlib.machinery
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    loader = importlib.machinery.ExtensionFileLoader(NAME, path)
    spec = importlib.util.spec_from_loader(NAME, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("reference torch_hash op:", p)
