"""TEST INFRASTRUCTURE ONLY - recipe that compiles the REFERENCE's own bulk-KV kernel for sm_100a,
from its sources where they lie under /root/reference, into oracle/_ref/ref_custom_ops.so
(git-ignored; travels to the GPU box with the gpurun snapshot).

The reference's build system (cmake, hard-coded arch list "7.0 ... 9.0+PTX",
csrc/custom_ops/CMakeLists.txt:8) is NOT run; the two translation units are compiled directly
with torch.utils.cpp_extension (ninja + nvcc), ~3 minutes.  No reference source is copied.

The resulting library registers ``arctic_inference::reshape_and_cache_flash_bulk`` - the same op
name as our drop-in - so tests load it in a separate process (tests/test_kv_write_gpu.py).

Run in the build container only:  python oracle/build_ref.py
"""
import glob
import os
import shutil

REF = "/root/reference/csrc/custom_ops"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(REF):
        print("reference sources not present; nothing to do")
        return
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    build_dir = "/tmp/ref_custom_ops_build"
    os.makedirs(build_dir, exist_ok=True)
    load(name="ref_custom_ops", sources=[f"{REF}/kernels.cu", f"{REF}/torch_bindings.cpp"],
         extra_include_paths=[REF], is_python_module=False, build_directory=build_dir,
         extra_cuda_cflags=["-O3"], verbose=True)
    so = glob.glob(os.path.join(build_dir, "ref_custom_ops*.so"))[0]
    shutil.copy(so, os.path.join(OUT, "ref_custom_ops.so"))
    print("built", os.path.join(OUT, "ref_custom_ops.so"))


if __name__ == "__main__":
    main()
