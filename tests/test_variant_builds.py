"""The kernel variants that round 2 measured and did not ship stay in the tree as build options.
They must keep compiling for sm_100a, and the compiled code must contain what the option promises."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC = "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC) or shutil.which("cuobjdump") is None, reason="needs the CUDA toolkit")
def test_tma_and_packed_deblock_variants_compile(tmp_path):
    obj = str(tmp_path / "kernels_variants.o")
    pkg = os.path.join(ROOT, "arrow-h264_b200")
    subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                    "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(pkg, "csrc"),
                    "-DH264R_DEBLOCK_PACKED=1", "-DH264R_INTER_TMA=1", "-c", os.path.join(pkg, "csrc", "kernels.cu"), "-o", obj],
                   check=True, cwd=pkg)
    sass = subprocess.run(["cuobjdump", "-sass", obj], check=True, capture_output=True, text=True).stdout
    assert "deblock4_kernel" in sass and "HSET2" in sass and "HMNMX2" in sass      # the fp16x2 edge filter
    assert "UTMALDG.2D" in sass and "UTMALDG.3D" in sass                           # the TMA window loads
    # and the product build has neither
    prod = subprocess.run(["cuobjdump", "-sass", os.path.join(pkg, "libh264recon.so")], check=True, capture_output=True, text=True).stdout
    assert "deblock4_kernel" not in prod and "UTMALDG" not in prod
