"""Build the sm_100a shared library (the C-ABI of include/marshmallow.h) in-tree with nvcc.

    python project-marshmallow_b200/build.py [--force] [--ptxas-v]

Output: project-marshmallow_b200/libmarshmallow_b200.so (git-ignored; travels to the GPU box with the
snapshot).  -fmad=false is part of the arithmetic contract of the march's decision path.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("MM_LIB_OUT") or os.path.join(HERE, "libmarshmallow_b200.so")      # MM_LIB_OUT: build a variant (with MM_NVCC_EXTRA) beside the product library
DEMO = os.path.join(HERE, "frame_demo")
CU_SOURCES = ["csrc/capi.cu", "csrc/cloud_march.cu", "csrc/cloud_march_fma.cu", "csrc/curl_noise.cu", "csrc/noise_volumes.cu", "csrc/tonemap.cu", "csrc/reproject.cu", "csrc/post_chain.cu"]
CPP_SOURCES = ["host/sky_camera.cpp"]
HEADERS = ["csrc/common.h", "csrc/cloud_march_x2.inl", "csrc/cloud_march_ray.inl", "csrc/cloud_march.cu", "csrc/post_chain_pixel.h", "csrc/reproject_pixel.h", "csrc/curl_noise_pixel.h", "csrc/curl_table.h", "csrc/noise_volume_pixel.h", "csrc/tonemap_pixel.h", "../include/marshmallow.h", "host/SkyManager.h", "host/Camera.h", "host/uniform_blocks.h", "host/ComputeShader.h", "host/frame_demo.cpp"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
          "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "--cudart", "static"] + os.environ.get("MM_NVCC_EXTRA", "").split()


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, s)) > t for s in CU_SOURCES + CPP_SOURCES + HEADERS + ["build.py"])


def build_library(force=False, verbose=False, ptxas_v=False):
    if not force and not _stale():
        return LIB
    objs = []
    bdir = "build" if not os.environ.get("MM_LIB_OUT") else "build_" + os.path.splitext(os.path.basename(LIB))[0]
    os.makedirs(os.path.join(HERE, bdir), exist_ok=True)
    procs = []
    for src in CU_SOURCES + CPP_SOURCES:
        obj = os.path.join(HERE, bdir, os.path.basename(src) + ".o")
        cmd = [NVCC] + ARCH + CFLAGS + (["-Xptxas", "-v"] if ptxas_v else []) + ["-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or ptxas_v or (verbose and out.strip()):
            print(f"--- {src}\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC] + ARCH + ["-shared", "--cudart", "static", "-o", LIB] + objs
    subprocess.run(cmd, check=True)
    if os.environ.get("MM_LIB_OUT"):
        return LIB
    # the headless C++ host demo (host/frame_demo.cpp) over the C++ mirror classes
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", DEMO, os.path.join(HERE, "host", "frame_demo.cpp"),
                    "-L" + HERE, "-lmarshmallow_b200", "-Wl,-rpath,$ORIGIN"], check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True, ptxas_v="--ptxas-v" in sys.argv))
