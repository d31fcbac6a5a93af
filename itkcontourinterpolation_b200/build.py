"""Builds libmci_b200.so (CUDA kernels + C ABI + host scheduler) in-tree for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so lives in itkcontourinterpolation_b200/_build/
(git-ignored, but shipped to the GPU box by gpurun).
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "_build")
LIB = os.path.join(OUT, "libmci_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"] + os.environ.get("MCI_NVCC_EXTRA", "").split()
SOURCES = ["mci_kernels.cu", "mci_cc.cu", "mci_dt.cu", "mci_dil.cu", "mci_median.cu", "mci_cabi.cu", "mci_filter.cpp"]
HEADERS = ["mci_types.h", "mci_kernels.cuh", "mci_launch.h", os.path.join("..", "..", "include", "mci_b200.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile and link; returns the path of the shared library.  No-op when up to date, and when the
    sources are absent but a prebuilt library is present."""
    os.makedirs(OUT, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    if not all(os.path.exists(s) for s in srcs):
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("libmci_b200 sources missing and no prebuilt library present")
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    for src in srcs:
        obj = os.path.join(OUT, os.path.splitext(os.path.basename(src))[0] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


DRIVER = os.path.join(OUT, "mci_filter_test")


def build_cpp_driver(verbose=False):
    """g++ build of tests/cpp/mci_filter_test.cpp: the C++ filter class of include/mci_b200_filter.hpp driven the
    way the reference's test driver drives itk::MorphologicalContourInterpolator.  Lives next to the library (rpath
    $ORIGIN) so that it travels to the GPU box."""
    lib = build(verbose=verbose)
    root = os.path.join(_HERE, "..")
    src = os.path.join(root, "tests", "cpp", "mci_filter_test.cpp")
    hdrs = [os.path.join(root, "include", "mci_b200_filter.hpp"), os.path.join(root, "include", "mci_b200_io.hpp"),
            os.path.join(root, "include", "mci_b200.h"), lib]
    if not os.path.exists(src):
        if os.path.exists(DRIVER):
            return DRIVER
        raise RuntimeError("mci_filter_test.cpp missing and no prebuilt driver present")
    if _newer(DRIVER, [src] + hdrs):
        cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", os.path.join(root, "include"), src,
               "-o", DRIVER, "-L", OUT, "-lmci_b200", "-lz", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return DRIVER


if __name__ == "__main__":
    print(build(verbose=True))
    print(build_cpp_driver(verbose=True))
