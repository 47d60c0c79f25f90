"""Build rtm_gpu_b200/librtm_b200.so (and the rtm_b200 driver executable) in-tree with nvcc.

sm_100a only: `-gencode arch=compute_100a,code=sm_100a -lineinfo`.  Host code is compiled
with -ffp-contract=off so that the FP64/FP32 host arithmetic keeps the reference's rounding.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "librtm_b200.so"
EXE = PKG / "rtm_b200"

CU_SOURCES = ["rtm_engine.cu"]
CXX_SOURCES = ["host_abi.cpp", "rtm_nccl.cpp", "driver.cpp", "host/fd_operator.cpp", "host/model.cpp",
               "host/config.cpp", "host/resample.cpp", "host/segy_io.cpp", "host/poststack.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-pthread", "--threads", "4"]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(d).stat().st_mtime <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = os.environ.get("NVCC", "nvcc")
    srcs = [CSRC / s for s in CU_SOURCES + CXX_SOURCES if (CSRC / s).exists()]
    main = CSRC / "main.cpp"
    deps = srcs + list(CSRC.glob("*.cuh")) + list(CSRC.glob("host/*.h")) + \
        [PKG.parent / "include" / "rtm_b200.h", Path(__file__)]
    if not force and _newer(LIB, deps) and (not main.exists() or _newer(EXE, deps + [main])):
        return LIB
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    for s in srcs:
        o = objdir / (s.name + ".o")
        if force or not _newer(o, deps):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        objs.append(str(o))
    subprocess.check_call([nvcc, "-shared", "-o", str(LIB), *objs, "-lcudart_static", "-ldl",
                           "-lpthread", "-lrt"])
    if main.exists():
        subprocess.check_call([nvcc, *NVCC_FLAGS, "-o", str(EXE), str(main), "-L" + str(PKG),
                               "-lrtm_b200", "-Xlinker", "-rpath=$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
