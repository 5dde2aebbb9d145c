"""Build the native libraries in-tree.

libesim_b200.so  - CUDA kernels + C ABI (include/esim.h), nvcc, sm_100a only.
libesim_host.so  - host-side population generator / sharding (include/esim_popgen.h), g++.

Both land next to this file so that they travel with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
INCLUDE = ROOT / "include"

CUDA_LIB = PKG / "libesim_b200.so"
HOST_LIB = PKG / "libesim_host.so"

CUDA_SOURCES = ["esim_kernels.cu", "esim_import.cu", "esim_hostcopy.cu", "esim_api.cu", "esim_popgen_device.cu"]
HOST_SOURCES = ["popgen.cpp", "population_io.cpp"]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(d).stat().st_mtime <= t for d in deps)


def _run(cmd):
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(map(str, cmd)), proc.stdout))
    return proc.stdout


def host_compiler() -> str:
    # the image exports CXX=/opt/gcc/bin/g++ (no OpenMP spec file); prefer the distribution compiler
    for cand in ("/usr/bin/g++", shutil.which("g++") or ""):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("no g++ found")


def build_host(force: bool = False) -> Path:
    srcs = [CSRC / s for s in HOST_SOURCES]
    deps = srcs + [INCLUDE / "esim.h", INCLUDE / "esim_popgen.h"] + sorted(CSRC.glob("*.h"))
    if not force and _newer(HOST_LIB, deps):
        return HOST_LIB
    _run([host_compiler(), "-O3", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-Wall", "-I", str(INCLUDE), "-I", str(CSRC), "-o", str(HOST_LIB)]
         + [str(s) for s in srcs])
    return HOST_LIB


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off"]
OBJ_DIR = PKG / "_obj"   # object files (git- and gpurun-ignored): one per source, so that an edit recompiles one file


def _compile_objects(tag: str, defines, force: bool, verbose: bool):
    """One nvcc -c per source, in parallel; an object is reused while it is newer than its source and every header."""
    from concurrent.futures import ThreadPoolExecutor
    OBJ_DIR.mkdir(exist_ok=True)
    headers = [INCLUDE / "esim.h", INCLUDE / "esim_popgen.h", INCLUDE / "esim_popgen_device.h", Path(__file__)] + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh"))
    jobs, objs = [], []
    for name in CUDA_SOURCES:
        src = CSRC / name
        obj = OBJ_DIR / ("%s%s.o" % (Path(name).stem, tag))
        objs.append(obj)
        if force or not _newer(obj, [src] + headers):
            cmd = [nvcc()] + NVCC_FLAGS + ["-ccbin", host_compiler(), "-I", str(INCLUDE), "-I", str(CSRC)] + ["-D" + d for d in defines]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            jobs.append(cmd + ["-c", str(src), "-o", str(obj)])
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        outs = list(ex.map(_run, jobs))
    if verbose:
        print("".join(outs))
    return objs


def _link(objs, out: Path):
    _run([nvcc(), "-shared", "-ccbin", host_compiler(), "-o", str(out)] + [str(o) for o in objs]
         + ["-lcudart", "-ldl", "-L", str(PKG), "-lesim_host", "-Xlinker", "-rpath=$ORIGIN"])


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in CUDA_SOURCES]
    build_host()   # libesim_b200.so links against the host library (sharding for multi-device handles)
    deps = srcs + [INCLUDE / "esim.h", INCLUDE / "esim_popgen.h", INCLUDE / "esim_popgen_device.h", HOST_LIB] + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh"))
    if not force and _newer(CUDA_LIB, deps):
        return CUDA_LIB
    _link(_compile_objects("", [], force, verbose), CUDA_LIB)
    return CUDA_LIB


def build_variant(name: str, defines) -> Path:
    """A/B builds of the CUDA library with other compile-time switches (scripts/kstep_ab.py selects one with ESIM_B200_LIB)."""
    out = PKG / ("libesim_b200_%s.so" % name)
    build_host()
    _link(_compile_objects("_" + name, list(defines), True, False), out)
    return out


DRIVER = ROOT / "drivers" / "esim_run"


def build_driver(force: bool = False) -> Path:
    """drivers/esim_run: the C++ stand-in for the reference's `run --simulate` binary, linked against both libraries."""
    src = ROOT / "drivers" / "esim_run.cpp"
    deps = [src, INCLUDE / "esim.h", INCLUDE / "esim_popgen.h", CUDA_LIB, HOST_LIB]
    if not force and _newer(DRIVER, deps):
        return DRIVER
    _run([host_compiler(), "-O2", "-std=c++17", "-Wall", "-I", str(INCLUDE), str(src), "-o", str(DRIVER),
          "-L", str(PKG), "-lesim_b200", "-lesim_host", "-Wl,-rpath,$ORIGIN/../epidemicsimulator_b200",
          "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-Wl,--allow-shlib-undefined"])
    return DRIVER


def build_all(force: bool = False, verbose: bool = False):
    return build_host(force), build_cuda(force, verbose), build_driver(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", HOST_LIB, CUDA_LIB)
