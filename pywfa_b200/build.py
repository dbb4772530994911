"""In-tree build of the C-ABI CUDA library ``pywfa_b200/libwfagpu.so`` (sm_100a only).

``python -m pywfa_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.
The built ``.so`` stays in-tree (git-ignored) so that it travels with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libwfagpu.so")
OBJ = os.path.join(HERE, "csrc", "build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets"]
CXX_FLAGS = ["-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-Wall"]


def _cuda_home() -> str:
    for c in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if c and os.path.exists(os.path.join(c, "bin", "nvcc")):
            return c
    nv = shutil.which("nvcc")
    if nv:
        return os.path.dirname(os.path.dirname(nv))
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    cuda = _cuda_home()
    nvcc = os.path.join(cuda, "bin", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "wfagpu.h"))
    jobs = [
        ("wfa_kernels.cu", [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])),
        ("wfa_pack.cu", [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])),
        ("wfa_reg_bytes.cu", [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])),
        ("wfa_vec_bytes.cu", [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])),
        ("wfagpu_api.cpp", ["g++"] + CXX_FLAGS + ["-I" + os.path.join(cuda, "include")]),
        ("pack.cpp", ["g++"] + CXX_FLAGS),
    ]
    objs, procs = [], []
    for src, cmd in jobs:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            procs.append((src, subprocess.Popen(cmd + ["-c", s, "-o", o], stdout=subprocess.PIPE,
                                                stderr=subprocess.STDOUT, text=True)))
    rebuilt = bool(procs)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"compiling {src} failed:\n{out}")
        if verbose and out:
            print(out)
    if rebuilt or not os.path.exists(OUT):
        r = subprocess.run([nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", OUT] + objs + ["-lpthread"],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"linking libwfagpu.so failed:\n{r.stdout}")
    return OUT


def build_cython(force: bool = False) -> str:
    """Compile the Cython host layer (``pywfa_b200/cy/align_cy.pyx`` over ``cy/wfagpu.pxd``) against
    ``libwfagpu.so`` -- the binding INTEGRATION.md describes, as a real extension module."""
    import sysconfig
    build_library()
    cy = os.path.join(HERE, "cy")
    pyx, pxd = os.path.join(cy, "align_cy.pyx"), os.path.join(cy, "wfagpu.pxd")
    out = os.path.join(cy, "align_cy" + sysconfig.get_config_var("EXT_SUFFIX"))
    csrc = os.path.join(cy, "build", "align_cy.c")
    os.makedirs(os.path.dirname(csrc), exist_ok=True)
    if force or _stale(out, [pyx, pxd, os.path.join(HERE, "..", "include", "wfagpu.h"), OUT]):
        from Cython.Compiler.Main import CompilationOptions, compile as cy_compile
        res = cy_compile(pyx, CompilationOptions(output_file=csrc, include_path=[cy], language_level=3))
        if res.num_errors:
            raise RuntimeError("cythonizing align_cy.pyx failed")
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-I" + sysconfig.get_paths()["include"],
               "-I" + os.path.join(HERE, "..", "include"), csrc, "-o", out,
               "-L" + HERE, "-lwfagpu", "-Wl,-rpath," + HERE]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"compiling the Cython host layer failed:\n{r.stdout}")
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--cython" in sys.argv:
        print(build_cython(force="--force" in sys.argv))
