"""In-tree build of the CUDA libraries for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA libraries cannot be built")


def _stale(out: str, srcs) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include")) if f.endswith(".h")]
    return d


def build_core(force: bool = False, verbose: bool = False) -> str:
    out = os.path.join(PKG, "libnmpc_b200.so")
    if force or _stale(out, _deps()):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, os.path.join(CSRC, "rti_kernels.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    return out


def build_acados_shims(force: bool = False) -> list:
    """libacados_ocp_solver_{diff2amr,omni4amr,tric3amr}.so + libacados.so: the link line of the
    reference (CMakeLists.txt:108-115)."""
    outs = []
    src = os.path.join(CSRC, "acados_shim.cpp")
    if not os.path.exists(src):
        return outs
    core = build_core(force)
    inc = os.path.join(ROOT, "include")
    cxx = shutil.which("g++") or "g++"
    out = os.path.join(PKG, "libacados.so")
    if force or _stale(out, _deps() + [core]):
        cmd = [cxx, "-std=c++14", "-O2", "-fPIC", "-shared", "-I", inc, "-o", out, src,
               "-L", PKG, "-lnmpc_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    outs.append(out)
    for m in ("diff2amr", "omni4amr", "tric3amr"):
        o = os.path.join(PKG, f"libacados_ocp_solver_{m}.so")
        s = os.path.join(CSRC, f"acados_solver_{m}.cpp")
        if force or _stale(o, _deps() + [out]):
            cmd = [cxx, "-std=c++14", "-O2", "-fPIC", "-shared", "-I", inc, "-o", o, s,
                   "-L", PKG, "-lacados", "-lnmpc_b200", "-Wl,-rpath,$ORIGIN"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
        outs.append(o)
    return outs


def build_alternate(out_dir: str) -> str:
    """libnmpc_b200.so for the horizon / default tables emitted into `out_dir` (emit.py --out-dir): the same sources with
    out_dir/include first on the include path and out_dir/model_defaults.inc as the table"""
    out = os.path.join(out_dir, "libnmpc_b200.so")
    hdr = open(os.path.join(out_dir, "include", "nmpc_horizon.h")).read()
    import re
    n = re.search(r"#define\s+NMPC_N\s+(\d+)", hdr).group(1)
    dt = re.search(r"#define\s+NMPC_DT\s+([0-9.eE+-]+)", hdr).group(1)
    flags = [f"-DNMPC_N={n}", f"-DNMPC_DT={dt}"]
    inc = os.path.join(out_dir, "model_defaults.inc")
    if os.path.exists(inc):
        flags.append(f'-DNMPC_MODEL_DEFAULTS_INC="{os.path.abspath(inc)}"')
    cmd = [_nvcc()] + NVCC_FLAGS + flags + ["-o", out, os.path.join(CSRC, "rti_kernels.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return out


def build_all(force: bool = False) -> None:
    build_core(force)
    build_acados_shims(force)


if __name__ == "__main__":
    build_all(force=True)
