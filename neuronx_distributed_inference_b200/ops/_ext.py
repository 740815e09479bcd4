"""Build / load the in-tree sm_100a extension (``neuronx_distributed_inference_b200/_C.so``).

Built with plain ``nvcc`` (cross-compiles without a GPU) + ninja-free incremental objects under
``build/``; the resulting ``.so`` lives next to the package so it travels with the source tree.
"""
from __future__ import annotations

import glob
import hashlib
import importlib.util
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CSRC = os.path.join(_PKG, "csrc")
_SO = os.path.join(_PKG, "_C.so")
_BUILD = os.path.join(os.path.dirname(_PKG), "build", "nxdi_b200")
_mod = None
_tried = False

GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]


def sources():
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")) + glob.glob(os.path.join(_CSRC, "*.cpp")))


def _include_flags():
    from torch.utils.cpp_extension import include_paths
    inc = include_paths(device_type="cuda") if "device_type" in include_paths.__code__.co_varnames else include_paths(cuda=True)
    inc.append(sysconfig.get_paths()["include"])
    inc.append(_CSRC)
    return [f"-I{p}" for p in inc]


def _digest(path, flags):
    h = hashlib.sha1()
    h.update(" ".join(flags).encode())
    for p in [path] + sorted(glob.glob(os.path.join(_CSRC, "*.cuh")) + glob.glob(os.path.join(_CSRC, "*.h"))):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile every source in csrc/ for sm_100a and link ``_C.so``.  Incremental."""
    os.makedirs(_BUILD, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "--expt-extended-lambda",
              "-diag-suppress", "177", "-diag-suppress", "550", "-DNDEBUG", f"-I{_CSRC}"] + GENCODE
    if verbose:
        common += ["-Xptxas", "-v"]
    objs, jobs = [], []
    for src in sources():
        obj = os.path.join(_BUILD, os.path.basename(src) + ".o")
        stamp = obj + ".sha1"
        dg = _digest(src, common)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dg:
            continue
        if src.endswith(".cpp"):
            cxx = os.environ.get("CXX", "g++")
            cmd = [cxx, "-O2", "-std=c++17", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_EXTENSION_NAME=_C",
                   "-DTORCH_API_INCLUDE_EXTENSION_H", "-DNDEBUG", "-Wno-deprecated-declarations",
                   "-I/usr/local/cuda/include"] + _include_flags() + ["-c", src, "-o", obj]
        else:
            cmd = [nvcc] + common + ["-c", src, "-o", obj]
        jobs.append((cmd, stamp, dg, src))

    def run(job):
        cmd, stamp, dg, src = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dg)
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    need_link = bool(jobs) or not os.path.exists(_SO) or force
    if need_link:
        tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
        cmd = [nvcc, "-shared", "-o", _SO] + objs + [f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu",
                                                    "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
                                                    f"-Xlinker=-rpath,{tlib}"] + GENCODE
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return _SO


def extension_available() -> bool:
    return load_extension(required=False) is not None


def load_extension(required: bool = True):
    global _mod, _tried
    if _mod is not None:
        return _mod
    if not _tried or required:
        _tried = True
        if os.path.exists(_SO):
            spec = importlib.util.spec_from_file_location("neuronx_distributed_inference_b200._C", _SO)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _mod = mod
            return _mod
    if required:
        raise RuntimeError(f"{_SO} not found: build it with __graft_entry__.build()")
    return None
