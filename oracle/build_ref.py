"""Build the UNMODIFIED reference CPU extension into oracle/_ref/ (test infrastructure only).

This is checker infrastructure, not product code: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may load what it produces.

Recipe = the reference's own `compile.sh -s CPU` flags (cpp_src/compile.sh:110-116):
sources cpu/*.cpp + tensor/{cpu_tensor,integral,bind}.cpp, `-O3 -fopenmp -std=c++17 -UGPU`.
`compile.sh` itself cannot be run (hostname switch raises NotImplementedError,
compile.sh:18-73), so the same sources/flags are handed to torch.utils.cpp_extension.load.

MAX_SORB_LEN is a compile-time constant in cpp_src/common/default.h:3 that is included by
relative path, so it cannot be overridden with -D/-I.  The build therefore works on a
throw-away copy of cpp_src under a temp dir (never inside this repo), rewrites that one
#define for L = 1, 2, 3, and keeps only the resulting shared objects:

    oracle/_ref/C_extension_L1.so, _L2.so, _L3.so      (git-ignored, travel with gpurun)

No reference source is copied into the repository.
"""
from __future__ import annotations

import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("PYNQS_REFERENCE_ROOT", "/root/reference")

SOURCES = [
    "cpu/hamiltonian.cpp",
    "cpu/onstate.cpp",
    "cpu/excitation.cpp",
    "tensor/cpu_tensor.cpp",
    "tensor/integral.cpp",
    "tensor/bind.cpp",
]


def ref_so_path(L: int) -> str:
    return os.path.join(REF_DIR, f"C_extension_L{L}.so")


def available(L: int = 1) -> bool:
    return os.path.exists(ref_so_path(L))


def build_one(L: int, verbose: bool = False) -> str:
    from torch.utils import cpp_extension

    src_root = os.path.join(REFERENCE_ROOT, "cpp_src")
    if not os.path.isdir(src_root):
        raise FileNotFoundError(f"{src_root} not present (reference is only mounted in the build container)")
    os.makedirs(REF_DIR, exist_ok=True)
    name = f"C_extension_L{L}"
    with tempfile.TemporaryDirectory(prefix="pynqs_ref_") as tmp:
        work = os.path.join(tmp, "cpp_src")
        shutil.copytree(src_root, work)
        hdr = os.path.join(work, "common", "default.h")
        os.chmod(hdr, 0o644)
        txt = open(hdr).read()
        txt, nsub = re.subn(r"#define MAX_SORB_LEN \d+", f"#define MAX_SORB_LEN {L}", txt, count=1)
        assert nsub == 1, "MAX_SORB_LEN define not found"
        open(hdr, "w").write(txt)
        build_dir = os.path.join(tmp, "build")
        os.makedirs(build_dir)
        cpp_extension.load(
            name=name,
            sources=[os.path.join(work, s) for s in SOURCES],
            extra_include_paths=[work],
            extra_cflags=["-O3", "-fopenmp", "-std=c++17", "-UGPU", "-w"],
            extra_ldflags=["-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lgomp"],
            build_directory=build_dir,
            verbose=verbose,
            is_python_module=False,
        )
        shutil.copy(os.path.join(build_dir, name + ".so"), ref_so_path(L))
    return ref_so_path(L)


def cuda_so_path(L: int) -> str:
    return os.path.join(REF_DIR, f"C_extension_cuda_L{L}.so")


def cuda_available(L: int = 1) -> bool:
    return os.path.exists(cuda_so_path(L))


def build_cuda_one(L: int, verbose: bool = False) -> str:
    """The UNMODIFIED reference CPU+GPU extension (`compile.sh -s GPU`, cpp_src/compile.sh:121-159:
    */*.cpp + */*.cu minus magma, cxx `-O3 -fopenmp -std=c++17 -DGPU=1`, nvcc `-O3 -dc
    --expt-relaxed-constexpr`, device link) cross-compiled for sm_100a.  Second comparator of the parity
    tests and of bench.py (GPU-vs-GPU); never on the product path."""
    import glob
    import subprocess
    import sysconfig

    import torch
    from torch.utils import cpp_extension

    src_root = os.path.join(REFERENCE_ROOT, "cpp_src")
    if not os.path.isdir(src_root):
        raise FileNotFoundError(f"{src_root} not present (reference is only mounted in the build container)")
    os.makedirs(REF_DIR, exist_ok=True)
    name = f"C_extension_cuda_L{L}"
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cuda_home = os.path.dirname(os.path.dirname(nvcc))
    with tempfile.TemporaryDirectory(prefix="pynqs_refcuda_") as tmp:
        work = os.path.join(tmp, "cpp_src")
        shutil.copytree(src_root, work)
        hdr = os.path.join(work, "common", "default.h")
        os.chmod(hdr, 0o644)
        txt = open(hdr).read()
        txt, nsub = re.subn(r"#define MAX_SORB_LEN \d+", f"#define MAX_SORB_LEN {L}", txt, count=1)
        assert nsub == 1, "MAX_SORB_LEN define not found"
        open(hdr, "w").write(txt)
        cpps = [f for f in glob.glob(os.path.join(work, "*", "*.cpp")) if "magma" not in f and os.sep + "test" + os.sep not in f]
        cus = [f for f in glob.glob(os.path.join(work, "*", "*.cu")) if "magma" not in f and os.sep + "test" + os.sep not in f]
        inc = [f"-I{work}"] + [f"-I{p}" for p in cpp_extension.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
        abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
        common = [f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H", abi]
        arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
        procs, objs = [], []
        for f in cpps:
            o = os.path.join(tmp, os.path.basename(f) + ".o")
            objs.append(o)
            procs.append(subprocess.Popen(["g++", "-O3", "-fopenmp", "-std=c++17", "-DGPU=1", "-fPIC", "-w", *common, *inc, "-c", f, "-o", o]))
        for f in cus:
            o = os.path.join(tmp, os.path.basename(f) + ".o")
            objs.append(o)
            procs.append(subprocess.Popen([nvcc, "-O3", "-dc", "--expt-relaxed-constexpr", "-std=c++17", "-w", *arch, "-Xcompiler", "-fPIC",
                                           *common, *inc, "-c", f, "-o", o]))
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("reference CUDA build failed")
        dlink = os.path.join(tmp, "dlink.o")
        subprocess.check_call([nvcc, "-dlink", *arch, "-Xcompiler", "-fPIC", *[o for o in objs if o.endswith(".cu.o")], "-o", dlink])
        tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
        subprocess.check_call(["g++", "-shared", "-o", cuda_so_path(L), *objs, dlink, f"-L{tlib}", f"-L{cuda_home}/lib64",
                               "-L/usr/lib/gcc/x86_64-linux-gnu/13", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
                               "-lcudart", "-lcudadevrt", "-lcurand", "-lgomp", f"-Wl,-rpath,{tlib}"])
    return cuda_so_path(L)


def build_all(force: bool = False, verbose: bool = False, cuda: bool = True) -> None:
    for L in (1, 2, 3):
        if force or not available(L):
            print(f"[oracle/_ref] building reference CPU extension, MAX_SORB_LEN={L}", flush=True)
            build_one(L, verbose=verbose)
    if cuda:
        for L in (1, 2):  # GPU-vs-GPU comparator: Fe2S2 / N2 (L = 1) and H50 (L = 2)
            if force or not cuda_available(L):
                print(f"[oracle/_ref] building reference CUDA extension (sm_100a), MAX_SORB_LEN={L}", flush=True)
                build_cuda_one(L, verbose=verbose)


def load_ref(L: int = 1, cuda: bool = False):
    """Import the prebuilt reference extension for MAX_SORB_LEN = L (pybind11 module);
    cuda=True: the CPU+GPU build (needs a CUDA device to do anything useful)."""
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    name = f"C_extension_cuda_L{L}" if cuda else f"C_extension_L{L}"
    if name in sys.modules:
        return sys.modules[name]
    path = cuda_so_path(L) if cuda else ref_so_path(L)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference is mounted")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    for L in (1, 2, 3):
        m = load_ref(L)
        print(L, m.MAX_SORB, m.MAX_SORB_LEN, m.MAX_NELE)
