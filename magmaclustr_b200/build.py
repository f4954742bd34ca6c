"""Build libmagma_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to
the GPU box with the repository snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmagma_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]
UNITS = [("tasks.cu", []), ("tasks3.cu", []), ("estep3.cu", []), ("pairs3.cu", []), ("mstep3.cu", ["-rdc=true"]), ("covbuild.cu", []), ("hostprep.cu", []), ("dense.cu", []), ("dense3.cu", []), ("dense_big.cu", []), ("bigtasks.cu", []), ("api.cu", []), ("lbfgsb_dev.cu", ["-fmad=false", "-rdc=true", "-maxrregcount=128"])]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build_timing():
    """Diagnostic variant: libmagma_b200_timing.so with per-phase clock64() marks in the per-task kernels
    (-DT3_TIMING).  Loaded instead of the product library when MB_LIB points at it (scripts/phase_timing.py)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    build()
    obj = os.path.join(LIBDIR, "tasks3_timing.o")
    subprocess.check_call([nvcc] + ARCH + COMMON + ["-DT3_TIMING", "-c", os.path.join(CSRC, "tasks3.cu"), "-o", obj])
    objs = [obj if name == "tasks3.cu" else os.path.join(LIBDIR, name.replace(".cu", ".o")) for name, _ in UNITS]
    out = os.path.join(LIBDIR, "libmagma_b200_timing.so")
    subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", out] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"])
    return out


def build(verbose=False, force=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "magma_b200.h"))
    objs, cmds = [], []
    for name, extra in UNITS:
        src = os.path.join(CSRC, name)
        obj = os.path.join(LIBDIR, name.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmds.append([nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
    if cmds:   # translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        def run(cmd):
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as ex:
            list(ex.map(run, cmds))
    if force or _newer(objs, LIB):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
