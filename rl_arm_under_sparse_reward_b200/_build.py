"""Build libbmi_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB_PATH = os.path.join(OUT_DIR, "libbmi_b200.so")
SOURCES = ["capi.cu", "her.cu", "normalizer.cu", "ddpg.cu", "comm.cu", "physics.cu"]


def _site_packages():
    return sysconfig.get_paths()["purelib"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC=/path/to/nvcc")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "bmi.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=True, extra_flags=(), lib_path=None, obj_suffix=""):
    """Compile every CUDA source for sm_100a and link the C-ABI shared library.
    extra_flags / lib_path / obj_suffix build an instrumented variant next to the product library (tools/)."""
    lib_path = lib_path or LIB_PATH
    if not force and lib_path == LIB_PATH and not needs_build():
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    sp = _site_packages()
    nccl_inc = os.path.join(sp, "nvidia", "nccl", "include")
    nccl_lib = os.path.join(sp, "nvidia", "nccl", "lib")
    cublas_lib = os.path.join(sp, "nvidia", "cublas", "lib")
    common = [
        "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC", "-I" + nccl_inc, "-Xptxas", "-v",
    ] + list(extra_flags)
    objs = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            raise RuntimeError("missing CUDA source " + path)
        obj = os.path.join(OUT_DIR, src.replace(".cu", obj_suffix + ".o"))
        objs.append(obj)
        cmd = [nvcc] + common + ["-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (src, out))
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on " + src)
    with open(os.path.join(OUT_DIR, "ptxas%s.log" % obj_suffix), "w") as f:
        f.write("\n".join(log))
    link = [
        nvcc, "-shared", "-o", lib_path] + objs + [
        "-L/usr/local/cuda/lib64", "-lcublasLt", "-L" + nccl_lib, "-l:libnccl.so.2",
        "-Xlinker", "-rpath," + nccl_lib, "-Xlinker", "-rpath," + cublas_lib,
        "-Xlinker", "-rpath,/usr/local/cuda/lib64",
    ]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    if verbose:
        print("built", lib_path)
    return lib_path


if __name__ == "__main__":
    build(force="--force" in sys.argv)
