"""Build libelo_b200.so in-tree with nvcc for sm_100a (no torch extension machinery, no JIT cache)."""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libelo_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(args):
    nvcc, src, obj, verbose, env = args
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    return src, cmd, proc.returncode, proc.stdout


def build(force=False, verbose=False):
    """Compile every CUDA source under csrc/ (one nvcc process per file, in parallel; an object is rebuilt when its
    source or any header is newer) and link them into csrc/libelo_b200.so.  Returns the library path."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that nvcc must not pick up as host compiler
    env.pop("CC", None)
    env.pop("CXX", None)
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    newest_header = max([os.path.getmtime(h) for h in headers] + [os.path.getmtime(__file__)])
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            jobs.append((nvcc, src, obj, verbose, env))
    for stale in set(glob.glob(os.path.join(objdir, "*.o"))) - set(objs):
        os.remove(stale)
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for src, cmd, rc, out in pool.map(_compile_one, jobs):
            if rc != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
            if verbose:
                print(out)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
    return LIB


TEST_SRC = os.path.join(os.path.dirname(HERE), "tests", "csrc")
TEST_LIB = os.path.join(TEST_SRC, "libelo_b200_test.so")


def build_test_lib(force=False):
    """The test hooks of tests/csrc/ (direct doors onto the tcgen05 primitives, used by tests/test_tc_gpu.py and
    tools/mma_bench.py) as a library of their own: nothing of them is linked into libelo_b200.so."""
    srcs = sorted(glob.glob(os.path.join(TEST_SRC, "*.cu"))) + [os.path.join(CSRC, "elo_common.cu")]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(TEST_SRC, "*.h"))
    if not force and os.path.exists(TEST_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(TEST_LIB) for d in deps):
        return TEST_LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    cmd = [nvcc] + NVCC_FLAGS + ["-o", TEST_LIB] + srcs
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
    return TEST_LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
