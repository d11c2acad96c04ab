"""In-tree build of libmacarons_b200.so (sm_100a only) with plain nvcc.

`python -m macarons_b200.build` or `macarons_b200.build.build()`; the library lands in
`macarons_b200/lib/` so it travels with the repo snapshot to the GPU box.  Sources are rebuilt only
when newer than the library.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libmacarons_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


# extra flags for debug builds, e.g. MAC_EXTRA_NVCC_FLAGS=-DMAC_LINEAR_PROFILE (then run with --force)
NVCC_FLAGS += os.environ.get("MAC_EXTRA_NVCC_FLAGS", "").split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "macarons_b200.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def up_to_date():
    out = lib_path()
    if not os.path.exists(out):
        return False
    t = os.path.getmtime(out)
    return all(os.path.getmtime(d) <= t for d in _deps())


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Returns the library path."""
    out = lib_path()
    if not force and up_to_date():
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build %s" % LIBNAME)
    os.makedirs(LIBDIR, exist_ok=True)
    logs = []
    # one object per source, compiled in parallel; an object is rebuilt when its source, any header of csrc/, the
    # public header or this script is newer than it
    headers = [d for d in _deps() if not d.endswith(".cu")]
    h_time = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(h_time, os.path.getmtime(src)):
            return obj, "", 0
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, "$ " + " ".join(cmd) + "\n" + res.stdout + res.stderr, res.returncode

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    objs = []
    for obj, log, rc in results:
        if log:
            logs.append(log)
        if rc != 0:
            raise RuntimeError("nvcc failed:\n" + log)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    logs.append("$ " + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + logs[-1])
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
