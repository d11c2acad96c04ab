"""Build recipe for the C part of the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

`python -m oracle.cbuild` or `oracle.cbuild.build()` compiles oracle/c/*.c with gcc + OpenMP into
oracle/_build/liboracle_f64.so (git-ignored; it travels to the GPU box with the snapshot and is rebuilt there on
demand: gcc is part of the image).  The reference itself is pure Python with no C sources, so there is no
`oracle/_ref` build: DESIGN.md section 2."""
import ctypes
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "sh_cov_f64.c")
OUT = os.path.join(HERE, "_build", "liboracle_f64.so")
_LIB = None


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: cannot build the C oracle")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [gcc, "-O2", "-fopenmp", "-shared", "-fPIC", "-o", OUT, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
    return OUT


def load():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        sig = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
               ctypes.c_int, ctypes.c_void_p]
        for name in ("mac_oracle_coverage_f64", "mac_oracle_visibility_f64"):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = ctypes.c_int, sig
        _LIB = lib
    return _LIB


if __name__ == "__main__":
    print(build(force=True))
