"""In-tree build of the C++ host layer + pybind11 module `ExpressionMatrix2` (reference API names).

    python -m expressionmatrix2_b200.hostmodule

Produces expressionmatrix2_b200/ExpressionMatrix2<ext-suffix>.so linked against libem2b200.so ($ORIGIN rpath).
Import it with `from expressionmatrix2_b200 import hostmodule; M = hostmodule.load()`.
"""
from __future__ import annotations

import importlib.util
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
HOST = os.path.join(HERE, "host")
SOURCES = ["ExpressionMatrix.cpp", "ExpressionMatrixSubset.cpp", "Lsh.cpp", "SimilarPairs.cpp", "Gpu.cpp", "PythonModule.cpp"]
TARGET = os.path.join(HERE, "ExpressionMatrix2" + sysconfig.get_config_var("EXT_SUFFIX"))
CXX = "/usr/bin/g++"


def _stale() -> bool:
    if not os.path.exists(TARGET):
        return True
    t = os.path.getmtime(TARGET)
    deps = [os.path.join(HOST, f) for f in os.listdir(HOST)] + [os.path.join(HERE, "..", "include", "em2b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False) -> str:
    from . import build as cuda_build
    cuda_build.build()
    if not force and not _stale():
        return TARGET
    import pybind11
    inc = ["-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"]]
    cmd = [CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off", *inc,
           *[os.path.join(HOST, s) for s in SOURCES], "-o", TARGET, "-L" + HERE, "-lem2b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return TARGET


def load():
    """Import the built module as `ExpressionMatrix2` (fails loudly if it was not built)."""
    if not os.path.exists(TARGET):
        raise ImportError(f"{TARGET} is missing: run `python -m expressionmatrix2_b200.hostmodule`")
    if "ExpressionMatrix2" in sys.modules:
        return sys.modules["ExpressionMatrix2"]
    spec = importlib.util.spec_from_file_location("ExpressionMatrix2", TARGET)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ExpressionMatrix2"] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
