"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/wolfd2_b200.h declares (no compute calls -- there is no GPU here), and the product fails
loudly instead of falling back when no device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "wolfd2_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", txt)
    return sorted({n for n in names if n.startswith("wolfd2_b200_") or n.endswith("_")} - {"reserved_"})


def test_library_builds_and_exports_every_declared_symbol():
    from wolfd2_b200 import build
    lib = C.CDLL(build.build())
    syms = _declared_symbols()
    assert len(syms) >= 30, syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/wolfd2_b200.h but not exported"
    for lit in ("nauxmomentum_", "xmomentum_", "ymomentum_", "alttridlu_", "ppe_", "divergence_", "project_",
                "velboundcond_", "presboundcond_", "veloutflowbcs_", "filter_", "diffmaxnorm_", "dmaxnorm_"):
        assert lit in syms


def test_signature_table_matches_header():
    """Every literal shim in the ctypes table has the same argument count as the header prototype."""
    from wolfd2_b200 import _abi
    txt = open(os.path.join(ROOT, "include", "wolfd2_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for name, (_, kinds) in _abi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\);", txt, flags=re.S)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip()]) == len(kinds), name


def test_header_is_plain_c_and_cxx(tmp_path):
    """The boundary is a C ABI: the header compiles on its own as C99 and as C++17, warnings as errors, and a
    C program can link the library with nothing but the header (no torch, no CUDA headers)."""
    import subprocess
    from wolfd2_b200 import build
    lib = build.build()
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "use.c"
    src.write_text('#include <stdio.h>\n#include "wolfd2_b200.h"\n'
                   'int main(void) { printf("%s %d\\n", wolfd2_b200_version(), wolfd2_b200_config(302, 302, 20, 10)); return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", str(src)])
    cxx = tmp_path / "use.cpp"
    cxx.write_text(src.read_text())
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", str(cxx)])
    exe = tmp_path / "use"
    d = os.path.dirname(lib)
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, "-o", str(exe), str(src), "-L" + d, "-lwolfd2_b200", "-Wl,-rpath," + d])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.split()[-1] == "0", (out.stdout, out.stderr)   # config needs no device


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from wolfd2_b200 import api, deck as dk
    d = dk.cavity(16, re=100.0, dt=0.01)
    with pytest.raises(api.Wolfd2Error, match="no usable CUDA device|CUDA"):
        api.Context(d)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under wolfd2_b200/ may reference it."""
    pkg = os.path.join(ROOT, "wolfd2_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                for bad in ("import oracle", "from oracle", "liboracle", "orc_", "oracle/", "/oracle"):
                    assert bad not in src, f"{os.path.join(dp, f)} references the oracle ({bad!r})"


def test_config_validation():
    from wolfd2_b200 import api
    with pytest.raises(api.Wolfd2Error):
        api.config(2, 2)
    with pytest.raises(api.Wolfd2Error):
        api.config(300, 300, 100, 100)   # mgri*mgrj beyond the table capacity
    api.config(302, 302, 20, 10)         # the reference defaults, include/config.f:28-31
