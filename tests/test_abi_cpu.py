"""CPU-side checks of the C-ABI boundary: the library loads (dlopen only) and exports every symbol that
include/demethify_b200.h declares; the ctypes mirrors agree with the header's struct layouts."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "demethify_b200.h")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from demethify_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    text = open(HEADER).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(dmf_[a-z0-9_]+)\s*\(", text, flags=re.M))
    assert declared, "no prototypes parsed from the header"
    lib = built.lib()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"symbols declared in the header but not exported: {missing}"
    assert declared == set(built.EXPORTS), "ctypes EXPORTS table and header prototypes differ"
    assert lib.dmf_abi_version() == built.ABI_VERSION == 4


def test_struct_layouts_match_header(built, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "demethify_b200.h"\nint main(){printf("%zu %zu %zu ",'
                   'sizeof(dmf_shape_t),sizeof(dmf_fit_desc_t),sizeof(dmf_fit_state_t));'
                   'printf("%zu\\n",sizeof(dmf_wls_desc_t));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    import ctypes as C
    assert sizes == [C.sizeof(built.Shape), C.sizeof(built.FitDesc), C.sizeof(built.FitState), C.sizeof(built.WlsDesc)]


def test_no_cpu_fallback_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from demethify_b200 import deconvolution as dec
    with pytest.raises(built.DmfError):
        dec.mdwbssmf_deconv(np.zeros((8, 1)), None, np.ones((2, 3)) / 2, np.zeros((8, 3)), np.ones((8, 3)), np.zeros((8, 1)), 1, 1, 1, 1e-3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "demethify_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
