"""CPU: the CUDA extension is built in-tree, loads, and exports every symbol include/*.h declares.
No compute call is made here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "memflow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_all_declared_symbols():
    from memflow_msci_project_b200 import _cabi, build
    build.build()
    L = _cabi.lib()
    names = declared_symbols()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/memflow_b200.h but not exported: {missing}"
    assert L.mf_abi_version() == 1


def test_desc_struct_layout_matches_header():
    """sizeof(mf_rambo_desc) as the C compiler sees it == ctypes mirror."""
    import subprocess
    import tempfile
    from memflow_msci_project_b200.phasespace._desc import RamboDesc
    src = '#include <stdio.h>\n#include "memflow_b200.h"\nint main(){printf("%zu %zu\\n", sizeof(mf_rambo_desc), __builtin_offsetof(mf_rambo_desc, rap_max_cut));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        size, off = map(int, subprocess.check_output([exe]).split())
    assert size == ctypes.sizeof(RamboDesc)
    assert off == RamboDesc.rap_max_cut.offset


def test_product_refuses_cpu_tensors():
    """The product path must fail loudly instead of falling back to a CPU implementation."""
    import torch
    from memflow_msci_project_b200 import _cabi
    from memflow_msci_project_b200.phasespace.rambo_generator import FlatInvertiblePhasespace
    gen = FlatInvertiblePhasespace([0.0, 0.0], torch.tensor([125.25, 172.5, 172.5, 1e-5]), 13000,
                                   pdf_active=True, uniform_x1x2=True, dev=torch.device("cpu"))
    with pytest.raises(_cabi.MemflowB200Error):
        gen.generateKinematics_batch(torch.rand(8, 10))
