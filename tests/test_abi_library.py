"""The C-ABI shared library loads and exports every symbol include/gate_b200.h declares
(no compute calls: there is no GPU in the CPU test tier), and fails loudly without a device."""
import ctypes
import os
import re

import pytest

from opengate_b200 import abi, core

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gate_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gate_b200_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    path = core.build_library()
    return ctypes.CDLL(path)


def test_header_and_python_export_lists_agree():
    assert declared_symbols() == sorted(core.EXPORTS)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert getattr(lib, name) is not None, name
    assert lib.gate_b200_abi_version() == abi.ABI_VERSION


def test_struct_sizes_match_the_c_layout():
    """ctypes mirrors must have the sizes the C compiler gives the header's structs."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "gate_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(gate_b200_tables),' \
          ' sizeof(gate_b200_source), sizeof(gate_b200_dose_desc), sizeof(gate_b200_phsp_desc),' \
          ' sizeof(gate_b200_phsp_columns), sizeof(gate_b200_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    got = [ctypes.sizeof(t) for t in (abi.Tables, abi.Source, abi.DoseDesc, abi.PhspDesc, abi.PhspColumns, abi.Stats)]
    assert got == sizes


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(core.GateB200Error, match="no CUDA device|CUDA"):
        core.Engine(0, 1)
