"""The reference-side artefacts under integration/ are real: the pybind11 binding compiles against
include/gate_b200.h, links libgate_b200.so, imports, and (CPU tier) fails loudly without a device; the
managers.py patch applies to the reference tree when it is present."""
import os
import subprocess
import sys
import sysconfig
import textwrap

import pytest

from opengate_b200 import abi, core

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pybind_binding_builds_links_and_loads(tmp_path):
    pybind11 = pytest.importorskip("pybind11")
    lib = core.build_library()
    wrapper = tmp_path / "module.cpp"
    wrapper.write_text(textwrap.dedent("""
        #include <pybind11/pybind11.h>
        namespace py = pybind11;
        void init_GateB200Engine(py::module &m);
        PYBIND11_MODULE(gate_b200_pybind_check, m) { init_GateB200Engine(m); }
    """))
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = tmp_path / ("gate_b200_pybind_check" + ext)
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", pybind11.get_include(),
           "-I", sysconfig.get_paths()["include"], str(wrapper), os.path.join(ROOT, "integration", "pyGateB200Engine.cpp"),
           lib, "-Wl,-rpath," + os.path.dirname(lib), "-o", str(out)]
    subprocess.check_call(cmd)
    code = textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {str(tmp_path)!r})
        import gate_b200_pybind_check as m
        assert m.GateB200AbiVersion() == {abi.ABI_VERSION}
        try:
            e = m.GateB200Engine(0, 123)
        except RuntimeError as ex:
            print("RAISED", ex)
        else:
            e.SetWorld([1000.0, 1000.0, 1000.0], 0) if False else None
            print("ENGINE", e.GetHandleAddress() != 0)
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    import torch
    if torch.cuda.is_available():
        assert "ENGINE True" in r.stdout
    else:
        assert "RAISED" in r.stdout and "no CUDA device" in r.stdout      # no CPU fallback through this door either


def test_engine_switch_patch_applies_to_the_reference_tree(tmp_path):
    ref = "/root/reference/opengate/managers.py"
    if not os.path.exists(ref):
        pytest.skip("the reference tree is not on this machine")
    work = tmp_path / "opengate"
    work.mkdir()
    (work / "managers.py").write_text(open(ref).read())
    subprocess.check_call(["patch", "-p1", "-s", "-d", str(tmp_path), "-i", os.path.join(ROOT, "integration", "opengate_engine_switch.patch")])
    txt = (work / "managers.py").read_text()
    assert '"engine": (' in txt and "GpuSimulationEngine" in txt
    compile(txt, "managers.py", "exec")
