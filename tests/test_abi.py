"""The C-ABI library loads and exports every symbol include/scvx_b200.h declares; the product never
touches the oracle; the device path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

from successiveconvexification_b200 import _lib
from successiveconvexification_b200.defns import CProbInfo

HEADER = os.path.join(ROOT, "include", "scvx_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scvx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"


def test_exported_symbols_are_plain_c():
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    scvx = [s for s in exported if s.startswith("scvx_")]
    assert sorted(scvx) == _declared_symbols()


def test_probinfo_layout():
    lib = _lib.load()
    assert lib.scvx_sizeof_probinfo() == ctypes.sizeof(CProbInfo) == 248
    assert lib.scvx_version() >= 1000
    from successiveconvexification_b200.defns import CDimProblem
    assert ctypes.sizeof(CDimProblem) == 216          # static_assert in scvx_api.cu


def test_julia_shim_binds_every_symbol():
    shim = open(os.path.join(ROOT, "julia", "SCvxB200.jl")).read()
    for name in _declared_symbols():
        assert f":{name}" in shim, f"julia shim does not ccall {name}"


def test_product_never_uses_the_oracle():
    pkg = os.path.join(ROOT, "successiveconvexification_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"


def test_fails_loudly_without_gpu():
    lib = _lib.load()
    if lib.scvx_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from successiveconvexification_b200.dynamics import DeviceContext
    with pytest.raises(_lib.ScvxError, match="no CPU fallback"):
        DeviceContext()


def test_julia_shim_overrides_sit_inside_the_module():
    """The replaced `Dynamics.*` methods use LinPoint / LinRes / IntegratorCache unqualified; master.jl exports them from
    RocketlandDefns but never brings them into Main, so the definitions must sit inside `module SCvxB200` (which does
    `using ..RocketlandDefns` and `import ..Dynamics`).  The shim has never run under Julia; this pins its structure."""
    shim = open(os.path.join(ROOT, "julia", "SCvxB200.jl")).read()
    start, end = shim.index("module SCvxB200"), shim.rindex("end # module")
    assert shim[end:].strip() == "end # module"
    assert "using ..RocketlandDefns" in shim[start:end] and "import ..Dynamics" in shim[start:end]
    for name in ("linearize_dynamics", "predict_state", "simulate_zygote", "sensitivity_zygote"):
        assert start < shim.index(f"function Dynamics.{name}(") < end
    code = "\n".join(l.split("#")[0] for l in shim[start:end].splitlines())
    assert "SCvxB200." not in code                      # no self-qualified names inside the module
    assert "live_mode" in code and "MODE_TEXTBOOK" in code
