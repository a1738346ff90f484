"""The C-ABI library loads on a machine without a GPU, exports every symbol
include/mppi_b200.h declares, agrees with the ctypes mirror of MppiConfig, and
refuses to create a handle when there is no device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mppi_b200.h")

from mppi_playground_b200 import _capi  # noqa: E402


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mppi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in mppi_b200.h but not exported"
        assert n in _capi.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_capi.PROTOTYPES) == names
    assert lib.mppi_abi_version() == _capi.ABI_VERSION


@pytest.mark.parametrize("name", ["MppiConfig", "MppiStepEpilogue", "MppiFp32Report"])
def test_struct_layouts_match_header(tmp_path, name):
    mirror = getattr(_capi, name)
    fields = [f[0] for f in mirror._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
            f'printf("%zu\\n", sizeof({name}));']
    prog += [f'printf("%zu\\n", offsetof({name}, {f}));' for f in fields]
    prog += ["return 0;}"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-o", str(exe), str(c)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(mirror)
    for f, off in zip(fields, out[1:]):
        assert getattr(mirror, f).offset == off, f


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    lib = _capi.load()
    out = (C.c_uint32 * 4)()
    lib.mppi_philox4x32_10((C.c_uint32 * 4)(0, 0, 0, 0), (C.c_uint32 * 2)(0, 0), out)
    assert list(out) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    lib.mppi_philox4x32_10((C.c_uint32 * 4)(f, f, f, f), (C.c_uint32 * 2)(f, f), out)
    assert list(out) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    lib.mppi_philox4x32_10((C.c_uint32 * 4)(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344),
                           (C.c_uint32 * 2)(0xA4093822, 0x299F31D0), out)
    assert list(out) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    lib = _capi.load()
    cfg = _capi.MppiConfig()
    cfg.abi_version, cfg.model, cfg.horizon, cfg.num_samples = _capi.ABI_VERSION, _capi.MODEL_PENDULUM, 10, 64
    cfg.dim_state, cfg.dim_control, cfg.lambda_ = 2, 1, 1.0
    h = C.c_void_p()
    rc = lib.mppi_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.mppi_last_error() or b"CUDA" in lib.mppi_last_error()
    import mppi_playground_b200 as m

    pm = m.PendulumModel()
    with pytest.raises(RuntimeError, match="no CPU"):
        m.MPPI(10, 64, 2, 1, pm.dynamics, pm.cost_func, torch.tensor([-2.0]), torch.tensor([2.0]),
               torch.tensor([1.0]), 1.0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under mppi_playground_b200/ may reference it."""
    pkg = os.path.join(ROOT, "mppi_playground_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(d, f)
    code = "import sys; import mppi_playground_b200; assert not any(m.split('.')[0]=='oracle' for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_pass1_loop_issue_budget_does_not_regress():
    """Static guard on the headline kernel's hot loop (profiles/sass_loop_budget.py): instructions per
    sample-timestep, and no local-memory traffic or constant-bank reloads inside it."""
    import importlib.util
    import shutil

    lib = os.path.join(ROOT, "mppi_playground_b200", "libmppi_b200.so")
    if not os.path.exists(lib) or shutil.which("cuobjdump") is None:
        pytest.skip("needs the built library and cuobjdump")
    spec = importlib.util.spec_from_file_location("sass_loop_budget", os.path.join(ROOT, "profiles", "sass_loop_budget.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    body, _ = mod.loop_body(lib)
    ops = [t.split()[0].split(".")[0] for t in body]
    per = len(body) / mod.SAMPLE_STEPS_PER_ITER
    assert per <= 140, f"{per} SASS instructions per sample-timestep (round 1: 184 scalar; round 2 paired loop: 138.5)"
    # ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 regardless of -fmad=false, which would break the
    # reference's separate roundings: every packed add must be in the uncontractable fma-by-one form (P2 operator+)
    assert ops.count("FADD2") == 0, "a plain packed add in the loop: ptxas may have fused a product into it"
    packed = sum(ops.count(k) for k in ("FFMA2", "FMUL2", "FADD2"))
    assert packed >= 120, f"only {packed} packed fp32 instructions (FFMA2/FMUL2/FADD2) in the paired loop"
    assert not {"LDL", "STL", "LD", "ST"} & set(ops), "local/generic memory traffic inside the pass-1 loop"
    assert ops.count("LDCU") + ops.count("LDC") <= 2, "loop constants are being re-read from the constant bank"
