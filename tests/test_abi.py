"""CPU-side checks of the drop-in boundary: the libraries load, export every declared symbol,
the plan struct layout matches the reference header, and the product fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

import common
from nfft_b200 import cabi, plan_abi as abi

ROOT = common.ROOT
LIBDIR = os.path.join(ROOT, "nfft_b200", "lib")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"

MEMBERS = [f[0] for f in abi.NfftPlanD._fields_]


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not (os.path.exists(os.path.join(LIBDIR, "libnfftcu.so"))
            and os.path.exists(os.path.join(LIBDIR, "libnfft3_b200.so"))):
        subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "nfft_b200", "csrc")], check=True)


def test_nfftcu_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nfftcu.h")).read()
    declared = set(re.findall(r"\b(nfftcu_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared == set(cabi.SYMBOLS), declared ^ set(cabi.SYMBOLS)
    L = C.CDLL(os.path.join(LIBDIR, "libnfftcu.so"))
    for s in declared:
        assert hasattr(L, s), s


def test_host_layer_exports_reference_plan_api():
    C.CDLL(os.path.join(LIBDIR, "libnfftcu.so"), mode=C.RTLD_GLOBAL)
    L = C.CDLL(os.path.join(LIBDIR, "libnfft3_b200.so"))
    for prefix in ("nfft_", "nfftf_"):
        for fn in abi.API_ALL_FUNCS:
            assert hasattr(L, prefix + fn), prefix + fn
        for hook in ("malloc_hook", "free_hook", "die_hook"):
            C.c_void_p.in_dll(L, prefix + hook)
    for prefix in ("solver_", "solverf_"):   # include/nfft3.h:782-786, device-resident (solver_host.c)
        for fn in ("init_advanced_complex", "init_complex", "before_loop_complex", "loop_one_step_complex",
                   "finalize_complex"):
            assert hasattr(L, prefix + fn), prefix + fn
    assert not hasattr(L, "nfftl_trafo")   # long double is deliberately not provided


def _offsets_from_c(include_dirs, header, extra=""):
    src = "#include <stdio.h>\n#include <stddef.h>\n#include <complex.h>\n" + extra
    src += f'#include "{header}"\nint main(void){{\n'
    for typ in ("nfft_plan", "nfftf_plan"):
        src += f'printf("{typ} sizeof %zu\\n", sizeof({typ}));\n'
        for mname in MEMBERS:
            src += f'printf("{typ} {mname} %zu\\n", offsetof({typ}, {mname}));\n'
    src += "return 0;}\n"
    with tempfile.TemporaryDirectory() as td:
        cfile, exe = os.path.join(td, "o.c"), os.path.join(td, "o")
        open(cfile, "w").write(src)
        subprocess.run([GCC, "-std=gnu99", "-w"] + [f"-I{d}" for d in include_dirs] + [cfile, "-o", exe],
                       check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return {(a, b): int(c) for a, b, c in (ln.split() for ln in out.splitlines())}


def _ctypes_offsets():
    out = {}
    for typ, st in (("nfft_plan", abi.NfftPlanD), ("nfftf_plan", abi.NfftPlanF)):
        out[(typ, "sizeof")] = C.sizeof(st)
        for mname in MEMBERS:
            out[(typ, mname)] = getattr(st, mname).offset
    return out


def test_plan_layout_ctypes_matches_our_header():
    assert _offsets_from_c([os.path.join(ROOT, "include")], "nfft3_b200.h") == _ctypes_offsets()


@pytest.mark.skipif(not os.path.exists("/root/reference/include/nfft3.h"),
                    reason="reference header only exists in the build container")
def test_plan_layout_matches_reference_header():
    ref = _offsets_from_c(["/root/reference/include", os.path.join(ROOT, "oracle", "refbuild")], "nfft3.h")
    assert ref == _ctypes_offsets()


SOLVER_MEMBERS = ("mv", "flags", "w", "w_hat", "y", "f_hat_iter", "r_iter", "z_hat_iter", "p_hat_iter", "v_iter",
                  "alpha_iter", "beta_iter", "dot_r_iter", "dot_r_iter_old", "dot_z_hat_iter", "dot_z_hat_iter_old",
                  "dot_p_hat_iter", "dot_v_iter")


def _solver_offsets(include_dirs, header):
    src = '#include <stdio.h>\n#include <stddef.h>\n#include <complex.h>\n#include "%s"\nint main(void){\n' % header
    for typ in ("solver_plan_complex", "solverf_plan_complex"):
        src += f'printf("{typ} sizeof %zu\\n", sizeof({typ}));\n'
        for mname in SOLVER_MEMBERS:
            src += f'printf("{typ} {mname} %zu\\n", offsetof({typ}, {mname}));\n'
    src += "return 0;}\n"
    with tempfile.TemporaryDirectory() as td:
        cfile, exe = os.path.join(td, "o.c"), os.path.join(td, "o")
        open(cfile, "w").write(src)
        subprocess.run([GCC, "-std=gnu99", "-w"] + [f"-I{d}" for d in include_dirs] + [cfile, "-o", exe], check=True)
        return subprocess.run([exe], check=True, capture_output=True, text=True).stdout


@pytest.mark.skipif(not os.path.exists("/root/reference/include/nfft3.h"),
                    reason="reference header only exists in the build container")
def test_solver_plan_layout_matches_reference_header():
    """solver_plan_complex of nfft3_b200.h (device-resident solver) vs include/nfft3.h:760-780."""
    ours = _solver_offsets([os.path.join(ROOT, "include")], "nfft3_b200.h")
    ref = _solver_offsets(["/root/reference/include", os.path.join(ROOT, "oracle", "refbuild")], "nfft3.h")
    assert ours == ref and "solverf_plan_complex dot_v_iter" in ours


def test_fails_loudly_without_gpu():
    """No CPU fallback: without a device nfft_init_guru must die with the library's message."""
    if cabi.lib().nfftcu_device_count() > 0:
        pytest.skip("a GPU is present")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from nfft_b200.plan import Plan\n"
            "Plan.init_guru(1, [32], 10, [64], 4, 0)\n"
            "print('UNREACHABLE')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    assert "UNREACHABLE" not in r.stdout
    assert "no usable CUDA device" in r.stderr
