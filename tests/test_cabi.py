"""The C ABI on a CPU-only box: the library loads, exports every symbol include/sdeb200.h declares, argument
structs have the declared layout, NVRTC compiles emitted sources for sm_100a without a GPU, and invalid
arguments are rejected with negative status codes (no compute calls)."""
import ctypes
import os
import re

import pytest

from stochastic_b200 import runtime, workloads
from stochastic_b200.sde_solver import build_source
from stochastic_b200.vectorized_I_generation import wiener_source

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "sdeb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdeb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = runtime.lib()
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdeb200.h but not exported"
    assert set(names) == set(runtime.EXPORTS)
    assert lib.sdeb200_abi_version() == 4


def test_nvrtc_is_bound_independently_of_the_host_process():
    import torch  # noqa: F401  (its wheel ships an older NVRTC under the same soname)
    major, minor, path = runtime.nvrtc_version()
    assert major >= 12 and path
    if os.path.exists("/usr/local/cuda/lib64/libnvrtc.so.12") and not os.environ.get("SDEB200_NVRTC"):
        assert path.startswith(os.environ.get("CUDA_HOME", "/usr/local/cuda"))


def test_struct_layouts():
    assert ctypes.sizeof(runtime.SolveArgs) == 120 and runtime.SolveArgs.out_hit.offset == 112
    assert runtime.SolveArgs.n_traj_total.offset == 24 and runtime.SolveArgs.x0.offset == 48
    assert runtime.SolveArgs.dt.offset == 64 and runtime.SolveArgs.moments.offset == 104
    assert ctypes.sizeof(runtime.WienerArgs) == 56


def test_nvrtc_compiles_for_sm100a_without_gpu():
    prob = workloads.gbm(f64=True)
    src = build_source(prob.a, prob.b, 1, "wagner_platen", True, "original")[0]
    key, cubin = runtime.compile_source(src, use_cache=False)
    assert cubin[:4] == b"\x7fELF" and b"sde_solve_kernel" in cubin
    src, block, smem = wiener_source(2, "wagner_platen", 10, True, "original")
    key, cubin = runtime.compile_source(src, use_cache=False)
    assert b"sde_wiener_kernel" in cubin and smem == (2 * 2 * 10 + 3 * 2) * block * 8


def test_nvrtc_error_is_reported():
    with pytest.raises(runtime.Sdeb200Error) as e:
        runtime.compile_source("this is not CUDA", use_cache=False)
    assert "NVRTC" in str(e.value) or "nvrtc" in str(e.value)


def test_invalid_arguments_are_rejected():
    lib = runtime.lib()
    assert lib.sdeb200_solve(None, None, 128, 0, 0, None) < 0
    assert b"null" in lib.sdeb200_last_error()
    assert lib.sdeb200_wiener_integrals(None, None, 128, 0, None) < 0
    assert lib.sdeb200_keys_split(0, 0, 10, 8, 4, 0, None, None) < 0          # out NULL
    assert lib.sdeb200_keys_split(0, 0, 10, 8, 4, 0, ctypes.c_void_p(8), None) < 0   # range exceeds n_total
    assert lib.sdeb200_normal(0, 0, 4, 1, 7, ctypes.c_void_p(8), None) < 0     # unknown layout
    assert lib.sdeb200_program_load(None, 0, None, None) < 0
    assert lib.sdeb200_compile(None, None, None, None, None, None) < 0


def test_stage_count_formula():
    assert runtime.stage_count(2, 10, "wagner_platen") == 46
    assert runtime.stage_count(1, 10, "wagner_platen") == 0
    assert runtime.stage_count(12, 10, "euler") == 12 and runtime.stage_count(2, 10, "milstein") == 0
    assert runtime.pick_block(2, 10, "wagner_platen", True) == 256
    assert runtime.pick_block(6, 10, "wagner_platen", True) * runtime.stage_count(6, 10, "wagner_platen") * 8 <= 200 * 1024


def test_plain_c_client_builds_against_the_header_and_compiles_a_kernel(tmp_path):
    """examples/cabi_client.c: the boundary used from C with no Python in the process.  Here (no GPU) it must link
    against libsdeb200.so, report the bound NVRTC and NVRTC-compile its hand-written step function for sm_100a."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h"):
        pytest.skip("gcc / CUDA runtime headers not available")
    runtime.lib()                                  # fails loudly if the library has not been built
    exe = str(tmp_path / "cabi_client")
    libdir = os.path.join(ROOT, "stochastic_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", os.path.join(ROOT, "examples", "cabi_client.c"),
                           "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include", "-L" + libdir, "-lsdeb200",
                           "-L/usr/local/cuda/lib64", "-lcudart", "-lm", "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe, os.path.join(libdir, "csrc"), "4096"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "sdeb200 ABI 4" in out.stdout and "compiled sde_solve_kernel for sm_100a" in out.stdout
    assert ("no CUDA device" in out.stdout) or ("E[x_T]" in out.stdout)
