"""The generated per-node programs must compile for sm_100a (nvcc cross-compiles
without a GPU) and the C-ABI library must export every declared symbol."""
import ctypes
import re
import subprocess
import tempfile
from pathlib import Path

import pytest

from helpers import build

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("case", ["general_lgr", "rocket_lgl_4x5", "quadrotor_lgl_14x6"])
def test_generated_cuda_compiles_for_sm100a(case):
    from pockit_b200 import plan as P

    S = build(case)
    dp = P.DevicePlan(S.lowering)
    for m in range(6):
        dp.mode(m)
    with tempfile.TemporaryDirectory() as tmp:
        for m in range(7):  # the five callbacks, the fused set pipeline, the error-estimate programs
            src = dp.finalize(m)["source"] if m < 6 else dp.error_estimate()["source"]
            cu = Path(tmp) / f"mode{m}.cu"
            cu.write_text(src)
            r = subprocess.run(
                ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "--fmad=false",
                 "-std=c++17", "-cubin", "-o", str(cu.with_suffix(".cubin")), str(cu)],
                capture_output=True, text=True,
            )
            assert r.returncode == 0, r.stderr[-2000:]


def test_library_exports_declared_symbols():
    import ctypes

    import __graft_entry__ as g

    g.build()
    header = (ROOT / "include" / "pockit_b200.h").read_text()
    names = set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 20
    lib = ctypes.CDLL(str(ROOT / "pockit_b200" / "libpockit_b200.so"))
    for n in sorted(names):
        assert hasattr(lib, n), n
    lib.pk_abi_version.restype = ctypes.c_int
    assert lib.pk_abi_version() == 2


def _nvrtc_compile(source: str, options):
    """Compile with the NVRTC library itself (what the engine does at run time; no GPU needed)."""
    lib = None
    for name in ("libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.12"):
        try:
            lib = ctypes.CDLL(name)
            break
        except OSError:
            continue
    if lib is None:
        pytest.skip("libnvrtc not found")
    prog = ctypes.c_void_p()
    assert lib.nvrtcCreateProgram(ctypes.byref(prog), source.encode(), b"pockit_b200_generated.cu", 0, None, None) == 0
    opts = (ctypes.c_char_p * len(options))(*[o.encode() for o in options])
    rc = lib.nvrtcCompileProgram(prog, len(options), opts)
    n = ctypes.c_size_t()
    lib.nvrtcGetProgramLogSize(prog, ctypes.byref(n))
    log = ctypes.create_string_buffer(n.value or 1)
    lib.nvrtcGetProgramLog(prog, log)
    size = ctypes.c_size_t()
    ok = rc == 0 and lib.nvrtcGetCUBINSize(prog, ctypes.byref(size)) == 0 and size.value > 0
    lib.nvrtcDestroyProgram(ctypes.byref(prog))
    return ok, log.value.decode(errors="replace")


@pytest.mark.parametrize("case,groups", [("general_lgr", 1), ("rocket_lgl_4x5", 1), ("robot_arm_lgr_6x20", 4)])
def test_generated_cuda_compiles_with_nvrtc(case, groups):
    """Same options as pk_engine.cu::compile_source; covers the per-callback programs, the fused set
    pipeline, the error-estimate programs and the expression-group variant."""
    from pockit_b200 import plan as P

    S = build(case)
    dp = P.DevicePlan(S.lowering, node_groups=groups)
    for m in range(6):
        dp.mode(m)
    sources = [dp.finalize(m)["source"] for m in range(6)] + [dp.error_estimate()["source"]]
    for k, src in enumerate(sources):
        ok, log = _nvrtc_compile(src, ["--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=false"])
        assert ok, f"program {k}: {log[-1500:]}"
