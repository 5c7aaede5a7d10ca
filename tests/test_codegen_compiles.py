"""The generated per-node programs must compile for sm_100a (nvcc cross-compiles
without a GPU) and the C-ABI library must export every declared symbol."""
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
