"""Host emulation of the device plan -- TEST INFRASTRUCTURE ONLY.

Lets the CPU-only test tier check the planner, the CUDA-C emitter and the job
records without a GPU: the generated per-node / system programs are compiled
*as C++* with a shim for the CUDA keywords and run thread by thread, and the job
records are interpreted with NumPy following the documented semantics of the
hand-written kernels (pockit_b200/csrc/pk_kernels.cuh).  The product never
imports this module; on a GPU the same plan runs through libpockit_b200.so.
"""
import ctypes
import hashlib
import subprocess
import tempfile
from pathlib import Path

import numpy as np

from pockit_b200 import plan as P

SHIM = r"""
#include <math.h>
#define __constant__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
struct pk_dim3 { long long x, y, z; };
static thread_local pk_dim3 blockIdx, blockDim, threadIdx;
"""

_CACHE = Path(tempfile.gettempdir()) / "pockit_b200_hostemu"


def _compile(source: str, kernels, sys_kernel):
    drivers = []
    for k in kernels:
        drivers.append(
            f'extern "C" void run_{k}(const double* X, const double* LAM, const double* FIX, const double* TM,'
            f" const double* WM, double* S, double* W, double* OUT, int B, const double* DP, const long long* IP,"
            f" long long threads) {{\n"
            f"  blockDim.x = 128;\n"
            f"  for (long long g = 0; g < threads; ++g) {{ blockIdx.x = g / 128; threadIdx.x = g % 128;\n"
            f"    {k}(X, LAM, FIX, TM, WM, S, W, OUT, B, DP, IP); }}\n}}\n"
        )
    if sys_kernel:
        drivers.append(
            f'extern "C" void run_{sys_kernel}(const double* X, double* S, double* OUT, int B) {{\n'
            f"  blockDim.x = 64;\n"
            f"  for (int g = 0; g < B; ++g) {{ blockIdx.x = g / 64; threadIdx.x = g % 64; {sys_kernel}(X, S, OUT, B); }}\n}}\n"
        )
    text = SHIM + source + "\n".join(drivers)
    _CACHE.mkdir(exist_ok=True)
    key = hashlib.sha256(text.encode()).hexdigest()[:24]
    so = _CACHE / f"{key}.so"
    if not so.exists():
        cpp = _CACHE / f"{key}.cpp"
        cpp.write_text(text)
        subprocess.check_call(
            ["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(cpp)]
        )
    return ctypes.CDLL(str(so))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class HostEmu:
    def __init__(self, system, batch=1, fixed=None, fused=False, shard=None, node_groups=1, set_subs=P.SET_ORDER):
        self.lo = system.lowering
        self.dp = P.DevicePlan(self.lo, batch, fused=fused, shard=shard, node_groups=node_groups, set_subs=set_subs)
        self.B = batch
        modes = range(6) if shard is None and not fused else range(5)  # + the fused set pipeline
        for m in modes:
            self.dp.mode(m)
            self.dp.source(m)
        self.fin = {m: self.dp.finalize(m) for m in modes}
        self.dpool, self.ipool = self.dp.pools.arrays()
        self.fixed = np.tile(self.dp.fixed_default, (batch, 1)) if fixed is None else np.asarray(fixed, float)
        self.libs = {}

    def run(self, mode, x, lam=None, sigma=None):
        lo, dp, B = self.lo, self.dp, self.B
        f = self.fin[mode]
        if mode not in self.libs:
            lib = _compile(f["source"], f["kernels"], f["sys_kernel"])
            tab = (ctypes.c_longlong * max(1, len(f["table"]))).in_dll(lib, f["table_symbol"])
            for k, v in enumerate(f["table"]):
                tab[k] = int(v)
            self.libs[mode] = lib
        lib = self.libs[mode]
        X = np.ascontiguousarray(np.asarray(x, float).reshape(B, lo.r_s))
        LAM = np.zeros((B, max(1, lo.m))) if lam is None else np.ascontiguousarray(np.asarray(lam, float).reshape(B, lo.m))
        SIG = np.zeros(B) if sigma is None else np.broadcast_to(np.asarray(sigma, float), (B,)).copy()
        S = np.zeros((B, max(1, f["n_scalar"])))
        W = np.full(max(1, f["n_table"]), np.nan)
        OUT = np.full((B, max(1, f["n_out"])), np.nan)
        FIX = np.ascontiguousarray(self.fixed.reshape(B, -1)) if self.fixed.size else np.zeros((B, 1))
        for pi, k in enumerate(f["kernels"]):
            Lm = lo.phases[pi].L_m
            tm = self.dpool[dp.tm_off[pi] : dp.tm_off[pi] + Lm].copy()
            wm = self.dpool[dp.wm_off[pi] : dp.wm_off[pi] + Lm].copy()
            getattr(lib, "run_" + k)(
                _ptr(X), _ptr(LAM), _ptr(FIX), _ptr(tm), _ptr(wm), _ptr(S), _ptr(W), _ptr(OUT),
                ctypes.c_int(B), _ptr(self.dpool), _ptr(self.ipool), ctypes.c_longlong(B * Lm * f["node_threads"][pi]),
            )
        jobs = f["jobs"]
        for jb in jobs[P.ST_REDUCE]:
            i = jb["i"]
            for b in range(B):
                row = W[i[0] + b * i[1] : i[0] + (b + 1) * i[1]]
                S[b, i[4]] = row[i[2] : i[3]].sum()
        if f["sys_kernel"]:
            getattr(lib, "run_" + f["sys_kernel"])(_ptr(X), _ptr(S), _ptr(OUT), ctypes.c_int(B))
        for jb in jobs[P.ST_DEFECT]:
            self._defect(jb, X, S, W, OUT)
        for jb in jobs[P.ST_GENERIC]:
            self._generic(jb, S, W, OUT, LAM, SIG)
        for jb in jobs[P.ST_EXPAND]:
            self._expand(jb, W, OUT, LAM)
        if f["grad_range"][1]:
            g0, gn = f["grad_range"]
            OUT[:, g0 : g0 + gn] = 0.0
            for jb in jobs[P.ST_GRAD_RANGE]:
                i = jb["i"]
                q = self.ipool[i[2] : i[2] + 4 * i[3]].reshape(-1, 4)
                for b in range(B):
                    acc = np.zeros(i[1])
                    for w0, lm, c_lo, sys in q:
                        acc = acc + W[w0 + b * lm + c_lo : w0 + b * lm + c_lo + i[1]] * S[b, sys]
                    OUT[b, i[0] : i[0] + i[1]] = acc
            for jb in jobs[P.ST_GRAD_SCALAR]:
                i = jb["i"]
                q = self.ipool[i[2] : i[2] + 2 * i[3]].reshape(-1, 2)
                for b in range(B):
                    acc = 0.0
                    for v, sys in q:
                        acc += S[b, v] * S[b, sys] if v >= 0 else S[b, sys]
                    OUT[b, i[0]] = acc
        out = OUT[:, : f["n_out"]]
        return out[0] if B == 1 else out

    # -- interpreters of the job records (semantics of pk_kernels.cuh) --------------------
    def _defect(self, jb, X, S, W, OUT):
        i = jb["i"]
        Lx, Lm, n_x, rows = i[1], i[2], i[3], i[4]
        ptr = self.ipool[i[5] : i[5] + rows + 1]
        nnz = ptr[-1]
        col = self.ipool[i[6] : i[6] + nnz]
        dat = self.dpool[i[7] : i[7] + nnz]
        cp = self.ipool[i[8] : i[8] + rows]
        cn = self.ipool[i[9] : i[9] + rows]
        for b in range(self.B):
            hdr = S[b, i[11] :]
            for s_ in range(n_x):
                f = W[i[10] + (s_ * self.B + b) * Lm : i[10] + (s_ * self.B + b + 1) * Lm]
                xv = X[b, i[0] + s_ * Lx : i[0] + (s_ + 1) * Lx].copy()
                xv[0] = hdr[1 + s_]
                xv[Lx - 1] = hdr[1 + n_x + s_]
                acc = np.zeros(rows)
                np.add.at(acc, np.repeat(np.arange(rows), np.diff(ptr)), dat * f[col])
                OUT[b, i[12] + s_ * rows : i[12] + (s_ + 1) * rows] = (xv[cp] - xv[cn]) - acc * hdr[0]

    def _src(self, b, S, W, scalar, unit, src, lm, c_lo, idx):
        if unit:
            return np.ones(len(idx))
        if scalar:
            return np.full(len(idx), S[b, src])
        return W[src + b * lm + c_lo + idx]

    def _generic(self, jb, S, W, OUT, LAM, SIG):
        i, fl, t = jb["i"], int(jb["flags"]), int(jb["type"])
        sign = jb["f"][0]
        cnt = i[1]
        e = np.arange(cnt)
        for b in range(self.B):
            sysv = S[b, i[3]] if i[3] >= 0 else 1.0
            post = {0: None, 1: SIG[b], 2: LAM[b, i[5]] if i[4] == 2 else None}[int(i[4])]
            if t == P.J_CONST:
                v = self.dpool[i[6] : i[6] + cnt]
            elif t == P.J_KRON:
                nb = i[7]
                a, q = e // nb, e % nb
                d = self.dpool[i[6] + a]
                if fl & P.F_LAM:
                    d = d * LAM[b, self.ipool[i[9] + a]]
                v = sign * (d * S[b, i[8] + q])
            elif t == P.J_EXPAND_TABLE:
                d = sign * self.dpool[i[8] : i[8] + cnt]
                if fl & P.F_LAM:
                    d = d * LAM[b, i[2] + self.ipool[i[6] : i[6] + cnt]]
                v = d * W[i[9] + b * i[10] + self.ipool[i[7] : i[7] + cnt]]
            elif t == P.J_SCALED:
                v = self._src(b, S, W, fl & P.F_A_SCALAR, 0, i[6], i[7], i[8], e) * sysv
                if post is not None:
                    v = v * post
            elif t == P.J_SYS:
                v = np.array([sysv]) if post is None else np.array([sysv * post])
            else:
                if t == P.J_OUTER:
                    ia, ib = e // i[12], e % i[12]
                else:
                    ia = ((np.sqrt(8.0 * e + 1) - 1) / 2).astype(np.int64)
                    ia -= ia * (ia + 1) // 2 > e
                    ia += (ia + 1) * (ia + 2) // 2 <= e
                    ib = e - ia * (ia + 1) // 2
                va = self._src(b, S, W, fl & P.F_A_SCALAR, fl & P.F_A_UNIT, i[6], i[7], i[8], ia)
                vb = self._src(b, S, W, fl & P.F_B_SCALAR, fl & P.F_B_UNIT, i[9], i[10], i[11], ib)
                v = (va * vb) * sysv
                if post is not None:
                    v = v * post
            OUT[b, i[0] : i[0] + cnt] = v

    def _expand(self, jb, W, OUT, LAM):
        i, fl = jb["i"], int(jb["flags"])
        sign = jb["f"][0]
        n, rows, step, c0 = i[3], i[4], i[5], i[6]
        bn = n * rows
        nK = i[11] // n
        e = np.arange(nK * bn)
        K, rem = e // bn, e % bn
        r, cc = rem // n, rem % n
        unit = self.dpool[i[7] : i[7] + bn]
        width = self.dpool[i[8] : i[8] + nK]
        lists = self.ipool[i[0] : i[0] + 2 * i[1]].reshape(-1, 2)
        for b in range(self.B):
            v = sign * ((unit[rem] * width[K]) / 2.0)
            if fl & P.F_LAM:
                v = v * LAM[b, i[2] + K * rows + r]
            for dst, wbase in lists:
                OUT[b, dst : dst + len(e)] = v * W[wbase + b * i[10] + c0 + K * step + cc]


def emulate_error_data(system, x):
    """Host emulation of the engine's continuous error-estimate path (``pk_eval_error_data``): the
    generated prep / node programs compiled as C++ and run thread by thread, the three CSR operators
    applied with SciPy (sequential row sums, like the device kernel).  B = 1."""
    lo = system.lowering
    dp = P.DevicePlan(lo)
    ee = dp.error_estimate()
    drivers = []
    for ph in ee["phases"]:
        drivers.append(
            f'extern "C" void run_{ph["prep"]}(const double* X, const double* FIX, double* XS, long long threads) {{\n'
            f"  blockDim.x = 128;\n  for (long long g = 0; g < threads; ++g) {{ blockIdx.x = g / 128; threadIdx.x = g % 128;\n"
            f'    {ph["prep"]}(X, FIX, XS, 1); }}\n}}\n'
            f'extern "C" void run_{ph["node"]}(const double* X, const double* XS, const double* XU, const double* TMA, double* WA,'
            f" long long threads) {{\n  blockDim.x = 128;\n"
            f"  for (long long g = 0; g < threads; ++g) {{ blockIdx.x = g / 128; threadIdx.x = g % 128;\n"
            f'    {ph["node"]}(X, XS, XU, TMA, WA, 1); }}\n}}\n'
        )
    text = SHIM + ee["source"] + "\n".join(drivers)
    _CACHE.mkdir(exist_ok=True)
    key = hashlib.sha256(text.encode()).hexdigest()[:24]
    so = _CACHE / f"{key}.so"
    if not so.exists():
        cpp = _CACHE / f"{key}.cpp"
        cpp.write_text(text)
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(cpp)])
    lib = ctypes.CDLL(str(so))
    X = np.ascontiguousarray(np.asarray(x, float))
    FIX = np.ascontiguousarray(dp.fixed_default) if dp.n_fixed else np.zeros(1)
    out = []
    for ph in ee["phases"]:
        XS = np.full(ph["L"], np.nan)
        getattr(lib, "run_" + ph["prep"])(_ptr(X), _ptr(FIX), _ptr(XS), ctypes.c_longlong(ph["L"]))
        XU = np.ascontiguousarray(ph["V"].dot(XS[: ph["L_xu"]]))
        WA = np.full(max(1, ph["n_x"] * ph["Lm_aug"]), np.nan)
        tm = np.ascontiguousarray(ph["tm_aug"])
        getattr(lib, "run_" + ph["node"])(_ptr(X), _ptr(XS), _ptr(XU), _ptr(tm), _ptr(WA), ctypes.c_longlong(ph["Lm_aug"]))
        TX = ph["T"].dot(XS[: ph["L_x_all"]]).reshape(ph["n_x"], -1)
        dt = XS[-1] - XS[-2]
        IF = np.array([ph["I"].dot(WA[i * ph["Lm_aug"] : (i + 1) * ph["Lm_aug"]]) * dt for i in range(ph["n_x"])]).reshape(ph["n_x"], -1)
        out.append((TX, IF))
    return out
