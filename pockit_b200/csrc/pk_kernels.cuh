// pk_kernels.cuh -- hand-written sm_100a kernels of the pockit-b200 engine.
//
// The per-node programs (function values / derivatives and their chain-rule
// products) are generated per model and JIT-compiled with NVRTC; everything
// that is *model independent* lives here and is driven by job records
// (include/pockit_b200.h, pk_job):
//
//   pk_reduce_rows    quadrature sums / np.add.at on broadcast columns   (phasebase.py:1004, systembase.py:654-656)
//   pk_defects        T.x - dt * (I.f), CSR row order                    (phasebase.py:1008-1012)
//   pk_expand_blocks  -I[r,c] * [lam_r] * list[c] for whole interval blocks   (phasebase.py:1120-1124, 1280-1285)
//   pk_generic_jobs   const / kron / table-expand / scaled / sys / outer / tril runs
//   pk_grad_range / pk_grad_scalar   gather-sum of the objective gradient (systembase.py:646-657)
//
// All of them are HBM-bound streaming kernels: every thread produces
// consecutive output slots (coalesced 8-byte stores, the runs are contiguous
// in the reference's slot order by construction), sources are the small,
// L2-resident node table W and scalar table S, and the constant pools.
// Arithmetic association follows the reference so that results agree to the
// last bit whenever the per-node leaves do (compiled with --fmad=false).
//
// Job field layout (i[] / f[] of pk_job):
//   REDUCE   i0 W base of the row, i1 L_m, i2 c_lo, i3 c_hi, i4 scalar slot
//   DEFECT   i0 x offset of the phase, i1 L_x, i2 L_m, i3 n_x, i4 rows per state,
//            i5 ipool row_ptr, i6 ipool cols, i7 dpool data (full operator, row-major),
//            i8 ipool +1 column per row, i9 ipool -1 column per row, i10 W base of f_0 (rows consecutive),
//            i11 scalar header (dt, front values, back values), i12 first output row,
//            i13 n | 0, flags rows per block, i14 dpool unit block (column-major), i15 dpool widths (same-order fast path)
//   GENERIC  i0 dst, i1 count, i2 lam row offset (-1: none), i3 system scalar slot (-1: none),
//            i4 post factor (0 none, 1 sigma, 2 lam[i5]), f0 sign, i15 division multiplier of count (set by the engine)
//     CONST         i6 dpool values
//     KRON          i6 dpool data[a], i7 nb, i8 first scalar slot (nb consecutive), i9 ipool lam rows[a] (F_LAM)
//     EXPAND_TABLE  i6 ipool rows, i7 ipool cols, i8 dpool data, i9 W base, i10 L_m
//     SCALED        i6 W base | scalar slot (F_A_SCALAR), i7 L_m, i8 c_lo
//     SYS           -
//     OUTER / TRIL  A: i6 i7 i8, B: i9 i10 i11 (as SCALED; F_*_SCALAR / F_*_UNIT), i12 length of B
//   EXPAND   i0 ipool list table (dst, W row base) x i1 lists, i2 lam row of the first block row (-1),
//            i3 n (block columns), i4 block rows,
//            i5 node step per interval, i6 first node, i7 dpool unit block, i8 dpool widths, i9 W base, i10 L_m,
//            i11 (interval, column) pairs = intervals * n, f0 sign, i12 / i13 division multipliers of pairs / n (set by the engine)
//   GRAD_RANGE   i0 dst, i1 count, i2 ipool contributions (W base, L_m, c_lo, system slot) x i3
//   GRAD_SCALAR  i0 dst, i2 ipool contributions (scalar slot | -1, system slot) x i3
#pragma once
#include <stdint.h>

#include "../../include/pockit_b200.h"

#define PK_F_A_SCALAR 1
#define PK_F_B_SCALAR 2
#define PK_F_A_UNIT 4
#define PK_F_B_UNIT 8
#define PK_F_LAM 16

#define PK_THREADS 256
#define PK_ITEMS 4  // consecutive output slots per thread
#define PK_CHUNK (PK_THREADS * PK_ITEMS)

// Result stores.  The values are never read again on the device (except by the opt-in compaction
// pass), so by default they are STREAMING stores (st.global.cs, evict-first): the write-once stream
// does not displace the node table / multipliers from L2 and drains to HBM sooner.  Measured on B200
// (tools/microbench_expand3.cu, round 2): column walk of 20x20 blocks, 102 MB: 20.96 -> 18.53 us
// (5.5 TB/s); LGL 9x10 blocks at unaligned slot offsets: 33.2 -> 26.9 us.  `stream` = 0 keeps plain
// stores (compaction re-reads the values from L2 right away).
__device__ __forceinline__ void pk_store(double* p, double v, int stream) {
  if (stream) __stcs(p, v);
  else *p = v;
}

// floor(x / d) for x, d < 2^32 with the pre-computed multiplier m = floor(2^64 / d) + 1 (0 encodes d == 1);
// the engine patches m into the job records at load time (pk_engine_load_mode) or passes it as an argument.
__device__ __forceinline__ unsigned pk_div(unsigned x, unsigned long long m) {
  return m ? (unsigned)__umul64hi((unsigned long long)x, m) : x;
}

struct PkCtx {
  const double* __restrict__ X;
  const double* __restrict__ LAM;
  const double* __restrict__ SIG;
  double* __restrict__ S;
  double* __restrict__ W;
  double* __restrict__ OUT;
  const double* __restrict__ dpool;
  const long long* __restrict__ ipool;
  long long L, m, n_scalar, n_out;
  int stream;  // 1: results are written with streaming stores (pk_store)
};

// ---------------------------------------------------------------------------------------------
// row sums, deterministic (fixed strides + fixed tree).  Short rows (batched small problems): one
// warp per (job, instance); long rows (fine meshes): one block per (job, instance).
__global__ void __launch_bounds__(PK_THREADS) pk_reduce_rows(PkCtx cx, const pk_job* __restrict__ jobs, int n_jobs, int B) {
  const int warp = (blockIdx.x * PK_THREADS + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_jobs * B) return;
  const int j = warp % n_jobs, b = warp / n_jobs;
  const pk_job& jb = jobs[j];
  const double* row = cx.W + jb.i[0] + (long long)b * jb.i[1];
  double acc = 0.0;
  for (long long c = jb.i[2] + lane; c < jb.i[3]; c += 32) acc += row[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) cx.S[(long long)b * cx.n_scalar + jb.i[4]] = acc;
}

// Long rows: PK_REDUCE_SPAN elements per block -> partial sums; the block that takes the last
// ticket of its (job, instance) adds the partials in index order, so the result does not depend
// on scheduling.  `partial` holds [job][instance][part], `ticket` one self-resetting counter each.
#define PK_REDUCE_SPAN 1024
__global__ void __launch_bounds__(PK_THREADS) pk_reduce_rows_block(PkCtx cx, const pk_job* __restrict__ jobs, int n_jobs, int B,
                                                                  int parts, double* __restrict__ partial,
                                                                  unsigned* __restrict__ ticket) {
  __shared__ double warp_sum[PK_THREADS / 32];
  __shared__ bool is_last;
  const int part = blockIdx.x % parts;
  const int jbi = blockIdx.x / parts;  // = b * n_jobs + j
  const int j = jbi % n_jobs, b = jbi / n_jobs;
  const pk_job& jb = jobs[j];
  const double* row = cx.W + jb.i[0] + (long long)b * jb.i[1];
  const long long lo = jb.i[2] + (long long)part * PK_REDUCE_SPAN;
  const long long hi = jb.i[3] < lo + PK_REDUCE_SPAN ? (long long)jb.i[3] : lo + PK_REDUCE_SPAN;
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < PK_REDUCE_SPAN / PK_THREADS; ++k) {
    const long long c = lo + threadIdx.x + k * PK_THREADS;
    if (c < hi) acc += row[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < PK_THREADS / 32; ++w) t += warp_sum[w];
    partial[(long long)jbi * parts + part] = t;
    __threadfence();
    is_last = atomicAdd(&ticket[jbi], 1u) == (unsigned)(parts - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    const int used = (int)((jb.i[3] - jb.i[2] + PK_REDUCE_SPAN - 1) / PK_REDUCE_SPAN);
    double t = 0.0;
    for (int q = threadIdx.x; q < used; q += 32) t += ((volatile double*)partial)[(long long)jbi * parts + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      cx.S[(long long)b * cx.n_scalar + jb.i[4]] = t;
      ticket[jbi] = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// one thread per defect row; the I.f sum runs over the row's entries in column order, dt is
// applied after the sum, exactly like `T_v.dot(x) - I_m.dot(f) * dt`.  Table-driven version for
// meshes with mixed orders (or blocks that lost exact zeros):
__global__ void __launch_bounds__(PK_THREADS) pk_defects(PkCtx cx, const pk_job* __restrict__ jobs, int n_jobs, int B) {
  const pk_job& jb = jobs[blockIdx.y];
  if (jb.i[13]) return;  // handled by pk_defects_blocks
  const long long rows = jb.i[4], n_x = jb.i[3];
  const long long gid = blockIdx.x * (long long)PK_THREADS + threadIdx.x;
  if (gid >= rows * n_x * B) return;
  const int b = (int)(gid / (rows * n_x));
  const long long rem = gid - (long long)b * rows * n_x;
  const int i = (int)(rem / rows);
  const long long r = rem - (long long)i * rows;
  const long long Lx = jb.i[1], Lm = jb.i[2];
  const double* Sb = cx.S + (long long)b * cx.n_scalar + jb.i[11];
  const double* xv = cx.X + (long long)b * cx.L + jb.i[0] + (long long)i * Lx;
  const double* f = cx.W + jb.i[10] + ((long long)i * B + b) * Lm;
  const long long* ptr = cx.ipool + jb.i[5];
  const long long* col = cx.ipool + jb.i[6];
  const double* dat = cx.dpool + jb.i[7];
  double acc = 0.0;
  for (long long k = ptr[r]; k < ptr[r + 1]; ++k) acc += dat[k] * f[col[k]];
  const long long cp = cx.ipool[jb.i[8] + r], cn = cx.ipool[jb.i[9] + r];
  // boundary nodes read the substituted values the node program left in the scalar header
  const double xp = cp == 0 ? Sb[1 + i] : (cp == Lx - 1 ? Sb[1 + n_x + i] : xv[cp]);
  const double xn = cn == 0 ? Sb[1 + i] : (cn == Lx - 1 ? Sb[1 + n_x + i] : xv[cn]);
  double tx = 0.0;
  tx += 1.0 * xp;
  tx += -1.0 * xn;
  cx.OUT[(long long)b * cx.n_out + jb.i[12] + (long long)i * rows + r] = tx - acc * Sb[0];
}

// Same-order meshes: the operator is never read from memory.  i13 = n (points per interval),
// flags = rows per block rb (= node step = offset of the -1 column: n-1 for LGL, n for LGR),
// i14 dpool unit block stored COLUMN-major (n x rows: entry [c * rb + r]), i15 dpool interval widths.
// Entries are (unit * width) / 2.  One thread per defect row over the flattened (instance, state, row)
// space (batches of small problems keep whole warps busy); the three index splits are multiplies by
// pre-computed reciprocals passed as kernel arguments (FAST: everything fits 32 bits), offsets inside an
// instance are 32-bit.  (ncu, round 2: with runtime divisions and 64-bit offsets the kernel executed 318
// instructions per row for a 6-term row sum and was issue-bound -- quadrotor B = 8192: 41 us; a 3-D grid
// without divisions was worse, 45 us: 49 k blocks of 70 live threads.)
// The unit block is read through L1 (__ldg): consecutive lanes own consecutive rows, so for a given
// column they read consecutive doubles; no shared-memory staging, no barrier.
struct PkDefectDiv {
  unsigned long long rows, n_x, rb;  // division multipliers (pk_div)
};
#define PK_DEF_JOBS 8
struct PkDefectDivs {
  PkDefectDiv d[PK_DEF_JOBS];  // one per defect job (= phase) of the launch: blockIdx.y
};

template <bool FAST>
__global__ void __launch_bounds__(PK_THREADS) pk_defects_blocks(PkCtx cx, const pk_job* __restrict__ jobs, int job0, int B,
                                                               const __grid_constant__ PkDefectDivs dvs) {
  const pk_job& jb = jobs[job0 + blockIdx.y];
  if (!jb.i[13]) return;  // a table-driven job (mixed orders): pk_defects
  const PkDefectDiv& dv = dvs.d[blockIdx.y];
  const int n = (int)jb.i[13];
  const unsigned rb = (unsigned)jb.flags;
  const unsigned rows = (unsigned)jb.i[4], n_x = (unsigned)jb.i[3];
  unsigned b, i, r, K;
  if (FAST) {
    const unsigned gid = blockIdx.x * PK_THREADS + threadIdx.x;
    if (gid >= rows * n_x * (unsigned)B) return;
    const unsigned pair = pk_div(gid, dv.rows);  // (instance, state)
    r = gid - pair * rows;
    b = pk_div(pair, dv.n_x);
    i = pair - b * n_x;
    K = pk_div(r, dv.rb);
  } else {
    const unsigned long long gid = blockIdx.x * (unsigned long long)PK_THREADS + threadIdx.x;
    if (gid >= (unsigned long long)rows * n_x * (unsigned long long)B) return;
    const unsigned long long pair = gid / rows;
    r = (unsigned)(gid - pair * rows);
    b = (unsigned)(pair / n_x);
    i = (unsigned)(pair - (unsigned long long)b * n_x);
    K = r / rb;
  }
  const unsigned rr = r - K * rb;
  const unsigned Lx = (unsigned)jb.i[1], Lm = (unsigned)jb.i[2];
  const double* __restrict__ Sb = cx.S + (long long)b * cx.n_scalar + jb.i[11];
  const double* __restrict__ xv = cx.X + (long long)b * cx.L + jb.i[0] + (long long)i * Lx;
  const double* __restrict__ f = cx.W + jb.i[10] + ((long long)i * B + b) * Lm + K * rb;
  const double w = cx.dpool[jb.i[15] + K];
  const double* __restrict__ u = cx.dpool + jb.i[14] + rr;
  double acc = 0.0;
#pragma unroll 4
  for (int c = 0; c < n; ++c) acc += ((__ldg(u + c * rb) * w) / 2.0) * f[c];
  const unsigned cp = K * rb + rr, cn = K * rb + rb;
  const double xp = cp == 0 ? Sb[1 + i] : (cp == Lx - 1 ? Sb[1 + n_x + i] : xv[cp]);
  const double xn = cn == 0 ? Sb[1 + i] : (cn == Lx - 1 ? Sb[1 + n_x + i] : xv[cn]);
  double tx = 0.0;
  tx += 1.0 * xp;
  tx += -1.0 * xn;
  pk_store(cx.OUT + (long long)b * cx.n_out + jb.i[12] + (long long)i * rows + r, tx - acc * Sb[0], cx.stream);
}

// ---------------------------------------------------------------------------------------------
// block-structured expansion: every interior interval contributes a dense (rows x n) block whose
// entries are (unit[r][c] * width_K) / 2 -- the reference's `I_lgl(n) * d / 2` -- so the operator
// never has to be read from memory: the (sign-folded) unit block sits in shared memory.
// A job covers ALL lists of one state (they share the block geometry and the multiplier rows).
// Work unit = (list, interval, column): one node of one list.  The kernel is PERSISTENT: the grid
// is exactly one resident wave (SMs x blocks/SM) and walks the flattened unit space of all jobs
// with a grid stride, so every SM streams stores until the very end (no partial last wave).
// A unit loads its node's list value once and walks down the block column: consecutive lanes
// write runs of n consecutive slots, ~6 instructions per 8-byte store.
// prefix[j] = first unit of job j; i11 = (interval, column) pairs of the job.
__global__ void __launch_bounds__(PK_THREADS) pk_expand_blocks(PkCtx cx, const pk_job* __restrict__ jobs, int n_jobs,
                                                              const long long* __restrict__ prefix, int uniform) {
  extern __shared__ double unit_s[];
  const int b = blockIdx.y;
  if (uniform) {  // all jobs share one unit block and sign: keep it (sign folded, exact) in shared memory
    const pk_job& j0 = jobs[0];
    const int bn0 = (int)(j0.i[3] * j0.i[4]);
    const double* unit = cx.dpool + j0.i[7];
    for (int t = threadIdx.x; t < bn0; t += PK_THREADS) unit_s[t] = j0.f[0] * unit[t];
    __syncthreads();
  }
  const long long total = prefix[n_jobs];
  int j = 0;
  for (long long uidx = blockIdx.x * (long long)PK_THREADS + threadIdx.x; uidx < total;
       uidx += (long long)gridDim.x * PK_THREADS) {
    if (n_jobs > 8 && uidx >= prefix[j + 1]) {  // many runs (hp-refined mesh): binary search
      int hi = n_jobs;
      while (hi - j > 1) {
        const int mid = (j + hi) >> 1;
        if (uidx >= prefix[mid]) j = mid; else hi = mid;
      }
    }
    while (uidx >= prefix[j + 1]) ++j;
    const pk_job& jb = jobs[j];
    const int n = (int)jb.i[3], rows = (int)jb.i[4];
    const unsigned pairs = (unsigned)jb.i[11];
    const unsigned v = (unsigned)(uidx - prefix[j]);
    const unsigned l = pk_div(v, (unsigned long long)jb.i[12]);  // v / pairs (multipliers patched in at load time)
    const unsigned t = v - l * pairs;
    const unsigned K = pk_div(t, (unsigned long long)jb.i[13]);  // t / n
    const unsigned cc = t - K * (unsigned)n;
    const long long* __restrict__ lists = cx.ipool + jb.i[0] + 2 * (long long)l;  // (dst, W row base)
    const double sv = cx.W[lists[1] + (long long)b * jb.i[10] + jb.i[6] + (long long)K * jb.i[5] + cc];
    const double w = cx.dpool[jb.i[8] + K];
    double* __restrict__ out = cx.OUT + (long long)b * cx.n_out + lists[0] + (long long)K * (n * rows) + cc;
    const double* u = (uniform ? unit_s : cx.dpool + jb.i[7]) + cc;
    const double sgn = uniform ? 1.0 : jb.f[0];
    if (jb.flags & PK_F_LAM) {
      const double* __restrict__ lam = cx.LAM + (long long)b * cx.m + jb.i[2] + (long long)K * rows;
#pragma unroll 4
      for (int r = 0; r < rows; ++r) pk_store(out + r * n, (((sgn * u[r * n]) * w) / 2.0 * lam[r]) * sv, cx.stream);
    } else {
#pragma unroll 4
      for (int r = 0; r < rows; ++r) pk_store(out + r * n, (((sgn * u[r * n]) * w) / 2.0) * sv, cx.stream);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Column walk with everything a thread needs before its first store in the kernel PARAMETERS
// (constant bank): job geometry and the (dst, W row) of every list.  The persistent kernel above
// spends its first microseconds on a chain of dependent L2 round trips (job record -> list table
// -> list value); here the only global loads ahead of the stores are the list value and the
// interval width, issued before the unit block is staged.  Non-persistent: grid = (chunks of
// PK_XC_THREADS (interval, column) pairs, lists, instances) -- measured on B200, a write stream
// of this shape (tools/microbench_write2.cu, "cols 1 unit/thread") is what gets closest to the
// plain-fill ceiling at ~100 MB.  Multipliers of the block's intervals are staged in shared memory.
// Same-order meshes only (all jobs share unit block, shape and sign), <= PK_XC_JOBS jobs,
// <= PK_XC_LISTS lists; everything else takes pk_expand_blocks.
#define PK_XC_THREADS 128
#define PK_XC_JOBS 8
#define PK_XC_LISTS 96
struct PkXcJob {
  long long lam0;    // multiplier row of the first block row (LAM only)
  long long node0;   // first node of the first interval
  long long width;   // dpool offset of the interval widths
  long long Lm;      // nodes per W row
  int step;          // nodes per interval step
  unsigned pairs;    // intervals * n
  unsigned long long m_pairs;  // division multiplier of pairs (pk_div)
  unsigned long long m_run;    // division multiplier of pairs * rows (slots of one list of one instance)
  int list0, n_lists;          // the job's lists are prm.list[list0 .. list0 + n_lists)
};
struct PkXcList {
  long long dst, wbase;
  int job, pad;
};
struct PkXcParams {
  int n, rows, n_lists, n_jobs;
  long long unit;  // dpool offset of the unit block
  double sign;
  unsigned long long m_n;  // division multiplier of n
  unsigned long long m_bn; // division multiplier of n * rows
  // batch kernels: blockIdx.y -> (job, first list, number of lists) of the group of lists a thread writes,
  // packed as job | first << 8 | count << 20 (filled at set-up: no division or search in the kernel)
  unsigned grp[PK_XC_LISTS];
  PkXcJob job[PK_XC_JOBS];
  PkXcList list[PK_XC_LISTS];
};

template <bool LAM>
__global__ void __launch_bounds__(PK_XC_THREADS) pk_expand_cols(PkCtx cx, const __grid_constant__ PkXcParams prm) {
  extern __shared__ double smem_d[];
  const int n = prm.n, rows = prm.rows;
  const int bn = n * rows;
  double* u_s = smem_d;        // [bn] sign-folded unit block
  double* lam_s = smem_d + bn; // multipliers of the intervals this block touches
  const PkXcList& L = prm.list[blockIdx.y];
  const PkXcJob& J = prm.job[L.job];
  const int b = blockIdx.z;
  const unsigned t0 = blockIdx.x * PK_XC_THREADS;
  if (t0 >= J.pairs) return;  // lists of a shorter job
  const unsigned t = t0 + threadIdx.x;
  const bool live = t < J.pairs;
  const unsigned K = pk_div(live ? t : J.pairs - 1, prm.m_n);  // / n by the pre-computed reciprocal
  const unsigned cc = (live ? t : J.pairs - 1) - K * (unsigned)n;
  const unsigned K0 = pk_div(t0, prm.m_n);
  // the two loads the stores depend on go out first
  const double sv = cx.W[L.wbase + (long long)b * J.Lm + J.node0 + (long long)K * J.step + cc];
  const double w = cx.dpool[J.width + K];
  {
    // short trip counts (1-4): keep the staging loops rolled -- unrolled, their set-up code alone was a third of
    // the instructions a thread executes (ncu source page of the bulk variant, round 2)
    const double* unit = cx.dpool + prm.unit;
#pragma unroll 1
    for (int q = threadIdx.x; q < bn; q += PK_XC_THREADS) u_s[q] = prm.sign * unit[q];
    if (LAM) {
      unsigned tl = t0 + PK_XC_THREADS - 1;
      if (tl >= J.pairs) tl = J.pairs - 1;
      const int n_lam = (int)(pk_div(tl, prm.m_n) - K0 + 1) * rows;
      const double* __restrict__ lam = cx.LAM + (long long)b * cx.m + J.lam0 + (long long)K0 * rows;
#pragma unroll 1
      for (int q = threadIdx.x; q < n_lam; q += PK_XC_THREADS) lam_s[q] = lam[q];
    }
  }
  __syncthreads();
  if (!live) return;
  double* __restrict__ out = cx.OUT + (long long)b * cx.n_out + L.dst + (long long)K * bn + cc;
  const double* u = u_s + cc;
  const double* lm = lam_s + (K - K0) * rows;
#pragma unroll 4
  for (int r = 0; r < rows; ++r) {
    double v = (u[r * n] * w) / 2.0;
    if (LAM) v = v * lm[r];
    pk_store(out + r * n, v * sv, cx.stream);
  }
}

// ---------------------------------------------------------------------------------------------
// Block expansion for BATCHES of small problems (BASELINE configs[4]: 8192 instances of a 14-interval
// mesh, 72 (interval, column) pairs and 5 x 6 blocks per job).  With blocks this small the cost is index
// arithmetic, not bytes: pk_expand_blocks executes ~260 instructions per (list, interval, column) unit
// of 5 stores (ncu, round 2: issue-bound, 2.9 TB/s).  All lists of a job share the block geometry, the
// unit block, the interval width and the multipliers, so a thread here owns one (instance, interval,
// column) of a JOB, forms P[r] = ((sign * unit[r][c]) * width) / 2 [* lambda_r] ONCE in registers and then
// writes the column of EVERY list of the job as P[r] * list value: one FP64 multiply and one store per
// slot, ~6 instructions per slot instead of ~50.  Geometry and list table come from the kernel
// parameters (constant bank), index splits are multiplies by pre-computed reciprocals, no shared memory,
// no barrier.  Same association as the other expansion kernels: bit-identical results.
// (Round 1 measured this "P in registers" shape on the one-big-mesh case -- 20 x 20 blocks, 16 lists --
// and found it slower there: that case is bound by the store streams, this one by instruction issue.)
#define PK_XM_ROWS 16
#define PK_XM_LISTS 2  // lists per thread
#define PK_XM_AHEAD 4  // list values loaded ahead of the stores
#define PK_XS_LISTS 8  // lists per thread of pk_expand_slots
// ROWS >= rows is the unrolled trip count of the row loops.  The default launch uses ROWS = 16 for every
// block shape: the exact instantiation (POCKIT_B200_BATCH_ROWS=exact) executes a third of the instructions
// for 5-row blocks and needs 32 registers instead of 48, and is SLOWER on configs[4] (Jacobian expansion
// 68 -> 83 us, set 262 -> 273 us): the kernel is bound by how the memory system takes ~3 KB output runs
// 34 KB apart, 16 resident blocks per SM push more of them at once than 10 do.  Capping the resident blocks
// further (POCKIT_B200_BATCH_SMEM) is slower again (8 per SM: 276 us, 4: 326 us) -- profiles/
// r02_call26_batch_rows_exact.log, r02_call27_batch_occupancy_cap.log.
// STREAM: results written with streaming stores (cx.stream, decided at compile time here: a run-time test
// costs two branches per store in these store-dominated loops)
template <bool LAM, int ROWS, bool STREAM>
__global__ void __launch_bounds__(PK_XC_THREADS) pk_expand_batch(PkCtx cx, const __grid_constant__ PkXcParams prm, unsigned B) {
  const int n = prm.n, rows = prm.rows;
  const int bn = n * rows;
  // blockIdx.y enumerates (job, group of lists; PK_XM_LISTS per group unless overridden): a thread writes at
  // most that many columns, so that jobs with many lists still spread over enough threads (8 lists per
  // thread: 590 k threads for the quadrotor Jacobian, 80 us; 2 per thread as in its Hessian: 32 us for half
  // the bytes)
  const unsigned gw = prm.grp[blockIdx.y];
  const PkXcJob& J = prm.job[gw & 255u];
  const int l_lo = (int)((gw >> 8) & 4095u);
  const int l_hi = l_lo + (int)(gw >> 20);
  const unsigned t = blockIdx.x * PK_XC_THREADS + threadIdx.x;  // B * pairs < 2^32 (checked at set-up)
  if (t >= J.pairs * B) return;
  const unsigned b = pk_div(t, J.m_pairs);
  const unsigned tp = t - b * J.pairs;
  const unsigned K = pk_div(tp, prm.m_n);
  const unsigned cc = tp - K * (unsigned)n;
  const double w = cx.dpool[J.width + K];
  const double* __restrict__ u = cx.dpool + prm.unit + cc;
  const double* __restrict__ lam = cx.LAM + (long long)b * cx.m + J.lam0 + (long long)K * rows;
  double P[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (r < rows) {
      double v = ((prm.sign * __ldg(u + r * n)) * w) / 2.0;
      if (LAM) v = v * __ldg(lam + r);
      P[r] = v;
    }
  }
  const long long wofs = (long long)b * J.Lm + J.node0 + (long long)K * J.step + cc;
  double* __restrict__ out_b = cx.OUT + (long long)b * cx.n_out + (long long)K * bn + cc;
  // the list values of up to PK_XM_AHEAD lists are loaded before the first store: the compiler keeps a load
  // behind the stores that precede it in program order (possible aliasing), which serialised one memory
  // latency per list (ncu round 2: 21-32 long-scoreboard stall cycles per issue)
  for (int l0 = l_lo; l0 < l_hi; l0 += PK_XM_AHEAD) {
    double sv[PK_XM_AHEAD];
#pragma unroll
    for (int k = 0; k < PK_XM_AHEAD; ++k)
      if (l0 + k < l_hi) sv[k] = cx.W[prm.list[l0 + k].wbase + wofs];
#pragma unroll
    for (int k = 0; k < PK_XM_AHEAD; ++k)
      if (l0 + k < l_hi) {
        double* __restrict__ out = out_b + prm.list[l0 + k].dst;
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
          if (r < rows) pk_store(out + r * n, P[r] * sv[k], STREAM);
      }
  }
}

// ---------------------------------------------------------------------------------------------
// Slot-order batch expansion: the default for batch groups whose jobs have three or more lists (configs[4] Jacobian).  pk_expand_batch gives a thread a
// block COLUMN: with 5 x 6 blocks a warp's store instruction then writes 48-byte pieces 240 bytes apart,
// every one of them a partial 32-byte sector (ncu, round 2: 2.15 sector writes per sector of output).  Here
// a thread owns one SLOT (instance, interval, row, column) of a job -- the slots of one list of one
// instance are contiguous in the output, so a warp writes 32 consecutive doubles per list: whole sectors
// but for the two ends.  The factor ((sign * unit[r][c]) * width) / 2 [* lambda_r] is formed once per thread
// and multiplies the list value of each of the job's lists (PK_XS_LISTS per thread): same association as
// every other expansion kernel, bit-identical (tests/test_gpu_baseline_sizes.py).
// Measured on B200, configs[4] set: first version (one load, one store, next load ...) 280-295 us against
// 263 us of the column mapping -- each list cost a full memory latency (ncu: 21 long-scoreboard stall
// cycles per issue, 207 instructions per thread, a third of them the search for the list group).  With the
// list values loaded ahead of the stores, the list groups from a parameter table and the streaming-store
// switch a template argument: 250 us (column mapping with the same changes: 264 us) --
// profiles/r02_call25_slot_order_batch.log, r02_ncu_batch_expansion_shapes.csv, r02_call29_batch_loads_ahead.log.
template <bool LAM, bool STREAM>
__global__ void __launch_bounds__(PK_XC_THREADS) pk_expand_slots(PkCtx cx, const __grid_constant__ PkXcParams prm, unsigned B) {
  const unsigned n = (unsigned)prm.n, rows = (unsigned)prm.rows;
  const unsigned bn = n * rows;
  const unsigned gw = prm.grp[blockIdx.y];  // (job, group of lists)
  const PkXcJob& J = prm.job[gw & 255u];
  const int l_lo = (int)((gw >> 8) & 4095u);
  const int l_hi = l_lo + (int)(gw >> 20);
  const unsigned run = J.pairs * rows;
  const unsigned t = blockIdx.x * PK_XC_THREADS + threadIdx.x;  // B * run < 2^32 (checked at set-up)
  if (t >= run * B) return;
  const unsigned b = pk_div(t, J.m_run);
  const unsigned s = t - b * run;
  const unsigned K = pk_div(s, prm.m_bn);
  const unsigned q = s - K * bn;  // r * n + c
  const unsigned r = pk_div(q, prm.m_n);
  const unsigned cc = q - r * n;
  const double* __restrict__ wp = cx.W + ((long long)b * J.Lm + J.node0 + (long long)K * J.step + cc);
  double v = ((prm.sign * __ldg(cx.dpool + prm.unit + q)) * cx.dpool[J.width + K]) / 2.0;
  if (LAM) v = v * __ldg(cx.LAM + (long long)b * cx.m + J.lam0 + (long long)K * rows + r);
  double* __restrict__ out_b = cx.OUT + (long long)b * cx.n_out + s;
  for (int l0 = l_lo; l0 < l_hi; l0 += PK_XM_AHEAD) {  // loads ahead of the stores, as in pk_expand_batch
    double sv[PK_XM_AHEAD];
#pragma unroll
    for (int k = 0; k < PK_XM_AHEAD; ++k)
      if (l0 + k < l_hi) sv[k] = wp[prm.list[l0 + k].wbase];
#pragma unroll
    for (int k = 0; k < PK_XM_AHEAD; ++k)
      if (l0 + k < l_hi) pk_store(out_b + prm.list[l0 + k].dst, v * sv[k], STREAM);
  }
}

// ---------------------------------------------------------------------------------------------
// Bulk-store variant of the parameter-driven column walk: the default for interval blocks that are not a
// whole number of 32-byte sectors (LGL n = 10: 9 x 10 = 90 slots), where it is 10-15 % faster than the
// column walk's 8-byte stores (humanoid set 128.7 -> 121.4 us); slower on sector-aligned 20 x 20 blocks,
// where the column walk stays (POCKIT_B200_EXPAND=bulk | params overrides).  Bit-identical to the other
// kernels (tests/test_gpu_baseline_sizes.py), memcheck / racecheck clean.
// A block owns PK_XB_PAIRS / n WHOLE intervals of one list, i.e. one contiguous run of output slots.
// Its threads compute the same values with the same association as pk_expand_cols, but store them
// into a shared-memory image of that run; one thread then hands the image to the TMA engine as a
// single 1-D bulk copy (cp.async.bulk.global.shared::cta): the LSU issues no global stores and the
// run reaches L2 as whole sectors whatever the 8-byte alignment of the reference's slot offsets
// (the image is placed in shared memory with the same 16-byte phase as its destination; an odd
// first / last element is stored normally).
#define PK_XB_THREADS 128
template <bool LAM>
__global__ void __launch_bounds__(PK_XB_THREADS) pk_expand_bulk(PkCtx cx, const __grid_constant__ PkXcParams prm, int per_block) {
  extern __shared__ __align__(16) double smem_d[];
  const int n = prm.n, rows = prm.rows;
  const int bn = n * rows;
  double* u_s = smem_d;                               // [bn] sign-folded unit block
  double* lam_s = u_s + ((bn + 1) & ~1);              // [per_block * rows]
  double* img = lam_s + ((per_block * rows + 1) & ~1);  // [per_block * bn + 2], 16-byte aligned
  const PkXcList& L = prm.list[blockIdx.y];
  const PkXcJob& J = prm.job[L.job];
  const int b = blockIdx.z;
  const unsigned nK = pk_div(J.pairs, prm.m_n);
  const unsigned K0 = blockIdx.x * (unsigned)per_block;
  if (K0 >= nK) return;  // lists of a shorter job
  const unsigned Kn = nK - K0 < (unsigned)per_block ? nK - K0 : (unsigned)per_block;  // intervals of this block
  const unsigned t = threadIdx.x;
  const bool live = t < Kn * (unsigned)n;
  const unsigned kk = live ? pk_div(t, prm.m_n) : 0;
  const unsigned cc = live ? t - kk * (unsigned)n : 0;
  const unsigned K = K0 + kk;
  double sv = 0.0, w = 0.0;
  if (live) {
    sv = cx.W[L.wbase + (long long)b * J.Lm + J.node0 + (long long)K * J.step + cc];
    w = cx.dpool[J.width + K];
  }
  {
    const double* unit = cx.dpool + prm.unit;
#pragma unroll 1
    for (int q = t; q < bn; q += PK_XB_THREADS) u_s[q] = prm.sign * unit[q];
    if (LAM) {
      const double* __restrict__ lam = cx.LAM + (long long)b * cx.m + J.lam0 + (long long)K0 * rows;
#pragma unroll 1
      for (int q = t; q < (int)Kn * rows; q += PK_XB_THREADS) lam_s[q] = lam[q];
    }
  }
  double* __restrict__ out = cx.OUT + (long long)b * cx.n_out + L.dst + (long long)K0 * bn;  // first slot of the run
  const unsigned phase = (unsigned)(((unsigned long long)(size_t)out >> 3) & 1ull);  // 1: the run starts on an odd double
  double* run = img + phase;  // image element j lives at run[j]: same 16-byte phase as out[j]
  __syncthreads();
  if (live) {
    double* o = run + kk * (unsigned)bn + cc;
    const double* u = u_s + cc;
    const double* lm = lam_s + kk * (unsigned)rows;
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
      double v = (u[r * n] * w) / 2.0;
      if (LAM) v = v * lm[r];
      o[r * n] = v * sv;
    }
  }
  // make the generic-proxy writes to shared memory visible to the async proxy, then let one thread copy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (t == 0) {
    const unsigned total = Kn * (unsigned)bn;
    unsigned j0 = phase;                    // first element on a 16-byte boundary
    if (j0) out[0] = run[0];
    unsigned cnt = (total - j0) & ~1u;      // whole 16-byte pairs
    if (j0 + cnt < total) out[total - 1] = run[total - 1];
    if (cnt) {
      const unsigned src = (unsigned)__cvta_generic_to_shared(run + j0);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + j0), "r"(src), "r"(cnt * 8u) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the image must outlive the copy's reads
    }
  }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pk_list_at(const PkCtx& cx, const double* Sb, int b, int scalar, int unit,
                                             long long src, long long lm, long long c_lo, long long k) {
  if (unit) return 1.0;
  if (scalar) return Sb[src];
  return cx.W[src + (long long)b * lm + c_lo + k];
}

// The (instance, slot) space of a job is flattened: idx = b * count + e, so that batches of small
// problems (runs of a few dozen slots) still fill whole warps; consecutive lanes write consecutive
// slots of a run and continue in the next instance's run.  idx -> (b, e) is a multiply by the
// pre-computed reciprocal of count (i15, patched in at load time) on the 32-bit path; the two job types
// that carry most slots -- constants of T, table-driven expansion of the irregular first / last
// interval -- have their own loops without the per-slot scalar-table / post-factor set-up.
// (ncu, round 2: the one-loop version executed 112 instructions per slot and was issue-bound.)
template <typename IT>
__device__ __forceinline__ void pk_split(IT idx, IT count, unsigned long long magic, unsigned& b, IT& e) {
  if (sizeof(IT) == 4) b = pk_div((unsigned)idx, magic);
  else b = (unsigned)(idx / count);
  e = idx - (IT)b * count;
}

template <typename IT>  // unsigned whenever every job has count * B < 2^32
__global__ void __launch_bounds__(PK_THREADS) pk_generic_jobs(PkCtx cx, const pk_job* __restrict__ jobs,
                                                             const int* __restrict__ blk_job,
                                                             const int* __restrict__ blk_chunk, int B) {
  const pk_job& jb = jobs[blk_job[blockIdx.x]];
  const IT count = (IT)jb.i[1];
  const IT total = count * (IT)B;
  const unsigned long long magic = (unsigned long long)jb.i[15];
  const double sign = jb.f[0];
  const bool use_lam = jb.flags & PK_F_LAM;
  const IT i0 = (IT)blk_chunk[blockIdx.x] * (IT)PK_CHUNK + threadIdx.x;
  const int type = jb.type;
  double* __restrict__ out0 = cx.OUT + jb.i[0];
  if (type == PK_JOB_CONST) {
    const double* __restrict__ src = cx.dpool + jb.i[6];
#pragma unroll
    for (int it = 0; it < PK_ITEMS; ++it) {
      const IT idx = i0 + (IT)it * PK_THREADS;
      if (idx >= total) break;
      unsigned b; IT e;
      pk_split<IT>(idx, count, magic, b, e);
      pk_store(out0 + (long long)b * cx.n_out + (long long)e, __ldg(src + e), cx.stream);
    }
    return;
  }
  if (type == PK_JOB_EXPAND_TABLE) {
    const long long* __restrict__ trow = cx.ipool + jb.i[6];
    const long long* __restrict__ tcol = cx.ipool + jb.i[7];
    const double* __restrict__ tdat = cx.dpool + jb.i[8];
    const double* __restrict__ w0 = cx.W + jb.i[9];
    const long long lm = jb.i[10];
#pragma unroll
    for (int it = 0; it < PK_ITEMS; ++it) {
      const IT idx = i0 + (IT)it * PK_THREADS;
      if (idx >= total) break;
      unsigned b; IT e;
      pk_split<IT>(idx, count, magic, b, e);
      double d = sign * __ldg(tdat + e);
      if (use_lam) d = d * cx.LAM[(long long)b * cx.m + jb.i[2] + __ldg(trow + e)];
      pk_store(out0 + (long long)b * cx.n_out + (long long)e, d * w0[(long long)b * lm + __ldg(tcol + e)], cx.stream);
    }
    return;
  }
  if (type == PK_JOB_SCALED && !(jb.flags & PK_F_A_SCALAR)) {  // integral lists x dF/dI [x sigma | lambda_i]: a node-table row scaled
    const double* __restrict__ w0 = cx.W + jb.i[6] + jb.i[8];
    const long long lm = jb.i[7];
    const long long sys_slot = jb.i[3], post_kind = jb.i[4], post_row = jb.i[5];
#pragma unroll
    for (int it = 0; it < PK_ITEMS; ++it) {
      const IT idx = i0 + (IT)it * PK_THREADS;
      if (idx >= total) break;
      unsigned b; IT e;
      pk_split<IT>(idx, count, magic, b, e);
      const double sysv = sys_slot >= 0 ? cx.S[(long long)b * cx.n_scalar + sys_slot] : 1.0;
      double v = w0[(long long)b * lm + (long long)e] * sysv;
      if (post_kind == 1) v = v * cx.SIG[b];
      if (post_kind == 2) v = v * cx.LAM[(long long)b * cx.m + post_row];
      pk_store(out0 + (long long)b * cx.n_out + (long long)e, v, cx.stream);
    }
    return;
  }
  for (int it = 0; it < PK_ITEMS; ++it) {
    const IT idx = i0 + (IT)it * PK_THREADS;
    if (idx >= total) break;
    unsigned bu; IT e;
    pk_split<IT>(idx, count, magic, bu, e);
    const int b = (int)bu;
    const double* Sb = cx.S + (long long)b * cx.n_scalar;
    const double* lam = cx.LAM + (long long)b * cx.m;
    const double sysv = jb.i[3] >= 0 ? Sb[jb.i[3]] : 1.0;
    double post = 1.0;
    if (jb.i[4] == 1) post = cx.SIG[b];
    if (jb.i[4] == 2) post = lam[jb.i[5]];
    double v;
    switch (type) {
      case PK_JOB_KRON: {
        const IT nb = (IT)jb.i[7];
        const IT a = e / nb, q = e - a * nb;
        double d = cx.dpool[jb.i[6] + a];
        if (use_lam) d = d * lam[cx.ipool[jb.i[9] + a]];
        v = sign * (d * Sb[jb.i[8] + q]);
      } break;
      case PK_JOB_SCALED:
        v = pk_list_at(cx, Sb, b, jb.flags & PK_F_A_SCALAR, 0, jb.i[6], jb.i[7], jb.i[8], (long long)e) * sysv;
        if (jb.i[4]) v = v * post;
        break;
      case PK_JOB_SYS:
        v = sysv;
        if (jb.i[4]) v = v * post;
        break;
      default: {  // OUTER / TRIL
        long long ia, ib;
        if (type == PK_JOB_OUTER) {
          ia = (long long)(e / (IT)jb.i[12]);
          ib = (long long)e - ia * jb.i[12];
        } else {
          ia = (long long)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
          while (ia * (ia + 1) / 2 > e) --ia;
          while ((ia + 1) * (ia + 2) / 2 <= e) ++ia;
          ib = (long long)e - ia * (ia + 1) / 2;
        }
        const double va = pk_list_at(cx, Sb, b, jb.flags & PK_F_A_SCALAR, jb.flags & PK_F_A_UNIT, jb.i[6], jb.i[7], jb.i[8], ia);
        const double vb = pk_list_at(cx, Sb, b, jb.flags & PK_F_B_SCALAR, jb.flags & PK_F_B_UNIT, jb.i[9], jb.i[10], jb.i[11], ib);
        v = (va * vb) * sysv;
        if (jb.i[4]) v = v * post;
      }
    }
    pk_store(out0 + (long long)b * cx.n_out + (long long)e, v, cx.stream);
  }
}

// ---------------------------------------------------------------------------------------------
// gradient: each output column sums its contributions in list order (np.add.at semantics).
// A thread owns PK_ITEMS columns PK_THREADS apart: the chains  contribution record -> node-table value
// of its columns are independent and in flight together (the one-column version was bound by that
// dependent-load latency: humanoid 100 k nodes 13.4 us for 20 MB of traffic).
template <typename IT>
__global__ void __launch_bounds__(PK_THREADS) pk_grad_range(PkCtx cx, const pk_job* __restrict__ jobs, int B) {
  const pk_job& jb = jobs[blockIdx.y];
  const IT cnt = (IT)jb.i[1];
  const IT total = cnt * (IT)B;
  const IT g0 = (IT)blockIdx.x * (IT)PK_CHUNK + threadIdx.x;
  if (g0 >= total) return;
  const long long* __restrict__ q0 = cx.ipool + jb.i[2];
  const int n_contrib = (int)jb.i[3];
  int b[PK_ITEMS];
  long long e[PK_ITEMS];
  double acc[PK_ITEMS];
#pragma unroll
  for (int it = 0; it < PK_ITEMS; ++it) {
    const IT gid = g0 + (IT)it * PK_THREADS;
    const IT g = gid < total ? gid : g0;  // dead items recompute item 0 (not stored)
    b[it] = (int)(g / cnt);
    e[it] = (long long)(g - (IT)b[it] * cnt);
    acc[it] = 0.0;
  }
  const long long* q = q0;
  for (int n = 0; n < n_contrib; ++n, q += 4) {
    const long long w0 = q[0], lm = q[1], c_lo = q[2], sys = q[3];
#pragma unroll
    for (int it = 0; it < PK_ITEMS; ++it)
      acc[it] += cx.W[w0 + (long long)b[it] * lm + c_lo + e[it]] * cx.S[(long long)b[it] * cx.n_scalar + sys];
  }
#pragma unroll
  for (int it = 0; it < PK_ITEMS; ++it)
    if (g0 + (IT)it * PK_THREADS < total) pk_store(cx.OUT + (long long)b[it] * cx.n_out + jb.i[0] + e[it], acc[it], cx.stream);
}

__global__ void __launch_bounds__(PK_THREADS) pk_grad_scalar(PkCtx cx, const pk_job* __restrict__ jobs, int n_jobs, int B) {
  const long long gid = blockIdx.x * (long long)PK_THREADS + threadIdx.x;
  if (gid >= (long long)n_jobs * B) return;
  const int j = (int)(gid % n_jobs), b = (int)(gid / n_jobs);
  const pk_job& jb = jobs[j];
  const double* Sb = cx.S + (long long)b * cx.n_scalar;
  const long long* q = cx.ipool + jb.i[2];
  double acc = 0.0;
  for (long long n = 0; n < jb.i[3]; ++n, q += 2) acc += (q[0] >= 0 ? Sb[q[0]] * Sb[q[1]] : Sb[q[1]]);
  cx.OUT[(long long)b * cx.n_out + jb.i[0]] = acc;
}

// ---------------------------------------------------------------------------------------------
// de-duplicated pattern (opt-in): the reference's COO patterns repeat a (row, col) pair once per
// contributing list entry and leave the summation to the consumer (Ipopt triplets, scipy.py:13-29).
// With compaction enabled the engine sums the duplicates itself -- OUTC[u] = sum of OUT[perm[j]] for
// j in [ptr[u], ptr[u+1]), in increasing slot order (the order a consumer walking the triplets would
// use) -- so only the unique entries cross PCIe.  The values were written a moment ago and are
// still L2-resident; one thread per unique entry, loads pipelined four deep.
__global__ void __launch_bounds__(PK_THREADS) pk_compact(const double* __restrict__ OUT, double* __restrict__ OUTC,
                                                        const unsigned* __restrict__ ptr, const unsigned* __restrict__ perm,
                                                        long long n_out, long long n_unique, int B) {
  const long long gid = blockIdx.x * (long long)PK_THREADS + threadIdx.x;
  if (gid >= n_unique * B) return;
  const int b = (int)(gid / n_unique);
  const long long u = gid - (long long)b * n_unique;
  const double* src = OUT + (long long)b * n_out;
  unsigned j = ptr[u];
  const unsigned hi = ptr[u + 1];
  double acc = 0.0;
  for (; j + 4 <= hi; j += 4) {
    const double v0 = src[perm[j]], v1 = src[perm[j + 1]], v2 = src[perm[j + 2]], v3 = src[perm[j + 3]];
    acc += v0; acc += v1; acc += v2; acc += v3;
  }
  for (; j < hi; ++j) acc += src[perm[j]];
  OUTC[(long long)b * n_unique + u] = acc;
}

// ---------------------------------------------------------------------------------------------
// CSR matrix-vector product with sequential row sums (the order of scipy.sparse csr_matvec):
// y[g][r] = (sum_k val[k] * x[g][idx[k]]) [* (hi[0] - lo[0])] for g < groups.  Used by the continuous
// error estimate (phasebase.py:1339-1366): interpolation to the augmented mesh, translation and
// augmented integration are applied exactly like the reference's csr.dot.
__global__ void __launch_bounds__(PK_THREADS) pk_csr_matvec(const long long* __restrict__ ptr, const long long* __restrict__ idx,
                                                           const double* __restrict__ val, const double* __restrict__ x,
                                                           double* __restrict__ y, long long n_rows, long long x_stride,
                                                           long long y_stride, int groups, const double* __restrict__ hi,
                                                           const double* __restrict__ lo) {
  const long long gid = blockIdx.x * (long long)PK_THREADS + threadIdx.x;
  if (gid >= n_rows * groups) return;
  const int g = (int)(gid / n_rows);
  const long long r = gid - (long long)g * n_rows;
  const double* xv = x + (long long)g * x_stride;
  double acc = 0.0;
  for (long long k = ptr[r]; k < ptr[r + 1]; ++k) acc += val[k] * xv[idx[k]];
  if (hi) acc = acc * (hi[0] - lo[0]);
  y[(long long)g * y_stride + r] = acc;
}

// L2 flush helper for timing hygiene
__global__ void pk_fill(double* p, long long n, double v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
