/*
 * pockit_b200.h -- C ABI of the B200 evaluation engine for pockit's NLP callbacks.
 *
 * This is the drop-in boundary for the hot path: every pk_eval_* entry point
 * replaces one SystemBase callback of the reference (file:line cited per
 * function, all in /root/reference/pockit/base/systembase.py).  Plain pointers
 * and sizes only; host buffers are caller-owned (NumPy arrays on the Python
 * side, pinned when obtained from pk_alloc_host).  All functions return 0 on
 * success; on failure pk_last_error() describes what happened.  An engine
 * owns one CUDA stream and is not thread-safe; independent engines may run
 * concurrently (batched sweeps, one process per GPU).
 *
 * The engine is *data driven*: the host planner (pockit_b200/plan.py) lowers a
 * model to
 *   - CUDA C for the per-node programs (JIT-compiled here with NVRTC for sm_100a),
 *   - a table of jobs for the hand-written expansion / reduction kernels,
 *   - constant pools (integration blocks, mesh fractions, weights).
 */
#ifndef POCKIT_B200_H
#define POCKIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_ABI_VERSION 2

typedef struct pk_engine pk_engine;

/* evaluation modes = the reference callbacks */
enum {
  PK_MODE_OBJECTIVE = 0,   /* SystemBase.objective   systembase.py:602 */
  PK_MODE_CONSTRAINTS = 1, /* SystemBase.constraints systembase.py:613 */
  PK_MODE_GRADIENT = 2,    /* SystemBase.gradient    systembase.py:646 */
  PK_MODE_JACOBIAN = 3,    /* SystemBase.jacobian    systembase.py:676 */
  PK_MODE_HESSIAN = 4,     /* SystemBase.hessian     systembase.py:820 (hessian_o :735, hessian_c :786) */
  PK_N_CALLBACKS = 5,
  /* several callbacks at one (x, lambda, sigma) as ONE pipeline: a single per-node program evaluates
   * every leaf once, one reduction and one system program feed all consumers; outputs in one buffer
   * [objective | gradient | constraints | Jacobian values | Hessian values] (the covered ones:
   * pk_mode_desc.sub_count > 0 -- all five, or the three small callbacks).  Used by pk_run_set /
   * pk_time_steps when every covered callback is requested and this mode is loaded. */
  PK_MODE_SET = 5,
  PK_N_MODES = 6
};

/* job stages (which hand-written kernel consumes the record) */
enum {
  PK_STAGE_REDUCE = 0,      /* row sums of the node table -> scalar table (quadrature, np.add.at on broadcast columns) */
  PK_STAGE_DEFECT = 1,      /* T.x - dt * (I.f)           phasebase.py:1008-1012 */
  PK_STAGE_GENERIC = 2,     /* const / kron / table-expand / scaled / sys / outer / tril slot runs */
  PK_STAGE_EXPAND = 3,      /* block-structured expansion through the integration operator, phasebase.py:1120-1124, 1280-1285 */
  PK_STAGE_GRAD_RANGE = 4,  /* gather-sum into per-node gradient columns, systembase.py:654-656 */
  PK_STAGE_GRAD_SCALAR = 5, /* gather-sum into scalar gradient columns */
  PK_N_STAGES = 6
};

/* generic job types (PK_STAGE_GENERIC) */
enum {
  PK_JOB_CONST = 0,
  PK_JOB_KRON = 1,
  PK_JOB_EXPAND_TABLE = 2,
  PK_JOB_SCALED = 3,
  PK_JOB_SYS = 4,
  PK_JOB_OUTER = 5,
  PK_JOB_TRIL = 6
};

/* One record for every stage; the meaning of i[] / f[] per stage and type is
 * documented next to the kernels in pockit_b200/csrc/pk_kernels.cuh. */
typedef struct {
  int32_t type;
  int32_t flags;
  int64_t i[16];
  double f[2];
} pk_job;

/* problem dimensions, fixed for the life of an engine */
typedef struct {
  int32_t abi_version; /* PK_ABI_VERSION */
  int32_t batch;       /* B: independent instances evaluated per call (>= 1) */
  int64_t L;           /* variables per instance                     (SystemBase.L) */
  int64_t m;           /* constraints per instance                   (len(c_lb)) */
  int64_t nnz_jac;     /* len(jacobianstructure()[0]) */
  int64_t nnz_hess;    /* len(hessianstructure()[0]), objective part first */
  int64_t n_scalar;    /* scalar-table slots per instance */
  int64_t n_table;     /* node-table doubles in total (all rows, all instances) */
  int64_t n_fixed;     /* FIXED boundary values per instance */
} pk_dims;

/* one per-node program (one phase) of a mode */
typedef struct {
  const char *kernel;  /* NVRTC kernel name */
  int64_t n_nodes;     /* collocation nodes L_m of the phase */
  int64_t tm_offset;   /* offsets into the double pool: mesh fractions, quadrature weights */
  int64_t wm_offset;
} pk_node_program;

typedef struct {
  const char *cuda_source;           /* all node programs + the system program of this mode */
  const char *const *nvrtc_options;  /* extra options (the engine adds -arch and the std ones) */
  int32_t n_nvrtc_options;
  int32_t n_node_programs;
  const pk_node_program *node_programs;
  const char *system_kernel;         /* NULL: no system-level function in this mode */
  const char *table_symbol;          /* __constant__ long long[] the programs index */
  const int64_t *table;
  int64_t n_table_entries;
  int64_t n_scalar;                  /* scalar-table stride of this mode (<= pk_dims.n_scalar) */
  int64_t n_out;                     /* outputs per instance */
  const pk_job *jobs[PK_N_STAGES];
  int64_t n_jobs[PK_N_STAGES];
  int64_t grad_offset, grad_count;     /* output slots the gradient gather-sums into (zeroed first); count 0: none */
  int64_t sub_offset[PK_N_CALLBACKS];  /* PK_MODE_SET: where each callback's values start in the combined output */
  int64_t sub_count[PK_N_CALLBACKS];   /* ... and how many there are (0 for the single-callback modes) */
} pk_mode_desc;

int pk_abi_version(void);
const char *pk_last_error(void);
int pk_device_count(int *count);

int pk_engine_create(const pk_dims *dims, int device, pk_engine **out);
int pk_engine_destroy(pk_engine *e);

/* constant pools shared by all modes (uploaded once per discretisation) */
int pk_engine_set_pools(pk_engine *e, const double *dpool, int64_t n_double, const int64_t *ipool, int64_t n_int);
/* FIXED boundary values, instance-major [batch][n_fixed] */
int pk_engine_set_fixed(pk_engine *e, const double *values);
/* compile (NVRTC, sm_100a) and install one mode; returns the build log through pk_last_error on failure */
int pk_engine_load_mode(pk_engine *e, int mode, const pk_mode_desc *desc);
/* cubin of a loaded mode (for caching / inspection); *size = 0 if not loaded */
int pk_engine_get_cubin(pk_engine *e, int mode, const void **data, size_t *size);

/* ---- host-to-host callbacks: x (and multipliers) in, values out; instance-major for batch > 1 ---- */
int pk_eval_objective(pk_engine *e, const double *x, double *f /* [B] */);
int pk_eval_gradient(pk_engine *e, const double *x, double *grad /* [B][L] */);
int pk_eval_constraints(pk_engine *e, const double *x, double *c /* [B][m] */);
int pk_eval_jacobian(pk_engine *e, const double *x, double *values /* [B][nnz_jac] */);
int pk_eval_hessian(pk_engine *e, const double *x, const double *lambda /* [B][m] */,
                    const double *sigma /* [B] */, double *values /* [B][nnz_hess] */);

/* All requested callbacks at one x in a single call (the "evaluate the whole set at this x" entry
 * point of an x-keyed cache in a solver adapter; Ipopt asks for f, grad f, g, J, H at the same x,
 * optimizer/ipopt.py:41-53): x and the multipliers cross PCIe once, the modes run concurrently and
 * every result is copied back as soon as its mode finishes.  outs[k] receives mode modes[k];
 * lambda / sigma are only read when PK_MODE_HESSIAN is among the modes.  x = NULL evaluates at the
 * point already resident on the device (the previous pk_eval_* / pk_upload_x): the cache's "same x,
 * another callback" case costs no upload. */
int pk_eval_set(pk_engine *e, const double *x, const double *lambda, const double *sigma, const int *modes,
                int n_modes, double *const *outs);
/* same, returning once everything is enqueued (results complete after pk_sync): for callers whose inputs
 * arrive in stages -- a mesh-shard worker starts the Jacobian when x is there, the Hessian when lambda is */
int pk_eval_set_async(pk_engine *e, const double *x, const double *lambda, const double *sigma, const int *modes,
                      int n_modes, double *const *outs);

/* ---- optional output shaping (both leave the reference pattern contract when enabled) ----
 * Mesh sharding (one fine mesh split over several GPUs): this engine was given only a
 * share of the slot-run jobs; `runs` = n_runs (offset, count) pairs of the output it computes.
 * pk_download / pk_eval_* then copy just those runs, to the same offsets of the host buffer
 * (a buffer shared by all ranks ends up complete).  n_runs = 0 restores the full copy. */
int pk_engine_set_output_runs(pk_engine *e, int mode, const int64_t *runs, int64_t n_runs);
/* De-duplicated pattern: the reference's COO patterns repeat (row, col) pairs and leave the sum to
 * the consumer (optimizer/scipy.py:13-29).  With a compaction table the engine sums the duplicates
 * on the device -- unique entry u = sum of slots perm[seg_ptr[u] .. seg_ptr[u+1]) in that order --
 * and the callbacks return n_unique values per instance.  n_unique = 0 switches it off. */
int pk_engine_set_compaction(pk_engine *e, int mode, int64_t n_unique, const int64_t *seg_ptr, const int64_t *perm);
int pk_out_size(pk_engine *e, int mode, int64_t *n); /* values per instance the host receives for this mode */

/* ---- continuous error-estimate data on the augmented mesh (the next host hot spot of an hp-adaptive
 * loop; PhaseBase._error_estimation_data_continuous, phasebase.py:1355-1366) ----
 * Per phase: xs = x with boundary values substituted (generated `prep` program), xu_aug = V.xs,
 * f_i at the augmented nodes (generated `node` program), T_x_aug = T.xs, I_f_aug = dt * (I.f_i).
 * The three operators are CSR matrices applied with sequential row sums, like the reference's csr.dot.
 * Engines with batch == 1 only. */
typedef struct {
  const char *prep_kernel, *node_kernel; /* NVRTC kernel names */
  int64_t x_offset;                      /* first entry of the phase in x (l_p) */
  int64_t L, L_xu, L_x_all;              /* phase vector length; state+control slots; state slots */
  int64_t n_x, n_u, Lm_aug, rows;        /* augmented nodes; rows per state of T.x and I.f */
  const double *tm_aug;                  /* [Lm_aug] mesh fractions of the augmented nodes */
  const int64_t *V_ptr, *V_idx; const double *V_val; /* ((n_x+n_u)*Lm_aug) x L_xu */
  const int64_t *T_ptr, *T_idx; const double *T_val; /* (n_x*rows) x L_x_all */
  const int64_t *I_ptr, *I_idx; const double *I_val; /* rows x Lm_aug */
} pk_aug_phase;
int pk_engine_load_error_estimate(pk_engine *e, const char *cuda_source, const char *const *nvrtc_options,
                                  int n_nvrtc_options, const pk_aug_phase *phases, int n_phases);
/* t_x / i_f: phases concatenated, each [n_x][rows] */
int pk_eval_error_data(pk_engine *e, const double *x, double *t_x, double *i_f);

/* ---- device-resident path (inputs already in HBM): upload once, run many, download ---- */
int pk_upload_x(pk_engine *e, const double *x);
int pk_upload_multipliers(pk_engine *e, const double *lambda, const double *sigma);
int pk_run(pk_engine *e, int mode);               /* enqueue the mode's kernels on the engine stream */
/* enqueue several callbacks at the same x as one CUDA graph: each mode on its own stream (private
 * tables and outputs), forked from / joined to the engine stream -- the "evaluate the whole set at
 * this x" entry point an x-keyed cache in the solver adapter calls */
int pk_run_set(pk_engine *e, const int *modes, int n_modes);
int pk_sync(pk_engine *e);
int pk_download(pk_engine *e, int mode, double *out);
/* `count` values from slot `offset` of every instance, packed [B][count]: hessian_o / hessian_c
 * (systembase.py:735, 786) are the head / tail of the Hessian values -- one evaluation, only the
 * requested part crosses PCIe */
int pk_download_range(pk_engine *e, int mode, int64_t offset, int64_t count, double *out);
/* device address of a mode's latest result for device-side consumers (e.g. an NCCL all-gather of
 * instance-sharded batches): *count values per instance, instances *stride doubles apart */
int pk_out_device_pointer(pk_engine *e, int mode, void **ptr, int64_t *count, int64_t *stride);
/* time `iters` back-to-back runs with CUDA events on the engine stream; ms_total covers the
 * whole mode, ms_stage[s] the kernels of each stage (s = 0..PK_N_STAGES-1 jobs, PK_N_STAGES = node
 * programs, PK_N_STAGES+1 = system program), measured in separate passes. */
int pk_time(pk_engine *e, int mode, int iters, float *ms_total, float *ms_stage /* [PK_N_STAGES+2] */);
/* one stage (bit s of stage_mask, as in pk_time) launch by launch, L2 flushed (untimed) before each:
 * the dominant kernel under the cache conditions of the whole-set measurement */
int pk_time_stage(pk_engine *e, int mode, unsigned stage_mask, int iters, int flush_l2, float *ms_each);
/* the same stage of several modes in turn, `rounds` times back to back (one event pair): e.g. the Jacobian
 * and the Hessian expansion alternating as inside a set, outputs together larger than L2 */
int pk_time_stage_alternating(pk_engine *e, const int *modes, int n_modes, unsigned stage_mask, int rounds, float *ms_total);
/* device-resident throughput: `steps` times { [flush L2, untimed]; event; pk_run_set(modes); event };
 * ms_steps[s] is the CUDA-event time of step s on the engine stream */
int pk_time_steps(pk_engine *e, const int *modes, int n_modes, int steps, int flush_l2, float *ms_steps);
/* device-side timeline of one evaluation set (per-mode streams, timing events around every launch):
 * rows[4*i..] = (mode, tag, edge, microseconds); tag = job stage 0..5, 6 node programs, 7 system
 * program, 8 compaction; edge 0 = before, 1 = after; the last row (mode -1) is the whole set */
int pk_timeline(pk_engine *e, const int *modes, int n_modes, double *rows, int max_rows, int *n_rows);
/* which block-expansion kernel a loaded mode uses: 0 none, 1 the persistent column walk for any
 * mix of orders [kernel pk_expand_blocks], 2 the parameter-driven column walk for same-order
 * meshes [kernel pk_expand_cols], 3 its opt-in TMA bulk-store variant [kernel pk_expand_bulk],
 * 4 the parameter-driven walk over the flattened (instance, pair) space of a batch [kernel pk_expand_batch;
 * POCKIT_B200_EXPAND=batch], 5 the slot-order batch kernel, a thread per output slot: the default for batch groups
 * of jobs with three or more lists each, 4 for the others [kernel pk_expand_slots] */
int pk_expand_variant(pk_engine *e, int mode, int *variant);
int pk_kernel_launches(pk_engine *e, int64_t *count);
/* process-wide cache of NVRTC results keyed by (architecture, options, source): the generated
 * programs take every mesh-dependent number from a table, so re-planning a re-meshed model compiles nothing */
int pk_cubin_cache_stats(int64_t *hits, int64_t *misses);
int pk_x_uploads(pk_engine *e, int64_t *count);       /* host-to-device copies of x so far */ /* kernels launched so far by this engine */
int pk_flush_l2(pk_engine *e);                        /* overwrite a buffer larger than L2 */

/* pinned host memory for zero-copy NumPy views */
void *pk_alloc_host(size_t bytes);
void pk_free_host(void *p);
/* page-lock / release a caller-owned range (a shared-memory mapping the ranks of a sharded mesh
 * copy their shares into); needs a current CUDA device */
int pk_host_register(void *p, size_t bytes);
int pk_host_unregister(void *p);

#ifdef __cplusplus
}
#endif
#endif /* POCKIT_B200_H */
