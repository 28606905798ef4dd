/*
 * sslap_b200 — C ABI of the B200-native (sm_100a) auction / Hopcroft-Karp path.
 *
 * This header is the drop-in boundary: plain C, pointers and sizes only.  Each entry point names the reference
 * interface it replaces (paths relative to the reference tree, OllieBoyne/sslap v0.2.5).  The Python package
 * `sslap_b200` binds it with ctypes (sslap_b200/_native.py); INTEGRATION.md shows the stub a maintainer of the
 * reference would add to route `sslap.auction_solve` / `sslap.hopcroft_solve` through it.
 *
 * Conventions
 *   - return value: 0 = ok; > 0 = a problem-level condition the reference reports as ValueError (see SSLAPB_E_*);
 *     < 0 = -(cudaError_t).  sslapb_last_error() gives the text.
 *   - `mem` flags say where caller buffers live (host by default).  Host buffers may be pageable or pinned.
 *   - one handle = one device + one stream + grow-only scratch in HBM; a handle is not thread-safe, distinct handles are.
 *   - indices are int32 or int64 (idx_bytes = 4 | 8), values are float64, exactly as the reference's typed buffers
 *     (auction_.pyx:601-602); caller memory is never modified (the reference negates `val` in place for 'min',
 *     auction_.pyx:236-237 — deliberately not reproduced).
 */
#ifndef SSLAP_B200_H
#define SSLAP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sslapb_handle sslapb_handle;

#define SSLAPB_ABI_VERSION 4      /* bumped whenever struct sslapb_meta or a prototype changes */

enum {
    SSLAPB_OK = 0,
    SSLAPB_E_FEWER_THAN_N = 1,   /* "Matrix is infeasible - Fewer than N valid values ..." auction_.pyx:559-560,604-605 */
    SSLAPB_E_CARDINALITY = 2,    /* "Matrix is infeasible (Maximum matching possible only involves c out of N rows.)" :565-566,611-612 */
    SSLAPB_E_UNSORTED = 3,       /* loc rows not non-decreasing: the reference's silent precondition (auction_.pyx:33-48) */
    SSLAPB_E_BAD_ARG = 4,
    SSLAPB_E_OUT_OF_RANGE = 5,   /* an index outside [0,n_rows) x [0,n_cols) */
    SSLAPB_E_EMPTY_ROW = 6,      /* cardinality_check off and some row has no entry (UB in the reference) */
    SSLAPB_E_ABORTED = 7         /* device watchdog fired */
};

enum {
    SSLAPB_MEM_HOST = 0,
    SSLAPB_MEM_DEVICE_IN = 1,    /* rows/cols/val (or mat) are device pointers on the handle's device */
    SSLAPB_MEM_DEVICE_OUT = 2    /* sol / pairings outputs are device pointers */
};

/* AuctionSolver.meta, auction_.pyx:264,297-304 — unrounded; the Python layer applies the reference's round(.,3). */
typedef struct sslapb_meta {
    float   start_eps;       /* :264 */
    float   final_eps;       /* :303 */
    float   target_eps;      /* :247 */
    int32_t eCE;             /* :297 */
    int32_t soln_found;      /* :300 */
    int64_t its;             /* :298 */
    int64_t nreductions;     /* :299 */
    int64_t n_assigned;      /* :301 */
    float   obj;             /* :302 get_obj() returns a C float (:489) */
    double  obj64;           /* the same sum before the float32 cast */
    float   setup_ms;        /* timer['setup'] (:206-207,265): CSR build + state init, device time */
    float   solve_ms;        /* timer['solve'] (:270,294): the persistent auction kernel, device time */
    float   hk_ms;           /* Hopcroft-Karp check, device+host loop time (not timed by the reference) */
    float   h2d_ms;          /* host->device staging of the inputs */
    int32_t cardinality;     /* result of the feasibility check, -1 if it did not run */
    int32_t n_rows, n_cols;  /* N, M actually used (inferred as max+1 when passed <= 0, auction_.pyx:209-210) */
    int64_t nnz;
    int64_t rounds_grid, rounds_warp, rounds_solo;   /* rounds executed per regime */
    float   prof_ms[8];      /* device time by section: grid bid, grid assign, grid compaction, warp regime, solo regime,
                                eCE + phase change, mid regime, grid barriers */
    int64_t prune_second_pass; /* grid-regime rows whose bound-pruned sweep needed the second (exactness) gather pass */
    int32_t stop_reason;     /* 1 target-eps CS holds (:275) | 2 eps < target (:280) | 3 max_iter (:309) */
    int32_t small_path;      /* 1 when the single-launch path for small problems ran (small.cu); the field replaces round 1's
                                rounds_cluster, same size and place */
    /* ---- ABI version 2 ---- */
    int32_t n_ranks, rank;   /* row-sharded solve (sslapb_comm_init): size of the communicator and this handle's rank; 1, 0 otherwise */
    int32_t row_lo, row_hi;  /* ... the nnz-balanced row range [row_lo, row_hi) this rank bids for in sharded rounds */
    int64_t rounds_sharded;  /* ... rounds (of rounds_grid) whose bidding step was split over the ranks */
    float   xchg_ms;         /* ... device time CTA 0 spent in the cross-GPU exchange barrier of those rounds (signal + wait) */
    float   sharded_ms;      /* ... device time of those rounds in total */
    float   sweep_insitu_us; /* mean device time of the bidding step of the full-frontier rounds (every person bids: first round
                                of each eps-phase) inside the persistent kernel, incl. the barrier that ends it */
    int32_t sweep_insitu_n;  /* number of such rounds */
    int32_t warm_start;      /* 1 when the solve started from caller-supplied prices (sslapb_set_prices) */
    int32_t strict;          /* 1 when the strict-optimality stop rule was on (option "strict") */
    /* ---- ABI version 3: hot lists (the 32 largest entries of every row decide a bid when that is provably exact) */
    int64_t hot_grid_bids, hot_grid_fallbacks;   /* grid-regime bids decided by the hot list / handed on to the full-row sweep */
    int64_t hot_tail_rounds, hot_tail_fallbacks; /* few-bidder + chain rounds run in hot form / their bids handed on to the full row */
    int64_t rounds_nohole;   /* rounds (of rounds_grid) that found no unowned object and no equal bids: no compaction, no final barrier */
    /* ---- ABI version 4: mid regime (33..t_mid bidders in a hot-list phase: CTA 0 alone, block barriers; prof_ms[6] is its time) */
    int64_t rounds_mid;      /* rounds executed by the mid regime; its = rounds_grid + rounds_mid + rounds_warp + rounds_solo */
} sslapb_meta;

int  sslapb_create(int device, sslapb_handle **out);
void sslapb_destroy(sslapb_handle *h);
const char *sslapb_last_error(const sslapb_handle *h);
/* Layout guards for foreign-language bindings: sizeof(struct sslapb_meta) of THIS build and SSLAPB_ABI_VERSION.  A stub must
   assert both before its first call (a shorter struct on the caller's side would be overrun by the library). */
size_t sslapb_meta_size(void);
int    sslapb_abi_version(void);
/* tuning knobs (none changes results): "t_small" (frontier size at or below which CTA 0 runs rounds alone, 0..32),
   "t_mid" (mid regime: in eps-phases whose bids the hot lists decide, CTA 0 also runs the rounds of 33..t_mid bidders alone,
   with block barriers instead of grid barriers; 0 = off, 33..256, default 128; only in effect with t_small = 32; in a
   row-sharded solve capped at t_shard: the ranks switch the hot form on and off independently, so only rounds that every
   rank runs in full may depend on it),
   "watchdog_ms" (device watchdog of a single barrier wait, default 120000),
   "t_shard" (row-sharded solves: rounds with more bidders than this are split over the ranks; default 16384),
   "max_ctas" (upper bound of the persistent kernel's grid, 0 = one CTA per SM; used to co-schedule several solves on one GPU),
   "hk_host_loop" (1: Hopcroft-Karp phases driven from the host with one read-back per BFS level, as in round 1, instead of the
   device-resident loop; A/B runs; default 0),
   "batch_v1" (1: round 1's batch kernel — a whole warp sweeps one bidder at a time — instead of the sub-warp kernel; A/B runs),
   "small_path" (0: never take the single-launch path for small problems — host buffers, default regime options, at most
   12288 entries; tests of the general path on small inputs; default 1), "small_max_n" (largest N, M that take it: 1..256,
   default 128),
   "hot" (0: never decide bids from the hot lists — A/B runs; default 1),
   "l2_persist" (1: persisting L2 access-policy window over the hot lists during a solve — A/B runs; measured no gain; default 0),
   "coop" (row-sharded solves only; 0: launch the persistent kernel without the cooperative attribute so that several of them
   can run side by side on ONE GPU — the driver runs one cooperative kernel at a time; only for the virtual-rank test, default 1),
   "strict" (1: strict-optimality stop rule — eps-CS is tested with eps = 1/(N+1) and zero tolerance and the eps schedule runs
   until eps < 1/(N+1), which makes the result provably optimal for integer costs; 0 = the reference's rule, default) */
int  sslapb_set_option(sslapb_handle *h, const char *name, int64_t value);

/* Pinned host memory for callers that want asynchronous staging (bench.py's e2e leg). */
void *sslapb_host_alloc(size_t bytes);
void  sslapb_host_free(void *p);

/*
 * auction_solve(loc=, val=) / auction_solve(coo_mat=)  — replaces _from_sparse + AuctionSolver.__init__ + .solve()
 * (auction_solve.py:45-50, auction_.pyx:575-617, :202-306).
 *   rows/cols : element k of the COO stream is rows[k*stride], cols[k*stride]  (stride 2 + cols == rows+1 for an
 *               interleaved (K,2) `loc`; stride 1 for separate arrays).  Must be row-sorted (else SSLAPB_E_UNSORTED).
 *   n_rows/n_cols <= 0 : infer max+1 like AuctionSolver.__init__ (:209-210).
 *   maximize  : 0 <=> problem == 'min'.   eps_start > 0 overrides C/2 (:251-252); `fast` is eps_start = float32(1/N).
 *   sol_out   : n_rows int32, -1 = unassigned (only when max_iter was hit).
 */
int sslapb_auction_coo(sslapb_handle *h, const void *rows, const void *cols, int idx_bytes, int64_t stride,
                       const double *val, int64_t nnz, int32_t n_rows, int32_t n_cols, int maximize, float eps_start,
                       int64_t max_iter, int cardinality_check, int mem, int32_t *sol_out, sslapb_meta *meta);

/* auction_solve(mat=) — replaces _from_matrix (auction_.pyx:528-571): row-major float64, entry valid iff >= 0. */
int sslapb_auction_dense(sslapb_handle *h, const double *mat, int32_t n_rows, int32_t n_cols, int maximize,
                         float eps_start, int64_t max_iter, int cardinality_check, int mem, int32_t *sol_out,
                         sslapb_meta *meta);

/*
 * hopcroft_solve(loc=) / c_hopcroft_solve — replaces HopcroftKarpSolverCython (feasibility_.pyx:95-225, :244-247).
 *   left_out (n_rows) / right_out (n_cols) may be NULL; *size_out = cardinality of a maximum matching.
 */
int sslapb_hopcroft_coo(sslapb_handle *h, const void *rows, const void *cols, int idx_bytes, int64_t stride,
                        int64_t nnz, int32_t n_rows, int32_t n_cols, int mem, int32_t *left_out, int32_t *right_out,
                        int32_t *size_out);

/* hopcroft_solve(mat=) — replaces the dense adapter (feasibility_.pyx:249-266). */
int sslapb_hopcroft_dense(sslapb_handle *h, const double *mat, int32_t n_rows, int32_t n_cols, int mem,
                          int32_t *left_out, int32_t *right_out, int32_t *size_out);

/*
 * Batch of independent problems (BASELINE.json configs[4]; the reference has no batch call — this replaces a Python loop
 * of n_problems auction_solve(loc=, val=, cardinality_check=False) calls, auction_solve.py:45-46).
 *   Problem p owns COO entries [nnz_offsets[p], nnz_offsets[p+1]) of rows/cols/val (indices LOCAL to the problem,
 *   row-sorted inside the problem) and has n_rows[p] x n_cols[p] shape; nnz_offsets[0] == 0.
 *   eps_start: per problem (NULL or <= 0 entries: C/2).  sol_out: concatenated, n_rows[p] entries per problem, local
 *   column ids.  metas: n_problems entries (timings are those of the whole batch).  One warp solves one problem; every
 *   problem's trajectory (sol, its, meta) is bit-identical to a single-problem call.  No feasibility check is run.
 */
int sslapb_auction_batch(sslapb_handle *h, int32_t n_problems, const int64_t *nnz_offsets, const int32_t *n_rows,
                         const int32_t *n_cols, const void *rows, const void *cols, int idx_bytes, int64_t stride,
                         const double *val, int maximize, const float *eps_start, int64_t max_iter, int mem,
                         int32_t *sol_out, sslapb_meta *metas);

/* Prices of the most recent solve on this handle (AuctionSolver.p, auction_.pyx:169,220) — n_cols doubles, host. */
int sslapb_get_prices(sslapb_handle *h, double *prices_out);

/*
 * Warm start: the NEXT sslapb_auction_coo / _dense call on this handle starts from these object
 * prices instead of AuctionSolver.__init__'s zeros (auction_.pyx:220); consumed by that one call.  n_cols must equal the
 * problem's column count (else SSLAPB_E_BAD_ARG at the solve); prices are in the solver's internal (maximisation) frame,
 * i.e. exactly what sslapb_get_prices returned for a related problem.  prices == NULL clears a pending warm start.
 * Combine with eps_start (e.g. the final eps of the previous solve) to skip the coarse eps-phases.
 */
int sslapb_set_prices(sslapb_handle *h, const double *prices, int32_t n_cols);

/*
 * Row-sharded solve of ONE problem over several GPUs of a node (the reference has no counterpart: it
 * replaces the serial merge loop auction_.pyx:367-385 across devices).  SPMD contract: n_ranks handles (one per GPU, in
 * one or several processes) call sslapb_comm_init, exchange the export blobs by any host-side means (torch.distributed,
 * MPI, a pipe), call sslapb_comm_connect with all n_ranks blobs in rank order, and from then on make the SAME sequence
 * of sslapb_auction_coo / _dense calls with the SAME full problem.  Every rank builds the whole CSR and keeps the whole
 * state; persons are split into nnz-balanced contiguous row ranges (computed on the device from the CSR offsets), and in
 * rounds whose frontier exceeds option "t_shard" each rank sweeps only its own bidders and stores their bids straight
 * into every rank's exchange buffer over NVLink (peer-mapped memory: cudaIpc handles across processes, plain peer access
 * inside one); one in-kernel exchange barrier (system-scope flags) later every rank merges all bids with the same 64-bit
 * atomicMax and applies the identical assignment, so prices, sol and its stay bit-identical to the single-GPU solve on
 * every rank.  Smaller frontiers (the latency-bound tail) run redundantly on every rank with no communication.
 *   capacity_rows: largest n_rows the communicator will see (sizes the exchange buffers: 24 bytes per row per rank).
 *   export_out:    SSLAPB_COMM_EXPORT_BYTES bytes, opaque.
 */
#define SSLAPB_COMM_EXPORT_BYTES 128
#define SSLAPB_COMM_MAX_RANKS 8
int sslapb_comm_init(sslapb_handle *h, int n_ranks, int rank, int64_t capacity_rows, void *export_out);
int sslapb_comm_connect(sslapb_handle *h, const void *all_exports);
int sslapb_comm_destroy(sslapb_handle *h);

/*
 * Kernel-level entry used by the parity tests and the roofline measurement: one bidding sweep
 * (bid_and_assign's bidding loop, auction_.pyx:339-365) over the CSR of the most recent problem on this handle.
 *   prices (n_cols, host, NULL = keep the handle's current prices), bidders (nb int32, host, NULL = persons 0..nb-1)
 *   merge bit 0: also perform the per-object atomicMax of the bids (:375-385); bit 1: disable the bound pruning of
 *   the price gathers (A/B measurement); bit 2 (only with bidders == NULL, nb == n_rows): the streamed TMA-ring variant
 *   of the sweep; bit 7: the per-row kernel (one warp per row); bit 3: the software-pipelined per-row kernel (bits 4-6: its
 *   CTA size / untrimmed stage); default: four rows per warp, 8 lanes per row — all variants return identical results
 *   and exist for A/B measurement;  iters >= 1 timed launches, with an L2
 *   flush (a write larger than L2) before each when flush_l2 != 0.
 *   jbest_out / bid_out (nb, host, may be NULL); *avg_ms_out = mean device time of one launch (CUDA events).
 */
int sslapb_bid_sweep(sslapb_handle *h, const double *prices, const int32_t *bidders, int32_t nb, float eps, int merge,
                     int iters, int flush_l2, int32_t *jbest_out, double *bid_out, float *avg_ms_out);

#ifdef __cplusplus
}
#endif
#endif
