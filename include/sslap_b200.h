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
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sslapb_handle sslapb_handle;

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
                                eCE + phase change, cluster regime, grid barriers */
    int64_t prune_second_pass; /* grid-regime rows whose bound-pruned sweep needed the second (exactness) gather pass */
    int32_t stop_reason;     /* 1 target-eps CS holds (:275) | 2 eps < target (:280) | 3 max_iter (:309) */
    int32_t rounds_cluster;  /* rounds run by cluster 0 alone (mid-sized frontiers); the others are in rounds_grid/warp/solo */
} sslapb_meta;

int  sslapb_create(int device, sslapb_handle **out);
void sslapb_destroy(sslapb_handle *h);
const char *sslapb_last_error(const sslapb_handle *h);
/* tuning knobs: "t_small" (frontier size at or below which CTA 0 runs rounds alone, 0..32), "t_cluster" (frontier size
   at or below which one thread-block cluster of 8 CTAs runs the rounds with hardware cluster barriers; 0 = off, the
   default — opt-in, honoured only while t_small is 32), "watchdog_ms" (device watchdog of a single barrier wait, default 120000) */
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
 * Kernel-level entry used by the parity tests and the roofline measurement: one bidding sweep
 * (bid_and_assign's bidding loop, auction_.pyx:339-365) over the CSR of the most recent problem on this handle.
 *   prices (n_cols, host, NULL = keep the handle's current prices), bidders (nb int32, host, NULL = persons 0..nb-1)
 *   merge bit 0: also perform the per-object atomicMax of the bids (:375-385); bit 1: disable the bound pruning of
 *   the price gathers (A/B measurement); bit 2 (only with bidders == NULL, nb == n_rows): the streamed TMA-ring variant
   of the sweep instead of the per-row kernel (same results);  iters >= 1 timed launches, with an L2
 *   flush (a write larger than L2) before each when flush_l2 != 0.
 *   jbest_out / bid_out (nb, host, may be NULL); *avg_ms_out = mean device time of one launch (CUDA events).
 */
int sslapb_bid_sweep(sslapb_handle *h, const double *prices, const int32_t *bidders, int32_t nb, float eps, int merge,
                     int iters, int flush_l2, int32_t *jbest_out, double *bid_out, float *avg_ms_out);

#ifdef __cplusplus
}
#endif
#endif
