// C ABI of sslap_b200 (see include/sslap_b200.h).  Host-side orchestration only: staging, CSR build, the
// Hopcroft-Karp phase loop, one cooperative launch of the persistent auction kernel, and the meta read-back.
#include "../../include/sslap_b200.h"
#include "auction.cuh"
#include "build.cuh"
#include "small.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <chrono>
#include <mutex>
#include <condition_variable>
#include <map>
#include <memory>
#include <unistd.h>

// ---- kernels' launchers (csr_build.cu / auction.cu / hopcroft.cu)


extern "C" {
cudaError_t sslapb_launch_coo_ingest(const void *, const void *, int, long long, const double *, long long, int, int, int,
                                     int *, double *, long long *, SslapbBuildFlags *, int, cudaStream_t);
cudaError_t sslapb_launch_coo_sort(const void *, const void *, int, long long, const double *, long long, int, unsigned *,
                                   unsigned *, unsigned *, unsigned *, long long *, int *, int *, double *, int, cudaStream_t);
cudaError_t sslapb_launch_rowmax(const long long *, const double *, long long, double *, int *, int, cudaStream_t);
cudaError_t sslapb_launch_hot_build(const long long *, const int *, const double *, int, SslapbHotEnt *, double *, int, cudaStream_t);
cudaError_t sslapb_launch_index_max(const void *, const void *, int, long long, long long, long long *, int, cudaStream_t);
cudaError_t sslapb_launch_dense_count(const double *, int, int, long long *, SslapbBuildFlags *, int, cudaStream_t);
cudaError_t sslapb_launch_dense_fill(const double *, int, int, int, const long long *, int *, double *,
                                     SslapbBuildFlags *, int, cudaStream_t);
cudaError_t sslapb_launch_auction_init(const SslapbAuctionParams *, int, int, cudaStream_t);
cudaError_t sslapb_launch_auction(const SslapbAuctionParams *, int, int, int, cudaStream_t);
cudaError_t sslapb_launch_row_split(const long long *, int, int, int *, cudaStream_t);
int sslapb_coop_row_entries();
cudaError_t sslapb_auction_grid_size(int, int *);
cudaError_t sslapb_launch_bid_sweep(const SslapbAuctionParams *, const int *, int, float, int, int, cudaStream_t);
cudaError_t sslapb_launch_sweep_plan(const SslapbAuctionParams *, int, int *, cudaStream_t);
cudaError_t sslapb_launch_bid_sweep_tma(const SslapbAuctionParams *, const int *, float, int, int, cudaStream_t);
cudaError_t sslapb_launch_price_bounds(const SslapbAuctionParams *, cudaStream_t);
cudaError_t sslapb_launch_bid_sweep2(const SslapbAuctionParams *, const int *, int, float, int, int, int, int, cudaStream_t);
cudaError_t sslapb_launch_bid_sweep4(const SslapbAuctionParams *, const int *, int, float, int, int, cudaStream_t);
cudaError_t sslapb_launch_bid_sweep_hot(const SslapbAuctionParams *, const int *, int, float, int, int, cudaStream_t);
cudaError_t sslapb_launch_hot_rest(const SslapbAuctionParams *, int, cudaStream_t);
cudaError_t sslapb_launch_small(const SslapbSmallArgs *, cudaStream_t);
cudaError_t sslapb_launch_bid_sweep_lean(const SslapbAuctionParams *, const int *, int, float, int, int, cudaStream_t);
cudaError_t sslapb_hk_launch_greedy(const long long *, const int *, int, int *, int *, SslapbHkFlags *, int, cudaStream_t);
cudaError_t sslapb_hk_launch_phase_init(int, int, const int *, int *, int *, int *, int *, SslapbHkFlags *, int, cudaStream_t);
cudaError_t sslapb_hk_launch_bfs_level(const long long *, const int *, int, int, const int *, int *, int *, int *, int *,
                                       SslapbHkFlags *, int, cudaStream_t);
cudaError_t sslapb_hk_launch_augment(int, const int *, const int *, int *, int *, SslapbHkFlags *, int, cudaStream_t);
cudaError_t sslapb_hk_persistent_grid(int, int *);
cudaError_t sslapb_hk_launch_persistent(const long long *, const int *, int, int, int *, int *, int *, int *, int *, int *, int *,
                                        int *, int *, void *, int, cudaStream_t);
}

#include "batch.cuh"
extern "C" {
cudaError_t sslapb_launch_batch_globalize(const void *, const void *, int, long long, const long long *, const long long *,
                                          const long long *, int, int *, int *, int *, int, cudaStream_t);
cudaError_t sslapb_launch_auction_batch(const SslapbBatchParams *, cudaStream_t);
cudaError_t sslapb_launch_auction_batch2(const SslapbBatchParams *, cudaStream_t);
}

namespace {

// Launch gate of a communicator whose ranks all live in THIS process (several handles driven by threads, e.g. the
// virtual-rank test on one GPU): every rank finishes its allocations, memsets, copies and set-up kernels, then waits here,
// and only then launches its persistent kernel.  Device memory allocation / memset between two launches implicitly
// serialises streams of one device, and a rank that is still allocating while another rank's kernel already waits for
// its bids would deadlock until the watchdog.
struct LocalGate {
    std::mutex m;
    std::condition_variable cv;
    int n = 0, arrived = 0;
    unsigned long long gen = 0;
    bool arrive_and_wait(long long timeout_ms)
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long long g0 = gen;
        if (++arrived == n) { arrived = 0; ++gen; cv.notify_all(); return true; }
        if (cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return gen != g0; })) return true;
        --arrived;
        return false;
    }
};
std::mutex g_gates_mu;
std::map<unsigned long long, std::shared_ptr<LocalGate>> g_gates;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = (bytes + 255) & ~(size_t)255;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace

struct sslapb_handle {
    std::recursive_mutex mu;       // every entry point holds it: concurrent calls on ONE handle are serialised (the reference is
                                   // re-entrant under the GIL; ctypes releases the GIL, so the library must lock)
    int device = 0;
    int sms = 0;
    int grid = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {};
    std::string err;
    int t_small = 32;
    int t_mid = SSLAPB_MID < 128 ? SSLAPB_MID : 128;               // mid regime: CTA 0 alone runs rounds of 33..t_mid bidders in hot-list phases (0: off; only with t_small == 32)
    long long watchdog_ms = 120000;
    int t_shard = 16384;           // row-sharded solves: rounds with more bidders than this are split over the ranks
    int max_ctas = 0;              // upper bound of the persistent kernel's grid (0: one CTA per SM)
    int strict = 0;                // strict-optimality stop rule (see the header)
    int coop = 1;                  // 0: launch the row-sharded persistent kernel without the cooperative attribute (virtual ranks)
    int hot = 1;                   // hot lists (hot.cu): 0 = off (A/B runs)
    int small_path = 1;            // single-launch path for small problems (small.cu): 0 = off
    int small_max_n = 128;         // ... taken for N, M up to this (measured: it wins up to the dense 110 x 110 that fits; the kernel holds up to SSLAPB_SMALL_MAXN)
    int l2_persist = 0;            // 1: L2 access-policy window (persisting) over the hot lists during a solve (measured: no gain at C3,
                                   // 233.0 vs 233.1 ms — the lists stay in L2 on their own; kept for A/B runs)
    size_t l2_persist_max = 0;     // cudaDevAttrMaxPersistingL2CacheSize
    int l2_window_max = 0;         // cudaDevAttrMaxAccessPolicyWindowSize
    // warm start: prices for the next solve (sslapb_set_prices)
    DevBuf warm;
    int warm_cols = 0;
    // row-sharded communicator (sslapb_comm_init / _connect)
    int n_ranks = 1, rank = 0;
    bool comm_connected = false, comm_broken = false;
    long long xcap = 0;
    DevBuf xbuf, xtab, rowsplit;   // exchange buffer (flags | bidv[2][xcap] | bidj[2][xcap]), peer address table, row boundaries
    void *peer_base[8] = {};       // peer-mapped base address of every rank's exchange buffer
    bool peer_ipc[8] = {};         // opened with cudaIpcOpenMemHandle (to be closed)
    unsigned xround = 0;           // sharded rounds completed on this communicator
    std::shared_ptr<LocalGate> gate;   // all ranks in this process: rendezvous right before the persistent kernel's launch
    unsigned long long gate_key = 0;
    // resident problem
    int N = 0, M = 0;
    int maxdeg = 0;                // longest row (read back after the row-maximum pass)
    long long nnz = 0;
    bool has_vals = false;
    DevBuf stage_idx, stage_val, stage_mat, cols, vals, rowptr, rowmax, flags;
    DevBuf hotbuf;                 // rest[N] | hthr[N] | hot lists N x 512 B (one allocation: one L2 access-policy window covers it)
    bool hot_valid = false;        // hotbuf holds the lists of the resident CSR
    DevBuf sort_keys, sort_idx, sort_hist, sort_rows, sort_cols, sort_val;   // only for unsorted input
    DevBuf b_off, b_rows, b_cols, b_eps, b_meta, b_bad, b_rec;               // batched problems
    int batch_v1 = 0;              // option "batch_v1": 1 = round 1's batch kernel (whole warp per bidder), A/B runs
    // auction state
    DevBuf price, owner, p2o, list, mover, bidj, bidv, bidkey, winpos, hole_count, chosen, ctrl, bidders, flush, sweep_plan;
    // HK state
    DevBuf pair_u, pair_v, dist, visited, cursor, pred, hkflags, hkq;
    int hk_grid = 0;
    int hk_host_loop = 0;          // option "hk_host_loop": 1 = round 1's host-driven phase loop (A/B runs)
    int hk_phases = 0, hk_levels = 0;   // instrumentation of the last run
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
            return -(int)e_;                                                                             \
        }                                                                                                \
    } while (0)

#define LOCK(h) std::lock_guard<std::recursive_mutex> lock_((h)->mu)

static int fail(sslapb_handle *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    return code;
}

extern "C" int sslapb_create(int device, sslapb_handle **out)
{
    if (!out) return SSLAPB_E_BAD_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return -(int)e;
    if (device < 0 || device >= count) return SSLAPB_E_BAD_ARG;
    sslapb_handle *h = new sslapb_handle();
    h->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess) { delete h; return -(int)e; }
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device);
    if (!coop) { delete h; return SSLAPB_E_BAD_ARG; }
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) { delete h; return -(int)e; }
    for (auto &ev : h->ev) cudaEventCreate(&ev);
    if ((e = sslapb_auction_grid_size(device, &h->grid)) != cudaSuccess) { delete h; return -(int)e; }
    {   // L2 persistence for the hot lists of the tail rounds (run_auction sets the window per launch)
        int pmax = 0, wmax = 0;
        cudaDeviceGetAttribute(&pmax, cudaDevAttrMaxPersistingL2CacheSize, device);
        cudaDeviceGetAttribute(&wmax, cudaDevAttrMaxAccessPolicyWindowSize, device);
        h->l2_persist_max = (size_t)(pmax > 0 ? pmax : 0); h->l2_window_max = wmax > 0 ? wmax : 0;
        cudaGetLastError();
    }
    if ((e = sslapb_hk_persistent_grid(device, &h->hk_grid)) != cudaSuccess) { delete h; return -(int)e; }
    *out = h;
    return SSLAPB_OK;
}

extern "C" void sslapb_destroy(sslapb_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf *all[] = {&h->b_rec, &h->b_off, &h->b_rows, &h->b_cols, &h->b_eps, &h->b_meta, &h->b_bad, &h->sort_keys, &h->sort_idx, &h->sort_hist, &h->sort_rows, &h->sort_cols, &h->sort_val, &h->stage_idx, &h->stage_val, &h->stage_mat, &h->cols, &h->vals, &h->rowptr, &h->rowmax, &h->flags, &h->price,
                     &h->owner, &h->p2o, &h->list, &h->mover, &h->bidj, &h->bidv, &h->bidkey, &h->winpos,
                     &h->hole_count, &h->chosen, &h->ctrl, &h->bidders, &h->flush, &h->sweep_plan, &h->pair_u, &h->pair_v, &h->dist,
                     &h->visited, &h->cursor, &h->pred, &h->hkflags, &h->hkq, &h->hotbuf};
    for (DevBuf *b : all) b->release();
    sslapb_comm_destroy(h);                                    // peer mappings, exchange buffer, the process-local launch gate
    h->rowsplit.release(); h->warm.release();
    for (auto &ev : h->ev) cudaEventDestroy(ev);
    cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char *sslapb_last_error(const sslapb_handle *h) { return h ? h->err.c_str() : "null handle"; }

// Layout guard for foreign-language bindings: a stub whose struct differs from this size was written against another header.
extern "C" size_t sslapb_meta_size(void) { return sizeof(sslapb_meta); }
extern "C" int sslapb_abi_version(void) { return SSLAPB_ABI_VERSION; }

extern "C" int sslapb_set_option(sslapb_handle *h, const char *name, int64_t value)
{
    if (!h || !name) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    if (!strcmp(name, "t_small")) { if (value < 0 || value > 32) return SSLAPB_E_BAD_ARG; h->t_small = (int)value; return 0; }
    if (!strcmp(name, "t_mid")) { if (value != 0 && (value < 33 || value > SSLAPB_MID)) return SSLAPB_E_BAD_ARG; h->t_mid = (int)value; return 0; }
    if (!strcmp(name, "watchdog_ms")) { if (value <= 0) return SSLAPB_E_BAD_ARG; h->watchdog_ms = value; return 0; }
    if (!strcmp(name, "t_shard")) { if (value < 0 || value > 0x7fffffff) return SSLAPB_E_BAD_ARG; h->t_shard = (int)value; return 0; }
    if (!strcmp(name, "max_ctas")) { if (value < 0 || value > 65535) return SSLAPB_E_BAD_ARG; h->max_ctas = (int)value; return 0; }
    if (!strcmp(name, "strict")) { if (value < 0 || value > 1) return SSLAPB_E_BAD_ARG; h->strict = (int)value; return 0; }
    if (!strcmp(name, "hk_host_loop")) { if (value < 0 || value > 1) return SSLAPB_E_BAD_ARG; h->hk_host_loop = (int)value; return 0; }
    if (!strcmp(name, "batch_v1")) { if (value < 0 || value > 1) return SSLAPB_E_BAD_ARG; h->batch_v1 = (int)value; return 0; }
    if (!strcmp(name, "l2_persist")) { if (value < 0 || value > 1) return SSLAPB_E_BAD_ARG; h->l2_persist = (int)value; return 0; }
    if (!strcmp(name, "small_max_n")) { if (value < 1 || value > SSLAPB_SMALL_MAXN) return SSLAPB_E_BAD_ARG; h->small_max_n = (int)value; return 0; }
    if (!strcmp(name, "small_path")) { if (value < 0 || value > 1) return SSLAPB_E_BAD_ARG; h->small_path = (int)value; return 0; }
    if (!strcmp(name, "hot")) { if (value < 0 || value > 1) return SSLAPB_E_BAD_ARG; h->hot = (int)value; return 0; }
    if (!strcmp(name, "coop")) { if (value < 0 || value > 1) return SSLAPB_E_BAD_ARG; h->coop = (int)value; return 0; }
    return fail(h, SSLAPB_E_BAD_ARG, std::string("unknown option ") + name);
}

extern "C" void *sslapb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void sslapb_host_free(void *p) { if (p) cudaFreeHost(p); }

// ----------------------------------------------------------------------------------------------------------------------
// CSR build
// ----------------------------------------------------------------------------------------------------------------------
static int build_from_coo(sslapb_handle *h, const void *rows, const void *cols, int idx_bytes, int64_t stride,
                          const double *val, int64_t nnz, int32_t &n_rows, int32_t &n_cols, int negate, int mem,
                          SslapbBuildFlags &F)
{
    if (!rows || !cols || nnz < 0 || (idx_bytes != 4 && idx_bytes != 8)) return fail(h, SSLAPB_E_BAD_ARG, "bad COO arguments");
    const bool interleaved = stride == 2 && (const char *)cols == (const char *)rows + idx_bytes;
    if (!(interleaved || stride == 1)) return fail(h, SSLAPB_E_BAD_ARG, "COO must be an interleaved (K,2) array or two contiguous arrays");
    if (nnz >= (1ll << 40)) return fail(h, SSLAPB_E_BAD_ARG, "nnz too large");
    const void *d_rows = rows, *d_cols = cols;
    const double *d_val = val;
    CK(cudaEventRecord(h->ev[0], h->stream));
    if (!(mem & SSLAPB_MEM_DEVICE_IN) && nnz > 0) {
        const size_t ib = (size_t)idx_bytes;
        CK(h->stage_idx.reserve(2 * (size_t)nnz * ib));
        char *s = h->stage_idx.as<char>();
        if (interleaved) {
            CK(cudaMemcpyAsync(s, rows, 2 * (size_t)nnz * ib, cudaMemcpyHostToDevice, h->stream));
            d_rows = s; d_cols = s + ib;
        } else {
            CK(cudaMemcpyAsync(s, rows, (size_t)nnz * ib, cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemcpyAsync(s + (size_t)nnz * ib, cols, (size_t)nnz * ib, cudaMemcpyHostToDevice, h->stream));
            d_rows = s; d_cols = s + (size_t)nnz * ib;
        }
        if (val) {
            CK(h->stage_val.reserve((size_t)nnz * sizeof(double)));
            CK(cudaMemcpyAsync(h->stage_val.p, val, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            d_val = h->stage_val.as<double>();
        }
    }
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(h->flags.reserve(sizeof(SslapbBuildFlags) + 2 * sizeof(long long)));
    CK(cudaMemsetAsync(h->flags.p, 0, sizeof(SslapbBuildFlags) + 2 * sizeof(long long), h->stream));
    if ((n_rows <= 0 || n_cols <= 0) && nnz > 0) {          // AuctionSolver.__init__ infers max+1 (auction_.pyx:209-210)
        long long *mx = reinterpret_cast<long long *>(h->flags.as<char>() + sizeof(SslapbBuildFlags));
        CK(sslapb_launch_index_max(d_rows, d_cols, idx_bytes, stride, nnz, mx, h->sms, h->stream));
        long long hm[2];
        CK(cudaMemcpyAsync(hm, mx, sizeof hm, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (hm[0] < 0 || hm[1] < 0 || hm[0] >= 0x7fffffff || hm[1] >= 0x7fffffff)
            return fail(h, SSLAPB_E_OUT_OF_RANGE, "negative or oversized index in loc");
        if (n_rows <= 0) n_rows = (int32_t)hm[0] + 1;
        if (n_cols <= 0) n_cols = (int32_t)hm[1] + 1;
    }
    if (n_rows <= 0 || n_cols <= 0) return fail(h, SSLAPB_E_BAD_ARG, "empty problem");
    CK(h->cols.reserve(((size_t)nnz + 8) * sizeof(int)));
    CK(h->rowptr.reserve(((size_t)n_rows + 2) * sizeof(long long)));
    CK(cudaMemsetAsync(h->cols.as<int>() + nnz, 0, 8 * sizeof(int), h->stream));
    CK(cudaMemsetAsync(h->rowptr.p, 0, ((size_t)n_rows + 2) * sizeof(long long), h->stream));
    if (val) {
        CK(h->vals.reserve(((size_t)nnz + 8) * sizeof(double)));
        CK(cudaMemsetAsync(h->vals.as<double>() + nnz, 0, 8 * sizeof(double), h->stream));
    }
    CK(sslapb_launch_coo_ingest(d_rows, d_cols, idx_bytes, stride, d_val, nnz, n_rows, n_cols, negate, h->cols.as<int>(),
                                val ? h->vals.as<double>() : nullptr, h->rowptr.as<long long>(),
                                h->flags.as<SslapbBuildFlags>(), h->sms, h->stream));
    CK(cudaMemcpyAsync(&F, h->flags.p, sizeof F, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(h->ev[2], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (nnz == 0) F.empty_rows = 1;
    h->N = n_rows; h->M = n_cols; h->nnz = nnz; h->has_vals = val != nullptr; h->hot_valid = false;
    if (F.out_of_range) return fail(h, SSLAPB_E_OUT_OF_RANGE, "loc holds an index outside the matrix");
    if (F.unsorted) {
        // The reference silently requires row-sorted input (auction_.pyx:33-48).  Superset behaviour: stable device
        // radix sort by row (entries of one row keep their input order), then the same ingest pass on the sorted stream.
        if (nnz >= 0xffffffffll) return fail(h, SSLAPB_E_UNSORTED, "unsorted loc with >= 2^32 entries is not supported");
        const long long ntiles = (nnz + 4095) / 4096;
        CK(h->sort_keys.reserve(2 * (size_t)nnz * 4)); CK(h->sort_idx.reserve(2 * (size_t)nnz * 4));
        CK(h->sort_hist.reserve((size_t)(256 * ntiles + 1024) * 8));
        CK(h->sort_rows.reserve((size_t)nnz * 4)); CK(h->sort_cols.reserve((size_t)nnz * 4));
        if (val) CK(h->sort_val.reserve((size_t)nnz * 8));
        CK(sslapb_launch_coo_sort(d_rows, d_cols, idx_bytes, stride, d_val, nnz, n_rows, h->sort_keys.as<unsigned>(),
                                  h->sort_keys.as<unsigned>() + nnz, h->sort_idx.as<unsigned>(), h->sort_idx.as<unsigned>() + nnz,
                                  h->sort_hist.as<long long>(), h->sort_rows.as<int>(), h->sort_cols.as<int>(),
                                  val ? h->sort_val.as<double>() : nullptr, h->sms, h->stream));
        CK(cudaMemsetAsync(h->flags.p, 0, sizeof(SslapbBuildFlags), h->stream));
        CK(cudaMemsetAsync(h->rowptr.p, 0, ((size_t)n_rows + 2) * sizeof(long long), h->stream));
        CK(sslapb_launch_coo_ingest(h->sort_rows.p, h->sort_cols.p, 4, 1, val ? h->sort_val.as<double>() : nullptr, nnz, n_rows,
                                    n_cols, negate, h->cols.as<int>(), val ? h->vals.as<double>() : nullptr,
                                    h->rowptr.as<long long>(), h->flags.as<SslapbBuildFlags>(), h->sms, h->stream));
        CK(cudaMemcpyAsync(&F, h->flags.p, sizeof F, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->ev[2], h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (F.unsorted || F.out_of_range) return fail(h, SSLAPB_E_UNSORTED, "device sort failed (internal error)");
    }
    if (val) {
        CK(h->rowmax.reserve(((size_t)n_rows + 1) * sizeof(double)));
        CK(sslapb_launch_rowmax(h->rowptr.as<long long>(), h->vals.as<double>(), n_rows, h->rowmax.as<double>(), &h->flags.as<SslapbBuildFlags>()->maxdeg, h->sms, h->stream));
        CK(cudaMemcpyAsync(&h->maxdeg, &h->flags.as<SslapbBuildFlags>()->maxdeg, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->ev[2], h->stream));
        CK(cudaStreamSynchronize(h->stream));                 // the longest row picks the kernel instance (run_auction)
    }
    return SSLAPB_OK;
}

static int build_from_dense(sslapb_handle *h, const double *mat, int32_t n_rows, int32_t n_cols, int negate, int mem,
                            bool want_vals, SslapbBuildFlags &F)
{
    if (!mat || n_rows <= 0 || n_cols <= 0) return fail(h, SSLAPB_E_BAD_ARG, "bad dense arguments");
    const double *d_mat = mat;
    CK(cudaEventRecord(h->ev[0], h->stream));
    if (!(mem & SSLAPB_MEM_DEVICE_IN)) {
        const size_t bytes = (size_t)n_rows * (size_t)n_cols * sizeof(double);
        CK(h->stage_mat.reserve(bytes));
        CK(cudaMemcpyAsync(h->stage_mat.p, mat, bytes, cudaMemcpyHostToDevice, h->stream));
        d_mat = h->stage_mat.as<double>();
    }
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(h->flags.reserve(sizeof(SslapbBuildFlags) + 2 * sizeof(long long)));
    CK(cudaMemsetAsync(h->flags.p, 0, sizeof(SslapbBuildFlags), h->stream));
    CK(h->rowptr.reserve(((size_t)n_rows + 2) * sizeof(long long)));
    CK(sslapb_launch_dense_count(d_mat, n_rows, n_cols, h->rowptr.as<long long>(), h->flags.as<SslapbBuildFlags>(), h->sms,
                                 h->stream));
    CK(cudaMemcpyAsync(&F, h->flags.p, sizeof F, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const long long nnz = F.nnz;
    CK(h->cols.reserve(((size_t)nnz + 8) * sizeof(int)));
    CK(cudaMemsetAsync(h->cols.as<int>() + nnz, 0, 8 * sizeof(int), h->stream));
    if (want_vals) {
        CK(h->vals.reserve(((size_t)nnz + 8) * sizeof(double)));
        CK(cudaMemsetAsync(h->vals.as<double>() + nnz, 0, 8 * sizeof(double), h->stream));
    }
    CK(sslapb_launch_dense_fill(d_mat, n_rows, n_cols, negate, h->rowptr.as<long long>(), h->cols.as<int>(),
                                want_vals ? h->vals.as<double>() : nullptr, h->flags.as<SslapbBuildFlags>(), h->sms,
                                h->stream));
    CK(cudaMemcpyAsync(&F, h->flags.p, sizeof F, cudaMemcpyDeviceToHost, h->stream));
    if (want_vals) {
        CK(h->rowmax.reserve(((size_t)n_rows + 1) * sizeof(double)));
        CK(sslapb_launch_rowmax(h->rowptr.as<long long>(), h->vals.as<double>(), n_rows, h->rowmax.as<double>(), &h->flags.as<SslapbBuildFlags>()->maxdeg, h->sms, h->stream));
        CK(cudaMemcpyAsync(&h->maxdeg, &h->flags.as<SslapbBuildFlags>()->maxdeg, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaEventRecord(h->ev[2], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->N = n_rows; h->M = n_cols; h->nnz = nnz; h->has_vals = want_vals; h->hot_valid = false;
    return SSLAPB_OK;
}

// ----------------------------------------------------------------------------------------------------------------------
// Hopcroft-Karp phase loop on the resident CSR (feasibility_.pyx:199-211)
// ----------------------------------------------------------------------------------------------------------------------
static int run_hopcroft(sslapb_handle *h, int32_t *card_out)
{
    const int N = h->N, M = h->M;
    CK(h->pair_u.reserve((size_t)N * 4)); CK(h->pair_v.reserve((size_t)M * 4));
    CK(h->dist.reserve((size_t)N * 4)); CK(h->visited.reserve((size_t)M * 4));      // visited = pred_v
    CK(h->cursor.reserve((size_t)N * 8)); CK(h->pred.reserve((size_t)N * 4));         // cursor = root | end_of_root
    CK(h->hkflags.reserve(sizeof(SslapbHkFlags) > sizeof(SslapbHkCtrl) ? sizeof(SslapbHkFlags) : sizeof(SslapbHkCtrl)));
    const long long *rowptr = h->rowptr.as<long long>();
    const int *cols = h->cols.as<int>();
    SslapbHkFlags *dF = h->hkflags.as<SslapbHkFlags>();
    int *root = h->cursor.as<int>(), *end_of_root = h->cursor.as<int>() + N, *pred_v = h->visited.as<int>();
    CK(cudaMemsetAsync(h->pair_u.p, 0xff, (size_t)N * 4, h->stream));
    CK(cudaMemsetAsync(h->pair_v.p, 0xff, (size_t)M * 4, h->stream));
    if (!h->hk_host_loop) {
        // device-resident phase loop: ONE cooperative launch, one read-back (feasibility_.pyx:199-211)
        CK(h->hkq.reserve((size_t)N * 8));
        SslapbHkCtrl hc;
        memset(&hc, 0, sizeof hc);
        CK(cudaMemcpyAsync(dF, &hc, sizeof hc, cudaMemcpyHostToDevice, h->stream));
        CK(sslapb_hk_launch_persistent(rowptr, cols, N, M, h->pair_u.as<int>(), h->pair_v.as<int>(), root, end_of_root, pred_v,
                                       h->dist.as<int>(), h->pred.as<int>(), h->hkq.as<int>(), h->hkq.as<int>() + N, dF,
                                       h->hk_grid, h->stream));
        CK(cudaMemcpyAsync(&hc, dF, sizeof hc, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (hc.watchdog) return fail(h, SSLAPB_E_ABORTED, "Hopcroft-Karp: device watchdog fired");
        h->hk_phases = hc.phases; h->hk_levels = hc.levels;
        if (getenv("SSLAPB_HK_TRACE")) fprintf(stderr, "hk persistent: matched=%d phases=%d levels=%d\n", hc.matched, hc.phases, hc.levels);
        *card_out = (int32_t)hc.matched;
        return SSLAPB_OK;
    }
    CK(cudaMemsetAsync(dF, 0, sizeof(SslapbHkFlags), h->stream));
    CK(sslapb_hk_launch_greedy(rowptr, cols, N, h->pair_u.as<int>(), h->pair_v.as<int>(), dF, h->sms, h->stream));
    SslapbHkFlags F;
    CK(cudaMemcpyAsync(&F, dF, sizeof F, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    long long matching = F.matched;
    const long long bound = N < M ? N : M;
    const bool trace = getenv("SSLAPB_HK_TRACE") != nullptr;
    if (trace) fprintf(stderr, "hk greedy: matched=%lld of %lld\n", matching, bound);
    while (matching < bound) {
        CK(sslapb_hk_launch_phase_init(N, M, h->pair_u.as<int>(), h->dist.as<int>(), root, end_of_root, pred_v, dF, h->sms, h->stream));
        int level = 0;
        bool found = false;
        for (;;) {                                            // level-synchronous BFS (feasibility_.pyx:128-168)
            CK(cudaMemsetAsync(&dF->grew, 0, sizeof(int), h->stream));
            CK(sslapb_hk_launch_bfs_level(rowptr, cols, N, level, h->pair_v.as<int>(), h->dist.as<int>(), root, end_of_root,
                                          pred_v, dF, h->sms, h->stream));
            CK(cudaMemcpyAsync(&F, dF, sizeof F, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            if (F.found) { found = true; break; }
            if (!F.grew) break;
            ++level;
        }
        if (!found) break;                                    // no augmenting path: maximum (feasibility_.pyx:202-203)
        CK(sslapb_hk_launch_augment(N, end_of_root, pred_v, h->pair_u.as<int>(), h->pair_v.as<int>(), dF, h->sms, h->stream));
        CK(cudaMemcpyAsync(&F, dF, sizeof F, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (F.augmented <= 0) return fail(h, SSLAPB_E_ABORTED, "Hopcroft-Karp phase made no progress (internal error)");
        matching += F.augmented;
        if (trace) fprintf(stderr, "hk phase: levels=%d augmented=%d matching=%lld/%lld\n", level + 1, F.augmented, matching, bound);
    }
    *card_out = (int32_t)matching;
    return SSLAPB_OK;
}

// ----------------------------------------------------------------------------------------------------------------------
// Auction on the resident CSR
// ----------------------------------------------------------------------------------------------------------------------
static int reserve_auction_state(sslapb_handle *h, SslapbAuctionParams &P)
{
    const size_t N = (size_t)h->N, M = (size_t)h->M;
    CK(h->price.reserve(M * 8)); CK(h->owner.reserve(M * sizeof(SslapbObjRec))); CK(h->p2o.reserve(N * 4));
    CK(h->list.reserve(N * 4)); CK(h->mover.reserve(N * 4)); CK(h->bidj.reserve(N * 4)); CK(h->bidv.reserve(N * 8));
    CK(h->bidkey.reserve(M * 8)); CK(h->winpos.reserve(M * 4)); CK(h->hole_count.reserve((size_t)h->sms * 4 + 64));
    CK(h->chosen.reserve(N * 8)); CK(h->ctrl.reserve(sizeof(SslapbCtrl)));
    P.N = h->N; P.M = h->M;
    P.rowptr = h->rowptr.as<long long>(); P.cols = h->cols.as<int>(); P.vals = h->vals.as<double>();
    P.rowmax = h->rowmax.as<double>();
    P.hot = nullptr; P.hthr = nullptr; P.rest = nullptr;
    P.price = h->price.as<double>(); P.rec = h->owner.as<SslapbObjRec>(); P.p2o = h->p2o.as<int>();
    P.list = h->list.as<int>(); P.mover = h->mover.as<int>(); P.bidj = h->bidj.as<int>(); P.bidv = h->bidv.as<double>();
    P.bidkey = h->bidkey.as<unsigned long long>(); P.winpos = h->winpos.as<int>();
    P.hole_count = h->hole_count.as<int>(); P.chosen = h->chosen.as<double>(); P.ctrl = h->ctrl.as<SslapbCtrl>();
    P.t_small = h->t_small;
    P.t_mid = h->t_small == 32 ? h->t_mid : 0; P.pad_mid = 0;   // a caller that sets t_small is exercising the other regimes' hand-overs
    P.watchdog_ns = (unsigned long long)h->watchdog_ms * 1000000ull;
    P.nranks = 1; P.rank = 0; P.t_shard = 0x7fffffff; P.rowsplit = nullptr; P.xtab = nullptr; P.xcap = 0; P.xround_base = 0;
    return SSLAPB_OK;
}

// Hot lists of the resident CSR (hot.cu): rest[N] | hthr[N] | N x 32 entries, one allocation; built once per resident problem.
static int attach_hot_lists(sslapb_handle *h, SslapbAuctionParams &P, size_t *total_out, char **base_out)
{
    const size_t N = (size_t)h->N;
    const size_t head = ((N * 16) + 511) & ~(size_t)511, total = head + N * 512;
    CK(h->hotbuf.reserve(total));
    char *hb = h->hotbuf.as<char>();
    P.rest = reinterpret_cast<double *>(hb); P.hthr = reinterpret_cast<double *>(hb) + N;
    SslapbHotEnt *hot = reinterpret_cast<SslapbHotEnt *>(hb + head);
    P.hot = hot;
    if (!h->hot_valid) {
        CK(sslapb_launch_hot_build(P.rowptr, P.cols, P.vals, h->N, hot, const_cast<double *>(P.hthr), h->sms, h->stream));
        h->hot_valid = true;
    }
    if (total_out) *total_out = total;
    if (base_out) *base_out = hb;
    return SSLAPB_OK;
}

static int run_auction(sslapb_handle *h, const SslapbBuildFlags &F, int maximize, float eps_start, int64_t max_iter,
                       int mem, int32_t *sol_out, sslapb_meta *meta)
{
    SslapbAuctionParams P;
    int rc = reserve_auction_state(h, P);
    if (rc) return rc;
    const int N = h->N;
    // eps schedule constants, auction_.pyx:242-252 (float32 arithmetic as in the generated C)
    double cmax;
    memcpy(&cmax, &F.maxabs, sizeof cmax);
    const float C = (float)cmax;
    float eps = (float)((double)C / 2.0);
    // strict mode: eps-CS with eps = 1/(N+1) (N * eps < 1: provably optimal for integer costs) and no tolerance
    const float target = h->strict ? (float)(1.0 / ((double)N + 1.0)) : (float)(1.0 / (double)N);
    if (eps_start > 0) eps = eps_start;
    const bool long_rows = h->maxdeg > sslapb_coop_row_entries();
    // ---- warm start: the caller's prices replace the zeros of AuctionSolver.__init__ (:220)
    const bool warm = h->warm_cols > 0;
    if (warm) {
        if (h->warm_cols != h->M) { h->warm_cols = 0; return fail(h, SSLAPB_E_BAD_ARG, "sslapb_set_prices: n_cols differs from the problem's column count"); }
        CK(cudaMemcpyAsync(P.price, h->warm.p, (size_t)h->M * 8, cudaMemcpyDeviceToDevice, h->stream));
        h->warm_cols = 0;
    }
    // ---- row-sharded solve
    const bool sharded = h->n_ranks > 1 && !long_rows;
    if (h->n_ranks > 1) {
        if (!h->comm_connected) return fail(h, SSLAPB_E_BAD_ARG, "sslapb_comm_init without sslapb_comm_connect");
        if (h->comm_broken) return fail(h, SSLAPB_E_ABORTED, "the communicator is out of step after an aborted solve: destroy and re-create it");
        if ((long long)N > h->xcap) return fail(h, SSLAPB_E_BAD_ARG, "n_rows exceeds the communicator's capacity_rows");
    }
    if (sharded) {
        CK(h->rowsplit.reserve((SSLAPB_MAX_RANKS + 1) * sizeof(int)));
        CK(sslapb_launch_row_split(P.rowptr, N, h->n_ranks, h->rowsplit.as<int>(), h->stream));
        P.nranks = h->n_ranks; P.rank = h->rank; P.t_shard = h->t_shard; P.rowsplit = h->rowsplit.as<int>();
        P.xtab = h->xtab.as<unsigned long long>(); P.xcap = h->xcap; P.xround_base = h->xround;
        // every rank switches the hot form — and with it the mid regime — on and off from the counters of its OWN rows, so the
        // ranks may disagree about it; that is harmless only while the mid regime's rounds are rounds every rank runs in full
        // (nu <= t_shard: no exchange).  A rank in the mid regime would never reach the exchange barrier of a sharded round.
        if (P.t_mid > h->t_shard) P.t_mid = h->t_shard > 32 ? h->t_shard : 0;
    }
    SslapbCtrl c;
    memset(&c, 0, sizeof c);
    c.nu = N; c.eps = eps; c.target_eps = target; c.theta = 0.15f; c.max_iter = max_iter; c.ece_final = -1;
    c.tol = h->strict ? 0.0 : 1e-7;                            // auction_.pyx:16
    c.pmin_key[0] = 0x8000000000000000ull;                     // all prices start at +0.0 (auction_.pyx:220)
    c.pmin_key[1] = ~0ull;
    c.pmax_key = 0x8000000000000000ull;
    // Exactness guard: the bound-pruned sweeps and the hot lists rest on prices never decreasing (bid = a - w + eps > p), i.e.
    // on eps being large against the rounding of a price.  Where the smallest eps of the schedule comes within ~100 ulps of
    // the cost magnitude (|a| * N beyond 1e14, or such an eps_start) every sweep reads and gathers the full row instead.
    const bool tiny_eps = cmax * (double)N > 1e14 || (eps_start > 0 && (double)eps_start < cmax * 1e-13);
    if (tiny_eps) c.pmax_key = 0xfff0000000000000ull;          // spread = +inf: no pruning
    CK(cudaMemcpyAsync(P.ctrl, &c, sizeof c, cudaMemcpyHostToDevice, h->stream));
    if (warm) CK(sslapb_launch_price_bounds(&P, h->stream));   // pruning bounds of the first phase from the caller's prices
    int grid = h->grid;
    // small problems: a grid barrier among 148 CTAs costs more than the few rows the extra CTAs would sweep (measured, dense
    // N = 10 / 32 / 100: 1 / 2-4 / 4-16 CTAs are the fastest, tools/gpu_small.py); one CTA per 8 persons, at least one
    if (!sharded && (N + 7) / 8 < grid) grid = (N + 7) / 8;
    if (h->max_ctas > 0 && h->max_ctas < grid) grid = h->max_ctas;
    // ---- hot lists (hot.cu): the 32 largest entries of every row + the bound of the rest
    bool l2_window = false;
    if (h->hot && N > 32 && !tiny_eps) {
        size_t total = 0;
        char *hb = nullptr;
        rc = attach_hot_lists(h, P, &total, &hb);
        if (rc) return rc;
        if (h->l2_persist && h->l2_persist_max > 0 && h->l2_window_max > 0) {   // loads from the hot lists (and rest[]) stay in L2 across the phase
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof av);
            const size_t win = total < (size_t)h->l2_window_max ? total : (size_t)h->l2_window_max;
            // the set-aside is taken from the ordinary L2 for as long as the limit stands: only while this solve runs
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, win < h->l2_persist_max ? win : h->l2_persist_max);
            av.accessPolicyWindow.base_ptr = hb;
            av.accessPolicyWindow.num_bytes = win;
            av.accessPolicyWindow.hitRatio = win <= h->l2_persist_max ? 1.0f : (float)((double)h->l2_persist_max / (double)win);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            l2_window = cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
            cudaGetLastError();
        }
    }
    CK(cudaEventRecord(h->ev[3], h->stream));
    CK(sslapb_launch_auction_init(&P, grid, warm ? 1 : 0, h->stream));
    if (sharded && h->gate) {                                  // in-process ranks: nobody launches before everybody is ready
        CK(cudaStreamSynchronize(h->stream));
        if (!h->gate->arrive_and_wait(h->watchdog_ms)) {
            h->comm_broken = true;
            return fail(h, SSLAPB_E_ABORTED, "a rank of this process did not reach the launch of the row-sharded solve");
        }
    }
    CK(sslapb_launch_auction(&P, grid, long_rows, h->coop, h->stream));
    CK(cudaEventRecord(h->ev[4], h->stream));
    if (l2_window) {                                           // later launches on this stream: no window, persisting lines released
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    int rs[SSLAPB_MAX_RANKS + 1] = {};
    if (sharded) CK(cudaMemcpyAsync(rs, h->rowsplit.p, sizeof rs, cudaMemcpyDeviceToHost, h->stream));
    std::vector<double> chosen((size_t)N);
    std::vector<int32_t> sol_tmp;
    int32_t *sol_host = sol_out;
    if (mem & SSLAPB_MEM_DEVICE_OUT) {
        CK(cudaMemcpyAsync(sol_out, P.p2o, (size_t)N * 4, cudaMemcpyDeviceToDevice, h->stream));
        sol_tmp.resize((size_t)N);
        sol_host = sol_tmp.data();
    }
    CK(cudaMemcpyAsync(sol_host, P.p2o, (size_t)N * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(chosen.data(), P.chosen, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&c, P.ctrl, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (l2_window) { cudaCtxResetPersistingL2Cache(); cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0); cudaGetLastError(); }
    if (sharded) {
        h->xround += (unsigned)c.rounds_sharded;
        if (c.abort_flag) h->comm_broken = true;               // the ranks may have left the solve at different rounds
    }
    if (c.abort_flag) {
        std::string msg = c.abort_flag == 2 ? "empty row reached the bidding kernel" : "device watchdog fired";
        if (sharded)
            msg += " [rank " + std::to_string(h->rank) + "/" + std::to_string(h->n_ranks) + " nu=" + std::to_string(c.nu) + " its=" +
                   std::to_string(c.its) + " sharded rounds=" + std::to_string(c.rounds_sharded) + " grid=" + std::to_string(grid) + "]";
        return fail(h, SSLAPB_E_ABORTED, msg);
    }
    double obj = 0.0;                                          // get_obj, auction_.pyx:489-523 (row order, double)
    long long assigned = 0;
    for (int i = 0; i < N; ++i) {
        if (sol_host[i] == -1) continue;
        ++assigned;
        if (maximize) obj += chosen[(size_t)i]; else obj -= chosen[(size_t)i];
    }
    if (meta) {
        meta->start_eps = eps; meta->final_eps = c.eps; meta->target_eps = target;
        meta->eCE = c.ece_final > 0; meta->soln_found = (c.nu == 0) && (c.ece_final > 0);
        meta->its = c.its; meta->nreductions = c.nreductions; meta->n_assigned = (int64_t)N - c.nu;
        meta->obj64 = obj; meta->obj = (float)obj;
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); meta->setup_ms = ms;
        cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]); meta->solve_ms = ms;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); meta->h2d_ms = ms;
        meta->n_rows = h->N; meta->n_cols = h->M; meta->nnz = h->nnz;
        meta->rounds_grid = c.rounds_grid; meta->rounds_warp = c.rounds_warp; meta->rounds_solo = c.rounds_solo;
        meta->small_path = 0;
        for (int k = 0; k < 8; ++k) meta->prof_ms[k] = (float)((double)c.prof[k] * 1e-6);
        meta->stop_reason = c.done;
        meta->prune_second_pass = c.prune_second_pass;
        meta->n_ranks = sharded ? h->n_ranks : 1; meta->rank = sharded ? h->rank : 0;
        meta->row_lo = sharded ? rs[h->rank] : 0; meta->row_hi = sharded ? rs[h->rank + 1] : N;
        meta->rounds_sharded = c.rounds_sharded;
        meta->xchg_ms = (float)((double)c.xchg_ns * 1e-6); meta->sharded_ms = (float)((double)c.sharded_ns * 1e-6);
        meta->sweep_insitu_n = (int32_t)c.sweep_ns[1];
        meta->sweep_insitu_us = c.sweep_ns[1] ? (float)((double)c.sweep_ns[0] * 1e-3 / (double)c.sweep_ns[1]) : 0.f;
        meta->warm_start = warm ? 1 : 0; meta->strict = h->strict;
        meta->hot_grid_bids = c.hot_grid[0]; meta->hot_grid_fallbacks = c.hot_grid[1];
        meta->hot_tail_rounds = c.hot_tail[0]; meta->hot_tail_fallbacks = c.hot_tail[1];
        meta->rounds_nohole = c.rounds_nohole;
        meta->rounds_mid = c.rounds_mid;
        (void)assigned;
    }
    return SSLAPB_OK;
}

static int solve_resident(sslapb_handle *h, const SslapbBuildFlags &F, int maximize, float eps_start, int64_t max_iter,
                          int cardinality_check, int mem, int32_t *sol_out, sslapb_meta *meta)
{
    if (meta) { memset(meta, 0, sizeof *meta); meta->cardinality = -1; meta->n_rows = h->N; meta->n_cols = h->M; meta->nnz = h->nnz; }
    if (h->nnz < h->N)                                         // auction_.pyx:559-560 / :604-605
        return fail(h, SSLAPB_E_FEWER_THAN_N, "fewer valid values than rows");
    if (cardinality_check) {                                   // :562-566 / :608-612
        auto t0 = std::chrono::steady_clock::now();
        int32_t card = 0;
        int rc = run_hopcroft(h, &card);
        if (rc) return rc;
        if (meta) {
            meta->cardinality = card;
            meta->hk_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        if (card < h->N) return fail(h, SSLAPB_E_CARDINALITY, "maximum matching smaller than the number of rows");
    } else if (F.empty_rows) {
        return fail(h, SSLAPB_E_EMPTY_ROW, "a row has no valid entry");
    }
    if (!sol_out) return fail(h, SSLAPB_E_BAD_ARG, "sol_out is NULL");
    const float hk_ms = meta ? meta->hk_ms : 0.f;
    const int32_t card = meta ? meta->cardinality : -1;
    int rc = run_auction(h, F, maximize, eps_start, max_iter, mem, sol_out, meta);
    if (meta) { meta->hk_ms = hk_ms; meta->cardinality = card; }
    return rc;
}

// ----------------------------------------------------------------------------------------------------------------------
// Small problems: one H2D copy, ONE launch (small.cu: build + feasibility + auction in shared memory), one D2H copy.
// *handled = false: not eligible (or the kernel sent the input to the general path: unsorted / out of range / too many
// valid entries) — the caller continues as before.  The resident CSR of the handle is invalidated (sweeps refuse).
// ----------------------------------------------------------------------------------------------------------------------
static int run_small(sslapb_handle *h, const double *mat, const void *rows, const void *cols, int idx_bytes, int64_t stride,
                     const double *val, int64_t nnz, int32_t N, int32_t M, int maximize, float eps_start, int64_t max_iter,
                     int cardinality_check, int mem, int32_t *sol_out, sslapb_meta *meta, bool *handled)
{
    *handled = false;
    // only with the default regime options: a caller that sets t_small is exercising the general path's regimes
    if (!h->small_path || h->t_small != 32 || h->n_ranks > 1 || h->strict || h->warm_cols > 0 || mem != SSLAPB_MEM_HOST || !sol_out)
        return SSLAPB_OK;
    if (N < 1 || M < 1 || N > h->small_max_n || M > h->small_max_n) return SSLAPB_OK;
    const bool dense = mat != nullptr;
    const bool interleaved = !dense && stride == 2 && (const char *)cols == (const char *)rows + 4;
    if (!dense && (idx_bytes != 4 || nnz < 1 || nnz > SSLAPB_SMALL_CAP || !(interleaved || stride == 1))) return SSLAPB_OK;
    SslapbSmallArgs A;
    memset(&A, 0, sizeof A);
    CK(cudaEventRecord(h->ev[0], h->stream));
    if (dense) {
        CK(h->stage_mat.reserve((size_t)N * M * 8));
        CK(cudaMemcpyAsync(h->stage_mat.p, mat, (size_t)N * M * 8, cudaMemcpyHostToDevice, h->stream));
        A.mat = h->stage_mat.as<double>(); A.dense = 1;
    } else {
        CK(h->stage_idx.reserve(2 * (size_t)nnz * 4));
        CK(h->stage_val.reserve((size_t)nnz * 8));
        int *s = h->stage_idx.as<int>();
        if (interleaved) {
            CK(cudaMemcpyAsync(s, rows, 2 * (size_t)nnz * 4, cudaMemcpyHostToDevice, h->stream));
            A.rows = s; A.cols_in = s + 1; A.stride = 2;
        } else {
            CK(cudaMemcpyAsync(s, rows, (size_t)nnz * 4, cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemcpyAsync(s + nnz, cols, (size_t)nnz * 4, cudaMemcpyHostToDevice, h->stream));
            A.rows = s; A.cols_in = s + nnz; A.stride = 1;
        }
        CK(cudaMemcpyAsync(h->stage_val.p, val, (size_t)nnz * 8, cudaMemcpyHostToDevice, h->stream));
        A.val = h->stage_val.as<double>(); A.nnz = (int)nnz;
    }
    // result block and `sol` side by side in one device buffer: ONE read-back
    struct SmallOut { SslapbSmallResult res; int sol[SSLAPB_SMALL_MAXN]; };
    CK(h->price.reserve((size_t)M * 8));
    CK(h->ctrl.reserve(sizeof(SslapbCtrl) > sizeof(SmallOut) ? sizeof(SslapbCtrl) : sizeof(SmallOut)));
    SmallOut *d_out = h->ctrl.as<SmallOut>();
    A.N = N; A.M = M; A.negate = !maximize; A.hk = cardinality_check ? 1 : 0; A.eps_start = eps_start; A.max_iter = max_iter;
    A.sol_out = d_out->sol; A.price_out = h->price.as<double>(); A.res = &d_out->res;
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(sslapb_launch_small(&A, h->stream));
    CK(cudaEventRecord(h->ev[2], h->stream));
    SmallOut out;
    CK(cudaMemcpyAsync(&out, d_out, offsetof(SmallOut, sol) + (size_t)N * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const SslapbSmallResult &R = out.res;
    if (R.status == 3) return SSLAPB_OK;                       // the general path takes it
    *handled = true;
    if (R.status == 0) memcpy(sol_out, out.sol, (size_t)N * 4);
    h->N = N; h->M = M; h->nnz = R.nnz; h->has_vals = false; h->hot_valid = false;   // no resident CSR: prices only
    if (meta) {
        memset(meta, 0, sizeof *meta);
        meta->cardinality = R.cardinality; meta->n_rows = N; meta->n_cols = M; meta->nnz = R.nnz;
        meta->n_ranks = 1; meta->row_hi = N; meta->small_path = 1;
    }
    if (R.status == 1) return fail(h, SSLAPB_E_FEWER_THAN_N, "fewer valid values than rows");
    if (R.status == 2) return fail(h, SSLAPB_E_CARDINALITY, "maximum matching smaller than the number of rows");
    if (R.status == 6) return fail(h, SSLAPB_E_EMPTY_ROW, "a row has no valid entry");
    if (meta) {
        meta->start_eps = R.start_eps; meta->final_eps = R.final_eps; meta->target_eps = R.target_eps;
        meta->eCE = R.eCE; meta->soln_found = R.soln_found; meta->its = R.its; meta->nreductions = R.nreductions;
        meta->n_assigned = R.n_assigned; meta->obj64 = R.obj64; meta->obj = (float)R.obj64;
        meta->stop_reason = R.its >= max_iter ? 3 : (R.eCE ? 1 : 2);
        meta->rounds_solo = R.its;                             // (all rounds run by the one CTA)
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); meta->solve_ms = ms;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); meta->h2d_ms = ms;
    }
    return SSLAPB_OK;
}

extern "C" int sslapb_auction_coo(sslapb_handle *h, const void *rows, const void *cols, int idx_bytes, int64_t stride,
                                  const double *val, int64_t nnz, int32_t n_rows, int32_t n_cols, int maximize,
                                  float eps_start, int64_t max_iter, int cardinality_check, int mem, int32_t *sol_out,
                                  sslapb_meta *meta)
{
    if (!h) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    if (!val) return fail(h, SSLAPB_E_BAD_ARG, "val is NULL");
    CK(cudaSetDevice(h->device));
    {
        bool handled = false;
        int rs = run_small(h, nullptr, rows, cols, idx_bytes, stride, val, nnz, n_rows, n_cols, maximize, eps_start, max_iter,
                           cardinality_check, mem, sol_out, meta, &handled);
        if (rs || handled) return rs;
    }
    SslapbBuildFlags F;
    memset(&F, 0, sizeof F);
    if (nnz == 0) return fail(h, SSLAPB_E_FEWER_THAN_N, "no entries");
    int rc = build_from_coo(h, rows, cols, idx_bytes, stride, val, nnz, n_rows, n_cols, !maximize, mem, F);
    if (rc) return rc;
    return solve_resident(h, F, maximize, eps_start, max_iter, cardinality_check, mem, sol_out, meta);
}

extern "C" int sslapb_auction_dense(sslapb_handle *h, const double *mat, int32_t n_rows, int32_t n_cols, int maximize,
                                    float eps_start, int64_t max_iter, int cardinality_check, int mem, int32_t *sol_out,
                                    sslapb_meta *meta)
{
    if (!h) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    CK(cudaSetDevice(h->device));
    bool small_fits = mat && (long long)n_rows * n_cols <= 4ll * SSLAPB_SMALL_CAP && mem == SSLAPB_MEM_HOST;
    if (small_fits && (long long)n_rows * n_cols > SSLAPB_SMALL_CAP) {        // count the valid entries before trying (a few microseconds)
        long long cnt = 0;
        for (long long k = 0, e = (long long)n_rows * n_cols; k < e; ++k) cnt += mat[k] >= 0.0;
        small_fits = cnt <= SSLAPB_SMALL_CAP;
    }
    if (small_fits) {
        bool handled = false;
        int rs = run_small(h, mat, nullptr, nullptr, 0, 0, nullptr, 0, n_rows, n_cols, maximize, eps_start, max_iter,
                           cardinality_check, mem, sol_out, meta, &handled);
        if (rs || handled) return rs;
    }
    SslapbBuildFlags F;
    memset(&F, 0, sizeof F);
    int rc = build_from_dense(h, mat, n_rows, n_cols, !maximize, mem, true, F);
    if (rc) return rc;
    return solve_resident(h, F, maximize, eps_start, max_iter, cardinality_check, mem, sol_out, meta);
}

static int hopcroft_finish(sslapb_handle *h, int mem, int32_t *left_out, int32_t *right_out, int32_t *size_out)
{
    int32_t card = 0;
    int rc = run_hopcroft(h, &card);
    if (rc) return rc;
    const cudaMemcpyKind kind = (mem & SSLAPB_MEM_DEVICE_OUT) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (left_out) CK(cudaMemcpyAsync(left_out, h->pair_u.p, (size_t)h->N * 4, kind, h->stream));
    if (right_out) CK(cudaMemcpyAsync(right_out, h->pair_v.p, (size_t)h->M * 4, kind, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (size_out) *size_out = card;
    return SSLAPB_OK;
}

extern "C" int sslapb_hopcroft_coo(sslapb_handle *h, const void *rows, const void *cols, int idx_bytes, int64_t stride,
                                   int64_t nnz, int32_t n_rows, int32_t n_cols, int mem, int32_t *left_out,
                                   int32_t *right_out, int32_t *size_out)
{
    if (!h) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    CK(cudaSetDevice(h->device));
    SslapbBuildFlags F;
    memset(&F, 0, sizeof F);
    int rc = build_from_coo(h, rows, cols, idx_bytes, stride, nullptr, nnz, n_rows, n_cols, 0, mem, F);
    if (rc) return rc;
    return hopcroft_finish(h, mem, left_out, right_out, size_out);
}

extern "C" int sslapb_hopcroft_dense(sslapb_handle *h, const double *mat, int32_t n_rows, int32_t n_cols, int mem,
                                     int32_t *left_out, int32_t *right_out, int32_t *size_out)
{
    if (!h) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    CK(cudaSetDevice(h->device));
    SslapbBuildFlags F;
    memset(&F, 0, sizeof F);
    int rc = build_from_dense(h, mat, n_rows, n_cols, 0, mem, false, F);
    if (rc) return rc;
    return hopcroft_finish(h, mem, left_out, right_out, size_out);
}

extern "C" int sslapb_get_prices(sslapb_handle *h, double *prices_out)
{
    if (!h || !prices_out) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    if (!h->price.p || h->M <= 0) return fail(h, SSLAPB_E_BAD_ARG, "no solve has run on this handle");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(prices_out, h->price.p, (size_t)h->M * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SSLAPB_OK;
}

extern "C" int sslapb_set_prices(sslapb_handle *h, const double *prices, int32_t n_cols)
{
    if (!h) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    if (!prices) { h->warm_cols = 0; return SSLAPB_OK; }
    if (n_cols <= 0) return fail(h, SSLAPB_E_BAD_ARG, "sslapb_set_prices: n_cols <= 0");
    CK(cudaSetDevice(h->device));
    CK(h->warm.reserve((size_t)n_cols * 8));
    CK(cudaMemcpyAsync(h->warm.p, prices, (size_t)n_cols * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->warm_cols = n_cols;
    return SSLAPB_OK;
}

// ----------------------------------------------------------------------------------------------------------------------
// Row-sharded communicator: one exchange buffer per rank, peer-mapped into every other rank (see the header)
// ----------------------------------------------------------------------------------------------------------------------
namespace {
struct CommExport {                 // SSLAPB_COMM_EXPORT_BYTES
    unsigned long long magic;
    long long pid;
    int device, rank, n_ranks, pad;
    unsigned long long base;        // device address of the exchange buffer (valid inside the exporting process)
    long long xcap;
    cudaIpcMemHandle_t ipc;         // 64 bytes
    char fill[SSLAPB_COMM_EXPORT_BYTES - 48 - 64];
};
static_assert(sizeof(CommExport) == SSLAPB_COMM_EXPORT_BYTES, "export blob size");
const unsigned long long COMM_MAGIC = 0x53534c4150423230ull;
size_t xbuf_bytes(long long cap) { return 128 + (size_t)cap * 2 * 8 + (size_t)cap * 2 * 4; }
}

extern "C" int sslapb_comm_destroy(sslapb_handle *h)
{
    if (!h) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int r = 0; r < 8; ++r) {
        if (h->peer_ipc[r] && h->peer_base[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
        h->peer_ipc[r] = false; h->peer_base[r] = nullptr;
    }
    h->xbuf.release(); h->xtab.release();
    if (h->gate) {
        std::lock_guard<std::mutex> g(g_gates_mu);
        h->gate.reset();
        auto it = g_gates.find(h->gate_key);
        if (it != g_gates.end() && it->second.use_count() == 1) g_gates.erase(it);
    }
    h->n_ranks = 1; h->rank = 0; h->comm_connected = false; h->comm_broken = false; h->xcap = 0; h->xround = 0;
    return SSLAPB_OK;
}

extern "C" int sslapb_comm_init(sslapb_handle *h, int n_ranks, int rank, int64_t capacity_rows, void *export_out)
{
    if (!h) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    if (n_ranks < 1 || n_ranks > SSLAPB_COMM_MAX_RANKS || rank < 0 || rank >= n_ranks || capacity_rows <= 0 || !export_out)
        return fail(h, SSLAPB_E_BAD_ARG, "sslapb_comm_init: bad arguments");
    sslapb_comm_destroy(h);
    CK(cudaSetDevice(h->device));
    CK(h->xbuf.reserve(xbuf_bytes(capacity_rows)));
    CK(cudaMemsetAsync(h->xbuf.p, 0, xbuf_bytes(capacity_rows), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->n_ranks = n_ranks; h->rank = rank; h->xcap = capacity_rows;
    CommExport e;
    memset(&e, 0, sizeof e);
    e.magic = COMM_MAGIC; e.pid = (long long)getpid(); e.device = h->device; e.rank = rank; e.n_ranks = n_ranks;
    e.base = (unsigned long long)h->xbuf.p; e.xcap = capacity_rows;
    if (n_ranks > 1) CK(cudaIpcGetMemHandle(&e.ipc, h->xbuf.p));
    memcpy(export_out, &e, sizeof e);
    return SSLAPB_OK;
}

extern "C" int sslapb_comm_connect(sslapb_handle *h, const void *all_exports)
{
    if (!h || !all_exports) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    if (h->xcap <= 0) return fail(h, SSLAPB_E_BAD_ARG, "sslapb_comm_connect before sslapb_comm_init");
    CK(cudaSetDevice(h->device));
    const CommExport *E = reinterpret_cast<const CommExport *>(all_exports);
    unsigned long long tab[3 * SSLAPB_MAX_RANKS] = {};
    for (int r = 0; r < h->n_ranks; ++r) {
        const CommExport &e = E[r];
        if (e.magic != COMM_MAGIC || e.rank != r || e.n_ranks != h->n_ranks || e.xcap != h->xcap)
            return fail(h, SSLAPB_E_BAD_ARG, "sslapb_comm_connect: export blob " + std::to_string(r) + " does not belong to this communicator");
        void *base = nullptr;
        if (r == h->rank) base = h->xbuf.p;
        else if (e.pid == (long long)getpid()) {               // same process: plain peer access
            base = (void *)e.base;
            if (e.device != h->device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, h->device, e.device));
                if (!can) return fail(h, SSLAPB_E_BAD_ARG, "no peer access between the devices of ranks " + std::to_string(h->rank) + " and " + std::to_string(r));
                cudaError_t pe = cudaDeviceEnablePeerAccess(e.device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CK(pe);
                cudaGetLastError();
            }
        } else {                                               // another process: CUDA IPC mapping (enables peer access lazily)
            CK(cudaIpcOpenMemHandle(&base, e.ipc, cudaIpcMemLazyEnablePeerAccess));
            h->peer_ipc[r] = true;
        }
        h->peer_base[r] = base;
        char *b = (char *)base;
        tab[3 * r] = (unsigned long long)b;                                            // flags
        tab[3 * r + 2] = (unsigned long long)(b + 128);                                // bidv[2][xcap]
        tab[3 * r + 1] = (unsigned long long)(b + 128 + (size_t)h->xcap * 2 * 8);      // bidj[2][xcap]
    }
    bool all_local = h->n_ranks > 1;
    for (int r = 0; r < h->n_ranks; ++r) all_local = all_local && E[r].pid == (long long)getpid();
    if (all_local) {
        std::lock_guard<std::mutex> g(g_gates_mu);
        h->gate_key = E[0].base;
        auto &slot = g_gates[h->gate_key];
        // a gate left behind by an earlier communicator whose rank 0 buffer had the same address is never reused:
        // only the map itself holding it (use_count 1), or a different size, means it is stale
        if (!slot || slot.use_count() == 1 || slot->n != h->n_ranks) { slot = std::make_shared<LocalGate>(); slot->n = h->n_ranks; }
        h->gate = slot;
    }
    CK(h->xtab.reserve(sizeof tab));
    CK(cudaMemcpyAsync(h->xtab.p, tab, sizeof tab, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->comm_connected = true; h->comm_broken = false; h->xround = 0;
    return SSLAPB_OK;
}

// L2 flush of the roofline measurement, second half: read a 256 MB buffer.  The memset before it evicts everything that was
// resident but leaves ~126 MB of DIRTY lines behind, whose write-back would be charged to the measured kernel's first
// misses; after this read pass the L2 holds clean, unrelated lines.
__global__ void __launch_bounds__(1024) sslapb_l2_drain_kernel(const uint4 *__restrict__ p, size_t n, unsigned *sink)
{
    unsigned acc = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = p[k];
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x9e3779b9u) *sink = acc;                       // never true for a zero-filled buffer: keeps the loads alive
}

extern "C" int sslapb_bid_sweep(sslapb_handle *h, const double *prices, const int32_t *bidders, int32_t nb, float eps,
                                int merge, int iters, int flush_l2, int32_t *jbest_out, double *bid_out,
                                float *avg_ms_out)
{
    if (!h) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    if (h->N <= 0 || !h->has_vals || nb <= 0 || iters < 1) return fail(h, SSLAPB_E_BAD_ARG, "no resident problem / bad arguments");
    if (!bidders && nb > h->N) return fail(h, SSLAPB_E_BAD_ARG, "nb > N");
    CK(cudaSetDevice(h->device));
    SslapbAuctionParams P;
    const bool fresh = h->price.cap < (size_t)h->M * 8;
    int rc = reserve_auction_state(h, P);
    if (rc) return rc;
    if (prices) CK(cudaMemcpyAsync(P.price, prices, (size_t)h->M * 8, cudaMemcpyHostToDevice, h->stream));
    else if (fresh) CK(cudaMemsetAsync(P.price, 0, (size_t)h->M * 8, h->stream));
    const int *d_bidders = nullptr;
    if (bidders) {
        CK(h->bidders.reserve((size_t)nb * 4));
        CK(cudaMemcpyAsync(h->bidders.p, bidders, (size_t)nb * 4, cudaMemcpyHostToDevice, h->stream));
        d_bidders = h->bidders.as<int>();
        if ((size_t)nb > (size_t)h->N) { CK(h->bidj.reserve((size_t)nb * 4)); CK(h->bidv.reserve((size_t)nb * 8)); P.bidj = h->bidj.as<int>(); P.bidv = h->bidv.as<double>(); }
    }
    CK(sslapb_launch_price_bounds(&P, h->stream));
    // bit 2 of `merge`: identity frontier through the streamed (TMA ring) sweep, sweep_tma.cu — bit-identical results,
    // measured slower than the per-row kernel on B200, kept for A/B runs
    const bool streamed = !d_bidders && nb == h->N && (merge & 4);
    // variants of the sweep, all bit-identical (A/B runs): default = the per-row kernel (one warp per row, 32
    // warps per SM — the fastest measured); bit 3 = the software-pipelined per-row kernel (sweep2.cu; bits 4-5: its CTA size
    // 768 / 1024 / 640 / 512, bit 6: untrimmed stage C); bit 7 = four rows per warp, 8 lanes per row (sweep4.cu)
    const bool pipelined = (merge & 8) != 0;
    const bool per_row = !pipelined && (merge & 128) == 0;
    // bit 8: hot form — every bidder from its hot list when provably exact (bounds taken at the prices of this call), else
    // the full-row sweep; what the persistent kernel's grid regime runs from the third eps-phase on
    const bool hot_form = (merge & 256) != 0 && h->N > 32 && !streamed;
    // bit 9: the lean full-row sweep (sweep_lean.cu: 40 registers, redo list for the uncommon rows) — every row read in full
    const bool lean_form = (merge & 512) != 0 && !streamed && !hot_form;
    if (hot_form) {
        rc = attach_hot_lists(h, P, nullptr, nullptr);
        if (rc) return rc;
        CK(sslapb_launch_hot_rest(&P, h->sms, h->stream));
    }
    static const int sw2_threads[4] = {768, 1024, 640, 512};
    const int threads2 = sw2_threads[(merge >> 4) & 3];
    const int lean2 = (merge & 64) ? 0 : 1;                    // bit 6: the untrimmed stage C (A/B)
    merge &= 3;
    if (streamed) {
        CK(h->sweep_plan.reserve(((size_t)h->sms + 2) * 4));
        CK(sslapb_launch_sweep_plan(&P, h->sms, h->sweep_plan.as<int>(), h->stream));
    }
    const size_t flush_bytes = (size_t)256 << 20;
    if (flush_l2) {
        CK(h->flush.reserve(2 * flush_bytes));
        CK(cudaMemsetAsync(h->flush.as<char>() + flush_bytes, 0, flush_bytes, h->stream));
    }
    float total = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (merge) CK(cudaMemsetAsync(P.bidkey, 0, (size_t)h->M * 8, h->stream));
        if (hot_form || lean_form) CK(cudaMemsetAsync(&P.ctrl->hot_probe_fail, 0, sizeof(int), h->stream));   // redo counter of the launch pairs
        if (flush_l2) {
            CK(cudaMemsetAsync(h->flush.p, it & 0xff, flush_bytes, h->stream));
            if (flush_l2 > 1)                                  // 2: memset, then a read pass (clean lines, see sslapb_l2_drain_kernel)
                sslapb_l2_drain_kernel<<<h->sms * 2, 1024, 0, h->stream>>>(reinterpret_cast<const uint4 *>(h->flush.as<char>() + flush_bytes),
                                                                           flush_bytes / 16, reinterpret_cast<unsigned *>(h->flush.p));
        }
        CK(cudaEventRecord(h->ev[3], h->stream));
        if (streamed) CK(sslapb_launch_bid_sweep_tma(&P, h->sweep_plan.as<int>(), eps, merge, h->sms, h->stream));
        else if (hot_form) CK(sslapb_launch_bid_sweep_hot(&P, d_bidders, nb, eps, merge, h->sms, h->stream));
        else if (lean_form) CK(sslapb_launch_bid_sweep_lean(&P, d_bidders, nb, eps, merge, h->sms, h->stream));
        else if (per_row) CK(sslapb_launch_bid_sweep(&P, d_bidders, nb, eps, merge, h->sms, h->stream));
        else if (pipelined) CK(sslapb_launch_bid_sweep2(&P, d_bidders, nb, eps, merge, threads2, lean2, h->sms, h->stream));
        else CK(sslapb_launch_bid_sweep4(&P, d_bidders, nb, eps, merge, h->sms, h->stream));
        CK(cudaEventRecord(h->ev[4], h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]));
        total += ms;
    }
    if (merge) CK(cudaMemsetAsync(P.bidkey, 0, (size_t)h->M * 8, h->stream));
    if (jbest_out) CK(cudaMemcpyAsync(jbest_out, P.bidj, (size_t)nb * 4, cudaMemcpyDeviceToHost, h->stream));
    if (bid_out) CK(cudaMemcpyAsync(bid_out, P.bidv, (size_t)nb * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (avg_ms_out) *avg_ms_out = total / (float)iters;
    return SSLAPB_OK;
}


// ----------------------------------------------------------------------------------------------------------------------
// Batch of independent problems (BASELINE.json configs[4]) — one warp per problem, see auction.cu
// ----------------------------------------------------------------------------------------------------------------------
extern "C" int sslapb_auction_batch(sslapb_handle *h, int32_t n_problems, const int64_t *nnz_offsets, const int32_t *n_rows,
                                    const int32_t *n_cols, const void *rows, const void *cols, int idx_bytes, int64_t stride,
                                    const double *val, int maximize, const float *eps_start, int64_t max_iter, int mem,
                                    int32_t *sol_out, sslapb_meta *metas)
{
    if (!h) return SSLAPB_E_BAD_ARG;
    LOCK(h);
    if (n_problems <= 0 || !nnz_offsets || !n_rows || !n_cols || !rows || !cols || !val || !sol_out)
        return fail(h, SSLAPB_E_BAD_ARG, "bad batch arguments");
    CK(cudaSetDevice(h->device));
    const int P = n_problems;
    std::vector<long long> off(3 * ((size_t)P + 1));
    long long *nnzoff = off.data(), *rowoff = nnzoff + P + 1, *coloff = rowoff + P + 1;
    rowoff[0] = coloff[0] = 0;
    for (int p = 0; p <= P; ++p) nnzoff[p] = nnz_offsets[p];
    for (int p = 0; p < P; ++p) {
        if (n_rows[p] <= 0 || n_cols[p] <= 0 || nnzoff[p + 1] < nnzoff[p]) return fail(h, SSLAPB_E_BAD_ARG, "bad problem shape");
        if (nnzoff[p + 1] - nnzoff[p] < n_rows[p]) return fail(h, SSLAPB_E_FEWER_THAN_N, "a problem has fewer values than rows");
        rowoff[p + 1] = rowoff[p] + n_rows[p];
        coloff[p + 1] = coloff[p] + n_cols[p];
    }
    const long long nnz = nnzoff[P] - nnzoff[0], R = rowoff[P], Cn = coloff[P];
    if (nnzoff[0] != 0 || R >= 0x7fffffffll || Cn >= 0x7fffffffll) return fail(h, SSLAPB_E_BAD_ARG, "batch too large or offsets not starting at 0");
    const bool interleaved = stride == 2 && (const char *)cols == (const char *)rows + idx_bytes;
    if (!(interleaved || stride == 1) || (idx_bytes != 4 && idx_bytes != 8)) return fail(h, SSLAPB_E_BAD_ARG, "bad COO layout");
    // stage the local COO (if on the host) and the offset tables
    const void *d_rows = rows, *d_cols = cols;
    const double *d_val = val;
    CK(cudaEventRecord(h->ev[0], h->stream));
    if (!(mem & SSLAPB_MEM_DEVICE_IN)) {
        const size_t ib = (size_t)idx_bytes;
        CK(h->stage_idx.reserve(2 * (size_t)nnz * ib));
        char *s = h->stage_idx.as<char>();
        if (interleaved) { CK(cudaMemcpyAsync(s, rows, 2 * (size_t)nnz * ib, cudaMemcpyHostToDevice, h->stream)); d_rows = s; d_cols = s + ib; }
        else {
            CK(cudaMemcpyAsync(s, rows, (size_t)nnz * ib, cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemcpyAsync(s + (size_t)nnz * ib, cols, (size_t)nnz * ib, cudaMemcpyHostToDevice, h->stream));
            d_rows = s; d_cols = s + (size_t)nnz * ib;
        }
        CK(h->stage_val.reserve((size_t)nnz * 8));
        CK(cudaMemcpyAsync(h->stage_val.p, val, (size_t)nnz * 8, cudaMemcpyHostToDevice, h->stream));
        d_val = h->stage_val.as<double>();
    }
    CK(h->b_off.reserve(off.size() * 8));
    CK(cudaMemcpyAsync(h->b_off.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, h->stream));
    const long long *d_nnzoff = h->b_off.as<long long>(), *d_rowoff = d_nnzoff + P + 1, *d_coloff = d_rowoff + P + 1;
    CK(h->b_rows.reserve((size_t)nnz * 4)); CK(h->b_cols.reserve((size_t)nnz * 4)); CK(h->b_bad.reserve(16));
    CK(cudaMemsetAsync(h->b_bad.p, 0, 16, h->stream));
    CK(sslapb_launch_batch_globalize(d_rows, d_cols, idx_bytes, stride, d_nnzoff, d_rowoff, d_coloff, P, h->b_rows.as<int>(),
                                     h->b_cols.as<int>(), h->b_bad.as<int>(), h->sms, h->stream));
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, h->b_bad.p, 4, cudaMemcpyDeviceToHost, h->stream));
    // one block-diagonal CSR through the ordinary build (sorts on the device if a problem arrives unsorted)
    SslapbBuildFlags F;
    memset(&F, 0, sizeof F);
    int32_t NR = (int32_t)R, NC = (int32_t)Cn;
    int rc = build_from_coo(h, h->b_rows.p, h->b_cols.p, 4, 1, d_val, nnz, NR, NC, !maximize, SSLAPB_MEM_DEVICE_IN, F);
    if (rc) return rc;
    if (bad) return fail(h, SSLAPB_E_OUT_OF_RANGE, "a problem holds an index outside its own shape");
    if (F.empty_rows) return fail(h, SSLAPB_E_EMPTY_ROW, "a row of some problem has no valid entry");
    // state
    CK(h->price.reserve((size_t)Cn * 8)); CK(h->owner.reserve((size_t)Cn * 4));
    CK(h->bidkey.reserve((size_t)Cn * 8)); CK(h->winpos.reserve((size_t)Cn * 4));
    CK(h->p2o.reserve((size_t)R * 4)); CK(h->list.reserve((size_t)R * 4)); CK(h->mover.reserve((size_t)R * 4));
    CK(h->bidj.reserve((size_t)R * 4)); CK(h->bidv.reserve((size_t)R * 8)); CK(h->chosen.reserve((size_t)R * 8));
    CK(h->b_meta.reserve((size_t)P * sizeof(SslapbBatchMeta)));
    SslapbBatchParams B;
    B.P = P; B.rowoff = d_rowoff; B.coloff = d_coloff;
    B.rowptr = h->rowptr.as<long long>(); B.cols = h->cols.as<int>(); B.vals = h->vals.as<double>();
    B.eps_start = nullptr;
    if (eps_start) {
        CK(h->b_eps.reserve((size_t)P * 4));
        CK(cudaMemcpyAsync(h->b_eps.p, eps_start, (size_t)P * 4, cudaMemcpyHostToDevice, h->stream));
        B.eps_start = h->b_eps.as<float>();
    }
    B.max_iter = max_iter;
    B.price = h->price.as<double>(); B.owner = h->owner.as<int>(); B.bestkey = h->bidkey.as<unsigned long long>();
    B.winpos = h->winpos.as<int>(); B.p2o = h->p2o.as<int>(); B.list = h->list.as<int>(); B.mover = h->mover.as<int>();
    B.bidj = h->bidj.as<int>(); B.bidv = h->bidv.as<double>(); B.chosen = h->chosen.as<double>();
    B.meta = h->b_meta.as<SslapbBatchMeta>();
    CK(h->b_rec.reserve((size_t)Cn * sizeof(SslapbBatchRec)));
    B.brec = h->b_rec.as<SslapbBatchRec>();
    CK(cudaEventRecord(h->ev[3], h->stream));
    if (h->batch_v1) CK(sslapb_launch_auction_batch(&B, h->stream));
    else CK(sslapb_launch_auction_batch2(&B, h->stream));
    CK(cudaEventRecord(h->ev[4], h->stream));
    std::vector<SslapbBatchMeta> bm((size_t)P);
    std::vector<double> chosen((size_t)R);
    std::vector<int32_t> sol_tmp;
    int32_t *sol_host = sol_out;
    if (mem & SSLAPB_MEM_DEVICE_OUT) {
        CK(cudaMemcpyAsync(sol_out, B.p2o, (size_t)R * 4, cudaMemcpyDeviceToDevice, h->stream));
        sol_tmp.resize((size_t)R); sol_host = sol_tmp.data();
    }
    CK(cudaMemcpyAsync(sol_host, B.p2o, (size_t)R * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(chosen.data(), B.chosen, (size_t)R * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(bm.data(), B.meta, (size_t)P * sizeof(SslapbBatchMeta), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float setup_ms = 0, solve_ms = 0, h2d_ms = 0;
    cudaEventElapsedTime(&setup_ms, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&solve_ms, h->ev[3], h->ev[4]);
    cudaEventElapsedTime(&h2d_ms, h->ev[0], h->ev[1]);
    if (metas) {
        for (int p = 0; p < P; ++p) {
            sslapb_meta &m = metas[p];
            memset(&m, 0, sizeof m);
            double obj = 0.0;
            for (long long i = rowoff[p]; i < rowoff[p + 1]; ++i) {
                if (sol_host[i] == -1) continue;
                if (maximize) obj += chosen[(size_t)i]; else obj -= chosen[(size_t)i];
            }
            m.start_eps = bm[p].start_eps; m.final_eps = bm[p].final_eps; m.target_eps = bm[p].target_eps;
            m.eCE = bm[p].eCE; m.soln_found = bm[p].soln_found; m.its = bm[p].its; m.nreductions = bm[p].nreductions;
            m.n_assigned = bm[p].n_assigned; m.obj64 = obj; m.obj = (float)obj;
            m.setup_ms = setup_ms; m.solve_ms = solve_ms; m.h2d_ms = h2d_ms;      // times of the WHOLE batch
            m.cardinality = -1; m.n_rows = n_rows[p]; m.n_cols = n_cols[p]; m.nnz = nnzoff[p + 1] - nnzoff[p];
            m.stop_reason = bm[p].stop_reason;
        }
    }
    h->has_vals = false;                                       // the resident CSR is block-diagonal: not a single problem
    return SSLAPB_OK;
}
