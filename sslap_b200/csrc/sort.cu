// Device-side stable sort of a COO stream by row (sm_100a) — the "device-side sort/segment" of the CSR build.
//
// The reference requires `loc` to arrive row-sorted (cumulative_idxs, /root/reference/sslap/auction_.pyx:33-48) and
// silently produces garbage otherwise.  Here an unsorted stream is detected by the ingest pass and sorted on the GPU:
// LSD radix sort on the row index, 8 bits per pass, STABLE — entries of one row keep their input order, which matters
// because ties inside a row are broken by position (auction_.pyx:351).  Payload = original entry index; a final gather
// produces the sorted (row, col, value) stream that the normal ingest pass consumes.
//
// Per pass: (1) per-tile digit histogram, (2) exclusive scan over (digit, tile), (3) stable scatter: a tile is walked in
// sub-steps of one CTA width; inside a sub-step a thread's rank among equal digits is popc(match_any & lower lanes) plus
// the counts of lower warps (shared memory) plus the counts of earlier sub-steps.
#include "common.cuh"

#define SORT_THREADS 256
#define SORT_ITEMS 16
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)

template <typename IT>
__global__ void __launch_bounds__(256) sslapb_sort_prep_kernel(const IT *__restrict__ rows, long long stride, long long nnz,
                                                               unsigned *__restrict__ keys, unsigned *__restrict__ idx)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long k = gtid; k < nnz; k += nth) {
        keys[k] = (unsigned)rows[k * stride];                 // range was checked by the ingest pass
        idx[k] = (unsigned)k;
    }
}

__global__ void __launch_bounds__(SORT_THREADS) sslapb_sort_hist_kernel(const unsigned *__restrict__ keys, long long nnz,
                                                                        int shift, long long ntiles, long long *hist)
{
    __shared__ unsigned s_cnt[256];
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const long long base = tile * SORT_TILE;
        for (int it = 0; it < SORT_ITEMS; ++it) {
            const long long k = base + (long long)it * SORT_THREADS + threadIdx.x;
            if (k < nnz) atomicAdd(&s_cnt[(keys[k] >> shift) & 255u], 1u);
        }
        __syncthreads();
        hist[(long long)threadIdx.x * ntiles + tile] = s_cnt[threadIdx.x];   // digit-major, tile-minor
        __syncthreads();
    }
}

// In-place exclusive scan of n int64 values by ONE CTA of 1024 threads.
__global__ void __launch_bounds__(1024) sslapb_sort_scan_kernel(long long *a, long long n)
{
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + tid;
        const long long v = i < n ? a[i] : 0;
        long long incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const long long o = __shfl_up_sync(SSLAPB_FULL, incl, off);
            if (lane >= off) incl += o;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const long long w = s_w[lane];
            long long wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const long long o = __shfl_up_sync(SSLAPB_FULL, wi, off);
                if (lane >= off) wi += o;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const long long carry = s_carry;
        if (i < n) a[i] = carry + s_w[warp] + incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_w[31] + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SORT_THREADS) sslapb_sort_scatter_kernel(const unsigned *__restrict__ keys_in,
                                                                           const unsigned *__restrict__ idx_in,
                                                                           unsigned *__restrict__ keys_out,
                                                                           unsigned *__restrict__ idx_out, long long nnz,
                                                                           int shift, long long ntiles,
                                                                           const long long *__restrict__ hist)
{
    __shared__ unsigned s_wcount[SORT_THREADS / 32][256];
    __shared__ unsigned s_wbase[SORT_THREADS / 32][256];
    __shared__ long long s_run[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        s_run[tid] = hist[(long long)tid * ntiles + tile];    // where this tile's run of digit `tid` starts
        const long long base = tile * SORT_TILE;
        for (int it = 0; it < SORT_ITEMS; ++it) {
            for (int w = 0; w < SORT_THREADS / 32; ++w) s_wcount[w][tid] = 0;
            __syncthreads();
            const long long k = base + (long long)it * SORT_THREADS + tid;
            const bool valid = k < nnz;
            unsigned key = 0, id = 0, d = 0x100u + lane;       // invalid threads get a private pseudo-digit
            if (valid) { key = keys_in[k]; id = idx_in[k]; d = (key >> shift) & 255u; }
            const unsigned peers = __match_any_sync(SSLAPB_FULL, d);
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            if (valid && rank == 0) s_wcount[warp][d] = __popc(peers);
            __syncthreads();
            {                                                  // thread `tid` owns digit `tid`: prefix over the warps
                long long run = s_run[tid];
                for (int w = 0; w < SORT_THREADS / 32; ++w) {
                    const unsigned c = s_wcount[w][tid];
                    s_wbase[w][tid] = (unsigned)(run - s_run[tid]);
                    run += c;
                }
                __syncthreads();
                if (valid) {
                    const long long pos = s_run[d] + s_wbase[warp][d] + rank;
                    keys_out[pos] = key;
                    idx_out[pos] = id;
                }
                __syncthreads();
                s_run[tid] = run;
            }
            __syncthreads();
        }
    }
}

template <typename IT>
__global__ void __launch_bounds__(256) sslapb_sort_gather_kernel(const unsigned *__restrict__ keys,
                                                                 const unsigned *__restrict__ idx,
                                                                 const IT *__restrict__ cols, long long stride,
                                                                 const double *__restrict__ val, long long nnz,
                                                                 int *__restrict__ rows_s, int *__restrict__ cols_s,
                                                                 double *__restrict__ val_s)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long k = gtid; k < nnz; k += nth) {
        const long long src = idx[k];
        rows_s[k] = (int)keys[k];
        cols_s[k] = (int)cols[src * stride];
        if (val) val_s[k] = val[src];
    }
}

// Sorts (rows, cols, val) by row, stable.  keys[2], idx[2]: nnz uint32 each; hist: 256 * ntiles int64.
// Outputs rows_s / cols_s (int32) and val_s; n_rows bounds the number of radix passes.
extern "C" cudaError_t sslapb_launch_coo_sort(const void *rows, const void *cols, int idx_bytes, long long stride,
                                              const double *val, long long nnz, int n_rows, unsigned *keys0,
                                              unsigned *keys1, unsigned *idx0, unsigned *idx1, long long *hist,
                                              int *rows_s, int *cols_s, double *val_s, int sms, cudaStream_t stream)
{
    if (nnz <= 0) return cudaSuccess;
    const long long ntiles = (nnz + SORT_TILE - 1) / SORT_TILE;
    const int grid = (int)(ntiles < (long long)sms * 8 ? ntiles : (long long)sms * 8);
    const int g256 = (int)(((nnz + 255) / 256) < (long long)sms * 8 ? ((nnz + 255) / 256) : (long long)sms * 8);
    if (idx_bytes == 4) sslapb_sort_prep_kernel<int><<<g256, 256, 0, stream>>>((const int *)rows, stride, nnz, keys0, idx0);
    else sslapb_sort_prep_kernel<long long><<<g256, 256, 0, stream>>>((const long long *)rows, stride, nnz, keys0, idx0);
    int bits = 1;
    while (bits < 32 && (1ll << bits) < (long long)n_rows) ++bits;
    unsigned *kin = keys0, *kout = keys1, *iin = idx0, *iout = idx1;
    for (int shift = 0; shift < bits; shift += 8) {
        sslapb_sort_hist_kernel<<<grid, SORT_THREADS, 0, stream>>>(kin, nnz, shift, ntiles, hist);
        sslapb_sort_scan_kernel<<<1, 1024, 0, stream>>>(hist, 256 * ntiles);
        sslapb_sort_scatter_kernel<<<grid, SORT_THREADS, 0, stream>>>(kin, iin, kout, iout, nnz, shift, ntiles, hist);
        unsigned *t = kin; kin = kout; kout = t;
        t = iin; iin = iout; iout = t;
    }
    if (idx_bytes == 4)
        sslapb_sort_gather_kernel<int><<<g256, 256, 0, stream>>>(kin, iin, (const int *)cols, stride, val, nnz, rows_s, cols_s, val_s);
    else
        sslapb_sort_gather_kernel<long long><<<g256, 256, 0, stream>>>(kin, iin, (const long long *)cols, stride, val, nnz, rows_s, cols_s, val_s);
    return cudaGetLastError();
}
