// Device-side COO / dense -> CSR build (sm_100a).
//
// Replaces the host-side adapters of the reference: _from_matrix (/root/reference/sslap/auction_.pyx:528-571),
// _from_sparse (:575-617), cumulative_idxs / diff / to_int_pointer (:33-95; gap-tolerant variant feasibility_.pyx:22-46),
// mult_ndarray_by (:112-118, WITHOUT mutating the caller's array) and max_val (:123-134).
//
// CSR layout in HBM: rowptr int64[N+1], cols int32[nnz+4], vals float64[nnz+4] (sign-folded: 'min' => negated).  The
// entry arrays are exactly the row-sorted COO stream, so the build is one streaming pass; the +4 slack lets the
// sweep kernels issue 16-byte-aligned vector loads that straddle row ends.
#include "common.cuh"
#include "build.cuh"


template <typename IT>
__global__ void __launch_bounds__(256) sslapb_coo_ingest_kernel(const IT *__restrict__ rows, const IT *__restrict__ cols,
                                                                long long stride, const double *__restrict__ val,
                                                                long long nnz, int N, int M, int negate,
                                                                int *__restrict__ cols32, double *__restrict__ vals,
                                                                long long *__restrict__ rowptr, SslapbBuildFlags *F)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    double amax = 0.0;
    for (long long k = gtid; k < nnz; k += nth) {
        const long long r = (long long)rows[k * stride];
        const long long c = (long long)cols[k * stride];
        if (r < 0 || r >= N || c < 0 || c >= M) { F->out_of_range = 1; continue; }
        const long long rp = k > 0 ? (long long)rows[(k - 1) * stride] : -1;
        if (rp > r) F->unsorted = 1;
        else if (rp < r && rp >= -1) {                 // segment boundary: rows rp+1..r start here
            for (long long rr = rp + 1; rr <= r; ++rr) rowptr[rr] = k;
            if (r - rp > 1) F->empty_rows = 1;
        }
        if (k == nnz - 1) {
            for (long long rr = r + 1; rr <= N; ++rr) rowptr[rr] = nnz;
            if (r < N - 1) F->empty_rows = 1;
        }
        cols32[k] = (int)c;
        if (val) {
            const double v = val[k];
            const double fv = negate ? v * -1.0 : v;   // mult_ndarray_by(val, -1)
            vals[k] = fv;
            amax = fmax(amax, fabs(fv));
        }
    }
    if (val) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) amax = fmax(amax, __shfl_xor_sync(SSLAPB_FULL, amax, off));
        if ((threadIdx.x & 31) == 0 && amax > 0.0) atomicMax(&F->maxabs, (unsigned long long)__double_as_longlong(amax));
    }
}

// ---- dense path: mat (N x M, row-major float64), entry valid iff v >= 0 (auction_.pyx:549, feasibility_.pyx:261)
__global__ void __launch_bounds__(256) sslapb_dense_count_kernel(const double *__restrict__ mat, int N, int M,
                                                                 long long *__restrict__ counts)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = gwarp; r < N; r += nwarps) {
        const double *row = mat + (long long)r * M;
        int cnt = 0;
        for (int c = lane; c < M; c += 32) cnt += (row[c] >= 0.0);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(SSLAPB_FULL, cnt, off);
        if (lane == 0) counts[r] = cnt;
    }
}

// In-place exclusive scan of n int64 counts (n = N, result has N+1 entries) by ONE CTA; total -> F->nnz.
__global__ void __launch_bounds__(1024) sslapb_scan_kernel(long long *a, int n, SslapbBuildFlags *F)
{
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_carry = 0; F->empty_rows = 0; }
    __syncthreads();
    int empty = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const long long v = i < n ? a[i] : 0;
        if (i < n && v == 0) empty = 1;
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
    if (tid == 0) { a[n] = s_carry; F->nnz = s_carry; }
    if (empty) F->empty_rows = 1;
}

__global__ void __launch_bounds__(256) sslapb_dense_fill_kernel(const double *__restrict__ mat, int N, int M, int negate,
                                                                const long long *__restrict__ rowptr,
                                                                int *__restrict__ cols32, double *__restrict__ vals,
                                                                SslapbBuildFlags *F)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    double amax = 0.0;
    for (int r = gwarp; r < N; r += nwarps) {
        const double *row = mat + (long long)r * M;
        long long pos = rowptr[r];
        for (int c0 = 0; c0 < M; c0 += 32) {               // row-major order is kept (the row-sorted precondition)
            const int c = c0 + lane;
            const double v = c < M ? row[c] : -1.0;
            const bool ok = v >= 0.0;
            const unsigned bal = __ballot_sync(SSLAPB_FULL, ok);
            if (ok) {
                const long long o = pos + __popc(bal & ((1u << lane) - 1u));
                const double fv = negate ? v * -1.0 : v;
                cols32[o] = c;
                if (vals) vals[o] = fv;
                amax = fmax(amax, fabs(fv));
            }
            pos += __popc(bal);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) amax = fmax(amax, __shfl_xor_sync(SSLAPB_FULL, amax, off));
    if (lane == 0 && amax > 0.0) atomicMax(&F->maxabs, (unsigned long long)__double_as_longlong(amax));
}

// max row / max col of a COO stream (N, M inference of AuctionSolver.__init__, auction_.pyx:209-210); out[0..1] start at 0
template <typename IT>
__global__ void __launch_bounds__(256) sslapb_index_max_kernel(const IT *__restrict__ rows, const IT *__restrict__ cols,
                                                               long long stride, long long nnz, long long *out)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    long long mr = -1, mc = -1, lo = 0;
    for (long long k = gtid; k < nnz; k += nth) {
        const long long r = (long long)rows[k * stride], c = (long long)cols[k * stride];
        mr = r > mr ? r : mr; mc = c > mc ? c : mc;
        lo = (r < lo) ? r : lo; lo = (c < lo) ? c : lo;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const long long a = __shfl_xor_sync(SSLAPB_FULL, mr, off), b = __shfl_xor_sync(SSLAPB_FULL, mc, off);
        const long long l = __shfl_xor_sync(SSLAPB_FULL, lo, off);
        mr = a > mr ? a : mr; mc = b > mc ? b : mc; lo = l < lo ? l : lo;
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo < 0) { mr = -1; mc = -1; atomicMin(out, -1ll); atomicMin(out + 1, -1ll); }   // negative index: poison
        else { atomicMax(out, mr); atomicMax(out + 1, mc); }
    }
}

extern "C" cudaError_t sslapb_launch_index_max(const void *rows, const void *cols, int idx_bytes, long long stride,
                                               long long nnz, long long *out, int sms, cudaStream_t stream)
{
    long long want = (nnz + 255) / 256;
    int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
    if (idx_bytes == 4)
        sslapb_index_max_kernel<int><<<grid, 256, 0, stream>>>((const int *)rows, (const int *)cols, stride, nnz, out);
    else
        sslapb_index_max_kernel<long long><<<grid, 256, 0, stream>>>((const long long *)rows, (const long long *)cols,
                                                                     stride, nnz, out);
    return cudaGetLastError();
}

// per-row maximum of the (sign-folded) values: warp per row
__global__ void __launch_bounds__(256) sslapb_rowmax_kernel(const long long *__restrict__ rowptr,
                                                            const double *__restrict__ vals, long long nrows,
                                                            double *__restrict__ rowmax, int *maxdeg)
{
    int dmax = 0;
    const int lane = threadIdx.x & 31;
    const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = gwarp; r < nrows; r += nwarps) {
        double m = SSLAPB_NEG_INF;
        for (long long e = rowptr[r] + lane; e < rowptr[r + 1]; e += 32) m = fmax(m, vals[e]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(SSLAPB_FULL, m, off));
        if (lane == 0) rowmax[r] = m;
        dmax = max(dmax, (int)min(rowptr[r + 1] - rowptr[r], 0x7fffffffll));
    }
    if (lane == 0 && dmax > 0) atomicMax(maxdeg, dmax);
}

extern "C" cudaError_t sslapb_launch_rowmax(const long long *rowptr, const double *vals, long long nrows, double *rowmax,
                                            int *maxdeg, int sms, cudaStream_t stream)
{
    sslapb_rowmax_kernel<<<sms * 8, 256, 0, stream>>>(rowptr, vals, nrows, rowmax, maxdeg);
    return cudaGetLastError();
}

extern "C" cudaError_t sslapb_launch_coo_ingest(const void *rows, const void *cols, int idx_bytes, long long stride,
                                                const double *val, long long nnz, int N, int M, int negate,
                                                int *cols32, double *vals, long long *rowptr, SslapbBuildFlags *F,
                                                int sms, cudaStream_t stream)
{
    if (nnz <= 0) return cudaSuccess;
    long long want = (nnz + 255) / 256;
    int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
    if (idx_bytes == 4)
        sslapb_coo_ingest_kernel<int><<<grid, 256, 0, stream>>>((const int *)rows, (const int *)cols, stride, val, nnz, N,
                                                                M, negate, cols32, vals, rowptr, F);
    else
        sslapb_coo_ingest_kernel<long long><<<grid, 256, 0, stream>>>((const long long *)rows, (const long long *)cols,
                                                                      stride, val, nnz, N, M, negate, cols32, vals,
                                                                      rowptr, F);
    return cudaGetLastError();
}

extern "C" cudaError_t sslapb_launch_dense_count(const double *mat, int N, int M, long long *rowptr,
                                                 SslapbBuildFlags *F, int sms, cudaStream_t stream)
{
    sslapb_dense_count_kernel<<<sms * 8, 256, 0, stream>>>(mat, N, M, rowptr);
    sslapb_scan_kernel<<<1, 1024, 0, stream>>>(rowptr, N, F);
    return cudaGetLastError();
}

extern "C" cudaError_t sslapb_launch_dense_fill(const double *mat, int N, int M, int negate, const long long *rowptr,
                                                int *cols32, double *vals, SslapbBuildFlags *F, int sms,
                                                cudaStream_t stream)
{
    sslapb_dense_fill_kernel<<<sms * 8, 256, 0, stream>>>(mat, N, M, negate, rowptr, cols32, vals, F);
    return cudaGetLastError();
}
