// Hot lists: per person the (up to) 32 entries of its CSR row with the largest values a_ij, 16 bytes each, 512 bytes per
// row — the rows the few-bidder / chain rounds and the mid-sized grid rounds read instead of the full CSR row.
//
// Why it is exact.  The bidding loop (/root/reference/sslap/auction_.pyx:339-365) needs the largest and second largest
// of v_k = a_ik - p_k over the row.  Let S be the hot entries and R the others, and rest_i >= fl(a_ik - p_k) for every
// k in R (taken at some earlier moment; prices never decrease, and IEEE subtraction is monotone, so the bound stays
// valid).  If the second largest value found inside S is strictly greater than rest_i, no entry of R can be the best
// or the second best, and none can tie with them, so (best entry, w_i) of S alone are those of the whole row — ties
// inside S are resolved by the original row index ("last maximal entry", :351).  Otherwise the caller sweeps the whole
// row.  rest_i is recomputed by every eps-CS sweep (end of each eps-phase, all prices gathered anyway).
//
// Measured on C3 (tools/analysis/hotlist_stats.py, the oracle with a statistics hook): from the third eps-phase on
// 98-100 % of the bids are decided by the hot list; in the first two phases (eps = C/2, 0.15 C/2: prices move by more
// than the gaps between a row's top values) practically none is, so the kernel probes once per phase and switches the
// hot path on or off for the rest of the phase.
#include "auction.cuh"
#include "rowsweep.cuh"

// S = { k : a_ik > v33 }, v33 = the 33rd largest value of the row (rows of at most 32 entries: S = the row, threshold
// NaN so that "a <= threshold" — the test for R — is false for every entry).  Warp per row; v33 by a bitwise radix
// select on the order-preserving keys (64 counting steps).
__global__ void __launch_bounds__(256) sslapb_hot_build_kernel(const long long *__restrict__ rowptr, const int *__restrict__ cols,
                                                               const double *__restrict__ vals, int N,
                                                               SslapbHotEnt *__restrict__ hot, double *__restrict__ hthr)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int i = gwarp; i < N; i += nwarps) {
        const long long st = __ldg(rowptr + i);
        const int deg = (int)(__ldg(rowptr + i + 1) - st);
        SslapbHotEnt pad;
        pad.col = 0; pad.idx = -1; pad.a = SSLAPB_NEG_INF;         // gathered like a real entry: -inf - p = -inf never wins
        if (deg <= 32) {
            SslapbHotEnt e = pad;
            if (lane < deg) { e.col = __ldg(cols + st + lane); e.idx = lane; e.a = __ldg(vals + st + lane); }
            hot[(long long)i * 32 + lane] = e;
            if (lane == 0) hthr[i] = __longlong_as_double(0x7ff8000000000000ll);
            continue;
        }
        unsigned long long T = 0ull;
        if (deg <= 128) {                                      // keys in registers
            unsigned long long k[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) k[q] = (lane + 32 * q < deg) ? sslapb_key_of(__ldg(vals + st + lane + 32 * q)) : 0ull;
            for (int bit = 63; bit >= 0; --bit) {
                const unsigned long long cand = T | (1ull << bit);
                const int c = (k[0] >= cand) + (k[1] >= cand) + (k[2] >= cand) + (k[3] >= cand);
                if (__reduce_add_sync(SSLAPB_FULL, c) >= 33) T = cand;
            }
        } else {
            for (int bit = 63; bit >= 0; --bit) {
                const unsigned long long cand = T | (1ull << bit);
                int c = 0;
                for (int e = lane; e < deg; e += 32) c += sslapb_key_of(__ldg(vals + st + e)) >= cand;
                if (__reduce_add_sync(SSLAPB_FULL, c) >= 33) T = cand;
            }
        }
        int base = 0;
        for (int e0 = 0; e0 < deg; e0 += 32) {                 // warp-uniform trip count
            const int e = e0 + lane;
            double a = 0.0;
            bool sel = false;
            if (e < deg) { a = __ldg(vals + st + e); sel = sslapb_key_of(a) > T; }
            const unsigned bal = __ballot_sync(SSLAPB_FULL, sel);
            if (sel) {
                SslapbHotEnt h;
                h.col = __ldg(cols + st + e); h.idx = e; h.a = a;
                hot[(long long)i * 32 + base + __popc(bal & lt_mask)] = h;
            }
            base += __popc(bal);
        }
        if (lane >= base) hot[(long long)i * 32 + lane] = pad;
        if (lane == 0) hthr[i] = sslapb_key2double(T);
    }
}

extern "C" cudaError_t sslapb_launch_hot_build(const long long *rowptr, const int *cols, const double *vals, int N,
                                               SslapbHotEnt *hot, double *hthr, int sms, cudaStream_t stream)
{
    sslapb_hot_build_kernel<<<sms * 8, 256, 0, stream>>>(rowptr, cols, vals, N, hot, hthr);
    return cudaGetLastError();
}
