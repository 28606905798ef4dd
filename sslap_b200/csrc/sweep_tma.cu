// Streamed full-frontier bidding sweep for sm_100a (EXPERIMENTAL, opt-in through bit 2 of sslapb_bid_sweep's `merge`):
// the CSR goes HBM -> shared memory by TMA bulk copies, rows are swept out of registers by 8-lane groups.
//
// This is the bidding loop of bid_and_assign (/root/reference/sslap/auction_.pyx:339-365) for the frontier "every person
// bids" (the first round of every eps-phase, :256-262) — same (object, bid) per person, bit for bit, as the per-row
// kernel (sslapb_bid_sweep_kernel) and the oracle (tests/test_gpu_parity.py::test_bid_sweep_streamed_bit_exact).
//
// Design
//   * one producer thread per CTA streams the CTA's contiguous slice of `cols`/`vals` through an 8-stage ring of 24 KB
//     stages with cp.async.bulk (UBLKCP) + mbarrier transaction counts; no load instruction and no HBM latency on the
//     consumer side;
//   * a consumer warp takes four rows per step (8 lanes per row, up to 4 chunks = 16 entries per lane), copies them
//     shared -> registers, hands the ring back at once (fence + mbarrier arrive), then runs masks, pruned price gathers
//     (all in flight together), a per-lane top-2 tournament and 3 shuffle steps; the lane holding the winner writes;
//   * rows are owned by the CTA whose chunk range contains their first entry; the boundaries (first row per CTA) are a
//     static function of the CSR (sslapb_sweep_plan_kernel).  Steps containing a row of more than 32 chunks, rows whose
//     pruned result is not proven exact and rows whose candidates are all at -inf take the exact generic sweep
//     (row_bid<32>) from global memory.
// Measured on B200 at C3 (profiles/r1_sweep_notes.md): 55-60 us against 49 us for the per-row kernel, at ~200 against
// 254 warp-instructions per row.  The ring takes 192 KB of the SM's 256 KB, which leaves ~60 KB of L1 for the price
// gathers (hit rate 7 %) and room for only 15 consumer warps at the ~120 registers the register-resident step needs;
// smaller rings are slower still (96 KB: 68 us, 48 KB: 86 us).  The per-row kernel, which keeps all of L1 for the gathers
// and runs 32 warps, stays the default.
#include "auction.cuh"
#include "rowsweep.cuh"

#ifndef TS_S
#define TS_S 512                          // 16-byte chunks (4 entries) per stage: 8 KB of columns + 16 KB of values
#endif
#ifndef TS_NS
#define TS_NS 8                           // stages in the ring
#endif
#define TS_RING (TS_S * TS_NS)            // chunks in the ring (power of two)
#ifndef TS_NCW
#define TS_NCW 15                         // consumer warps; warp TS_NCW is the producer
#endif
#define TS_THREADS ((TS_NCW + 1) * 32)
#define TS_SMEM (TS_RING * 48 + 2 * TS_NS * 8)

__device__ __forceinline__ unsigned ts_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ts_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ts_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ts_mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool ts_mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// try_wait suspends the thread in hardware for a bounded time; a barrier that never completes is a bug -> trap instead
// of hanging the device
__device__ __forceinline__ void ts_mbar_wait(unsigned bar, unsigned parity)
{
    for (unsigned spin = 0; !ts_mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 20)) __trap();
}
__device__ __forceinline__ void ts_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// First row of every CTA's slice: cta_row[b] = first row whose first entry lies at or after chunk TS_S * (nst * b / G).
__global__ void sslapb_sweep_plan_kernel(const long long *__restrict__ rowptr, int N, int G, int *__restrict__ cta_row)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > G) return;
    const long long nnz = rowptr[N];
    const long long nst = (((nnz + 3) >> 2) + TS_S - 1) / TS_S;
    const long long target = 4ll * TS_S * (nst * b / G);
    int lo = 0, hi = N;                                        // first i in [0, N] with rowptr[i] >= target
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rowptr[mid] >= target) hi = mid; else lo = mid + 1;
    }
    cta_row[b] = lo;
}

__device__ __forceinline__ int4 ts_lds_i4(unsigned a)
{
    int4 r;
    asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ double2 ts_lds_d2(unsigned a)
{
    double2 r;
    asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a));
    return r;
}

// Top-2 of one 4-entry chunk merged into the lane's running (best, second, row index of best).  Entries are in row
// order and a later entry wins an equal value ("last maximal entry", auction_.pyx:351); masked slots carry -inf and
// never become the best (`> -inf` test; a row whose candidates are all at -inf goes to the exact generic sweep).
__device__ __forceinline__ void ts_fold(double v0, double v1, double v2, double v3, int off, double &bv, double &sv, int &bi)
{
    const SslapbLaneTop lt = sslapb_lane_top2(v0, v1, v2, v3);
    const bool take = lt.b > SSLAPB_NEG_INF && lt.b >= bv;
    const double lose = take ? bv : lt.b;
    sv = fmax(take ? lt.s : sv, lose);
    bv = take ? lt.b : bv;
    bi = take ? off + lt.w : bi;
}

__global__ void __launch_bounds__(TS_THREADS, 1) sslapb_bid_sweep_tma_kernel(SslapbAuctionParams P, const int *__restrict__ cta_row,
                                                                            float eps_f, int merge)
{
    constexpr int W = 8, RPW = 4, MAXT = 4;                    // lanes per row, rows per warp step, chunks per lane
    extern __shared__ __align__(128) unsigned char ts_smem[];
    const unsigned cols_u32 = ts_smem_u32(ts_smem), vals_u32 = cols_u32 + TS_RING * 16;
    const unsigned full_u32 = cols_u32 + TS_RING * 48, empty_u32 = full_u32 + TS_NS * 8;

    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(SSLAPB_FULL, tid >> 5, 0);
    const int G = gridDim.x, b = blockIdx.x;
    const int R0 = cta_row[b], R1 = cta_row[b + 1];
    if (R0 >= R1) return;                                      // CTA-uniform
    const long long nnz = __ldg(P.rowptr + P.N);
    const long long TC = (nnz + 3) >> 2;                       // chunks in the CSR
    const long long nst = (TC + TS_S - 1) / TS_S;
    const long long Cs = (nst * b / G) * TS_S;                 // first chunk of this CTA's slice
    const long long E0 = Cs << 2;                              // its first entry: everything below is relative to it (int32)
    const int nload = (int)((((__ldg(P.rowptr + R1) + 3) >> 2) - Cs + TS_S - 1) / TS_S);   // through the end of the last owned row

    if (tid == 0) {
        for (int s = 0; s < TS_NS; ++s) { ts_mbar_init(full_u32 + 8 * s, 1); ts_mbar_init(empty_u32 + 8 * s, TS_NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == TS_NCW) {                                      // ---- producer
        if (lane == 0) {
            for (int s = 0; s < nload; ++s) {
                const int slot = s & (TS_NS - 1);
                if (s >= TS_NS) ts_mbar_wait(empty_u32 + 8 * slot, ((s / TS_NS) - 1) & 1);
                const long long cb = Cs + (long long)s * TS_S;
                const unsigned n = (unsigned)min((long long)TS_S, TC - cb);
                ts_mbar_expect_tx(full_u32 + 8 * slot, n * 48u);
                ts_bulk_g2s(cols_u32 + slot * (TS_S * 16), P.cols + 4 * cb, n * 16u, full_u32 + 8 * slot);
                ts_bulk_g2s(vals_u32 + slot * (TS_S * 32), P.vals + 4 * cb, n * 32u, full_u32 + 8 * slot);
            }
        }
        return;
    }

    // ---- consumers: group g (8 lanes) of warp w sweeps row R0 + 4 * (w + TS_NCW * t) + g
    const int g = lane / W, q = lane % W;
    const double eps = (double)eps_f;
    const bool prune = (merge & 2) == 0;
    merge &= 1;
    const double *__restrict__ price = P.price;
    const double pmin = sslapb_key2double(P.ctrl->pmin_key[0]);
    double spread = sslapb_key2double(P.ctrl->pmax_key) - pmin;
    if (!(spread < 1.7e308)) spread = __longlong_as_double(0x7ff0000000000000ll);
    const double NEG = SSLAPB_NEG_INF;
    int acq = 0, rel = 0, n2nd = 0;

    // stages [rel, upto) are handed back to the producer.  A stage is always acquired before it is released, needed or
    // not: an arrival for stage s + TS_NS must not be counted in the phase of stage s.
    auto release_upto = [&](int upto) {
        for (; rel < upto; ++rel) {
            if (acq <= rel) { ts_mbar_wait(full_u32 + 8 * (rel & (TS_NS - 1)), (rel / TS_NS) & 1); acq = rel + 1; }
            if (lane == 0) ts_mbar_arrive(empty_u32 + 8 * (rel & (TS_NS - 1)));
        }
    };

    int i = R0 + RPW * warp + g;
    if (R0 + RPW * warp >= R1) i = -1;                         // this warp has no row at all
    long long st64 = E0, en64 = E0;                            // row bounds (absolute; converted when used so that the
    double rmax = 0.0;                                         // loads issued one step ahead are not waited for early)
    if (i >= 0 && i < R1) { st64 = __ldg(P.rowptr + i); en64 = __ldg(P.rowptr + i + 1); rmax = __ldg(P.rowmax + i); }
    while (i >= 0) {                                           // warp-uniform: i < 0 only when the whole step is absent
        const bool valid = i < R1;
        // next step's offsets, one step ahead
        int in = i + RPW * TS_NCW;
        if (in - g >= R1) in = -1;
        long long stn = E0, enn = E0;
        double rmaxn = 0.0;
        if (in >= 0 && in < R1) { stn = __ldg(P.rowptr + in); enn = __ldg(P.rowptr + in + 1); rmaxn = __ldg(P.rowmax + in); }
        const int st = (int)(st64 - E0), en = (int)(en64 - E0);   // relative to the CTA's first entry

        const int c0 = st >> 2, c1 = (en + 3) >> 2;            // chunks of the row, relative to Cs
        const int nch = valid ? c1 - c0 : 0;
        const int deg = en - st;
        bool redo = false;
        if (!__any_sync(SSLAPB_FULL, nch > W * MAXT)) {
            // ---- the step's chunks: shared memory -> registers, then the ring is free again
            release_upto(min((__shfl_sync(SSLAPB_FULL, st, 0) >> 2) / TS_S, nload));
            const int s_need = __reduce_max_sync(SSLAPB_FULL, valid ? (c1 + TS_S - 1) / TS_S : 0);
            while (acq < s_need) { ts_mbar_wait(full_u32 + 8 * (acq & (TS_NS - 1)), (acq / TS_NS) & 1); ++acq; }
            int4 cj[MAXT];
            double2 va[MAXT], vb[MAXT];
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                const int c = c0 + q + W * t;
                // unconditional: chunks past the row (or of an absent row, deg = 0) are masked slot by slot below
                const unsigned rc = (unsigned)c & (TS_RING - 1);
                cj[t] = ts_lds_i4(cols_u32 + (rc << 4)); va[t] = ts_lds_d2(vals_u32 + (rc << 5)); vb[t] = ts_lds_d2(vals_u32 + (rc << 5) + 16);
            }
            // hand the ring back before the long part of the step.  The fence keeps the shared-memory loads above the
            // arrival (without it ptxas sinks / re-executes them below the release: observed as rare wrong rows)
            __threadfence_block();
            __syncwarp();
            {   // first stage the NEXT step touches (its offsets were requested at the top of this step)
                const int in0 = __shfl_sync(SSLAPB_FULL, in, 0);
                release_upto(in0 < 0 ? nload : min((int)(((__shfl_sync(SSLAPB_FULL, stn, 0) - E0) >> 2) / TS_S), nload));
            }
            // ---- 16 candidates per lane: masks, pruned price gathers (all in flight together), top-2
            const double thr = prune ? rmax - spread : NEG;
            const int off0 = ((c0 + q) << 2) - st;             // row index of slot 0 of chunk 0 (may be negative)
            double v[MAXT][4];
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                const int off = off0 + 4 * W * t;              // slots outside [0, deg) belong to other rows / absent chunks
                const bool m0 = (unsigned)off < (unsigned)deg && va[t].x >= thr, m1 = (unsigned)(off + 1) < (unsigned)deg && va[t].y >= thr;
                const bool m2 = (unsigned)(off + 2) < (unsigned)deg && vb[t].x >= thr, m3 = (unsigned)(off + 3) < (unsigned)deg && vb[t].y >= thr;
                v[t][0] = NEG; v[t][1] = NEG; v[t][2] = NEG; v[t][3] = NEG;
                if (m0) v[t][0] = va[t].x - price[cj[t].x];
                if (m1) v[t][1] = va[t].y - price[cj[t].y];
                if (m2) v[t][2] = vb[t].x - price[cj[t].z];
                if (m3) v[t][3] = vb[t].y - price[cj[t].w];
            }
            double bv = NEG, sv = NEG;                         // best / second-best value seen by this lane
            int bi = -1;                                       // row index of the best
#pragma unroll
            for (int t = 0; t < MAXT; ++t) ts_fold(v[t][0], v[t][1], v[t][2], v[t][3], off0 + 4 * W * t, bv, sv, bi);
            // combine the 8 lanes of the group: lexicographic max of (value, row index), second = best of the rest
#pragma unroll
            for (int o = 1; o < W; o <<= 1) {
                const double ob = __shfl_xor_sync(SSLAPB_FULL, bv, o), os = __shfl_xor_sync(SSLAPB_FULL, sv, o);
                const int oi = __shfl_xor_sync(SSLAPB_FULL, bi, o);
                const bool take = (ob > bv) || (ob == bv && oi > bi);
                sv = take ? fmax(os, bv) : fmax(sv, ob);
                bv = take ? ob : bv;
                bi = take ? oi : bi;
            }
            const bool proven = !(thr > NEG) || ((thr - pmin) < sv);   // no skipped entry can matter (row_bid_pruned)
            if (valid && !proven) ++n2nd;
            redo = valid && !(bi >= 0 && bv > NEG && proven);  // also: every candidate at -inf -> exact generic sweep
            // the lane that holds the winning entry writes the bid
            const int wc = ((st + bi) >> 2) - c0;              // chunk of the winner inside the row
            if (valid && !redo && (wc & (W - 1)) == q) {
                const int wt = wc / W, k = (st + bi) & 3;
                int4 wj = cj[0]; double2 wa = va[0], wb = vb[0];
#pragma unroll
                for (int t = 1; t < MAXT; ++t) if (wt == t) { wj = cj[t]; wa = va[t]; wb = vb[t]; }
                const int j = (k & 2) ? ((k & 1) ? wj.w : wj.z) : ((k & 1) ? wj.y : wj.x);
                const double a = (k & 2) ? ((k & 1) ? wb.y : wb.x) : ((k & 1) ? wa.y : wa.x);
                const double bid = (a - sv) + eps;             // :360 (sv = -inf for a single-candidate row, :344)
                P.bidj[i] = j;
                P.bidv[i] = bid;
                if (merge) atomicMax(P.bidkey + j, sslapb_ord64(bid));
            }
        } else {
            redo = valid;                                      // a row of more than 32 chunks in this step
        }
        // exact generic sweep from global memory, one flagged row at a time by the whole warp
        unsigned rm = __ballot_sync(SSLAPB_FULL, redo && q == 0);
        while (rm) {
            const int src = __ffs(rm) - 1;
            rm &= rm - 1;
            const int ri = __shfl_sync(SSLAPB_FULL, i, src);
            const long long rst = E0 + __shfl_sync(SSLAPB_FULL, st, src), ren = E0 + __shfl_sync(SSLAPB_FULL, en, src);
            const double rrmax = __shfl_sync(SSLAPB_FULL, rmax, src);
            const bool rlong = (((ren + 3) >> 2) - (rst >> 2)) > W * MAXT;
            int j; double bid;
            row_bid<32>(P.cols, P.vals, P.price, rst, ren, lane, eps, j, bid, pmin, (rlong && prune) ? rrmax - spread : NEG);
            if (lane == 0) {
                P.bidj[ri] = j;
                P.bidv[ri] = bid;
                if (merge && j >= 0) atomicMax(P.bidkey + j, sslapb_ord64(bid));
            }
        }
        i = in; st64 = stn; en64 = enn; rmax = rmaxn;
    }
    release_upto(nload);
    if (n2nd && q == 0) atomicAdd((unsigned long long *)&P.ctrl->prune_second_pass, (unsigned long long)n2nd);
}

extern "C" cudaError_t sslapb_launch_sweep_plan(const SslapbAuctionParams *P, int grid, int *cta_row, cudaStream_t stream)
{
    sslapb_sweep_plan_kernel<<<(grid + 1 + 127) / 128, 128, 0, stream>>>(P->rowptr, P->N, grid, cta_row);
    return cudaGetLastError();
}

extern "C" cudaError_t sslapb_launch_bid_sweep_tma(const SslapbAuctionParams *P, const int *cta_row, float eps, int merge,
                                                   int grid, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(sslapb_bid_sweep_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM);
    if (e != cudaSuccess) return e;
    sslapb_bid_sweep_tma_kernel<<<grid, TS_THREADS, TS_SMEM, stream>>>(*P, cta_row, eps, merge);
    return cudaGetLastError();
}
