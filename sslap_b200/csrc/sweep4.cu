// Four-rows-per-warp bidding sweep for sm_100a: 8 lanes per person, up to 4 aligned 16-byte chunks per lane.
//
// Same result per bidder, bit for bit, as the per-row kernel (auction.cu: sslapb_bid_sweep_kernel) and the oracle — the
// bidding loop of bid_and_assign (/root/reference/sslap/auction_.pyx:339-365): top-2 of a_ij - p_j over the CSR row, the
// LAST maximal entry wins (:351), w_i = second largest of the multiset (-inf for a single candidate, :344),
// bid = a_ibest - w_i + eps (:360), optional per-object atomicMax merge (:375-385).
//
// Why another schedule.  The per-row kernel (one warp = one row, 25 of 32 lanes busy at C3's ~100-entry rows) spends
// ~250 warp-instructions per row and keeps ONE row's 1.2 KB in flight per warp; ncu shows it neither bandwidth- nor
// issue-saturated (DRAM 40 %, issue 53 %): too few bytes in flight, too many instructions per byte.  Here a warp takes
// FOUR consecutive rows at once: lane (g, t) = (lane / 8, lane % 8) owns chunks t, t+8, t+16, t+24 of row g, requests all
// of them up front (up to 192 B per lane, ~5 KB per warp in flight), and every instruction of the per-lane part and of
// the cross-lane part (3-step shuffle butterflies inside the 8-lane groups instead of 5 REDUX over the warp) works for
// four rows.  Rows of more than 32 chunks, rows whose bound-pruned result is not proven exact and rows whose candidates
// are all at -inf take the exact generic sweep (row_bid<32>) afterwards, as in the other kernels.
#include "auction.cuh"
#include "rowsweep.cuh"

#ifndef SW4_THREADS
#define SW4_THREADS 512
#endif

struct Sw4Chunk { int4 cj; double va, vb, vc, vd; };

__device__ __forceinline__ Sw4Chunk sw4_load(const int *__restrict__ cols, const double *__restrict__ vals, long long ch, bool on)
{
    Sw4Chunk c;
    c.cj = make_int4(0, 0, 0, 0); c.va = c.vb = c.vc = c.vd = 0.0;
    if (on) {
        c.cj = sslapb_ldg_stream_i4(reinterpret_cast<const int4 *>(cols) + ch);
        asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(c.va), "=d"(c.vb), "=d"(c.vc), "=d"(c.vd) : "l"(vals + 4 * ch));
    }
    return c;
}

// running top-2 of one lane over its chunks (later chunks hold later entries: they win equal values, :351)
struct Sw4Lane { double b, s, bc; int bi, bj; };

__device__ __forceinline__ void sw4_chunk(Sw4Lane &L, const Sw4Chunk &c, const double *price, int off, int deg, double thr)
{
    const bool g0 = ((unsigned)off < (unsigned)deg) && (c.va >= thr);
    const bool g1 = ((unsigned)(off + 1) < (unsigned)deg) && (c.vb >= thr);
    const bool g2 = ((unsigned)(off + 2) < (unsigned)deg) && (c.vc >= thr);
    const bool g3 = ((unsigned)(off + 3) < (unsigned)deg) && (c.vd >= thr);
    double v0 = SSLAPB_NEG_INF, v1 = SSLAPB_NEG_INF, v2 = SSLAPB_NEG_INF, v3 = SSLAPB_NEG_INF;
    if (g0) v0 = c.va - price[c.cj.x];
    if (g1) v1 = c.vb - price[c.cj.y];
    if (g2) v2 = c.vc - price[c.cj.z];
    if (g3) v3 = c.vd - price[c.cj.w];
    const SslapbLaneTop lt = sslapb_lane_top2(v0, v1, v2, v3);
    if (lt.b >= L.b) {                                         // this chunk's best is at least as good and LATER in the row
        L.s = L.b > lt.s ? L.b : lt.s;
        L.b = lt.b;
        L.bi = off + lt.w;
        L.bc = (lt.w & 2) ? ((lt.w & 1) ? c.vd : c.vc) : ((lt.w & 1) ? c.vb : c.va);
        L.bj = (lt.w & 2) ? ((lt.w & 1) ? c.cj.w : c.cj.z) : ((lt.w & 1) ? c.cj.y : c.cj.x);
    } else {
        L.s = lt.b > L.s ? lt.b : L.s;
    }
}

__global__ void __launch_bounds__(SW4_THREADS, 1) sslapb_bid_sweep4_kernel(SslapbAuctionParams P, const int *__restrict__ bidders,
                                                                          int nb, float eps_f, int merge)
{
    const int lane = threadIdx.x & 31, g = lane >> 3, t = lane & 7;
    const int wpc = blockDim.x >> 5;
    const int gwarp = blockIdx.x * wpc + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * wpc;
    const double eps = (double)eps_f;
    const bool prune = (merge & 2) == 0;
    const double pmin = sslapb_key2double(P.ctrl->pmin_key[0]);
    double spread = sslapb_key2double(P.ctrl->pmax_key) - pmin;
    if (!(spread < 1.7e308) || !prune) spread = __longlong_as_double(0x7ff0000000000000ll);   // +inf: thr = -inf, no pruning
    merge &= 1;
    int n2nd = 0;
    // row info of the NEXT iteration is requested one iteration ahead
    int a = 4 * gwarp + g;
    long long st = 0; int deg = 0; double rmax = 0.0;
    if (a < nb) {
        const int i = bidders ? __ldg(bidders + a) : a;
        st = __ldg(P.rowptr + i); deg = (int)(__ldg(P.rowptr + i + 1) - st); rmax = __ldg(P.rowmax + i);
    }
    for (; 4 * (a >> 2) < nb; ) {                              // warp-uniform: the first position of the warp's group of four
        const int an = a + 4 * nwarps;
        long long stn = 0; int degn = 0; double rmaxn = 0.0;
        if (an < nb) {
            const int i = bidders ? __ldg(bidders + an) : an;
            stn = __ldg(P.rowptr + i); degn = (int)(__ldg(P.rowptr + i + 1) - stn); rmaxn = __ldg(P.rowmax + i);
        }
        const bool live = a < nb;
        const long long c0 = st >> 2, c1 = (st + deg + 3) >> 2;
        const bool fits = (c1 - c0) <= 32;                     // group-uniform
        const long long ch = c0 + t;
        const bool on = live && fits;
        // all chunks of the four rows are requested before anything is used
        const Sw4Chunk k0 = sw4_load(P.cols, P.vals, ch, on && ch < c1);
        const Sw4Chunk k1 = sw4_load(P.cols, P.vals, ch + 8, on && ch + 8 < c1);
        const Sw4Chunk k2 = sw4_load(P.cols, P.vals, ch + 16, on && ch + 16 < c1);
        const Sw4Chunk k3 = sw4_load(P.cols, P.vals, ch + 24, on && ch + 24 < c1);
        const double thr = rmax - spread;
        const int off = 4 * t - (int)(st & 3);                 // row index of slot 0 of chunk t (may be negative)
        Sw4Lane Ln;
        Ln.b = SSLAPB_NEG_INF; Ln.s = SSLAPB_NEG_INF; Ln.bc = 0.0; Ln.bi = -1; Ln.bj = -1;
        const int dg = on ? deg : 0;
        sw4_chunk(Ln, k0, P.price, off, dg, thr);
        sw4_chunk(Ln, k1, P.price, off + 32, dg, thr);
        sw4_chunk(Ln, k2, P.price, off + 64, dg, thr);
        sw4_chunk(Ln, k3, P.price, off + 96, dg, thr);
        // ---- cross-lane inside the 8-lane group: lexicographic max of (value, row index), then the second largest value
        const bool has = Ln.b > SSLAPB_NEG_INF;
        const int bi = has ? Ln.bi : -1;
        const unsigned long long bk = has ? sslapb_ord64(Ln.b + 0.0) : 0ull;     // + 0.0 folds -0.0 into +0.0
        const unsigned long long sk = has ? sslapb_ord64(Ln.s + 0.0) : 0ull;
        unsigned long long mk = bk;
        int mi = bi;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(SSLAPB_FULL, mk, o);
            const int oi = __shfl_xor_sync(SSLAPB_FULL, mi, o);
            const bool take = (ok > mk) || (ok == mk && oi > mi);
            mk = take ? ok : mk;
            mi = take ? oi : mi;
        }
        const bool iswin = has && bk == mk && bi == mi;
        unsigned long long ck = iswin ? sk : bk;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(SSLAPB_FULL, ck, o);
            ck = ok > ck ? ok : ck;
        }
        const unsigned own = (__ballot_sync(SSLAPB_FULL, iswin) >> (lane & 24)) & 0xffu;
        const int src = own ? ((lane & 24) + __ffs(own) - 1) : lane;
        const double bc = __shfl_sync(SSLAPB_FULL, Ln.bc, src);
        const int bj = __shfl_sync(SSLAPB_FULL, Ln.bj, src);
        const double wi = ck > SSLAPB_KEY_NEG_INF ? sslapb_key2double(ck) : SSLAPB_NEG_INF;   // :344
        double bid = (bc - wi) + eps;                          // :360
        const bool proven = !(thr > SSLAPB_NEG_INF) || ((thr - pmin) < wi);
        int j = (own && proven) ? bj : -1;
        if (on && own && !proven && t == 0) ++n2nd;
        // ---- rows the group pass cannot decide (long, unproven, all candidates at -inf): exact generic sweep, whole warp
        unsigned redo = __ballot_sync(SSLAPB_FULL, live && t == 0 && j < 0);
        while (redo) {
            const int gl = __ffs(redo) - 1;                    // lane 8 * group
            redo &= redo - 1;
            const long long rst = __shfl_sync(SSLAPB_FULL, st, gl);
            const int rdg = __shfl_sync(SSLAPB_FULL, deg, gl);
            const double rthr = __shfl_sync(SSLAPB_FULL, thr, gl);
            const bool rfits = __shfl_sync(SSLAPB_FULL, (int)fits, gl) != 0;
            int rj; double rbid;
            // a long row keeps the bound pruning of the generic sweep; a short one was either unproven or all at -inf
            row_bid<32>(P.cols, P.vals, P.price, rst, rst + rdg, lane, eps, rj, rbid, pmin, rfits ? SSLAPB_NEG_INF : rthr);
            if ((lane >> 3) == (gl >> 3)) { j = rj; bid = rbid; }
        }
        if (live && t == 0) {
            P.bidj[a] = j;
            P.bidv[a] = bid;
            if (merge && j >= 0) atomicMax(P.bidkey + j, sslapb_ord64(bid));
        }
        a = an; st = stn; deg = degn; rmax = rmaxn;
    }
    if (n2nd) atomicAdd((unsigned long long *)&P.ctrl->prune_second_pass, (unsigned long long)n2nd);
}

extern "C" cudaError_t sslapb_launch_bid_sweep4(const SslapbAuctionParams *P, const int *bidders, int nb, float eps,
                                                int merge, int grid, cudaStream_t stream)
{
    sslapb_bid_sweep4_kernel<<<grid, SW4_THREADS, 0, stream>>>(*P, bidders, nb, eps, merge);
    return cudaGetLastError();
}
