// Lean full-row bidding sweep for sm_100a: the per-row kernel (auction.cu: sslapb_bid_sweep_kernel) rebuilt for occupancy.
//
// Same result per bidder, bit for bit — the bidding loop of bid_and_assign (/root/reference/sslap/auction_.pyx:339-365):
// top-2 of a_ij - p_j over the CSR row, the LAST maximal entry wins (:351), w_i = second largest of the multiset (:344),
// bid = a_ibest - w_i + eps (:360), per-object atomicMax merge (:375-385).
//
// ncu on the per-row kernel (profiles/r2_bid_sweep_stream_raw.csv): 271 warp-instructions per row, 64 registers = 32 warps
// per SM, issue slots 58 % busy, DRAM 40 % — between issue-bound and latency-bound.  The hot-form pass (auction.cu) showed
// what this access pattern wants: few registers, many warps.  Here every row is still read in full (12 B per entry), but
//   * the four entries of a lane are folded into a running (best, second, index) one after the other instead of a
//     two-level tournament over four live values — fewer live registers (40: 48 warps per SM), ~half the instructions;
//   * everything that is not the common case — rows of more than one warp pass, bound-pruned results that are not proven
//     exact, rows whose candidates are all at -inf — goes to a redo list and through the per-row kernel's exact code in
//     a second, tiny launch (sslapb_bid_sweep_redo_kernel), so none of that code costs registers here.
#include "auction.cuh"
#include "rowsweep.cuh"

__global__ void __launch_bounds__(512, 3) sslapb_bid_sweep_lean_kernel(SslapbAuctionParams P, const int *__restrict__ bidders, int nb,
                                                                      float eps_f, int merge)
{
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int gwarp = blockIdx.x * wpc + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * wpc;
    const double eps = (double)eps_f;
    const bool prune = (merge & 2) == 0;
    const double pmin = sslapb_key2double(P.ctrl->pmin_key[0]);
    double spread = sslapb_key2double(P.ctrl->pmax_key) - pmin;
    if (!(spread < 1.7e308) || !prune) spread = __longlong_as_double(0x7ff0000000000000ll);   // +inf: thr = -inf, no pruning
    merge &= 1;
    for (int a = gwarp; a < nb; a += nwarps) {
        const int i = bidders ? __ldg(bidders + a) : a;
        const long long st = __ldg(P.rowptr + i);
        const int deg = (int)(__ldg(P.rowptr + i + 1) - st);
        const double thr = __ldg(P.rowmax + i) - spread;
        const long long c0 = st >> 2, c1 = (st + deg + 3) >> 2;
        int j = -1;
        double bid = 0.0;
        if (c1 - c0 <= 32) {                                   // warp-uniform
            const long long ch = c0 + lane;
            int4 cj = make_int4(0, 0, 0, 0);
            double2 va = make_double2(0.0, 0.0), vb = va;
            if (ch < c1) {
                cj = sslapb_ldg_stream_i4(reinterpret_cast<const int4 *>(P.cols) + ch);
                va = sslapb_ldg_stream_d2(reinterpret_cast<const double2 *>(P.vals) + 2 * ch);
                vb = sslapb_ldg_stream_d2(reinterpret_cast<const double2 *>(P.vals) + 2 * ch + 1);
            }
            const int off = 4 * lane - (int)(st & 3);          // row index of slot 0 (may be negative)
            // all four gathers go out together (predicated); the values are then folded one after the other into a running
            // (best, second, slot) — a later slot wins an equal value ("last maximal entry", :351)
            const bool g0 = ((unsigned)off < (unsigned)deg) && (va.x >= thr);
            const bool g1 = ((unsigned)(off + 1) < (unsigned)deg) && (va.y >= thr);
            const bool g2 = ((unsigned)(off + 2) < (unsigned)deg) && (vb.x >= thr);
            const bool g3 = ((unsigned)(off + 3) < (unsigned)deg) && (vb.y >= thr);
            double v0 = SSLAPB_NEG_INF, v1 = SSLAPB_NEG_INF, v2 = SSLAPB_NEG_INF, v3 = SSLAPB_NEG_INF;
            if (g0) v0 = va.x - P.price[cj.x];
            if (g1) v1 = va.y - P.price[cj.y];
            if (g2) v2 = vb.x - P.price[cj.z];
            if (g3) v3 = vb.y - P.price[cj.w];
            double b = v0, s = SSLAPB_NEG_INF;
            int w = 0;
            if (v1 >= b) { s = b; b = v1; w = 1; } else s = v1;
            if (v2 >= b) { s = b; b = v2; w = 2; } else if (v2 > s) s = v2;
            if (v3 >= b) { s = b; b = v3; w = 3; } else if (v3 > s) s = v3;
            const int bi = off + w;
            const double bc = (w & 2) ? ((w & 1) ? vb.y : vb.x) : ((w & 1) ? va.y : va.x);
            const int bj = (w & 2) ? ((w & 1) ? cj.w : cj.z) : ((w & 1) ? cj.y : cj.x);
            const bool has = b > SSLAPB_NEG_INF;               // (real -inf candidates: the redo pass decides those rows)
            const unsigned long long bk = has ? sslapb_ord64(b + 0.0) : 0ull;   // + 0.0 folds -0.0 into +0.0
            const unsigned bh = (unsigned)(bk >> 32), bl = (unsigned)bk;
            const unsigned khi = __reduce_max_sync(SSLAPB_FULL, bh);
            const unsigned hm = __ballot_sync(SSLAPB_FULL, (bh == khi) & has);
            bool iswin;
            unsigned own;
            if ((hm & (hm - 1u)) == 0u) {                      // at most one lane holds the maximal high word
                own = hm;
                iswin = (bh == khi) & has;
            } else {
                const unsigned klo = __reduce_max_sync(SSLAPB_FULL, bh == khi ? bl : 0u);
                const bool top = (bh == khi) & (bl == klo) & has;
                const int widx = __reduce_max_sync(SSLAPB_FULL, top ? bi : -1);
                iswin = top & (bi == widx);
                own = __ballot_sync(SSLAPB_FULL, iswin);
            }
            if (own) {
                const unsigned long long sk = (has & (s > SSLAPB_NEG_INF)) ? sslapb_ord64(s + 0.0) : 0ull;
                const unsigned long long cand = iswin ? sk : bk;
                const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
                const unsigned shi = __reduce_max_sync(SSLAPB_FULL, chh);
                const unsigned slo = __reduce_max_sync(SSLAPB_FULL, chh == shi ? chl : 0u);
                const unsigned long long skey = ((unsigned long long)shi << 32) | slo;
                int src;
                asm("bfind.u32 %0, %1;" : "=r"(src) : "r"(own));
                const double wbc = __shfl_sync(SSLAPB_FULL, bc, src);
                const int wbj = __shfl_sync(SSLAPB_FULL, bj, src);
                const double wi = skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(skey) : SSLAPB_NEG_INF;   // :344
                const bool proven = !(thr > SSLAPB_NEG_INF) || ((thr - pmin) < wi);
                if (proven) { j = wbj; bid = (wbc - wi) + eps; }                                           // :360
            }
        }
        if (lane == 0) {
            if (j >= 0) {
                P.bidj[a] = j;
                P.bidv[a] = bid;
                if (merge) atomicMax(P.bidkey + j, sslapb_ord64(bid));
            } else {
                P.mover[atomicAdd(&P.ctrl->hot_probe_fail, 1)] = a;     // redo list (sslapb_bid_sweep_redo_kernel)
            }
        }
    }
    asm volatile("griddepcontrol.launch_dependents;");         // see sslapb_launch_bid_sweep_redo
    // (rows of more than one warp pass folded in here by a loop over trips — instead of the redo list — spill at 40 registers
    // and cost 60 us: measured and dropped)
}

extern "C" cudaError_t sslapb_launch_bid_sweep_redo(const SslapbAuctionParams *P, const int *bidders, float eps, int merge,
                                                    int grid, cudaStream_t stream);
extern "C" cudaError_t sslapb_launch_bid_sweep_lean(const SslapbAuctionParams *P, const int *bidders, int nb, float eps,
                                                    int merge, int grid, cudaStream_t stream)
{
    // the redo counter (ctrl->hot_probe_fail) is zeroed by the caller before every launch pair
    sslapb_bid_sweep_lean_kernel<<<grid * 3, 512, 0, stream>>>(*P, bidders, nb, eps, merge);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return sslapb_launch_bid_sweep_redo(P, bidders, eps, merge, grid, stream);
}
