// Software-pipelined full-frontier bidding sweep for sm_100a: the roofline kernel of the CSR traversal.
//
// Same result per bidder, bit for bit, as sslapb_bid_sweep_kernel (auction.cu) and the oracle: the bidding loop of
// bid_and_assign (/root/reference/sslap/auction_.pyx:339-365) — top-2 of a_ij - p_j over the CSR row, last maximal entry
// wins, bid = a_ibest - w_i + eps — with the per-object atomicMax merge (:375-385) when asked.
//
// What differs from the round-1 kernel is the schedule, not the arithmetic.  ncu showed that kernel half latency bound
// and half issue bound: one row per warp at a time, a dependent chain offsets -> entries (HBM) -> prices (L2) -> five
// REDUX per row, issue slots 53 % busy, 32 warps of 64 registers per SM.  Here every warp keeps THREE rows in flight:
//   stage A  (two rows ahead)  bidder id, row offsets, row maximum            [3 small loads]
//   stage B  (one row ahead)   the row's 16-byte chunks: int4 columns + 2 x double2 values per lane  [HBM stream]
//   stage C  (current row)     bound-pruned price gathers, per-lane top-2, cross-lane top-2, bid, store / atomicMax
// so the HBM latency of row m+1 is covered by the arithmetic of row m inside the same warp instead of by other warps.
// Rows that do not fit one warp pass (> 32 chunks), rows whose pruned result is not proven exact and rows whose
// candidates are all at -inf take the exact generic sweep (row_bid<32>), as before.
#include "auction.cuh"
#include "rowsweep.cuh"


struct Sw2Row { long long st; int deg; double rmax; };

__device__ __forceinline__ Sw2Row sw2_load_row(const SslapbAuctionParams &P, const int *__restrict__ bidders, int a, int nb)
{
    Sw2Row r;
    r.st = 0; r.deg = 0; r.rmax = 0.0;
    if (a < nb) {
        const int i = bidders ? __ldg(bidders + a) : a;
        const long long st = __ldg(P.rowptr + i);
        r.st = st;
        r.deg = (int)(__ldg(P.rowptr + i + 1) - st);
        r.rmax = __ldg(P.rowmax + i);
    }
    return r;
}

__device__ __forceinline__ bool sw2_single(const Sw2Row &r) { return (((r.st + r.deg + 3) >> 2) - (r.st >> 2)) <= 32; }

template <int SW2_THREADS, bool LEAN>
__global__ void __launch_bounds__(SW2_THREADS, 1) sslapb_bid_sweep2_kernel(SslapbAuctionParams P, const int *__restrict__ bidders,
                                                                          int nb, float eps_f, int merge)
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
    int n2nd = 0;
    int a = gwarp;
    if (a >= nb) return;
    Sw2Row cur = sw2_load_row(P, bidders, a, nb);
    Sw2Row nxt = sw2_load_row(P, bidders, a + nwarps, nb);
    SslapbStreamChunk cc;
    cc.cj = make_int4(0, 0, 0, 0); cc.va = make_double2(0.0, 0.0); cc.vb = cc.va;
    if (sw2_single(cur)) cc = sslapb_stream_chunk(P.cols, P.vals, cur.st, cur.st + cur.deg, lane);
    for (;;) {
        const int an = a + nwarps;
        // stage B of the next row, stage A of the row after it: issued before this row's arithmetic
        SslapbStreamChunk cn;
        cn.cj = make_int4(0, 0, 0, 0); cn.va = make_double2(0.0, 0.0); cn.vb = cn.va;
        if (an < nb && sw2_single(nxt)) cn = sslapb_stream_chunk(P.cols, P.vals, nxt.st, nxt.st + nxt.deg, lane);
        const Sw2Row nn = sw2_load_row(P, bidders, an + nwarps, nb);
        // stage C
        const long long st = cur.st, en = cur.st + cur.deg;
        int j; double bid;
        if (sw2_single(cur)) {
            const SslapbBid o = LEAN ? row_bid_pruned_lean(cc, P.price, st, cur.deg, lane, eps, pmin, cur.rmax - spread, n2nd)
                                     : row_bid_pruned(cc, P.price, st, en, lane, eps, pmin, cur.rmax - spread, n2nd);
            j = o.j; bid = o.bid;
            if (j < 0) row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid);   // all candidates at -inf / unproven
        } else {
            row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid, pmin, cur.rmax - spread);
        }
        if (lane == 0) {
            P.bidj[a] = j;
            P.bidv[a] = bid;
            if (merge && j >= 0) atomicMax(P.bidkey + j, sslapb_ord64(bid));
        }
        if (an >= nb) break;
        a = an; cur = nxt; nxt = nn; cc = cn;
    }
    if (n2nd && lane == 0) atomicAdd((unsigned long long *)&P.ctrl->prune_second_pass, (unsigned long long)n2nd);
}

// threads: 768 (default: 24 warps of <= 85 registers per SM), or 1024 / 640 / 512 for A/B runs; lean: the trimmed stage C
extern "C" cudaError_t sslapb_launch_bid_sweep2(const SslapbAuctionParams *P, const int *bidders, int nb, float eps,
                                                int merge, int threads, int lean, int grid, cudaStream_t stream)
{
#define SW2_LAUNCH(T) do { if (lean) sslapb_bid_sweep2_kernel<T, true><<<grid, T, 0, stream>>>(*P, bidders, nb, eps, merge); \
                           else sslapb_bid_sweep2_kernel<T, false><<<grid, T, 0, stream>>>(*P, bidders, nb, eps, merge); } while (0)
    switch (threads) {
    case 1024: SW2_LAUNCH(1024); break;
    case 640: SW2_LAUNCH(640); break;
    case 512: SW2_LAUNCH(512); break;
    default: SW2_LAUNCH(768); break;
    }
#undef SW2_LAUNCH
    return cudaGetLastError();
}
