// Batched small problems, round 2 (BASELINE.json configs[4]: thousands of independent 512 x 512 problems): ONE WARP PER
// PROBLEM as before (every auction is a chain of dependent rounds; throughput = chains in flight), but the warp no longer
// sweeps one bidder at a time with 32 lanes:
//   * sub-warp bidding (north star: "one warp or sub-warp per unassigned person, chosen by row degree"): rows of up to
//     16 aligned chunks (>= 61 entries; C5 rows have ~26) are swept by 8 lanes, FOUR bidders per pass, with the price and
//     the current owner of every candidate object gathered in one 16-byte record and a 3-step shuffle butterfly inside
//     the group instead of 5 REDUX over the warp (reference loop: auction_.pyx:337-365);
//   * once the frontier fits the warp (nu <= 32: the large majority of all rounds) the unassigned list lives in
//     registers — lane n = list position n — and merge (:375-385), assignment (:394-427) and push_all_left (:137-162)
//     are done with match_any / ballots, without the global atomics and the six dependent L2 round trips of the
//     wide-frontier path.  A round then costs two memory round trips (row entries -> object records) per pass.
// The trajectory of every problem (sol, its, nreductions, prices, meta) is the reference's, bit for bit, exactly as in
// the round-1 kernel (auction.cu, kept behind option "batch_v1" for A/B runs).
#include "auction.cuh"
#include "rowsweep.cuh"
#include "batch.cuh"

#define B2_NCH 2                          // chunks per lane of an 8-lane group: rows of up to 8 * B2_NCH aligned chunks

struct B2Chunks { int4 cj[B2_NCH]; double va[B2_NCH], vb[B2_NCH], vc[B2_NCH], vd[B2_NCH]; };

__device__ __forceinline__ B2Chunks b2_load(const int *__restrict__ cols, const double *__restrict__ vals, long long st,
                                            int deg, int t)
{
    B2Chunks C;
    const long long c0 = st >> 2, c1 = (st + deg + 3) >> 2;
#pragma unroll
    for (int k = 0; k < B2_NCH; ++k) {
        C.cj[k] = make_int4(0, 0, 0, 0); C.va[k] = C.vb[k] = C.vc[k] = C.vd[k] = 0.0;
        const long long ch = c0 + t + 8 * k;
        if (deg > 0 && ch < c1) {
            C.cj[k] = __ldg(reinterpret_cast<const int4 *>(cols) + ch);
            asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(C.va[k]), "=d"(C.vb[k]), "=d"(C.vc[k]), "=d"(C.vd[k]) : "l"(vals + 4 * ch));
        }
    }
    return C;
}

struct B2Rec { double price; int owner; };
__device__ __forceinline__ B2Rec b2_ld_rec(const SslapbBatchRec *p, bool on)
{
    B2Rec r;
    r.price = 0.0; r.owner = -1;
    if (on) {
        unsigned long long pb, ob;
        asm volatile("ld.global.v2.b64 {%0,%1}, [%2];" : "=l"(pb), "=l"(ob) : "l"(p) : "memory");
        r.price = __longlong_as_double((long long)pb);
        r.owner = (int)(unsigned)ob;
    }
    return r;
}
__device__ __forceinline__ void b2_st_rec(SslapbBatchRec *p, double price, int owner)
{
    asm volatile("st.global.v2.b64 [%0], {%1,%2};" :: "l"(p), "l"((unsigned long long)__double_as_longlong(price)),
                 "l"((unsigned long long)(unsigned)owner) : "memory");
}

struct B2Bid { int j; double bid; int powner; };

// One row per 8-lane group (lane t of the group owns chunks t, t + 8, ...): top-2 of a_ij - p_j with the reference's tie
// rule (last maximal entry, :351), bid = a_ibest - w_i + eps (:360), plus the current owner of the object bid on.
// j = -1: the row has no candidate above -inf (empty group, or every object priced +inf) — the caller decides it exactly.
__device__ __forceinline__ B2Bid b2_group_bid(const B2Chunks &C, const SslapbBatchRec *brec, long long st, int deg, int t,
                                              int lane, double eps)
{
    const int off = 4 * t - (int)(st & 3);
    double b = SSLAPB_NEG_INF, s = SSLAPB_NEG_INF, bc = 0.0;
    int bi = -1, bj = -1, bo = -1;
#pragma unroll
    for (int k = 0; k < B2_NCH; ++k) {
        const int o = off + 32 * k;
        const bool m0 = (unsigned)o < (unsigned)deg, m1 = (unsigned)(o + 1) < (unsigned)deg;
        const bool m2 = (unsigned)(o + 2) < (unsigned)deg, m3 = (unsigned)(o + 3) < (unsigned)deg;
        const B2Rec r0 = b2_ld_rec(brec + C.cj[k].x, m0), r1 = b2_ld_rec(brec + C.cj[k].y, m1);
        const B2Rec r2 = b2_ld_rec(brec + C.cj[k].z, m2), r3 = b2_ld_rec(brec + C.cj[k].w, m3);
        const double v0 = m0 ? C.va[k] - r0.price : SSLAPB_NEG_INF, v1 = m1 ? C.vb[k] - r1.price : SSLAPB_NEG_INF;
        const double v2 = m2 ? C.vc[k] - r2.price : SSLAPB_NEG_INF, v3 = m3 ? C.vd[k] - r3.price : SSLAPB_NEG_INF;
        const SslapbLaneTop lt = sslapb_lane_top2(v0, v1, v2, v3);
        if (lt.b >= b) {                                       // later chunk: wins equal values
            s = b > lt.s ? b : lt.s;
            b = lt.b;
            bi = o + lt.w;
            bc = (lt.w & 2) ? ((lt.w & 1) ? C.vd[k] : C.vc[k]) : ((lt.w & 1) ? C.vb[k] : C.va[k]);
            bj = (lt.w & 2) ? ((lt.w & 1) ? C.cj[k].w : C.cj[k].z) : ((lt.w & 1) ? C.cj[k].y : C.cj[k].x);
            bo = (lt.w & 2) ? ((lt.w & 1) ? r3.owner : r2.owner) : ((lt.w & 1) ? r1.owner : r0.owner);
        } else {
            s = lt.b > s ? lt.b : s;
        }
    }
    const bool has = b > SSLAPB_NEG_INF;
    const int mybi = has ? bi : -1;
    const unsigned long long bk = has ? sslapb_ord64(b + 0.0) : 0ull;      // + 0.0 folds -0.0 into +0.0
    const unsigned long long sk = has ? sslapb_ord64(s + 0.0) : 0ull;
    unsigned long long mk = bk;
    int mi = mybi;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(SSLAPB_FULL, mk, o);
        const int oi = __shfl_xor_sync(SSLAPB_FULL, mi, o);
        const bool take = (ok > mk) || (ok == mk && oi > mi);
        mk = take ? ok : mk;
        mi = take ? oi : mi;
    }
    const bool iswin = has && bk == mk && mybi == mi;
    unsigned long long ck = iswin ? sk : bk;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(SSLAPB_FULL, ck, o);
        ck = ok > ck ? ok : ck;
    }
    const unsigned own = (__ballot_sync(SSLAPB_FULL, iswin) >> (lane & 24)) & 0xffu;
    const int src = own ? ((lane & 24) + __ffs(own) - 1) : lane;
    const double wbc = __shfl_sync(SSLAPB_FULL, bc, src);
    const int wbj = __shfl_sync(SSLAPB_FULL, bj, src);
    const int wbo = __shfl_sync(SSLAPB_FULL, bo, src);
    const double wi = ck > SSLAPB_KEY_NEG_INF ? sslapb_key2double(ck) : SSLAPB_NEG_INF;   // :344
    B2Bid R;
    R.j = own ? wbj : -1;
    R.bid = (wbc - wi) + eps;                                  // :360
    R.powner = wbo;
    return R;
}

__device__ __forceinline__ bool b2_fits(long long st, int deg) { return (((st + deg + 3) >> 2) - (st >> 2)) <= 8 * B2_NCH; }

// Bids of the four persons (pst[g], pdg[g] known to every lane of group g; dg = 0: no person) of one pass.  Rows that do
// not fit a group, and rows without a candidate above -inf, are decided by the exact whole-warp sweep on price[].
__device__ __forceinline__ B2Bid b2_pass(const SslapbBatchParams &B, long long st, int deg, int lane, double eps)
{
    const int t = lane & 7;
    const bool fits = b2_fits(st, deg);
    const B2Chunks C = b2_load(B.cols, B.vals, st, fits ? deg : 0, t);
    B2Bid R = b2_group_bid(C, B.brec, st, fits ? deg : 0, t, lane, eps);
    unsigned redo = __ballot_sync(SSLAPB_FULL, deg > 0 && t == 0 && R.j < 0);
    while (redo) {
        const int gl = __ffs(redo) - 1;
        redo &= redo - 1;
        const long long rst = __shfl_sync(SSLAPB_FULL, st, gl);
        const int rdg = __shfl_sync(SSLAPB_FULL, deg, gl);
        int rj; double rbid;
        row_bid<32>(B.cols, B.vals, B.price, rst, rst + rdg, lane, eps, rj, rbid);
        const int ro = rj >= 0 ? B.brec[rj].owner : -1;
        if ((lane >> 3) == (gl >> 3)) { R.j = rj; R.bid = rbid; R.powner = ro; }
    }
    return R;
}

__device__ __forceinline__ bool b2_ece(const SslapbBatchParams &B, long long r0, int N, int lane, float teps)
{
    const double eps_t = (double)teps;
    bool viol = false;
    for (int i = 0; i < N; ++i) {
        const long long g = r0 + i;
        const int j = B.p2o[g];
        double vmax, choice, csum;
        row_ece(B.cols, B.vals, B.price, __ldg(B.rowptr + g), __ldg(B.rowptr + g + 1), lane, j, vmax, choice, csum);
        if (((choice - B.price[j]) + 1e-7) < vmax - eps_t) viol = true;
    }
    return !__any_sync(SSLAPB_FULL, viol);
}

__global__ void __launch_bounds__(128, 7) sslapb_auction_batch2_kernel(SslapbBatchParams B)
{
    const int lane = threadIdx.x & 31, g = lane >> 3;
    const int p = __shfl_sync(SSLAPB_FULL, (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), 0);
    if (p >= B.P) return;
    const long long r0 = B.rowoff[p], c0 = B.coloff[p];
    const int N = (int)(B.rowoff[p + 1] - r0), M = (int)(B.coloff[p + 1] - c0);
    // ---- AuctionSolver.__init__ (auction_.pyx:220-261)
    for (int j = lane; j < M; j += 32) {
        b2_st_rec(B.brec + c0 + j, 0.0, -1);
        B.price[c0 + j] = 0.0; B.bestkey[c0 + j] = 0ull; B.winpos[c0 + j] = 0x7fffffff;
    }
    for (int i = lane; i < N; i += 32) { B.p2o[r0 + i] = -1; B.list[r0 + i] = (int)(r0 + i); }
    double cmax = 0.0;                                         // max_val (:123-134)
    for (long long e = __ldg(B.rowptr + r0) + lane; e < __ldg(B.rowptr + r0 + N); e += 32) cmax = fmax(cmax, fabs(B.vals[e]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cmax = fmax(cmax, __shfl_xor_sync(SSLAPB_FULL, cmax, off));
    float eps = (float)((double)((float)cmax) / 2.0);          // :242-246
    const float target = (float)(1.0 / (double)N), theta = 0.15f;
    if (B.eps_start && B.eps_start[p] > 0) eps = B.eps_start[p];
    const float start_eps = eps;
    __syncwarp();

    int nu = N, stop = 0, nred = 0, last_opt = -1;
    long long its = 0;
    // register-resident list entry of position `lane` (valid while inreg)
    bool inreg = false;
    int me = -1, mdg = 0;
    long long mst = 0;
    for (;;) {
        const double epsd = (double)eps;
        if (nu <= 32) {
            // ================= the frontier fits the warp: list in registers, merge / assignment / compaction by ballots =================
            if (!inreg) {
                me = -1; mst = 0; mdg = 0;
                if (lane < nu) {
                    me = B.list[r0 + lane];
                    mst = __ldg(B.rowptr + me);
                    mdg = (int)(__ldg(B.rowptr + me + 1) - mst);
                }
                inreg = true;
            }
            int bj = -1, bo = -1;
            double bv = 0.0;
            for (int base = 0; base < nu; base += 4) {         // bidding (:339-365), four bidders per pass
                const int src = base + g;                      // list position swept by my group (< 32)
                const long long st = __shfl_sync(SSLAPB_FULL, mst, src);
                const int dg = __shfl_sync(SSLAPB_FULL, mdg, src);
                const B2Bid R = b2_pass(B, st, src < nu ? dg : 0, lane, epsd);
                const int from = ((lane - base) & 3) * 8;      // lane base + q receives the result of group q
                const int tj = __shfl_sync(SSLAPB_FULL, R.j, from);
                const double tb = __shfl_sync(SSLAPB_FULL, R.bid, from);
                const int to = __shfl_sync(SSLAPB_FULL, R.powner, from);
                if (lane >= base && lane < base + 4) { bj = tj; bv = tb; bo = to; }
            }
            const bool act = lane < nu;
            bool win = act && bj >= 0;
            // the evicted owner's row is requested now; the merge below overlaps the load
            long long pst = 0;
            int pdg = 0;
            if (win && bo >= 0) { pst = __ldg(B.rowptr + bo); pdg = (int)(__ldg(B.rowptr + bo + 1) - pst); }
            // merge (:375-385): only bidders on the same object have anything to settle
            const unsigned peers = __match_any_sync(SSLAPB_FULL, win ? bj : (-1 - lane));
            if (__any_sync(SSLAPB_FULL, peers != (1u << lane))) {
                for (int s = 0; s < nu; ++s) {
                    const int oj = __shfl_sync(SSLAPB_FULL, bj, s);
                    const double ob = __shfl_sync(SSLAPB_FULL, bv, s);
                    // strict '>' at :379: the earliest bidder in list order keeps an equal bid
                    if (act && s != lane && oj == bj && (ob > bv || (ob == bv && s < lane))) win = false;
                }
            }
            int nv = act ? me : -1;
            long long nst = mst;
            int ndg = mdg;
            if (win) {                                         // assignment (:394-427)
                b2_st_rec(B.brec + bj, bv, me);
                B.price[bj] = bv;
                B.p2o[me] = bj;
                if (bo >= 0) B.p2o[bo] = -1;
                nv = bo; nst = pst; ndg = pdg;                 // evicted owner takes the slot (:409) or hole (:412)
            }
            __syncwarp();
            // push_all_left (:137-162): the k-th hole left of the new count takes the k-th live entry right of it
            const unsigned valid = nu >= 32 ? SSLAPB_FULL : ((1u << nu) - 1u);
            const unsigned holes = __ballot_sync(SSLAPB_FULL, act && nv < 0);
            const int new_nu = nu - __popc(holes);
            if (holes) {
                const unsigned leftm = new_nu >= 32 ? SSLAPB_FULL : ((1u << new_nu) - 1u);
                const unsigned left_holes = holes & leftm, right_live = valid & ~holes & ~leftm;
                int src = lane;
                if ((left_holes >> lane) & 1u) {
                    const int k = __popc(left_holes & ((1u << lane) - 1u));
                    unsigned m = right_live;
                    for (int q = 0; q < k; ++q) m &= m - 1u;
                    src = __ffs(m) - 1;
                }
                const int v = __shfl_sync(SSLAPB_FULL, nv, src & 31);
                const long long vs = __shfl_sync(SSLAPB_FULL, nst, src & 31);
                const int vd = __shfl_sync(SSLAPB_FULL, ndg, src & 31);
                me = lane < new_nu ? v : -1; mst = vs; mdg = vd;
            } else {
                me = nv; mst = nst; mdg = ndg;
            }
            nu = new_nu;
        } else {
            // ================= wide frontier: list in global memory (auction_.pyx:313-430 literally), four bidders per pass =================
            inreg = false;
            for (int base = 0; base < nu; base += 4) {
                const int n = base + g;
                long long st = 0;
                int dg = 0;
                if (n < nu) {
                    const int i = B.list[r0 + n];
                    st = __ldg(B.rowptr + i);
                    dg = (int)(__ldg(B.rowptr + i + 1) - st);
                }
                const B2Bid R = b2_pass(B, st, dg, lane, epsd);
                if (n < nu && (lane & 7) == 0) { B.bidj[r0 + n] = R.j; B.bidv[r0 + n] = R.bid; }
            }
            __syncwarp();
            // merge (:375-385): per-object maximum of the order-preserving bid, earliest list position on equal bids
            bool tie = false;
            for (int base = 0; base < nu; base += 32) {
                const int n = base + lane;
                if (n < nu) {
                    const int j = B.bidj[r0 + n];
                    if (j >= 0) {
                        const unsigned long long key = sslapb_ord64(B.bidv[r0 + n]);
                        if (atomicMax(B.bestkey + j, key) == key) tie = true;
                    }
                }
            }
            tie = __any_sync(SSLAPB_FULL, tie);
            if (tie) {
                for (int base = 0; base < nu; base += 32) {
                    const int n = base + lane;
                    if (n < nu) {
                        const int j = B.bidj[r0 + n];
                        if (j >= 0 && __ldcg(B.bestkey + j) == sslapb_ord64(B.bidv[r0 + n])) atomicMin(B.winpos + j, n);
                    }
                }
            }
            __syncwarp();
            // assignment (:394-427)
            int holes = 0;
            for (int base = 0; base < nu; base += 32) {
                const int n = base + lane;
                bool hole = false;
                if (n < nu) {
                    const int j = B.bidj[r0 + n];
                    const double bid = B.bidv[r0 + n];
                    const bool win = j >= 0 && (__ldcg(B.bestkey + j) == sslapb_ord64(bid)) && (!tie || __ldcg(B.winpos + j) == n);
                    if (win) {
                        const int i = B.list[r0 + n];
                        const int prev = B.brec[j].owner;
                        b2_st_rec(B.brec + j, bid, i);
                        B.price[j] = bid;
                        B.p2o[i] = j;
                        if (prev >= 0) B.p2o[prev] = -1;
                        B.list[r0 + n] = prev;
                        hole = prev < 0;
                    }
                }
                holes += __popc(__ballot_sync(SSLAPB_FULL, hole));
            }
            // reset best-bid slots of every object that received a bid (auction_.pyx:421-422)
            for (int base = 0; base < nu; base += 32) {
                const int n = base + lane;
                if (n < nu) {
                    const int j = B.bidj[r0 + n];
                    if (j >= 0) { B.bestkey[j] = 0ull; B.winpos[j] = 0x7fffffff; }
                }
            }
            __syncwarp();
            // push_all_left (:137-162)
            const int new_nu = nu - holes;
            if (holes && new_nu > 0) {
                int k = 0;
                for (int base = new_nu & ~31; base < nu; base += 32) {
                    const int n = base + lane;
                    const bool live = n >= new_nu && n < nu && B.list[r0 + n] >= 0;
                    const unsigned bal = __ballot_sync(SSLAPB_FULL, live);
                    if (live) B.mover[r0 + k + __popc(bal & ((1u << lane) - 1u))] = B.list[r0 + n];
                    k += __popc(bal);
                }
                __syncwarp();
                int q = 0;
                for (int base = 0; base < new_nu && q < k; base += 32) {
                    const int n = base + lane;
                    const bool hole = n < new_nu && B.list[r0 + n] < 0;
                    const unsigned bal = __ballot_sync(SSLAPB_FULL, hole);
                    if (hole) B.list[r0 + n] = B.mover[r0 + q + __popc(bal & ((1u << lane) - 1u))];
                    q += __popc(bal);
                }
                __syncwarp();
            }
            nu = new_nu;
        }
        ++its;
        last_opt = -1;
        // ---- terminate() / eps-scaling (:275-292)
        if (its >= B.max_iter) { stop = 3; break; }
        if (nu == 0) {
            __syncwarp();
            last_opt = b2_ece(B, r0, N, lane, target) ? 1 : 0;
            if (last_opt) { stop = 1; break; }
            if (eps < target) { stop = 2; break; }
            eps = eps * theta;
            for (int j = lane; j < M; j += 32) B.brec[c0 + j].owner = -1;
            for (int i = lane; i < N; i += 32) { B.p2o[r0 + i] = -1; B.list[r0 + i] = (int)(r0 + i); }
            __syncwarp();
            nu = N;
            inreg = false;
            ++nred;
        }
    }
    __syncwarp();
    if (last_opt < 0) last_opt = (nu == 0 && b2_ece(B, r0, N, lane, target)) ? 1 : 0;
    // ---- get_obj (:489-523): per-person chosen values (summed on the host in row order) and local column ids
    for (int i = 0; i < N; ++i) {
        const long long gi = r0 + i;
        const int j = B.p2o[gi];
        double csum = 0.0;
        if (j >= 0) {
            double vmax, choice;
            row_ece(B.cols, B.vals, B.price, __ldg(B.rowptr + gi), __ldg(B.rowptr + gi + 1), lane, j, vmax, choice, csum);
        }
        if (lane == 0) { B.chosen[gi] = csum; B.p2o[gi] = j >= 0 ? (int)(j - c0) : -1; }
    }
    if (lane == 0) {
        SslapbBatchMeta m;
        m.start_eps = start_eps; m.final_eps = eps; m.target_eps = target;
        m.eCE = last_opt; m.soln_found = (nu == 0 && last_opt) ? 1 : 0; m.stop_reason = stop;
        m.its = its; m.nreductions = nred; m.n_assigned = N - nu;
        B.meta[p] = m;
    }
}

extern "C" cudaError_t sslapb_launch_auction_batch2(const SslapbBatchParams *B, cudaStream_t stream)
{
    const int warps_per_cta = 4;
    const int grid = (B->P + warps_per_cta - 1) / warps_per_cta;
    sslapb_auction_batch2_kernel<<<grid, 32 * warps_per_cta, 0, stream>>>(*B);
    return cudaGetLastError();
}
