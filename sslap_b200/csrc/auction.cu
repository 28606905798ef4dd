// Persistent, device-resident Jacobi forward auction with eps-scaling for sm_100a.
//
// Replaces AuctionSolver.solve / bid_and_assign / eCE_satisfied / get_obj (/root/reference/sslap/auction_.pyx:268-523).
// One cooperative launch runs every round of every eps-phase; the host never synchronises per round.
//
// The trajectory is the reference's, bit for bit (same prices, `its`, `nreductions`, `sol`), ties included:
//   * within a row the LAST maximal entry wins and w_i is the second largest of the multiset (:346-358);
//   * between equal bids on one object the bidder EARLIEST in the unassigned list wins (strict '>' at :379) — so the
//     list is kept in exactly the reference's order: an evicted owner takes the winner's slot (:401-409), other
//     winners leave holes (:411-413), and push_all_left (:137-162) fills the k-th hole left of the new count with the
//     k-th live entry right of it.  All of that is order-independent per round, hence data-parallel.
//
// Three regimes, chosen by the frontier size nu (monotone non-increasing inside an eps-phase):
//   grid  (nu > t_small): all CTAs; warp per bidder; 64-bit atomicMax of the order-preserving bid per object (+ an
//                         atomicMin of the list position only in rounds where an equal bid was seen); 3 grid barriers.
//   warp  (4 < nu <= 32): CTA 0 only; warp w sweeps the row of list position w; warp 0 merges through shuffles.
//   solo  (nu <= 4)     : warp 0 of CTA 0 only; 32/16/8 lanes per bidder; no block barrier at all.
#include "auction.cuh"

#define SOLO_MAX 4

// ----------------------------------------------------------------------------------------------------------------------
// Row sweep: top-2 of (a_ij - p_j) over one CSR row by a group of W lanes (bidding loop, auction_.pyx:346-358).
// Lane t of the group owns the 16-byte-aligned chunks t, t+W, ... of 4 consecutive entries (one int4 of columns, two
// double2 of values); entries of the chunk outside [start,end) belong to neighbouring rows and are masked.
// Returns (in every lane of the group) the object and the bid (a_ibest - w_i + eps, :360); jbest = -1 for an empty row.
// ----------------------------------------------------------------------------------------------------------------------
#define SSLAPB_VISIT(M_, VAL_, PR_, COL_, IDX_)                                          \
    if (M_) {                                                                            \
        const double vi_ = (VAL_) - (PR_);                                               \
        if (vi_ >= b) { s = b; b = vi_; bc = (VAL_); bi = (IDX_); bj = (COL_); }         \
        else if (vi_ > s) s = vi_;                                                       \
    }

template <int W>
__device__ __forceinline__ void row_bid(const int *__restrict__ cols, const double *__restrict__ vals,
                                        const double *price, long long start, long long end, int t, double eps,
                                        int &jbest, double &bid)
{
    const int4 *c4 = reinterpret_cast<const int4 *>(cols);
    const double2 *v2 = reinterpret_cast<const double2 *>(vals);
    double b = SSLAPB_NEG_INF, s = SSLAPB_NEG_INF, bc = 0.0;
    int bi = -1, bj = -1;
    const long long c1 = (end + 3) >> 2;
#pragma unroll 2
    for (long long ch = (start >> 2) + t; ch < c1; ch += W) {
        const int4 cj = sslapb_ldg_stream_i4(c4 + ch);
        const double2 va = sslapb_ldg_stream_d2(v2 + 2 * ch);
        const double2 vb = sslapb_ldg_stream_d2(v2 + 2 * ch + 1);
        const long long e0 = ch << 2;
        const bool m0 = (e0 >= start) & (e0 < end);
        const bool m1 = (e0 + 1 >= start) & (e0 + 1 < end);
        const bool m2 = (e0 + 2 >= start) & (e0 + 2 < end);
        const bool m3 = (e0 + 3 >= start) & (e0 + 3 < end);
        const double p0 = m0 ? price[cj.x] : 0.0;
        const double p1 = m1 ? price[cj.y] : 0.0;
        const double p2 = m2 ? price[cj.z] : 0.0;
        const double p3 = m3 ? price[cj.w] : 0.0;
        const int i0 = (int)(e0 - start);
        SSLAPB_VISIT(m0, va.x, p0, cj.x, i0)
        SSLAPB_VISIT(m1, va.y, p1, cj.y, i0 + 1)
        SSLAPB_VISIT(m2, vb.x, p2, cj.z, i0 + 2)
        SSLAPB_VISIT(m3, vb.y, p3, cj.w, i0 + 3)
    }
    const int mine = bi;
#pragma unroll
    for (int off = W / 2; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(SSLAPB_FULL, b, off);
        const double os = __shfl_xor_sync(SSLAPB_FULL, s, off);
        const int oi = __shfl_xor_sync(SSLAPB_FULL, bi, off);
        const bool ow = (ob > b) || (ob == b && oi > bi);      // larger row index wins equal values (:351)
        s = ow ? fmax(os, b) : fmax(s, ob);
        b = ow ? ob : b;
        bi = ow ? oi : bi;
    }
    const unsigned gmask = (W == 32) ? SSLAPB_FULL : (((1u << (W & 31)) - 1u) << ((threadIdx.x & 31) & ~(W - 1)));
    const unsigned own = __ballot_sync(SSLAPB_FULL, mine >= 0 && mine == bi) & gmask;
    const int src = own ? (__ffs(own) - 1) : (threadIdx.x & 31);
    bc = __shfl_sync(SSLAPB_FULL, bc, src);
    bj = __shfl_sync(SSLAPB_FULL, bj, src);
    jbest = own ? bj : -1;
    bid = (bc - s) + eps;                                      // :360
}

// eCE / objective sweep of one row by a full warp (auction_.pyx:460-483 and :504-521).
//   vmax   = max_k (a_ik - p_k)
//   choice = value of the LAST entry whose column is jsel (:467-471)
//   csum   = sum of the values of ALL entries whose column is jsel (get_obj adds every match, :514-521)
__device__ __forceinline__ void row_ece(const int *__restrict__ cols, const double *__restrict__ vals,
                                        const double *price, long long start, long long end, int lane, int jsel,
                                        double &vmax, double &choice, double &csum)
{
    const int4 *c4 = reinterpret_cast<const int4 *>(cols);
    const double2 *v2 = reinterpret_cast<const double2 *>(vals);
    double vm = SSLAPB_NEG_INF, ch_v = 0.0, cs = 0.0;
    int ch_i = -1;
    const long long c1 = (end + 3) >> 2;
    for (long long ch = (start >> 2) + lane; ch < c1; ch += 32) {
        const int4 cj = sslapb_ldg_stream_i4(c4 + ch);
        const double2 va = sslapb_ldg_stream_d2(v2 + 2 * ch);
        const double2 vb = sslapb_ldg_stream_d2(v2 + 2 * ch + 1);
        const long long e0 = ch << 2;
        const int cc[4] = {cj.x, cj.y, cj.z, cj.w};
        const double vv[4] = {va.x, va.y, vb.x, vb.y};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long e = e0 + k;
            if (e >= start && e < end) {
                const double v = vv[k] - price[cc[k]];
                vm = fmax(vm, v);
                if (cc[k] == jsel) { ch_v = vv[k]; ch_i = (int)(e - start); cs += vv[k]; }
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        vm = fmax(vm, __shfl_xor_sync(SSLAPB_FULL, vm, off));
        const int oi = __shfl_xor_sync(SSLAPB_FULL, ch_i, off);
        const double ov = __shfl_xor_sync(SSLAPB_FULL, ch_v, off);
        if (oi > ch_i) { ch_i = oi; ch_v = ov; }
        cs += __shfl_xor_sync(SSLAPB_FULL, cs, off);
    }
    vmax = vm; choice = ch_v; csum = cs;
}

// ----------------------------------------------------------------------------------------------------------------------
// Grid barrier (sense by generation).  Returns false when the solve was aborted (watchdog / internal assert).
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool grid_barrier(SslapbCtrl *c, unsigned nblk, unsigned long long watchdog_ns)
{
    __shared__ int s_abort;
    __syncthreads();
    if (threadIdx.x == 0) {
        int ab = 0;
        const unsigned gen = *(volatile unsigned *)&c->bar_gen;
        __threadfence();
        const unsigned prev = atomicAdd(&c->bar_count, 1u);
        if (prev == nblk - 1) {
            *(volatile unsigned *)&c->bar_count = 0u;
            __threadfence();
            atomicAdd(&c->bar_gen, 1u);
        } else {
            unsigned polls = 0, ns = 64;
            unsigned long long t0 = 0;
            while (sslapb_ld_acquire_u32(&c->bar_gen) == gen) {
                if (++polls < 256) continue;                 // short waits (grid rounds): pure spin
                if (polls == 256) t0 = sslapb_globaltimer();
                if (sslapb_ld_volatile_s32(&c->abort_flag)) { ab = 1; break; }
                __nanosleep(ns);                             // long waits (CTA 0 runs the tail alone): back off
                if (ns < 2048) ns <<= 1;
                if (sslapb_globaltimer() - t0 > watchdog_ns) { *(volatile int *)&c->abort_flag = 1; ab = 1; break; }
            }
        }
        __threadfence();
        s_abort = ab | sslapb_ld_volatile_s32(&c->abort_flag);
    }
    __syncthreads();
    return s_abort == 0;
}

// ----------------------------------------------------------------------------------------------------------------------
// Warp-list regime: merge + assignment + list compaction for nu <= 32 bidders, executed by ONE warp.
// Lane a < nu holds list position a: person `li`, its object `j` and bid `bid`.  Returns the new count; `li` becomes
// the new list entry of position `lane` (-1 beyond the new count).  Restates auction_.pyx:375-430.
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_resolve(const SslapbAuctionParams &P, int nu, int &li, int j, double bid)
{
    const int lane = threadIdx.x & 31;
    const bool act = lane < nu;
    bool win = act && j >= 0;
    if (nu > 1) {
        for (int s = 0; s < nu; ++s) {
            const int oj = __shfl_sync(SSLAPB_FULL, j, s);
            const double ob = __shfl_sync(SSLAPB_FULL, bid, s);
            if (act && s != lane && oj == j && (ob > bid || (ob == bid && s < lane))) win = false;
        }
    }
    int nv = act ? li : -1;
    if (win) {
        const int prev = P.owner[j];
        P.price[j] = bid;                                      // :397
        P.owner[j] = li;                                       // :418
        P.p2o[li] = j;                                         // :417
        if (prev >= 0) P.p2o[prev] = -1;                       // :404
        nv = prev;                                             // evicted owner takes the slot (:409) or hole (:412)
    }
    __syncwarp();
    const unsigned valid = nu >= 32 ? SSLAPB_FULL : ((1u << nu) - 1u);
    const unsigned holes = __ballot_sync(SSLAPB_FULL, act && nv < 0);
    const int new_nu = nu - __popc(holes);                     // :429
    const unsigned leftm = new_nu >= 32 ? SSLAPB_FULL : ((1u << new_nu) - 1u);
    const unsigned left_holes = holes & leftm;
    const unsigned right_live = valid & ~holes & ~leftm;
    int src = lane;
    if ((left_holes >> lane) & 1u) {                           // push_all_left (:137-162)
        const int k = __popc(left_holes & ((1u << lane) - 1u));
        unsigned m = right_live;
        for (int q = 0; q < k; ++q) m &= m - 1u;               // drop the k lowest live entries
        src = __ffs(m) - 1;
    }
    const int v = __shfl_sync(SSLAPB_FULL, nv, src & 31);
    li = lane < new_nu ? v : -1;
    return new_nu;
}

// CTA 0 finishes the eps-phase alone once nu <= t_small (nu only shrinks inside a phase).
__device__ __noinline__ void small_regime(const SslapbAuctionParams &P, SslapbCtrl *C, int nu, float eps_f,
                                          long long its, long long max_iter)
{
    __shared__ int s_list[32];
    __shared__ int s_j[32];
    __shared__ double s_bid[32];
    __shared__ int s_nu, s_done;
    __shared__ long long s_its, s_rw, s_rs;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double eps = (double)eps_f;
    int li = -1, done = 0;
    long long rw = 0, rs = 0;
    if (warp == 0) {
        if (lane < nu) {
            const int v = P.list[lane];
            li = v < -1 ? P.mover[-(v + 2)] : v;               // grid regime leaves rank-encoded holes, see below
            s_list[lane] = li;
        }
    }
    __syncthreads();

    // ---- warp regime: one warp per bidder, 2 block barriers per round
    while (nu > SOLO_MAX && !done) {
        if (warp < nu) {
            const int i = s_list[warp];
            const long long st = __ldg(P.rowptr + i), en = __ldg(P.rowptr + i + 1);
            int j; double bid;
            row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid);
            if (lane == 0) { s_j[warp] = j; s_bid[warp] = bid; }
        }
        __syncthreads();
        if (warp == 0) {
            const int j = lane < nu ? s_j[lane] : -1;
            const double bid = lane < nu ? s_bid[lane] : 0.0;
            nu = warp_resolve(P, nu, li, j, bid);
            ++its; ++rw;
            if (its >= max_iter) done = 3;
            if (lane < 32) s_list[lane] = li;
            if (lane == 0) { s_nu = nu; s_done = done; s_its = its; }
        }
        __syncthreads();
        nu = s_nu; done = s_done; its = s_its;
    }

    // ---- solo regime: warp 0 alone, 32 / 16 / 8 lanes per bidder, no block barrier
    if (warp == 0) {
        while (nu > 0 && !done) {
            int j, pj; double bid, pb;
            if (nu == 1) {
                const int i = __shfl_sync(SSLAPB_FULL, li, 0);
                const long long st = __ldg(P.rowptr + i), en = __ldg(P.rowptr + i + 1);
                row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid);
                pj = j; pb = bid;
            } else if (nu == 2) {
                const int g = lane >> 4;
                const int i = __shfl_sync(SSLAPB_FULL, li, g);
                const long long st = __ldg(P.rowptr + i), en = __ldg(P.rowptr + i + 1);
                row_bid<16>(P.cols, P.vals, P.price, st, en, lane & 15, eps, j, bid);
                pj = __shfl_sync(SSLAPB_FULL, j, (lane << 4) & 31);
                pb = __shfl_sync(SSLAPB_FULL, bid, (lane << 4) & 31);
            } else {
                const int g = lane >> 3;
                const int i = __shfl_sync(SSLAPB_FULL, li, g);
                long long st = 0, en = 0;
                if (g < nu) { st = __ldg(P.rowptr + i); en = __ldg(P.rowptr + i + 1); }
                row_bid<8>(P.cols, P.vals, P.price, st, en, lane & 7, eps, j, bid);
                pj = __shfl_sync(SSLAPB_FULL, j, (lane << 3) & 31);
                pb = __shfl_sync(SSLAPB_FULL, bid, (lane << 3) & 31);
            }
            nu = warp_resolve(P, nu, li, pj, pb);
            ++its; ++rs;
            if (its >= max_iter) done = 3;
        }
        if (lane == 0) { s_nu = nu; s_done = done; s_its = its; s_rw = rw; s_rs = rs; }
        if (nu > 0 && lane < nu) P.list[lane] = li;            // only reachable through max_iter
    }
    __syncthreads();
    if (tid == 0) {
        C->nu = s_nu;
        C->its = s_its;
        if (s_done) C->done = s_done;
        C->rounds_warp += s_rw;
        C->rounds_solo += s_rs;
    }
}

// Block-wide exclusive prefix of a 0/1 flag over the threads of the CTA (in thread order); returns the CTA total in
// `total`.  Uses warp ballots + one pass over the warp totals.
__device__ __forceinline__ int block_excl_scan_flag(bool flag, int &total)
{
    __shared__ int s_wtot[32];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned bal = __ballot_sync(SSLAPB_FULL, flag);
    const int inwarp = __popc(bal & ((1u << lane) - 1u));
    __syncthreads();                                           // protect s_wtot from the previous call
    if (lane == 0) s_wtot[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        int v = lane < nw ? s_wtot[lane] : 0;
        int incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(SSLAPB_FULL, incl, off);
            if (lane >= off) incl += o;
        }
        s_wtot[lane] = incl - v;
        if (lane == 31) s_total = incl;
    }
    __syncthreads();
    total = s_total;
    return s_wtot[warp] + inwarp;
}

#define GB() do { if (!grid_barrier(C, nblk, P.watchdog_ns)) return; } while (0)

__global__ void __launch_bounds__(1024, 1) sslapb_auction_kernel(SslapbAuctionParams P)
{
    SslapbCtrl *C = P.ctrl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nblk = gridDim.x;
    const int wpc = blockDim.x >> 5;
    const int gwarp = blockIdx.x * wpc + warp;
    const int nwarps = nblk * wpc;
    const int gtid = blockIdx.x * blockDim.x + tid;
    const int nthreads = nblk * blockDim.x;
    __shared__ int s_red;
    __shared__ int s_hpre[3];

    if (gtid == 0) C->t_begin = sslapb_globaltimer();

    for (;;) {
        // ---- loop top: every CTA arrives here right after a grid barrier; the control block is stable
        int nu = *(volatile int *)&C->nu;
        int done = *(volatile int *)&C->done;
        const float eps_f = *(volatile float *)&C->eps;
        const long long its = *(volatile long long *)&C->its;
        const long long max_iter = *(volatile long long *)&C->max_iter;
        if (done) break;

        if (nu > P.t_small) {
            // ================================ grid regime: one round ================================
            const double eps = (double)eps_f;
            // (1) bidding: warp per list position (auction_.pyx:339-365) + per-object atomicMax merge (:375-385)
            for (int a = gwarp; a < nu; a += nwarps) {
                int v = P.list[a];
                if (v < -1) {                                  // hole filled by last round's compaction: k-th mover
                    v = P.mover[-(v + 2)];
                    if (lane == 0) P.list[a] = v;
                }
                const long long st = __ldg(P.rowptr + v), en = __ldg(P.rowptr + v + 1);
                int j; double bid;
                row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid);
                if (lane == 0) {
                    P.bidj[a] = j;
                    P.bidv[a] = bid;
                    if (j >= 0) {
                        const unsigned long long key = sslapb_ord64(bid);
                        const unsigned long long old = atomicMax(P.bidkey + j, key);
                        if (old == key) *(volatile int *)&C->tie_flag = 1;
                    } else {
                        *(volatile int *)&C->abort_flag = 2;   // empty row: rejected at CSR build, cannot happen
                    }
                }
            }
            GB();
            const int tie = *(volatile int *)&C->tie_flag;
            const int L = (nu + (int)nblk - 1) / (int)nblk;    // each CTA owns a contiguous chunk of positions
            const int lo = min(nu, (int)blockIdx.x * L), hi = min(nu, lo + L);
            if (tie) {                                         // (1b) equal best bids: earliest list position wins (:379)
                for (int a = lo + tid; a < hi; a += blockDim.x) {
                    const int j = P.bidj[a];
                    if (P.bidkey[j] == sslapb_ord64(P.bidv[a])) atomicMin(P.winpos + j, a);
                }
                GB();
            }
            // (2) assignment (:394-427), by the winner's own list position
            int myholes = 0;
            for (int a = lo + tid; a < hi; a += blockDim.x) {
                const int j = P.bidj[a];
                const double bid = P.bidv[a];
                const bool win = (P.bidkey[j] == sslapb_ord64(bid)) && (!tie || P.winpos[j] == a);
                if (win) {
                    const int i = P.list[a];
                    const int prev = P.owner[j];
                    P.price[j] = bid;
                    P.owner[j] = i;
                    P.p2o[i] = j;
                    if (prev >= 0) P.p2o[prev] = -1; else ++myholes;
                    P.list[a] = prev;                          // evicted owner takes the slot, or -1 = hole
                    P.bidkey[j] = 0ull;
                    if (tie) P.winpos[j] = 0x7fffffff;
                }
            }
            if (tid == 0) s_red = 0;
            __syncthreads();
            if (myholes) atomicAdd(&s_red, myholes);
            __syncthreads();
            if (tid == 0) P.hole_count[blockIdx.x] = s_red;
            GB();
            // (3) push_all_left (:137-162): k-th hole left of new_nu <- k-th live entry right of it
            if (warp == 0) {                                   // per-CTA hole counts -> total, my prefix, prefix of the split chunk
                int ht = 0, hp = 0;
                for (unsigned b = lane; b < nblk; b += 32) {
                    const int h = P.hole_count[b];
                    ht += h;
                    if (b < blockIdx.x) hp += h;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    ht += __shfl_xor_sync(SSLAPB_FULL, ht, off);
                    hp += __shfl_xor_sync(SSLAPB_FULL, hp, off);
                }
                const int cb0 = (nu - ht) / L;
                int h2 = 0;
                for (int b = lane; b < cb0; b += 32) h2 += P.hole_count[b];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) h2 += __shfl_xor_sync(SSLAPB_FULL, h2, off);
                if (lane == 0) { s_hpre[0] = h2; s_hpre[1] = hp; s_hpre[2] = ht; s_red = 0; }
            }
            __syncthreads();
            const int H = s_hpre[2], hpre = s_hpre[1];
            const int new_nu = nu - H;
            const int cb = new_nu / L;                         // chunk that contains the split point
            int cnt = 0;
            for (int a = cb * L + tid; a < new_nu; a += blockDim.x) cnt += (P.list[a] < 0);
            if (cnt) atomicAdd(&s_red, cnt);
            __syncthreads();
            const int Hsplit = s_hpre[0] + s_red;              // holes in [0,new_nu)
            __syncthreads();
            int run = hpre;                                    // holes in [0, tile start)
            for (int base = lo; base < hi; base += blockDim.x) {
                const int a = base + tid;
                int v = 0;
                bool hole = false;
                if (a < hi) { v = P.list[a]; hole = v < 0; }
                int ttot;
                const int before = run + block_excl_scan_flag(hole, ttot);
                if (a < hi) {
                    if (a < new_nu) {
                        if (hole) P.list[a] = -(before + 2);   // rank-encoded; decoded by the next reader
                    } else if (!hole) {
                        P.mover[(a - new_nu) - (before - Hsplit)] = v;
                    }
                }
                run += ttot;
            }
            if (gtid == 0) {
                C->nu = new_nu;
                C->its = its + 1;
                C->tie_flag = 0;
                C->rounds_grid += 1;
                if (its + 1 >= max_iter) C->done = 3;
            }
            GB();
        } else {
            // ================================ warp-list regimes: CTA 0 finishes the phase ================================
            if (blockIdx.x == 0) small_regime(P, C, nu, eps_f, its, max_iter);
            GB();
        }

        nu = *(volatile int *)&C->nu;
        done = *(volatile int *)&C->done;
        if (done || nu != 0) continue;

        // ================================ full assignment reached: terminate() / eps-scaling (:275-292) ================================
        {
            const float teps = *(volatile float *)&C->target_eps;
            const double eps_t = (double)teps;
            bool viol = false;
            for (int i = gwarp; i < P.N; i += nwarps) {        // eCE_satisfied(target_eps), :443-485
                const int j = P.p2o[i];
                const long long st = __ldg(P.rowptr + i), en = __ldg(P.rowptr + i + 1);
                double vmax, choice, csum;
                row_ece(P.cols, P.vals, P.price, st, en, lane, j, vmax, choice, csum);
                const double lhs = (choice - P.price[j]) + 1e-7;
                if (lhs < vmax - eps_t) viol = true;
            }
            if (viol && lane == 0) *(volatile int *)&C->ece_viol = 1;
            GB();
            const int v = *(volatile int *)&C->ece_viol;
            const float eps_now = *(volatile float *)&C->eps;
            const bool stop_opt = (v == 0);
            const bool stop_eps = !stop_opt && (eps_now < teps);   // :280
            if (!stop_opt && !stop_eps) {                      // :283-292 next phase: prices kept, everything else reset
                for (int i = gtid; i < P.N; i += nthreads) { P.p2o[i] = -1; P.list[i] = i; }
                for (int j = gtid; j < P.M; j += nthreads) P.owner[j] = -1;
            }
            GB();
            if (gtid == 0) {
                if (stop_opt) { C->done = 1; C->ece_final = 1; }
                else if (stop_eps) { C->done = 2; C->ece_final = 0; }
                else {
                    C->eps = eps_now * C->theta;               // float32 product (:283)
                    C->nreductions += 1;
                    C->nu = P.N;
                }
                C->ece_viol = 0;
            }
            GB();
        }
    }

    // ================================ epilogue: meta['eCE'] (:297) and per-person chosen values (get_obj) ================================
    {
        const int nu = *(volatile int *)&C->nu;
        const bool need_ece = (*(volatile int *)&C->ece_final < 0) && nu == 0;   // max_iter hit on a full assignment
        const double eps_t = (double)(*(volatile float *)&C->target_eps);
        bool viol = false;
        for (int i = gwarp; i < P.N; i += nwarps) {
            const int j = P.p2o[i];
            double csum = 0.0;
            if (j >= 0) {
                const long long st = __ldg(P.rowptr + i), en = __ldg(P.rowptr + i + 1);
                double vmax, choice;
                row_ece(P.cols, P.vals, P.price, st, en, lane, j, vmax, choice, csum);
                if (need_ece && ((choice - P.price[j]) + 1e-7 < vmax - eps_t)) viol = true;
            }
            if (lane == 0) P.chosen[i] = csum;
        }
        if (viol && lane == 0) *(volatile int *)&C->ece_viol = 1;
        GB();
        if (gtid == 0) {
            if (*(volatile int *)&C->ece_final < 0) C->ece_final = (nu == 0 && C->ece_viol == 0) ? 1 : 0;
            C->t_end = sslapb_globaltimer();
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// Stand-alone bidding sweep (non-cooperative): the grid regime's step (1) for an explicit bidder list.  Used for
// kernel-level parity (bit-exact (jbest, bid) against the oracle) and for the HBM-roofline measurement of the CSR sweep.
// ----------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) sslapb_bid_sweep_kernel(SslapbAuctionParams P, const int *bidders, int nb,
                                                                  float eps_f, int merge)
{
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int gwarp = blockIdx.x * wpc + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * wpc;
    const double eps = (double)eps_f;
    for (int a = gwarp; a < nb; a += nwarps) {
        const int i = bidders ? bidders[a] : a;
        const long long st = __ldg(P.rowptr + i), en = __ldg(P.rowptr + i + 1);
        int j; double bid;
        row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid);
        if (lane == 0) {
            P.bidj[a] = j;
            P.bidv[a] = bid;
            if (merge && j >= 0) atomicMax(P.bidkey + j, sslapb_ord64(bid));
        }
    }
}

// Initial state of a solve (AuctionSolver.__init__, auction_.pyx:220-261).
__global__ void sslapb_auction_init_kernel(SslapbAuctionParams P)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
    for (int i = gtid; i < P.N; i += n) { P.p2o[i] = -1; P.list[i] = i; }
    for (int j = gtid; j < P.M; j += n) { P.owner[j] = -1; P.price[j] = 0.0; P.bidkey[j] = 0ull; P.winpos[j] = 0x7fffffff; }
}

extern "C" cudaError_t sslapb_launch_auction(const SslapbAuctionParams *P, int grid, cudaStream_t stream)
{
    sslapb_auction_init_kernel<<<grid, 1024, 0, stream>>>(*P);
    void *args[] = {(void *)P};
    return cudaLaunchCooperativeKernel((const void *)sslapb_auction_kernel, dim3(grid), dim3(1024), args, 0, stream);
}

extern "C" cudaError_t sslapb_auction_grid_size(int device, int *grid)
{
    int per_sm = 0, sms = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sslapb_auction_kernel, 1024, 0);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid = sms;                                               // one persistent CTA per SM
    return cudaSuccess;
}

extern "C" cudaError_t sslapb_launch_bid_sweep(const SslapbAuctionParams *P, const int *bidders, int nb, float eps,
                                               int merge, int grid, cudaStream_t stream)
{
    sslapb_bid_sweep_kernel<<<grid, 1024, 0, stream>>>(*P, bidders, nb, eps, merge);
    return cudaGetLastError();
}
