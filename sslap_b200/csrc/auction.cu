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
// Regimes, chosen by the frontier size nu (monotone non-increasing inside an eps-phase):
//   grid      (nu > t_small = 32): all CTAs; warp per bidder; 64-bit atomicMax of the order-preserving bid per object
//                                  (+ an atomicMin of the list position only in rounds where an equal bid was seen);
//                                  3 grid barriers (spread_round).
//   33..t_mid CTA 0 only, in eps-phases whose bids the hot lists decide (t_mid = 128): positions strided over its 16 warps
//             with pipelined record gathers, merge by all 512 threads, block barriers instead of grid barriers (mid_regime).
//   17..32    CTA 0 only; list positions strided over its 16 warps, warp 0 merges through shuffles (warp_resolve).
//   3..16     CTA 0 only; warp a owns position a; ONE named barrier per round, outcome derived redundantly in registers
//             (multi_rounds; the instance of auction_long.cu keeps a two-barrier form with a whole-CTA sweep of very
//             long rows).
//   2         warps 0 and 1 of CTA 0 (duo rounds inside multi_rounds).
//   1         warp 0 of CTA 0, no barrier at all (chain_rounds; coop_chain_rounds when the row is long).
// Round 2, on top of the regimes:
//   hot lists     (hot.cu) every regime first tries to decide a bid from the person's 32 largest entries (512 bytes, one
//                 gather per lane) and falls back to the full row when that is not provably exact; switched on and off
//                 per eps-phase by the measured share of undecided bids (row_bid_hot, sweep_hot, *_rounds_hot);
//   no-hole exit  a grid round in which no unowned object was won and no equal bids were seen ends after its second
//                 barrier (no compaction, no final barrier); frontiers of at most 512 are compacted by CTA 0 alone;
//   lean loops    the single-warp loops are bound by instruction latency, not memory: the chain loop of
//                 the decided rounds carries no fallback code and is unrolled by two;
//   person_to_object is rebuilt from the records' owners when it is needed (rebuild_p2o_owners), not kept per round.
// This file is compiled three times (plain, SSLAPB_LONG_ROWS, SSLAPB_SHARDED): see the two wrapper units.  (Round 1's
// opt-in cluster regime — mid-sized frontiers on one 8-CTA cluster — was removed in round 2: DESIGN.md 4.1b.)
#include "auction.cuh"

#define SSLAPB_THREADS 512       // persistent kernel: one CTA of 16 warps per SM (128 registers per thread)

#include "rowsweep.cuh"


// ----------------------------------------------------------------------------------------------------------------------
// Grid barrier: ONE release-add per CTA on a monotonically increasing counter, then acquire-polling of the same word
// until it reaches `target` (= barriers passed so far * #CTAs; every CTA counts its barriers in a register, so there is
// no reset and no generation word).  Returns false when the solve was aborted (watchdog / internal assert).
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool grid_barrier(SslapbCtrl *c, unsigned nblk, unsigned &epoch, unsigned long long watchdog_ns)
{
    __shared__ int s_abort;
    epoch += nblk;
    __syncthreads();
    if (threadIdx.x == 0) {
        int ab = 0;
        const unsigned target = epoch;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(&c->bar_count) : "memory");
        unsigned polls = 0, ns = 64;
        unsigned long long t0 = 0;
        while ((int)(sslapb_ld_acquire_u32(&c->bar_count) - target) < 0) {
            if (++polls < 256) continue;                     // short waits (grid rounds): pure spin
            if (polls == 256) t0 = sslapb_globaltimer();
            if (sslapb_ld_volatile_s32(&c->abort_flag)) { ab = 1; break; }
            __nanosleep(ns);                                 // long waits (CTA 0 runs the tail alone): back off
            if (ns < 2048) ns <<= 1;
            if (sslapb_globaltimer() - t0 > watchdog_ns) { *(volatile int *)&c->abort_flag = 1; ab = 1; break; }
        }
        s_abort = ab | sslapb_ld_volatile_s32(&c->abort_flag);
    }
    __syncthreads();
    return s_abort == 0;
}

// ----------------------------------------------------------------------------------------------------------------------
// Warp-list regime: merge + assignment + list compaction for nu <= 32 bidders, executed by ONE warp.
// Lane a < nu holds list position a: person `li` with CSR row [lst, lst+ldg), and its bid `B` (object, value, record of
// the object's current owner).  Returns the new count; (li,lst,ldg) become the new list entry of position `lane`.
// Restates auction_.pyx:375-430.
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_resolve(const SslapbAuctionParams &P, int nu, int &li, long long &lst, int &ldg,
                                            const SslapbBid &B)
{
    const int lane = threadIdx.x & 31;
    const bool act = lane < nu;
    const int j = B.j;
    const double bid = B.bid;
    bool win = act && j >= 0;
    // bidders on the same object: only then is there anything to merge (rare for small frontiers)
    const unsigned peers = __match_any_sync(SSLAPB_FULL, win ? j : (-1 - lane));
    if (__any_sync(SSLAPB_FULL, peers != (1u << lane))) {
        for (int s = 0; s < nu; ++s) {
            const int oj = __shfl_sync(SSLAPB_FULL, j, s);
            const double ob = __shfl_sync(SSLAPB_FULL, bid, s);
            // strict '>' at :379: the earliest bidder in list order keeps an equal bid
            if (act && s != lane && oj == j && (ob > bid || (ob == bid && s < lane))) win = false;
        }
    }
    int nv = act ? li : -1;
    long long nst = lst;
    int ndg = ldg;
    if (win) {
        sslapb_st_rec256(P.rec + j, (unsigned long long)lst, ((unsigned long long)(unsigned)ldg << 32) | (unsigned)li,
                         (unsigned long long)__double_as_longlong(bid));   // owner + its row (:418), price (:397)
        P.price[j] = bid;
        // (person_to_object, :404 / :417, is not kept up to date round by round: rebuild_p2o derives it from the records)
        nv = B.powner;                                         // evicted owner takes the slot (:409) or hole (:412)
        nst = B.pstart; ndg = B.pdeg;
    }
    __syncwarp();
    const unsigned valid = nu >= 32 ? SSLAPB_FULL : ((1u << nu) - 1u);
    const unsigned holes = __ballot_sync(SSLAPB_FULL, act && nv < 0);
    const int new_nu = nu - __popc(holes);                     // :429
    if (holes == 0u) { li = nv; lst = nst; ldg = ndg; return new_nu; }
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
    const long long vs = __shfl_sync(SSLAPB_FULL, nst, src & 31);
    const int vd = __shfl_sync(SSLAPB_FULL, ndg, src & 31);
    li = lane < new_nu ? v : -1;
    lst = vs; ldg = vd;
    return new_nu;
}

// Loads of one 4-entry chunk of a row (hop 1 of a round).
struct SslapbChunk { int4 cj; double2 va, vb; };
__device__ __forceinline__ SslapbChunk sslapb_load_chunk(const int *__restrict__ cols, const double *__restrict__ vals,
                                                         long long ch, bool active)
{
    SslapbChunk c;
    c.cj = make_int4(0, 0, 0, 0); c.va = make_double2(0.0, 0.0); c.vb = c.va;
    if (active) {
        c.cj = __ldg(reinterpret_cast<const int4 *>(cols) + ch);
        c.va = __ldg(reinterpret_cast<const double2 *>(vals) + 2 * ch);
        c.vb = __ldg(reinterpret_cast<const double2 *>(vals) + 2 * ch + 1);
    }
    return c;
}

// One bidder whose row fits a single warp pass (<= 32 aligned chunks, i.e. up to 125..128 entries), row entries already
// in registers (`cur`).  Gathers the object records, finds the winning entry (3 REDUX) and — before finishing the bid —
// requests the row of the object's current owner (`nxt`), the person who bids next if this bid wins.  The second-best
// reduction and the bid arithmetic overlap that load.  Returns false for an empty row.
__device__ __forceinline__ bool sweep_single(const SslapbAuctionParams &P, const SslapbChunk &cur, long long st, int dg,
                                             double eps, bool want_next, SslapbBid &B, SslapbChunk &nxt, bool &nxt_single)
{
    const int lane = threadIdx.x & 31;
    const long long ch = (st >> 2) + lane;
    const int off = (int)((ch << 2) - st);                    // row index of slot 0 (may be negative)
    const bool m0 = (unsigned)off < (unsigned)dg, m1 = (unsigned)(off + 1) < (unsigned)dg;
    const bool m2 = (unsigned)(off + 2) < (unsigned)dg, m3 = (unsigned)(off + 3) < (unsigned)dg;
    const int4 cj = cur.cj;
    SslapbRec256 q0, q1, q2, q3;
    q0.start = q1.start = q2.start = q3.start = 0ull;
    q0.owner_deg = q1.owner_deg = q2.owner_deg = q3.owner_deg = 0xffffffffull;      // owner = -1, deg = 0
    q0.price_bits = q1.price_bits = q2.price_bits = q3.price_bits = 0ull;
    if (m0) q0 = sslapb_ld_rec256(P.rec + cj.x);
    if (m1) q1 = sslapb_ld_rec256(P.rec + cj.y);
    if (m2) q2 = sslapb_ld_rec256(P.rec + cj.z);
    if (m3) q3 = sslapb_ld_rec256(P.rec + cj.w);
    // a_ij - p_j in float64 as the reference; slots of neighbouring rows carry -inf
    const double v0 = m0 ? cur.va.x - __longlong_as_double((long long)q0.price_bits) : SSLAPB_NEG_INF;
    const double v1 = m1 ? cur.va.y - __longlong_as_double((long long)q1.price_bits) : SSLAPB_NEG_INF;
    const double v2 = m2 ? cur.vb.x - __longlong_as_double((long long)q2.price_bits) : SSLAPB_NEG_INF;
    const double v3 = m3 ? cur.vb.y - __longlong_as_double((long long)q3.price_bits) : SSLAPB_NEG_INF;
    const SslapbLaneTop lt = sslapb_lane_top2(v0, v1, v2, v3);
    // a lane whose best value is -inf reports nothing: real -inf candidates (objects priced +inf) cannot be told from the
    // absent slots here, so rows whose candidates are ALL at -inf are left to the exact generic sweep (caller)
    const int bi = ((unsigned)(off + lt.w) < (unsigned)dg && lt.b > SSLAPB_NEG_INF) ? off + lt.w : -1;
    const unsigned long long qs = (lt.w & 2) ? ((lt.w & 1) ? q3.start : q2.start) : ((lt.w & 1) ? q1.start : q0.start);
    const unsigned long long qo = (lt.w & 2) ? ((lt.w & 1) ? q3.owner_deg : q2.owner_deg) : ((lt.w & 1) ? q1.owner_deg : q0.owner_deg);
    // top-1 across the warp -> whose object -> request the owner's row right away
    const unsigned long long bk = bi >= 0 ? sslapb_key_of(lt.b) : 0ull;
    const unsigned bh = (unsigned)(bk >> 32), bl = (unsigned)bk;
    const unsigned khi = __reduce_max_sync(SSLAPB_FULL, bh);
    // Almost always ONE lane holds the maximal high word (two candidates closer than 2^-20 relative, or exact ties, are
    // the exception): then that lane is the winner and the low-word and row-index reductions — two dependent REDUX on
    // the path to the next bidder's row request — are not needed.
    const unsigned hm = __ballot_sync(SSLAPB_FULL, (bh == khi) & (bi >= 0));
    bool iswin;
    unsigned own;
    if (__popc(hm) <= 1) {
        own = hm;
        iswin = (hm >> lane) & 1u;
    } else {
        const unsigned klo = __reduce_max_sync(SSLAPB_FULL, bh == khi ? bl : 0u);
        const bool top = (bh == khi) & (bl == klo);
        const int widx = __reduce_max_sync(SSLAPB_FULL, top ? bi : -1);
        iswin = top & (bi == widx) & (bi >= 0);
        own = __ballot_sync(SSLAPB_FULL, iswin);
    }
    if (own == 0u) return false;
    const int src = __ffs(own) - 1;
    const unsigned long long wo = __shfl_sync(SSLAPB_FULL, qo, src);
    B.pstart = (long long)__shfl_sync(SSLAPB_FULL, qs, src);
    B.powner = (int)(unsigned)wo;
    B.pdeg = (int)(wo >> 32);
    const long long n0 = B.pstart >> 2, n1 = (B.pstart + B.pdeg + 3) >> 2;
    nxt_single = (n1 - n0) <= 32;
    nxt = sslapb_load_chunk(P.cols, P.vals, n0 + lane, want_next && B.powner >= 0 && nxt_single && (n0 + lane < n1));
    // ---- everything below overlaps the load above
    const unsigned long long sk = bi >= 0 ? sslapb_key_of(lt.s) : 0ull;
    const unsigned long long cand = iswin ? sk : bk;
    const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
    const unsigned shi = __reduce_max_sync(SSLAPB_FULL, chh);
    const unsigned slo = __reduce_max_sync(SSLAPB_FULL, chh == shi ? chl : 0u);
    const unsigned long long skey = ((unsigned long long)shi << 32) | slo;
    const double myc = (lt.w & 2) ? ((lt.w & 1) ? cur.vb.y : cur.vb.x) : ((lt.w & 1) ? cur.va.y : cur.va.x);
    const int myj = (lt.w & 2) ? ((lt.w & 1) ? cj.w : cj.z) : ((lt.w & 1) ? cj.y : cj.x);
    const double bc = __shfl_sync(SSLAPB_FULL, myc, src);
    B.j = __shfl_sync(SSLAPB_FULL, myj, src);
    const double wi = skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(skey) : SSLAPB_NEG_INF;   // :344
    B.bid = (bc - wi) + eps;                                   // :360
    return true;
}

// ---- hot-list form of the same step (hot.cu): lane t holds hot entry t of the bidder — ONE record gather per lane, no
// per-lane tournament, a 512-byte row that stays in L2.  Returns false when the hot list cannot prove its answer (the
// second-best value found is not above the bound of the entries outside the list) or finds no candidate: the caller then
// sweeps the full row.  The next bidder's hot row is requested as early as in sweep_single.
struct SslapbHotRow { int col, idx; double a, rest; };
__device__ __forceinline__ SslapbHotRow sslapb_load_hot(const SslapbAuctionParams &P, int person, bool active)
{
    SslapbHotRow h;
    h.col = 0; h.idx = -1; h.a = SSLAPB_NEG_INF; h.rest = __longlong_as_double(0x7ff0000000000000ll);
    if (active) {
        const int4 q = __ldg(reinterpret_cast<const int4 *>(P.hot) + (long long)person * 32 + (threadIdx.x & 31));
        h.col = q.x; h.idx = q.y; h.a = __hiloint2double(q.w, q.z);
        h.rest = P.rest[person];                               // any earlier value of the bound is still a valid bound
    }
    return h;
}
// NEXT: request the hot row of the object's current owner (unconditionally — index clamped to 0 when there is none; the
// callers never use `nxt` then) as soon as the winner is known.  `nxt` is written only when the function returns true.
template <bool NEXT>
__device__ __forceinline__ bool sweep_hot(const SslapbAuctionParams &P, const SslapbHotRow &cur, double eps, SslapbBid &B,
                                          SslapbHotRow &nxt)
{
    const int lane = threadIdx.x & 31;
    // padding entries carry (column 0, a = -inf): gathered like the others, their value -inf - p = -inf never wins
    const SslapbRec256 q = sslapb_ld_rec256(P.rec + cur.col);
    // (measured and dropped: every lane prefetching its candidate's owner's hot row into L1 ahead of the reduction — 128
    // prefetches per round cost more than the L2 round trip they hide: C3 233.7 -> 248.5 ms)
    const double v = (cur.a - __longlong_as_double((long long)q.price_bits)) + 0.0;   // + 0.0 folds -0.0 into +0.0
    const unsigned long long bk = sslapb_ord64(v);             // -inf (padding, objects priced +inf) -> SSLAPB_KEY_NEG_INF
    const unsigned bh = (unsigned)(bk >> 32), bl = (unsigned)bk;
    const unsigned khi = __reduce_max_sync(SSLAPB_FULL, bh);
    if (khi <= (unsigned)(SSLAPB_KEY_NEG_INF >> 32)) return false;     // every candidate at -inf: the exact generic sweep decides
    const unsigned hm = __ballot_sync(SSLAPB_FULL, bh == khi);
    bool iswin;
    unsigned own;
    if ((hm & (hm - 1u)) == 0u) {                              // one lane holds the maximal high word: it is the winner
        own = hm;
        iswin = bh == khi;
    } else {
        const unsigned klo = __reduce_max_sync(SSLAPB_FULL, bh == khi ? bl : 0u);
        const bool top = (bh == khi) & (bl == klo);
        const int widx = __reduce_max_sync(SSLAPB_FULL, top ? cur.idx : -1);   // equal values: the later row entry wins (:351)
        iswin = top & (cur.idx == widx);
        own = __ballot_sync(SSLAPB_FULL, iswin);
    }
    int src;                                                   // the one lane of `own`
    asm("bfind.u32 %0, %1;" : "=r"(src) : "r"(own));
    const unsigned long long wo = __shfl_sync(SSLAPB_FULL, q.owner_deg, src);
    B.pstart = (long long)__shfl_sync(SSLAPB_FULL, q.start, src);
    B.powner = (int)(unsigned)wo;
    B.pdeg = (int)(wo >> 32);
    SslapbHotRow nx;
    nx.col = 0; nx.idx = -1; nx.a = SSLAPB_NEG_INF; nx.rest = SSLAPB_NEG_INF;
    if (NEXT) {
        const unsigned who = B.powner >= 0 ? (unsigned)B.powner : 0u;
        const int4 h4 = __ldg(reinterpret_cast<const int4 *>(reinterpret_cast<const char *>(P.hot) + ((unsigned long long)who << 9)) + lane);
        nx.col = h4.x; nx.idx = h4.y; nx.a = __hiloint2double(h4.w, h4.z);
        nx.rest = *reinterpret_cast<const double *>(reinterpret_cast<const char *>(P.rest) + ((unsigned long long)who << 3));
    }
    // ---- everything below overlaps the load above
    const unsigned long long cand = iswin ? 0ull : bk;         // one candidate per lane: second best = best of the other lanes
    const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
    const unsigned shi = __reduce_max_sync(SSLAPB_FULL, chh);
    const unsigned slo = __reduce_max_sync(SSLAPB_FULL, chh == shi ? chl : 0u);
    const unsigned long long skey = ((unsigned long long)shi << 32) | slo;
    const double bc = __shfl_sync(SSLAPB_FULL, cur.a, src);
    B.j = __shfl_sync(SSLAPB_FULL, cur.col, src);
    const double wi = skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(skey) : SSLAPB_NEG_INF;   // :344
    B.bid = (bc - wi) + eps;                                   // :360
    if (!((wi > cur.rest) || (cur.rest == SSLAPB_NEG_INF))) return false;
    nxt = nx;
    return true;
}

// One bidder of the few-bidder rounds in hot form: decide from the hot list, otherwise sweep the full row with the exact
// generic sweep (bound-pruned when the row is longer than one warp pass) and request the next occupant's hot row.
// (Measured: moving this fallback out of line, with the publication written by either path and read back after the
// barrier, made these rounds slower — 1.20 against 1.06 us — although it removed 50 instructions from the loop; the
// single-warp chain, by contrast, gained 11 % from exactly that treatment, see chain_rounds_hot.)
// `fell` counts the bids the hot list could not decide.  Returns false for a row without any entry.
__device__ __forceinline__ bool hot_bid(const SslapbAuctionParams &P, const SslapbHotRow &cur, int me, long long st, int dg,
                                        double eps, const double *s_bounds, bool want_next, SslapbBid &B, SslapbHotRow &nxt,
                                        int &fell)
{
    if (sweep_hot<true>(P, cur, eps, B, nxt)) return true;
    ++fell;
    const bool single = (((st + dg + 3) >> 2) - (st >> 2)) <= 32;
    B = row_bid_rec<32>(P.cols, P.vals, P.rec, st, st + dg, threadIdx.x & 31, eps, s_bounds[0],
                        single ? SSLAPB_NEG_INF : __ldg(P.rowmax + me) - s_bounds[1]);
    const bool ok = B.j >= 0;
    nxt = sslapb_load_hot(P, B.powner, ok && want_next && B.powner >= 0);
    return ok;
}

// The winner's writes (auction_.pyx:397-418) — executed by one lane.
__device__ __forceinline__ void commit_win(const SslapbAuctionParams &P, int person, long long st, int dg, const SslapbBid &B)
{
    // the whole 32-byte record in ONE store (a reader's single 256-bit load can never see it torn)
    sslapb_st_rec256(P.rec + B.j, (unsigned long long)st, ((unsigned long long)(unsigned)dg << 32) | (unsigned)person,
                     (unsigned long long)__double_as_longlong(B.bid));   // owner + its row (:418), price (:397)
    P.price[B.j] = B.bid;
    // person_to_object (:404, :417) is NOT written here: nothing reads it inside an eps-phase, and two of a commit's four
    // stores (with their address arithmetic) are a measurable share of a single-warp round — rebuild_p2o derives it from
    // the records' owners before every eps-CS sweep and in the epilogue
    (void)person;
}

// Single-bidder chain (52 % of all rounds at N = 100k): person i bids, wins (there is no competitor), evicts the owner i'
// of the object, i' bids, ...  One warp, no barrier; dependent chain per round = row entries -> object records -> top-1
// (see sweep_single).  Rows longer than one warp pass take the generic sweep.  Returns the new count (0 or 1).
__device__ __forceinline__ int chain_rounds(const SslapbAuctionParams &P, double eps, int &li, long long &lst, int &ldg,
                                            long long &its, long long max_iter, int &done, long long &rounds)
{
    const int lane = threadIdx.x & 31;
    li = __shfl_sync(SSLAPB_FULL, li, 0); lst = __shfl_sync(SSLAPB_FULL, lst, 0); ldg = __shfl_sync(SSLAPB_FULL, ldg, 0);
    int nu = 1;
    bool single = (((lst + ldg + 3) >> 2) - (lst >> 2)) <= 32;
    SslapbChunk cur = sslapb_load_chunk(P.cols, P.vals, (lst >> 2) + lane, single && ((lst >> 2) + lane < ((lst + ldg + 3) >> 2)));
    while (nu == 1 && !done && single) {                       // long rows are swept by the whole CTA (coop_chain_rounds)
        SslapbBid B;
        SslapbChunk nxt;
        bool nsingle = false;
        const bool lean = sweep_single(P, cur, lst, ldg, eps, its + 1 < max_iter, B, nxt, nsingle);
        if (!lean) {                                           // every candidate at -inf: exact generic sweep
            B = row_bid_rec<32>(P.cols, P.vals, P.rec, lst, lst + ldg, lane, eps);
            if (B.j < 0) { done = 4; break; }
            const long long n0 = B.pstart >> 2, n1 = (B.pstart + B.pdeg + 3) >> 2;
            nsingle = (n1 - n0) <= 32;
            nxt = sslapb_load_chunk(P.cols, P.vals, n0 + lane, B.powner >= 0 && nsingle && (n0 + lane < n1));
        }
        if (lane == 0) commit_win(P, li, lst, ldg, B);         // the only bidder wins (:379-385, :394-427)
        __syncwarp();
        ++its; ++rounds;
        if (its >= max_iter) done = 3;
        if (B.powner < 0) { nu = 0; li = -1; break; }          // nobody evicted: the frontier is empty
        li = B.powner; lst = B.pstart; ldg = B.pdeg;           // the evicted owner is the next (and only) bidder
        single = nsingle; cur = nxt;
    }
    return nu;
}

// 2..16 bidders: warp a owns list position a for as long as the frontier stays this small.  Per round every active
// warp sweeps its bidder (row entries usually already in registers: the next occupant of a slot is either the same
// person — it lost — or the owner it evicted, whose row was requested during the sweep), publishes (object, bid), and
// after one barrier decides by itself whether it won (no serial merge); winners commit; after a second barrier every
// warp derives the compacted list (push_all_left, auction_.pyx:137-162) redundantly from shared memory.
struct SslapbDuoPub { double bid; long long pstart, st; int j, powner, pdeg, me, dg, pad; };   // 48 B, see the duo rounds

#ifdef SSLAPB_LONG_ROWS
// Very long rows (more than 8 warp passes; only in the kernel instance built with SSLAPB_LONG_ROWS, auction_long.cu): the
// warp that owns such a bidder does not sweep it alone.  It flags the row as pending; the round's first barrier carries
// the vote (BAR.RED.OR); if any row is pending ALL warps of the CTA sweep each pending row together — warp w takes the
// 32-chunk trips w, w+16, ... (rotated per bidder), the partial top-2 go through shared memory — and the owner combines
// them and publishes its bid (two more barriers, only in such rounds).  A dense 3162-entry row costs each warp 1-2 trips
// per bidder instead of 25 dependent trips on the owner.
#define SSLAPB_COOP_CHUNKS 256    // rows of more chunks (more than 8 warp passes, > ~1000 entries) are swept by the whole CTA
__device__ __forceinline__ void coop_multi_pending(int a, int me, long long st, int dg, int *s_j);
__device__ __forceinline__ void coop_multi_phase(const SslapbAuctionParams &P, double eps, const double *s_bounds, int nu, int a,
                                                 bool pend, int me, long long st, int dg, int *s_j, double *s_bidv,
                                                 SslapbBid &B, SslapbChunk &nxt, bool &nsingle, int &done);
#endif
__device__ __forceinline__ int multi_rounds(const SslapbAuctionParams &P, double eps, const double *s_bounds, int nu, int *s_list,
                                            long long *s_start, int *s_deg, int *s_j, double *s_bidv, long long &its,
                                            long long max_iter, int &done, long long &rounds, int &me, long long &st,
                                            int &dg)
{
    const int lane = threadIdx.x & 31, a = __shfl_sync(SSLAPB_FULL, (int)(threadIdx.x >> 5), 0);
    bool active = a < nu;
    bool single = false;
    SslapbChunk cur;
    cur.cj = make_int4(0, 0, 0, 0); cur.va = make_double2(0.0, 0.0); cur.vb = cur.va;
    me = -1; st = 0; dg = 0;
    if (active) {
        me = s_list[a]; st = s_start[a]; dg = s_deg[a];
        single = (((st + dg + 3) >> 2) - (st >> 2)) <= 32;
        cur = sslapb_load_chunk(P.cols, P.vals, (st >> 2) + lane, single && ((st >> 2) + lane < ((st + dg + 3) >> 2)));
    }
#ifdef SSLAPB_LONG_ROWS
    while (nu > 1 && nu <= SSLAPB_THREADS / 32 && !done) {
        SslapbBid B;
        B.j = -1; B.bid = 0.0; B.powner = -1; B.pdeg = 0; B.pstart = 0;
        SslapbChunk nxt = cur;
        bool nsingle = false;
#ifdef SSLAPB_LONG_ROWS
        bool pend = false;
#endif
        if (active) {
            bool ok = single;
            if (ok) ok = sweep_single(P, cur, st, dg, eps, true, B, nxt, nsingle);
#ifdef SSLAPB_LONG_ROWS
            if (!ok && (((st + dg + 3) >> 2) - (st >> 2)) > SSLAPB_COOP_CHUNKS) {
                pend = true;
                if (lane == 0) coop_multi_pending(a, me, st, dg, s_j);
            } else {
#endif
            if (!ok) {                                         // long row (bound-pruned), or every candidate at -inf (exact, unpruned)
                B = row_bid_rec<32>(P.cols, P.vals, P.rec, st, st + dg, lane, eps, s_bounds[0],
                                    single ? SSLAPB_NEG_INF : __ldg(P.rowmax + me) - s_bounds[1]);
                ok = B.j >= 0;
                const long long n0 = B.pstart >> 2, n1 = (B.pstart + B.pdeg + 3) >> 2;
                nsingle = (n1 - n0) <= 32;
                nxt = sslapb_load_chunk(P.cols, P.vals, n0 + lane, ok && B.powner >= 0 && nsingle && (n0 + lane < n1));
            }
            if (!ok) { B.j = -1; done = 4; }
            if (lane == 0) { s_j[a] = B.j; s_bidv[a] = B.bid; }
#ifdef SSLAPB_LONG_ROWS
            }
#endif
        }
#ifdef SSLAPB_LONG_ROWS
        if (__syncthreads_or(pend)) coop_multi_phase(P, eps, s_bounds, nu, a, pend, me, st, dg, s_j, s_bidv, B, nxt, nsingle, done);
#else
        __syncthreads();
#endif
        int nme = me; long long nst = st; int ndg = dg; bool won = false;
        if (active) {                                          // merge (:375-385): do I hold the best bid on my object?
            const int oj = lane < nu ? s_j[lane] : -1;
            const double ob = lane < nu ? s_bidv[lane] : 0.0;
            const bool beaten = (lane < nu) && (lane != a) && (oj == B.j) && (ob > B.bid || (ob == B.bid && lane < a));
            won = (B.j >= 0) && !__any_sync(SSLAPB_FULL, beaten);
            if (won) {
                if (lane == 0) commit_win(P, me, st, dg, B);
                nme = B.powner; nst = B.pstart; ndg = B.pdeg;  // evicted owner takes the slot (:409) or hole (:412)
            }
            if (lane == 0) { s_list[a] = nme; s_start[a] = nst; s_deg[a] = ndg; }
        }
        if (__any_sync(SSLAPB_FULL, lane < nu && s_j[lane] < 0)) done = 4;   // uniform across the CTA (same smem)
        ++its; ++rounds;
        if (its >= max_iter && !done) done = 3;
        __syncthreads();
        // compaction (:429-430), identical in every warp
        const int v = lane < nu ? s_list[lane] : 0;
        const unsigned holes = __ballot_sync(SSLAPB_FULL, lane < nu && v < 0);
        const int new_nu = nu - __popc(holes);
        int src = a;
        if (holes) {
            const unsigned valid = (1u << nu) - 1u, leftm = (1u << new_nu) - 1u;
            const unsigned left_holes = holes & leftm, right_live = valid & ~holes & ~leftm;
            if (a < new_nu && ((left_holes >> a) & 1u)) {
                unsigned m = right_live;
                for (int q = __popc(left_holes & ((1u << a) - 1u)); q > 0; --q) m &= m - 1u;
                src = __ffs(m) - 1;
            }
        }
        const int old_me = me;
        nu = new_nu;
        active = a < nu;
        if (active) {
            if (src != a) { nme = s_list[src]; nst = s_start[src]; ndg = s_deg[src]; }
            me = nme; st = nst; dg = ndg;
            if (src == a && won) { cur = nxt; single = nsingle; }               // the evicted owner: row already requested
            else if (src == a && me == old_me) { /* lost: same person, same row, still in registers */ }
            else {
                single = (((st + dg + 3) >> 2) - (st >> 2)) <= 32;
                cur = sslapb_load_chunk(P.cols, P.vals, (st >> 2) + lane, single && ((st >> 2) + lane < ((st + dg + 3) >> 2)));
            }
        } else {
            me = -1;
        }
        // no third barrier: s_list / s_j are rewritten only after the next round's first barrier / after this round's second
    }
#else
    // ---- 3..16 bidders, ONE named barrier per round among the active warps only.  Every active warp publishes its bid
    // together with its list entry and the record of the object's owner; after the barrier lane b of EVERY active warp
    // holds publication b and the whole outcome is derived redundantly, in registers: who is beaten (:375-385: a higher
    // bid on the same object, or an equal one from an earlier position), the next occupant of every position (evicted
    // owner, the loser itself, or a hole, :401-413), the compaction (push_all_left, :137-162).  Winners commit; a warp whose
    // position falls off the end leaves for the caller's block barrier.  Publications are double-buffered: a warp can be
    // at most one round ahead of any warp that took part in the previous round (leavers included, see below).
    __shared__ SslapbDuoPub s_pub[2][SSLAPB_THREADS / 32];
    int par = 0;
    int bar_warps = nu;                                        // warps expected at the next named barrier
    while (active && nu > 2 && !done) {
        SslapbBid B;
        B.j = -1; B.bid = 0.0; B.powner = -1; B.pdeg = 0; B.pstart = 0;
        SslapbChunk nxt = cur;
        bool nsingle = false;
        bool ok = single;
        if (ok) ok = sweep_single(P, cur, st, dg, eps, true, B, nxt, nsingle);
        if (!ok) {                                             // long row (bound-pruned), or every candidate at -inf (exact, unpruned)
            B = row_bid_rec<32>(P.cols, P.vals, P.rec, st, st + dg, lane, eps, s_bounds[0],
                                single ? SSLAPB_NEG_INF : __ldg(P.rowmax + me) - s_bounds[1]);
            ok = B.j >= 0;
            const long long n0 = B.pstart >> 2, n1 = (B.pstart + B.pdeg + 3) >> 2;
            nsingle = (n1 - n0) <= 32;
            nxt = sslapb_load_chunk(P.cols, P.vals, n0 + lane, ok && B.powner >= 0 && nsingle && (n0 + lane < n1));
        }
        if (!ok) B.j = -1;
        if (lane == 0) {
            SslapbDuoPub pb;
            pb.bid = B.bid; pb.pstart = B.pstart; pb.st = st; pb.j = B.j; pb.powner = B.powner; pb.pdeg = B.pdeg;
            pb.me = me; pb.dg = dg; pb.pad = 0;
            s_pub[par][a] = pb;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(bar_warps * 32) : "memory");
        SslapbDuoPub O;
        O.bid = 0.0; O.pstart = 0; O.st = 0; O.j = -1 - lane; O.powner = -1; O.pdeg = 0; O.me = -1; O.dg = 0; O.pad = 0;
        if (lane < nu) O = s_pub[par][lane];
        par ^= 1;
        if (__any_sync(SSLAPB_FULL, lane < nu && O.j < 0)) { done = 4; break; }   // empty row: rejected at CSR build
        bool beaten = false;
        for (int c = 0; c < nu; ++c) {
            const int cj = __shfl_sync(SSLAPB_FULL, O.j, c);
            const double cb = __shfl_sync(SSLAPB_FULL, O.bid, c);
            beaten |= (c != lane) && (cj == O.j) && (cb > O.bid || (cb == O.bid && c < lane));
        }
        const bool wonb = (lane < nu) && !beaten;              // position `lane` wins its object
        const int nme_b = wonb ? O.powner : O.me;              // next occupant of position `lane`; -1 = hole
        const long long nst_b = wonb ? O.pstart : O.st;
        const int ndg_b = wonb ? O.pdeg : O.dg;
        const bool won = __shfl_sync(SSLAPB_FULL, (int)wonb, a) != 0;
        // Commits (:397-418).  EVERY active warp stores EVERY winner's record (lane b commits publication b; all warps
        // write identical values), so that the loads of a warp's next sweep are ordered after the commits by its own
        // program order + __syncwarp — no second barrier, and no window in which a loser re-sweeps an object whose
        // winner (another warp) has not stored the new price yet.
        if (wonb) {
            SslapbBid Wb;
            Wb.j = O.j; Wb.bid = O.bid; Wb.powner = O.powner; Wb.pdeg = O.pdeg; Wb.pstart = O.pstart;
            commit_win(P, O.me, O.st, O.dg, Wb);
        }
        __syncwarp();
        ++its; ++rounds;
        if (its >= max_iter) done = 3;
        // compaction (:429-430): the k-th hole left of the new count takes the k-th live entry right of it
        const unsigned holes = __ballot_sync(SSLAPB_FULL, lane < nu && nme_b < 0);
        const int new_nu = nu - __popc(holes);
        int src = a;
        if (holes) {
            const unsigned valid = (1u << nu) - 1u, leftm = (1u << new_nu) - 1u;
            const unsigned left_holes = holes & leftm, right_live = valid & ~holes & ~leftm;
            if (a < new_nu && ((left_holes >> a) & 1u)) {
                unsigned m = right_live;
                for (int q = __popc(left_holes & ((1u << a) - 1u)); q > 0; --q) m &= m - 1u;
                src = __ffs(m) - 1;
            }
        }
        const int e_me = __shfl_sync(SSLAPB_FULL, nme_b, src);
        const long long e_st = __shfl_sync(SSLAPB_FULL, nst_b, src);
        const int e_dg = __shfl_sync(SSLAPB_FULL, ndg_b, src);
        // the NEXT barrier still counts this round's leavers: they arrive (without waiting) only after their look at
        // this round's buffer above, so no stayer can get two rounds ahead and overwrite it under them
        bar_warps = nu;
        nu = new_nu;
        active = a < nu;
        if (!active) {                                         // my position fell off the end
            if (nu > 2 && !done) asm volatile("bar.arrive 1, %0;" ::"r"(bar_warps * 32) : "memory");
            me = -1;
            break;
        }
        me = e_me; st = e_st; dg = e_dg;
        if (src == a && won) { cur = nxt; single = nsingle; }  // the evicted owner: row already requested
        else if (src == a) { /* lost: same person, same row, still in registers */ }
        else {
            single = (((st + dg + 3) >> 2) - (st >> 2)) <= 32;
            cur = sslapb_load_chunk(P.cols, P.vals, (st >> 2) + lane, single && ((st >> 2) + lane < ((st + dg + 3) >> 2)));
        }
    }
#endif
#ifndef SSLAPB_LONG_ROWS
    // ---- exactly two bidders (45 % of the few-bidder rounds at C3): warps 0 and 1 alone, ONE named barrier per round.
    // Each publishes its bid together with its own list entry and the record of the object's owner; after the barrier
    // both warps derive the whole outcome from the two publications (who wins a contested object: higher bid, position
    // 0 on a tie, :379; whose slot turns into a hole; push_all_left for two positions) — no second barrier, no list in
    // shared memory, the other 14 warps wait at the caller's block barrier.  The publications are double-buffered: a warp
    // can be at most one round ahead of the other.
    if (nu == 2 && !done && a < 2) {
        __shared__ SslapbDuoPub s_duo[2][2];
        int par = 0;
        while (nu == 2 && !done) {
            SslapbBid B;
            B.j = -1; B.bid = 0.0; B.powner = -1; B.pdeg = 0; B.pstart = 0;
            SslapbChunk nxt = cur;
            bool nsingle = false;
            bool ok = single;
            if (ok) ok = sweep_single(P, cur, st, dg, eps, true, B, nxt, nsingle);
            if (!ok) {                                         // long row (bound-pruned), or every candidate at -inf (exact, unpruned)
                B = row_bid_rec<32>(P.cols, P.vals, P.rec, st, st + dg, lane, eps, s_bounds[0],
                                    single ? SSLAPB_NEG_INF : __ldg(P.rowmax + me) - s_bounds[1]);
                ok = B.j >= 0;
                const long long n0 = B.pstart >> 2, n1 = (B.pstart + B.pdeg + 3) >> 2;
                nsingle = (n1 - n0) <= 32;
                nxt = sslapb_load_chunk(P.cols, P.vals, n0 + lane, ok && B.powner >= 0 && nsingle && (n0 + lane < n1));
            }
            if (!ok) B.j = -1;
            if (lane == 0) {
                SslapbDuoPub pb;
                pb.bid = B.bid; pb.pstart = B.pstart; pb.st = st; pb.j = B.j; pb.powner = B.powner; pb.pdeg = B.pdeg;
                pb.me = me; pb.dg = dg; pb.pad = 0;
                s_duo[par][a] = pb;
            }
            asm volatile("bar.sync 1, 64;" ::: "memory");
            const SslapbDuoPub O = s_duo[par][a ^ 1];
            par ^= 1;
            if (B.j < 0 || O.j < 0) { done = 4; break; }       // empty row: rejected at CSR build, cannot happen
            const bool contested = O.j == B.j;
            const bool won = !contested || B.bid > O.bid || (B.bid == O.bid && a == 0);
            const bool owon = !contested || !won;
            // both warps store BOTH commits (identical values): each warp's next sweep is then ordered after them by
            // its own program order + __syncwarp (see the 3..16 loop)
            if ((lane == 0 && won) || (lane == 1 && owon)) {
                SslapbBid Wb;
                if (lane == 0) Wb = B; else { Wb.j = O.j; Wb.bid = O.bid; Wb.powner = O.powner; Wb.pdeg = O.pdeg; Wb.pstart = O.pstart; }
                commit_win(P, lane == 0 ? me : O.me, lane == 0 ? st : O.st, lane == 0 ? dg : O.dg, Wb);
            }
            __syncwarp();
            // next occupant of my slot / of the other slot (-1 = hole)
            const int nme = won ? B.powner : me, ome = owon ? O.powner : O.me;
            ++its; ++rounds;
            if (its >= max_iter) done = 3;
            if (nme >= 0 && ome >= 0) {                        // both slots stay live
                if (won) { me = B.powner; st = B.pstart; dg = B.pdeg; cur = nxt; single = nsingle; }
                continue;                                      // lost: same person, same row, still in registers
            }
            // push_all_left (:137-162) for two positions: the survivor, if any, ends up at position 0 = warp 0
            nu = (nme >= 0) + (ome >= 0);
            if (nme >= 0) { if (won) { me = B.powner; st = B.pstart; dg = B.pdeg; } }
            else if (ome >= 0) { me = ome; st = owon ? O.pstart : O.st; dg = owon ? O.pdeg : O.dg; }
            else me = -1;
        }
    }
#endif
    return nu;
}

// ----------------------------------------------------------------------------------------------------------------------
// Hot-list forms of chain_rounds / multi_rounds (hot.cu), used in the eps-phases whose probing round found the hot lists
// decisive.  Same round structure, same publications, same commits — only the sweep differs: 512 bytes and one record
// gather per lane instead of the full row and four, and whatever the hot list cannot prove goes to the exact generic sweep.
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int chain_rounds_hot(const SslapbAuctionParams &P, double eps, const double *s_bounds, int &li,
                                                long long &lst, int &ldg, long long &its, long long max_iter, int &done,
                                                long long &rounds, int &fell, bool &hot)
{
    const int lane = threadIdx.x & 31;
    li = __shfl_sync(SSLAPB_FULL, li, 0); lst = __shfl_sync(SSLAPB_FULL, lst, 0); ldg = __shfl_sync(SSLAPB_FULL, ldg, 0);
    int nu = 1;
    const long long rounds0 = rounds;
    int undecided_here = 0;
    // One warp issues one instruction at a time and waits out every dependent latency, so a round costs what its
    // instruction count costs (~4.4 cycles each, measured: tools/gpu_tailprobe.py) — the loop of the decided rounds is
    // kept free of the fallback's code and of the register moves that merging two paths would put on the common one.
    for (;;) {
        SslapbHotRow cur = sslapb_load_hot(P, li, true);
        bool undecided = false;
#pragma unroll 2
        while (!done) {
            SslapbBid B;
            SslapbHotRow nxt;
            if (!sweep_hot<true>(P, cur, eps, B, nxt)) { undecided = true; break; }
            if (lane == 0) commit_win(P, li, lst, ldg, B);     // the only bidder wins (:379-385, :394-427)
            __syncwarp();
            ++its; ++rounds;
            if (its >= max_iter) done = 3;
            if (B.powner < 0) { nu = 0; li = -1; break; }      // nobody evicted: the frontier is empty
            li = B.powner; lst = B.pstart; ldg = B.pdeg;       // the evicted owner is the next (and only) bidder
            cur = nxt;
        }
        if (!undecided) break;
        // ---- the hot list could not decide this bid (rare): rows of more than one warp pass are left to the whole CTA
        // (coop_chain_rounds), the others are swept exactly (every entry gathered) right here
        ++fell; ++undecided_here;
        // a chain the hot lists keep failing on (price wars over objects priced +inf: every candidate at -inf) goes back to
        // the full-row loop for the rest of the phase
        if (undecided_here >= 32 && 2 * (long long)undecided_here > rounds - rounds0) { hot = false; break; }
        if ((((lst + ldg + 3) >> 2) - (lst >> 2)) > 32) break;
        const SslapbBid B = row_bid_rec<32>(P.cols, P.vals, P.rec, lst, lst + ldg, lane, eps);
        if (B.j < 0) { done = 4; break; }
        if (lane == 0) commit_win(P, li, lst, ldg, B);
        __syncwarp();
        ++its; ++rounds;
        if (its >= max_iter) done = 3;
        if (B.powner < 0) { nu = 0; li = -1; break; }
        li = B.powner; lst = B.pstart; ldg = B.pdeg;
        if (done) break;
    }
    return nu;
}

__device__ __forceinline__ int multi_rounds_hot(const SslapbAuctionParams &P, double eps, const double *s_bounds, int nu,
                                                const int *s_list, const long long *s_start, const int *s_deg,
                                                long long &its, long long max_iter, int &done, long long &rounds, int &me,
                                                long long &st, int &dg, int &fell)
{
    const int lane = threadIdx.x & 31, a = __shfl_sync(SSLAPB_FULL, (int)(threadIdx.x >> 5), 0);
    bool active = a < nu;
    SslapbHotRow cur = sslapb_load_hot(P, 0, false);
    me = -1; st = 0; dg = 0;
    if (active) {
        me = s_list[a]; st = s_start[a]; dg = s_deg[a];
        cur = sslapb_load_hot(P, me, true);
    }
    // ---- 3..16 bidders, one named barrier per round among the active warps (see multi_rounds)
    __shared__ SslapbDuoPub s_hpub[2][SSLAPB_THREADS / 32];
    int par = 0;
    int bar_warps = nu;
    while (active && nu > 2 && !done) {
        SslapbBid B;
        B.j = -1; B.bid = 0.0; B.powner = -1; B.pdeg = 0; B.pstart = 0;
        SslapbHotRow nxt;
        if (!hot_bid(P, cur, me, st, dg, eps, s_bounds, true, B, nxt, fell)) B.j = -1;
        if (lane == 0) {
            SslapbDuoPub pb;
            pb.bid = B.bid; pb.pstart = B.pstart; pb.st = st; pb.j = B.j; pb.powner = B.powner; pb.pdeg = B.pdeg;
            pb.me = me; pb.dg = dg; pb.pad = 0;
            s_hpub[par][a] = pb;
        }
        asm volatile("bar.sync 2, %0;" ::"r"(bar_warps * 32) : "memory");
        SslapbDuoPub O;
        O.bid = 0.0; O.pstart = 0; O.st = 0; O.j = -1 - lane; O.powner = -1; O.pdeg = 0; O.me = -1; O.dg = 0; O.pad = 0;
        if (lane < nu) O = s_hpub[par][lane];
        par ^= 1;
        if (__any_sync(SSLAPB_FULL, lane < nu && O.j < 0)) { done = 4; break; }   // empty row: rejected at CSR build
        // who is beaten (:375-385): only positions that share their object with another one can be (inactive lanes carry
        // unique negative ids), and most rounds have none
        bool beaten = false;
        const unsigned peers = __match_any_sync(SSLAPB_FULL, O.j);
        unsigned contested = __ballot_sync(SSLAPB_FULL, peers != (1u << lane));
        while (contested) {
            const int c = __ffs(contested) - 1;
            contested &= contested - 1u;
            const int cj = __shfl_sync(SSLAPB_FULL, O.j, c);
            const double cb = __shfl_sync(SSLAPB_FULL, O.bid, c);
            beaten |= (c != lane) && (cj == O.j) && (cb > O.bid || (cb == O.bid && c < lane));
        }
        const bool wonb = (lane < nu) && !beaten;              // position `lane` wins its object
        const int nme_b = wonb ? O.powner : O.me;              // next occupant of position `lane`; -1 = hole
        const long long nst_b = wonb ? O.pstart : O.st;
        const int ndg_b = wonb ? O.pdeg : O.dg;
        const bool won = __shfl_sync(SSLAPB_FULL, (int)wonb, a) != 0;
        if (wonb) {                                            // every active warp stores every winner's record (see multi_rounds)
            SslapbBid Wb;
            Wb.j = O.j; Wb.bid = O.bid; Wb.powner = O.powner; Wb.pdeg = O.pdeg; Wb.pstart = O.pstart;
            commit_win(P, O.me, O.st, O.dg, Wb);
        }
        __syncwarp();
        ++its; ++rounds;
        if (its >= max_iter) done = 3;
        const unsigned holes = __ballot_sync(SSLAPB_FULL, lane < nu && nme_b < 0);
        const int new_nu = nu - __popc(holes);
        int src = a;
        if (holes) {
            const unsigned valid = (1u << nu) - 1u, leftm = (1u << new_nu) - 1u;
            const unsigned left_holes = holes & leftm, right_live = valid & ~holes & ~leftm;
            if (a < new_nu && ((left_holes >> a) & 1u)) {
                unsigned m = right_live;
                for (int q = __popc(left_holes & ((1u << a) - 1u)); q > 0; --q) m &= m - 1u;
                src = __ffs(m) - 1;
            }
        }
        const int e_me = __shfl_sync(SSLAPB_FULL, nme_b, src);
        const long long e_st = __shfl_sync(SSLAPB_FULL, nst_b, src);
        const int e_dg = __shfl_sync(SSLAPB_FULL, ndg_b, src);
        bar_warps = nu;
        nu = new_nu;
        active = a < nu;
        if (!active) {                                         // my position fell off the end
            if (nu > 2 && !done) asm volatile("bar.arrive 2, %0;" ::"r"(bar_warps * 32) : "memory");
            me = -1;
            break;
        }
        me = e_me; st = e_st; dg = e_dg;
        if (src == a && won) cur = nxt;                        // the evicted owner: hot row already requested
        else if (src != a) cur = sslapb_load_hot(P, me, true); // moved here by the compaction
        // (lost: same person, same hot row, still in registers)
    }
    // ---- exactly two bidders: warps 0 and 1, closed-form outcome (see the duo rounds of multi_rounds)
    if (nu == 2 && !done && a < 2) {
        __shared__ SslapbDuoPub s_hduo[2][2];
        int dpar = 0;
        while (nu == 2 && !done) {
            SslapbBid B;
            B.j = -1; B.bid = 0.0; B.powner = -1; B.pdeg = 0; B.pstart = 0;
            SslapbHotRow nxt;
            if (!hot_bid(P, cur, me, st, dg, eps, s_bounds, true, B, nxt, fell)) B.j = -1;
            if (lane == 0) {
                SslapbDuoPub pb;
                pb.bid = B.bid; pb.pstart = B.pstart; pb.st = st; pb.j = B.j; pb.powner = B.powner; pb.pdeg = B.pdeg;
                pb.me = me; pb.dg = dg; pb.pad = 0;
                s_hduo[dpar][a] = pb;
            }
            asm volatile("bar.sync 3, 64;" ::: "memory");
            const SslapbDuoPub O = s_hduo[dpar][a ^ 1];
            dpar ^= 1;
            if (B.j < 0 || O.j < 0) { done = 4; break; }       // empty row: rejected at CSR build, cannot happen
            const bool contested = O.j == B.j;
            const bool won = !contested || B.bid > O.bid || (B.bid == O.bid && a == 0);
            const bool owon = !contested || !won;
            if ((lane == 0 && won) || (lane == 1 && owon)) {   // both warps store both commits (identical values)
                SslapbBid Wb;
                if (lane == 0) Wb = B; else { Wb.j = O.j; Wb.bid = O.bid; Wb.powner = O.powner; Wb.pdeg = O.pdeg; Wb.pstart = O.pstart; }
                commit_win(P, lane == 0 ? me : O.me, lane == 0 ? st : O.st, lane == 0 ? dg : O.dg, Wb);
            }
            __syncwarp();
            const int nme = won ? B.powner : me, ome = owon ? O.powner : O.me;
            ++its; ++rounds;
            if (its >= max_iter) done = 3;
            if (nme >= 0 && ome >= 0) {                        // both slots stay live
                if (won) { me = B.powner; st = B.pstart; dg = B.pdeg; cur = nxt; }
                continue;                                      // lost: same person, same hot row, still in registers
            }
            nu = (nme >= 0) + (ome >= 0);
            if (nme >= 0) { if (won) { me = B.powner; st = B.pstart; dg = B.pdeg; } }
            else if (ome >= 0) { me = ome; st = owon ? O.pstart : O.st; dg = owon ? O.pdeg : O.dg; }
            else me = -1;
        }
    }
    return nu;
}
// ---- long rows (more than one warp pass, e.g. dense inputs) in the single-bidder chain: the WHOLE CTA sweeps the row,
// warp w taking the 32-chunk trips w, w+16, ...; the per-warp top-2 go through shared memory and warp 0 combines them,
// commits and publishes the next bidder.  Two block barriers per round, independent of the row length up to 2048 entries
// per pass (a 10^4-entry row takes 5 trips per warp instead of 79 dependent trips on one warp).
struct SslapbPartial {
    unsigned long long bk, sk;    // best / second-best key seen by this warp (0 = none)
    double bc;                    // value a_ij of the best entry
    int bi, bj;                   // its row index / column
    int4 br;                      // its object's record {start lo, start hi, owner, deg}
};
struct SslapbCoopRow { long long st; double thr; int me, dg, state; };   // state: 0 sweep this (long) row, 1 next row is short,
                                                                        // 2 frontier empty, 3 max_iter, 4 abort

__device__ __forceinline__ SslapbPartial row_partial_rec(const int *__restrict__ cols, const double *__restrict__ vals,
                                                         const SslapbObjRec *rec, long long start, long long end, int lane,
                                                         int trip0, int tstride, double thr)
{
    const int4 *c4 = reinterpret_cast<const int4 *>(cols);
    const double2 *v2 = reinterpret_cast<const double2 *>(vals);
    unsigned long long b = 0ull, s = 0ull;
    double bc = 0.0;
    int bi = -1, bj = -1;
    int4 br = make_int4(0, 0, -1, 0);
    const long long c0 = start >> 2, c1 = (end + 3) >> 2;
    const int trips = (int)((c1 - c0 + 31) >> 5);
#ifdef SSLAPB_LONG_ROWS
    // software pipeline: the chunk of this warp's next trip is requested before the current one is processed (the rows
    // of this instance take several trips per warp; the entries come from HBM, the records from L2)
    long long chn = c0 + (long long)trip0 * 32 + lane;
    int4 cjn = make_int4(0, 0, 0, 0);
    double2 van = make_double2(0.0, 0.0), vbn = van;
    if (trip0 < trips && chn < c1) { cjn = __ldg(c4 + chn); van = __ldg(v2 + 2 * chn); vbn = __ldg(v2 + 2 * chn + 1); }
#pragma unroll 1
    for (int it = trip0; it < trips; it += tstride) {
        const long long ch = chn;
        const int4 cj = cjn;
        const double2 va = van, vb = vbn;
        chn = ch + (long long)tstride * 32;
        if (it + tstride < trips && chn < c1) { cjn = __ldg(c4 + chn); van = __ldg(v2 + 2 * chn); vbn = __ldg(v2 + 2 * chn + 1); }
        if (ch >= c1) continue;
#else
#pragma unroll 1
    for (int it = trip0; it < trips; it += tstride) {
        const long long ch = c0 + (long long)it * 32 + lane;
        if (ch >= c1) continue;
        const int4 cj = __ldg(c4 + ch);
        const double2 va = __ldg(v2 + 2 * ch), vb = __ldg(v2 + 2 * ch + 1);
#endif
        const int lo = (int)(start - (ch << 2)), hi = (int)min(end - (ch << 2), 4ll);
        const bool m0 = (0 >= lo) & (0 < hi) & (va.x >= thr), m1 = (1 >= lo) & (1 < hi) & (va.y >= thr);   // bound pruning
        const bool m2 = (2 >= lo) & (2 < hi) & (vb.x >= thr), m3 = (3 >= lo) & (3 < hi) & (vb.y >= thr);
        SslapbRec256 q0, q1, q2, q3;
        q0.start = q1.start = q2.start = q3.start = 0ull;
        q0.owner_deg = q1.owner_deg = q2.owner_deg = q3.owner_deg = 0xffffffffull;
        q0.price_bits = q1.price_bits = q2.price_bits = q3.price_bits = 0ull;
        if (m0) q0 = sslapb_ld_rec256(rec + cj.x);
        if (m1) q1 = sslapb_ld_rec256(rec + cj.y);
        if (m2) q2 = sslapb_ld_rec256(rec + cj.z);
        if (m3) q3 = sslapb_ld_rec256(rec + cj.w);
        const unsigned long long k0 = sslapb_vkey(va.x, __longlong_as_double((long long)q0.price_bits), m0);
        const unsigned long long k1 = sslapb_vkey(va.y, __longlong_as_double((long long)q1.price_bits), m1);
        const unsigned long long k2 = sslapb_vkey(vb.x, __longlong_as_double((long long)q2.price_bits), m2);
        const unsigned long long k3 = sslapb_vkey(vb.y, __longlong_as_double((long long)q3.price_bits), m3);
        const bool w01 = k1 >= k0, w23 = k3 >= k2;
        const unsigned long long b01 = w01 ? k1 : k0, l01 = w01 ? k0 : k1;
        const unsigned long long b23 = w23 ? k3 : k2, l23 = w23 ? k2 : k3;
        const bool wf = b23 >= b01;
        const unsigned long long b4 = wf ? b23 : b01;
        const unsigned long long s4 = wf ? (b01 > l23 ? b01 : l23) : (b23 > l01 ? b23 : l01);
        const int w4 = wf ? (w23 ? 3 : 2) : (w01 ? 1 : 0);
        if (b4 >= b && b4 != 0ull) {                           // later trips hold later entries: they win equal keys (:351)
            s = b > s4 ? b : s4;
            b = b4;
            bi = (int)((ch << 2) - start) + w4;
            bc = (w4 & 2) ? ((w4 & 1) ? vb.y : vb.x) : ((w4 & 1) ? va.y : va.x);
            bj = (w4 & 2) ? ((w4 & 1) ? cj.w : cj.z) : ((w4 & 1) ? cj.y : cj.x);
            const SslapbRec256 &q = (w4 & 2) ? ((w4 & 1) ? q3 : q2) : ((w4 & 1) ? q1 : q0);
            br = make_int4((int)q.start, (int)(q.start >> 32), (int)q.owner_deg, (int)(q.owner_deg >> 32));
        } else {
            s = b4 > s ? b4 : s;
        }
    }
    // warp-level top-2 (keys) + payload of the winning lane
    const unsigned bh = (unsigned)(b >> 32), bl = (unsigned)b;
    const unsigned hi = __reduce_max_sync(SSLAPB_FULL, bh);
    const unsigned lo = __reduce_max_sync(SSLAPB_FULL, bh == hi ? bl : 0u);
    const bool top = (bh == hi) & (bl == lo);
    const int widx = __reduce_max_sync(SSLAPB_FULL, top ? bi : -1);
    const bool iswin = top & (bi == widx) & (bi >= 0);
    const unsigned long long cand = iswin ? s : b;
    const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
    const unsigned shi = __reduce_max_sync(SSLAPB_FULL, chh);
    const unsigned slo = __reduce_max_sync(SSLAPB_FULL, chh == shi ? chl : 0u);
    const unsigned own = __ballot_sync(SSLAPB_FULL, iswin);
    const int src = own ? (__ffs(own) - 1) : lane;
    SslapbPartial o;
    o.bk = own ? (((unsigned long long)hi << 32) | lo) : 0ull;
    o.sk = ((unsigned long long)shi << 32) | slo;
    o.bi = own ? widx : -1;
    o.bc = __shfl_sync(SSLAPB_FULL, bc, src);
    o.bj = __shfl_sync(SSLAPB_FULL, bj, src);
    o.br.x = __shfl_sync(SSLAPB_FULL, br.x, src); o.br.y = __shfl_sync(SSLAPB_FULL, br.y, src);
    o.br.z = __shfl_sync(SSLAPB_FULL, br.z, src); o.br.w = __shfl_sync(SSLAPB_FULL, br.w, src);
    return o;
}

#ifdef SSLAPB_LONG_ROWS
struct SslapbCoopMulti {
    SslapbPartial part[(SSLAPB_THREADS / 32) * (SSLAPB_THREADS / 32)];   // [bidder][sweeping warp]
    long long st[SSLAPB_THREADS / 32];
    int me[SSLAPB_THREADS / 32], dg[SSLAPB_THREADS / 32];
};
__device__ __forceinline__ SslapbCoopMulti *coop_multi_store()
{
    __shared__ SslapbCoopMulti s_cm;
    return &s_cm;
}
__device__ __forceinline__ void coop_multi_pending(int a, int me, long long st, int dg, int *s_j)
{
    SslapbCoopMulti *cm = coop_multi_store();
    cm->me[a] = me; cm->st[a] = st; cm->dg[a] = dg;
    s_j[a] = -2;
}
// Warp-level combination of the per-warp partials of ONE row (lane l holds partial l; lanes without one hold "none").
// Returns 0 and the bid in B; 1 when the bound-pruned result is not proven exact (fl(thr - pmin) < second-best fails: the
// row must be swept again with every entry gathered); 2 when the row has no entry.
__device__ __forceinline__ int combine_partials(const SslapbPartial &q, double eps, double thr, double pmin, SslapbBid &B)
{
    const unsigned bh = (unsigned)(q.bk >> 32), bl = (unsigned)q.bk;
    const unsigned hi = __reduce_max_sync(SSLAPB_FULL, bh);
    const unsigned lo = __reduce_max_sync(SSLAPB_FULL, bh == hi ? bl : 0u);
    const bool top = (bh == hi) & (bl == lo);
    const int widx = __reduce_max_sync(SSLAPB_FULL, top ? q.bi : -1);
    const bool iswin = top & (q.bi == widx) & (q.bi >= 0);
    const unsigned long long cand = iswin ? q.sk : q.bk;
    const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
    const unsigned shi = __reduce_max_sync(SSLAPB_FULL, chh);
    const unsigned slo = __reduce_max_sync(SSLAPB_FULL, chh == shi ? chl : 0u);
    const unsigned long long skey = ((unsigned long long)shi << 32) | slo;
    const unsigned own = __ballot_sync(SSLAPB_FULL, iswin);
    const double wi = skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(skey) : SSLAPB_NEG_INF;   // :344
    if (thr > SSLAPB_NEG_INF && !((thr - pmin) < wi)) return 1;
    if (own == 0u) return 2;
    const int src = __ffs(own) - 1;
    const double bc = __shfl_sync(SSLAPB_FULL, q.bc, src);
    B.j = __shfl_sync(SSLAPB_FULL, q.bj, src);
    const int sx = __shfl_sync(SSLAPB_FULL, q.br.x, src), sy = __shfl_sync(SSLAPB_FULL, q.br.y, src);
    B.powner = __shfl_sync(SSLAPB_FULL, q.br.z, src);
    B.pdeg = __shfl_sync(SSLAPB_FULL, q.br.w, src);
    B.pstart = (long long)(((unsigned long long)(unsigned)sy << 32) | (unsigned)sx);
    B.bid = (bc - wi) + eps;                                   // :360
    return 0;
}
// after the round's first barrier, by every warp of the CTA (a = warp index = list position it owns, if any)
__device__ __forceinline__ void coop_multi_phase(const SslapbAuctionParams &P, double eps, const double *s_bounds, int nu, int a,
                                                 bool pend, int me, long long st, int dg, int *s_j, double *s_bidv,
                                                 SslapbBid &B, SslapbChunk &nxt, bool &nsingle, int &done)
{
    constexpr int NW = SSLAPB_THREADS / 32;
    const int lane = threadIdx.x & 31;
    SslapbCoopMulti *cm = coop_multi_store();
    for (int a2 = 0; a2 < nu; ++a2) {
        if (s_j[a2] != -2) continue;
        const long long st2 = cm->st[a2];
        const SslapbPartial part = row_partial_rec(P.cols, P.vals, P.rec, st2, st2 + cm->dg[a2], lane, (a + 5 * a2) & (NW - 1), NW,
                                                   __ldg(P.rowmax + cm->me[a2]) - s_bounds[1]);
        if (lane == 0) cm->part[a2 * NW + a] = part;
    }
    __syncthreads();
    if (pend) {
        SslapbPartial q;
        q.bk = 0ull; q.sk = 0ull; q.bc = 0.0; q.bi = -1; q.bj = -1; q.br = make_int4(0, 0, -1, 0);
        if (lane < NW) q = cm->part[a * NW + lane];
        if (combine_partials(q, eps, __ldg(P.rowmax + me) - s_bounds[1], s_bounds[0], B))
            B = row_bid_rec<32>(P.cols, P.vals, P.rec, st, st + dg, lane, eps);   // pruning unproven: gather everything
        const bool ok = B.j >= 0;
        const long long n0 = B.pstart >> 2, n1 = (B.pstart + B.pdeg + 3) >> 2;
        nsingle = (n1 - n0) <= 32;
        nxt = sslapb_load_chunk(P.cols, P.vals, n0 + lane, ok && B.powner >= 0 && nsingle && (n0 + lane < n1));
        if (!ok) { B.j = -1; done = 4; }
        if (lane == 0) { s_j[a] = B.j; s_bidv[a] = B.bid; }
    }
    __syncthreads();
}
#endif

// executed by ALL warps of CTA 0 while the single bidder's row is long; warp 0 carries the list entry in (li, lst, ldg)
__device__ __forceinline__ int coop_chain_rounds(const SslapbAuctionParams &P, double eps, double pmin, double spread,
                                                 SslapbPartial *s_part, SslapbCoopRow *s_row, int &li, long long &lst,
                                                 int &ldg, long long &its, long long max_iter, int &done, long long &rounds)
{
    const int lane = threadIdx.x & 31, warp = __shfl_sync(SSLAPB_FULL, (int)(threadIdx.x >> 5), 0);
    constexpr int NW = SSLAPB_THREADS / 32;
    int nu = 1;
    for (;;) {
        const long long st = s_row->st;
        const int dg = s_row->dg;
        const double thr = s_row->thr;
        const SslapbPartial part = row_partial_rec(P.cols, P.vals, P.rec, st, st + dg, lane, warp, NW, thr);
        if (lane == 0) s_part[warp] = part;
        __syncthreads();
        if (warp == 0) {
            SslapbPartial q;
            q.bk = 0ull; q.sk = 0ull; q.bc = 0.0; q.bi = -1; q.bj = -1; q.br = make_int4(0, 0, -1, 0);
            if (lane < NW) q = s_part[lane];
            const unsigned bh = (unsigned)(q.bk >> 32), bl = (unsigned)q.bk;
            const unsigned hi = __reduce_max_sync(SSLAPB_FULL, bh);
            const unsigned lo = __reduce_max_sync(SSLAPB_FULL, bh == hi ? bl : 0u);
            const bool top = (bh == hi) & (bl == lo);
            const int widx = __reduce_max_sync(SSLAPB_FULL, top ? q.bi : -1);
            const bool iswin = top & (q.bi == widx) & (q.bi >= 0);
            const unsigned long long cand = iswin ? q.sk : q.bk;
            const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
            const unsigned shi = __reduce_max_sync(SSLAPB_FULL, chh);
            const unsigned slo = __reduce_max_sync(SSLAPB_FULL, chh == shi ? chl : 0u);
            const unsigned long long skey = ((unsigned long long)shi << 32) | slo;
            const unsigned own = __ballot_sync(SSLAPB_FULL, iswin);
            const double wi0 = skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(skey) : SSLAPB_NEG_INF;
            int state;
            double nthr = SSLAPB_NEG_INF;
            if (thr > SSLAPB_NEG_INF && !((thr - pmin) < wi0)) {
                state = 0;                                     // the skipped entries might matter: same row again, gather all
            } else if (own == 0u) { state = 4; done = 4; }     // empty row: rejected at CSR build, cannot happen
            else {
                const int src = __ffs(own) - 1;
                SslapbBid B;
                const double bc = __shfl_sync(SSLAPB_FULL, q.bc, src);
                B.j = __shfl_sync(SSLAPB_FULL, q.bj, src);
                const int sx = __shfl_sync(SSLAPB_FULL, q.br.x, src), sy = __shfl_sync(SSLAPB_FULL, q.br.y, src);
                B.powner = __shfl_sync(SSLAPB_FULL, q.br.z, src);
                B.pdeg = __shfl_sync(SSLAPB_FULL, q.br.w, src);
                B.pstart = (long long)(((unsigned long long)(unsigned)sy << 32) | (unsigned)sx);
                const double wi = skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(skey) : SSLAPB_NEG_INF;   // :344
                B.bid = (bc - wi) + eps;                       // :360
                if (lane == 0) commit_win(P, li, lst, ldg, B); // the only bidder wins (:379-385, :394-427)
                __syncwarp();
                ++its; ++rounds;
                if (B.powner < 0) { nu = 0; li = -1; state = 2; }
                else {
                    li = B.powner; lst = B.pstart; ldg = B.pdeg;
                    state = ((((lst + ldg + 3) >> 2) - (lst >> 2)) > 32) ? 0 : 1;
                    if (state == 0) nthr = __ldg(P.rowmax + li) - spread;
                }
                if (its >= max_iter) { done = 3; state = 3; }
            }
            if (lane == 0) { s_row->st = lst; s_row->dg = ldg; s_row->me = li; s_row->state = state; s_row->thr = nthr; }
        }
        __syncthreads();
        if (s_row->state != 0) break;
        __syncthreads();                                       // s_row is rewritten only after everybody has read the state
    }
    return nu;
}

// CTA 0 finishes the eps-phase alone once nu <= t_small (nu only shrinks inside a phase).
__device__ __forceinline__ void small_regime(const SslapbAuctionParams &P, SslapbCtrl *C, int nu, float eps_f,
                                             long long its, long long max_iter, double pmin, double spread, bool hot)
{
    __shared__ int s_list[32], s_deg[32], s_j[32];
    __shared__ long long s_start[32];
    __shared__ double s_bidv[32];
    __shared__ SslapbBid s_bid[32];
    __shared__ int s_nu, s_done;
    __shared__ long long s_its, s_rw, s_rs;
    __shared__ SslapbPartial s_part[SSLAPB_THREADS / 32];
    __shared__ SslapbCoopRow s_row;
    __shared__ double s_bounds[2];                             // {pmin, spread}: only the rare long-row paths read them

    // warp index through a shuffle: tells the compiler it is warp-uniform (same idiom as cutlass::canonical_warp_idx_sync)
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(SSLAPB_FULL, tid >> 5, 0);
    const double eps = (double)eps_f;
    int li = -1, ldg = 0, done = 0;
    long long lst = 0;
    long long rw = 0, rs = 0;
    int fell = 0;                                              // hot mode: bids of this warp the hot list could not decide
    if (tid == 0) { s_bounds[0] = pmin; s_bounds[1] = spread; }
    if (warp == 0) {
        if (lane < nu) {
            const int v = P.list[lane];
            li = v < -1 ? P.mover[-(v + 2)] : v;               // grid regime leaves rank-encoded holes, see below
            lst = __ldg(P.rowptr + li);
            ldg = (int)(__ldg(P.rowptr + li + 1) - lst);
            s_list[lane] = li; s_start[lane] = lst; s_deg[lane] = ldg;
        }
    }
    __syncthreads();

    unsigned long long tw0 = sslapb_globaltimer();
    // ---- 17..32 bidders: positions strided over the warps, warp 0 merges; 2 block barriers per round
    while (nu > SSLAPB_THREADS / 32 && !done) {
        for (int a = warp; a < nu; a += SSLAPB_THREADS / 32) {
            const long long st = s_start[a];
            SslapbBid b;
            bool ok = false;
            if (hot) {
                SslapbHotRow unused;
                ok = sweep_hot<false>(P, sslapb_load_hot(P, s_list[a], true), eps, b, unused);
                if (!ok) ++fell;
            }
            if (!ok) b = row_bid_rec<32>(P.cols, P.vals, P.rec, st, st + s_deg[a], lane, eps, s_bounds[0],
                                         __ldg(P.rowmax + s_list[a]) - s_bounds[1]);
            if (lane == 0) s_bid[a] = b;
        }
        __syncthreads();
        if (warp == 0) {
            SslapbBid b;
            if (lane < nu) b = s_bid[lane]; else { b.j = -1; b.bid = 0.0; b.powner = -1; b.pdeg = 0; b.pstart = 0; }
            nu = warp_resolve(P, nu, li, lst, ldg, b);
            ++its; ++rw;
            if (its >= max_iter) done = 3;
            s_list[lane] = li; s_start[lane] = lst; s_deg[lane] = ldg;
            if (lane == 0) { s_nu = nu; s_done = done; s_its = its; }
        }
        __syncthreads();
        nu = __shfl_sync(SSLAPB_FULL, s_nu, 0); done = __shfl_sync(SSLAPB_FULL, s_done, 0);
        its = __shfl_sync(SSLAPB_FULL, s_its, 0);
    }
    // ---- 2..16 bidders: one warp per list position, distributed merge
    if (nu > 1 && !done && hot) nu = multi_rounds_hot(P, eps, s_bounds, nu, s_list, s_start, s_deg, its, max_iter, done, rw, li, lst, ldg, fell);
    else if (nu > 1 && !done) nu = multi_rounds(P, eps, s_bounds, nu, s_list, s_start, s_deg, s_j, s_bidv, its, max_iter, done, rw, li, lst, ldg);

    unsigned long long tw1 = sslapb_globaltimer();
    // ---- single-bidder chain: warp 0 alone, no block barrier (after multi_rounds warp a holds position a in registers);
    // whenever the bidder's row is longer than one warp pass the whole CTA sweeps it (coop_chain_rounds)
    for (;;) {
        if (warp == 0) {
            if (nu == 1 && !done && hot) nu = chain_rounds_hot(P, eps, s_bounds, li, lst, ldg, its, max_iter, done, rs, fell, hot);
            else if (nu == 1 && !done) nu = chain_rounds(P, eps, li, lst, ldg, its, max_iter, done, rs);
            const bool longrow = nu == 1 && !done;             // chain_rounds stops in front of a long row
            if (lane == 0) {
                s_row.st = lst; s_row.dg = ldg; s_row.me = li; s_row.state = longrow ? 0 : 1;
                s_row.thr = longrow ? __ldg(P.rowmax + li) - s_bounds[1] : SSLAPB_NEG_INF;
                s_nu = nu; s_done = done; s_its = its; s_rw = rw; s_rs = rs;
            }
        }
        __syncthreads();
        if (s_row.state != 0) break;
        __syncthreads();
        const int cnu = coop_chain_rounds(P, eps, s_bounds[0], s_bounds[1], s_part, &s_row, li, lst, ldg, its, max_iter, done, rs);
        if (warp == 0) nu = cnu;
        __syncthreads();
    }
    if (s_nu > 0 && warp < s_nu && lane == 0 && li >= 0 && (s_nu <= SSLAPB_THREADS / 32)) P.list[warp] = li;   // max_iter exit only
    if (tid == 0) {
        C->nu = s_nu;
        C->its = s_its;
        if (s_done == 4) *(volatile int *)&C->abort_flag = 2;
        else if (s_done) C->done = s_done;
        C->rounds_warp += s_rw;
        C->rounds_solo += s_rs;
        C->prof[3] += tw1 - tw0;
        C->prof[4] += sslapb_globaltimer() - tw1;
        if (hot) C->hot_tail[0] += s_rw + s_rs;                // rounds run in hot form
    }
    if (fell && lane == 0) atomicAdd((unsigned long long *)&C->hot_tail[1], (unsigned long long)fell);
}

// ----------------------------------------------------------------------------------------------------------------------
// Mid regime: 32 < nu <= t_mid (at most SSLAPB_MID) bidders in a phase whose bids the hot lists decide.  CTA 0 runs these
// rounds alone with block barriers instead of grid barriers (a grid round of 33..128 bidders costs ~7.6 us at C3, of which
// 3.4 us are barriers; 72 % of C3's grid rounds are of this size).
//   bidding   list positions strided over the 16 warps, SSLAPB_MID_BATCH positions of a warp in flight: their hot rows
//             are requested together and the record gather of the next position is issued before the reduction of the
//             current one (sweep_hot_q); the rare bid a hot list cannot decide takes the exact full-row sweep afterwards;
//             a decided bid prefetches the hot row of the object's owner, who takes the position if the bid wins;
//   merge     2 threads per position (4 when nu <= 128) scan the round's bids, four per shared-memory load, for a
//             competitor on the same object (strict '>' at auction_.pyx:379: the earliest bidder in list order keeps an
//             equal bid); winners commit (:397-418), the evicted owner takes the winner's slot (:409) or leaves a hole (:412);
//   compaction push_all_left (:137-162) by ballot words, only in rounds that produced a hole (15 % at C3).
// The list is double-buffered in shared memory and handed to small_regime through P.list and the control block once
// nu <= 32.  Measured at C3 (profiles/r2_mid_regime_ab.log, -DSSLAPB_MID_PROF builds): 3,898 rounds at 3.65 us — bidding
// 2.1, merge 1.4, compaction 0.1 — against 7.6 us in the grid regime; the solve 199.6 -> 185.4 ms on the same box.
// ----------------------------------------------------------------------------------------------------------------------
// sweep_hot<false> with the record gather taken out (the caller issues it — for the NEXT position before this one is
// reduced — so that the gathers of a warp's positions overlap).  Same arithmetic, same exactness test.
__device__ __forceinline__ bool sweep_hot_q(const SslapbHotRow &cur, const SslapbRec256 &q, double eps, SslapbBid &B)
{
    const double v = (cur.a - __longlong_as_double((long long)q.price_bits)) + 0.0;   // + 0.0 folds -0.0 into +0.0
    const unsigned long long bk = sslapb_ord64(v);
    const unsigned bh = (unsigned)(bk >> 32), bl = (unsigned)bk;
    const unsigned khi = __reduce_max_sync(SSLAPB_FULL, bh);
    if (khi <= (unsigned)(SSLAPB_KEY_NEG_INF >> 32)) return false;     // every candidate at -inf: the exact generic sweep decides
    const unsigned hm = __ballot_sync(SSLAPB_FULL, bh == khi);
    bool iswin;
    unsigned own;
    if ((hm & (hm - 1u)) == 0u) {                              // one lane holds the maximal high word: it is the winner
        own = hm;
        iswin = bh == khi;
    } else {
        const unsigned klo = __reduce_max_sync(SSLAPB_FULL, bh == khi ? bl : 0u);
        const bool top = (bh == khi) & (bl == klo);
        const int widx = __reduce_max_sync(SSLAPB_FULL, top ? cur.idx : -1);   // equal values: the later row entry wins (:351)
        iswin = top & (cur.idx == widx);
        own = __ballot_sync(SSLAPB_FULL, iswin);
    }
    int src;                                                   // the one lane of `own`
    asm("bfind.u32 %0, %1;" : "=r"(src) : "r"(own));
    const unsigned long long wo = __shfl_sync(SSLAPB_FULL, q.owner_deg, src);
    B.pstart = (long long)__shfl_sync(SSLAPB_FULL, q.start, src);
    B.powner = (int)(unsigned)wo;
    B.pdeg = (int)(wo >> 32);
    const unsigned long long cand = iswin ? 0ull : bk;         // one candidate per lane: second best = best of the other lanes
    const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
    const unsigned shi = __reduce_max_sync(SSLAPB_FULL, chh);
    const unsigned slo = __reduce_max_sync(SSLAPB_FULL, chh == shi ? chl : 0u);
    const unsigned long long skey = ((unsigned long long)shi << 32) | slo;
    const double bc = __shfl_sync(SSLAPB_FULL, cur.a, src);
    B.j = __shfl_sync(SSLAPB_FULL, cur.col, src);
    const double wi = skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(skey) : SSLAPB_NEG_INF;   // :344
    B.bid = (bc - wi) + eps;                                   // :360
    return (wi > cur.rest) || (cur.rest == SSLAPB_NEG_INF);
}
// Inlined, and called from the LAST branch of the kernel's regime chain.  Both were measured (tools/gpu_mid_ab.sh, same box):
// out of line, the compiler no longer takes the warps for converged after the call and every warp collective of the code
// that follows gets a divergence check; inlined in front of small_regime, the few-bidder loops move 17 KB down the binary —
// either way they slow down from 1.03 to 1.10-1.16 us per round (7-12 ms at C3, more than the regime gains).
#ifndef SSLAPB_MID_PROF
#define SSLAPB_MID_PROF 0        // profiling builds: prof[6] = 1 bidding / 2 merge / 3 compaction part of the rounds, 4 counters
#endif
#ifndef SSLAPB_MID_BATCH
#define SSLAPB_MID_BATCH 4       // positions a warp keeps in flight (hot rows requested together, gathers software-pipelined)
#endif
static __device__ __forceinline__ void mid_regime(const SslapbAuctionParams &Pk, SslapbCtrl *C, int nu, float eps_f, long long its,
                                                     long long max_iter, double pmin, double spread)
{
    int done = 0;
    nu = __shfl_sync(SSLAPB_FULL, nu, 0);                      // warp-uniform for the compiler (as is `warp` below)
    SslapbAuctionParams P;                                     // the fields the rounds use (the list itself is read and written
    P.hot = Pk.hot; P.rest = Pk.rest; P.rec = Pk.rec;          // through Pk, once each)
    P.price = Pk.price; P.cols = Pk.cols; P.vals = Pk.vals; P.rowmax = Pk.rowmax;
    __shared__ int m_li[2][SSLAPB_MID], m_dg[2][SSLAPB_MID];
    __shared__ long long m_st[2][SSLAPB_MID];
    __shared__ __align__(16) int m_j[SSLAPB_MID];
    __shared__ int m_po[SSLAPB_MID], m_pd[SSLAPB_MID];
    __shared__ double m_bid[SSLAPB_MID];
    __shared__ long long m_ps[SSLAPB_MID];
    __shared__ unsigned m_hw[SSLAPB_MID / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(SSLAPB_FULL, tid >> 5, 0);
    constexpr int NW = SSLAPB_THREADS / 32;
    const double eps = (double)eps_f;
    const unsigned long long t0 = sslapb_globaltimer();
    long long rounds = 0;
    int c = 0;                                                 // current half of the double-buffered list
    if (tid < nu) {
        const int v = Pk.list[tid];
        const int li = v < -1 ? Pk.mover[-(v + 2)] : v;        // the grid regime leaves rank-encoded holes
        const long long st = __ldg(Pk.rowptr + li);
        m_li[0][tid] = li; m_st[0][tid] = st; m_dg[0][tid] = (int)(__ldg(Pk.rowptr + li + 1) - st);
    }
    __syncthreads();
#if SSLAPB_MID_PROF
    unsigned long long tacc = 0, ts = 0;                       // profiling builds: prof[6] = 1 bidding / 2 merge / 3 compaction part only
#endif
    while (nu > 32 && !done) {
#if SSLAPB_MID_PROF == 1
        ts = sslapb_globaltimer();
#elif SSLAPB_MID_PROF == 4
        tacc += nu;                                            // (+ bids / 1e6 in the same figure)
#endif
        // ---- bidding: warp w sweeps positions w, w + 16, ...; SSLAPB_MID_BATCH at a time: their hot rows are requested
        // together, and the record gather of the next position is issued before the reduction of the current one
        for (int a0 = warp; a0 < nu; a0 += SSLAPB_MID_BATCH * NW) {
            SslapbHotRow row[SSLAPB_MID_BATCH];
#pragma unroll
            for (int k = 0; k < SSLAPB_MID_BATCH; ++k) {
                const int a = a0 + k * NW;
                row[k] = sslapb_load_hot(P, a < nu ? m_li[c][a] : 0, a < nu);   // beyond the list: padding row (column 0)
            }
            SslapbRec256 q = sslapb_ld_rec256(P.rec + row[0].col);
            unsigned undecided = 0u;                           // positions the hot list could not decide (rare): swept below,
#pragma unroll                                                 // out of the pipelined part (no call inside it)
            for (int k = 0; k < SSLAPB_MID_BATCH; ++k) {
                const int a = a0 + k * NW;
                if (a >= nu) break;
                SslapbRec256 qn = q;
                if (k < SSLAPB_MID_BATCH - 1) qn = sslapb_ld_rec256(P.rec + row[k + 1].col);
                SslapbBid b;
                if (!sweep_hot_q(row[k], q, eps, b)) undecided |= 1u << k;
                else {
                    if (lane == 0) { m_j[a] = b.j; m_bid[a] = b.bid; m_po[a] = b.powner; m_pd[a] = b.pdeg; m_ps[a] = b.pstart; }
                    // the owner takes this position when the bid wins: its hot row (4 lines) is on its way for the next round
                    if (lane < 4 && b.powner >= 0)
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char *>(P.hot) +
                                                                         ((unsigned long long)(unsigned)b.powner << 9) + (lane << 7)));
                }
                q = qn;
            }
#if SSLAPB_MID_PROF == 4
            if (lane == 0 && undecided) atomicAdd(&C->prof[6], 1000ull * __popc(undecided));   // prof_ms[6] * 1000 = undecided bids
#endif
            while (undecided) {
                const int a = a0 + (__ffs(undecided) - 1) * NW;
                undecided &= undecided - 1u;
                const long long st = m_st[c][a];
                const SslapbBid b = row_bid_rec<32>(P.cols, P.vals, P.rec, st, st + m_dg[c][a], lane, eps, pmin,
                                                    __ldg(P.rowmax + m_li[c][a]) - spread);
                if (lane == 0) { m_j[a] = b.j; m_bid[a] = b.bid; m_po[a] = b.powner; m_pd[a] = b.pdeg; m_ps[a] = b.pstart; }
            }
        }
        __syncthreads();
#if SSLAPB_MID_PROF == 1
        tacc += sslapb_globaltimer() - ts;
#elif SSLAPB_MID_PROF == 2
        ts = sslapb_globaltimer();
#endif
        // ---- merge + assignment: 2 threads per position (4 when nu <= 128), each scans every 2nd (4th) bid
        const int gs = nu <= SSLAPB_THREADS / 4 ? 2 : 1;       // log2(threads per position)
        const int a = tid >> gs, q = tid & ((1 << gs) - 1);
        const bool act = a < nu;
        const int j = act ? m_j[a] : -1;
        const double bid = act ? m_bid[a] : 0.0;
        bool beaten = false;
        if (j >= 0) {
            // four positions per shared-memory load; entries at or beyond nu are stale (excluded in `other`)
            auto other = [&](int s) { const double ob = m_bid[s]; return s != a && s < nu && (ob > bid || (ob == bid && s < a)); };
            for (int ch = q; ch * 4 < nu; ch += 1 << gs) {
                const int4 v = reinterpret_cast<const int4 *>(m_j)[ch];
                if (v.x == j && other(4 * ch)) beaten = true;
                if (v.y == j && other(4 * ch + 1)) beaten = true;
                if (v.z == j && other(4 * ch + 2)) beaten = true;
                if (v.w == j && other(4 * ch + 3)) beaten = true;
            }
        }
        const unsigned bb = __ballot_sync(SSLAPB_FULL, beaten);
        const bool win = act && j >= 0 && ((bb >> (lane & ~((1 << gs) - 1))) & ((1u << (1 << gs)) - 1u)) == 0u;
        bool hole = false;
        if (act && q == 0) {
            int nv = m_li[c][a], ndg = m_dg[c][a];
            long long nst = m_st[c][a];
            if (win) {
                sslapb_st_rec256(P.rec + j, (unsigned long long)nst, ((unsigned long long)(unsigned)ndg << 32) | (unsigned)nv,
                                 (unsigned long long)__double_as_longlong(bid));   // owner + its row (:418), price (:397)
                P.price[j] = bid;
                nv = m_po[a]; nst = m_ps[a]; ndg = m_pd[a];    // evicted owner takes the slot (:409) or hole (:412)
            }
            m_li[c ^ 1][a] = nv; m_st[c ^ 1][a] = nst; m_dg[c ^ 1][a] = ndg;
            hole = nv < 0;
        }
        ++its; ++rounds;
        if (__shfl_sync(SSLAPB_FULL, (int)(its >= max_iter), 0)) done = 3;
        c ^= 1;
        const int any_hole = __syncthreads_or(hole);
#if SSLAPB_MID_PROF == 2
        tacc += sslapb_globaltimer() - ts;
#elif SSLAPB_MID_PROF == 3
        ts = sslapb_globaltimer();
#endif
        if (any_hole) {
            // ---- push_all_left: the k-th hole left of the new count takes the k-th live entry right of it
            const unsigned hb = __ballot_sync(SSLAPB_FULL, tid < nu && m_li[c][tid] < 0);
            if (lane == 0 && warp < SSLAPB_MID / 32) m_hw[warp] = hb;
            __syncthreads();
            int nholes = 0;
            for (int w = 0; w < SSLAPB_MID / 32; ++w) nholes += __popc(m_hw[w]);
            const int new_nu = nu - nholes;                    // :429
            int src = -1;
            if (tid < new_nu && ((m_hw[warp] >> lane) & 1u)) {
                int k = __popc(m_hw[warp] & ((1u << lane) - 1u));
                for (int w = 0; w < warp; ++w) k += __popc(m_hw[w]);
                for (int w = new_nu >> 5; w < SSLAPB_MID / 32; ++w) {
                    const int lo = w << 5;
                    const unsigned valid = nu >= lo + 32 ? SSLAPB_FULL : (nu <= lo ? 0u : (1u << (nu - lo)) - 1u);
                    const unsigned right = new_nu <= lo ? SSLAPB_FULL : ~((1u << (new_nu - lo)) - 1u);
                    unsigned m = valid & right & ~m_hw[w];
                    const int cnt = __popc(m);
                    if (k < cnt) {
                        for (int r = 0; r < k; ++r) m &= m - 1u;   // drop the k lowest live entries
                        src = lo + __ffs(m) - 1;
                        break;
                    }
                    k -= cnt;
                }
            }
            if (src >= 0) { m_li[c][tid] = m_li[c][src]; m_st[c][tid] = m_st[c][src]; m_dg[c][tid] = m_dg[c][src]; }
            nu = __shfl_sync(SSLAPB_FULL, new_nu, 0);
            __syncthreads();
        }
#if SSLAPB_MID_PROF == 3
        tacc += sslapb_globaltimer() - ts;
#endif
    }
    if (tid < nu) Pk.list[tid] = m_li[c][tid];                 // small_regime (or the epilogue after max_iter) reads it there
    if (tid == 0) {
        C->nu = nu;
        C->its = its;
        if (done) C->done = done;
        C->rounds_mid += rounds;
#if SSLAPB_MID_PROF
        C->prof[6] += tacc;
#else
        C->prof[6] += sslapb_globaltimer() - t0;
#endif
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

#ifdef SSLAPB_SHARDED
// ----------------------------------------------------------------------------------------------------------------------
// Cross-GPU exchange barrier of a row-sharded round (round number k of this communicator, monotone).  Before it every
// rank has stored the bids of ITS bidders into every rank's exchange buffer (plain stores to peer-mapped memory over
// NVLink).  Protocol: local grid barrier arrive (each CTA fences at system scope first, so the release is cumulative over
// its warps' peer stores) -> CTA 0, once all local CTAs have arrived, publishes k in word `rank` of every peer's flag
// block (st.release.sys) -> every CTA waits until all of its OWN flag words show k (ld.acquire.sys on local memory).
// A rank can be at most one sharded round ahead of any other (it needs their flag for round k to leave round k), which is
// why the exchange buffers have two parity halves.  Returns false when the solve was aborted (watchdog).
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cross_barrier(const SslapbAuctionParams &P, SslapbCtrl *c, unsigned nblk, unsigned &epoch, unsigned k)
{
    __shared__ int s_xabort;
    epoch += nblk;
    __syncthreads();
    if (threadIdx.x == 0) {
        int ab = 0;
        const unsigned target = epoch;
        __threadfence_system();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(&c->bar_count) : "memory");
        unsigned polls = 0;
        unsigned long long t0 = sslapb_globaltimer();
        while ((int)(sslapb_ld_acquire_u32(&c->bar_count) - target) < 0) {
            if (++polls < 1024) continue;
            polls = 0;
            if (sslapb_ld_volatile_s32(&c->abort_flag)) { ab = 1; break; }
            if (sslapb_globaltimer() - t0 > P.watchdog_ns) { *(volatile int *)&c->abort_flag = 1; ab = 1; break; }
        }
        unsigned long long tx0 = 0;
        if (blockIdx.x == 0) {
            tx0 = sslapb_globaltimer();
            __threadfence_system();
            for (int r = 0; r < P.nranks; ++r) {
                if (r == P.rank) continue;
                unsigned *pf = reinterpret_cast<unsigned *>(__ldg(P.xtab + 3 * r)) + P.rank;
                asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(pf), "r"(k) : "memory");
            }
        }
        const unsigned *mf = reinterpret_cast<const unsigned *>(__ldg(P.xtab + 3 * P.rank));
        for (int r = 0; r < P.nranks && !ab; ++r) {
            if (r == P.rank) continue;
            polls = 0;
            for (;;) {
                unsigned v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mf + r) : "memory");
                if ((int)(v - k) >= 0) break;
                if (++polls < 1024) continue;
                polls = 0;
                if (sslapb_ld_volatile_s32(&c->abort_flag)) { ab = 1; break; }
                if (sslapb_globaltimer() - t0 > P.watchdog_ns) { *(volatile int *)&c->abort_flag = 1; ab = 1; break; }
            }
        }
        if (blockIdx.x == 0) c->xchg_ns += sslapb_globaltimer() - tx0;
        s_xabort = ab | sslapb_ld_volatile_s32(&c->abort_flag);
    }
    __syncthreads();
    return s_xabort == 0;
}
#endif

// One Jacobi round with the bidders spread over all CTAs of the grid (auction_.pyx:337-430).
struct SslapbScope { int blk, nblk, gwarp, nwarps; bool lead; };   // CTA rank / count, warp rank / count, the one reporting thread
// Returns 0 = aborted, 1 = round done behind a final grid barrier (the caller re-reads the control block), 2 = round done
// WITHOUT that barrier: no bidder found an unowned object (no hole, so the list keeps its size and order and nothing is
// compacted) and no equal bids were seen (tie_flag stays 0) — every CTA knows the next round's inputs (same nu, its + 1)
// and nothing is written after the third barrier that the next round reads.  67 % of the grid rounds of a C3 solve.
__device__ __forceinline__ int spread_round(const SslapbAuctionParams &P, SslapbCtrl *C, const SslapbScope S, int nu, float eps_f,
                                             long long its, long long max_iter, double pmin, double spread, unsigned &bar_epoch,
                                             unsigned &xround, bool hot_mode)
{
    __shared__ int s_red, s_tie;
    __shared__ int s_hpre[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(SSLAPB_FULL, tid >> 5, 0);
    {
    
    const double eps = (double)eps_f;
    unsigned long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0, tp4 = 0, tp5 = 0;
    if (S.lead) tp0 = sslapb_globaltimer();
    int n2nd = 0;
    int *bidj = P.bidj;                                // per list position: object / bid of this round
    double *bidv = P.bidv;
#ifdef SSLAPB_SHARDED
    // Row-sharded round (several GPUs, large frontier): this rank sweeps only the bidders of its own row range and stores
    // each bid into every rank's exchange buffer; after the exchange barrier every rank holds all nu bids and performs the
    // merge, the assignment and the compaction identically (the whole state is replicated, the trajectory is unchanged).
    const bool sharded = P.nranks > 1 && nu > P.t_shard;
    int row_lo = 0, row_hi = 0x7fffffff;
    unsigned xk = 0;
    long long xoff = 0;
    if (sharded) {
        xk = ++xround;
        xoff = (long long)(xk & 1u) * P.xcap;
        row_lo = __ldg(P.rowsplit + P.rank); row_hi = __ldg(P.rowsplit + P.rank + 1);
        bidj = reinterpret_cast<int *>(__ldg(P.xtab + 3 * P.rank + 1)) + xoff;
        bidv = reinterpret_cast<double *>(__ldg(P.xtab + 3 * P.rank + 2)) + xoff;
    }
#endif
    // (1) bidding: warp per list position (auction_.pyx:339-365) + per-object atomicMax merge (:375-385)
    auto emit_bid = [&](int a, int j, double bid) {
#ifdef SSLAPB_SHARDED
                        if (sharded) {                 // lane r stores the bid into rank r's buffer (one instruction for all ranks)
                            if (lane < P.nranks) {
                                reinterpret_cast<int *>(__ldg(P.xtab + 3 * lane + 1))[xoff + a] = j;
                                reinterpret_cast<double *>(__ldg(P.xtab + 3 * lane + 2))[xoff + a] = bid;
                            }
                            return;
                        }
#endif
                        if (lane == 0) {
                            bidj[a] = j;
                            bidv[a] = bid;
                            if (j >= 0) {
                                const unsigned long long key = sslapb_ord64(bid);
                                const unsigned long long old = atomicMax(P.bidkey + j, key);
                                if (old == key) *(volatile int *)&C->tie_flag = 1;
                            } else {
                                *(volatile int *)&C->abort_flag = 2;   // empty row: rejected at CSR build
                            }
                        }
                    };
    // Hot lists (hot.cu): in the first round of an eps-phase (every person bids) every row is tried — that is the probe
    // which decides whether the rest of the phase, tail included, runs in hot form; afterwards only if it does.
    const bool hot_probe = P.hot != nullptr && nu == P.N;
    const bool hot_on = P.hot != nullptr && (hot_mode || hot_probe);
    int n_hot_ok = 0, n_hot_fell = 0;
    // end of round, lead thread: hot form on/off for the rest of the phase.  The probing round decides first (optimistic:
    // its bounds are as fresh as they get), afterwards the share of bids the hot lists failed to decide since the last look.
    auto hot_bookkeeping = [&]() {
        if (P.hot == nullptr) return;
        const long long ok = *(volatile long long *)&C->hot_grid[0], fl = *(volatile long long *)&C->hot_grid[1];
        if (hot_probe) {
            int probed = P.N;
#ifdef SSLAPB_SHARDED
            if (P.nranks > 1 && nu > P.t_shard) probed = __ldg(P.rowsplit + P.rank + 1) - __ldg(P.rowsplit + P.rank);   // own rows only
#endif
            C->hot_mode = ((long long)C->hot_probe_fail * 16 < (long long)probed) ? 1 : 0;
            C->hot_last[0] = ok; C->hot_last[1] = fl;
        } else if (hot_mode) {
            const long long d = (ok + fl) - (C->hot_last[0] + C->hot_last[1]), df = fl - C->hot_last[1];
            if (d >= 64) {
                if (df * 8 > d) C->hot_mode = 0;
                C->hot_last[0] = ok; C->hot_last[1] = fl;
            }
        }
    };
    // (a software pipeline across rows was measured to buy nothing: the sweep is instruction-issue bound)
    for (int a = S.gwarp; a < nu; a += S.nwarps) {
        int v = P.list[a];
        if (v < -1) { v = P.mover[-(v + 2)]; if (lane == 0) P.list[a] = v; }   // hole filled by the last compaction
#ifdef SSLAPB_SHARDED
        if (v < row_lo || v >= row_hi) continue;       // another rank's person (every rank still decodes every hole above)
#endif
        int j = -1; double bid = 0.0;
        if (hot_on) {
            const SslapbBid o = row_bid_hot(P.hot, P.rest, P.price, v, lane, eps);
            j = o.j; bid = o.bid;
            if (j >= 0) ++n_hot_ok; else ++n_hot_fell;
        }
        if (j < 0) {
            const long long st = __ldg(P.rowptr + v), en = __ldg(P.rowptr + v + 1);
            if ((((en + 3) >> 2) - (st >> 2)) <= 32) {
                const SslapbStreamChunk c = sslapb_stream_chunk(P.cols, P.vals, st, en, lane);
                const SslapbBid o = row_bid_pruned(c, P.price, st, en, lane, eps, pmin, __ldg(P.rowmax + v) - spread, n2nd);
                j = o.j; bid = o.bid;
                // every candidate at -inf (objects priced +inf by single-choice bidders): the exact generic sweep decides
                if (j < 0) row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid);
            } else {                                   // long row: multi-trip sweep, same bound pruning
                row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid, pmin, __ldg(P.rowmax + v) - spread);
            }
        }
        emit_bid(a, j, bid);
    }
    if (hot_on && lane == 0 && (n_hot_ok | n_hot_fell)) {
        atomicAdd((unsigned long long *)&C->hot_grid[0], (unsigned long long)n_hot_ok);
        atomicAdd((unsigned long long *)&C->hot_grid[1], (unsigned long long)n_hot_fell);
        if (hot_probe) atomicAdd(&C->hot_probe_fail, n_hot_fell);
    }
    if (n2nd && lane == 0) atomicAdd((unsigned long long *)&C->prune_second_pass, (unsigned long long)n2nd);
    if (S.lead) tp1 = sslapb_globaltimer();
#ifdef SSLAPB_SHARDED
    if (sharded) {
        if (!cross_barrier(P, C, (unsigned)S.nblk, bar_epoch, xk)) return 0;
        // merge (:375-385) over ALL nu bids, identically on every rank
        for (int a = blockIdx.x * blockDim.x + tid; a < nu; a += S.nblk * blockDim.x) {
            const int j = bidj[a];
            if (j >= 0) {
                const unsigned long long key = sslapb_ord64(bidv[a]);
                const unsigned long long old = atomicMax(P.bidkey + j, key);
                if (old == key) *(volatile int *)&C->tie_flag = 1;
            } else {
                *(volatile int *)&C->abort_flag = 2;   // empty row: rejected at CSR build
            }
        }
    }
#endif
    if (!grid_barrier(C, (unsigned)S.nblk, bar_epoch, P.watchdog_ns)) return 0;
    if (S.lead) {
        tp2 = sslapb_globaltimer();
        // in-situ full-frontier bidding step (every person bids: first round of an eps-phase), up to the barrier that
        // proves every CTA has finished its rows — no launch ramp, no tail outside the barrier
        if (nu == P.N) { C->sweep_ns[0] += tp2 - tp0; C->sweep_ns[1] += 1; }
    }
    if (tid == 0) s_tie = *(volatile int *)&C->tie_flag;
    __syncthreads();
    const int tie = __shfl_sync(SSLAPB_FULL, s_tie, 0);
    const int L = (nu + S.nblk - 1) / S.nblk;    // each CTA owns a contiguous chunk of positions
    const int lo = min(nu, S.blk * L), hi = min(nu, lo + L);
    if (tie) {                                         // (1b) equal best bids: earliest list position wins (:379)
        for (int a = lo + tid; a < hi; a += blockDim.x) {
            const int j = bidj[a];
            if (P.bidkey[j] == sslapb_ord64(bidv[a])) atomicMin(P.winpos + j, a);
        }
        if (!grid_barrier(C, (unsigned)S.nblk, bar_epoch, P.watchdog_ns)) return 0;
    }
    // (2) assignment (:394-427), by the winner's own list position
    int myholes = 0;
    for (int a = lo + tid; a < hi; a += blockDim.x) {
        const int j = bidj[a];
        const double bid = bidv[a];
        const bool win = (P.bidkey[j] == sslapb_ord64(bid)) && (!tie || P.winpos[j] == a);
        if (win) {
            const int i = P.list[a];
            const int prev = P.rec[j].owner;
            const long long rst = __ldg(P.rowptr + i);
            const int rdg = (int)(__ldg(P.rowptr + i + 1) - rst);
            sslapb_st_rec256(P.rec + j, (unsigned long long)rst, ((unsigned long long)(unsigned)rdg << 32) | (unsigned)i,
                             (unsigned long long)__double_as_longlong(bid));
            P.price[j] = bid;
            if (prev < 0) ++myholes;                   // (person_to_object: see rebuild_p2o)
            P.list[a] = prev;                          // evicted owner takes the slot, or -1 = hole
            P.bidkey[j] = 0ull;
            if (tie) P.winpos[j] = 0x7fffffff;
        }
    }
    // Frontiers of at most one position per thread of a CTA (92 % of the grid rounds of a C3 solve): no per-CTA hole
    // counts — after the barrier EVERY CTA reads the whole list (2 KB) and counts the holes itself, so all of them know H
    // (and take the no-hole exit together), and CTA 0 alone compacts with one block scan.
    const bool small_round = nu <= SSLAPB_THREADS;
    if (!small_round) {
        if (tid == 0) s_red = 0;
        __syncthreads();
        if (myholes) atomicAdd(&s_red, myholes);
        __syncthreads();
        if (tid == 0) P.hole_count[S.blk] = s_red;
    }
    if (S.lead) tp3 = sslapb_globaltimer();
    if (!grid_barrier(C, (unsigned)S.nblk, bar_epoch, P.watchdog_ns)) return 0;
    if (S.lead) tp4 = sslapb_globaltimer();
    // (3) push_all_left (:137-162): k-th hole left of new_nu <- k-th live entry right of it
    int sr_v = 0, sr_total = 0;
    bool sr_hole = false;
    if (small_round) {
        if (tid < nu) { sr_v = P.list[tid]; sr_hole = sr_v < 0; }
        sr_total = __syncthreads_count(sr_hole);       // ONE barrier instruction: every CTA knows H (most rounds have none)
    } else
    if (warp == 0) {                                   // per-CTA hole counts -> total, my prefix, prefix of the split chunk
        int ht = 0, hp = 0;
        for (int b = lane; b < S.nblk; b += 32) {
            const int h = P.hole_count[b];
            ht += h;
            if (b < S.blk) hp += h;
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
    if (!small_round) __syncthreads();
    const int H = small_round ? sr_total : __shfl_sync(SSLAPB_FULL, s_hpre[2], 0);
    const int hpre = small_round ? 0 : __shfl_sync(SSLAPB_FULL, s_hpre[1], 0);
    if (H == 0 && !tie && !hot_probe && its + 2 < max_iter) {  // no hole, no tie: nothing to compact, nothing to reset
        if (S.lead) {
            C->its = its + 1;
            C->rounds_grid += 1;
            C->rounds_nohole += 1;
            hot_bookkeeping();
            tp5 = sslapb_globaltimer();
#ifdef SSLAPB_SHARDED
            if (sharded) { C->rounds_sharded += 1; C->sharded_ns += tp5 - tp0; }
#endif
            C->prof[0] += tp1 - tp0; C->prof[1] += tp3 - tp2; C->prof[2] += tp5 - tp4;
            C->prof[7] += (tp2 - tp1) + (tp4 - tp3);
        }
        __syncthreads();                               // s_hpre / s_red / s_tie are rewritten by the next round
        return 2;
    }
    const int new_nu = nu - H;
    if (small_round) {
        if (S.blk == 0 && H > 0) {                     // thread t = position t; holes before position new_nu = Hsplit
            int tot2;
            const int sr_before = block_excl_scan_flag(sr_hole, tot2);
            if (tid == new_nu) s_hpre[0] = sr_before;  // (new_nu < nu <= blockDim whenever there is a hole)
            __syncthreads();
            const int Hsplit = new_nu < nu ? s_hpre[0] : H;
            if (tid < nu) {
                if (tid < new_nu) {
                    if (sr_hole) P.list[tid] = -(sr_before + 2);   // rank-encoded; decoded by the next reader
                } else if (!sr_hole) {
                    P.mover[(tid - new_nu) - (sr_before - Hsplit)] = sr_v;
                }
            }
        }
    } else {
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
    }
    if (S.lead) {
        C->nu = new_nu;
        C->its = its + 1;
        C->tie_flag = 0;
        C->rounds_grid += 1;
        hot_bookkeeping();
        if (its + 1 >= max_iter) C->done = 3;
        tp5 = sslapb_globaltimer();
#ifdef SSLAPB_SHARDED
        if (sharded) { C->rounds_sharded += 1; C->sharded_ns += tp5 - tp0; }
#endif
    }
    if (!grid_barrier(C, (unsigned)S.nblk, bar_epoch, P.watchdog_ns)) return 0;
    if (S.lead) {
        C->prof[0] += tp1 - tp0; C->prof[1] += tp3 - tp2; C->prof[2] += tp5 - tp4;
        C->prof[7] += (tp2 - tp1) + (tp4 - tp3) + (sslapb_globaltimer() - tp5);
    }
    }
    return 1;
}

#define GB() do { if (!grid_barrier(C, nblk, bar_epoch, P.watchdog_ns)) return; } while (0)

// person_to_object (auction_.pyx:231) from object_to_person: every owned object names its person.  Called by all CTAs right
// after a grid barrier; the caller follows it with another one.  `clear` first resets every person (needed when some are
// unassigned: max_iter hit inside a phase); with a full assignment every entry is overwritten anyway.
__device__ __forceinline__ void rebuild_p2o_owners(const SslapbAuctionParams &P, int gtid, int nthreads)
{
    for (int j = gtid; j < P.M; j += nthreads) {
        const int o = P.rec[j].owner;
        if (o >= 0) P.p2o[o] = j;
    }
}

#ifdef SSLAPB_LONG_ROWS
#define sslapb_auction_kernel sslapb_auction_kernel_long   // second instance of the kernel, see auction_long.cu
#endif
#ifdef SSLAPB_SHARDED
#define sslapb_auction_kernel sslapb_auction_kernel_sharded   // row-sharded multi-GPU instance, see auction_sharded.cu
#endif
__global__ void __launch_bounds__(SSLAPB_THREADS, 1) sslapb_auction_kernel(const __grid_constant__ SslapbAuctionParams P)
{
    SslapbCtrl *C = P.ctrl;
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(SSLAPB_FULL, tid >> 5, 0);
    const unsigned nblk = gridDim.x;
    const int wpc = blockDim.x >> 5;
    const int gwarp = blockIdx.x * wpc + warp;
    const int nwarps = nblk * wpc;
    const int gtid = blockIdx.x * blockDim.x + tid;
    const int nthreads = nblk * blockDim.x;
    __shared__ struct { int nu, done, nred, tie, hot; float eps; long long its, max_iter; unsigned long long pmin0, pmin1, pmax; } s_top;
    unsigned bar_epoch = 0;                                    // barriers passed so far * #CTAs (wraps harmlessly)
    unsigned xround = P.xround_base;                           // row-sharded rounds of this communicator so far (all CTAs agree)

    if (gtid == 0) C->t_begin = sslapb_globaltimer();

    for (;;) {
        // ---- loop top: every CTA arrives here right after a grid barrier; the control block is stable.  ONE thread per
        // CTA reads it (all 75k threads loading the same line serialise in its L2 slice) and shares it through smem.
        if (tid == 0) {
            s_top.nu = *(volatile int *)&C->nu;
            s_top.done = *(volatile int *)&C->done;
            s_top.eps = *(volatile float *)&C->eps;
            s_top.its = *(volatile long long *)&C->its;
            s_top.max_iter = *(volatile long long *)&C->max_iter;
            s_top.nred = *(volatile int *)&C->nreductions;
            s_top.hot = *(volatile int *)&C->hot_mode;
            s_top.pmin0 = *(volatile unsigned long long *)&C->pmin_key[0];
            s_top.pmin1 = *(volatile unsigned long long *)&C->pmin_key[1];
            s_top.pmax = *(volatile unsigned long long *)&C->pmax_key;
        }
        __syncthreads();
        int nu = __shfl_sync(SSLAPB_FULL, s_top.nu, 0);        // shuffles: warp-uniform for the compiler
        int done = __shfl_sync(SSLAPB_FULL, s_top.done, 0);
        const float eps_f = s_top.eps;
        const long long its = s_top.its;
        const long long max_iter = s_top.max_iter;
        const int phase_slot = s_top.nred & 1;
        const bool hot_mode = __shfl_sync(SSLAPB_FULL, s_top.hot, 0) != 0;
        // price bounds for the pruned sweep, taken at the start of the eps-phase: pmin stays a valid lower bound all phase
        // long (prices never decrease); the spread is only a heuristic for which candidates to gather first
        const double pmin = sslapb_key2double(phase_slot ? s_top.pmin1 : s_top.pmin0);
        double spread = (sslapb_key2double(s_top.pmax) - pmin) + 2.0 * (double)eps_f;
        if (!(spread < 1.7e308)) spread = __longlong_as_double(0x7ff0000000000000ll);   // inf / NaN (infinite prices): no pruning
        __syncthreads();                                       // s_top is rewritten only after every thread has read it
        if (done) break;

        // mid regime (CTA 0 alone, block barriers) only in phases whose bids the hot lists decide: a bid is then one 512-byte
        // row + one record gather, and 16 warps finish a round of up to t_mid bidders before three grid barriers would
        const bool hot_small = hot_mode && P.hot != nullptr;
        const int t_lim = (hot_small && P.t_mid > P.t_small) ? P.t_mid : P.t_small;
        if (nu > t_lim) {
            // ================================ grid regime: one round ================================
            const SslapbScope S = {(int)blockIdx.x, (int)nblk, gwarp, nwarps, gtid == 0};
            long long its_l = its;
            for (;;) {
                const int r = spread_round(P, C, S, nu, eps_f, its_l, max_iter, pmin, spread, bar_epoch, xround, hot_mode);
                if (r == 0) return;
                if (r == 1) break;
                ++its_l;                                       // r == 2: same frontier, next round straight away
            }
        } else if (nu <= 32) {
            // ================================ warp-list regimes: CTA 0 finishes the phase ================================
            if (blockIdx.x == 0) small_regime(P, C, nu, eps_f, its, max_iter, pmin, spread, hot_small);
            GB();
        } else {
            // ================================ mid regime: CTA 0 runs rounds until nu <= 32 ================================
            // (the LAST branch of the chain, so that its code lies behind small_regime's in the binary, with its own barrier
            // and a trip through the loop top before small_regime takes over: the few-bidder loops are sensitive to the code
            // around their call site and to their placement)
            if (blockIdx.x == 0) mid_regime(P, C, nu, eps_f, its, max_iter, pmin, spread);
            GB();
        }

        if (tid == 0) { s_top.nu = *(volatile int *)&C->nu; s_top.done = *(volatile int *)&C->done; }
        __syncthreads();
        nu = __shfl_sync(SSLAPB_FULL, s_top.nu, 0);
        done = __shfl_sync(SSLAPB_FULL, s_top.done, 0);
        __syncthreads();
        if (done || nu != 0) continue;

        // ================================ full assignment reached: terminate() / eps-scaling (:275-292) ================================
        {
            unsigned long long te0 = 0;
            if (gtid == 0) te0 = sslapb_globaltimer();
            const float teps = *(volatile float *)&C->target_eps;
            const double eps_t = (double)teps, tol = *(volatile double *)&C->tol;
            bool viol = false;
            rebuild_p2o_owners(P, gtid, nthreads);             // nu == 0: every person owns exactly one object
            GB();
            for (int i = gwarp; i < P.N; i += nwarps) {        // eCE_satisfied(target_eps), :443-485
                const int j = P.p2o[i];
                const long long st = __ldg(P.rowptr + i), en = __ldg(P.rowptr + i + 1);
                double vmax, choice, csum;
                if (P.hot) {
                    // the sweep gathers every price of the row anyway: it also renews the bound of everything outside the
                    // person's hot list, for the next eps-phase, and touches the hot row (L2-persisting window, api.cu)
                    double rst;
                    row_ece<true>(P.cols, P.vals, P.price, st, en, lane, j, vmax, choice, csum, P.hthr[i], &rst);
                    if (lane == 0) P.rest[i] = rst;
                    const int4 touch = __ldg(reinterpret_cast<const int4 *>(P.hot) + (long long)i * 32 + lane);
                    asm volatile("" :: "r"(touch.x));
                } else {
                    row_ece(P.cols, P.vals, P.price, st, en, lane, j, vmax, choice, csum);
                }
                const double lhs = (choice - P.price[j]) + tol;
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
                unsigned long long kmin = ~0ull, kmax = 0ull;  // price range: the next phase's pruning bounds
                for (int j = gtid; j < P.M; j += nthreads) {
                    P.rec[j].owner = -1;
                    const unsigned long long k = sslapb_ord64(P.price[j]);
                    kmin = k < kmin ? k : kmin;
                    kmax = k > kmax ? k : kmax;
                }
                const unsigned mh = __reduce_min_sync(SSLAPB_FULL, (unsigned)(kmin >> 32));
                const unsigned ml = __reduce_min_sync(SSLAPB_FULL, (unsigned)(kmin >> 32) == mh ? (unsigned)kmin : 0xffffffffu);
                const unsigned xh = __reduce_max_sync(SSLAPB_FULL, (unsigned)(kmax >> 32));
                if (lane == 0) {
                    atomicMin(&C->pmin_key[phase_slot ^ 1], ((unsigned long long)mh << 32) | ml);
                    const unsigned long long kx = ((unsigned long long)xh << 32) | 0xffffffffull;   // rounded up: heuristic only
                    atomicMax(&C->pmax_key, kx > 0xfff0000000000000ull ? 0xfff0000000000000ull : kx);   // never beyond +inf
                }
            }
            GB();
            if (gtid == 0) {
                if (stop_opt) { C->done = 1; C->ece_final = 1; }
                else if (stop_eps) { C->done = 2; C->ece_final = 0; }
                else {
                    C->eps = eps_now * C->theta;               // float32 product (:283)
                    C->pmin_key[phase_slot] = ~0ull;           // recycled two phases from now
                    C->nreductions += 1;
                    C->nu = P.N;
                    C->hot_mode = 0; C->hot_probe_fail = 0;    // the next phase probes again
                }
                C->ece_viol = 0;
                C->prof[5] += sslapb_globaltimer() - te0;
            }
            GB();
        }
    }

    // ================================ epilogue: meta['eCE'] (:297) and per-person chosen values (get_obj) ================================
    {
        const int nu = *(volatile int *)&C->nu;
        const bool need_ece = (*(volatile int *)&C->ece_final < 0) && nu == 0;   // max_iter hit on a full assignment
        const double eps_t = (double)(*(volatile float *)&C->target_eps), tol = *(volatile double *)&C->tol;
        bool viol = false;
        for (int i = gtid; i < P.N; i += nthreads) P.p2o[i] = -1;              // the final person_to_object (`sol`)
        GB();
        rebuild_p2o_owners(P, gtid, nthreads);
        GB();
        for (int i = gwarp; i < P.N; i += nwarps) {
            const int j = P.p2o[i];
            double csum = 0.0;
            if (j >= 0) {
                const long long st = __ldg(P.rowptr + i), en = __ldg(P.rowptr + i + 1);
                double vmax, choice;
                row_ece(P.cols, P.vals, P.price, st, en, lane, j, vmax, choice, csum);
                if (need_ece && ((choice - P.price[j]) + tol < vmax - eps_t)) viol = true;
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

// Cooperative launch of this translation unit's kernel instance.
// coop = 0: an ordinary launch.  The driver runs one cooperative kernel at a time, so several persistent kernels that must
// make progress TOGETHER on one GPU (the virtual-rank test of the row-sharded solve: K kernels of sms/K CTAs each, one CTA
// per SM by construction) are launched without the cooperative attribute; co-residency then rests on grid <= free SMs, and
// the barrier watchdog turns a violation into an error instead of a hang.
static cudaError_t launch_persistent(const SslapbAuctionParams *P, int grid, cudaStream_t stream, int coop = 1)
{
    void *args[] = {(void *)P};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(SSLAPB_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = coop ? 1 : 0;
    return cudaLaunchKernelExC(&cfg, (const void *)sslapb_auction_kernel, args);
}
#if defined(SSLAPB_LONG_ROWS)
// (the state is initialised by sslapb_launch_auction, auction.cu)
extern "C" cudaError_t sslapb_launch_auction_long(const SslapbAuctionParams *P, int grid, cudaStream_t stream)
{
    return launch_persistent(P, grid, stream);
}
extern "C" int sslapb_coop_row_entries() { return 4 * SSLAPB_COOP_CHUNKS - 3; }
#elif defined(SSLAPB_SHARDED)
extern "C" cudaError_t sslapb_launch_auction_sharded(const SslapbAuctionParams *P, int grid, int coop, cudaStream_t stream)
{
    return launch_persistent(P, grid, stream, coop);
}
// nnz-balanced contiguous row split (the device-side counterpart of cumulative_idxs' row partition, auction_.pyx:33-48):
// boundary r = first row whose CSR offset reaches r * nnz / parts (binary search over rowptr, one thread per boundary).
__global__ void sslapb_row_split_kernel(const long long *__restrict__ rowptr, int N, int parts, int *__restrict__ split)
{
    const int r = threadIdx.x;
    if (r > parts) return;
    if (r == parts) { split[r] = N; return; }
    const long long nnz = rowptr[N], target = (nnz / parts) * r + ((nnz % parts) * r) / parts;
    int lo = 0, hi = N;                                        // first row with rowptr[row] >= target
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rowptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    split[r] = lo;
}
extern "C" cudaError_t sslapb_launch_row_split(const long long *rowptr, int N, int parts, int *split, cudaStream_t stream)
{
    sslapb_row_split_kernel<<<1, 32, 0, stream>>>(rowptr, N, parts, split);
    return cudaGetLastError();
}
#else
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
    const bool prune = (merge & 2) == 0;                       // bit 1 of `merge` switches the bound pruning off (A/B measurement)
    const double pmin = sslapb_key2double(P.ctrl->pmin_key[0]);
    double spread = sslapb_key2double(P.ctrl->pmax_key) - pmin;
    if (!(spread < 1.7e308)) spread = __longlong_as_double(0x7ff0000000000000ll);
    int n2nd = 0;
    merge &= 1;
    // one row per warp at a time; the NEXT row's offsets and row maximum are requested one iteration ahead (one of the
    // three dependent round trips per row: offsets -> entries -> prices)
    int a = gwarp;
    if (a >= nb) return;
    int i = bidders ? bidders[a] : a;
    long long st = __ldg(P.rowptr + i), en = __ldg(P.rowptr + i + 1);
    double rmax = __ldg(P.rowmax + i);
    for (;;) {
        const int an = a + nwarps;
        int in = 0;
        long long stn = 0, enn = 0;
        double rmaxn = 0.0;
        if (an < nb) {
            in = bidders ? bidders[an] : an;
            stn = __ldg(P.rowptr + in); enn = __ldg(P.rowptr + in + 1);
            rmaxn = __ldg(P.rowmax + in);
        }
        int j; double bid;
        if ((((en + 3) >> 2) - (st >> 2)) <= 32) {
            const SslapbStreamChunk c = sslapb_stream_chunk(P.cols, P.vals, st, en, lane);
            const SslapbBid o = row_bid_pruned(c, P.price, st, en, lane, eps, pmin, prune ? rmax - spread : SSLAPB_NEG_INF, n2nd);
            j = o.j; bid = o.bid;
            if (j < 0) row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid);   // all candidates at -inf / unproven
        } else {
            row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid, pmin, prune ? rmax - spread : SSLAPB_NEG_INF);
        }
        if (lane == 0) {
            P.bidj[a] = j;
            P.bidv[a] = bid;
            if (merge && j >= 0) atomicMax(P.bidkey + j, sslapb_ord64(bid));
        }
        if (an >= nb) break;
        a = an; st = stn; en = enn; rmax = rmaxn;
    }
    if (n2nd && lane == 0) atomicAdd((unsigned long long *)&P.ctrl->prune_second_pass, (unsigned long long)n2nd);
}

// ----------------------------------------------------------------------------------------------------------------------
// Stand-alone bidding sweep in hot form: every bidder is decided from its hot list (512 B + one price gather per lane) when
// that is provably exact, else by the full-row sweep of the kernel above — the same two-step the grid regime of the
// persistent kernel runs from the third eps-phase on.  Results are bit-identical to the full-row kernels.
// ----------------------------------------------------------------------------------------------------------------------
// Pass 1: hot lists only — light enough (<= 32 registers) for 64 resident warps per SM, which is what a latency-bound
// gather kernel wants.  Bidders the hot list cannot decide are appended to a redo list (P.mover, counter in ctrl).
__global__ void __launch_bounds__(1024, 2) sslapb_bid_sweep_hot_kernel(SslapbAuctionParams P, const int *__restrict__ bidders, int nb,
                                                                      float eps_f, int merge)
{
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int gwarp = blockIdx.x * wpc + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * wpc;
    const double eps = (double)eps_f;
    merge &= 1;
    // (requesting the next bidder's hot row one iteration ahead was measured and dropped: 36.9 against 34.7 us)
    for (int a = gwarp; a < nb; a += nwarps) {
        const int i = bidders ? __ldg(bidders + a) : a;
        const SslapbBid o = row_bid_hot(P.hot, P.rest, P.price, i, lane, eps);
        if (lane == 0) {
            P.bidj[a] = o.j;
            P.bidv[a] = o.bid;
            if (o.j >= 0) { if (merge) atomicMax(P.bidkey + o.j, sslapb_ord64(o.bid)); }
            else P.mover[atomicAdd(&P.ctrl->hot_probe_fail, 1)] = a;
        }
    }
    // programmatic dependent launch: the redo pass may be scheduled while this grid drains (it waits for our writes itself)
    asm volatile("griddepcontrol.launch_dependents;");
}

// Pass 2: the redo list through the full-row sweep (bound-pruned, exact) — the per-row kernel's body.
__global__ void __launch_bounds__(512, 2) sslapb_bid_sweep_redo_kernel(SslapbAuctionParams P, const int *__restrict__ bidders,
                                                                      float eps_f, int merge)
{
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int gwarp = blockIdx.x * wpc + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * wpc;
    asm volatile("griddepcontrol.wait;" ::: "memory");        // the pass that wrote the redo list has completed and is visible
    const int nredo = *(volatile int *)&P.ctrl->hot_probe_fail;
    if (gwarp >= nredo) return;
    const double eps = (double)eps_f;
    const bool prune = (merge & 2) == 0;
    const double pmin = sslapb_key2double(P.ctrl->pmin_key[0]);
    double spread = sslapb_key2double(P.ctrl->pmax_key) - pmin;
    if (!(spread < 1.7e308)) spread = __longlong_as_double(0x7ff0000000000000ll);
    int n2nd = 0;
    merge &= 1;
    for (int k = gwarp; k < nredo; k += nwarps) {
        const int a = P.mover[k];
        const int i = bidders ? __ldg(bidders + a) : a;
        const long long st = __ldg(P.rowptr + i), en = __ldg(P.rowptr + i + 1);
        const double thr = prune ? __ldg(P.rowmax + i) - spread : SSLAPB_NEG_INF;
        int j; double bid;
        if ((((en + 3) >> 2) - (st >> 2)) <= 32) {
            const SslapbStreamChunk c = sslapb_stream_chunk(P.cols, P.vals, st, en, lane);
            const SslapbBid f = row_bid_pruned(c, P.price, st, en, lane, eps, pmin, thr, n2nd);
            j = f.j; bid = f.bid;
            if (j < 0) row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid);
        } else {
            row_bid<32>(P.cols, P.vals, P.price, st, en, lane, eps, j, bid, pmin, thr);
        }
        if (lane == 0) {
            P.bidj[a] = j;
            P.bidv[a] = bid;
            if (merge && j >= 0) atomicMax(P.bidkey + j, sslapb_ord64(bid));
        }
    }
    if (lane == 0) {
        if (n2nd) atomicAdd((unsigned long long *)&P.ctrl->prune_second_pass, (unsigned long long)n2nd);
        if (gwarp == 0) P.ctrl->hot_grid[1] += nredo;
    }
}

// rest[i] = upper bound of a_ik - p_k over the entries outside person i's hot list, at the current prices (what every
// eps-CS sweep of the persistent kernel leaves behind); stand-alone sweep only
__global__ void __launch_bounds__(256) sslapb_hot_rest_kernel(SslapbAuctionParams P)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    if (gwarp == 0 && lane == 0) P.ctrl->hot_grid[1] = 0;
    for (int i = gwarp; i < P.N; i += nwarps) {
        double vmax, choice, csum, rst;
        row_ece<true>(P.cols, P.vals, P.price, __ldg(P.rowptr + i), __ldg(P.rowptr + i + 1), lane, -1, vmax, choice, csum, P.hthr[i], &rst);
        if (lane == 0) P.rest[i] = rst;
    }
}

extern "C" cudaError_t sslapb_launch_hot_rest(const SslapbAuctionParams *P, int sms, cudaStream_t stream)
{
    sslapb_hot_rest_kernel<<<sms * 8, 256, 0, stream>>>(*P);
    return cudaGetLastError();
}

extern "C" cudaError_t sslapb_launch_bid_sweep_redo(const SslapbAuctionParams *P, const int *bidders, float eps, int merge,
                                                    int grid, cudaStream_t stream)
{
    // launched with programmatic stream serialization: its CTAs may be scheduled while the producing pass drains
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid * 2); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, sslapb_bid_sweep_redo_kernel, *P, bidders, eps, merge);
}

extern "C" cudaError_t sslapb_launch_bid_sweep_hot(const SslapbAuctionParams *P, const int *bidders, int nb, float eps,
                                                   int merge, int grid, cudaStream_t stream)
{
    // the redo counter (ctrl->hot_probe_fail) is zeroed by the caller before every launch pair
    sslapb_bid_sweep_hot_kernel<<<grid * 2, 1024, 0, stream>>>(*P, bidders, nb, eps, merge);
    sslapb_bid_sweep_redo_kernel<<<grid * 2, 512, 0, stream>>>(*P, bidders, eps, merge);
    return cudaGetLastError();
}

// Initial state of a solve (AuctionSolver.__init__, auction_.pyx:220-261).
// warm != 0: price[] already holds the caller's start prices (sslapb_set_prices) — the reference starts from zeros (:220)
__global__ void sslapb_auction_init_kernel(SslapbAuctionParams P, int warm)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
    for (int i = gtid; i < P.N; i += n) {
        P.p2o[i] = -1; P.list[i] = i;
        if (P.rest) P.rest[i] = __longlong_as_double(0x7ff0000000000000ll);   // no bound yet: the first eps-CS sweep writes it
    }
    for (int j = gtid; j < P.M; j += n) {
        const double p0 = warm ? P.price[j] : 0.0;
        SslapbObjRec r; r.start = 0; r.owner = -1; r.deg = 0; r.price = p0; r.pad = 0;
        P.rec[j] = r; P.price[j] = p0; P.bidkey[j] = 0ull; P.winpos[j] = 0x7fffffff;
    }
}

extern "C" cudaError_t sslapb_launch_auction_long(const SslapbAuctionParams *P, int grid, cudaStream_t stream);
// Three instances of the persistent kernel: this one (lean), auction_long.cu (the longest row exceeds
// sslapb_coop_row_entries() entries) and auction_sharded.cu (row-sharded multi-GPU solve, nranks > 1).
extern "C" cudaError_t sslapb_launch_auction_sharded(const SslapbAuctionParams *P, int grid, int coop, cudaStream_t stream);
extern "C" cudaError_t sslapb_launch_auction_init(const SslapbAuctionParams *P, int grid, int warm, cudaStream_t stream)
{
    sslapb_auction_init_kernel<<<grid, 1024, 0, stream>>>(*P, warm);
    return cudaGetLastError();
}
extern "C" cudaError_t sslapb_launch_auction(const SslapbAuctionParams *P, int grid, int long_rows, int coop, cudaStream_t stream)
{
    if (long_rows) return sslapb_launch_auction_long(P, grid, stream);
    if (P->nranks > 1) return sslapb_launch_auction_sharded(P, grid, coop, stream);
    return launch_persistent(P, grid, stream);
}

extern "C" cudaError_t sslapb_auction_grid_size(int device, int *grid)
{
    int per_sm = 0, sms = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sslapb_auction_kernel, SSLAPB_THREADS, 0);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid = sms;                                               // one persistent CTA per SM
    return cudaSuccess;
}

// min / max of the current prices into ctrl (stand-alone sweep only; the persistent kernel tracks them itself)
__global__ void sslapb_price_bounds_kernel(SslapbAuctionParams P, int reset)
{
    SslapbCtrl *C = P.ctrl;
    if (reset) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { C->pmin_key[0] = ~0ull; C->pmin_key[1] = ~0ull; C->pmax_key = 0ull; C->prune_second_pass = 0; }
        return;
    }
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
    unsigned long long kmin = ~0ull, kmax = 0ull;
    for (int j = gtid; j < P.M; j += n) {
        const unsigned long long k = sslapb_ord64(P.price[j]);
        kmin = k < kmin ? k : kmin; kmax = k > kmax ? k : kmax;
    }
    atomicMin(&C->pmin_key[0], kmin);
    atomicMax(&C->pmax_key, kmax);
}

extern "C" cudaError_t sslapb_launch_price_bounds(const SslapbAuctionParams *P, cudaStream_t stream)
{
    sslapb_price_bounds_kernel<<<1, 32, 0, stream>>>(*P, 1);
    sslapb_price_bounds_kernel<<<64, 256, 0, stream>>>(*P, 0);
    return cudaGetLastError();
}

extern "C" cudaError_t sslapb_launch_bid_sweep(const SslapbAuctionParams *P, const int *bidders, int nb, float eps,
                                               int merge, int grid, cudaStream_t stream)
{
    sslapb_bid_sweep_kernel<<<grid, 1024, 0, stream>>>(*P, bidders, nb, eps, merge);
    return cudaGetLastError();
}

// ======================================================================================================================
// Batched small problems (BASELINE.json configs[4]: thousands of independent 512 x 512 problems): ONE WARP PER PROBLEM.
//
// Every auction is a chain of dependent rounds, so a single small problem cannot use more than a sliver of the GPU;
// throughput comes from running thousands of those chains at once (148 SMs x 32 resident warps = 4736 problems in
// flight).  Each warp executes the reference's round literally on its own problem — bidding over the unassigned list
// (row sweep by the 32 lanes, auction_.pyx:339-365), merge by atomicMax on the order-preserving bid + list-position
// tie-break (:375-385), winner-driven assignment (:394-427), push_all_left in list order (:137-162), eps-scaling and the
// eps-CS test (:268-309, :443-485) — so `sol`, `its` and every meta key are bit-identical to a reference call per problem.
// The problems are stored block-diagonally in ONE CSR (global row/column ids), built by the ordinary ingest pass.
// ======================================================================================================================
#include "batch.cuh"

__device__ __forceinline__ bool batch_ece(const SslapbBatchParams &B, long long r0, int N, int lane, float teps)
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

__global__ void __launch_bounds__(128) sslapb_auction_batch_kernel(SslapbBatchParams B)
{
    const int lane = threadIdx.x & 31;
    const int p = __shfl_sync(SSLAPB_FULL, (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), 0);
    if (p >= B.P) return;
    const long long r0 = B.rowoff[p], c0 = B.coloff[p];
    const int N = (int)(B.rowoff[p + 1] - r0), M = (int)(B.coloff[p + 1] - c0);
    // ---- AuctionSolver.__init__ (auction_.pyx:220-261)
    for (int j = lane; j < M; j += 32) { B.price[c0 + j] = 0.0; B.owner[c0 + j] = -1; B.bestkey[c0 + j] = 0ull; B.winpos[c0 + j] = 0x7fffffff; }
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
    int dummy2nd = 0;
    for (;;) {
        const double epsd = (double)eps;
        // ---- bidding (:339-365): the unassigned list in order, one row sweep at a time
        for (int n = 0; n < nu; ++n) {
            const int i = B.list[r0 + n];
            const long long st = __ldg(B.rowptr + i), en = __ldg(B.rowptr + i + 1);
            int j; double bid;
            if ((((en + 3) >> 2) - (st >> 2)) <= 32) {
                const SslapbStreamChunk c = sslapb_stream_chunk(B.cols, B.vals, st, en, lane);
                const SslapbBid o = row_bid_pruned(c, B.price, st, en, lane, epsd, 0.0, SSLAPB_NEG_INF, dummy2nd);
                j = o.j; bid = o.bid;
                if (j < 0) row_bid<32>(B.cols, B.vals, B.price, st, en, lane, epsd, j, bid);   // all candidates at -inf
            } else {
                row_bid<32>(B.cols, B.vals, B.price, st, en, lane, epsd, j, bid);
            }
            if (lane == 0) { B.bidj[r0 + n] = j; B.bidv[r0 + n] = bid; }
        }
        __syncwarp();
        // ---- merge (:375-385): per-object maximum of the order-preserving bid, earliest list position on equal bids
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
        // ---- assignment (:394-427)
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
                    const int prev = B.owner[j];
                    B.price[j] = bid;
                    B.owner[j] = i;
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
        // ---- push_all_left (:137-162): k-th hole left of the new count <- k-th live entry right of it
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
        ++its;
        last_opt = -1;
        // ---- terminate() / eps-scaling (:275-292)
        if (its >= B.max_iter) { stop = 3; break; }
        if (nu == 0) {
            last_opt = batch_ece(B, r0, N, lane, target) ? 1 : 0;
            if (last_opt) { stop = 1; break; }
            if (eps < target) { stop = 2; break; }
            eps = eps * theta;
            for (int j = lane; j < M; j += 32) B.owner[c0 + j] = -1;
            for (int i = lane; i < N; i += 32) { B.p2o[r0 + i] = -1; B.list[r0 + i] = (int)(r0 + i); }
            __syncwarp();
            nu = N;
            ++nred;
        }
    }
    if (last_opt < 0) last_opt = (nu == 0 && batch_ece(B, r0, N, lane, target)) ? 1 : 0;
    // ---- get_obj (:489-523): per-person chosen values (summed on the host in row order) and local column ids
    for (int i = 0; i < N; ++i) {
        const long long g = r0 + i;
        const int j = B.p2o[g];
        double csum = 0.0;
        if (j >= 0) {
            double vmax, choice;
            row_ece(B.cols, B.vals, B.price, __ldg(B.rowptr + g), __ldg(B.rowptr + g + 1), lane, j, vmax, choice, csum);
        }
        if (lane == 0) { B.chosen[g] = csum; B.p2o[g] = j >= 0 ? (int)(j - c0) : -1; }
    }
    if (lane == 0) {
        SslapbBatchMeta m;
        m.start_eps = start_eps; m.final_eps = eps; m.target_eps = target;
        m.eCE = last_opt; m.soln_found = (nu == 0 && last_opt) ? 1 : 0; m.stop_reason = stop;
        m.its = its; m.nreductions = nred; m.n_assigned = N - nu;
        B.meta[p] = m;
    }
}

// local (row, col) -> global block-diagonal ids; one CTA per problem
template <typename IT>
__global__ void __launch_bounds__(256) sslapb_batch_globalize_kernel(const IT *__restrict__ rows, const IT *__restrict__ cols,
                                                                     long long stride, const long long *__restrict__ nnzoff,
                                                                     const long long *__restrict__ rowoff,
                                                                     const long long *__restrict__ coloff, int P,
                                                                     int *__restrict__ rows_g, int *__restrict__ cols_g, int *bad)
{
    for (int p = blockIdx.x; p < P; p += gridDim.x) {
        const long long r0 = rowoff[p], c0 = coloff[p];
        const long long N = rowoff[p + 1] - r0, M = coloff[p + 1] - c0;
        for (long long k = nnzoff[p] + threadIdx.x; k < nnzoff[p + 1]; k += blockDim.x) {
            const long long r = (long long)rows[k * stride], c = (long long)cols[k * stride];
            if (r < 0 || r >= N || c < 0 || c >= M) { *bad = 1; rows_g[k] = (int)r0; cols_g[k] = (int)c0; }
            else { rows_g[k] = (int)(r0 + r); cols_g[k] = (int)(c0 + c); }
        }
    }
}

extern "C" cudaError_t sslapb_launch_batch_globalize(const void *rows, const void *cols, int idx_bytes, long long stride,
                                                     const long long *nnzoff, const long long *rowoff, const long long *coloff,
                                                     int P, int *rows_g, int *cols_g, int *bad, int sms, cudaStream_t stream)
{
    const int grid = P < sms * 8 ? P : sms * 8;
    if (idx_bytes == 4)
        sslapb_batch_globalize_kernel<int><<<grid, 256, 0, stream>>>((const int *)rows, (const int *)cols, stride, nnzoff, rowoff, coloff, P, rows_g, cols_g, bad);
    else
        sslapb_batch_globalize_kernel<long long><<<grid, 256, 0, stream>>>((const long long *)rows, (const long long *)cols, stride, nnzoff, rowoff, coloff, P, rows_g, cols_g, bad);
    return cudaGetLastError();
}

extern "C" cudaError_t sslapb_launch_auction_batch(const SslapbBatchParams *B, cudaStream_t stream)
{
    const int warps_per_cta = 4;
    const int grid = (B->P + warps_per_cta - 1) / warps_per_cta;
    sslapb_auction_batch_kernel<<<grid, 32 * warps_per_cta, 0, stream>>>(*B);
    return cudaGetLastError();
}

#endif  // !SSLAPB_LONG_ROWS
