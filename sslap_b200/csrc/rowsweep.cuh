// Row sweeps shared by the persistent auction kernel (auction.cu) and the streamed full-frontier sweep (sweep_tma.cu).
#pragma once
#include "auction.cuh"

// ----------------------------------------------------------------------------------------------------------------------
// Row sweep: top-2 of (a_ij - p_j) over one CSR row by a group of W lanes (bidding loop, auction_.pyx:346-358).
// Lane t of the group owns the 16-byte-aligned chunks t, t+W, ... of 4 consecutive entries (one int4 of columns, two
// double2 of values); entries of the chunk outside [start,end) belong to neighbouring rows and are masked.
// Returns (in every lane of the group) the object and the bid (a_ibest - w_i + eps, :360); jbest = -1 for an empty row.
// ----------------------------------------------------------------------------------------------------------------------
// Key of one candidate: order-preserving 64-bit image of v = a_ij - p_j (computed in float64 exactly as the reference
// does).  All comparisons of the sweep then run on the integer pipe (FP64 compares are ~5x slower on the critical path).
// -0.0 is folded into +0.0 first so that integer order reproduces the reference's floating-point ties; 0 = "no entry".
__device__ __forceinline__ unsigned long long sslapb_vkey(double val, double p, bool m)
{
    long long bits = __double_as_longlong(val - p);
    if ((bits << 1) == 0) bits = 0;
    const unsigned long long k = (unsigned long long)(bits ^ ((bits >> 63) | (long long)0x8000000000000000ull));
    return m ? k : 0ull;
}
__device__ __forceinline__ double sslapb_key2double(unsigned long long k)
{
    const long long b = (long long)k;
    return __longlong_as_double(b ^ (((~b) >> 63) | (long long)0x8000000000000000ull));
}

// What a bidder knows after its sweep: the object, the bid and (warp-list regimes only) the record of the object's
// current owner, fetched speculatively for every candidate so that no dependent load follows the reduction.
struct SslapbBid {
    int j;                    // object bid on (-1: empty row)
    double bid;
    int powner;               // current owner of j (-1 none)
    int pdeg;                 // its row
    long long pstart;
};

template <int W>
__device__ __forceinline__ unsigned sslapb_group_mask()
{
    return (W == 32) ? SSLAPB_FULL : (((1u << (W & 31)) - 1u) << ((threadIdx.x & 31) & ~(W - 1)));
}

// The sweep.  REC = false: prices gathered from price[] (grid regime, 8 B per candidate).  REC = true: prices and owner
// records gathered from the 32-byte object records (warp-list regimes).
template <int W, bool REC>
__device__ __forceinline__ SslapbBid row_bid_core(const int *__restrict__ cols, const double *__restrict__ vals,
                                                  const double *price, const SslapbObjRec *rec, long long start,
                                                  long long end, int t, double eps, double pmin = 0.0,
                                                  double thr = -__builtin_huge_val())
{
    // (pmin, thr): bound pruning as in row_bid_pruned — entries with a < thr are not gathered; the uniform test
    // fl(thr - pmin) < second-best proves them irrelevant, otherwise the row is swept again with thr = -inf.
    SslapbBid o;
#pragma unroll 1
    for (int pass = 0;; ++pass) {
    const int4 *c4 = reinterpret_cast<const int4 *>(cols);
    const double2 *v2 = reinterpret_cast<const double2 *>(vals);
    unsigned long long b = 0ull, s = 0ull;          // best / second-best key of this lane (0 = none)
    double bc = 0.0;                                // value a_ij of the best entry
    int bi = -1, bj = -1;                           // its index inside the row / its column
    int4 br = make_int4(0, 0, -1, 0);               // its object's record (REC)
    // Warp-uniform trip count (the widest group decides) keeps the loop convergent.
    const long long c0 = start >> 2, c1 = (end + 3) >> 2;
    int trips = (int)((c1 - c0 + (W - 1)) / W);
    trips = (W == 32) ? __shfl_sync(SSLAPB_FULL, trips, 0) : __reduce_max_sync(SSLAPB_FULL, trips);
#pragma unroll 1
    for (int it = 0; it < trips; ++it) {
        const long long ch = c0 + (long long)it * W + t;
        if (ch >= c1) continue;
        const int4 cj = REC ? __ldg(c4 + ch) : sslapb_ldg_stream_i4(c4 + ch);
        const double2 va = REC ? __ldg(v2 + 2 * ch) : sslapb_ldg_stream_d2(v2 + 2 * ch);
        const double2 vb = REC ? __ldg(v2 + 2 * ch + 1) : sslapb_ldg_stream_d2(v2 + 2 * ch + 1);
        const int lo = (int)(start - (ch << 2)), hi = (int)min(end - (ch << 2), 4ll);
        const bool m0 = (0 >= lo) & (0 < hi) & (va.x >= thr), m1 = (1 >= lo) & (1 < hi) & (va.y >= thr);
        const bool m2 = (2 >= lo) & (2 < hi) & (vb.x >= thr), m3 = (3 >= lo) & (3 < hi) & (vb.y >= thr);
        double p0, p1, p2, p3;
        int4 r0, r1, r2, r3;
        if (REC) {
            const int4 z = make_int4(0, 0, -1, 0);
            r0 = m0 ? *reinterpret_cast<const int4 *>(rec + cj.x) : z;
            r1 = m1 ? *reinterpret_cast<const int4 *>(rec + cj.y) : z;
            r2 = m2 ? *reinterpret_cast<const int4 *>(rec + cj.z) : z;
            r3 = m3 ? *reinterpret_cast<const int4 *>(rec + cj.w) : z;
            p0 = m0 ? rec[cj.x].price : 0.0;
            p1 = m1 ? rec[cj.y].price : 0.0;
            p2 = m2 ? rec[cj.z].price : 0.0;
            p3 = m3 ? rec[cj.w].price : 0.0;
        } else {
            p0 = m0 ? price[cj.x] : 0.0;
            p1 = m1 ? price[cj.y] : 0.0;
            p2 = m2 ? price[cj.z] : 0.0;
            p3 = m3 ? price[cj.w] : 0.0;
        }
        // top-2 of the four candidates by a two-level tournament (later entry wins equal keys: last maximal, :351)
        const unsigned long long k0 = sslapb_vkey(va.x, p0, m0), k1 = sslapb_vkey(va.y, p1, m1);
        const unsigned long long k2 = sslapb_vkey(vb.x, p2, m2), k3 = sslapb_vkey(vb.y, p3, m3);
        const bool w01 = k1 >= k0, w23 = k3 >= k2;
        const unsigned long long b01 = w01 ? k1 : k0, l01 = w01 ? k0 : k1;
        const unsigned long long b23 = w23 ? k3 : k2, l23 = w23 ? k2 : k3;
        const bool wf = b23 >= b01;
        const unsigned long long b4 = wf ? b23 : b01;
        const unsigned long long s4 = wf ? (b01 > l23 ? b01 : l23) : (b23 > l01 ? b23 : l01);
        const int w4 = wf ? (w23 ? 3 : 2) : (w01 ? 1 : 0);
        if (b4 >= b && b4 != 0ull) {                 // this chunk holds later entries: it wins equal keys
            s = b > s4 ? b : s4;
            b = b4;
            bi = (int)((ch << 2) - start) + w4;
            bc = (w4 & 2) ? ((w4 & 1) ? vb.y : vb.x) : ((w4 & 1) ? va.y : va.x);
            bj = (w4 & 2) ? ((w4 & 1) ? cj.w : cj.z) : ((w4 & 1) ? cj.y : cj.x);
            if (REC) br = (w4 & 2) ? ((w4 & 1) ? r3 : r2) : ((w4 & 1) ? r1 : r0);
        } else {
            s = b4 > s ? b4 : s;
        }
    }
    // cross-lane: lexicographic max of (key, row index) and the second-largest key, on the redux unit
    const unsigned gm = sslapb_group_mask<W>();
    const unsigned bh = (unsigned)(b >> 32), bl = (unsigned)b;
    const unsigned hi = __reduce_max_sync(gm, bh);
    const unsigned lo = __reduce_max_sync(gm, bh == hi ? bl : 0u);
    const bool top = (bh == hi) & (bl == lo);
    const int widx = __reduce_max_sync(gm, top ? bi : -1);
    const bool iswin = top & (bi == widx) & (bi >= 0);
    const unsigned long long cand = iswin ? s : b;
    const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
    const unsigned shi = __reduce_max_sync(gm, chh);
    const unsigned slo = __reduce_max_sync(gm, chh == shi ? chl : 0u);
    const unsigned long long skey = ((unsigned long long)shi << 32) | slo;
    const unsigned own = __ballot_sync(SSLAPB_FULL, iswin) & gm;
    const int src = own ? (__ffs(own) - 1) : (threadIdx.x & 31);
    bc = __shfl_sync(SSLAPB_FULL, bc, src);
    bj = __shfl_sync(SSLAPB_FULL, bj, src);
    o.j = own ? bj : -1;
    if (REC) {
        const int sx = __shfl_sync(SSLAPB_FULL, br.x, src), sy = __shfl_sync(SSLAPB_FULL, br.y, src);
        o.powner = __shfl_sync(SSLAPB_FULL, br.z, src);
        o.pdeg = __shfl_sync(SSLAPB_FULL, br.w, src);
        o.pstart = (long long)(((unsigned long long)(unsigned)sy << 32) | (unsigned)sx);
    } else {
        o.powner = -1; o.pdeg = 0; o.pstart = 0;
    }
    const double wi = skey ? sslapb_key2double(skey) : SSLAPB_NEG_INF;   // w_i = -inf for a single-entry row (:344)
    o.bid = (bc - wi) + eps;                                   // :360
    if (pass || !(thr > SSLAPB_NEG_INF) || ((thr - pmin) < wi)) break;
    thr = SSLAPB_NEG_INF;                                      // the skipped entries might matter: sweep again, gather all
    }
    return o;
}

struct SslapbStreamChunk { int4 cj; double2 va, vb; };
__device__ __forceinline__ SslapbStreamChunk sslapb_stream_chunk(const int *__restrict__ cols,
                                                                 const double *__restrict__ vals, long long start,
                                                                 long long end, int lane)
{
    SslapbStreamChunk c;
    c.cj = make_int4(0, 0, 0, 0); c.va = make_double2(0.0, 0.0); c.vb = c.va;
    const long long ch = (start >> 2) + lane;
    if (ch < ((end + 3) >> 2)) {
        c.cj = sslapb_ldg_stream_i4(reinterpret_cast<const int4 *>(cols) + ch);
        c.va = sslapb_ldg_stream_d2(reinterpret_cast<const double2 *>(vals) + 2 * ch);
        c.vb = sslapb_ldg_stream_d2(reinterpret_cast<const double2 *>(vals) + 2 * ch + 1);
    }
    return c;
}

// Top-2 of four values by a two-level tournament in float64 (the later slot wins equal values: "last maximal entry",
// auction_.pyx:351; -0.0 == +0.0 exactly as in the reference).  Absent slots carry -inf.
struct SslapbLaneTop { double b, s; int w; };
__device__ __forceinline__ SslapbLaneTop sslapb_lane_top2(double v0, double v1, double v2, double v3)
{
    const bool t01 = v1 >= v0, t23 = v3 >= v2;
    const double b01 = t01 ? v1 : v0, l01 = t01 ? v0 : v1;
    const double b23 = t23 ? v3 : v2, l23 = t23 ? v2 : v3;
    const bool tf = b23 >= b01;
    SslapbLaneTop r;
    r.b = tf ? b23 : b01;
    const double x = tf ? b01 : b23, y = tf ? l23 : l01;
    r.s = x > y ? x : y;
    r.w = tf ? (t23 ? 3 : 2) : (t01 ? 1 : 0);
    return r;
}

// order-preserving key with -0.0 folded into +0.0 (cross-lane ties must behave like the float compare)
__device__ __forceinline__ unsigned long long sslapb_key_of(double v)
{
    long long bits = __double_as_longlong(v);
    if ((bits << 1) == 0) bits = 0;
    return (unsigned long long)(bits ^ ((bits >> 63) | (long long)0x8000000000000000ull));
}

// Cross-lane part shared by the single-pass sweeps: lexicographic maximum of (value, row index) and the second largest
// value of the row, from each lane's (best, second, index) — five REDUX on the integer images of the values.
struct SslapbRowTop { bool iswin; unsigned own; unsigned long long skey; };
__device__ __forceinline__ SslapbRowTop sslapb_row_top2(double b, double s, int bi)
{
    const unsigned long long bk = bi >= 0 ? sslapb_key_of(b) : 0ull;
    const unsigned long long sk = bi >= 0 ? sslapb_key_of(s) : 0ull;
    const unsigned bh = (unsigned)(bk >> 32), bl = (unsigned)bk;
    const unsigned khi = __reduce_max_sync(SSLAPB_FULL, bh);
    const unsigned klo = __reduce_max_sync(SSLAPB_FULL, bh == khi ? bl : 0u);
    const bool top = (bh == khi) & (bl == klo);
    const int widx = __reduce_max_sync(SSLAPB_FULL, top ? bi : -1);
    SslapbRowTop r;
    r.iswin = top & (bi == widx) & (bi >= 0);
    const unsigned long long cand = r.iswin ? sk : bk;
    const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
    const unsigned shi = __reduce_max_sync(SSLAPB_FULL, chh);
    const unsigned slo = __reduce_max_sync(SSLAPB_FULL, chh == shi ? chl : 0u);
    r.skey = ((unsigned long long)shi << 32) | slo;
    r.own = __ballot_sync(SSLAPB_FULL, r.iswin);
    return r;
}
// key of "-inf": a second-best key at or below it means the row has no second entry (w_i = -inf, auction_.pyx:344)
#define SSLAPB_KEY_NEG_INF 0x000fffffffffffffull

// Bound-pruned sweep of one row that fits a single warp pass (grid regime).  Exact, but most price gathers are skipped:
//   v_k = a_k - p_k <= a_k - L for any lower bound L of the prices.  Only the entries with a_k >= thr are gathered, where
//   thr = (static row maximum of a) - (price spread at the start of the phase) is a heuristic.  Every other entry has
//   a_k < thr, hence fl(a_k - p_k) <= fl(a_k - L) <= fl(thr - L) (rounding is monotone), so ONE uniform comparison
//   fl(thr - L) < (second-best value found) proves that none of them can be the best or the second best.  If it fails
//   (rare: the spread is stale because prices moved inside the phase) the row is redone with every entry gathered.
// The random price gathers, not the 12 B/entry stream, are what loads L1TEX/L2 in this kernel (one wavefront and one
// 32-byte sector per 8-byte price).  Entries of the chunk that belong to neighbouring rows are replaced by a = -inf.
__device__ __forceinline__ SslapbBid row_bid_pruned(const SslapbStreamChunk &C, const double *price, long long start,
                                                    long long end, int lane, double eps, double pmin, double thr,
                                                    int &second_pass)
{
    const long long ch = (start >> 2) + lane;
    const int off = (int)((ch << 2) - start);                 // row index of slot 0 (may be negative)
    const int deg = (int)(end - start);
    const int4 cj = C.cj;
    const double a0 = (unsigned)off < (unsigned)deg ? C.va.x : SSLAPB_NEG_INF;
    const double a1 = (unsigned)(off + 1) < (unsigned)deg ? C.va.y : SSLAPB_NEG_INF;
    const double a2 = (unsigned)(off + 2) < (unsigned)deg ? C.vb.x : SSLAPB_NEG_INF;
    const double a3 = (unsigned)(off + 3) < (unsigned)deg ? C.vb.y : SSLAPB_NEG_INF;
    double v0 = SSLAPB_NEG_INF, v1 = SSLAPB_NEG_INF, v2 = SSLAPB_NEG_INF, v3 = SSLAPB_NEG_INF;
    if (a0 >= thr) v0 = a0 - price[cj.x];                      // thr > -inf: absent slots are never gathered;
    if (a1 >= thr) v1 = a1 - price[cj.y];                      // thr = -inf (no pruning): they give -inf - p = -inf
    if (a2 >= thr) v2 = a2 - price[cj.z];
    if (a3 >= thr) v3 = a3 - price[cj.w];
    const SslapbLaneTop lt = sslapb_lane_top2(v0, v1, v2, v3);
    const int bi = ((unsigned)(off + lt.w) < (unsigned)deg && lt.b > SSLAPB_NEG_INF) ? off + lt.w : -1;
    const SslapbRowTop rt = sslapb_row_top2(lt.b, lt.s, bi);
    const int src = rt.own ? (__ffs(rt.own) - 1) : lane;
    const double myc = (lt.w & 2) ? ((lt.w & 1) ? a3 : a2) : ((lt.w & 1) ? a1 : a0);
    const int myj = (lt.w & 2) ? ((lt.w & 1) ? cj.w : cj.z) : ((lt.w & 1) ? cj.y : cj.x);
    const double bc = __shfl_sync(SSLAPB_FULL, myc, src);
    const int bj = __shfl_sync(SSLAPB_FULL, myj, src);
    const double wi = rt.skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(rt.skey) : SSLAPB_NEG_INF;   // :344
    SslapbBid o;
    o.powner = -1; o.pdeg = 0; o.pstart = 0;
    o.bid = (bc - wi) + eps;                                   // :360
    // j = -1 sends the caller to the exact generic sweep: every candidate at -inf, or (uniform test) a skipped entry
    // might matter — rare: the spread is stale because prices moved inside the phase
    const bool proven = !(thr > SSLAPB_NEG_INF) || ((thr - pmin) < wi);
    if (!proven) ++second_pass;
    o.j = (rt.own && proven) ? bj : -1;
    return o;
}

// Same function as row_bid_pruned with fewer instructions on the per-lane path (used by the pipelined sweep, sweep2.cu):
//   * the in-row test is folded into the gather predicate (DSETP.AND) instead of substituting -inf for absent slots first;
//   * -0.0 is folded into +0.0 by one addition of +0.0 (exact, IEEE: -0 + +0 = +0) instead of compare + selects;
//   * the winning slot is always a gathered, hence in-row, slot: no second range test.
__device__ __forceinline__ SslapbBid row_bid_pruned_lean(const SslapbStreamChunk &C, const double *price, long long start,
                                                         int deg, int lane, double eps, double pmin, double thr,
                                                         int &second_pass)
{
    const int off = 4 * lane - (int)(start & 3);              // row index of slot 0 (may be negative)
    const int4 cj = C.cj;
    const bool g0 = ((unsigned)off < (unsigned)deg) && (C.va.x >= thr);
    const bool g1 = ((unsigned)(off + 1) < (unsigned)deg) && (C.va.y >= thr);
    const bool g2 = ((unsigned)(off + 2) < (unsigned)deg) && (C.vb.x >= thr);
    const bool g3 = ((unsigned)(off + 3) < (unsigned)deg) && (C.vb.y >= thr);
    double v0 = SSLAPB_NEG_INF, v1 = SSLAPB_NEG_INF, v2 = SSLAPB_NEG_INF, v3 = SSLAPB_NEG_INF;
    if (g0) v0 = C.va.x - price[cj.x];
    if (g1) v1 = C.va.y - price[cj.y];
    if (g2) v2 = C.vb.x - price[cj.z];
    if (g3) v3 = C.vb.y - price[cj.w];
    const SslapbLaneTop lt = sslapb_lane_top2(v0, v1, v2, v3);
    const bool has = lt.b > SSLAPB_NEG_INF;
    const int bi = has ? off + lt.w : -1;
    const unsigned long long bk = has ? sslapb_ord64(lt.b + 0.0) : 0ull;
    const unsigned long long sk = has ? sslapb_ord64(lt.s + 0.0) : 0ull;
    const unsigned bh = (unsigned)(bk >> 32), bl = (unsigned)bk;
    const unsigned khi = __reduce_max_sync(SSLAPB_FULL, bh);
    const unsigned klo = __reduce_max_sync(SSLAPB_FULL, bh == khi ? bl : 0u);
    const bool top = (bh == khi) & (bl == klo);
    const int widx = __reduce_max_sync(SSLAPB_FULL, top ? bi : -1);
    const bool iswin = top & (bi == widx) & has;
    const unsigned long long cand = iswin ? sk : bk;
    const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
    const unsigned shi = __reduce_max_sync(SSLAPB_FULL, chh);
    const unsigned slo = __reduce_max_sync(SSLAPB_FULL, chh == shi ? chl : 0u);
    const unsigned long long skey = ((unsigned long long)shi << 32) | slo;
    const unsigned own = __ballot_sync(SSLAPB_FULL, iswin);
    const int src = own ? (__ffs(own) - 1) : lane;
    const double myc = (lt.w & 2) ? ((lt.w & 1) ? C.vb.y : C.vb.x) : ((lt.w & 1) ? C.va.y : C.va.x);
    const int myj = (lt.w & 2) ? ((lt.w & 1) ? cj.w : cj.z) : ((lt.w & 1) ? cj.y : cj.x);
    const double bc = __shfl_sync(SSLAPB_FULL, myc, src);
    const int bj = __shfl_sync(SSLAPB_FULL, myj, src);
    const double wi = skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(skey) : SSLAPB_NEG_INF;   // :344
    SslapbBid o;
    o.powner = -1; o.pdeg = 0; o.pstart = 0;
    o.bid = (bc - wi) + eps;                                   // :360
    const bool proven = !(thr > SSLAPB_NEG_INF) || ((thr - pmin) < wi);
    if (!proven) ++second_pass;
    o.j = (own && proven) ? bj : -1;
    return o;
}

// Bid of one person from its hot list (hot.cu): lane t holds hot entry t; one price gather per lane, no per-lane
// tournament.  Exact when the second-best value found is strictly above `rest` (the bound of everything outside the
// list), or when nothing is outside (rest = -inf); otherwise j = -1 sends the caller to the full-row sweep.
struct SslapbHotIn { int4 q; double rest; };                  // lane t: hot entry t of the person + the row's bound
__device__ __forceinline__ SslapbHotIn sslapb_hot_in(const SslapbHotEnt *__restrict__ hot, const double *rest_arr, int person, int lane)
{
    SslapbHotIn in;
    in.q = __ldg(reinterpret_cast<const int4 *>(hot) + (long long)person * 32 + lane);
    in.rest = rest_arr[person];
    return in;
}
__device__ __forceinline__ SslapbBid row_bid_hot_from(const SslapbHotIn &in, const double *price, int lane, double eps);
__device__ __forceinline__ SslapbBid row_bid_hot(const SslapbHotEnt *__restrict__ hot, const double *rest_arr, const double *price,
                                                 int person, int lane, double eps)
{
    return row_bid_hot_from(sslapb_hot_in(hot, rest_arr, person, lane), price, lane, eps);
}
__device__ __forceinline__ SslapbBid row_bid_hot_from(const SslapbHotIn &in, const double *price, int lane, double eps)
{
    const int4 q = in.q;
    const double rest = in.rest;
    const double a = __hiloint2double(q.w, q.z);
    const double v = a - price[q.x];                           // padding: column 0, a = -inf -> v = -inf
    const bool has = v > SSLAPB_NEG_INF;                       // real -inf candidates: left to the generic sweep
    const unsigned long long bk = has ? sslapb_key_of(v) : 0ull;
    const unsigned bh = (unsigned)(bk >> 32), bl = (unsigned)bk;
    const unsigned khi = __reduce_max_sync(SSLAPB_FULL, bh);
    const unsigned hm = __ballot_sync(SSLAPB_FULL, (bh == khi) & has);
    bool iswin;
    unsigned own;
    if (__popc(hm) <= 1) {                                     // one lane holds the maximal high word: it is the winner
        own = hm;
        iswin = (hm >> lane) & 1u;
    } else {
        const unsigned klo = __reduce_max_sync(SSLAPB_FULL, bh == khi ? bl : 0u);
        const bool top = (bh == khi) & (bl == klo) & has;
        const int widx = __reduce_max_sync(SSLAPB_FULL, top ? q.y : -1);   // equal values: the later row entry wins (:351)
        iswin = top & (q.y == widx);
        own = __ballot_sync(SSLAPB_FULL, iswin);
    }
    const unsigned long long cand = iswin ? 0ull : bk;         // one candidate per lane: the second best is the best of the others
    const unsigned chh = (unsigned)(cand >> 32), chl = (unsigned)cand;
    const unsigned shi = __reduce_max_sync(SSLAPB_FULL, chh);
    const unsigned slo = __reduce_max_sync(SSLAPB_FULL, chh == shi ? chl : 0u);
    const unsigned long long skey = ((unsigned long long)shi << 32) | slo;
    const int src = own ? (__ffs(own) - 1) : lane;
    const double bc = __shfl_sync(SSLAPB_FULL, a, src);
    const int bj = __shfl_sync(SSLAPB_FULL, q.x, src);
    const double wi = skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(skey) : SSLAPB_NEG_INF;   // :344
    SslapbBid o;
    o.powner = -1; o.pdeg = 0; o.pstart = 0;
    o.bid = (bc - wi) + eps;                                   // :360
    const bool proven = (wi > rest) || (rest == SSLAPB_NEG_INF);
    o.j = (own && proven) ? bj : -1;
    return o;
}

template <int W>
__device__ __forceinline__ void row_bid(const int *__restrict__ cols, const double *__restrict__ vals,
                                        const double *price, long long start, long long end, int t, double eps,
                                        int &jbest, double &bid, double pmin = 0.0, double thr = -__builtin_huge_val())
{
    const SslapbBid o = row_bid_core<W, false>(cols, vals, price, nullptr, start, end, t, eps, pmin, thr);
    jbest = o.j;
    bid = o.bid;
}

// Out of line on purpose: the warp-list loops are executed by one to a few warps, so their speed is set by instruction
// fetch as much as by memory; keeping the rare long-row sweep out of the loop bodies keeps those bodies inside the
// instruction cache's first level.
template <int W>
__device__ __noinline__ SslapbBid row_bid_rec(const int *__restrict__ cols, const double *__restrict__ vals,
                                              const SslapbObjRec *rec, long long start, long long end, int t,
                                              double eps, double pmin = 0.0, double thr = -__builtin_huge_val())
{
    return row_bid_core<W, true>(cols, vals, nullptr, rec, start, end, t, eps, pmin, thr);
}

// eCE / objective sweep of one row by a full warp (auction_.pyx:460-483 and :504-521).
//   vmax   = max_k (a_ik - p_k)
//   choice = value of the LAST entry whose column is jsel (:467-471)
//   csum   = sum of the values of ALL entries whose column is jsel (get_obj adds every match, :514-521)
//   REST: additionally rest = max_k (a_ik - p_k) over the entries with a_ik <= hthr — the entries outside the person's
//   hot list (hot.cu); -inf when there is none
template <bool REST = false>
__device__ __forceinline__ void row_ece(const int *__restrict__ cols, const double *__restrict__ vals,
                                        const double *price, long long start, long long end, int lane, int jsel,
                                        double &vmax, double &choice, double &csum, double hthr = 0.0, double *rest = nullptr)
{
    const int4 *c4 = reinterpret_cast<const int4 *>(cols);
    const double2 *v2 = reinterpret_cast<const double2 *>(vals);
    double vm = SSLAPB_NEG_INF, ch_v = 0.0, cs = 0.0, rm = SSLAPB_NEG_INF;
    int ch_i = -1;
    const long long c0 = start >> 2, c1 = (end + 3) >> 2;
    const int trips = __shfl_sync(SSLAPB_FULL, (int)((c1 - c0 + 31) / 32), 0);
    for (int it = 0; it < trips; ++it) {
        const long long ch = c0 + (long long)it * 32 + lane;
        if (ch >= c1) continue;
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
                if (REST && vv[k] <= hthr) rm = fmax(rm, v);
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
        if (REST) rm = fmax(rm, __shfl_xor_sync(SSLAPB_FULL, rm, off));
    }
    if (REST) *rest = rm;
    vmax = vm; choice = ch_v; csum = cs;
}

