// Small problems in ONE launch (sm_100a): CSR build, feasibility check and the whole eps-scaling auction by a single CTA
// with every array in shared memory.
//
// Why.  The reference's own benchmark (benchmarking.py:84-147) starts at 10 x 10; it solves that in 20-100 us on one
// core.  The general path costs ~0.26 ms there whatever the kernels do: eight launches, a cooperative launch among them,
// three or four stream synchronisations.  This path is one H2D copy, ONE launch, one D2H copy — and a round costs
// shared-memory latencies instead of L2 round trips.
//
// What it computes is the reference's trajectory, bit for bit (sol, its, nreductions, prices, meta), exactly as the
// batch kernel does with one warp per problem (auction.cu), here with 16 warps on one problem:
//   bidding    warp per unassigned person, top-2 of a_ij - p_j over its row, LAST maximal entry wins (:346-358)
//   merge      per-object maximum of the order-preserving bid (shared-memory atomicMax), earliest list position on equal
//              bids (:375-385)
//   assignment winner-driven, evicted owner takes the winner's slot (:394-427)
//   compaction push_all_left in list order (:137-162) by one block scan
//   eps-CS     eCE_satisfied / terminate / eps-scaling in float32 (:268-309, :443-485), get_obj (:489-523)
// Feasibility (cardinality_check): greedy + Kuhn's augmenting paths by one warp (the cardinality of a maximum matching
// does not depend on the algorithm; feasibility_.pyx:95-211 is Hopcroft-Karp).
// Limits: N, M <= 256 and at most SSLAPB_SMALL_CAP entries; anything else — and unsorted or out-of-range COO input — is
// reported back (status 3) and takes the general path.
#include "auction.cuh"
#include "rowsweep.cuh"
#include "small.cuh"

#define SM_THREADS 512
#define SM_WARPS (SM_THREADS / 32)

struct SmallShared {
    double vals[SSLAPB_SMALL_CAP];
    int cols[SSLAPB_SMALL_CAP];
    double price[SSLAPB_SMALL_MAXN], bidv[SSLAPB_SMALL_MAXN], chosen[SSLAPB_SMALL_MAXN];
    unsigned long long bestkey[SSLAPB_SMALL_MAXN];
    int rowptr[SSLAPB_SMALL_MAXN + 1];
    int owner[SSLAPB_SMALL_MAXN], p2o[SSLAPB_SMALL_MAXN], list[SSLAPB_SMALL_MAXN], mover[SSLAPB_SMALL_MAXN];
    int bidj[SSLAPB_SMALL_MAXN], winpos[SSLAPB_SMALL_MAXN];
    int wtot[32];
    int tie, misc, hsplit;
    long long chain_rounds;
    unsigned long long amax_bits;
};

// block-wide exclusive prefix of a flag over the threads (thread order) + total; three barriers
__device__ __forceinline__ int small_scan_flag(SmallShared &S, bool flag, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(SSLAPB_FULL, flag);
    const int inwarp = __popc(bal & ((1u << lane) - 1u));
    __syncthreads();
    if (lane == 0) S.wtot[warp] = __popc(bal);
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < SM_WARPS; ++w) { const int c = S.wtot[w]; if (w < warp) pre += c; tot += c; }
    __syncthreads();
    total = tot;
    return pre + inwarp;
}

// eps-CS test of one person by one warp (eCE_satisfied, :460-483) + the chosen value of get_obj (:504-521)
__device__ __forceinline__ bool small_ece_row(const SmallShared &S, int i, int lane, double eps_t, double tol, double &csum_out)
{
    const int j = S.p2o[i];
    const int st = S.rowptr[i], en = S.rowptr[i + 1];
    double vm = SSLAPB_NEG_INF, ch_v = 0.0, cs = 0.0;
    int ch_i = -1;
    for (int e = st + lane; e < en; e += 32) {
        const int c = S.cols[e];
        const double a = S.vals[e];
        vm = fmax(vm, a - S.price[c]);
        if (c == j) { ch_v = a; ch_i = e - st; cs += a; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        vm = fmax(vm, __shfl_xor_sync(SSLAPB_FULL, vm, off));
        const int oi = __shfl_xor_sync(SSLAPB_FULL, ch_i, off);
        const double ov = __shfl_xor_sync(SSLAPB_FULL, ch_v, off);
        if (oi > ch_i) { ch_i = oi; ch_v = ov; }
        cs += __shfl_xor_sync(SSLAPB_FULL, cs, off);
    }
    csum_out = cs;
    if (j < 0) return false;
    return ((ch_v - S.price[j]) + tol) < vm - eps_t;            // :475, :482 — true = violated
}

__global__ void __launch_bounds__(SM_THREADS, 1) sslapb_small_kernel(SslapbSmallArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallShared &S = *reinterpret_cast<SmallShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = A.N, M = A.M;
    SslapbSmallResult R;
    R.status = 0; R.cardinality = -1; R.start_eps = R.final_eps = R.target_eps = 0.f; R.eCE = 0; R.soln_found = 0;
    R.its = 0; R.nreductions = 0; R.n_assigned = 0; R.obj64 = 0.0; R.nnz = 0;

    // ------------------------------------------------------------------ CSR build in shared memory
    if (tid == 0) { S.tie = 0; S.misc = 0; S.amax_bits = 0ull; }
    for (int i = tid; i <= N; i += SM_THREADS) S.rowptr[i] = 0;
    __syncthreads();
    double amax = 0.0;
    if (A.dense) {
        // _from_matrix (auction_.pyx:546-553): entries >= 0 are valid, row-major order kept
        for (int r = warp; r < N; r += SM_WARPS) {
            int cnt = 0;
            for (int c = lane; c < M; c += 32) cnt += (A.mat[(long long)r * M + c] >= 0.0);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(SSLAPB_FULL, cnt, off);
            if (lane == 0) S.rowptr[r + 1] = cnt;
        }
        __syncthreads();
        if (tid == 0) { int acc = 0; for (int r = 0; r < N; ++r) { acc += S.rowptr[r + 1]; S.rowptr[r + 1] = acc; } S.misc = acc; }
        __syncthreads();
        const int nnz = S.misc;
        if (nnz > SSLAPB_SMALL_CAP) { if (tid == 0) { R.status = 3; *A.res = R; } return; }
        for (int r = warp; r < N; r += SM_WARPS) {
            int pos = S.rowptr[r];
            for (int c0 = 0; c0 < M; c0 += 32) {
                const int c = c0 + lane;
                const double v = c < M ? A.mat[(long long)r * M + c] : -1.0;
                const bool ok = v >= 0.0;
                const unsigned bal = __ballot_sync(SSLAPB_FULL, ok);
                if (ok) {
                    const int o = pos + __popc(bal & ((1u << lane) - 1u));
                    const double fv = A.negate ? v * -1.0 : v;
                    S.cols[o] = c; S.vals[o] = fv;
                    amax = fmax(amax, fabs(fv));
                }
                pos += __popc(bal);
            }
        }
    } else {
        // _from_sparse (auction_.pyx:575-617): row-sorted COO; entry k keeps position k
        const int nnz = A.nnz;
        int bad = 0;
        for (int k = tid; k < nnz; k += SM_THREADS) {
            const int r = A.rows[(long long)k * A.stride], c = A.cols_in[(long long)k * A.stride];
            if (r < 0 || r >= N || c < 0 || c >= M) { bad = 1; continue; }
            if (k > 0 && A.rows[(long long)(k - 1) * A.stride] > r) bad = 1;            // not row-sorted
            const double fv = A.negate ? A.val[k] * -1.0 : A.val[k];
            S.cols[k] = c; S.vals[k] = fv;
            amax = fmax(amax, fabs(fv));
            atomicAdd(&S.rowptr[r + 1], 1);
        }
        if (__syncthreads_or(bad)) { if (tid == 0) { R.status = 3; *A.res = R; } return; }
        if (tid == 0) { int acc = 0; for (int r = 0; r < N; ++r) { acc += S.rowptr[r + 1]; S.rowptr[r + 1] = acc; } S.misc = acc; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) amax = fmax(amax, __shfl_xor_sync(SSLAPB_FULL, amax, off));
    if (lane == 0 && amax > 0.0) atomicMax(&S.amax_bits, (unsigned long long)__double_as_longlong(amax));
    __syncthreads();
    const int nnz = S.misc;
    R.nnz = nnz;
    if (nnz < N) { if (tid == 0) { R.status = 1; *A.res = R; } return; }        // auction_.pyx:559-560 / :604-605
    int empty = 0;
    for (int r = tid; r < N; r += SM_THREADS) empty |= (S.rowptr[r + 1] == S.rowptr[r]);
    empty = __syncthreads_or(empty);

    // ------------------------------------------------------------------ feasibility (:562-566 / :608-612)
    if (A.hk) {
        int *pair_u = S.p2o, *pair_v = S.owner, *visited = S.winpos, *stk_u = S.list, *stk_v = S.mover, *ptr = S.bidj;
        for (int i = tid; i < N; i += SM_THREADS) pair_u[i] = -1;
        for (int j = tid; j < M; j += SM_THREADS) pair_v[j] = -1;
        __syncthreads();
        // greedy maximal matching by all warps (every left vertex grabs its first free neighbour, shared-memory CAS) ...
        for (int u = warp; u < N; u += SM_WARPS) {
            bool done = false;
            for (int base = S.rowptr[u]; base < S.rowptr[u + 1] && !done; base += 32) {
                const int e = base + lane;
                int v = -1;
                bool fr = false;
                if (e < S.rowptr[u + 1]) { v = S.cols[e]; fr = *(volatile int *)(pair_v + v) == -1; }
                unsigned cand = __ballot_sync(SSLAPB_FULL, fr);
                while (cand && !done) {
                    const int l = __ffs(cand) - 1;
                    cand &= cand - 1;
                    int ok = 0;
                    if (lane == l) ok = atomicCAS(pair_v + v, -1, u) == -1;
                    ok = __shfl_sync(SSLAPB_FULL, ok, l);
                    if (ok) { if (lane == l) pair_u[u] = v; done = true; }
                }
            }
        }
        __syncthreads();
        // ... then Kuhn's augmenting searches for the vertices it left free, by one warp; a vertex that has a free
        // neighbour ends the search at once (look-ahead over its whole adjacency before descending)
        if (warp == 0) {
            int card = 0;
            for (int u = lane; u < N; u += 32) card += pair_u[u] >= 0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) card += __shfl_xor_sync(SSLAPB_FULL, card, off);
            for (int u0 = 0; u0 < N && card < N; ++u0) {
                if (pair_u[u0] >= 0) continue;
                for (int j = lane; j < M; j += 32) visited[j] = 0;
                __syncwarp();
                int d = 0;
                if (lane == 0) { stk_u[0] = u0; ptr[0] = S.rowptr[u0]; }
                __syncwarp();
                bool ok = false;
                while (d >= 0) {
                    const int u = stk_u[d];
                    const int p0 = ptr[d], st = S.rowptr[u], en = S.rowptr[u + 1];
                    if (p0 == st) {                            // first visit: is any neighbour free?
                        int vfree = -1;
                        for (int base = st; base < en && vfree < 0; base += 32) {
                            const int e = base + lane;
                            int v = -1;
                            bool fr = false;
                            if (e < en) { v = S.cols[e]; fr = pair_v[v] < 0; }
                            const unsigned bal = __ballot_sync(SSLAPB_FULL, fr);
                            if (bal) vfree = __shfl_sync(SSLAPB_FULL, v, __ffs(bal) - 1);
                        }
                        if (vfree >= 0) {
                            if (lane == 0) {
                                stk_v[d] = vfree;
                                for (int k = d; k >= 0; --k) { pair_v[stk_v[k]] = stk_u[k]; pair_u[stk_u[k]] = stk_v[k]; }
                            }
                            __syncwarp();
                            ok = true;
                            break;
                        }
                    }
                    if (p0 >= en) { --d; continue; }
                    const int e = p0 + lane;
                    int v = -1;
                    bool cand = false;
                    if (e < en) { v = S.cols[e]; cand = visited[v] == 0; }
                    const unsigned bal = __ballot_sync(SSLAPB_FULL, cand);
                    if (bal == 0u) { if (lane == 0) ptr[d] = p0 + 32; __syncwarp(); continue; }
                    const int l = __ffs(bal) - 1;
                    const int vs = __shfl_sync(SSLAPB_FULL, v, l);
                    const int w = pair_v[vs];                  // matched (a free one would have ended the search above)
                    if (lane == 0) { visited[vs] = 1; ptr[d] = p0 + l + 1; stk_v[d] = vs; }
                    __syncwarp();
                    ++d;
                    if (lane == 0) { stk_u[d] = w; ptr[d] = S.rowptr[w]; }
                    __syncwarp();
                }
                card += ok ? 1 : 0;
            }
            if (lane == 0) S.misc = card;
        }
        __syncthreads();
        R.cardinality = S.misc;
        if (R.cardinality < N) { if (tid == 0) { R.status = 2; *A.res = R; } return; }
    } else if (empty) {
        if (tid == 0) { R.status = 6; *A.res = R; }
        return;
    }
    __syncthreads();

    // ------------------------------------------------------------------ AuctionSolver.__init__ (:220-261)
    for (int j = tid; j < M; j += SM_THREADS) { S.price[j] = 0.0; S.owner[j] = -1; S.bestkey[j] = 0ull; S.winpos[j] = 0x7fffffff; }
    for (int i = tid; i < N; i += SM_THREADS) { S.p2o[i] = -1; S.list[i] = i; }
    const float Cf = (float)__longlong_as_double((long long)S.amax_bits);       // max_val (:123-134), float32 (:242-243)
    float eps = (float)((double)Cf / 2.0);                                       // :246
    const float target = (float)(1.0 / (double)N), theta = 0.15f;               // :247-248
    if (A.eps_start > 0.f) eps = A.eps_start;                                    // :251-252
    R.start_eps = eps; R.target_eps = target;
    const double eps_t = (double)target, tol = 1e-7;                             // :16
    __syncthreads();

    int nu = N, nred = 0, last_opt = -1;
    long long its = 0;
    for (;;) {
        const double epsd = (double)eps;
        if (nu == 1) {
            // ---- single-bidder chain (about half of all rounds): the bidder wins (nobody to merge with), evicts the owner,
            // who bids next ... — warp 0 alone, no block barrier, until the frontier is empty or max_iter is reached
            if (warp == 0) {
                int i = S.list[0];
                long long done_rounds = 0;
                int left = 1;
                for (;;) {
                    const int st = S.rowptr[i], en = S.rowptr[i + 1];
                    double b = SSLAPB_NEG_INF, s = SSLAPB_NEG_INF, bc = 0.0;
                    int bi = -1, bj = -1;
                    for (int e = st + lane; e < en; e += 32) {
                        const int c = S.cols[e];
                        const double a = S.vals[e];
                        const double v = a - S.price[c];
                        if (v >= b || bi < 0) { s = b; b = v; bi = e - st; bc = a; bj = c; }
                        else if (v > s) s = v;
                    }
                    const SslapbRowTop rt = sslapb_row_top2(b, s, bi);
                    const int src = rt.own ? (__ffs(rt.own) - 1) : 0;
                    const double wbc = __shfl_sync(SSLAPB_FULL, bc, src);
                    const int j = __shfl_sync(SSLAPB_FULL, bj, src);
                    const double wi = rt.skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(rt.skey) : SSLAPB_NEG_INF;
                    const double bid = (wbc - wi) + epsd;
                    const int prev = S.owner[j];
                    __syncwarp();
                    if (lane == 0) {
                        S.price[j] = bid; S.owner[j] = i; S.p2o[i] = j;
                        if (prev >= 0) S.p2o[prev] = -1;
                    }
                    __syncwarp();
                    ++done_rounds;
                    if (prev < 0) { left = 0; break; }         // nobody evicted: the frontier is empty
                    i = prev;
                    if (its + done_rounds >= A.max_iter) break;
                }
                if (lane == 0) { S.list[0] = left ? i : -1; S.misc = left; S.chain_rounds = done_rounds; }
            }
            __syncthreads();
            nu = S.misc;
            its += S.chain_rounds;
            __syncthreads();
        } else {
        // ---- bidding (:339-365): warp per list position
        for (int n = warp; n < nu; n += SM_WARPS) {
            const int i = S.list[n];
            const int st = S.rowptr[i], en = S.rowptr[i + 1];
            double b = SSLAPB_NEG_INF, s = SSLAPB_NEG_INF, bc = 0.0;
            int bi = -1, bj = -1;
            for (int e = st + lane; e < en; e += 32) {
                const int c = S.cols[e];
                const double a = S.vals[e];
                const double v = a - S.price[c];
                if (v >= b || bi < 0) { s = b; b = v; bi = e - st; bc = a; bj = c; }        // last maximal entry wins (:351)
                else if (v > s) s = v;
            }
            const SslapbRowTop rt = sslapb_row_top2(b, s, bi);
            const int src = rt.own ? (__ffs(rt.own) - 1) : 0;
            const double wbc = __shfl_sync(SSLAPB_FULL, bc, src);
            const int wbj = __shfl_sync(SSLAPB_FULL, bj, src);
            const double wi = rt.skey > SSLAPB_KEY_NEG_INF ? sslapb_key2double(rt.skey) : SSLAPB_NEG_INF;   // :344
            const double bid = (wbc - wi) + epsd;                                                           // :360
            if (lane == 0) {
                S.bidj[n] = wbj; S.bidv[n] = bid;
                const unsigned long long key = sslapb_ord64(bid);
                if (atomicMax(&S.bestkey[wbj], key) == key) S.tie = 1;
            }
        }
        __syncthreads();
        // ---- merge (:375-385): equal best bids -> the earliest list position
        const int tie = S.tie;
        int myj = -1;
        double mybid = 0.0;
        if (tid < nu) { myj = S.bidj[tid]; mybid = S.bidv[tid]; }
        if (tie) {
            if (tid < nu && S.bestkey[myj] == sslapb_ord64(mybid)) atomicMin(&S.winpos[myj], tid);
            __syncthreads();
        }
        // ---- assignment (:394-427), thread n = list position n
        bool hole = false;
        if (tid < nu) {
            const bool win = (S.bestkey[myj] == sslapb_ord64(mybid)) && (!tie || S.winpos[myj] == tid);
            if (win) {
                const int i = S.list[tid];
                const int prev = S.owner[myj];
                S.price[myj] = mybid; S.owner[myj] = i; S.p2o[i] = myj;
                if (prev >= 0) S.p2o[prev] = -1;
                S.list[tid] = prev;                            // evicted owner takes the slot, or -1 = hole
                hole = prev < 0;
            }
        }
        const int H = __syncthreads_count(hole);
        if (tid < nu) { S.bestkey[myj] = 0ull; S.winpos[myj] = 0x7fffffff; }     // :421-422
        if (tid == 0) S.tie = 0;
        // ---- push_all_left (:137-162): k-th hole left of the new count <- k-th live entry right of it
        const int new_nu = nu - H;
        if (H > 0 && new_nu > 0) {
            int v = 0;
            bool ishole = false;
            if (tid < nu) { v = S.list[tid]; ishole = v < 0; }
            int tot;
            const int before = small_scan_flag(S, ishole, tot);
            if (tid == new_nu) S.hsplit = before;
            __syncthreads();
            const int hsplit = S.hsplit;
            if (tid >= new_nu && tid < nu && !ishole) S.mover[(tid - new_nu) - (before - hsplit)] = v;
            __syncthreads();
            if (tid < new_nu && ishole) S.list[tid] = S.mover[before];
        }
        __syncthreads();
        nu = new_nu;
        ++its;
        }
        last_opt = -1;
        // ---- terminate() / eps-scaling (:275-292)
        if (its >= A.max_iter) break;
        if (nu == 0) {
            int viol = 0;
            for (int i = warp; i < N; i += SM_WARPS) { double cs; viol |= small_ece_row(S, i, lane, eps_t, tol, cs) ? 1 : 0; }
            viol = __syncthreads_or(viol);
            last_opt = viol ? 0 : 1;
            if (last_opt) break;
            if (eps < target) break;                           // :280-281
            eps = eps * theta;                                 // float32 product (:283)
            for (int j = tid; j < M; j += SM_THREADS) S.owner[j] = -1;
            for (int i = tid; i < N; i += SM_THREADS) { S.p2o[i] = -1; S.list[i] = i; }
            __syncthreads();
            nu = N;
            ++nred;
        }
    }
    // ---- meta (:297-304) and get_obj (:489-523)
    {
        int viol = 0;
        for (int i = warp; i < N; i += SM_WARPS) {
            double cs;
            const bool v = small_ece_row(S, i, lane, eps_t, tol, cs);
            viol |= v ? 1 : 0;
            if (lane == 0) S.chosen[i] = cs;
        }
        viol = __syncthreads_or(viol);
        if (last_opt < 0) last_opt = (nu == 0 && !viol) ? 1 : 0;
    }
    for (int i = tid; i < N; i += SM_THREADS) A.sol_out[i] = S.p2o[i];
    for (int j = tid; j < M; j += SM_THREADS) A.price_out[j] = S.price[j];
    if (tid == 0) {
        double obj = 0.0;
        for (int i = 0; i < N; ++i) {
            if (S.p2o[i] < 0) continue;
            if (A.negate) obj -= S.chosen[i]; else obj += S.chosen[i];
        }
        R.final_eps = eps; R.eCE = last_opt; R.soln_found = (nu == 0 && last_opt) ? 1 : 0;
        R.its = its; R.nreductions = nred; R.n_assigned = N - nu; R.obj64 = obj;
        *A.res = R;
    }
}

extern "C" size_t sslapb_small_smem_bytes() { return sizeof(SmallShared); }

extern "C" cudaError_t sslapb_launch_small(const SslapbSmallArgs *A, cudaStream_t stream)
{
    static bool configured[64] = {};                            // per device: the attribute belongs to the device's context
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(sslapb_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallShared));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    sslapb_small_kernel<<<1, SM_THREADS, sizeof(SmallShared), stream>>>(*A);
    return cudaGetLastError();
}
