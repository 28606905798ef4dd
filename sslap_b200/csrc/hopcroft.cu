// Hopcroft-Karp-style maximum-cardinality bipartite matching on the device (sm_100a).
//
// Replaces HopcroftKarpSolverCython (/root/reference/sslap/feasibility_.pyx:95-211).  The reference alternates a BFS
// that layers the graph from all free left vertices (:128-168) with a recursive DFS that augments along the layers
// (:170-197).  A DFS is a serial walk; with a handful of free vertices left it explores the whole level graph on one
// warp (measured: 55 ms for 8 paths on a 10 M edge graph).  Here both halves are level-synchronous and parallel:
//   * BFS: every right vertex is claimed by ONE left vertex of the frontier (atomicCAS on pred_v), so the search
//     forest is a set of vertex-disjoint trees, one per free left vertex; the first free right vertex a tree reaches
//     is recorded as that tree's endpoint (atomicCAS per root).  The phase stops after the level that found endpoints
//     (shortest augmenting paths, as in Hopcroft-Karp).
//   * augmentation: one thread per tree with an endpoint walks pred_v / pair_u back to the root and flips the path;
//     the trees are disjoint, so all flips are independent.
// A phase augments >= 1 path whenever an augmenting path exists (alternating BFS from all free left vertices reaches
// every right vertex that any alternating path reaches), so the loop ends exactly at a maximum matching (Berge):
// `size` is bit-exact with the reference; the pairings are a (generally different) valid maximum matching.
#include "common.cuh"
#include "build.cuh"

#define HK_INF 0x7fffffff

// Cheap maximal-matching initialisation: every left vertex grabs its first free neighbour (warp per vertex: the lanes scan
// the adjacency list together, the free neighbours are tried in adjacency order).
__global__ void __launch_bounds__(256) sslapb_hk_greedy_kernel(const long long *__restrict__ rowptr,
                                                               const int *__restrict__ cols, int N, int *pair_u,
                                                               int *pair_v, SslapbHkFlags *F)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    int got = 0;
    for (int u = gwarp; u < N; u += nwarps) {
        const long long st = rowptr[u], en = rowptr[u + 1];
        bool done = false;
        for (long long base = st; base < en && !done; base += 32) {
            const long long e = base + lane;
            int v = -1;
            bool fr = false;
            if (e < en) { v = cols[e]; fr = *(volatile int *)(pair_v + v) == -1; }
            unsigned cand = __ballot_sync(SSLAPB_FULL, fr);
            while (cand && !done) {
                const int l = __ffs(cand) - 1;
                cand &= cand - 1;
                int ok = 0;
                if (lane == l) ok = atomicCAS(pair_v + v, -1, u) == -1;
                ok = __shfl_sync(SSLAPB_FULL, ok, l);
                if (ok) { if (lane == l) pair_u[u] = v; done = true; ++got; }
            }
        }
    }
    if (got && lane == 0) atomicAdd(&F->matched, got);
}

__global__ void __launch_bounds__(256) sslapb_hk_phase_init_kernel(int N, int M, const int *__restrict__ pair_u,
                                                                   int *dist, int *root, int *end_of_root, int *pred_v,
                                                                   SslapbHkFlags *F)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int u = gtid; u < N; u += nth) {
        const bool fr = pair_u[u] == -1;
        dist[u] = fr ? 0 : HK_INF;                             // feasibility_.pyx:136-143
        root[u] = fr ? u : -1;
        end_of_root[u] = -1;
    }
    for (int v = gtid; v < M; v += nth) pred_v[v] = -1;
    if (gtid == 0) { F->found = 0; F->grew = 0; F->augmented = 0; }
}

// One BFS level (feasibility_.pyx:147-166): warp per left vertex of the current level; the lanes scan its adjacency list
// together.  A right vertex joins the tree of the first left vertex that claims it.
__global__ void __launch_bounds__(256) sslapb_hk_bfs_level_kernel(const long long *__restrict__ rowptr,
                                                                  const int *__restrict__ cols, int N, int level,
                                                                  const int *__restrict__ pair_v, int *dist, int *root,
                                                                  int *end_of_root, int *pred_v, SslapbHkFlags *F)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    int found = 0, grew = 0;
    for (int u = gwarp; u < N; u += nwarps) {
        if (dist[u] != level) continue;
        const int r = root[u];
        const long long st = rowptr[u], en = rowptr[u + 1];
        for (long long e = st + lane; e < en; e += 32) {
            const int v = cols[e];
            if (*(volatile int *)(pred_v + v) != -1) continue;
            if (atomicCAS(pred_v + v, -1, u) != -1) continue;  // somebody else's tree
            const int pu = pair_v[v];
            if (pu == -1) {                                    // free right vertex: endpoint of tree r (first one wins)
                if (atomicCAS(end_of_root + r, -1, v) == -1) found = 1;
            } else {                                           // matched: its partner joins the next level of tree r
                dist[pu] = level + 1;
                root[pu] = r;
                grew = 1;
            }
        }
    }
    if (found) F->found = 1;
    if (grew) F->grew = 1;
}

// Flip the tree path root -> ... -> endpoint of every tree that found one (feasibility_.pyx:189-190).
__global__ void __launch_bounds__(256) sslapb_hk_augment_kernel(int N, const int *__restrict__ end_of_root,
                                                                const int *__restrict__ pred_v, int *pair_u, int *pair_v,
                                                                SslapbHkFlags *F)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    int wins = 0;
    for (int r = gtid; r < N; r += nth) {
        int v = end_of_root[r];
        if (v < 0) continue;
        for (;;) {
            const int u = pred_v[v];
            const int old_v = pair_u[u];
            pair_u[u] = v;
            pair_v[v] = u;
            if (u == r) break;
            v = old_v;
        }
        ++wins;
    }
    if (wins) atomicAdd(&F->augmented, wins);
}

extern "C" cudaError_t sslapb_hk_launch_greedy(const long long *rowptr, const int *cols, int N, int *pair_u, int *pair_v,
                                               SslapbHkFlags *F, int sms, cudaStream_t s)
{
    sslapb_hk_greedy_kernel<<<sms * 8, 256, 0, s>>>(rowptr, cols, N, pair_u, pair_v, F);
    return cudaGetLastError();
}
extern "C" cudaError_t sslapb_hk_launch_phase_init(int N, int M, const int *pair_u, int *dist, int *root, int *end_of_root,
                                                   int *pred_v, SslapbHkFlags *F, int sms, cudaStream_t s)
{
    sslapb_hk_phase_init_kernel<<<sms * 4, 256, 0, s>>>(N, M, pair_u, dist, root, end_of_root, pred_v, F);
    return cudaGetLastError();
}
extern "C" cudaError_t sslapb_hk_launch_bfs_level(const long long *rowptr, const int *cols, int N, int level,
                                                  const int *pair_v, int *dist, int *root, int *end_of_root, int *pred_v,
                                                  SslapbHkFlags *F, int sms, cudaStream_t s)
{
    sslapb_hk_bfs_level_kernel<<<sms * 8, 256, 0, s>>>(rowptr, cols, N, level, pair_v, dist, root, end_of_root, pred_v, F);
    return cudaGetLastError();
}
extern "C" cudaError_t sslapb_hk_launch_augment(int N, const int *end_of_root, const int *pred_v, int *pair_u, int *pair_v,
                                                SslapbHkFlags *F, int sms, cudaStream_t s)
{
    sslapb_hk_augment_kernel<<<sms * 4, 256, 0, s>>>(N, end_of_root, pred_v, pair_u, pair_v, F);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------------
// Device-resident phase loop (feasibility_.pyx:199-211): ONE cooperative launch runs the greedy initialisation and every
// BFS level / augmentation of every phase, with grid barriers instead of a host read-back per level (round 1 paid one
// cudaStreamSynchronize per BFS level: 37 ms for a 100k-vertex deficient graph that needs ~10^3 levels in total).
// The BFS keeps explicit frontier queues (a level costs O(frontier), not O(N)); the search forest, the endpoint rule and
// the path flips are those of the kernels above, so the result is the same maximum matching size.
// ----------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool hk_grid_barrier(SslapbHkCtrl *c, unsigned nblk, unsigned &epoch)
{
    __shared__ int s_ab;
    epoch += nblk;
    __syncthreads();
    if (threadIdx.x == 0) {
        int ab = 0;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(&c->bar) : "memory");
        unsigned polls = 0;
        unsigned long long t0 = 0;
        while ((int)(sslapb_ld_acquire_u32(&c->bar) - epoch) < 0) {
            if (++polls < 4096) continue;
            if (polls == 4096) t0 = sslapb_globaltimer();
            if (sslapb_ld_volatile_s32(&c->watchdog)) { ab = 1; break; }
            __nanosleep(200);
            if (sslapb_globaltimer() - t0 > 20000000000ull) { *(volatile int *)&c->watchdog = 1; ab = 1; break; }
        }
        s_ab = ab | sslapb_ld_volatile_s32(&c->watchdog);
    }
    __syncthreads();
    return s_ab == 0;
}

#define HK_THREADS 512
__global__ void __launch_bounds__(HK_THREADS, 2) sslapb_hk_persistent_kernel(const long long *__restrict__ rowptr,
                                                                            const int *__restrict__ cols, int N, int M,
                                                                            int *pair_u, int *pair_v, int *root,
                                                                            int *end_of_root, int *pred_v, int *q0, int *q1,
                                                                            int *q2, int *roots, SslapbHkCtrl *C)
{
    const int lane = threadIdx.x & 31;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int gwarp = gtid >> 5, nwarps = nth >> 5;
    const unsigned nblk = gridDim.x;
    unsigned epoch = 0;
    int *const Q[3] = {q0, q1, q2};
    // ---- greedy maximal matching (same rule as sslapb_hk_greedy_kernel)
    {
        int got = 0;
        for (int u = gwarp; u < N; u += nwarps) {
            const long long st = rowptr[u], en = rowptr[u + 1];
            bool done = false;
            for (long long base = st; base < en && !done; base += 32) {
                const long long e = base + lane;
                int v = -1;
                bool fr = false;
                if (e < en) { v = cols[e]; fr = *(volatile int *)(pair_v + v) == -1; }
                unsigned cand = __ballot_sync(SSLAPB_FULL, fr);
                while (cand && !done) {
                    const int l = __ffs(cand) - 1;
                    cand &= cand - 1;
                    int ok = 0;
                    if (lane == l) ok = atomicCAS(pair_v + v, -1, u) == -1;
                    ok = __shfl_sync(SSLAPB_FULL, ok, l);
                    if (ok) { if (lane == l) pair_u[u] = v; done = true; ++got; }
                }
            }
        }
        if (got && lane == 0) atomicAdd(&C->matched, got);
    }
    if (!hk_grid_barrier(C, nblk, epoch)) return;
    const int bound = N < M ? N : M;
    for (;;) {
        if (*(volatile int *)&C->matched >= bound) break;
        // ---- phase init: free left vertices -> level 0 and the list of roots; pred_v cleared
        for (int base = blockIdx.x * blockDim.x; base < N; base += nth) {
            const int u = base + threadIdx.x;
            const bool fr = u < N && pair_u[u] == -1;
            const unsigned bal = __ballot_sync(SSLAPB_FULL, fr);
            int pos = 0;
            if (lane == 0 && bal) pos = atomicAdd(&C->cnt[0], __popc(bal));
            pos = __shfl_sync(SSLAPB_FULL, pos, 0) + __popc(bal & ((1u << lane) - 1u));
            if (fr) { root[u] = u; end_of_root[u] = -1; q0[pos] = u; roots[pos] = u; }
        }
        for (int v = gtid; v < M; v += nth) pred_v[v] = -1;
        if (!hk_grid_barrier(C, nblk, epoch)) return;
        const int nroots = *(volatile int *)&C->cnt[0];
        if (gtid == 0) { C->nroots = nroots; C->phases += 1; }
        // ---- level-synchronous BFS over explicit frontiers (feasibility_.pyx:128-168)
        int level = 0, found = 0;
        for (;;) {
            const int ncur = *(volatile int *)&C->cnt[level % 3];
            const int *qc = Q[level % 3];
            int *qn = Q[(level + 1) % 3];
            int *cn = &C->cnt[(level + 1) % 3];
            if (gtid == 0) C->cnt[(level + 2) % 3] = 0;        // the level after next: nobody reads or writes it now
            for (int idx = gwarp; idx < ncur; idx += nwarps) {
                const int u = qc[idx];
                const int r = root[u];
                const long long st = rowptr[u], en = rowptr[u + 1];
                for (long long e = st + lane; e < en; e += 32) {
                    const int v = cols[e];
                    if (*(volatile int *)(pred_v + v) != -1) continue;
                    if (atomicCAS(pred_v + v, -1, u) != -1) continue;      // somebody else's tree
                    const int pu = pair_v[v];
                    if (pu == -1) {                                        // free right vertex: endpoint of tree r (first wins)
                        // `found` carries the level (+1): a CTA that is already one level ahead must not make a slower
                        // CTA, which has not read the flag for the previous level yet, leave the loop one level early
                        if (atomicCAS(end_of_root + r, -1, v) == -1) *(volatile int *)&C->found = level + 1;
                    } else {                                               // matched: its partner joins the next level of tree r
                        root[pu] = r;
                        qn[atomicAdd(cn, 1)] = pu;
                    }
                }
            }
            if (!hk_grid_barrier(C, nblk, epoch)) return;
            const int fl = *(volatile int *)&C->found;
            found = fl != 0 && fl <= level + 1;                // an endpoint was found at this level (or, impossible, an earlier one)
            const int nnext = *(volatile int *)cn;
            ++level;
            if (found || nnext == 0) break;
        }
        if (gtid == 0) C->levels += level;
        if (!found) break;                                     // no augmenting path: maximum (feasibility_.pyx:202-203)
        // ---- flip the tree path of every root that found an endpoint (feasibility_.pyx:189-190)
        int wins = 0;
        for (int k = gtid; k < nroots; k += nth) {
            const int r = roots[k];
            int v = end_of_root[r];
            if (v < 0) continue;
            for (;;) {
                const int u = pred_v[v];
                const int old_v = pair_u[u];
                pair_u[u] = v;
                pair_v[v] = u;
                if (u == r) break;
                v = old_v;
            }
            ++wins;
        }
        const unsigned wb = __ballot_sync(SSLAPB_FULL, wins > 0);
        if (wins) atomicAdd(&C->matched, wins);
        (void)wb;
        if (!hk_grid_barrier(C, nblk, epoch)) return;
        if (gtid == 0) { C->found = 0; C->cnt[0] = 0; C->cnt[1] = 0; C->cnt[2] = 0; }
        if (!hk_grid_barrier(C, nblk, epoch)) return;
    }
}

extern "C" cudaError_t sslapb_hk_persistent_grid(int device, int *grid)
{
    int per_sm = 0, sms = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sslapb_hk_persistent_kernel, HK_THREADS, 0);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid = sms * (per_sm > 2 ? 2 : per_sm);
    return cudaSuccess;
}

extern "C" cudaError_t sslapb_hk_launch_persistent(const long long *rowptr, const int *cols, int N, int M, int *pair_u,
                                                   int *pair_v, int *root, int *end_of_root, int *pred_v, int *q0, int *q1,
                                                   int *q2, int *roots, void *ctrl, int grid, cudaStream_t s)
{
    SslapbHkCtrl *C = reinterpret_cast<SslapbHkCtrl *>(ctrl);
    void *args[] = {(void *)&rowptr, (void *)&cols, (void *)&N, (void *)&M, (void *)&pair_u, (void *)&pair_v, (void *)&root,
                    (void *)&end_of_root, (void *)&pred_v, (void *)&q0, (void *)&q1, (void *)&q2, (void *)&roots, (void *)&C};
    return cudaLaunchCooperativeKernel((const void *)sslapb_hk_persistent_kernel, dim3(grid), dim3(HK_THREADS), args, 0, s);
}
