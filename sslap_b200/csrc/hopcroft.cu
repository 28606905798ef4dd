// Hopcroft-Karp maximum-cardinality bipartite matching on the device (sm_100a).
//
// Replaces HopcroftKarpSolverCython (/root/reference/sslap/feasibility_.pyx:95-211): level-synchronous frontier BFS from
// all free left vertices (:128-168) followed by parallel, vertex-disjoint augmentation along the level graph (:170-197,
// recursive DFS in the reference; iterative with an atomic claim per right vertex here).  The cardinality of a maximum
// matching is unique, so `size` is bit-exact with the reference; the pairings are a (generally different) valid
// maximum matching.
//
// Why the parallel augmentation always makes progress: a right vertex is claimed (atomicExch on visited[]) only after the
// level test passed, so a claim can only block searches that arrive from the same BFS level; a search blocked by a
// claim blames a search working strictly deeper in the level graph, and the deepest level ends in free right vertices
// whose first claimant completes its path.  Hence every phase in which BFS reached a free right vertex augments >= 1
// path, and the loop ends exactly when no augmenting path exists (Berge) — i.e. at a maximum matching.
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
                                                                   int *dist, int *visited, SslapbHkFlags *F)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int u = gtid; u < N; u += nth) dist[u] = (pair_u[u] == -1) ? 0 : HK_INF;   // feasibility_.pyx:136-143
    for (int v = gtid; v < M; v += nth) visited[v] = 0;
    if (gtid == 0) { F->found = 0; F->grew = 0; F->augmented = 0; }
}

// One BFS level (feasibility_.pyx:147-166): warp per left vertex of the current level.
__global__ void __launch_bounds__(256) sslapb_hk_bfs_level_kernel(const long long *__restrict__ rowptr,
                                                                  const int *__restrict__ cols, int N, int level,
                                                                  const int *__restrict__ pair_v, int *dist,
                                                                  SslapbHkFlags *F)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    int found = 0, grew = 0;
    for (int u = gwarp; u < N; u += nwarps) {
        if (dist[u] != level) continue;
        const long long st = rowptr[u], en = rowptr[u + 1];
        for (long long e = st + lane; e < en; e += 32) {
            const int pu = pair_v[cols[e]];
            if (pu == -1) found = 1;
            else if (dist[pu] == HK_INF) { dist[pu] = level + 1; grew = 1; }   // same value from every writer
        }
    }
    if (found) F->found = 1;
    if (grew) F->grew = 1;
}

// Augmentation: one WARP per free left vertex walks the level graph depth-first; the 32 lanes scan the adjacency list of
// the current vertex together (coalesced), the eligible neighbours are claimed one at a time in adjacency order (a claim
// that is not used would block other searches for nothing).  Left vertices are reached only through their (claimed)
// partner, so cursor[]/pred[] entries are private to the walking warp.
__global__ void __launch_bounds__(128) sslapb_hk_augment_kernel(const long long *__restrict__ rowptr,
                                                                const int *__restrict__ cols, int N, int dist_nil,
                                                                int *pair_u, int *pair_v, int *dist, int *visited,
                                                                long long *cursor, int *pred, SslapbHkFlags *F)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    int wins = 0;
    for (int root = gwarp; root < N; root += nwarps) {
        if (*(volatile int *)(dist + root) != 0) continue;     // free left vertices are exactly the level-0 ones
        int cur = root;
        if (lane == 0) { cursor[cur] = rowptr[cur]; pred[cur] = -1; }
        __syncwarp();
        for (;;) {
            const long long en = rowptr[cur + 1];
            const int want = *(volatile int *)(dist + cur) + 1;
            long long base = *(volatile long long *)(cursor + cur);
            int next = -2, via = -1;                           // -2 none, -1 free right vertex, >= 0 left vertex to descend to
            long long resume = en;
            while (base < en && next == -2) {
                const long long e = base + lane;
                int v = -1, pu = -1;
                bool ok = false;
                if (e < en) {
                    v = cols[e];
                    pu = *(volatile int *)(pair_v + v);
                    const int dpu = (pu == -1) ? dist_nil : *(volatile int *)(dist + pu);
                    ok = (dpu == want) && (*(volatile int *)(visited + v) == 0);   // feasibility_.pyx:186
                }
                unsigned cand = __ballot_sync(SSLAPB_FULL, ok);
                while (cand && next == -2) {                   // claim in adjacency order, one at a time
                    const int l = __ffs(cand) - 1;
                    cand &= cand - 1;
                    int got = 0;
                    if (lane == l) got = (atomicExch(visited + v, 1) == 0);
                    got = __shfl_sync(SSLAPB_FULL, got, l);
                    if (got) {
                        next = __shfl_sync(SSLAPB_FULL, pu, l);
                        via = __shfl_sync(SSLAPB_FULL, v, l);
                        resume = base + l + 1;
                    }
                }
                base += 32;
            }
            if (lane == 0) cursor[cur] = resume;
            if (next == -2) {                                  // dead end (:194)
                if (lane == 0) dist[cur] = HK_INF;
                if (cur == root) break;
                cur = pred[cur];
                __syncwarp();
                continue;
            }
            if (next == -1) {                                  // free right vertex: flip the path (:189-190)
                if (lane == 0) {
                    int u = cur, v = via;
                    for (;;) {
                        const int old_v = pair_u[u];
                        pair_u[u] = v;
                        pair_v[v] = u;
                        if (u == root) break;
                        v = old_v;
                        u = pred[u];
                    }
                }
                __syncwarp();
                ++wins;
                break;
            }
            if (lane == 0) { pred[next] = cur; cursor[next] = rowptr[next]; }
            __syncwarp();
            cur = next;
        }
    }
    if (wins && lane == 0) atomicAdd(&F->augmented, wins);
}

extern "C" cudaError_t sslapb_hk_launch_greedy(const long long *rowptr, const int *cols, int N, int *pair_u, int *pair_v,
                                               SslapbHkFlags *F, int sms, cudaStream_t s)
{
    sslapb_hk_greedy_kernel<<<sms * 8, 256, 0, s>>>(rowptr, cols, N, pair_u, pair_v, F);
    return cudaGetLastError();
}
extern "C" cudaError_t sslapb_hk_launch_phase_init(int N, int M, const int *pair_u, int *dist, int *visited,
                                                   SslapbHkFlags *F, int sms, cudaStream_t s)
{
    sslapb_hk_phase_init_kernel<<<sms * 4, 256, 0, s>>>(N, M, pair_u, dist, visited, F);
    return cudaGetLastError();
}
extern "C" cudaError_t sslapb_hk_launch_bfs_level(const long long *rowptr, const int *cols, int N, int level,
                                                  const int *pair_v, int *dist, SslapbHkFlags *F, int sms,
                                                  cudaStream_t s)
{
    sslapb_hk_bfs_level_kernel<<<sms * 8, 256, 0, s>>>(rowptr, cols, N, level, pair_v, dist, F);
    return cudaGetLastError();
}
extern "C" cudaError_t sslapb_hk_launch_augment(const long long *rowptr, const int *cols, int N, int dist_nil, int *pair_u,
                                                int *pair_v, int *dist, int *visited, long long *cursor, int *pred,
                                                SslapbHkFlags *F, int sms, cudaStream_t s)
{
    sslapb_hk_augment_kernel<<<sms * 8, 128, 0, s>>>(rowptr, cols, N, dist_nil, pair_u, pair_v, dist, visited, cursor,
                                                     pred, F);
    return cudaGetLastError();
}
