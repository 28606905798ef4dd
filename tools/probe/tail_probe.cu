// Probe: the latency floor of one single-bidder ("chain") round of the auction tail on this GPU, by parts.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tail_probe tail_probe.cu && ./tail_probe
// One warp, one SM, dependent steps exactly as chain_rounds_hot chains them:
//   (a) hot row: 32 lanes x 16 B from a 51 MB table (row index depends on the previous step)
//   (b) record gather: 32 lanes x 32 B (LDG.256) from a 3.2 MB table, addresses from (a)
//   (c) the reduction that names the next row: REDUX + ballot + ffs + shuffle
// and the same with everything resident in L2 (second pass over the same random walk) or not (first pass, table flushed).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
struct __align__(32) Rec { unsigned long long start, owner_deg, price, pad; };
__device__ __forceinline__ Rec ldrec(const Rec *p) { Rec r; asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.start), "=l"(r.owner_deg), "=l"(r.price), "=l"(r.pad) : "l"(p)); return r; }

// mode bits: 1 = hot row load, 2 = record gather, 4 = reduction chain
template <int MODE>
__global__ void __launch_bounds__(32) chain_kernel(const int4 *hot, const Rec *rec, int nrows, int nobj, int iters, int start, long long *out)
{
    const int lane = threadIdx.x;
    int row = start;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        int col = (row * 37 + lane * 101) % nobj;
        if (MODE & 1) { const int4 h = __ldg(hot + (long long)row * 32 + lane); col = h.x; acc += h.y; }
        unsigned nxt = (unsigned)col * 2654435761u + it;
        if (MODE & 2) { const Rec r = ldrec(rec + col); nxt = (unsigned)r.owner_deg; acc += (unsigned)r.price; }
        if (MODE & 4) {
            const unsigned m = __reduce_max_sync(0xffffffffu, nxt >> 8);
            const unsigned b = __ballot_sync(0xffffffffu, (nxt >> 8) == m);
            const int src = __ffs(b) - 1;
            nxt = __shfl_sync(0xffffffffu, nxt, src);
        } else {
            nxt = __shfl_sync(0xffffffffu, nxt, it & 31);
        }
        row = (int)(nxt % (unsigned)nrows);
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = acc + row; }
}

int main()
{
    const int nrows = 100000, nobj = 100000, iters = 20000;
    std::vector<int4> hot((size_t)nrows * 32);
    std::vector<Rec> rec(nobj);
    srand(1);
    for (size_t k = 0; k < hot.size(); ++k) hot[k] = make_int4(rand() % nobj, rand(), rand(), rand());
    for (int j = 0; j < nobj; ++j) { rec[j].start = rand(); rec[j].owner_deg = ((unsigned long long)rand() << 16) ^ rand(); rec[j].price = rand(); rec[j].pad = 0; }
    int4 *dh; Rec *dr; long long *dout; char *flush;
    CK(cudaMalloc(&dh, hot.size() * 16)); CK(cudaMalloc(&dr, rec.size() * 32)); CK(cudaMalloc(&dout, 64)); CK(cudaMalloc(&flush, 512 << 20));
    CK(cudaMemcpy(dh, hot.data(), hot.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dr, rec.data(), rec.size() * 32, cudaMemcpyHostToDevice));
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SM clock (attr) %.0f MHz; %d dependent rounds by one warp\n", clk / 1e3, iters);
    auto run = [&](const char *name, void (*k)(const int4 *, const Rec *, int, int, int, int, long long *), bool cold) {
        for (int rep = 0; rep < 2; ++rep) {
            if (cold) CK(cudaMemset(flush, rep, 512 << 20));
            k<<<1, 32>>>(dh, dr, nrows, nobj, iters, 17, dout);
            CK(cudaDeviceSynchronize());
            long long o[2]; CK(cudaMemcpy(o, dout, 16, cudaMemcpyDeviceToHost));
            printf("%-64s %s rep%d  %7.1f cycles/round\n", name, cold ? "L2 flushed before" : "warm (same walk)  ", rep, (double)o[0] / iters);
        }
    };
    run("shuffle only (loop overhead)", chain_kernel<0>, false);
    run("reduction chain only (REDUX + ballot + ffs + shfl)", chain_kernel<4>, false);
    run("hot row (32 x 16 B) + shfl", chain_kernel<1>, true);
    run("hot row (32 x 16 B) + shfl", chain_kernel<1>, false);
    run("record gather (32 x 32 B, 3.2 MB table) + shfl", chain_kernel<2>, true);
    run("record gather (32 x 32 B, 3.2 MB table) + shfl", chain_kernel<2>, false);
    run("hot row -> record gather -> shfl", chain_kernel<3>, true);
    run("hot row -> record gather -> shfl", chain_kernel<3>, false);
    run("hot row -> record gather -> reduction chain (a full round's skeleton)", chain_kernel<7>, true);
    run("hot row -> record gather -> reduction chain (a full round's skeleton)", chain_kernel<7>, false);
    return 0;
}
