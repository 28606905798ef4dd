// Feasibility probe: cooperative + cluster launch of a 512-thread persistent kernel, and the latency of a cluster barrier.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(512, 1) probe_kernel(long long *out, int iters)
{
    cg::cluster_group cl = cg::this_cluster();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (t1 - t0) / iters; out[1] = cl.num_blocks(); }
}

int main()
{
    int dev = 0, sms = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    printf("SMs %d\n", sms);
    long long *d; cudaMalloc(&d, 64);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 0;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 2;
        if (cs > 8) cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int ncl = 0;
        cfg.gridDim = dim3(sms / cs * cs);
        cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, probe_kernel, &cfg);
        printf("cluster %2d: max active clusters %d (%s) -> %d CTAs\n", cs, ncl, cudaGetErrorString(e), ncl * cs);
        if (e != cudaSuccess || ncl == 0) { cudaGetLastError(); continue; }
        int grid = ncl * cs; if (grid > sms / cs * cs) grid = sms / cs * cs;
        cfg.gridDim = dim3(grid);
        int iters = 2000;
        e = cudaLaunchKernelEx(&cfg, probe_kernel, d, iters);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("   coop+cluster launch grid %d: %s / %s ; barrier = %lld cycles, cluster blocks %lld\n", grid, cudaGetErrorString(e), cudaGetErrorString(e2), h[0], h[1]);
        cudaGetLastError();
    }
    return 0;
}
