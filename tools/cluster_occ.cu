// How many clusters of a given size are co-resident for a CTA of 352 threads and ~185 kB of shared memory (one CTA per SM):
//   nvcc -arch=sm_100a -o /tmp/cluster_occ tools/cluster_occ.cu && /tmp/cluster_occ
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(352, 1) dummy(int* p) { extern __shared__ int s[]; if (p) p[0] = s[threadIdx.x]; }
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s: %d SMs\n", pr.name, pr.multiProcessorCount);
    const int smem = 185 * 1024;
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs = 1; cs <= 16; cs++) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(cs, 148); cfg.blockDim = dim3(352); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
        printf("cluster size %2d: %3d clusters = %3d SMs  (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
        cudaGetLastError();
    }
    return 0;
}
