// How many thread-block clusters of size 2 / 4 / 8 with one 227 KB CTA per SM can be co-resident on this GPU?
// (decides whether a persistent 4-CTA-cluster GEMM would strand SMs)   nvcc -arch=sm_100a -o cluster_occupancy cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main()
{
    const int smem = 230000;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s SMs=%d\n", pr.name, pr.multiProcessorCount);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 148); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: max active clusters %d -> %d SMs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
    }
    return 0;
}
