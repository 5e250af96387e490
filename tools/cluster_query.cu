// nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_query tools/cluster_query.cu
// How many clusters of size 2 / 4 / 8 of a (512 threads, 218 KB smem) CTA can be resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) dummy(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  const int smem = 218 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: max active clusters %d (%d SMs)  %s\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
