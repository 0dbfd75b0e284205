// How many clusters of each size can be co-resident with one fat CTA per SM (200 KB of shared memory)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int sm[]; if (p) p[0] = sm[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int smems[3] = {64 * 1024, 120 * 1024, 204 * 1024};
  for (int si = 0; si < 3; ++si)
    for (int cs = 1; cs <= 16; cs *= 2) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smems[si];
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %3d KB cluster %2d: max active clusters %d (%d CTAs) %s\n", smems[si] / 1024, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
