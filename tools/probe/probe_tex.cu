// probe: which CUDA array flag combinations accept tld4 (texture gather) on this driver
#include <cuda_runtime.h>
#include <stdio.h>
int main() {
  cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
  int n = 32;
  struct { const char *name; unsigned flags; int depth; } tries[] = {
      {"layered|gather", cudaArrayLayered | cudaArrayTextureGather, n},
      {"layered", cudaArrayLayered, n},
      {"gather 2D (depth 0)", cudaArrayTextureGather, 0},
      {"3D default", cudaArrayDefault, n},
  };
  for (auto &t : tries) {
    cudaArray_t a = nullptr;
    cudaError_t e = cudaMalloc3DArray(&a, &fd, make_cudaExtent(n, n, t.depth), t.flags);
    printf("%-24s -> %s\n", t.name, cudaGetErrorString(e));
    cudaGetLastError();
    if (a) cudaFreeArray(a);
  }
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("maxTexture2DGather %d %d  maxTexture2DLayered %d %d %d  l2 %d MB  smem/blk optin %zu  sms %d\n", p.maxTexture2DGather[0], p.maxTexture2DGather[1],
         p.maxTexture2DLayered[0], p.maxTexture2DLayered[1], p.maxTexture2DLayered[2], p.l2CacheSize >> 20, p.sharedMemPerBlockOptin, p.multiProcessorCount);
  return 0;
}
