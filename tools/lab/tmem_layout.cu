// Lab probe (not part of the product): register <-> (lane, column) mapping of the tcgen05.ld/st shapes.
#include "../../s3od_b200/csrc/common.cuh"
#include <cstdio>
using namespace s3od;
__global__ void probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<64>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + ((uint32_t)(warp * 32) << 16);
  uint32_t r[32];
  for (int i = 0; i < 32; ++i) r[i] = (warp * 32 + lane) * 256 + i;      // value = lane * 256 + column
  tmem_st_32x32(tm, r);
  tmem_st_wait();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 1) {
    uint32_t a[4], b[2], c[2], d[4];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(tm + 8));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(tm + 8));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(c[0]), "=r"(c[1]) : "r"(tm + 8 + (16u << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(tm + 8 + (16u << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 4; ++i) out[lane * 12 + i] = a[i];
    for (int i = 0; i < 2; ++i) out[lane * 12 + 4 + i] = b[i];
    for (int i = 0; i < 2; ++i) out[lane * 12 + 6 + i] = c[i];
    for (int i = 0; i < 4; ++i) out[lane * 12 + 8 + i] = d[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(slot); }
}
int main() {
  uint32_t* d; cudaMalloc(&d, 32 * 12 * 4);
  probe<<<1, 128>>>(d);
  uint32_t h[32 * 12]; cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%s\nwarp 1 (lanes 32..63), column base 8. entries are lane.column\n", cudaGetErrorString(e));
  const char* nm[4] = {"16x256b.x1      ", "16x128b.x1      ", "16x128b.x1 +16ln", "16x256b.x1 +16ln"};
  const int off[4] = {0, 4, 6, 8}, n[4] = {4, 2, 2, 4};
  for (int k = 0; k < 4; ++k) {
    printf("%s\n", nm[k]);
    for (int t = 0; t < 32; ++t) {
      printf("  t%02d:", t);
      for (int i = 0; i < n[k]; ++i) printf(" %2u.%-2u", h[t * 12 + off[k] + i] >> 8, h[t * 12 + off[k] + i] & 255);
      if (t % 4 == 3) printf("\n");
    }
  }
  return 0;
}
