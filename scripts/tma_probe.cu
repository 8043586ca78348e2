// Stand-alone probe for the TMA 3-D box load used by k_wavefront (debug aid; not part of the product).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int B>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap *gmap, int use_g, int cz, int cy, int cx, uint32_t *out, int coop) {
  __shared__ __align__(128) uint32_t buf[B * B * B];
  __shared__ __align__(8) uint64_t mbar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(B * B * B * 4) : "memory");
    const CUtensorMap *d = use_g ? gmap : &tmap;
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(buf)),
                 "l"((unsigned long long)d), "r"(cz), "r"(cy), "r"(cx), "r"(smem_u32(&mbar))
                 : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&mbar)), "r"(0) : "memory");
  if (coop) cg::this_grid().sync();
  for (int k = threadIdx.x; k < B * B * B; k += blockDim.x) out[k] = buf[k];
}
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int B>
int run(int gz, int gy, int gx, int pz, int use_g, int coop, int cz, int cy, int cx) {
  size_t n = (size_t)gx * gy * pz;
  std::vector<uint32_t> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (uint32_t)i + 1;
  uint32_t *d, *o;
  cudaMalloc(&d, n * 4); cudaMalloc(&o, B * B * B * 4);
  cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)gz, (cuuint64_t)gy, (cuuint64_t)gx};
  cuuint64_t str[2] = {(cuuint64_t)pz * 4, (cuuint64_t)pz * gy * 4};
  cuuint32_t box[3] = {B, B, B}, es[3] = {1, 1, 1};
  CUresult r = ((PFN)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("  encode failed %d\n", (int)r); return 1; }
  CUtensorMap *gm; cudaMalloc(&gm, sizeof(tm)); cudaMemcpy(gm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  cudaError_t e;
  if (coop) {
    void *args[] = {&tm, &gm, &use_g, &cz, &cy, &cx, &o, &coop};
    e = cudaLaunchCooperativeKernel((void *)probe<B>, dim3(4), dim3(128), args, 0, 0);
  } else {
    probe<B><<<4, 128>>>(tm, gm, use_g, cz, cy, cx, o, coop);
    e = cudaGetLastError();
  }
  cudaError_t e2 = cudaDeviceSynchronize();
  if (e != cudaSuccess || e2 != cudaSuccess) { printf("  FAIL launch=%s sync=%s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); return 1; }
  std::vector<uint32_t> ho(B * B * B);
  cudaMemcpy(ho.data(), o, B * B * B * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int x = 0; x < B; ++x) for (int y = 0; y < B; ++y) for (int z = 0; z < B; ++z) {
    int X = cx + x, Y = cy + y, Z = cz + z;
    uint32_t exp = (X < 0 || Y < 0 || Z < 0 || X >= gx || Y >= gy || Z >= gz) ? 0u : h[((size_t)X * gy + Y) * pz + Z];
    if (ho[(x * B + y) * B + z] != exp) ++bad;
  }
  printf("  ok, mismatches=%d\n", bad);
  return bad != 0;
}
int main(int argc, char **argv) {
  int v = argc > 1 ? atoi(argv[1]) : 0;
  switch (v) {
    case 0: printf("B=12 param desc, plain launch, interior\n"); return run<12>(64, 64, 64, 64, 0, 0, 6, 6, 6);
    case 1: printf("B=12 param desc, plain launch, negative coords\n"); return run<12>(64, 64, 64, 64, 0, 0, -2, -2, -2);
    case 2: printf("B=12 global desc, plain launch\n"); return run<12>(64, 64, 64, 64, 1, 0, 6, 6, 6);
    case 3: printf("B=12 param desc, cooperative\n"); return run<12>(64, 64, 64, 64, 0, 1, 6, 6, 6);
    case 4: printf("B=16 param desc, plain\n"); return run<16>(64, 64, 64, 64, 0, 0, 6, 6, 6);
    case 5: printf("B=8 param desc, plain\n"); return run<8>(64, 64, 64, 64, 0, 0, 6, 6, 6);
    case 6: printf("B=12 gz=33 pz=36 negative + high OOB\n"); return run<12>(33, 33, 33, 36, 0, 0, 26, 26, -2);
    case 7: printf("B=16 global desc\n"); return run<16>(64, 64, 64, 64, 1, 0, 6, 6, 6);
    case 8: printf("B=8 global desc, aligned coords\n"); return run<8>(64, 64, 64, 64, 1, 0, 8, 8, 8);
  }
  return 0;
}
