// Stand-alone check of the TMA path used by InterKernel: a 2-D byte tensor, a 32x21 box at an arbitrary
// (x, y), descriptor read from GLOBAL memory (mode 0) or from a __grid_constant__ parameter (mode 1).
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

struct alignas(128) Map { unsigned char b[128]; };

__device__ __forceinline__ unsigned Smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// stage 1: barrier only (expect 0 bytes); stage 2: 1-D bulk copy of 672 bytes; (the tensor load is Probe below)
__global__ void ProbeStage(int stage, const unsigned char *src, unsigned char *out) {
  __shared__ __align__(128) unsigned char tile[21 * 32];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(Smem(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    const unsigned bytes = stage == 1 ? 0 : 21 * 32;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(Smem(&bar)), "r"(bytes) : "memory");
    if (stage == 2)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(Smem(tile)), "l"(src),
                   "r"(bytes), "r"(Smem(&bar))
                   : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(Smem(&bar))
      : "memory");
  for (int i = threadIdx.x; i < 21 * 32; i += 32) out[i] = tile[i];
}

template <int MODE>
__global__ void Probe(const Map *gmap, const __grid_constant__ Map pmap, int x, int y, unsigned char *out, int fence) {
  __shared__ __align__(128) unsigned char tile[21 * 32];
  __shared__ unsigned long long bar;
  const Map *m = MODE == 0 ? gmap : &pmap;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(Smem(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    if (fence) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(m) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(Smem(&bar)), "r"(21 * 32) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(Smem(tile)),
                 "l"(m), "r"(x), "r"(y), "r"(Smem(&bar))
                 : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(Smem(&bar))
      : "memory");
  for (int i = threadIdx.x; i < 21 * 32; i += 32) out[i] = tile[i];
}

int main() {
  const int W = 240, H = 208;
  unsigned char *h = new unsigned char[W * H];
  for (int i = 0; i < W * H; ++i) h[i] = (unsigned char)((i * 7 + (i / W) * 13) & 0xff);
  unsigned char *d, *dout;
  cudaMalloc(&d, W * H);
  cudaMalloc(&dout, 21 * 32);
  cudaMemcpy(d, h, W * H, cudaMemcpyHostToDevice);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
  Map hm;
  cuuint64_t dims[2] = {W, H}, strides[1] = {W};
  cuuint32_t box[2] = {32, 21}, es[2] = {1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap *>(&hm), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  // drivers up to CUDA 13.1 set a descriptor bit for tensors below 128 KB that makes the load trap (the same
  // work-around as in CUTLASS, cute/atom/copy_traits_sm90_tma.hpp)
  int drv = 0;
  cudaDriverGetVersion(&drv);
  printf("driver %d\n", drv);
  if (drv <= 13010 && (size_t)W * H < 131072) reinterpret_cast<uint64_t *>(&hm)[1] &= ~(1ull << 21);
  Map *dm;
  cudaMalloc(&dm, sizeof(Map));
  cudaMemcpy(dm, &hm, sizeof(Map), cudaMemcpyHostToDevice);
  for (int stage = 1; stage <= 2; ++stage) {
    ProbeStage<<<1, 32>>>(stage, d, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("stage %d: %s\n", stage, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  unsigned char got[21 * 32];
  for (int mode = 1; mode >= 0; --mode)
    for (int fence = 0; fence < 2; ++fence) {
      if (mode == 1 && fence) continue;
      const int x = 37, y = 11;
      cudaMemset(dout, 0, sizeof(got));
      if (mode == 0) Probe<0><<<1, 32>>>(dm, hm, x, y, dout, fence);
      else Probe<1><<<1, 32>>>(dm, hm, x, y, dout, fence);
      cudaError_t e = cudaDeviceSynchronize();
      printf("mode %d (%s) fence %d: %s", mode, mode ? "grid_constant" : "global", fence, cudaGetErrorString(e));
      if (e == cudaSuccess) {
        cudaMemcpy(got, dout, sizeof(got), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < 21; ++r2)
          for (int c = 0; c < 32; ++c) bad += got[r2 * 32 + c] != h[(y + r2) * W + x + c];
        printf(", %d wrong bytes", bad);
      }
      printf("\n");
      if (e != cudaSuccess) return 1;
    }
  return 0;
}
