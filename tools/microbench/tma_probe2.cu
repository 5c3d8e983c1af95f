// usage: tma_probe2 <elem_bytes 1|2|4> <box_w> <box_h> <x> <y> <l2promo 0..3> <clearbit 0|1>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
struct alignas(128) Map { unsigned char b[128]; };
__device__ __forceinline__ unsigned Smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void Probe(const __grid_constant__ Map pmap, int x, int y, unsigned bytes, unsigned char *out) {
  extern __shared__ __align__(1024) unsigned char tile[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(Smem(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(Smem(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(Smem(tile)),
                 "l"(&pmap), "r"(x), "r"(y), "r"(Smem(&bar))
                 : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(Smem(&bar))
      : "memory");
  for (unsigned i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char **argv) {
  const int eb = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]), promo = atoi(argv[6]), clr = atoi(argv[7]);
  const int W = 1024, H = 512;  // elements
  const size_t bytes_total = size_t(W) * H * eb;
  unsigned char *h = new unsigned char[bytes_total];
  for (size_t i = 0; i < bytes_total; ++i) h[i] = (unsigned char)((i * 7 + (i >> 9) * 13) & 0xff);
  unsigned char *d, *dout;
  cudaMalloc(&d, bytes_total);
  const unsigned box_bytes = unsigned(bw) * bh * eb;
  cudaMalloc(&dout, box_bytes);
  cudaMemcpy(d, h, bytes_total, cudaMemcpyHostToDevice);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
  Map hm;
  cuuint64_t dims[2] = {cuuint64_t(W), cuuint64_t(H)}, strides[1] = {cuuint64_t(W) * eb};
  cuuint32_t box[2] = {cuuint32_t(bw), cuuint32_t(bh)}, es[2] = {1, 1};
  const CUtensorMapDataType dt = eb == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32);
  CUresult r = enc(reinterpret_cast<CUtensorMap *>(&hm), dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (clr) reinterpret_cast<uint64_t *>(&hm)[1] &= ~(1ull << 21);
  printf("eb %d box %dx%d at (%d,%d) promo %d clr %d: encode %d; ", eb, bw, bh, x, y, promo, clr, (int)r);
  const uint64_t *wds = reinterpret_cast<const uint64_t *>(&hm);
  Probe<<<1, 128, box_bytes + 1024>>>(hm, x, y, box_bytes, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    unsigned char *got = new unsigned char[box_bytes];
    cudaMemcpy(got, dout, box_bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < bh; ++r2)
      for (int c = 0; c < bw * eb; ++c) bad += got[r2 * bw * eb + c] != h[(size_t(y + r2) * W + x) * eb + c];
    printf(", %d wrong bytes", bad);
  }
  printf("  desc[0..3] = %016llx %016llx %016llx %016llx\n", (unsigned long long)wds[0], (unsigned long long)wds[1], (unsigned long long)wds[2],
         (unsigned long long)wds[3]);
  return e != cudaSuccess;
}
