// Throughput of the integer / SIMD-video instructions the reconstruction kernels lean on, per SM and clock.
// Eight independent dependency chains per thread, 8 warps per SM sub-partition: issue-bound, not latency-bound.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

template <int OP>
__device__ __forceinline__ uint32_t Step(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  if (OP == 0) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == 1) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == 2) asm volatile("vabsdiff4.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0));
  else if (OP == 3) asm volatile("{.reg .b32 t; add.s16x2 t, %1, %2; min.s16x2.relu %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == 4) asm volatile("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else if (OP == 5) asm volatile("{.reg .b32 t; max.u16x2 t, %1, %2; max.u16x2 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == 6) asm volatile("add.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else if (OP == 7) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == 8) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else if (OP == 9) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == 10) asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == 11) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (OP == 12) asm volatile("min.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else if (OP == 13) asm volatile("shl.b32 %0, %1, 8;" : "=r"(d) : "r"(a));
  else if (OP == 14) asm volatile("add.u32 %0, %1, %2;\n\tadd.u32 %0, %0, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));  // iadd3 candidate
  else d = a;
  return d;
}

template <int OP, int OP2>
__global__ void __launch_bounds__(256) Bench(uint32_t *out, uint32_t seed, long long *cycles) {
  uint32_t x[CHAINS], y = seed ^ threadIdx.x, z = seed * 3 + blockIdx.x;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) x[i] = seed + i * 977 + threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      x[i] = Step<OP>(x[i], y, z);
      if (OP2 >= 0) x[i] = Step<OP2>(x[i], z, y);
    }
  }
  const long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) r ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP, int OP2>
void Run(const char *name, uint32_t *d_out, long long *d_cyc, int sms) {
  const int ctas = sms * 4;  // 4 CTAs x 8 warps = 32 warps per SM, 8 per sub-partition
  Bench<OP, OP2><<<ctas, 256>>>(d_out, 12345u, d_cyc);
  Bench<OP, OP2><<<ctas, 256>>>(d_out, 12345u, d_cyc);
  cudaDeviceSynchronize();
  static long long h[4096];
  cudaMemcpy(h, d_cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < ctas; ++i) avg += h[i];
  avg /= ctas;
  const double warp_instr_per_sm = 32.0 * ITERS * CHAINS * (OP2 >= 0 ? 2 : 1);
  printf("%-34s %6.3f warp-instructions / clock / SM   (%.2f clocks per warp-instruction per sub-partition)\n", name,
         warp_instr_per_sm / avg, avg / (warp_instr_per_sm / 4));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  uint32_t *d_out;
  long long *d_cyc;
  cudaMalloc(&d_out, sizeof(uint32_t) * p.multiProcessorCount * 4 * 256);
  cudaMalloc(&d_cyc, sizeof(long long) * 4096);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  Run<0, -1>("LOP3", d_out, d_cyc, p.multiProcessorCount);
  Run<1, -1>("PRMT", d_out, d_cyc, p.multiProcessorCount);
  Run<2, -1>("VABSDIFF4.U8", d_out, d_cyc, p.multiProcessorCount);
  Run<3, -1>("VIADDMNMX.S16x2.RELU", d_out, d_cyc, p.multiProcessorCount);
  Run<4, -1>("VIMNMX.U16x2", d_out, d_cyc, p.multiProcessorCount);
  Run<5, -1>("VIMNMX3.U16x2", d_out, d_cyc, p.multiProcessorCount);
  Run<6, -1>("VIADD.16x2", d_out, d_cyc, p.multiProcessorCount);
  Run<7, -1>("IMAD", d_out, d_cyc, p.multiProcessorCount);
  Run<8, -1>("IADD (compiler's choice of pipe)", d_out, d_cyc, p.multiProcessorCount);
  Run<9, -1>("SHF", d_out, d_cyc, p.multiProcessorCount);
  Run<10, -1>("IDP.4A", d_out, d_cyc, p.multiProcessorCount);
  Run<11, -1>("I2IP (cvt.pack.sat)", d_out, d_cyc, p.multiProcessorCount);
  Run<12, -1>("VIMNMX (32-bit min)", d_out, d_cyc, p.multiProcessorCount);
  Run<13, -1>("SHL by 8", d_out, d_cyc, p.multiProcessorCount);
  Run<0, 7>("LOP3 + IMAD interleaved", d_out, d_cyc, p.multiProcessorCount);
  Run<1, 7>("PRMT + IMAD interleaved", d_out, d_cyc, p.multiProcessorCount);
  Run<3, 7>("VIADDMNMX + IMAD interleaved", d_out, d_cyc, p.multiProcessorCount);
  Run<2, 7>("VABSDIFF4 + IMAD interleaved", d_out, d_cyc, p.multiProcessorCount);
  Run<1, 0>("PRMT + LOP3 interleaved", d_out, d_cyc, p.multiProcessorCount);
  Run<3, 1>("VIADDMNMX + PRMT interleaved", d_out, d_cyc, p.multiProcessorCount);
  Run<10, 0>("IDP.4A + LOP3 interleaved", d_out, d_cyc, p.multiProcessorCount);
  Run<10, 7>("IDP.4A + IMAD interleaved", d_out, d_cyc, p.multiProcessorCount);
  return 0;
}
