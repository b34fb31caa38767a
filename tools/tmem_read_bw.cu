// Micro-benchmark: how fast can an SM read its tensor memory back into registers?
// The fused score+top-k kernel has to pull every fp32 accumulator through tcgen05.ld, so this rate
// bounds it.  One CTA per SM allocates all 512 TMEM columns; W warps (warp w reads
// lane quarter w % 4) loop over tcgen05.ld.32x32b.x32 (32 lanes x 32 columns x 4 B = 4 KB per
// instruction), with D loads in flight before each tcgen05.wait::ld.  No MMA is running.
// Build+run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/tmem_read_bw.cu -o /tmp/tm && /tmp/tm
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

template <int DEPTH>
__global__ void __launch_bounds__(512, 1) tmem_read(int iters, uint32_t* sink, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[DEPTH][32];
#pragma unroll
    for (int dd = 0; dd < DEPTH; ++dd) tc_ld32(base + (uint32_t)(((it * DEPTH + dd) * 32) & 511), r[dd]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int dd = 0; dd < DEPTH; ++dd) acc ^= r[dd][0] ^ r[dd][31];   // 2 ALU ops per 32 cells: the reads are the cost
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int DEPTH>
static void run(int warps, uint32_t* sink, long long* cyc) {
  const int iters = 4096;
  tmem_read<DEPTH><<<148, warps * 32>>>(iters, sink, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  tmem_read<DEPTH><<<148, warps * 32>>>(iters, sink, cyc);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double bytes_sm = (double)warps * iters * DEPTH * 4096.0;
  cudaError_t e = cudaGetLastError();
  printf("%2d warps, %d loads in flight: %8.1f B/clk/SM (%lld clk)  %6.2f TB/s chip  [%s]\n", warps, DEPTH,
         bytes_sm / (double)c, c, bytes_sm * 148 / (ms * 1e-3) / 1e12, cudaGetErrorString(e));
}

int main() {
  uint32_t* sink; long long* cyc;
  cudaMalloc(&sink, 4096); cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 16}) {
    run<1>(w, sink, cyc);
    run<2>(w, sink, cyc);
    if (w < 16) run<3>(w, sink, cyc);   // 16 warps x 3 x 32 registers would not fit the register file
  }
  return 0;
}
