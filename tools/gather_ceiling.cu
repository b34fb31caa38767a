// Micro-benchmark: the best random 256-byte-row gather rate this GPU sustains from an L2-resident
// (18 MB) or HBM-resident (1.5 GB) fp32 table — the practical ceiling of the SpMM's neighbour gather.
// Every half-warp reads one random row per load (16 lanes x 16 B), UNR independent loads in flight.
// Build+run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/gather_ceiling.cu -o /tmp/gc && /tmp/gc
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

template <int UNR>
__global__ void __launch_bounds__(256) gather(const float4* __restrict__ tab, const int* __restrict__ idx,
                                              long n_idx, int row_f4, float4* __restrict__ out) {
  const int lane16 = threadIdx.x & 15;
  const long group = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const long n_groups = ((long)gridDim.x * blockDim.x) >> 4;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long e = group * UNR; e + UNR <= n_idx; e += n_groups * UNR) {
    int c[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) c[u] = __ldg(idx + e + u);
    float4 v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) v[u] = __ldg(tab + (long)c[u] * row_f4 + lane16);
#pragma unroll
    for (int u = 0; u < UNR; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x == 123.456f) out[threadIdx.x] = acc;
}

template <int UNR>
static void run(const float4* tab, const int* idx, long n_idx, float4* out, int blocks_per_sm, const char* what) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * blocks_per_sm;
  gather<UNR><<<grid, 256>>>(tab, idx, n_idx, 16, out);
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) gather<UNR><<<grid, 256>>>(tab, idx, n_idx, 16, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  printf("%s unroll %d, %d CTAs/SM: %.1f us for %ld rows -> %.2f TB/s (%.1f G rows/s)\n", what, UNR, blocks_per_sm,
         ms * 1e3, n_idx, n_idx * 256.0 / ms / 1e9, n_idx / ms / 1e6);
}

int main() {
  for (int big = 0; big < 2; ++big) {
    const long rows = big ? 6000000 : 70000;     // 1.5 GB vs 18 MB
    const long n_idx = big ? 40000000 : 2000000 * 8;
    float4* tab; int* idx; float4* out;
    cudaMalloc(&tab, rows * 256); cudaMemset(tab, 0, rows * 256);
    cudaMalloc(&idx, n_idx * 4); cudaMalloc(&out, 4096 * 16);
    std::vector<int> h(n_idx);
    unsigned long long s = 88172645463325252ull;
    for (long i = 0; i < n_idx; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % rows); }
    cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
    const char* what = big ? "HBM table (1.5 GB)" : "L2 table (18 MB) ";
    run<2>(tab, idx, n_idx, out, 8, what);
    run<4>(tab, idx, n_idx, out, 8, what);
    run<8>(tab, idx, n_idx, out, 8, what);
    run<8>(tab, idx, n_idx, out, 4, what);
    run<16>(tab, idx, n_idx, out, 4, what);
    cudaFree(tab); cudaFree(idx); cudaFree(out);
  }
  return 0;
}
