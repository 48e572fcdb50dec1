// Micro-benchmark (development aid, not on the product path): issue / execution rate of
// tcgen05.mma M=128, K=16, bf16, no-swizzle K-major operands, as a function of N, and the
// tcgen05.ld drain rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../mvsnet_b200/csrc/umma.cuh"

using namespace mvsb200::umma;

// mode 0: A start cycles through 9 tap offsets of a padded tile (PX cells per row); mode 1: same A every time
template <int N, int UNROLL>
__global__ void __launch_bounds__(160, 1) rate_kernel(int iters, int mode, int px, int ps, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 4) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16_f32(128, N);
      const uint32_t a16 = smem_u32(smem) >> 4;
      const uint32_t b16 = (smem_u32(smem) + 128 * 1024) >> 4;
      const uint64_t hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;
      const uint32_t a_lbo = ((uint32_t)ps >> 4) << 16, b_lbo = ((uint32_t)(N * 16) >> 4) << 16;
      uint32_t aoff[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int tap = u % 9;
        aoff[u] = mode == 0 ? (uint32_t)((tap / 3) * px + tap % 3) : 0u;
      }
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          mma_bf16(tmem + (u & 1) * N, hi | (uint64_t)(a16 + aoff[u] + a_lbo), hi | (uint64_t)(b16 + b_lbo), idesc, 1u);
        }
      }
      const long long t1 = clock64();
      mma_commit(&s_bar);
      mbar_wait(&s_bar, 0);
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// tcgen05.ld drain: 4 warps each read 32 lanes x COLS columns, `iters` times
template <int COLS>
__global__ void __launch_bounds__(128, 1) ld_kernel(int iters, long long* out, float* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem + ((uint32_t)(warp * 32) << 16);
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[COLS];
#pragma unroll
    for (int c = 0; c < COLS; c += 16) tmem_ld16(tmem + ((it & 3) * COLS) + c, r + c);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < COLS; ++c) acc += __uint_as_float(r[c]);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_tmem, 512);
}

template <int N>
void run(int grid, int mode, int px, int ps, long long* d_out) {
  const int iters = 200, UN = 18;
  cudaFuncSetAttribute(rate_kernel<N, UN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) rate_kernel<N, UN><<<grid, 160, 200 * 1024>>>(iters, mode, px, ps, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d grid=%3d mode=%d px=%d ps=%d: issue %.1f clk/MMA, complete %.1f clk/MMA  (%s)\n", N, grid, mode, px, ps,
         (double)h[0] / (iters * UN), (double)h[1] / (iters * UN), cudaGetErrorString(e));
}

template <int COLS>
void run_ld(long long* d_out, float* sink) {
  const int iters = 1000;
  for (int rep = 0; rep < 2; ++rep) ld_kernel<COLS><<<148, 128>>>(iters, d_out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[1];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("tcgen05.ld 4 warps x 32 lanes x %d cols: %.1f clk per round = %.1f B/clk/SM (%s)\n", COLS, (double)h[0] / iters,
         128.0 * COLS * 4 * iters / (double)h[0], cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  float* sink;
  cudaMalloc(&d_out, 64);
  cudaMalloc(&sink, 4096);
  for (int grid : {1, 148}) {
    for (int mode : {0, 1}) {
      run<16>(grid, mode, 14, 14 * 11 * 16 + 256, d_out);
      run<32>(grid, mode, 14, 14 * 11 * 16 + 256, d_out);
      run<64>(grid, mode, 14, 14 * 11 * 16 + 256, d_out);
      run<96>(grid, mode, 14, 14 * 11 * 16 + 256, d_out);
      run<128>(grid, mode, 14, 14 * 11 * 16 + 256, d_out);
      run<256>(grid, mode, 14, 14 * 11 * 16 + 256, d_out);
    }
  }
  // plane stride variants (bank alignment of the two K halves)
  run<32>(148, 0, 14, 4096, d_out);
  run<32>(148, 0, 14, 4096 + 32, d_out);
  run<32>(148, 0, 14, 4096 + 64, d_out);
  run<32>(148, 0, 26, 26 * 21 * 16, d_out);
  run_ld<16>(d_out, sink);
  run_ld<32>(d_out, sink);
  run_ld<64>(d_out, sink);
  return 0;
}
