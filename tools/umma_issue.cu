// Micro-benchmark (development aid): cost per tcgen05.mma of different ways to feed the issuing thread with
// per-op descriptor words (N = 32, M = 128, K = 16).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
#include <cstdio>
#include <cuda_runtime.h>
#include "../mvsnet_b200/csrc/umma.cuh"
using namespace mvsb200::umma;

constexpr int NOPS = 108, N = 32;
struct Tab { uint4 ops[NOPS]; };

template <int VARIANT>
__global__ void __launch_bounds__(192, 1) issue_kernel(const __grid_constant__ Tab tab, int steps, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  __shared__ uint4 s_ops[NOPS];
  for (int i = threadIdx.x; i < 150 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < NOPS; i += blockDim.x) s_ops[i] = tab.ops[i];
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = make_idesc_bf16_f32(128, N);
  const uint32_t a16 = smem_u32(smem) >> 4;
  const uint64_t hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;
  const int nissuers = VARIANT == 5 ? 2 : 1;
  if (warp >= 4 && warp < 4 + nissuers) {
    const int w = warp - 4;
    if (elect_one()) {
      const long long t0 = clock64();
      for (int t = 0; t < steps; ++t) {
        const uint32_t sl = a16 + (uint32_t)(t & 3) * 64u;
        const uint32_t d_base = tmem + (uint32_t)((t & 1) * 64);
        if (VARIANT == 1) {
#pragma unroll 6
          for (int o = 0; o < NOPS; ++o) {
            const uint4 e = s_ops[o];
            mma_bf16(d_base + e.z, hi | (uint64_t)(e.x + sl), hi | (uint64_t)e.y, idesc, e.w);
          }
        } else if (VARIANT == 2) {
          // software pipelined: records of the next group are loaded before the current group is issued
          uint4 cur[6], nxt[6];
#pragma unroll
          for (int u = 0; u < 6; ++u) cur[u] = s_ops[u];
          for (int o = 0; o < NOPS; o += 6) {
            const int on = o + 6 < NOPS ? o + 6 : 0;
#pragma unroll
            for (int u = 0; u < 6; ++u) nxt[u] = s_ops[on + u];
#pragma unroll
            for (int u = 0; u < 6; ++u)
              mma_bf16(d_base + cur[u].z, hi | (uint64_t)(cur[u].x + sl), hi | (uint64_t)cur[u].y, idesc, cur[u].w);
#pragma unroll
            for (int u = 0; u < 6; ++u) cur[u] = nxt[u];
          }
        } else if (VARIANT == 3) {
          // constant bank (kernel parameter) table, uniform index
#pragma unroll 6
          for (int o = 0; o < NOPS; ++o) {
            const uint4 e = tab.ops[o];
            mma_bf16(d_base + e.z, hi | (uint64_t)(e.x + sl), hi | (uint64_t)e.y, idesc, e.w);
          }
        } else if (VARIANT == 4) {
          // two row blocks per op, inner loop unrolled
#pragma unroll 3
          for (int o = 0; o < NOPS / 2; ++o) {
            const uint4 e = s_ops[o];
            const uint32_t a_lo = e.x + sl, d = d_base + e.z;
            const uint64_t db = hi | (uint64_t)e.y;
            mma_bf16(d, hi | (uint64_t)a_lo, db, idesc, e.w);
            mma_bf16(d + 32, hi | (uint64_t)(a_lo + 128u), db, idesc, e.w);
          }
        } else if (VARIANT == 5) {
          // two issuing warps, each its own accumulator block and half of the MMAs
#pragma unroll 6
          for (int o = 0; o < NOPS / 2; ++o) {
            const uint4 e = s_ops[o];
            mma_bf16(d_base + e.z + w * 32, hi | (uint64_t)(e.x + sl + w * 128u), hi | (uint64_t)e.y, idesc, e.w);
          }
        } else if (VARIANT == 6) {
          // 64-bit descriptor halves straight from the table (A absolute, slot folded in via 4 table copies)
#pragma unroll 6
          for (int o = 0; o < NOPS; ++o) {
            const uint4 e = s_ops[o];
            mma_bf16(d_base + e.z, hi | (uint64_t)e.x, hi | (uint64_t)e.y, idesc, 1u);
          }
        }
      }
      const long long t1 = clock64();
      mma_commit(&s_bar[w]);
      mbar_wait(&s_bar[w], 0);
      const long long t2 = clock64();
      if (blockIdx.x == 0 && w == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int V>
void run(const Tab& tab, long long* d_out, const char* what) {
  const int steps = 100;
  cudaFuncSetAttribute(issue_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) issue_kernel<V><<<148, 192, 160 * 1024>>>(tab, steps, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("variant %d (%s): issue %.1f clk/MMA, complete %.1f clk/MMA (%s)\n", V, what, (double)h[0] / (steps * NOPS),
         (double)h[1] / (steps * NOPS), cudaGetErrorString(e));
}

int main() {
  Tab tab;
  for (int o = 0; o < NOPS; ++o) {
    const int tap = o % 9;
    tab.ops[o] = make_uint4((uint32_t)((tap / 3) * 14 + tap % 3) | ((2720u >> 4) << 16),
                            (uint32_t)((131072 + (o % 18) * 1024) >> 4) | ((512u >> 4) << 16), 0u, 1u);
  }
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<1>(tab, d_out, "smem table, unroll 6");
  run<2>(tab, d_out, "smem table, software pipelined");
  run<3>(tab, d_out, "constant-bank table");
  run<4>(tab, d_out, "smem table, 2 row blocks per op");
  run<5>(tab, d_out, "two issuing warps");
  run<6>(tab, d_out, "smem table, no adds");
  return 0;
}
