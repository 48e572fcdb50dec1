// Diagnostic: one 128 x N x (16*kblocks) tcgen05 GEMM tile on caller-provided shared-memory images.
// It pins the descriptor semantics the implicit-GEMM convolution relies on (no-swizzle K-major
// operands with arbitrary 16-byte-aligned start, leading and stride byte offsets) against a host
// reference.  Not on the product path.
#include "common.cuh"
#include "umma.cuh"

namespace mvsb200 {
using namespace umma;

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint4* __restrict__ a_image, int a_bytes, const uint4* __restrict__ b_image, int b_bytes, int n,
                  int kblocks, int a_kblock_stride, int a_start, int a_lbo, int a_sbo, int b_kblock_stride, int b_lbo,
                  int b_sbo, float* __restrict__ d_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  unsigned char* sa = smem;
  unsigned char* sb = smem + ((a_bytes + 127) / 128) * 128;
  for (int i = threadIdx.x; i < a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sa)[i] = a_image[i];
  for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sb)[i] = b_image[i];
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&s_tmem, 256);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16_f32(128, n);
    for (int j = 0; j < kblocks; ++j) {
      uint64_t da = make_smem_desc(smem_u32(sa) + a_start + j * a_kblock_stride, a_lbo, a_sbo);
      uint64_t db = make_smem_desc(smem_u32(sb) + j * b_kblock_stride, b_lbo, b_sbo);
      mma_bf16(tmem, da, db, idesc, j > 0 ? 1u : 0u);
    }
    mma_commit(&s_bar);
  }
  mbar_wait(&s_bar, 0);
  tc_fence_after();
  const int row = threadIdx.x;   // TMEM lane == D row for M = 128, cta_group::1
  for (int c0 = 0; c0 < n; c0 += 8) {
    uint32_t r[8];
    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 8; ++k) d_out[row * n + c0 + k] = __uint_as_float(r[k]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_umma_probe(const void* a_image, int a_bytes, const void* b_image, int b_bytes, int n,
                                  int kblocks, int a_kblock_stride, int a_start, int a_lbo, int a_sbo,
                                  int b_kblock_stride, int b_lbo, int b_sbo, float* d_out, void* stream) {
  MVS_CHECK_ARG(a_image && b_image && d_out, "umma_probe: NULL pointer");
  MVS_CHECK_ARG(a_bytes > 0 && b_bytes > 0 && a_bytes % 16 == 0 && b_bytes % 16 == 0, "umma_probe: bad image sizes");
  MVS_CHECK_ARG(n >= 16 && n <= 256 && n % 16 == 0 && kblocks >= 1, "umma_probe: bad n/kblocks");
  size_t smem = (size_t)((a_bytes + 127) / 128) * 128 + b_bytes;
  MVS_CHECK_ARG(smem <= 200 * 1024, "umma_probe: images too large");
  MVS_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint4*)a_image, a_bytes, (const uint4*)b_image,
                                                            b_bytes, n, kblocks, a_kblock_stride, a_start, a_lbo,
                                                            a_sbo, b_kblock_stride, b_lbo, b_sbo, d_out);
  MVS_LAUNCH_CHECK("umma_probe_kernel");
  return MVSB200_OK;
}
