// Test/diagnostic only: raw tcgen05 probe.  The host supplies a verbatim shared-memory image, a list of
// (A descriptor, B descriptor, TMEM column, accumulate flag) MMAs and the instruction descriptor; one CTA
// copies the image to 1024-byte-aligned shared memory, issues the MMAs from one elected thread and
// returns the TMEM accumulator columns plus the cycles the list took.  Shared-memory matrix-descriptor
// conventions (swizzle on absolute address bits, Toeplitz views with overlapping rows, no-swizzle
// layouts with SBO = 128) are therefore tested on the host side of the test, against numpy.
// Not called by the product path.
#include "common.cuh"

#define PROBE_MAX_MMA 256

__global__ void __launch_bounds__(128, 1)
umma_raw_kernel(const uint8_t* __restrict__ image, int image_bytes, const uint64_t* __restrict__ a_desc,
                const uint64_t* __restrict__ b_desc, const uint32_t* __restrict__ d_col,
                const uint32_t* __restrict__ acc, const uint32_t* __restrict__ idescs, int n_mma, int n_cols_out,
                float* __restrict__ d_out, int reps, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t s_a[PROBE_MAX_MMA], s_b[PROBE_MAX_MMA];
  __shared__ uint32_t s_col[PROBE_MAX_MMA], s_acc[PROBE_MAX_MMA], s_id[PROBE_MAX_MMA];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < image_bytes / 16; i += 128)
    reinterpret_cast<uint4*>(sm)[i] = reinterpret_cast<const uint4*>(image)[i];
  for (int i = threadIdx.x; i < n_mma; i += 128) {
    // the host's start-address field is an offset into the image; rebase it (14-bit field, units of 16 B)
    const uint64_t lo_mask = 0x3FFFull;
    uint64_t a = a_desc[i], b = b_desc[i];
    a = (a & ~lo_mask) | (((a & lo_mask) + (base >> 4)) & lo_mask);
    b = (b & ~lo_mask) | (((b & lo_mask) + (base >> 4)) & lo_mask);
    s_a[i] = a; s_b[i] = b; s_col[i] = d_col[i]; s_acc[i] = acc[i]; s_id[i] = idescs[i];
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    uint32_t phase = 0;
    long long total = 0;
    for (int rep = 0; rep < reps; ++rep) {
      const long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < n_mma; ++i) umma_f16(tmem + s_col[i], s_a[i], s_b[i], s_id[i], (rep > 0) ? 1u : s_acc[i]);
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      total += clock64() - t0;
    }
    if (threadIdx.x == 0 && cycles) cycles[0] = total;
  }
  tc_fence_after();
  __syncthreads();
  if (reps == 1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int nb = 0; nb < n_cols_out / 32; ++nb) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + nb * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) d_out[(size_t)(warp * 32 + lane) * n_cols_out + nb * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

extern "C" int zns_dbg_umma_raw(const void* image, int image_bytes, const uint64_t* a_desc, const uint64_t* b_desc,
                                const uint32_t* d_col, const uint32_t* acc, const uint32_t* idescs, int n_mma,
                                int n_cols_out, float* d_out, int reps, long long* cycles, void* stream) {
  ZNS_REQUIRE(image && a_desc && b_desc && d_col && acc && idescs && d_out, "NULL argument");
  ZNS_REQUIRE(image_bytes > 0 && image_bytes % 16 == 0 && image_bytes <= 200 * 1024, "image must be 16..204800 bytes");
  ZNS_REQUIRE(n_mma >= 1 && n_mma <= PROBE_MAX_MMA, "1..%d MMAs", PROBE_MAX_MMA);
  ZNS_REQUIRE(n_cols_out % 32 == 0 && n_cols_out >= 32 && n_cols_out <= 512 && reps >= 1, "bad output columns / reps");
  ZNS_CHECK_CUDA(cudaFuncSetAttribute(umma_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 206 * 1024));
  umma_raw_kernel<<<1, 128, (size_t)image_bytes + 1024, (cudaStream_t)stream>>>(
      (const uint8_t*)image, image_bytes, a_desc, b_desc, d_col, acc, idescs, n_mma, n_cols_out, d_out, reps, cycles);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}
