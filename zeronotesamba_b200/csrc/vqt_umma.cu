// VQT / CQT pyramid on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// Replaces the per-octave work of librosa.vqt / librosa.cqt (0.8.1) + resampy 0.4.2 as the reference calls
// them at /root/reference/zeroNoteSamba/processing/input_rep.py:27-34,42-49: for every octave i
//   (a) the framed filterbank  C_i[k,t] = sum_n g_i[k,n] ypad_i[t hop_i + n - n_fft_i/2]   (+ |.|, 1/sqrt(L), log)
//   (b) the 2:1 "kaiser_fast" decimation  y_{i+1}[t] = sqrt(2) sum_{|j|<=31} h[|j|] y_i[2t + j]
// Both are contractions of a Toeplitz view of the level signal with a small constant matrix, so ONE level
// kernel runs them as tcgen05.mma on the same shared-memory image of the signal:
//
//  * fp32 samples are split into two fp16 terms x = x1 + x2/2048 (22 mantissa bits), coefficients likewise;
//    the three leading products accumulate in fp32 in TMEM (x1 g1 | x1 g2 + x2 g1), as the mma.sync
//    filterbank of round 1 did.
//  * the level signal lives in shared memory "chunk-major": 16-byte chunk c (8 samples) of row r (R = 8 q
//    samples) at  c * LBO + 16 r.  With the no-swizzle K-major descriptor and SBO = 128 this gives MMA rows at a
//    16-byte pitch, so a shift of the operand window by whole rows / chunks is just a different descriptor
//    start address: every filter tap window and every decimator window is a VIEW of the same image -- nothing is
//    materialised per frame (profiles/r02_umma_view_probe.txt validates the descriptor semantics).
//  * filterbank: row = hop block (or 2 / 4 frames per row, their shifted filters stacked on N);
//    decimator: banded Toeplitz, per 16-sample k-step one 48-output window of a shared tap tile.
//  * epilogue: TMEM -> registers; filterbank rows -> log-magnitude -> out[b][bin][frame]; decimator rows ->
//    the next level's signal, already split into its two fp16 terms.
//  * frames whose window crosses a clip edge need reflect padding while the decimator needs zero extension:
//    the level kernels zero-extend and a tiny SIMT kernel recomputes those (<= 4 per edge and octave) frames.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <cuda_fp16.h>

#include "common.cuh"
#include "vqt_plan.h"

// ---------------------------------------------------------------------------------------------
// host: level plans
// ---------------------------------------------------------------------------------------------
static inline uint16_t hbits(__half h) { return __half_as_ushort(h); }

static void split_half(float v, uint16_t* h1, uint16_t* h2) {
  const __half a = __float2half_rn(v);
  const __half b = __float2half_rn((v - __half2float(a)) * 2048.f);
  *h1 = hbits(a);
  *h2 = hbits(b);
}

// power-of-two scale that brings the largest coefficient magnitude into [0.5, 1)
float vqt_coef_scale(const float* re, const float* im, int n) {
  float gmax = 0.f;
  for (int k = 0; k < n; ++k) gmax = std::max(gmax, std::max(fabsf(re[k]), fabsf(im[k])));
  return gmax > 0.f ? exp2f(-1.f - floorf(log2f(gmax))) : 1.f;
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// element (row n, k) of a chunk-major coefficient tile with `rows` rows and 16 k: halfword index
static inline size_t tile_idx(size_t tile_off_bytes, int rows, int n, int k) {
  return (tile_off_bytes + (size_t)(k / 8) * 16 * rows + (size_t)16 * n) / 2 + (k % 8);
}

#define DEC_T_ROWS 112   // decimator tap tile rows n = -32 .. 79
#define DEC_T_N0 (-32)
#define DEC_TAP_SCALE 2.0f

static int build_level(zns_vqt_plan* p, int lvl, double fmin, double gamma_in, const double* taps) {
  VqtLevelDev& L = p->level[lvl];
  memset(&L, 0, sizeof(L));
  const int hop = p->hop >> lvl;
  const int nf = p->n_fft[lvl];
  if (hop < 2 || nf > 128 || nf < 16 || p->bpo != 12) return 1;
  const int R = hop >= 32 ? hop : (hop >= 8 ? 32 : 8);
  const int q = R / 8, fpr = R / hop;
  if (!(q == 1 || q == 4 || q == 8 || q == 16 || q == 32) || fpr > 4) return 1;
  const bool last = (lvl == p->n_oct - 1);
  L.q = q; L.fpr = fpr; L.hop = hop; L.n_fft = nf; L.bin0 = p->n_bins - p->bpo * (lvl + 1);
  L.dec_w = last ? 0 : R / 2;
  L.wacc = last ? 0 : ((L.dec_w + 15) / 16) * 16;
  L.dec_scale = (float)(sqrt(2.0) / DEC_TAP_SCALE);
  L.fb_scale = p->coef_inv_scale[lvl];

  // ---- k-steps (16 samples at offset X0 from the row start) ----
  std::vector<int> fb_x0, dec_x0;
  const int x_lo = floordiv(-nf / 2, 16) * 16, x_hi = (fpr - 1) * hop + nf / 2;
  for (int x = x_lo; x < x_hi; x += 16) fb_x0.push_back(x);
  if (!last) for (int x = -32; x < R + 32; x += 16) dec_x0.push_back(x);
  int smin = 0, smax = 0;
  auto track = [&](int x0) {
    if (q == 1) { smin = std::min(smin, floordiv(x0, 8)); smax = std::max(smax, floordiv(x0, 8) + 1); }
    else { smin = std::min(smin, floordiv(x0, R)); smax = std::max(smax, floordiv(x0 + 15, R)); }
  };
  for (int x : fb_x0) track(x);
  for (int x : dec_x0) track(x);
  L.hb = -smin; L.ha = smax;
  int rtot = 128 + L.hb + L.ha;
  if (q >= 8) { if (rtot % 2 == 0) ++rtot; }
  else if (q == 4) { while (rtot % 8 != 2) ++rtot; }
  L.rtot = rtot;
  L.a_lbo = (q == 1) ? 16 : 16 * rtot;
  auto a_off = [&](int x0) -> uint32_t {
    if (q == 1) return (uint32_t)(16 * (floordiv(x0, 8) + L.hb));
    const int rs = floordiv(x0, R), p0 = (x0 - rs * R) / 8;
    return (uint32_t)(p0 * L.a_lbo + 16 * (rs + L.hb));
  };

  // ---- TMEM columns ----
  const int n1 = fpr * 48, n2 = std::max(32, fpr * 24);
  L.fb_b_stride = 24;
  L.dec_a_col = 0; L.dec_b_col = L.wacc; L.fb_a_col = 2 * L.wacc; L.fb_b_col = L.fb_a_col + n1;
  const int used = L.fb_b_col + n2;
  int alloc = 32;
  while (alloc < used) alloc *= 2;
  if (alloc > 512) return 1;
  L.tmem_cols = alloc;

  // ---- coefficient image ----
  std::vector<float> re((size_t)p->bpo * 1024), im((size_t)p->bpo * 1024);
  int nf2 = 0;
  int rc = zns_vqt_basis_host(p->sr, p->n_bins, p->bpo, fmin, gamma_in, lvl, re.data(), im.data(), &nf2);
  if (rc || nf2 != nf) return 1;
  const float gs = 1.f / p->coef_inv_scale[lvl];
  const size_t fb_tile1 = (size_t)n1 * 32, fb_tile2 = (size_t)n2 * 32;
  const size_t dec_base = fb_x0.size() * (fb_tile1 + fb_tile2);
  const size_t total = dec_base + (last ? 0 : 2 * (size_t)DEC_T_ROWS * 32);
  std::vector<uint16_t>* img = new std::vector<uint16_t>(total / 2, 0);
  for (size_t s = 0; s < fb_x0.size(); ++s) {
    const size_t t1 = s * (fb_tile1 + fb_tile2), t2 = t1 + fb_tile1;
    for (int j = 0; j < fpr; ++j)
      for (int col = 0; col < 24; ++col)
        for (int k = 0; k < 16; ++k) {
          const int n = fb_x0[s] + k - (j * hop - nf / 2);
          if (n < 0 || n >= nf) continue;
          const int bin = col / 2;
          const float v = ((col & 1) ? im[(size_t)bin * nf + n] : re[(size_t)bin * nf + n]) * gs;
          uint16_t h1, h2;
          split_half(v, &h1, &h2);
          (*img)[tile_idx(t1, n1, j * 48 + col, k)] = h1;
          (*img)[tile_idx(t1, n1, j * 48 + 24 + col, k)] = h2;
          (*img)[tile_idx(t2, n2, j * 24 + col, k)] = h1;
        }
  }
  if (!last) {
    for (int n = DEC_T_N0; n < DEC_T_N0 + DEC_T_ROWS; ++n)
      for (int k = 0; k < 16; ++k) {
        const int j = abs(k + 32 - 2 * n);
        if (j > 31) continue;
        uint16_t h1, h2;
        split_half((float)(taps[j] * DEC_TAP_SCALE), &h1, &h2);
        (*img)[tile_idx(dec_base, DEC_T_ROWS, n - DEC_T_N0, k)] = h1;
        (*img)[tile_idx(dec_base + (size_t)DEC_T_ROWS * 32, DEC_T_ROWS, n - DEC_T_N0, k)] = h2;
      }
  }
  L.b_bytes = (int)total;
  p->h_bimg[lvl] = img;

  // ---- MMA list ----
  int m = 0;
  auto push = [&](uint32_t ao, size_t bo, int n, int col, int term, int brows) -> bool {
    if (m >= ZNS_VQT_MAX_MMA) return false;
    L.mma[m++] = VqtMma{ao, (uint32_t)bo, (uint16_t)n, (uint16_t)col, (uint16_t)term, (uint16_t)brows};
    return true;
  };
  for (size_t s = 0; s < fb_x0.size(); ++s) {
    const size_t t1 = s * (fb_tile1 + fb_tile2), t2 = t1 + fb_tile1;
    if (!push(a_off(fb_x0[s]), t1, n1, L.fb_a_col, 0, n1)) return 1;
    if (!push(a_off(fb_x0[s]), t2, n2, L.fb_b_col, 1, n2)) return 1;
  }
  for (int x0 : dec_x0) {
    // outputs o = x0/2 - 16 + n; taps are non-zero for n in [1, 39]
    const int o_lo = std::max(0, x0 / 2 - 15), o_hi = std::min(L.dec_w - 1, x0 / 2 + 23);
    if (o_lo > o_hi) continue;
    int o_start = (o_lo / 8) * 8;
    int n = ((o_hi + 1 - o_start + 15) / 16) * 16;
    if (o_start + n > L.wacc) o_start = L.wacc - n;
    if (o_start < 0) { o_start = 0; n = L.wacc; }
    const int n0 = o_start - x0 / 2 + 16;
    if (n0 < DEC_T_N0 || n0 + n > DEC_T_N0 + DEC_T_ROWS) return 1;
    const size_t row_off = (size_t)16 * (n0 - DEC_T_N0);
    const size_t tt1 = dec_base + row_off, tt2 = dec_base + (size_t)DEC_T_ROWS * 32 + row_off;
    if (!push(a_off(x0), tt1, n, L.dec_a_col + o_start, 0, DEC_T_ROWS)) return 1;
    if (!push(a_off(x0), tt2, n, L.dec_b_col + o_start, 0, DEC_T_ROWS)) return 1;
    if (!push(a_off(x0), tt1, n, L.dec_b_col + o_start, 1, DEC_T_ROWS)) return 1;
  }
  L.n_mma = m;
  return 0;
}

void vqt_umma_free(zns_vqt_plan* p) {
  for (int i = 0; i < ZNS_VQT_MAX_OCT; ++i) {
    if (p->d_bimg[i]) cudaFree(p->d_bimg[i]);
    if (p->d_hi[i]) cudaFree(p->d_hi[i]);
    if (p->d_lo[i]) cudaFree(p->d_lo[i]);
    delete p->h_bimg[i];
    p->d_bimg[i] = nullptr; p->d_hi[i] = nullptr; p->d_lo[i] = nullptr; p->h_bimg[i] = nullptr;
  }
}

static size_t level_smem(const VqtLevelDev& L) { return (size_t)2 * L.q * L.rtot * 16 + (size_t)L.b_bytes + 16; }

int vqt_umma_build(zns_vqt_plan* p, double fmin, double gamma_in) {
  p->umma_ok = false;
  double taps[32];
  zns_vqt_decimator_taps_host(taps);
  for (int i = 0; i < p->n_oct; ++i) {
    if (build_level(p, i, fmin, gamma_in, taps) != 0 || level_smem(p->level[i]) > 220 * 1024) {
      vqt_umma_free(p);
      return ZNS_OK;   // unsupported geometry: the legacy kernels stay in charge
    }
  }
  size_t n = (size_t)p->max_samples;
  for (int i = 0; i < p->n_oct; ++i) {
    ZNS_CHECK_CUDA(cudaMalloc(&p->d_bimg[i], p->level[i].b_bytes));
    ZNS_CHECK_CUDA(cudaMemcpy(p->d_bimg[i], p->h_bimg[i]->data(), p->level[i].b_bytes, cudaMemcpyHostToDevice));
    if (i > 0) {
      n = (n + 1) / 2;
      p->sig_stride[i] = (long long)((n + 63) / 64) * 64;
      const size_t bytes = (size_t)p->max_batch * p->sig_stride[i] * sizeof(uint16_t);
      ZNS_CHECK_CUDA(cudaMalloc(&p->d_hi[i], bytes));
      ZNS_CHECK_CUDA(cudaMalloc(&p->d_lo[i], bytes));
      ZNS_CHECK_CUDA(cudaMemset(p->d_hi[i], 0, bytes));
      ZNS_CHECK_CUDA(cudaMemset(p->d_lo[i], 0, bytes));
    }
  }
  p->umma_ok = true;
  return ZNS_OK;
}

// ---------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// zero 32 lanes x 16 columns
__device__ __forceinline__ void tmem_zero_32x16(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// no-swizzle K-major descriptor, rows at 16-byte pitch (SBO = 128), leading-dimension offset lbo
__device__ __forceinline__ uint64_t desc_rows16(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}

// x = x1 + x2 / 2048 for eight values -> two 16-byte chunks
__device__ __forceinline__ void split8(const float* v, uint4& c1, uint4& c2) {
  uint32_t a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h1 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 f1 = __half22float2(h1);
    const __half2 h2 = __floats2half2_rn((v[2 * i] - f1.x) * 2048.f, (v[2 * i + 1] - f1.y) * 2048.f);
    a[i] = *reinterpret_cast<const uint32_t*>(&h1);
    b[i] = *reinterpret_cast<const uint32_t*>(&h2);
  }
  c1 = make_uint4(a[0], a[1], a[2], a[3]);
  c2 = make_uint4(b[0], b[1], b[2], b[3]);
}

#define VQT_LEVEL_THREADS 256

template <bool SRC_F32>
__global__ void __launch_bounds__(VQT_LEVEL_THREADS)
vqt_level_kernel(const __grid_constant__ VqtLevelDev L, const float* __restrict__ y32,
                 const uint16_t* __restrict__ src_hi, const uint16_t* __restrict__ src_lo, int n_sig,
                 long long src_stride, const uint16_t* __restrict__ bimg, const float* __restrict__ inv_sqrt_len,
                 float* __restrict__ out, int n_frames, int n_bins, uint16_t* __restrict__ dst_hi,
                 uint16_t* __restrict__ dst_lo, int n_valid, long long dst_stride) {
  extern __shared__ __align__(16) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int q = L.q, R = 8 * q;
  const uint32_t plane_bytes = (uint32_t)q * L.rtot * 16;
  uint8_t* sA1 = sm;
  uint8_t* sA2 = sm + plane_bytes;
  uint8_t* sB = sm + 2 * plane_bytes;
  const int clip = blockIdx.z;
  const int row0 = blockIdx.x * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), (uint32_t)L.tmem_cols); tmem_relinquish(); }

  // coefficient image
  for (int i = threadIdx.x; i < L.b_bytes / 16; i += VQT_LEVEL_THREADS)
    reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(bimg) + i);

  // signal rows row0 - hb .. row0 + 127 + ha, zero extended outside [0, n_sig)
  {
    const int n_rows = 128 + L.hb + L.ha;
    const int n_chunks = n_rows * q;
    const long long first = ((long long)row0 - L.hb) * R;     // sample index of chunk 0
    const float* yb = SRC_F32 ? y32 + (size_t)clip * src_stride : nullptr;
    const uint16_t* hb_ = SRC_F32 ? nullptr : src_hi + (size_t)clip * src_stride;
    const uint16_t* lb_ = SRC_F32 ? nullptr : src_lo + (size_t)clip * src_stride;
    const bool vec_ok = SRC_F32 ? ((reinterpret_cast<uintptr_t>(yb) & 15) == 0) : true;
    for (int i = threadIdx.x; i < n_chunks; i += VQT_LEVEL_THREADS) {
      const int r = i / q, c = i - r * q;
      const long long s0 = first + 8LL * i;
      uint4 c1 = make_uint4(0, 0, 0, 0), c2 = c1;
      if (s0 + 8 > 0 && s0 < n_sig) {
        if (SRC_F32) {
          float v[8];
          if (s0 >= 0 && s0 + 8 <= n_sig && vec_ok) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(yb + s0));
            const float4 b = __ldg(reinterpret_cast<const float4*>(yb + s0) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (s0 + e >= 0 && s0 + e < n_sig) ? __ldg(yb + s0 + e) : 0.f;
          }
          split8(v, c1, c2);
        } else if (s0 >= 0) {
          c1 = __ldg(reinterpret_cast<const uint4*>(hb_ + s0));
          c2 = __ldg(reinterpret_cast<const uint4*>(lb_ + s0));
          if (s0 + 8 > n_sig) {     // the tail of the buffer may hold a longer earlier signal
            uint16_t* p1 = reinterpret_cast<uint16_t*>(&c1);
            uint16_t* p2 = reinterpret_cast<uint16_t*>(&c2);
#pragma unroll
            for (int e = 0; e < 8; ++e) if (s0 + e >= n_sig) { p1[e] = 0; p2[e] = 0; }
          }
        }
      }
      const uint32_t off = (q == 1) ? (uint32_t)(16 * r) : (uint32_t)(c * L.a_lbo + 16 * r);
      *reinterpret_cast<uint4*>(sA1 + off) = c1;
      *reinterpret_cast<uint4*>(sA2 + off) = c2;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int used_cols = L.fb_b_col + max(32, L.fpr * 24);

  // zero the accumulators (every MMA accumulates: the decimator windows overlap)
  if (warp < 4) {
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < used_cols; c += 16) tmem_zero_32x16(tl + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t a_base[2] = {smem_u32(sA1), smem_u32(sA2)};
      const uint32_t b_base = smem_u32(sB);
#pragma unroll 1
      for (int i = 0; i < L.n_mma; ++i) {
        const VqtMma m = L.mma[i];
        umma_f16(tmem + m.d_col, desc_rows16(a_base[m.term] + m.a_off, (uint32_t)L.a_lbo),
                 desc_rows16(b_base + m.b_off, 16u * m.b_rows), umma_idesc_f16(128, m.n), 1u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();

  // ---- epilogue: warp -> TMEM lane quadrant (warp & 3), column share (warp >> 2) ----
  const int quad = warp & 3, half = warp >> 2;
  const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
  const long long g = (long long)row0 + quad * 32 + lane;      // global row
  if (L.dec_w > 0) {
    uint16_t* dh = dst_hi + (size_t)clip * dst_stride;
    uint16_t* dl = dst_lo + (size_t)clip * dst_stride;
    const int n_blk = (L.dec_w + 15) / 16;
    for (int kb = half; kb < n_blk; kb += 2) {
      uint32_t a[16], b[16];
      tmem_ld_32x16(tl + L.dec_a_col + 16 * kb, a);
      tmem_ld_32x16(tl + L.dec_b_col + 16 * kb, b);
      tmem_ld_wait();
      const long long t0 = g * L.dec_w + 16 * kb;
      float yv[16];
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        const float v = (__uint_as_float(a[o]) + __uint_as_float(b[o]) * (1.f / 2048.f)) * L.dec_scale;
        yv[o] = (t0 + o < n_valid) ? v : 0.f;
      }
      uint4 h1a, h2a, h1b, h2b;
      split8(yv, h1a, h2a);
      split8(yv + 8, h1b, h2b);
      if (L.dec_w >= 16) {
        if (t0 + 16 <= dst_stride) {
          reinterpret_cast<uint4*>(dh + t0)[0] = h1a; reinterpret_cast<uint4*>(dh + t0)[1] = h1b;
          reinterpret_cast<uint4*>(dl + t0)[0] = h2a; reinterpret_cast<uint4*>(dl + t0)[1] = h2b;
        } else if (t0 + 8 <= dst_stride) {
          reinterpret_cast<uint4*>(dh + t0)[0] = h1a;
          reinterpret_cast<uint4*>(dl + t0)[0] = h2a;
        }
      } else if (t0 + 4 <= dst_stride) {      // dec_w == 4: four outputs per row
        *reinterpret_cast<uint2*>(dh + t0) = make_uint2(h1a.x, h1a.y);
        *reinterpret_cast<uint2*>(dl + t0) = make_uint2(h2a.x, h2a.y);
      }
    }
  }
  for (int j = half; j < L.fpr; j += 2) {
    const long long f = g * L.fpr + j;
    uint32_t fa[48], fb[24];
    tmem_ld_32x32(tl + L.fb_a_col + 48 * j, fa);
    tmem_ld_32x16(tl + L.fb_a_col + 48 * j + 32, fa + 32);
    tmem_ld_32x16(tl + L.fb_b_col + 24 * j, fb);
    tmem_ld_32x8(tl + L.fb_b_col + 24 * j + 16, fb + 16);
    tmem_ld_wait();
    if (f < n_frames) {
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const float re = __uint_as_float(fa[2 * k]) + (__uint_as_float(fa[24 + 2 * k]) + __uint_as_float(fb[2 * k])) * (1.f / 2048.f);
        const float im = __uint_as_float(fa[2 * k + 1]) + (__uint_as_float(fa[25 + 2 * k]) + __uint_as_float(fb[2 * k + 1])) * (1.f / 2048.f);
        const int bin = L.bin0 + k;
        out[((size_t)clip * n_bins + bin) * n_frames + f] =
            logf(sqrtf(re * re + im * im) * (__ldg(inv_sqrt_len + bin) * L.fb_scale) + 1e-9f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, (uint32_t)L.tmem_cols); }
}

// ---- frames whose window crosses a clip edge: reflect padding, direct fp32 evaluation -------------
struct VqtEdgeParams {
  int n_oct, bpo, n_bins, n_frames;
  int n_fft[ZNS_VQT_MAX_OCT], hop[ZNS_VQT_MAX_OCT], n_sig[ZNS_VQT_MAX_OCT];
  long long stride[ZNS_VQT_MAX_OCT];
  const float* coef[ZNS_VQT_MAX_OCT];      // [n][2][bpo/2][2]
  const uint16_t* hi[ZNS_VQT_MAX_OCT];
  const uint16_t* lo[ZNS_VQT_MAX_OCT];
  float scale[ZNS_VQT_MAX_OCT];
};

__device__ __forceinline__ int reflect_idx(long long qq, int n) {
  if (n == 1) return 0;
  const long long period = 2LL * (n - 1);
  qq %= period;
  if (qq < 0) qq += period;
  return (int)(qq < n ? qq : period - qq);
}

__global__ void __launch_bounds__(192)
vqt_edge_kernel(const __grid_constant__ VqtEdgeParams P, const float* __restrict__ y32, long long y_stride,
                const float* __restrict__ inv_sqrt_len, float* __restrict__ out) {
  const int oct = blockIdx.x, clip = blockIdx.z;
  const int nf = P.n_fft[oct], hop = P.hop[oct], n = P.n_sig[oct], F = P.n_frames;
  // left: frames t with t*hop < nf/2 ; right: frames with t*hop + nf/2 > n
  const int n_left = min(F, (nf / 2 + hop - 1) / hop);
  int t_right = (n >= nf / 2) ? (n - nf / 2) / hop + 1 : 0;
  t_right = max(t_right, n_left);
  const int n_edge = n_left + max(0, F - t_right);
  const int hb = P.bpo / 2;
  for (int w = threadIdx.x; w < n_edge * P.bpo; w += blockDim.x) {
    const int e = w / P.bpo, k = w - e * P.bpo;
    const int t = e < n_left ? e : t_right + (e - n_left);
    float re = 0.f, im = 0.f;
    const float* cf = P.coef[oct] + ((size_t)(k / hb) * hb + k % hb) * 2;
    for (int i = 0; i < nf; ++i) {
      const int idx = reflect_idx((long long)t * hop + i - nf / 2, n);
      float s;
      if (oct == 0) s = __ldg(y32 + (size_t)clip * y_stride + idx);
      else {
        const size_t o = (size_t)clip * P.stride[oct] + idx;
        s = __half2float(__ushort_as_half(P.hi[oct][o])) + __half2float(__ushort_as_half(P.lo[oct][o])) * (1.f / 2048.f);
      }
      const float* cn = cf + (size_t)i * P.bpo * 2;
      re = fmaf(cn[0], s, re);
      im = fmaf(cn[1], s, im);
    }
    const int bin = P.n_bins - P.bpo * (oct + 1) + k;
    out[((size_t)clip * P.n_bins + bin) * F + t] = logf(sqrtf(re * re + im * im) * __ldg(inv_sqrt_len + bin) + 1e-9f);
  }
}

// ---------------------------------------------------------------------------------------------
// host: forward
// ---------------------------------------------------------------------------------------------
int vqt_umma_forward(zns_vqt_plan* p, const float* y, int batch, int n_samples, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int n_frames = zns_vqt_num_frames(n_samples, p->hop);
  static bool attr_set[64] = {};
  int dev = 0;
  ZNS_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(vqt_level_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(vqt_level_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set[dev] = true;
  }
  VqtEdgeParams E;
  memset(&E, 0, sizeof(E));
  E.n_oct = p->n_oct; E.bpo = p->bpo; E.n_bins = p->n_bins; E.n_frames = n_frames;
  int n_cur = n_samples;
  for (int i = 0; i < p->n_oct; ++i) {
    const VqtLevelDev& L = p->level[i];
    const int R = 8 * L.q;
    const int n_valid = n_cur / 2, n_next = (n_cur + 1) / 2;
    int rows = (n_frames + L.fpr - 1) / L.fpr;
    if (L.dec_w > 0) rows = std::max(rows, (n_cur + R - 1) / R);
    dim3 grid((rows + 127) / 128, 1, batch);
    const size_t smem = level_smem(L);
    const bool last = (i == p->n_oct - 1);
    if (i == 0)
      vqt_level_kernel<true><<<grid, VQT_LEVEL_THREADS, smem, st>>>(
          L, y, nullptr, nullptr, n_cur, (long long)n_samples, p->d_bimg[i], p->d_inv_sqrt_len, out, n_frames, p->n_bins,
          last ? nullptr : p->d_hi[i + 1], last ? nullptr : p->d_lo[i + 1], n_valid, last ? 0 : p->sig_stride[i + 1]);
    else
      vqt_level_kernel<false><<<grid, VQT_LEVEL_THREADS, smem, st>>>(
          L, nullptr, p->d_hi[i], p->d_lo[i], n_cur, p->sig_stride[i], p->d_bimg[i], p->d_inv_sqrt_len, out, n_frames,
          p->n_bins, last ? nullptr : p->d_hi[i + 1], last ? nullptr : p->d_lo[i + 1], n_valid,
          last ? 0 : p->sig_stride[i + 1]);
    ZNS_CHECK_LAUNCH();
    E.n_fft[i] = L.n_fft; E.hop[i] = L.hop; E.n_sig[i] = n_cur; E.stride[i] = p->sig_stride[i];
    E.coef[i] = p->d_coef[i]; E.hi[i] = p->d_hi[i]; E.lo[i] = p->d_lo[i];
    n_cur = n_next;
  }
  vqt_edge_kernel<<<dim3(p->n_oct, 1, batch), 192, 0, st>>>(E, y, (long long)n_samples, p->d_inv_sqrt_len, out);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

// Test hook (host only, no CUDA call): the level plan and its coefficient image for the reference's VQT / CQT
// configuration, so that the MMA list can be replayed in numpy against the oracle on a machine without a GPU.
extern "C" int zns_dbg_vqt_level_plan(int sr, int hop, int n_bins, int bpo, double fmin, double gamma_in, int level,
                                      void* level_struct, int struct_bytes, uint16_t* bimg, int bimg_halfwords) {
  ZNS_REQUIRE(level_struct && bpo > 0 && n_bins % bpo == 0, "bad arguments");
  ZNS_REQUIRE(struct_bytes == (int)sizeof(VqtLevelDev), "struct size mismatch (%d vs %d)", struct_bytes, (int)sizeof(VqtLevelDev));
  zns_vqt_plan* p = (zns_vqt_plan*)calloc(1, sizeof(zns_vqt_plan));
  if (!p) return zns_set_error(ZNS_ERR_ALLOC, "out of host memory");
  p->sr = sr; p->hop = hop; p->n_bins = n_bins; p->bpo = bpo; p->n_oct = n_bins / bpo;
  int rc = ZNS_OK;
  if (level < 0 || level >= p->n_oct || p->n_oct > ZNS_VQT_MAX_OCT) rc = zns_set_error(ZNS_ERR_INVALID, "bad level");
  std::vector<float> re((size_t)bpo * 1024), im((size_t)bpo * 1024);
  if (!rc) {
    int nf = 0;
    rc = zns_vqt_basis_host(sr, n_bins, bpo, fmin, gamma_in, level, re.data(), im.data(), &nf);
    if (!rc) {
      p->n_fft[level] = nf;
      p->coef_inv_scale[level] = 1.f / vqt_coef_scale(re.data(), im.data(), bpo * nf);
      double taps[32];
      zns_vqt_decimator_taps_host(taps);
      if (build_level(p, level, fmin, gamma_in, taps) != 0) rc = zns_set_error(ZNS_ERR_INVALID, "level geometry not supported");
    }
  }
  if (!rc) {
    memcpy(level_struct, &p->level[level], sizeof(VqtLevelDev));
    if (bimg) {
      if (bimg_halfwords < p->level[level].b_bytes / 2) rc = zns_set_error(ZNS_ERR_INVALID, "coefficient buffer too small");
      else memcpy(bimg, p->h_bimg[level]->data(), p->level[level].b_bytes);
    }
  }
  vqt_umma_free(p);
  free(p);
  return rc;
}
