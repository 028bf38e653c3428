// Test-only SIMT reference convolutions on the act layout (fp32 accumulation, no tensor cores,
// no shared-memory staging): deliberately the most literal loops, used by tests to localise errors
// of the tcgen05 kernels at sizes the CPU oracle cannot reach.  Not on the product path.
#include <algorithm>

#include "common.cuh"

typedef __nv_bfloat16 bf16;

// element i of a 16-bit tensor that is fp16 (f16 != 0) or bf16
__device__ __forceinline__ float ld16(const bf16* p, size_t i, int f16) {
  return f16 ? __half2float(reinterpret_cast<const __half*>(p)[i]) : __bfloat162float(p[i]);
}

__global__ void conv_fwd_simt_kernel(const bf16* __restrict__ in, const bf16* __restrict__ wpk,
                                     const float* __restrict__ bias, const bf16* __restrict__ mask,
                                     bf16* __restrict__ out, int G, int H, int W, int Cin, int Cout, int kh, int kw,
                                     int relu, float scale, size_t total, int fmt) {
  const int ph = kh / 2, pw = kw / 2;
  const int in16 = fmt & ZNS_FMT_IN_F16, w16 = fmt & ZNS_FMT_W_F16, out16 = fmt & ZNS_FMT_OUT_F16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int n = (int)(r % Cout); r /= Cout;
    const int b8 = (int)(r % 8); r /= 8;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int g = (int)r;
    float acc = bias ? bias[n] : 0.f;
    for (int rr = 0; rr < kh; ++rr) {
      const int hh = h + rr - ph;
      if (hh < 0 || hh >= H) continue;
      for (int ss = 0; ss < kw; ++ss) {
        const int ww = w + ss - pw;
        if (ww < 0 || ww >= W) continue;
        const bf16* xp = in + zns_act_index(g, hh, ww, b8, 0, H, W, Cin);
        const bf16* wp = wpk + ((size_t)(rr * kw + ss) * Cout + n) * Cin;
        for (int c = 0; c < Cin; ++c) acc = fmaf(ld16(xp, c, in16), ld16(wp, c, w16), acc);
      }
    }
    if (relu) acc = fmaxf(acc, 0.f);
    if (mask && !act_bits_positive(reinterpret_cast<const uint16_t*>(mask)[i])) acc = 0.f;
    if (out16) reinterpret_cast<__half*>(out)[i] = __float2half_rn(acc * scale);
    else out[i] = __float2bfloat16(acc * scale);
  }
}

extern "C" int zns_dbg_conv_fwd_simt(const zns_conv_desc* d, const void* in, const void* wpk, const float* bias,
                                     const void* mask, void* out, void* stream) {
  ZNS_REQUIRE(d && in && wpk && out, "NULL argument");
  const int G = zns_groups(d->batch);
  const size_t total = (size_t)G * d->H * d->W * 8 * d->c_out;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 64);
  conv_fwd_simt_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)in, (const bf16*)wpk, bias,
                                                                 (const bf16*)mask, (bf16*)out, G, d->H, d->W, d->c_in,
                                                                 d->c_out, d->kh, d->kw, d->relu,
                                                                 d->out_scale == 0.f ? 1.f : d->out_scale, total, d->fmt);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

// dwpk[tap][n][c] += sum_p dy[p][n] * x[p + tap][c]; one warp per output element, lanes stride the
// positions.
__global__ void conv_wgrad_simt_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                       float* __restrict__ dwpk, int G, int H, int W, int Cin, int Cout, int kh, int kw,
                                       size_t total) {
  const int ph = kh / 2, pw = kw / 2;
  const int lane = threadIdx.x & 31;
  const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t n_pos = (size_t)G * H * W * 8;
  for (size_t i = warp_global; i < total; i += n_warps) {
    size_t r = i;
    const int c = (int)(r % Cin); r /= Cin;
    const int n = (int)(r % Cout); r /= Cout;
    const int tap = (int)r;
    const int rr = tap / kw, ss = tap - rr * kw;
    float acc = 0.f;
    for (size_t p = lane; p < n_pos; p += 32) {
      size_t q = p;
      const int b8 = (int)(q % 8); q /= 8;
      const int w = (int)(q % W); q /= W;
      const int h = (int)(q % H); q /= H;
      const int g = (int)q;
      const int hh = h + rr - ph, ww = w + ss - pw;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      acc = fmaf(__bfloat162float(dy[p * Cout + n]), __bfloat162float(x[zns_act_index(g, hh, ww, b8, c, H, W, Cin)]), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dwpk[i] += acc;
  }
}

extern "C" int zns_dbg_conv_wgrad_simt(const zns_conv_desc* d, const void* x, const void* dy, float* dwpk, void* stream) {
  ZNS_REQUIRE(d && x && dy && dwpk, "NULL argument");
  const int G = zns_groups(d->batch);
  const size_t total = (size_t)d->kh * d->kw * d->c_out * d->c_in;
  const int blocks = (int)std::min<size_t>((total * 32 + 255) / 256, 148 * 32);
  conv_wgrad_simt_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, dwpk, G, d->H, d->W,
                                                                   d->c_in, d->c_out, d->kh, d->kw, total);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}
