// Bandwidth-bound pieces of the encoder path on the act layout (bf16 [G][H][W][8][C]):
// cv1 (C_in = 1) forward / weight gradient, frequency max-pool + ReLU + dropout and its backward,
// the fc1 + sigmoid head and its backward, layout converters, weight packing, bias gradients,
// Down_CNN merge and the flat Adam update.
// Reference semantics: /root/reference/zeroNoteSamba/models/models.py:16-74,85-103,139-150 and
// torch.optim.Adam as constructed at /root/reference/zeroNoteSamba/pretext.py:202.
#include <algorithm>

#include "common.cuh"

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t dropout_threshold(float p) {
  // keep iff hash >= threshold  (P[drop] = p)
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

// ---------------------------------------------------------------------------------------------
// cv1 forward: x fp32 [B][H][W] -> act [G][H][W][8][64], 3x11 taps, pad (1,5), ReLU, dropout
// ---------------------------------------------------------------------------------------------
#define C1_KH 3
#define C1_KW 11
#define C1_TAPS 33
#define C1_CO 64

// Per-branch pointers of the kernels that run the anchor and the positive encoder in ONE launch (grid.z or grid.y carries
// the branch): the reference's two DS_CNN branches (models.py:114-124) have identical geometry.
struct C1FwdBr { const float* x; const float* weight; const float* bias; bf16* out; bf16* out2; };
struct C1FwdArgs { C1FwdBr br[2]; int G; };

// Two adjacent frames per thread: a weight vector read from shared memory feeds both (the first version, one frame per
// thread, ran at 87 % of the L1 / shared-memory pipe with 35 % occupancy: profiles/r02_ncu_cv1_summary.txt) and their 3 x 12
// input patch overlaps in 10 of 11 columns.
#define C1_NPOS 2
__global__ void __launch_bounds__(256) conv1_fwd_kernel(const __grid_constant__ C1FwdArgs a, long long clip_stride, long long row_stride,
                                                        int B, int H, int W, float drop_p, uint32_t seed,
                                                        const uint32_t* __restrict__ seed_dev, uint32_t stream_id, int f16) {
  __shared__ __align__(16) float ws[C1_TAPS][C1_CO];
  const int branch = blockIdx.z / a.G;
  const float* __restrict__ x = a.br[branch].x;
  const float* __restrict__ weight = a.br[branch].weight;
  const float* __restrict__ bias = a.br[branch].bias;
  bf16* __restrict__ out = a.br[branch].out;
  bf16* __restrict__ out2 = a.br[branch].out2;
  stream_id += branch;
  if (seed_dev) seed ^= __ldg(seed_dev) * 0x9E3779B9u;
  __shared__ float bs[C1_CO];
  for (int i = threadIdx.x; i < C1_TAPS * C1_CO; i += 256) {
    int c = i / C1_TAPS, t = i - c * C1_TAPS;  // weight is [64][1][3][11]
    ws[t][c] = weight[i];
  }
  if (threadIdx.x < C1_CO) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int b8 = threadIdx.x & 7, wl = threadIdx.x >> 3;
  const int g = blockIdx.z - branch * a.G, h = blockIdx.y, w = blockIdx.x * (32 * C1_NPOS) + C1_NPOS * wl;
  if (w >= W) return;
  const int n_pos = min(C1_NPOS, W - w);
  const int b = g * 8 + b8;
  const size_t e0 = zns_act_index(g, h, w, b8, 0, H, W, C1_CO);        // frame w; frame w + 1 is 8 * 64 elements further
  const size_t e_step = 8 * C1_CO;
  if (b >= B) {
    for (int p = 0; p < n_pos; ++p) {
      uint4* dst = reinterpret_cast<uint4*>(out + e0 + p * e_step);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
      if (out2) {
        uint4* dst2 = reinterpret_cast<uint4*>(out2 + e0 + p * e_step);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst2[i] = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }
  float xin[C1_KH][C1_KW + C1_NPOS - 1];
  const float* xb = x + (size_t)b * clip_stride;
#pragma unroll
  for (int r = 0; r < C1_KH; ++r)
#pragma unroll
    for (int s = 0; s < C1_KW + C1_NPOS - 1; ++s) {
      int hh = h + r - 1, ww = w + s - 5;
      xin[r][s] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xb + (size_t)hh * row_stride + ww) : 0.f;
    }
  const uint32_t thr = dropout_threshold(drop_p);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#pragma unroll 1
  for (int c0 = 0; c0 < C1_CO; c0 += 8) {
    float acc[C1_NPOS][8];
#pragma unroll
    for (int p = 0; p < C1_NPOS; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[p][i] = bs[c0 + i];
#pragma unroll
    for (int r = 0; r < C1_KH; ++r)
#pragma unroll
      for (int sx = 0; sx < C1_KW; ++sx) {
        const float4 w0 = *reinterpret_cast<const float4*>(&ws[r * C1_KW + sx][c0]);
        const float4 w1 = *reinterpret_cast<const float4*>(&ws[r * C1_KW + sx][c0 + 4]);
#pragma unroll
        for (int p = 0; p < C1_NPOS; ++p) {
          const float v = xin[r][sx + p];
          acc[p][0] = fmaf(w0.x, v, acc[p][0]); acc[p][1] = fmaf(w0.y, v, acc[p][1]);
          acc[p][2] = fmaf(w0.z, v, acc[p][2]); acc[p][3] = fmaf(w0.w, v, acc[p][3]);
          acc[p][4] = fmaf(w1.x, v, acc[p][4]); acc[p][5] = fmaf(w1.y, v, acc[p][5]);
          acc[p][6] = fmaf(w1.z, v, acc[p][6]); acc[p][7] = fmaf(w1.w, v, acc[p][7]);
        }
      }
#pragma unroll
    for (int p = 0; p < C1_NPOS; ++p) {
      if (p < n_pos) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v = fmaxf(acc[p][i], 0.f);
          if (drop_p > 0.f) v = (zns_hash32(e0 + p * e_step + c0 + i, seed, stream_id) >= thr) ? v * keep_scale : 0.f;
          acc[p][i] = v;
        }
        *reinterpret_cast<uint4*>(out + e0 + p * e_step + c0) =
            make_uint4(pack_act2(acc[p][0], acc[p][1], f16), pack_act2(acc[p][2], acc[p][3], f16),
                       pack_act2(acc[p][4], acc[p][5], f16), pack_act2(acc[p][6], acc[p][7], f16));
        if (out2)   // bf16 copy: the x operand of cv2's weight gradient
          *reinterpret_cast<uint4*>(out2 + e0 + p * e_step + c0) =
              make_uint4(pack_bf16x2(acc[p][0], acc[p][1]), pack_bf16x2(acc[p][2], acc[p][3]),
                         pack_bf16x2(acc[p][4], acc[p][5]), pack_bf16x2(acc[p][6], acc[p][7]));
      }
    }
  }
}

extern "C" int zns_conv1_fwd_nbr(int n_br, const float* const* x, long long clip_stride, long long row_stride,
                                 const float* const* weight, const float* const* bias, void* const* out_act, int batch, int H,
                                 int W, float drop_p, uint32_t seed, const uint32_t* seed_dev, uint32_t rng_stream, int out_f16,
                                 void* const* out_act_bf16, void* stream) {
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(x && weight && bias && out_act, "NULL argument");
  ZNS_REQUIRE(batch > 0 && H > 0 && W > 0 && drop_p >= 0.f && drop_p < 1.f, "bad conv1 geometry");
  C1FwdArgs a;
  memset(&a, 0, sizeof(a));
  a.G = zns_groups(batch);
  for (int b = 0; b < n_br; ++b) {
    ZNS_REQUIRE(x[b] && weight[b] && bias[b] && out_act[b], "NULL tensor for branch %d", b);
    a.br[b] = C1FwdBr{x[b], weight[b], bias[b], (bf16*)out_act[b], out_act_bf16 ? (bf16*)out_act_bf16[b] : nullptr};
  }
  dim3 grid((W + 32 * C1_NPOS - 1) / (32 * C1_NPOS), H, a.G * n_br);
  conv1_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, clip_stride, row_stride, batch, H, W, drop_p, seed, seed_dev,
                                                           rng_stream, out_f16);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

extern "C" int zns_conv1_fwd(const float* x, long long clip_stride, long long row_stride, const float* weight,
                             const float* bias, void* out_act, int batch, int H, int W, float drop_p, uint32_t seed,
                             const uint32_t* seed_dev, uint32_t rng_stream, int out_f16, void* out_act_bf16,
                             void* stream) {
  return zns_conv1_fwd_nbr(1, &x, clip_stride, row_stride, &weight, &bias, &out_act, batch, H, W, drop_p, seed, seed_dev,
                           rng_stream, out_f16, out_act_bf16 ? &out_act_bf16 : nullptr, stream);
}

// ---------------------------------------------------------------------------------------------
// cv1 weight / bias gradient
// ---------------------------------------------------------------------------------------------
// block = (group, row h, 32 frames); thread = (output channel n, kernel row r): the 42 input samples
// of row r are held in registers and slide under the 11 taps, so the inner loop is 11 FMAs per
// shared-memory load of dy.
#define C1W_TILES 5   // 32-frame tiles per block: fewer same-address atomics on the 64 x 33 gradient

struct C1WgBr { const bf16* dy; const float* x; float* dw; float* db; };
struct C1WgArgs { C1WgBr br[2]; int G; };

__global__ void __launch_bounds__(192) conv1_wgrad_kernel(const __grid_constant__ C1WgArgs a, long long clip_stride, long long row_stride,
                                                          int B, int H, int W) {
  __shared__ float xs[C1_KH][8][32 + C1_KW - 1];
  __shared__ __align__(16) bf16 dys[32 * 8 * C1_CO];      // [frame][clip slot][64 channels], 32 KB
  const int branch = blockIdx.z / a.G;
  const bf16* __restrict__ dy = a.br[branch].dy;
  const float* __restrict__ x = a.br[branch].x;
  float* __restrict__ dw = a.br[branch].dw;
  float* __restrict__ db = a.br[branch].db;
  const int g = blockIdx.z - branch * a.G, h = blockIdx.y;
  const int n = threadIdx.x & 63, r = threadIdx.x >> 6;   // r = 0..2
  float acc[C1_KW];
#pragma unroll
  for (int j = 0; j < C1_KW; ++j) acc[j] = 0.f;
  float accb = 0.f;
  for (int tile = 0; tile < C1W_TILES; ++tile) {
    const int w0 = (blockIdx.x * C1W_TILES + tile) * 32;
    if (w0 >= W) break;
    const int wmax = min(32, W - w0);
    if (tile > 0) __syncthreads();
    {
      // the 32-frame dy tile of this (group, row) is one contiguous run: all 16-byte loads in flight at once
      const uint4* src = reinterpret_cast<const uint4*>(dy + zns_act_index(g, h, w0, 0, 0, H, W, C1_CO));
      const int n_vec = wmax * 8 * C1_CO / 8;
      uint4 v[11];
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const int i = threadIdx.x + k * 192;
        v[k] = i < n_vec ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const int i = threadIdx.x + k * 192;
        if (i < 2048) reinterpret_cast<uint4*>(dys)[i] = v[k];
      }
    }
    for (int i = threadIdx.x; i < C1_KH * 8 * 42; i += 192) {
      int rr = i / (8 * 42), rem = i - rr * 8 * 42;
      int b8 = rem / 42, j = rem - b8 * 42;
      int hh = h + rr - 1, ww = w0 + j - 5, b = g * 8 + b8;
      xs[rr][b8][j] = (b < B && hh >= 0 && hh < H && ww >= 0 && ww < W)
                          ? __ldg(x + (size_t)b * clip_stride + (size_t)hh * row_stride + ww) : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int b8 = 0; b8 < 8; ++b8) {
      float xr[32 + C1_KW - 1];
#pragma unroll
      for (int j = 0; j < 32 + C1_KW - 1; ++j) xr[j] = xs[r][b8][j];
#pragma unroll
      for (int wl = 0; wl < 32; ++wl) {
        const float d = __bfloat162float(dys[(wl * 8 + b8) * C1_CO + n]);   // zero beyond wmax
        accb += d;
#pragma unroll
        for (int sft = 0; sft < C1_KW; ++sft) acc[sft] = fmaf(d, xr[wl + sft], acc[sft]);
      }
    }
  }
#pragma unroll
  for (int sft = 0; sft < C1_KW; ++sft) atomicAdd(dw + n * C1_TAPS + r * C1_KW + sft, acc[sft]);
  if (r == 0) atomicAdd(db + n, accb);
}

extern "C" int zns_conv1_wgrad_nbr(int n_br, const void* const* dy_act, const float* const* x, long long clip_stride,
                                   long long row_stride, float* const* dw, float* const* db, int batch, int H, int W,
                                   void* stream) {
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(dy_act && x && dw && db, "NULL argument");
  C1WgArgs a;
  memset(&a, 0, sizeof(a));
  a.G = zns_groups(batch);
  for (int b = 0; b < n_br; ++b) {
    ZNS_REQUIRE(dy_act[b] && x[b] && dw[b] && db[b], "NULL tensor for branch %d", b);
    a.br[b] = C1WgBr{(const bf16*)dy_act[b], x[b], dw[b], db[b]};
  }
  dim3 grid((W + 32 * C1W_TILES - 1) / (32 * C1W_TILES), H, a.G * n_br);
  conv1_wgrad_kernel<<<grid, 192, 0, (cudaStream_t)stream>>>(a, clip_stride, row_stride, batch, H, W);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

extern "C" int zns_conv1_wgrad(const void* dy_act, const float* x, long long clip_stride, long long row_stride, float* dw,
                               float* db, int batch, int H, int W, void* stream) {
  return zns_conv1_wgrad_nbr(1, &dy_act, &x, clip_stride, row_stride, &dw, &db, batch, H, W, stream);
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d((pool,1)) -> ReLU -> Dropout, and its backward (first arg-max routing)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float* f, bool f16 = false) {
  if (f16) {
    const __half2* p = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 v = __half22float2(p[i]);
      f[2 * i] = v.x;
      f[2 * i + 1] = v.y;
    }
    return;
  }
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 v = __bfloat1622float2(p[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f, bool f16 = false) {
  return make_uint4(pack_act2(f[0], f[1], f16), pack_act2(f[2], f[3], f16), pack_act2(f[4], f[5], f16), pack_act2(f[6], f[7], f16));
}

struct PoolBr { const uint4* y; uint4* out; uint4* out2; const uint4* dp; uint4* dy; };
struct PoolArgs { PoolBr br[2]; };

__global__ void pool_fwd_kernel(const __grid_constant__ PoolArgs a, size_t n_vec, size_t row_vec,
                                int Hp, int pool, float drop_p, uint32_t seed, const uint32_t* __restrict__ seed_dev,
                                uint32_t stream_id, int f16) {
  const uint4* __restrict__ y = a.br[blockIdx.y].y;
  uint4* __restrict__ out = a.br[blockIdx.y].out;
  uint4* __restrict__ out2 = a.br[blockIdx.y].out2;
  stream_id += blockIdx.y;
  if (seed_dev) seed ^= __ldg(seed_dev) * 0x9E3779B9u;
  // vec index = ((g*Hp + hp) * row_vec + r), row_vec = W*8*C/8
  const uint32_t thr = dropout_threshold(drop_p);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
    const size_t ghp = i / row_vec, r = i - ghp * row_vec;
    const size_t g = ghp / Hp, hp = ghp - g * Hp;
    const uint4* src = y + ((g * Hp + hp) * pool) * row_vec + r;
    float m[8], v[8];
    unpack8(__ldg(src), m, f16);
    for (int k = 1; k < pool; ++k) {
      unpack8(__ldg(src + (size_t)k * row_vec), v, f16);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaxf(m[j], 0.f);
      if (drop_p > 0.f) t = (zns_hash32(i * 8 + j, seed, stream_id) >= thr) ? t * keep_scale : 0.f;
      m[j] = t;
    }
    out[i] = pack8(m, f16);
    if (out2) out2[i] = pack8(m, false);   // bf16 copy: the x operand of the next layer's weight gradient
  }
}

extern "C" int zns_pool_fwd_nbr(int n_br, const void* const* y_act, void* const* out_act, int batch, int H, int W, int C,
                                int pool, float drop_p, uint32_t seed, const uint32_t* seed_dev, uint32_t rng_stream, int act_f16,
                                void* const* out_act_bf16, void* stream) {
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(y_act && out_act, "NULL argument");
  ZNS_REQUIRE(pool >= 1 && H % pool == 0 && C % 8 == 0, "pool %d must divide H %d", pool, H);
  PoolArgs a;
  memset(&a, 0, sizeof(a));
  for (int b = 0; b < n_br; ++b) {
    ZNS_REQUIRE(y_act[b] && out_act[b], "NULL tensor for branch %d", b);
    a.br[b].y = (const uint4*)y_act[b]; a.br[b].out = (uint4*)out_act[b];
    a.br[b].out2 = out_act_bf16 ? (uint4*)out_act_bf16[b] : nullptr;
  }
  const int G = zns_groups(batch), Hp = H / pool;
  const size_t row_vec = (size_t)W * 8 * C / 8;
  const size_t n_vec = (size_t)G * Hp * row_vec;
  const int blocks = (int)std::min<size_t>((n_vec + 255) / 256, 148 * 16);
  pool_fwd_kernel<<<dim3(blocks, n_br), 256, 0, (cudaStream_t)stream>>>(a, n_vec, row_vec, Hp, pool, drop_p, seed, seed_dev,
                                                                        rng_stream, act_f16);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

extern "C" int zns_pool_fwd(const void* y_act, void* out_act, int batch, int H, int W, int C, int pool, float drop_p,
                            uint32_t seed, const uint32_t* seed_dev, uint32_t rng_stream, int act_f16, void* out_act_bf16,
                            void* stream) {
  return zns_pool_fwd_nbr(1, &y_act, &out_act, batch, H, W, C, pool, drop_p, seed, seed_dev, rng_stream, act_f16,
                          out_act_bf16 ? &out_act_bf16 : nullptr, stream);
}

__global__ void pool_bwd_kernel(const __grid_constant__ PoolArgs a, size_t n_vec, size_t row_vec, int Hp, int pool, int y_f16) {
  const uint4* __restrict__ y = a.br[blockIdx.y].y;
  const uint4* __restrict__ dp = a.br[blockIdx.y].dp;
  uint4* __restrict__ dy = a.br[blockIdx.y].dy;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
    const size_t ghp = i / row_vec, r = i - ghp * row_vec;
    const size_t g = ghp / Hp, hp = ghp - g * Hp;
    const size_t base = ((g * Hp + hp) * pool) * row_vec + r;
    float m[8], v[8], d[8];
    int arg[8];
    unpack8(__ldg(y + base), m, y_f16);
#pragma unroll
    for (int j = 0; j < 8; ++j) arg[j] = 0;
    for (int k = 1; k < pool; ++k) {
      unpack8(__ldg(y + base + (size_t)k * row_vec), v, y_f16);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[j] > m[j]) { m[j] = v[j]; arg[j] = k; }
    }
    unpack8(__ldg(dp + i), d);
    for (int k = 0; k < pool; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (arg[j] == k) ? d[j] : 0.f;
      dy[base + (size_t)k * row_vec] = pack8(v);
    }
  }
}

extern "C" int zns_pool_bwd_nbr(int n_br, const void* const* y_act, const void* const* dpool_act, void* const* dy_act, int batch,
                                int H, int W, int C, int pool, int y_f16, void* stream) {
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(y_act && dpool_act && dy_act, "NULL argument");
  ZNS_REQUIRE(pool >= 1 && H % pool == 0 && C % 8 == 0, "pool %d must divide H %d", pool, H);
  PoolArgs a;
  memset(&a, 0, sizeof(a));
  for (int b = 0; b < n_br; ++b) {
    ZNS_REQUIRE(y_act[b] && dpool_act[b] && dy_act[b], "NULL tensor for branch %d", b);
    a.br[b].y = (const uint4*)y_act[b]; a.br[b].dp = (const uint4*)dpool_act[b]; a.br[b].dy = (uint4*)dy_act[b];
  }
  const int G = zns_groups(batch), Hp = H / pool;
  const size_t row_vec = (size_t)W * 8 * C / 8;
  const size_t n_vec = (size_t)G * Hp * row_vec;
  const int blocks = (int)std::min<size_t>((n_vec + 255) / 256, 148 * 16);
  pool_bwd_kernel<<<dim3(blocks, n_br), 256, 0, (cudaStream_t)stream>>>(a, n_vec, row_vec, Hp, pool, y_f16);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

extern "C" int zns_pool_bwd(const void* y_act, const void* dpool_act, void* dy_act, int batch, int H, int W, int C,
                            int pool, int y_f16, void* stream) {
  return zns_pool_bwd_nbr(1, &y_act, &dpool_act, &dy_act, batch, H, W, C, pool, y_f16, stream);
}

// Backward of the pooling fused into a convolution epilogue (zns_conv_pool_fwd): the forward pass kept one byte per pooled
// element, the row of the first maximum inside its window; dy[row] = dp if row == arg else 0.  Reads dp (2 B) + arg (1 B) per
// pooled element instead of the whole pre-pool activation.
struct UnpoolBr { const uint4* dp; const uint2* arg; uint4* dy; };
struct UnpoolArgs { UnpoolBr br[2]; };

__global__ void pool_bwd_arg_kernel(const __grid_constant__ UnpoolArgs a, size_t n_vec, size_t row_vec, int Hp, int pool) {
  const uint4* __restrict__ dp = a.br[blockIdx.y].dp;
  const uint2* __restrict__ arg = a.br[blockIdx.y].arg;
  uint4* __restrict__ dy = a.br[blockIdx.y].dy;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
    const size_t ghp = i / row_vec, r = i - ghp * row_vec;
    const size_t base = (ghp * pool) * row_vec + r;          // (g * Hp + hp) * pool rows of the unpooled tensor
    const uint4 d = __ldg(dp + i);
    const uint2 am = __ldg(arg + i);                          // eight row indices
    const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
    for (int k = 0; k < pool; ++k) {
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t a0 = ((q < 2 ? am.x : am.y) >> (16 * (q & 1))) & 0xFFu, a1 = ((q < 2 ? am.x : am.y) >> (16 * (q & 1) + 8)) & 0xFFu;
        o[q] = (a0 == (uint32_t)k ? (dw[q] & 0xFFFFu) : 0u) | (a1 == (uint32_t)k ? (dw[q] & 0xFFFF0000u) : 0u);
      }
      dy[base + (size_t)k * row_vec] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

extern "C" int zns_pool_bwd_arg_nbr(int n_br, const void* const* argmax, const void* const* dpool_act, void* const* dy_act,
                                    int batch, int H, int W, int C, int pool, void* stream) {
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(argmax && dpool_act && dy_act, "NULL argument");
  ZNS_REQUIRE(pool >= 1 && pool <= 255 && H % pool == 0 && C % 8 == 0, "pool %d must divide H %d", pool, H);
  UnpoolArgs a;
  memset(&a, 0, sizeof(a));
  for (int b = 0; b < n_br; ++b) {
    ZNS_REQUIRE(argmax[b] && dpool_act[b] && dy_act[b], "NULL tensor for branch %d", b);
    a.br[b].dp = (const uint4*)dpool_act[b]; a.br[b].arg = (const uint2*)argmax[b]; a.br[b].dy = (uint4*)dy_act[b];
  }
  const int G = zns_groups(batch), Hp = H / pool;
  const size_t row_vec = (size_t)W * 8 * C / 8;
  const size_t n_vec = (size_t)G * Hp * row_vec;
  const int blocks = (int)std::min<size_t>((n_vec + 255) / 256, 148 * 16);
  pool_bwd_arg_kernel<<<dim3(blocks, n_br), 256, 0, (cudaStream_t)stream>>>(a, n_vec, row_vec, Hp, pool);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

// ---------------------------------------------------------------------------------------------
// head: Conv1d(128, 1, k=1) + Sigmoid + flatten, forward and backward
// block = 256 threads = 32 positions x 8 lanes; a lane owns 16 channels
// ---------------------------------------------------------------------------------------------
#define HD_C 128

struct HeadBr { const bf16* x; const float* w; const float* bias; float* emb; const float* d_emb; float* dw; float* dbias; bf16* dy; };
struct HeadArgs { HeadBr br[2]; };

__global__ void __launch_bounds__(256) head_fwd_kernel(const __grid_constant__ HeadArgs a, int B, int T, int n_pos, int f16) {
  const bf16* __restrict__ x = a.br[blockIdx.y].x;
  const float* __restrict__ w = a.br[blockIdx.y].w;
  const float* __restrict__ bias = a.br[blockIdx.y].bias;
  float* __restrict__ emb = a.br[blockIdx.y].emb;
  const int part = threadIdx.x & 7;
  const int pos = blockIdx.x * 32 + (threadIdx.x >> 3);
  float acc = 0.f;
  if (pos < n_pos) {
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)pos * HD_C + part * 16);
    float v[8];
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      unpack8(__ldg(src + hlf), v, f16);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(v[j], __ldg(w + part * 16 + hlf * 8 + j), acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (part == 0 && pos < n_pos) {
    // pos = (g*T + t)*8 + b8
    const int b8 = pos & 7, gt = pos >> 3;
    const int g = gt / T, t = gt - g * T;
    const int b = g * 8 + b8;
    if (b < B) emb[(size_t)b * T + t] = 1.f / (1.f + expf(-(acc + __ldg(bias))));
  }
}

extern "C" int zns_head_fwd_nbr(int n_br, const void* const* x_act, const float* const* w128, const float* const* bias1,
                                float* const* emb, int batch, int T, int x_f16, void* stream) {
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(x_act && w128 && bias1 && emb, "NULL argument");
  HeadArgs a;
  memset(&a, 0, sizeof(a));
  for (int b = 0; b < n_br; ++b) {
    ZNS_REQUIRE(x_act[b] && w128[b] && bias1[b] && emb[b], "NULL tensor for branch %d", b);
    a.br[b].x = (const bf16*)x_act[b]; a.br[b].w = w128[b]; a.br[b].bias = bias1[b]; a.br[b].emb = emb[b];
  }
  const int n_pos = zns_groups(batch) * T * 8;
  head_fwd_kernel<<<dim3((n_pos + 31) / 32, n_br), 256, 0, (cudaStream_t)stream>>>(a, batch, T, n_pos, x_f16);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

extern "C" int zns_head_fwd(const void* x_act, const float* w128, const float* bias1, float* emb, int batch, int T,
                            int x_f16, void* stream) {
  return zns_head_fwd_nbr(1, &x_act, &w128, &bias1, &emb, batch, T, x_f16, stream);
}

__global__ void __launch_bounds__(256) head_bwd_kernel(const __grid_constant__ HeadArgs a, int B, int T, int n_pos, float out_scale, int x_f16) {
  const bf16* __restrict__ x = a.br[blockIdx.y].x;
  const float* __restrict__ emb = a.br[blockIdx.y].emb;
  const float* __restrict__ d_emb = a.br[blockIdx.y].d_emb;
  const float* __restrict__ w = a.br[blockIdx.y].w;
  float* __restrict__ dw = a.br[blockIdx.y].dw;
  float* __restrict__ dbias = a.br[blockIdx.y].dbias;
  bf16* __restrict__ dy = a.br[blockIdx.y].dy;
  __shared__ float sdw[HD_C];
  __shared__ float sdb;
  if (threadIdx.x < HD_C) sdw[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) sdb = 0.f;
  __syncthreads();
  const int part = threadIdx.x & 7;
  const int pos = blockIdx.x * 32 + (threadIdx.x >> 3);
  if (pos < n_pos) {
    const int b8 = pos & 7, gt = pos >> 3;
    const int g = gt / T, t = gt - g * T;
    const int b = g * 8 + b8;
    float dz = 0.f;
    if (b < B) {
      const float e = emb[(size_t)b * T + t];
      dz = d_emb[(size_t)b * T + t] * e * (1.f - e);
    }
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)pos * HD_C + part * 16);
    uint4* dst = reinterpret_cast<uint4*>(dy + (size_t)pos * HD_C + part * 16);
    float v[8], o[8];
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      unpack8(__ldg(src + hlf), v, x_f16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = part * 16 + hlf * 8 + j;
        o[j] = v[j] > 0.f ? dz * __ldg(w + c) * out_scale : 0.f;
        if (dz != 0.f) atomicAdd(&sdw[c], dz * v[j]);
      }
      dst[hlf] = pack8(o);
    }
    if (part == 0 && dz != 0.f) atomicAdd(&sdb, dz);
  }
  __syncthreads();
  if (threadIdx.x < HD_C) atomicAdd(dw + threadIdx.x, sdw[threadIdx.x]);
  if (threadIdx.x == 0) atomicAdd(dbias, sdb);
}

extern "C" int zns_head_bwd_nbr(int n_br, const void* const* x_act, const float* const* emb, const float* const* d_emb,
                                const float* const* w128, float* const* dw128, float* const* dbias1, void* const* dy_act,
                                int batch, int T, float out_scale, int x_f16, void* stream) {
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(x_act && emb && d_emb && w128 && dw128 && dbias1 && dy_act, "NULL argument");
  HeadArgs a;
  memset(&a, 0, sizeof(a));
  for (int b = 0; b < n_br; ++b) {
    ZNS_REQUIRE(x_act[b] && emb[b] && d_emb[b] && w128[b] && dw128[b] && dbias1[b] && dy_act[b], "NULL tensor for branch %d", b);
    a.br[b].x = (const bf16*)x_act[b]; a.br[b].emb = const_cast<float*>(emb[b]); a.br[b].d_emb = d_emb[b]; a.br[b].w = w128[b];
    a.br[b].dw = dw128[b]; a.br[b].dbias = dbias1[b]; a.br[b].dy = (bf16*)dy_act[b];
  }
  const int n_pos = zns_groups(batch) * T * 8;
  head_bwd_kernel<<<dim3((n_pos + 31) / 32, n_br), 256, 0, (cudaStream_t)stream>>>(a, batch, T, n_pos, out_scale, x_f16);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

extern "C" int zns_head_bwd(const void* x_act, const float* emb, const float* d_emb, const float* w128, float* dw128,
                            float* dbias1, void* dy_act, int batch, int T, float out_scale, int x_f16, void* stream) {
  return zns_head_bwd_nbr(1, &x_act, &emb, &d_emb, &w128, &dw128, &dbias1, &dy_act, batch, T, out_scale, x_f16, stream);
}

// ---------------------------------------------------------------------------------------------
// Down_CNN merge
// ---------------------------------------------------------------------------------------------
__global__ void merge_kernel(const float* a, const float* b, float* out, long long n, int mode) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = mode == 0 ? fmaxf(a[i], b[i]) : (a[i] + b[i]) / 2.f;
}

extern "C" int zns_merge(const float* a, const float* b, float* out, long long n, int mode, void* stream) {
  ZNS_REQUIRE(a && b && out && n >= 0, "NULL argument");
  ZNS_REQUIRE(mode == 0 || mode == 1, "merge mode must be 0 (max) or 1 (mean)");
  if (n == 0) return ZNS_OK;
  merge_kernel<<<(int)std::min<long long>((n + 255) / 256, 1184), 256, 0, (cudaStream_t)stream>>>(a, b, out, n, mode);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

// ---------------------------------------------------------------------------------------------
// layout converters
// ---------------------------------------------------------------------------------------------
__global__ void act_from_nchw_kernel(const float* __restrict__ x, bf16* __restrict__ act, int B, int C, int H, int W,
                                     size_t total, int f16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int c = (int)(r % C); r /= C;
    const int b8 = (int)(r % 8); r /= 8;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int b = (int)r * 8 + b8;
    const float v = b < B ? x[(((size_t)b * C + c) * H + h) * W + w] : 0.f;
    if (f16) reinterpret_cast<__half*>(act)[i] = __float2half_rn(v);
    else act[i] = __float2bfloat16(v);
  }
}
__global__ void act_to_nchw_kernel(const bf16* __restrict__ act, float* __restrict__ x, int B, int C, int H, int W,
                                   size_t total, int f16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int c = (int)(r % C); r /= C;
    const int b = (int)r;
    const size_t e = zns_act_index(b / 8, h, w, b % 8, c, H, W, C);
    x[i] = f16 ? __half2float(reinterpret_cast<const __half*>(act)[e]) : __bfloat162float(act[e]);
  }
}

extern "C" int zns_act_from_nchw(const float* x, void* act, int batch, int C, int H, int W, int act_f16, void* stream) {
  ZNS_REQUIRE(x && act, "NULL argument");
  const size_t total = (size_t)zns_groups(batch) * H * W * 8 * C;
  act_from_nchw_kernel<<<(int)std::min<size_t>((total + 255) / 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(
      x, (bf16*)act, batch, C, H, W, total, act_f16);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}
extern "C" int zns_act_to_nchw(const void* act, float* x, int batch, int C, int H, int W, int act_f16, void* stream) {
  ZNS_REQUIRE(x && act, "NULL argument");
  const size_t total = (size_t)batch * C * H * W;
  act_to_nchw_kernel<<<(int)std::min<size_t>((total + 255) / 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)act, x, batch, C, H, W, total, act_f16);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 [co][ci][tap] -> bf16 wf [tap][co][ci] and wd [ntaps-1-tap][ci][co]
// (each through a shared-memory tile so that both sides stay coalesced)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_wf_block(int bx, int by, const float* __restrict__ w, bf16* __restrict__ wf, int c_out,
                                              int c_in, int ntaps, int f16) {
  extern __shared__ float tile[];  // [64 ci][ntaps]
  const int co = by, ci0 = bx * 64;
  const float* src = w + ((size_t)co * c_in + ci0) * ntaps;
  const int n = 64 * ntaps;
  for (int i = threadIdx.x; i < n; i += 256) tile[i] = __ldg(src + i);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 256) {
    const int t = i / 64, ci = i - t * 64;
    const size_t o = ((size_t)t * c_out + co) * c_in + ci0 + ci;
    if (f16) reinterpret_cast<__half*>(wf)[o] = __float2half_rn(tile[ci * ntaps + t]);
    else wf[o] = __float2bfloat16(tile[ci * ntaps + t]);
  }
}
__device__ __forceinline__ void pack_wd_block(int bx, int by, const float* __restrict__ w, bf16* __restrict__ wd, int c_out,
                                              int c_in, int ntaps) {
  extern __shared__ float tile[];  // [64 co][ntaps]
  const int ci = by, co0 = bx * 64;
  const int n = 64 * ntaps;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int co = i / ntaps, t = i - co * ntaps;
    tile[i] = __ldg(w + ((size_t)(co0 + co) * c_in + ci) * ntaps + t);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 256) {
    const int t = i / 64, co = i - t * 64;
    wd[((size_t)(ntaps - 1 - t) * c_in + ci) * c_out + co0 + co] = __float2bfloat16(tile[co * ntaps + t]);
  }
}

__device__ __forceinline__ void unpack_grads_block(int bx, int by, float* __restrict__ gpk, float* __restrict__ g, int c_out,
                                                   int c_in, int ntaps, float scale, int accumulate, int zero_src) {
  extern __shared__ float tile[];  // [64 ci][ntaps]
  const int co = by, ci0 = bx * 64;
  const int n = 64 * ntaps;
  // loads and the clearing stores in separate loops: a store to *src inside the load loop orders every later load behind
  // it (same pointer), which leaves ONE load in flight per thread (measured: 1.15 ms instead of ~0.1 ms for all layers)
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += 256) {
    const int t = i / 64, ci = i - t * 64;
    tile[ci * ntaps + t] = gpk[((size_t)t * c_out + co) * c_in + ci0 + ci];
  }
  if (zero_src) {                      // the packed accumulator is clean for the next step's atomics (no memset launch)
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += 256) {
      const int t = i / 64, ci = i - t * 64;
      gpk[((size_t)t * c_out + co) * c_in + ci0 + ci] = 0.f;
    }
  }
  __syncthreads();
  float* dst = g + ((size_t)co * c_in + ci0) * ntaps;
  for (int i = threadIdx.x; i < n; i += 256) dst[i] = (accumulate ? dst[i] : 0.f) + scale * tile[i];
}

// Multi-tensor launches: one kernel packs (or unpacks) every convolution weight of both encoders.  A job is one
// (tensor, direction); block b of the grid finds its job by scanning the (<= 32 entry) table in the kernel parameters.
#define ZNS_MT_MAX_JOBS 32
struct MtJob { const float* src; void* dst; int c_out, c_in, ntaps, kind, blk0, nbx; };   // kind 0 forward pack, 1 flipped pack, 2 unpack
struct MtArgs { MtJob job[ZNS_MT_MAX_JOBS]; int n_jobs, f16, accumulate, zero_src; float scale; };

__global__ void __launch_bounds__(256) multi_tensor_kernel(const __grid_constant__ MtArgs a) {
  // (the table stays in the constant bank: __grid_constant__ -- a by-value parameter indexed with a run-time subscript is
  // copied to local memory by every thread, which made this kernel 60x slower than its work)
  const int b = blockIdx.x;
  int j = 0;
  while (j + 1 < a.n_jobs && b >= a.job[j + 1].blk0) ++j;
  const MtJob& J = a.job[j];
  const int lb = b - J.blk0, bx = lb % J.nbx, by = lb / J.nbx;
  if (J.kind == 0) pack_wf_block(bx, by, J.src, (bf16*)J.dst, J.c_out, J.c_in, J.ntaps, a.f16);
  else if (J.kind == 1) pack_wd_block(bx, by, J.src, (bf16*)J.dst, J.c_out, J.c_in, J.ntaps);
  else unpack_grads_block(bx, by, const_cast<float*>(J.src), (float*)J.dst, J.c_out, J.c_in, J.ntaps, a.scale, a.accumulate, a.zero_src);
}

static int launch_multi(MtArgs& a, int total_blocks, int max_taps, cudaStream_t st) {
  const size_t smem = (size_t)64 * max_taps * sizeof(float);
  ZNS_REQUIRE(smem <= 48 * 1024, "filter with %d taps not supported", max_taps);
  if (total_blocks == 0) return ZNS_OK;
  multi_tensor_kernel<<<total_blocks, 256, smem, st>>>(a);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

extern "C" int zns_pack_weights_multi(int n, const float* const* w, const int* c_out, const int* c_in, const int* kh, const int* kw,
                                      void* const* wf, void* const* wd, int wf_f16, void* stream) {
  ZNS_REQUIRE(n >= 0 && 2 * n <= ZNS_MT_MAX_JOBS, "at most %d tensors per call", ZNS_MT_MAX_JOBS / 2);
  ZNS_REQUIRE(n == 0 || (w && c_out && c_in && kh && kw), "NULL argument");
  MtArgs a;
  memset(&a, 0, sizeof(a));
  a.f16 = wf_f16;
  int blk = 0, max_taps = 1;
  for (int i = 0; i < n; ++i) {
    ZNS_REQUIRE(w[i], "NULL weight %d", i);
    ZNS_REQUIRE(c_out[i] % 64 == 0 && c_in[i] % 64 == 0, "channels must be multiples of 64 (got %d, %d)", c_out[i], c_in[i]);
    const int ntaps = kh[i] * kw[i];
    max_taps = std::max(max_taps, ntaps);
    if (wf && wf[i]) {
      a.job[a.n_jobs++] = MtJob{w[i], wf[i], c_out[i], c_in[i], ntaps, 0, blk, c_in[i] / 64};
      blk += (c_in[i] / 64) * c_out[i];
    }
    if (wd && wd[i]) {
      a.job[a.n_jobs++] = MtJob{w[i], wd[i], c_out[i], c_in[i], ntaps, 1, blk, c_out[i] / 64};
      blk += (c_out[i] / 64) * c_in[i];
    }
  }
  return launch_multi(a, blk, max_taps, (cudaStream_t)stream);
}

extern "C" int zns_pack_weights(const float* w, int c_out, int c_in, int kh, int kw, void* wf, void* wd, int wf_f16,
                                void* stream) {
  ZNS_REQUIRE(w, "NULL weight");
  return zns_pack_weights_multi(1, &w, &c_out, &c_in, &kh, &kw, &wf, &wd, wf_f16, stream);
}

extern "C" int zns_unpack_grads_multi(int n, float* const* gpk, const int* c_out, const int* c_in, const int* kh, const int* kw,
                                      float scale, int accumulate, int zero_packed, float* const* g, void* stream) {
  ZNS_REQUIRE(n >= 0 && n <= ZNS_MT_MAX_JOBS, "at most %d tensors per call", ZNS_MT_MAX_JOBS);
  ZNS_REQUIRE(n == 0 || (gpk && g && c_out && c_in && kh && kw), "NULL argument");
  MtArgs a;
  memset(&a, 0, sizeof(a));
  a.scale = scale; a.accumulate = accumulate; a.zero_src = zero_packed;
  int blk = 0, max_taps = 1;
  for (int i = 0; i < n; ++i) {
    ZNS_REQUIRE(gpk[i] && g[i], "NULL tensor %d", i);
    ZNS_REQUIRE(c_out[i] >= 1 && c_in[i] % 64 == 0, "c_in must be a multiple of 64");
    const int ntaps = kh[i] * kw[i];
    max_taps = std::max(max_taps, ntaps);
    a.job[a.n_jobs++] = MtJob{gpk[i], g[i], c_out[i], c_in[i], ntaps, 2, blk, c_in[i] / 64};
    blk += (c_in[i] / 64) * c_out[i];
  }
  return launch_multi(a, blk, max_taps, (cudaStream_t)stream);
}

extern "C" int zns_unpack_grads(const float* gpk, int c_out, int c_in, int kh, int kw, float scale, int accumulate,
                                float* g, void* stream) {
  ZNS_REQUIRE(gpk && g, "NULL argument");
  float* src = const_cast<float*>(gpk);
  return zns_unpack_grads_multi(1, &src, &c_out, &c_in, &kh, &kw, scale, accumulate, 0, &g, stream);
}

// Zero `bytes` bytes on the stream (cudaMemsetAsync: a memset node in a captured graph, not a kernel launch).
extern "C" int zns_zero(void* p, long long bytes, void* stream) {
  ZNS_REQUIRE(p && bytes >= 0, "bad argument");
  ZNS_CHECK_CUDA(cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream));
  return ZNS_OK;
}

// ---------------------------------------------------------------------------------------------
// bias gradient: db[c] += sum_p dy[p][c]
// ---------------------------------------------------------------------------------------------
struct BiasBr { const uint4* dy; float* db; };
struct BiasArgs { BiasBr br[2]; };

__global__ void __launch_bounds__(256) bias_grad_kernel(const __grid_constant__ BiasArgs a, size_t n_pos, int C) {
  const uint4* __restrict__ dy = a.br[blockIdx.y].dy;
  float* __restrict__ db = a.br[blockIdx.y].db;
  // a thread owns 8 consecutive channels (one 16-byte load per position); C/8 threads span a position
  __shared__ float red[256 * 8];
  const int tpr = C / 8;                     // threads per position: 8, 16 or 32
  const int cg = threadIdx.x % tpr;          // channel group
  const int lane_row = threadIdx.x / tpr;
  const int rows_per_it = 256 / tpr;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  // four independent 16-byte loads per thread in flight (one per trip left the kernel latency bound: 19.7 us per launch on
  // average for 374 MB per step over seven launches)
  const size_t p_step = (size_t)gridDim.x * rows_per_it;
  size_t p = (size_t)blockIdx.x * rows_per_it + lane_row;
  for (; p + 3 * p_step < n_pos; p += 4 * p_step) {
    uint4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = __ldg(dy + (p + u * p_step) * tpr + cg);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float v[8];
      unpack8(q[u], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
  for (; p < n_pos; p += p_step) {
    float v[8];
    unpack8(__ldg(dy + p * tpr + cg), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < C) {
    const int g = threadIdx.x / 8, j = threadIdx.x % 8;
    float s = 0.f;
    for (int r = 0; r < rows_per_it; ++r) s += red[(r * tpr + g) * 8 + j];
    atomicAdd(db + threadIdx.x, s);
  }
}

extern "C" int zns_bias_grad_nbr(int n_br, const void* const* dy_act, int batch, int H, int W, int C, float* const* db,
                                 void* stream) {
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(dy_act && db, "NULL argument");
  ZNS_REQUIRE(C == 64 || C == 128 || C == 256, "bias_grad supports C in {64,128,256}");
  BiasArgs a;
  memset(&a, 0, sizeof(a));
  for (int b = 0; b < n_br; ++b) {
    ZNS_REQUIRE(dy_act[b] && db[b], "NULL tensor for branch %d", b);
    a.br[b].dy = (const uint4*)dy_act[b]; a.br[b].db = db[b];
  }
  const size_t n_pos = (size_t)zns_groups(batch) * H * W * 8;
  const int rows_per_it = 256 / (C / 8);
  const int blocks = (int)std::min<size_t>((n_pos + (size_t)rows_per_it * 8 - 1) / ((size_t)rows_per_it * 8), 148 * 8);
  bias_grad_kernel<<<dim3(std::max(blocks, 1), n_br), 256, 0, (cudaStream_t)stream>>>(a, n_pos, C);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

extern "C" int zns_bias_grad(const void* dy_act, int batch, int H, int W, int C, float* db, void* stream) {
  return zns_bias_grad_nbr(1, &dy_act, batch, H, W, C, &db, stream);
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults: no weight decay, no amsgrad), flat fp32 buffers
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                                   float b1, float b2, float eps, float inv_bc1, float inv_sqrt_bc2,
                                                   const uint32_t* __restrict__ step_dev, float gscale) {
  if (step_dev) {
    const double st = (double)__ldg(step_dev);
    inv_bc1 = (float)(1.0 / (1.0 - pow((double)b1, st)));
    inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, st)));
  }
  const long long n4 = n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define ZNS_ADAM1(f)                                                        \
  {                                                                         \
    const float gr = gg.f * gscale;                                         \
    mm.f = b1 * mm.f + (1.f - b1) * gr;                                     \
    vv.f = b2 * vv.f + (1.f - b2) * gr * gr;                                \
    const float denom = sqrtf(vv.f) * inv_sqrt_bc2 + eps;                   \
    pp.f -= (lr * inv_bc1) * (mm.f / denom);                                \
  }
    ZNS_ADAM1(x) ZNS_ADAM1(y) ZNS_ADAM1(z) ZNS_ADAM1(w)
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gr = g[i] * gscale;
    const float mm = b1 * m[i] + (1.f - b1) * gr;
    const float vv = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mm;
    v[i] = vv;
    p[i] -= (lr * inv_bc1) * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
  }
}

extern "C" int zns_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                             float beta2, float eps, int step, const uint32_t* step_dev, float grad_scale, void* stream) {
  ZNS_REQUIRE(p && g && m && v, "NULL argument");
  ZNS_REQUIRE((step >= 1 || step_dev) && n >= 0, "Adam step counts from 1");
  if (step < 1) step = 1;
  ZNS_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, "Adam buffers must be 16-byte aligned");
  if (n == 0) return ZNS_OK;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  const int blocks = (int)std::min<long long>((n / 4 + 255) / 256 + 1, 148 * 8);
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, (float)(1.0 / bc1),
                                                        (float)(1.0 / sqrt(bc2)), step_dev, grad_scale);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

__global__ void counter_add_kernel(uint32_t* ctr, uint32_t inc) { *ctr += inc; }

extern "C" int zns_counter_add(uint32_t* ctr, uint32_t inc, void* stream) {
  ZNS_REQUIRE(ctr, "NULL counter");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctr, inc);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused gradient reduce-scatter + Adam + parameter all-gather over NVLink peer memory.
//   Every rank owns one contiguous shard of the flat parameter vector.  For its shard it reads the
//   gradient of EVERY rank straight from that rank's HBM (peer loads over NVLink/NVSwitch), sums
//   them in rank order (the same order on every owner -> replicas stay bit-identical), applies
//   Adam (moments exist only for the owned shard) and stores the new parameters into every rank's
//   parameter buffer (peer stores).  One launch replaces all-reduce + optimizer: per rank
//   2 (W-1)/W of the parameter bytes cross NVLink instead of the ring's 2 (W-1)/W * 2, and the
//   transfers overlap the arithmetic element by element.  The caller brackets the launch with
//   cross-rank barriers (gradients complete before, parameters visible after).
// ---------------------------------------------------------------------------------------------
#define P2P_MAX_WORLD 16

struct P2PPtrs {
  const float* g[P2P_MAX_WORLD];
  float* p[P2P_MAX_WORLD];
};

__global__ void __launch_bounds__(256) adam_p2p_kernel(P2PPtrs ptrs, int world, int rank, float* __restrict__ m,
                                                       float* __restrict__ v, long long lo4, long long hi4, float lr, float b1,
                                                       float b2, float eps, float inv_bc1, float inv_sqrt_bc2,
                                                       const uint32_t* __restrict__ step_dev, float gscale) {
  if (step_dev) {
    const double st = (double)__ldg(step_dev);
    inv_bc1 = (float)(1.0 / (1.0 - pow((double)b1, st)));
    inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, st)));
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride) {
    float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int r = 0; r < world; ++r) {
      const float4 t = __ldcv(reinterpret_cast<const float4*>(ptrs.g[r]) + i);   // peer (or local) HBM, not cached
      gg.x += t.x; gg.y += t.y; gg.z += t.z; gg.w += t.w;
    }
    float4 pp = reinterpret_cast<float4*>(ptrs.p[rank])[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define ZNS_ADAM2(f)                                                        \
  {                                                                         \
    const float gr = gg.f * gscale;                                         \
    mm.f = b1 * mm.f + (1.f - b1) * gr;                                     \
    vv.f = b2 * vv.f + (1.f - b2) * gr * gr;                                \
    const float denom = sqrtf(vv.f) * inv_sqrt_bc2 + eps;                   \
    pp.f -= (lr * inv_bc1) * (mm.f / denom);                                \
  }
    ZNS_ADAM2(x) ZNS_ADAM2(y) ZNS_ADAM2(z) ZNS_ADAM2(w)
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(ptrs.p[r])[i] = pp;   // all-gather by peer stores
  }
}

extern "C" int zns_adam_p2p(int world, int rank, const void* const* g_peers_host, void* const* p_peers_host, float* m, float* v,
                            long long n, float lr, float beta1, float beta2, float eps, int step, const uint32_t* step_dev,
                            void* stream) {
  ZNS_REQUIRE(g_peers_host && p_peers_host && m && v, "NULL argument");
  ZNS_REQUIRE(world >= 1 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world, "bad world/rank");
  ZNS_REQUIRE(n % 4 == 0, "flat buffers must hold a multiple of 4 elements");
  ZNS_REQUIRE(step >= 1 || step_dev, "Adam step counts from 1");
  if (step < 1) step = 1;
  P2PPtrs ptrs;
  for (int r = 0; r < world; ++r) {
    ZNS_REQUIRE(g_peers_host[r] && p_peers_host[r], "NULL peer pointer for rank %d", r);
    ptrs.g[r] = (const float*)g_peers_host[r];
    ptrs.p[r] = (float*)p_peers_host[r];
  }
  const long long n4 = n / 4;
  const long long per = (n4 + world - 1) / world;
  const long long lo4 = std::min<long long>(n4, per * rank), hi4 = std::min<long long>(n4, per * (rank + 1));
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  if (hi4 > lo4) {
    const int blocks = (int)std::min<long long>((hi4 - lo4 + 255) / 256, 148 * 8);
    adam_p2p_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ptrs, world, rank, m, v, lo4, hi4, lr, beta1, beta2, eps,
                                                              (float)(1.0 / bc1), (float)(1.0 / sqrt(bc2)), step_dev,
                                                              1.0f / (float)world);
    ZNS_CHECK_LAUNCH();
  }
  return ZNS_OK;
}
