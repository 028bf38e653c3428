// NT-Xent forward + backward in one launch (one CTA; the whole problem lives in shared memory).
//
// Replaces NTXent.forward (/root/reference/zeroNoteSamba/models/loss_functions.py:24-55), which
// runs ~20 tiny kernels and 3 host syncs per row, and its autograd backward.  Closed form
// (SURVEY.md appendix C): a^ = a/max(|a|,1e-8), p^ likewise, S = A^ P^T, Z = S/tau,
// loss = (1/batch_len) sum_i (logsumexp_j Z_ij - Z_ii); with G = (softmax(Z) - I)/(tau batch_len):
// dL/dA^ = G P^, dL/dP^ = G^T A^, and through the normalisation
// dL/da_i = (dL/da^_i - (dL/da^_i . a^_i) a^_i) / max(|a_i|,1e-8).
#include "common.cuh"

#define NT_THREADS 256
#define NT_MAX_ROWS 64

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(NT_THREADS)
ntxent_kernel(const float* __restrict__ anchors, const float* __restrict__ poss, int n, int D, int batch_len, float tau,
              float* __restrict__ result3, float* __restrict__ d_anchors, float* __restrict__ d_poss) {
  extern __shared__ float sm[];
  float* A = sm;                    // [n][D] normalised anchors
  float* P = A + (size_t)n * D;     // [n][D] normalised positives
  float* dA = P + (size_t)n * D;    // [n][D]
  float* dP = dA + (size_t)n * D;   // [n][D]
  float* S = dP + (size_t)n * D;    // [n][n] cosine, then G
  float* inv_na = S + n * n;        // [n]
  float* inv_np = inv_na + n;       // [n]
  float* row_loss = inv_np + n;     // [n]
  float* row_pos = row_loss + n;    // [n]
  float* row_neg = row_pos + n;     // [n]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT_THREADS / 32;

  for (int i = tid; i < n * D; i += NT_THREADS) {
    A[i] = anchors[i];
    P[i] = poss[i];
  }
  __syncthreads();
  for (int r = warp; r < 2 * n; r += nwarps) {
    const float* row = (r < n) ? A + (size_t)r * D : P + (size_t)(r - n) * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(row[d], row[d], s);
    s = warp_sum(s);
    if (lane == 0) (r < n ? inv_na[r] : inv_np[r - n]) = 1.f / fmaxf(sqrtf(s), 1e-8f);
  }
  __syncthreads();
  for (int i = tid; i < n * D; i += NT_THREADS) {
    const int r = i / D;
    A[i] *= inv_na[r];
    P[i] *= inv_np[r];
  }
  __syncthreads();
  for (int pr = warp; pr < n * n; pr += nwarps) {
    const int i = pr / n, j = pr - i * n;
    const float* a = A + (size_t)i * D;
    const float* p = P + (size_t)j * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(a[d], p[d], s);
    s = warp_sum(s);
    if (lane == 0) S[pr] = s;
  }
  __syncthreads();
  // row-wise log-softmax; S is overwritten by G
  for (int i = warp; i < n; i += nwarps) {
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, S[i * n + j] / tau);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f, ssum = 0.f;
    for (int j = lane; j < n; j += 32) {
      se += expf(S[i * n + j] / tau - mx);
      ssum += S[i * n + j];
    }
    se = warp_sum(se);
    ssum = warp_sum(ssum);
    const float sii = S[i * n + i];
    const float lse = mx + logf(se);
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
      const float w = expf(S[i * n + j] / tau - lse);
      S[i * n + j] = (w - (j == i ? 1.f : 0.f)) / (tau * (float)batch_len);
    }
    if (lane == 0) {
      row_loss[i] = lse - sii / tau;
      row_pos[i] = sii;
      row_neg[i] = (ssum - sii) / (float)(batch_len - 1);
    }
  }
  __syncthreads();
  if (tid == 0) {
    float l = 0.f, cp = 0.f, cn = 0.f;
    for (int i = 0; i < n; ++i) {
      l += row_loss[i];
      cp += row_pos[i];
      cn += row_neg[i];
    }
    result3[0] = l / (float)batch_len;
    result3[1] = cp / (float)batch_len;
    result3[2] = cn / (float)batch_len;
  }
  if (d_anchors == nullptr && d_poss == nullptr) return;
  // dA^ = G P^ ; dP^ = G^T A^
  for (int e = tid; e < n * D; e += NT_THREADS) {
    const int r = e / D, d = e - r * D;
    float ga = 0.f, gp = 0.f;
    for (int j = 0; j < n; ++j) {
      ga = fmaf(S[r * n + j], P[(size_t)j * D + d], ga);
      gp = fmaf(S[j * n + r], A[(size_t)j * D + d], gp);
    }
    dA[e] = ga;
    dP[e] = gp;
  }
  __syncthreads();
  // projection through the normalisation; row_loss / row_pos are reused for the dot products
  for (int r = warp; r < 2 * n; r += nwarps) {
    const float* gr = (r < n) ? dA + (size_t)r * D : dP + (size_t)(r - n) * D;
    const float* xr = (r < n) ? A + (size_t)r * D : P + (size_t)(r - n) * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(gr[d], xr[d], s);
    s = warp_sum(s);
    if (lane == 0) (r < n ? row_loss[r] : row_pos[r - n]) = s;
  }
  __syncthreads();
  for (int e = tid; e < n * D; e += NT_THREADS) {
    const int r = e / D;
    if (d_anchors) d_anchors[e] = (dA[e] - row_loss[r] * A[e]) * inv_na[r];
    if (d_poss) d_poss[e] = (dP[e] - row_pos[r] * P[e]) * inv_np[r];
  }
}

extern "C" int zns_ntxent_fwd_bwd(const float* anchors, const float* poss, int n_rows, int dim, int batch_len,
                                  float temperature, float* result3, float* d_anchors, float* d_poss, void* stream) {
  ZNS_REQUIRE(anchors && poss && result3, "NULL argument");
  ZNS_REQUIRE(n_rows >= 1 && n_rows <= NT_MAX_ROWS && dim >= 1, "NT-Xent supports 1..%d rows", NT_MAX_ROWS);
  ZNS_REQUIRE(batch_len >= 2 && n_rows <= batch_len, "batch_len must be >= 2 and >= rows (got %d rows, batch_len %d)",
              n_rows, batch_len);
  ZNS_REQUIRE(temperature > 0.f, "temperature must be positive");
  const size_t smem = ((size_t)4 * n_rows * dim + (size_t)n_rows * n_rows + 5 * n_rows) * sizeof(float);
  ZNS_REQUIRE(smem <= 200 * 1024, "NT-Xent problem %d x %d does not fit shared memory", n_rows, dim);
  static unsigned long long attr_devs = 0;          // one bit per device ordinal: function attributes are per device
  if (zns_first_use_on_device(&attr_devs)) {
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(ntxent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  ntxent_kernel<<<1, NT_THREADS, smem, (cudaStream_t)stream>>>(anchors, poss, n_rows, dim, batch_len, temperature,
                                                               result3, d_anchors, d_poss);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}


// ---------------------------------------------------------------------------------------------
// Binary cross entropy of the downstream beat head (torch.nn.BCELoss, mean reduction; loader.py:20, epochs.py:52-54),
// forward and backward in ONE launch: loss = -mean(t log o + (1 - t) log(1 - o)) with torch's clamps (log >= -100;
// gradient denominator >= 1e-12), d_out = (o - t) / max(o (1 - o), 1e-12) / n.  One CTA: n is a clip's frame count.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bce_kernel(const float* __restrict__ o, const float* __restrict__ t, long long n,
                                                  float* __restrict__ loss, float* __restrict__ d_out) {
  __shared__ float red[8];
  const float inv_n = 1.f / (float)n;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const float x = o[i], y = t[i];
    acc -= y * fmaxf(logf(x), -100.f) + (1.f - y) * fmaxf(log1pf(-x), -100.f);
    if (d_out) d_out[i] = (x - y) / fmaxf(x * (1.f - x), 1e-12f) * inv_n;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int w = 0; w < 8; ++w) sum += red[w];
    *loss = sum * inv_n;
  }
}

extern "C" int zns_bce_fwd_bwd(const float* out, const float* target, long long n, float* loss, float* d_out, void* stream) {
  ZNS_REQUIRE(out && target && loss && n > 0, "bad argument");
  bce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(out, target, n, loss, d_out);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}
