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
//  * one persistent CTA per SM, warp specialised: 8 loader warps (cp.async into a ring of plane-group slots, fp32 ->
//    two-term fp16 conversion in place for level 0), 4 MMA-issuing warps (the tiles' MMAs are many and small -- N <= 48,
//    ~55 clocks each -- so a single issuing thread, at ~190 clocks of scalar work per MMA, is the bottleneck; the
//    accumulator units are dealt to four threads, each owning its units' TMEM columns), 8 epilogue warps; TMEM
//    accumulator rings with two stages so that the epilogue of a tile overlaps the MMAs of the next.
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
#define VQT_ZERO_BYTES 128

struct PendingMma { VqtMma m; int group; int issuer; };

static int build_level(zns_vqt_plan* p, int lvl, double fmin, double gamma_in, const double* taps) {
  VqtLevelDev& L = p->level[lvl];
  memset(&L, 0, sizeof(L));
  const int hop = p->hop >> lvl;
  const int nf = p->n_fft[lvl];
  if (hop < 2 || nf > 256 || nf < 16 || p->bpo != 12) return 1;
  const int R = hop >= 32 ? hop : (hop >= 8 ? 32 : 8);
  const int q = R / 8, fpr = R / hop;
  if (!(q == 1 || q == 4 || q == 8 || q == 16 || q == 32) || fpr > 4) return 1;
  const bool last = (lvl == p->n_oct - 1);
  L.q = q; L.fpr = fpr; L.hop = hop; L.n_fft = nf; L.bin0 = p->n_bins - p->bpo * (lvl + 1);
  L.dec_w = last ? 0 : R / 2;
  L.n_pass = last ? 0 : (L.dec_w + 63) / 64;
  L.dec_wp = last ? 0 : ((std::min(L.dec_w, 64) + 15) / 16) * 16;
  L.dec_scale = (float)(sqrt(2.0) / DEC_TAP_SCALE);
  L.fb_scale = p->coef_inv_scale[lvl];
  L.pg = (q == 1) ? 1 : (q >= 8 ? 8 : 4);     // bigger groups = fewer pipeline hand-offs per tile
  L.gpt = q / L.pg;
  if (L.gpt > ZNS_VQT_MAX_GROUPS) return 1;

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
  L.slot_term_bytes = L.pg * rtot * 16;
  // group (slot) and byte offset inside the slot's term half of the k-step at x0
  auto a_group = [&](int x0) -> int {
    if (q == 1) return 0;
    const int rs = floordiv(x0, R);
    return ((x0 - rs * R) / 8) / L.pg;
  };
  auto a_off = [&](int x0) -> uint32_t {
    if (q == 1) return (uint32_t)(16 * (floordiv(x0, 8) + L.hb));
    const int rs = floordiv(x0, R), p0 = (x0 - rs * R) / 8;
    return (uint32_t)((p0 % L.pg) * L.a_lbo + 16 * (rs + L.hb));
  };

  // ---- TMEM rings: [0] filterbank, [1] decimator ----
  // Filterbank accumulator of a stage.  One frame per row: [x1.g1 | x1.g2] (48 columns) and x2.g1 in 24 (padded to 32) columns
  // of its own, two accumulator units issued by different threads.  Several frames per row ("merged"): columns
  // [g1 of frame 0 .. fpr-1 | g2 of frame 0 .. fpr-1]; x1 . [g1; g2] is one MMA over all 48 fpr columns, x2 . g1 a second one
  // that ACCUMULATES ONTO the g2 columns (both carry the 1/2048 scale) and reads the first 24 fpr rows of the same coefficient
  // tile -- 48 fpr instead of 72 fpr columns per stage, so the four-frame levels get two stages too.
  const bool merged = fpr >= 2;
  const int n1 = fpr * 48, n2 = merged ? 0 : std::max(32, fpr * 24);
  L.fb_n1 = n1; L.fb_n2 = n2;
  L.ring_width[0] = n1 + n2;
  L.ring_width[1] = 2 * L.dec_wp;
  L.ring_stages[1] = last ? 0 : 2;
  L.ring_base[1] = 0;
  L.ring_base[0] = L.ring_stages[1] * L.ring_width[1];
  L.ring_stages[0] = (L.ring_base[0] + 2 * L.ring_width[0] <= 512) ? 2 : 1;
  if (L.ring_base[0] + L.ring_stages[0] * L.ring_width[0] > 512) return 1;

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
          if (merged) {
            (*img)[tile_idx(t1, n1, j * 24 + col, k)] = h1;
            (*img)[tile_idx(t1, n1, fpr * 24 + j * 24 + col, k)] = h2;
          } else {
            (*img)[tile_idx(t1, n1, j * 48 + col, k)] = h1;
            (*img)[tile_idx(t1, n1, j * 48 + 24 + col, k)] = h2;
            (*img)[tile_idx(t2, n2, j * 24 + col, k)] = h1;
          }
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

  // ---- MMAs, tagged with the plane group they read and their accumulator unit (job, part) ----
  std::vector<PendingMma> pend;
  auto push = [&](int x0, size_t bo, int n, int col, int term, int brows, int job, int part) {
    PendingMma pm;
    pm.m = VqtMma{a_off(x0), (uint32_t)bo, (uint16_t)n, (uint16_t)col, (uint8_t)term, (uint8_t)job, (uint8_t)part,
                  (uint8_t)(brows / 8)};
    pm.group = a_group(x0);
    pm.issuer = 0;
    pend.push_back(pm);
  };
  for (size_t s = 0; s < fb_x0.size(); ++s) {
    const size_t t1 = s * (fb_tile1 + fb_tile2), t2 = t1 + fb_tile1;
    push(fb_x0[s], t1, n1, 0, 0, n1, 0, 0);
    if (merged) push(fb_x0[s], t1, fpr * 24, fpr * 24, 1, n1, 0, 0);     // same unit: issued behind the x1 MMA by the same thread
    else push(fb_x0[s], t2, n2, n1, 1, n2, 0, 1);
  }
  for (int pass = 0; pass < L.n_pass; ++pass) {
    const int c0 = 64 * pass, c1 = std::min(L.dec_w, c0 + 64);   // output columns of this pass
    for (int x0 : dec_x0) {
      // outputs o = x0/2 - 16 + n; taps are non-zero for n in [1, 39]
      const int o_lo = std::max(c0, x0 / 2 - 15), o_hi = std::min(c1 - 1, x0 / 2 + 23);
      if (o_lo > o_hi) continue;
      int o_start = c0 + ((o_lo - c0) / 8) * 8;
      int n = ((o_hi + 1 - o_start + 15) / 16) * 16;
      if (o_start + n > c0 + L.dec_wp) o_start = c0 + L.dec_wp - n;
      if (o_start < c0) { o_start = c0; n = L.dec_wp; }
      const int n0 = o_start - x0 / 2 + 16;
      if (n0 < DEC_T_N0 || n0 + n > DEC_T_N0 + DEC_T_ROWS) return 1;
      const size_t row_off = (size_t)16 * (n0 - DEC_T_N0);
      const size_t tt1 = dec_base + row_off, tt2 = dec_base + (size_t)DEC_T_ROWS * 32 + row_off;
      const int col = o_start - c0;
      push(x0, tt1, n, col, 0, DEC_T_ROWS, 1 + pass, 0);
      push(x0, tt2, n, L.dec_wp + col, 0, DEC_T_ROWS, 1 + pass, 1);
      push(x0, tt1, n, L.dec_wp + col, 1, DEC_T_ROWS, 1 + pass, 1);
    }
  }
  L.n_jobs = 1 + L.n_pass;

  // ---- deal the accumulator units to the issuer warps (longest first) ----
  const int n_units = 2 * L.n_jobs;
  int weight[2 * ZNS_VQT_MAX_JOBS] = {0}, unit_issuer[2 * ZNS_VQT_MAX_JOBS] = {0}, load[ZNS_VQT_ISSUERS] = {0};
  for (const PendingMma& pm : pend) weight[2 * pm.m.job + pm.m.part] += 1;
  std::vector<int> by_w(n_units);
  for (int u = 0; u < n_units; ++u) by_w[u] = u;
  std::stable_sort(by_w.begin(), by_w.end(), [&](int a, int b) { return weight[a] > weight[b]; });
  for (int u : by_w) {
    int best = 0;
    for (int k = 1; k < ZNS_VQT_ISSUERS; ++k) if (load[k] < load[best]) best = k;
    unit_issuer[u] = best;
    load[best] += weight[u];
  }
  for (PendingMma& pm : pend) pm.issuer = unit_issuer[2 * pm.m.job + pm.m.part];

  // ---- group order: start with the group the filterbank's first (row - 1) k-step reads ----
  const int g_first = a_group(fb_x0[0]);
  for (int i = 0; i < L.gpt; ++i) L.g_order[i] = (g_first + i) % L.gpt;

  // ---- segments: (issuer, position, unit) runs ----
  int m = 0, n_seg = 0;
  int last_seg[2 * ZNS_VQT_MAX_JOBS], last_mma[ZNS_VQT_MAX_JOBS];
  bool started[2 * ZNS_VQT_MAX_JOBS];
  for (int u = 0; u < 2 * ZNS_VQT_MAX_JOBS; ++u) { last_seg[u] = -1; started[u] = false; }
  for (int j = 0; j < ZNS_VQT_MAX_JOBS; ++j) last_mma[j] = -1;
  for (int isr = 0; isr < ZNS_VQT_ISSUERS; ++isr) {
    for (int pos = 0; pos < L.gpt; ++pos) {
      L.seg_begin[isr][pos] = n_seg;
      for (int u = 0; u < n_units; ++u) {
        if (unit_issuer[u] != isr) continue;
        int seg_first = -1, seg_count = 0;
        for (const PendingMma& pm : pend) {
          if (pm.group != L.g_order[pos] || 2 * pm.m.job + pm.m.part != u) continue;
          if (m >= ZNS_VQT_MAX_MMA) return 1;
          if (seg_first < 0) seg_first = m;
          L.mma[m] = pm.m;
          last_mma[pm.m.job] = std::max(last_mma[pm.m.job], pos);
          ++m;
          ++seg_count;
        }
        if (seg_count > 0) {
          if (n_seg >= ZNS_VQT_MAX_SEGS) return 1;
          L.seg[n_seg] = VqtSeg{(uint16_t)seg_first, (uint16_t)seg_count, (uint8_t)(u / 2), (uint8_t)(u % 2),
                                (uint8_t)(started[u] ? 0 : 1), 0};
          started[u] = true;
          last_seg[u] = n_seg;
          ++n_seg;
        }
      }
    }
    L.seg_begin[isr][L.gpt] = n_seg;
  }
  L.n_mma = m;
  L.n_seg = n_seg;
  // Hand-off invariant of the slot barriers: an issuer that works at all must have MMAs at EVERY position of a tile.  One that
  // idles at position p but works at a later one skips ahead inside the tile, and with a one-tile ring its next-tile arrival
  // on slot p could complete the slot's "empty" phase before the others have read it (issuers without any unit are not counted).
  for (int isr = 0; isr < ZNS_VQT_ISSUERS; ++isr) {
    if (L.seg_begin[isr][L.gpt] == L.seg_begin[isr][0]) continue;
    for (int pos = 0; pos < L.gpt; ++pos)
      if (L.seg_begin[isr][pos + 1] == L.seg_begin[isr][pos]) return 1;
  }
  for (int u = 0; u < n_units; ++u) {
    if (last_seg[u] < 0) {
      if (merged && u == 1) continue;         // merged filterbank: one unit, its "full" barrier expects one commit
      return 1;                               // every other unit must appear: the "full" barriers expect two commits per job
    }
    L.seg[last_seg[u]].flags |= 2;
  }
  // epilogue order = completion order (ties: filterbank first)
  int order[ZNS_VQT_MAX_JOBS] = {0, 1, 2};
  std::stable_sort(order, order + L.n_jobs, [&](int a, int b) { return last_mma[a] < last_mma[b]; });
  for (int j = 0; j < L.n_jobs; ++j) L.ep_job[j] = order[j];

  for (int i = 0; i < L.n_mma; ++i) {
    const VqtMma& mm = L.mma[i];
    VqtMmaPacked& pk = L.pk[i];
    pk.a_lo = ((mm.a_off + (uint32_t)mm.term * (uint32_t)L.slot_term_bytes) >> 4) | (((uint32_t)L.a_lbo >> 4) << 16);
    pk.b_lo = (mm.b_off >> 4) | (((128u * mm.b_rows8) >> 4) << 16);
    pk.idesc = umma_idesc_f16(128, mm.n);
    pk.col = (uint32_t)(L.ring_base[mm.job ? 1 : 0] + mm.d_col);
  }

  // ---- shared-memory ring ----
  const size_t slot = 2 * (size_t)L.slot_term_bytes;
  const size_t budget = 200 * 1024 - (size_t)L.b_bytes - VQT_ZERO_BYTES;
  int n_slots = (int)(budget / slot);
  n_slots = n_slots >= 8 ? 8 : (n_slots >= 4 ? 4 : (n_slots >= 2 ? 2 : n_slots));     // 8 loader warps share the slots evenly
  if (n_slots < L.gpt || n_slots % L.gpt != 0) return 1;                            // whole tiles
  L.n_slots = n_slots;
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

static size_t level_smem(const VqtLevelDev& L) {
  return (size_t)L.n_slots * 2 * L.slot_term_bytes + (size_t)L.b_bytes + VQT_ZERO_BYTES + 128;
}

int vqt_umma_build(zns_vqt_plan* p, double fmin, double gamma_in) {
  p->umma_ok = false;
  double taps[32];
  zns_vqt_decimator_taps_host(taps);
  for (int i = 0; i < p->n_oct; ++i) {
    if (build_level(p, i, fmin, gamma_in, taps) != 0 || level_smem(p->level[i]) > 220 * 1024) {
      vqt_umma_free(p);
      return ZNS_OK;   // unsupported geometry: the round-1 kernels stay in charge
    }
  }
  size_t n = (size_t)p->max_samples;
  const size_t f_max = 1 + (size_t)p->max_samples / p->hop;
  for (int i = 0; i < p->n_oct; ++i) {
    ZNS_CHECK_CUDA(cudaMalloc(&p->d_bimg[i], p->level[i].b_bytes));
    ZNS_CHECK_CUDA(cudaMemcpy(p->d_bimg[i], p->h_bimg[i]->data(), p->level[i].b_bytes, cudaMemcpyHostToDevice));
    if (i > 0) {
      n = (n + 1) / 2;
      const VqtLevelDev& L = p->level[i];
      const size_t R = 8 * (size_t)L.q;
      const size_t rows = std::max((f_max + L.fpr - 1) / L.fpr, (n + R - 1) / R);
      const size_t tiles = (rows + 127) / 128;
      p->sig_cap[i] = (long long)(tiles * 128 * R);
      // tiled: (tiles + 1) images of q planes x rtot rows x 8 samples (the extra one takes the "halo before" copy of the last
      // tile's rows); linear (q == 1): the samples behind a 1024-sample zero pad, one pad tile of slack behind
      p->sig_stride[i] = L.q == 1 ? (long long)((tiles + 2) * 1024) : (long long)((tiles + 1) * L.q * L.rtot * 8);
      const size_t bytes = (size_t)p->max_batch * p->sig_stride[i] * sizeof(uint16_t);
      ZNS_CHECK_CUDA(cudaMalloc(&p->d_hi[i], bytes));
      ZNS_CHECK_CUDA(cudaMalloc(&p->d_lo[i], bytes));
      ZNS_CHECK_CUDA(cudaMemset(p->d_hi[i], 0, bytes));
      ZNS_CHECK_CUDA(cudaMemset(p->d_lo[i], 0, bytes));
    }
  }
  p->umma_last_n = 0;
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

__device__ __forceinline__ float log_approx(float x) {        // MUFU.LG2 * ln 2 for x >= 1e-9 (no denormal pre-scaling as in __logf)
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.69314718055994530942f;
}
__device__ __forceinline__ float sqrt_approx(float x) {       // MUFU.SQRT (2 ulp): the magnitude only feeds a log
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// 16-byte asynchronous global -> shared copy; bytes beyond src_bytes are zero filled
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Global layout of a level signal (levels >= 1; written by the previous level's epilogue).  q >= 4 ("tiled"): the buffer holds,
// tile after tile and plane group after plane group, exactly the shared-memory image the loader needs -- per tile t and plane c
// `rtot` rows of 16 bytes: hb halo rows (the last rows of tile t - 1), the 128 rows of the tile, ha halo rows (the first rows of
// tile t + 1), padding up to rtot.  A plane group of a tile is therefore ONE contiguous block per fp16 term and its load is one
// bulk copy (the TMA unit's request rate, not bytes, bounded the earlier layout with ~48 small copies per group).  The
// producer writes the (<= 3 of 128) halo rows twice.  Row rr of plane c of tile t: chunk index (t q + c) rtot + rr.
// q == 1: linear, one 1024-sample zero pad in front.
__host__ __device__ __forceinline__ long long level_index(long long p, int q, int rtot, int hb) {
  if (q == 1) return 1024 + p;
  const int R = 8 * q;
  const long long row = p / R;
  const int c = (int)(p - row * R) >> 3, e = (int)(p & 7);
  const long long t = row >> 7;
  const int r = (int)(row & 127);
  return ((t * q + c) * rtot + hb + r) * 8 + e;
}

// bulk asynchronous copy global -> shared (TMA, no tensor map), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
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

// Warp roles of the persistent level kernel (20 / 21 warps): epilogue = warps 0..7, MMA issuers = warps 8..11 (one per scheduler
// sub-partition), warps 12..20 depend on the source:
//   fp32 source (level 0): warps 12..19 load (global -> registers -> two-term split -> shared memory);
//   levels >= 1: the source is already split and tile ordered, a group is two bulk copies -- warp 12 alone issues them and
//   warps 13..20 are EIGHT MORE EPILOGUE warps (four per TMEM lane quadrant instead of two): the small levels are bound by the
//   per-tile latency chain of the epilogue (tcgen05.ld -> FP -> store), which only more warps shorten.
#define VQT_EPI_WARPS 8          // base epilogue warps (all kernels)
#define VQT_LOAD_WARPS 8
#ifndef VQT_PF_DIST
#define VQT_PF_DIST 1            // level-0 loader: L2 prefetch distance in tiles (2: 0.495 ms, 3: 0.522 ms vs 0.461 at cfg2 -- L2 thrash)
#endif
#ifndef VQT_LD_U
#define VQT_LD_U 6               // level-0 loader: 32-byte chunks per lane in flight
#endif
#define VQT_ISSUE_WARP0 VQT_EPI_WARPS
#define VQT_LOAD_WARP0 (VQT_EPI_WARPS + ZNS_VQT_ISSUERS)
#define VQT_EXTRA_WARP0 (VQT_LOAD_WARP0 + 1)      // levels >= 1: first of the eight extra epilogue warps
// 20 warps at level 0, 21 behind it (a sixth warp on one scheduler sub-partition caps the kernel at 80 registers: level 0 needs 96)
#define VQT_LEVEL_THREADS(src_f32) (32 * (VQT_LOAD_WARP0 + VQT_LOAD_WARPS + ((src_f32) ? 0 : 1)))

struct VqtLevelArgs {
  const float* y32;            // level 0 source [batch][src_stride]
  const uint16_t* src_hi;      // level >= 1 source, two fp16 terms, tiled chunk-major layout (zero beyond n_sig)
  const uint16_t* src_lo;
  int n_sig;
  long long src_stride;
  const uint16_t* bimg;
  const float* inv_sqrt_len;
  float* out;                  // [batch][n_bins][n_frames]
  int n_frames, n_bins;
  uint16_t* dst_hi;            // next level (tiled chunk-major layout, see level_index)
  uint16_t* dst_lo;
  int n_valid;
  long long dst_stride;
  int dst_q;                   // planes per row of the next level (1: linear layout)
  int dst_rtot, dst_hb, dst_ha; // next level: rows per plane image, halo rows before / after (tiled layout, see level_index)
  long long dst_cap;           // samples the next level's buffer holds per clip
  int tiles_per_clip, n_tiles;
  long long* dbg;              // optional per-role cycle counters of CTA 0 (zns_dbg_vqt_timing), else NULL
  int mode;                    // timing build only (ZNS_VQT_DBG_MODE): knock-out bits, see VQT_KO
};

// Cycle counters of the roles (zns_dbg_vqt_timing) exist only in builds with -DZNS_VQT_TIMING: the product kernels carry
// no diagnostic code (instruction-cache footprint).
#ifdef ZNS_VQT_TIMING
#define VQT_CLOCK() clock64()
#define VQT_TIMING_ON true
// knock-out experiments (results are garbage, only the time is of interest): 1 loader skips the fp32 -> fp16 split,
// 2 epilogue skips everything but the handshakes, 4 issuers skip the MMAs, 8 loaders skip the copies, 16 epilogue skips
// the global stores only
#define VQT_KO(bit) ((A.mode & (bit)) != 0)
// timeline of CTA 0 of the fp32-source (level 0) kernel, tiles 4..7: absolute clock64 stamps at dbg[320 + i]
// (loaders: [tile][warp][wait end, arrive]; issuers: [tile][issuer][pos][operands ready, after commit]; epilogue warps 0 / 4:
// [tile][half][job][accumulator ready, arrive])
#define VQT_TL(cond, i) do { if (SRC_F32 && A.dbg && blockIdx.x == 0 && (cond)) A.dbg[320 + (i)] = clock64(); } while (0)
#else
#define VQT_TL(cond, i) do { } while (0)
#define VQT_CLOCK() 0LL
#define VQT_TIMING_ON false
#define VQT_KO(bit) false
#endif

// Wait flavour of the level kernels (A/B at build time, -DZNS_VQT_WAIT=n): 0 parked (nanosleep back-off 20..80 ns),
// 1 plain try_wait spin (the instruction itself suspends the thread for a while), 2 nanosleep fixed at 20 ns
#ifndef ZNS_VQT_WAIT
#define ZNS_VQT_WAIT 0
#endif
#ifndef ZNS_VQT_HINT
#define ZNS_VQT_HINT 100000
#endif
#ifndef ZNS_VQT_SLEEP
#define ZNS_VQT_SLEEP 200
#endif
__device__ __forceinline__ void vqt_wait(uint32_t bar, uint32_t parity) {
#if ZNS_VQT_WAIT == 0
  mbar_wait_parked(bar, parity);
#elif ZNS_VQT_WAIT == 1
#pragma unroll 1
  for (int it = 0; it < 400000000; ++it)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
#elif ZNS_VQT_WAIT == 2
#pragma unroll 1
  for (int it = 0; it < 100000000; ++it) {
    if (mbar_try_wait(bar, parity)) return;
    __nanosleep(20);
  }
  __trap();
#elif ZNS_VQT_WAIT == 3
  // try_wait with an explicit suspend-time hint: the hardware parks the thread until the phase completes or the hint expires
#pragma unroll 1
  for (int it = 0; it < 4000000; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(ZNS_VQT_HINT)
        : "memory");
    if (ok) return;
  }
  __trap();
#else
#pragma unroll 1
  for (int it = 0; it < 100000000; ++it) {
    if (mbar_try_wait(bar, parity)) return;
    __nanosleep(ZNS_VQT_SLEEP);
  }
  __trap();
#endif
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// FPR1: one frame per signal row (levels whose hop is >= 32 samples: the four big levels) -- a compile-time switch so that the
// register needs of the multi-frame epilogue (72 accumulator columns per thread) do not spill the hot single-frame one.
// GROUPED (levels >= 1 whose accumulator rings have two stages): the sixteen epilogue warps form two groups of eight and
// group g drains accumulator stage g, i.e. every other tile -- a tile's drain is one latency chain (barrier -> tcgen05.ld ->
// FP -> store -> barrier, ~3 K clocks whether eight or sixteen warps share it), so two tiles in flight double the rate.
template <bool SRC_F32, bool FPR1, bool GROUPED>
__global__ void __launch_bounds__(VQT_LEVEL_THREADS(SRC_F32), 1)
vqt_level_kernel(const __grid_constant__ VqtLevelDev L, const __grid_constant__ VqtLevelArgs A) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar_full[8], bar_empty[8], bar_acc_full[4], bar_acc_empty[4];   // acc barriers: [type * 2 + stage]
  __shared__ uint32_t tmem_slot;
  const int q = L.q, R = 8 * q;
  const uint32_t slot_bytes = 2u * (uint32_t)L.slot_term_bytes;
  uint8_t* sZero = sm;                                   // 128 zero bytes (both operands of the clearing MMAs)
  uint8_t* sB = sm + VQT_ZERO_BYTES;
  uint8_t* sRing = sB + ((L.b_bytes + 127) / 128) * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  static_assert(!(SRC_F32 && GROUPED), "level 0 has eight epilogue warps");
  constexpr int EPQ = (SRC_F32 || GROUPED) ? 2 : 4;      // epilogue warps per TMEM lane quadrant that share a tile (must divide the twelve bins)
  const bool is_epi = warp < VQT_EPI_WARPS || (!SRC_F32 && warp >= VQT_EXTRA_WARP0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < L.n_slots; ++i) {
      // fp32 source: VQT_LOAD_WARPS / n_slots loader warps fill a slot and all their lanes arrive; bulk copies: one arrive.expect_tx
      mbar_init(smem_u32(&bar_full[i]), SRC_F32 ? 32 * (VQT_LOAD_WARPS / L.n_slots) : 1);
      // every issuer that owns accumulator units commits once per group.  An issuer WITHOUT units must stay out of the
      // count: nothing throttles it (it never waits for operands), so it would run through all its tiles at once and its
      // arrivals of later ring turns would complete a slot's "empty" phase before the working issuers have left the slot.
      int n_active = 0;
      for (int k = 0; k < ZNS_VQT_ISSUERS; ++k) n_active += L.seg_begin[k][L.gpt] > L.seg_begin[k][0] ? 1 : 0;
      mbar_init(smem_u32(&bar_empty[i]), n_active);
    }
    for (int t = 0; t < 4; ++t) {
      mbar_init(smem_u32(&bar_acc_full[t]), (t < 2 && L.fb_n2 == 0) ? 1 : 2);   // accumulator units (parts) per job
      mbar_init(smem_u32(&bar_acc_empty[t]), 32 * 4 * EPQ);
    }
    mbar_fence_init();
  }
  // Programmatic dependent launch: the level kernels of a pass are launched back to back with the "programmatic stream
  // serialization" attribute, so the CTAs of level i + 1 start on an SM as soon as level i's CTA has left it and run this
  // prologue (barriers, TMEM, the constant coefficient image) under level i's tail; griddepcontrol.wait then holds them until
  // level i has completed and its level signal is visible.  Both instructions are no-ops in a launch without the attribute.
  pdl_launch_dependents();
  if (warp == VQT_ISSUE_WARP0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  if (threadIdx.x < VQT_ZERO_BYTES / 4) reinterpret_cast<uint32_t*>(sZero)[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < L.b_bytes / 16; i += VQT_LEVEL_THREADS(SRC_F32))
    reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(A.bimg) + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int n_my_tiles = (A.n_tiles > (int)blockIdx.x) ? (A.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int sh0 = L.ring_stages[0] == 2 ? 1 : 0, sh1 = L.ring_stages[1] == 2 ? 1 : 0;   // log2(stages)

  if (warp >= VQT_ISSUE_WARP0 && warp < VQT_LOAD_WARP0) {
    // =================================== MMA issuers: one elected thread each, O(1) bookkeeping ===================================
    const int isr = warp - VQT_ISSUE_WARP0;
    if (L.seg_begin[isr][L.gpt] > L.seg_begin[isr][0] && elect_one()) {     // issuers without accumulator units sit the level out
      const uint32_t ring16 = smem_u32(sRing) >> 4, b16 = smem_u32(sB) >> 4, z16 = smem_u32(sZero) >> 4;
      const uint32_t slot16 = slot_bytes >> 4;
      const uint64_t hi_norm = (uint64_t)((128u >> 4) | (1u << 14)) << 32, hi_zero = (uint64_t)(1u << 14) << 32;
      const uint64_t zdesc = hi_zero | z16;
      const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
      const uint32_t accf0 = smem_u32(&bar_acc_full[0]), acce0 = smem_u32(&bar_acc_empty[0]);
      int slot = 0;
      uint32_t full_par = 0;
      uint32_t dec_inst0 = 0;
      long long t_wait_full = 0, t_wait_acc = 0, n_mma_issued = 0;
      const long long t_begin = VQT_CLOCK();
#pragma unroll 1
      for (int ts = 0; ts < n_my_tiles; ++ts, dec_inst0 += (uint32_t)L.n_pass) {
#pragma unroll 1
        for (int pos = 0; pos < L.gpt; ++pos) {
          const int s_begin = L.seg_begin[isr][pos], s_end = L.seg_begin[isr][pos + 1];
          VQT_TL(ts >= 4 && ts < 8 && isr == 0, 240 + ((ts - 4) * 4 + pos) * 2);     // issuer 0: before the operand wait
          if (s_begin < s_end) {
            const long long t0 = VQT_CLOCK();
            vqt_wait(full0 + 8 * slot, full_par);
            t_wait_full += VQT_CLOCK() - t0;
            tc_fence_after();
          }
          VQT_TL(ts >= 4 && ts < 8, 64 + ((((ts - 4) * 4 + isr) * 4 + pos) * 2));
          const uint32_t a16 = ring16 + (uint32_t)slot * slot16;
#pragma unroll 1
          for (int si = s_begin; si < s_end; ++si) {
            const VqtSeg sg = L.seg[si];
            // accumulator stage of this job instance: ring slot, its barrier, the parity of its "drained" phase
            const int type = sg.job ? 1 : 0;
            const uint32_t inst = type ? dec_inst0 + (uint32_t)(sg.job - 1) : (uint32_t)ts;
            const int sh = type ? sh1 : sh0;
            const uint32_t stage = inst & (uint32_t)sh;
            const uint32_t bidx = (uint32_t)(2 * type) + stage;
            const uint32_t cb = tmem + stage * (uint32_t)L.ring_width[type];
            if (sg.flags & 1) {
              const long long t0 = VQT_CLOCK();
              vqt_wait(acce0 + 8 * bidx, ((inst >> sh) & 1) ^ 1);
              t_wait_acc += VQT_CLOCK() - t0;
              tc_fence_after();
              VQT_TL(ts >= 4 && ts < 8 && isr == 0, 240 + ((ts - 4) * 4 + pos) * 2 + 1);   // issuer 0: accumulator stage granted (last one of the position)
              // clear this unit's columns: accumulate = 0 with the zero block as both operands
              const int w0 = type ? L.dec_wp : L.fb_n1, w1 = type ? L.dec_wp : L.fb_n2;
              const int c_lo = sg.part ? w0 : 0, c_hi = sg.part ? w0 + w1 : w0;
              for (int c = c_lo; c < c_hi; c += 256)
                umma_f16(cb + (uint32_t)(L.ring_base[type] + c), zdesc, zdesc, umma_idesc_f16(128, min(256, c_hi - c)), 0u);
            }
            const int i_end = sg.begin + sg.count;
#pragma unroll 4
            for (int i = sg.begin; i < i_end; ++i) {
              if (VQT_KO(4)) break;
              const VqtMmaPacked m = L.pk[i];
              umma_f16(cb + m.col, hi_norm | (uint64_t)(m.a_lo + a16), hi_norm | (uint64_t)(m.b_lo + b16), m.idesc, 1u);
            }
            n_mma_issued += sg.count;
            if (sg.flags & 2) umma_commit(accf0 + 8 * bidx);
          }
          umma_commit(empty0 + 8 * slot);       // this issuer no longer reads the slot (immediate when it had no MMAs)
          VQT_TL(ts >= 4 && ts < 8, 64 + ((((ts - 4) * 4 + isr) * 4 + pos) * 2) + 1);
          if (++slot == L.n_slots) { slot = 0; full_par ^= 1; }
        }
      }
      if (VQT_TIMING_ON && A.dbg && blockIdx.x == 0) {
        long long* d = A.dbg + 4 * isr;
        d[0] = VQT_CLOCK() - t_begin; d[1] = t_wait_full; d[2] = t_wait_acc; d[3] = n_mma_issued;
      }
    }
    __syncwarp();
  } else if (is_epi) {
    // =================================== epilogue ===================================
    // EPQ warps per TMEM lane quadrant; warp `sub` of a quadrant takes every EPQ-th 16-column decimator block and
    // its share of the filterbank bins / frames.  The roles run one dependent chain per warp (tcgen05.ld -> FP -> pack ->
    // store), so throughput comes from the number of warps, not from the instruction count.
    // (a warp reads the TMEM lanes of quadrant warp % 4: the extra warps 13..20 are quadrants 1, 2, 3, 0, 1, 2, 3, 0)
    const int quad = warp & 3, sub_all = warp < VQT_EPI_WARPS ? warp >> 2 : 2 + ((warp - VQT_EXTRA_WARP0) >> 2);
    const int sub = GROUPED ? (sub_all & 1) : sub_all;     // index among the warps of the quadrant that share a tile
    const int group = GROUPED ? (sub_all >> 1) : 0;        // GROUPED: warps 0..7 drain the even tiles (stage 0), warps 13..20 the odd ones
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    long long t_wait_ep = 0, t_ld = 0, t_math = 0, t_st = 0;
    const long long t_begin_ep = VQT_CLOCK();
    int dst_shift = 3;                                      // log2(samples per row) of the next level
    while ((8 << (dst_shift - 3)) < 8 * A.dst_q) ++dst_shift;
    constexpr int kBinsPerWarp = 12 / EPQ;              // fpr == 1: bins of this warp
    float isl[kBinsPerWarp];                                // 1 / sqrt(L_k) / (coefficient scale) of the bins this thread writes
#pragma unroll
    for (int k = 0; k < kBinsPerWarp; ++k) isl[k] = __ldg(A.inv_sqrt_len + L.bin0 + kBinsPerWarp * sub + k) * L.fb_scale;
    // everything the per-job code needs from the level description, read ONCE with compile-time subscripts (uniform loads):
    // a run-time subscript into the kernel parameters is an indexed constant load of several hundred clocks, and the job loop
    // used to issue a chain of them per job
    const int n_jobs = L.n_jobs, n_pass = L.n_pass, dec_w = L.dec_w, dec_wp = L.dec_wp, fb_n1 = L.fb_n1, fpr = L.fpr;
    const int job_of0 = L.ep_job[0], job_of1 = L.ep_job[1], job_of2 = L.ep_job[2];
    const int rbase0 = L.ring_base[0], rbase1 = L.ring_base[1], rwid0 = L.ring_width[0], rwid1 = L.ring_width[1];
    const uint32_t accf = smem_u32(&bar_acc_full[0]), acce = smem_u32(&bar_acc_empty[0]);
    const int n_frames = A.n_frames, n_valid = A.n_valid;
    const bool out_vec = (n_frames % L.fpr) == 0 && (reinterpret_cast<uintptr_t>(A.out) & 15) == 0;   // rows of `out` keep vector alignment
    const long long dst_cap = A.dst_cap;
    const float s_a = L.dec_scale, s_b = L.dec_scale * (1.f / 2048.f);
    const int dst_q = A.dst_q, dst_rtot = A.dst_rtot, dst_hb = A.dst_hb, dst_ha = A.dst_ha;
    const int o_half = dst_q == 1 ? 8 : dst_rtot * 8;      // the next chunk of a row: next plane (tiled) / adjacent (linear)
    const long long o_dup = (long long)(dst_q * dst_rtot - 128) * 8;   // a row's copy in the neighbouring tile's halo
    int tau = (int)blockIdx.x, clip = tau / A.tiles_per_clip, tile_in_clip = tau - clip * A.tiles_per_clip;
    const int clip_step = (int)gridDim.x / A.tiles_per_clip, tile_step = (int)gridDim.x - clip_step * A.tiles_per_clip;
#pragma unroll 1
    for (int ts = 0; ts < n_my_tiles; ++ts) {
      if (GROUPED && (ts & 1) != group) {                  // the other group's tile
        clip += clip_step; tile_in_clip += tile_step;
        if (tile_in_clip >= A.tiles_per_clip) { tile_in_clip -= A.tiles_per_clip; ++clip; }
        continue;
      }
      const int row_first = tile_in_clip * 128;
      const int g = row_first + quad * 32 + lane;          // row of this thread inside the clip (< 2^24)
      // every decimator output of the tile lies inside the signal: no per-element masking (all but the last tile of a clip)
      const bool tile_valid = (long long)(row_first + 128) * dec_w <= n_valid;
#pragma unroll
      for (int jj = 0; jj < ZNS_VQT_MAX_JOBS; ++jj) {
        if (jj >= n_jobs) break;
        const int job = jj == 0 ? job_of0 : (jj == 1 ? job_of1 : job_of2);
        const int type = job ? 1 : 0;
        const uint32_t inst = type ? (uint32_t)(ts * n_pass + (job - 1)) : (uint32_t)ts;
        const int sh = type ? sh1 : sh0;
        const uint32_t stage = inst & (uint32_t)sh;
        const uint32_t bidx = (uint32_t)(type * 2) + stage;
        {
          const long long t0 = VQT_CLOCK();
          vqt_wait(accf + 8 * bidx, (inst >> sh) & 1);     // every lane polls (see the loader: no single-lane code before the drain)
          t_wait_ep += VQT_CLOCK() - t0;
        }
        tc_fence_after();
        VQT_TL(ts >= 4 && ts < 8 && quad == 0 && lane == 0, 192 + ((((ts - 4) * 2 + sub) * 3 + jj) * 2));
        const uint32_t acc = tl + (uint32_t)((type ? rbase1 : rbase0) + (int)stage * (type ? rwid1 : rwid0));
        if (VQT_KO(2)) {
        } else if (type == 1) {
          const int c0 = 64 * (job - 1);
          const int w = min(64, dec_w - c0);
          uint16_t* dh = A.dst_hi + (size_t)clip * A.dst_stride;
          uint16_t* dl = A.dst_lo + (size_t)clip * A.dst_stride;
          const int n_blk = (w + 15) / 16;
          const int t_row = g * dec_w + c0;                            // first output of this row in this pass (< 2^31)
#pragma unroll 1
          for (int kb = sub; kb < n_blk; kb += EPQ) {
            uint32_t a[16], b[16];
            const long long c0_ = VQT_CLOCK();
            tmem_ld_32x16(acc + 16 * kb, a);
            tmem_ld_32x16(acc + dec_wp + 16 * kb, b);
            tmem_ld_wait();
            const long long c1_ = VQT_CLOCK();
            t_ld += c1_ - c0_;
            const int t0 = t_row + 16 * kb;
            float yv[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) yv[o] = fmaf(__uint_as_float(b[o]), s_b, __uint_as_float(a[o]) * s_a);
            if (!tile_valid) {
#pragma unroll
              for (int o = 0; o < 16; ++o) if (t0 + o >= n_valid) yv[o] = 0.f;
            }
            uint4 h1a, h2a, h1b, h2b;
            split8(yv, h1a, h2a);
            split8(yv + 8, h1b, h2b);
            const long long c2_ = VQT_CLOCK();
            t_math += c2_ - c1_;
            if (t0 < dst_cap && !VQT_KO(16)) {
              if (dst_q == 1) {
                const long long o = 1024 + (long long)t0;
                if (dec_w >= 16) {         // 16 adjacent samples
                  *reinterpret_cast<uint4*>(dh + o) = h1a; *reinterpret_cast<uint4*>(dh + o + 8) = h1b;
                  *reinterpret_cast<uint4*>(dl + o) = h2a; *reinterpret_cast<uint4*>(dl + o + 8) = h2b;
                } else {                   // dec_w == 4: four outputs per row
                  *reinterpret_cast<uint2*>(dh + o) = make_uint2(h1a.x, h1a.y);
                  *reinterpret_cast<uint2*>(dl + o) = make_uint2(h2a.x, h2a.y);
                }
              } else {
                // 16 outputs = two chunks (adjacent planes) of one row of the next level's tile image
                const int row2 = t0 >> dst_shift, c2 = (t0 & ((1 << dst_shift) - 1)) >> 3;
                const int t2 = row2 >> 7, r2 = row2 & 127;
                const long long o = (((long long)t2 * dst_q + c2) * dst_rtot + dst_hb + r2) * 8;
                *reinterpret_cast<uint4*>(dh + o) = h1a; *reinterpret_cast<uint4*>(dh + o + o_half) = h1b;
                *reinterpret_cast<uint4*>(dl + o) = h2a; *reinterpret_cast<uint4*>(dl + o + o_half) = h2b;
                if (r2 < dst_ha && t2 > 0) {            // also the "halo after" rows of the previous tile
                  *reinterpret_cast<uint4*>(dh + o - o_dup) = h1a; *reinterpret_cast<uint4*>(dh + o - o_dup + o_half) = h1b;
                  *reinterpret_cast<uint4*>(dl + o - o_dup) = h2a; *reinterpret_cast<uint4*>(dl + o - o_dup + o_half) = h2b;
                }
                if (r2 >= 128 - dst_hb) {               // and the "halo before" rows of the next tile
                  *reinterpret_cast<uint4*>(dh + o + o_dup) = h1a; *reinterpret_cast<uint4*>(dh + o + o_dup + o_half) = h1b;
                  *reinterpret_cast<uint4*>(dl + o + o_dup) = h2a; *reinterpret_cast<uint4*>(dl + o + o_dup + o_half) = h2b;
                }
              }
            }
            t_st += VQT_CLOCK() - c2_;
          }
        } else if (FPR1) {
          // one frame per row: the warps of a lane quadrant share the twelve bins
          const int f = g;
          constexpr int NC = 2 * kBinsPerWarp;             // accumulator columns of this warp's bins (re, im interleaved)
          constexpr int NL = NC <= 8 ? 8 : 16;             // columns loaded (NC of them are used; the rest are allocated TMEM)
          uint32_t fa[NL], fg[NL], fb[NL];
          const int c0 = NC * sub;
          if (NL == 8) {
            tmem_ld_32x8(acc + c0, fa);                    // x1 . g1
            tmem_ld_32x8(acc + 24 + c0, fg);               // x1 . g2
            tmem_ld_32x8(acc + fb_n1 + c0, fb);          // x2 . g1
          } else {
            tmem_ld_32x16(acc + c0, fa);
            tmem_ld_32x16(acc + 24 + c0, fg);
            tmem_ld_32x16(acc + fb_n1 + c0, fb);
          }
          tmem_ld_wait();
          if (f < n_frames && !VQT_KO(16)) {
            float* op = A.out + ((size_t)clip * A.n_bins + L.bin0 + kBinsPerWarp * sub) * n_frames + f;
#pragma unroll
            for (int k = 0; k < kBinsPerWarp; ++k) {
              const float re = __uint_as_float(fa[2 * k]) + (__uint_as_float(fg[2 * k]) + __uint_as_float(fb[2 * k])) * (1.f / 2048.f);
              const float im = __uint_as_float(fa[2 * k + 1]) + (__uint_as_float(fg[2 * k + 1]) + __uint_as_float(fb[2 * k + 1])) * (1.f / 2048.f);
              op[(size_t)k * n_frames] = log_approx(fmaf(sqrt_approx(fmaf(re, re, im * im)), isl[k], 1e-9f));
            }
          }
        } else if (!FPR1) {
          // several frames per row (hop < 32 samples): a thread owns the fpr CONSECUTIVE frames of its row, so the warps of a
          // quadrant split the BINS and a thread writes its frames of one bin as one 8 / 16-byte store -- consecutive lanes,
          // consecutive addresses (one store per frame and bin was a 16-byte-strided scatter, 4 x the store instructions)
          // Three bins (six accumulator columns, one 32x8 load per product term) at a time: 36 live registers instead of 72.
          constexpr int KB = 12 / EPQ;                     // bins of this warp
          const int f0 = g * fpr;
          const bool whole = f0 + fpr <= n_frames && out_vec;
#pragma unroll
          for (int kh = 0; kh < KB; kh += 3) {
            const int k0 = KB * sub + kh;
            float res[4][3];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < fpr) {
                uint32_t fa[8], fg[8];
                tmem_ld_32x8(acc + 24 * j + 2 * k0, fa);                // x1 . g1
                tmem_ld_32x8(acc + 24 * fpr + 24 * j + 2 * k0, fg);     // x1 . g2 + x2 . g1 (merged columns, see build_level)
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                  const float re = fmaf(__uint_as_float(fg[2 * k]), 1.f / 2048.f, __uint_as_float(fa[2 * k]));
                  const float im = fmaf(__uint_as_float(fg[2 * k + 1]), 1.f / 2048.f, __uint_as_float(fa[2 * k + 1]));
                  res[j][k] = log_approx(fmaf(sqrt_approx(fmaf(re, re, im * im)), isl[kh + k], 1e-9f));
                }
              }
            }
            if (f0 < n_frames && !VQT_KO(16)) {
              float* op = A.out + ((size_t)clip * A.n_bins + L.bin0 + k0) * n_frames + f0;
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                float* o = op + (size_t)k * n_frames;
                if (whole && fpr == 4) *reinterpret_cast<float4*>(o) = make_float4(res[0][k], res[1][k], res[2][k], res[3][k]);
                else if (whole && fpr == 2) *reinterpret_cast<float2*>(o) = make_float2(res[0][k], res[1][k]);
                else {
#pragma unroll
                  for (int j = 0; j < 4; ++j) if (j < fpr && f0 + j < n_frames) o[j] = res[j][k];
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(acce + 8 * bidx);              // every lane, once it has read its accumulator rows
        VQT_TL(ts >= 4 && ts < 8 && quad == 0 && lane == 0, 192 + ((((ts - 4) * 2 + sub) * 3 + jj) * 2) + 1);
      }
      // next tile of this CTA: tau += gridDim.x without a division
      clip += clip_step; tile_in_clip += tile_step;
      if (tile_in_clip >= A.tiles_per_clip) { tile_in_clip -= A.tiles_per_clip; ++clip; }
    }
    if (VQT_TIMING_ON && A.dbg && blockIdx.x == 0 && (warp == 0 || warp == 4) && lane == 0) {
      A.dbg[16 + 2 * sub] = VQT_CLOCK() - t_begin_ep; A.dbg[17 + 2 * sub] = t_wait_ep;
      if (sub == 0) { A.dbg[27] = t_ld; A.dbg[28] = t_math; A.dbg[29] = t_st; }
    }
  } else if (!SRC_F32) {
    // =================================== loader of the levels >= 1: one warp, two bulk copies per plane group ===================================
    // The source is stored tile by tile in plane order (see level_index): a plane group of a tile is ONE contiguous image per
    // fp16 term (q == 1: the rows are consecutive chunks of the linear signal), so a slot fill is two bulk copies.
    if (warp == VQT_LOAD_WARP0) {
      const int n_rows = 128 + L.hb + L.ha;
      const int n_groups = n_my_tiles * L.gpt;
      const uint32_t tb = L.pg > 1 ? (uint32_t)L.slot_term_bytes : (uint32_t)(n_rows * 16);
      int slot = 0, pos = 0, tau = (int)blockIdx.x;
      uint32_t par = 1;                           // first use of a slot passes immediately
      long long t_wait_ld = 0;
      const long long t_begin_ld = VQT_CLOCK();
#pragma unroll 1
      for (int gi = 0; gi < n_groups; ++gi) {
        const int clip = tau / A.tiles_per_clip;
        const long long t = tau - clip * A.tiles_per_clip;                 // tile inside the clip
        const uint32_t s1 = smem_u32(sRing) + (uint32_t)slot * slot_bytes;
        const uint32_t bfull = smem_u32(&bar_full[slot]);
        {
          const long long t0 = VQT_CLOCK();
          vqt_wait(smem_u32(&bar_empty[slot]), par);                       // every lane polls
          t_wait_ld += VQT_CLOCK() - t0;
        }
        if (VQT_KO(8)) {
          if (lane == 0) mbar_arrive(bfull);
        } else {
          if (lane == 0) mbar_expect_tx(bfull, 2u * tb);
          __syncwarp();
          if (lane < 2) {
            const uint16_t* base = (lane ? A.src_lo : A.src_hi) + (size_t)clip * A.src_stride;
            const uint16_t* src = L.pg > 1 ? base + ((t * q + L.g_order[pos] * L.pg) * (long long)L.rtot) * 8
                                           : base + 1024 + (t * 128 - L.hb) * 8;
            bulk_g2s(s1 + (lane ? (uint32_t)L.slot_term_bytes : 0u), src, tb, bfull);
          }
        }
        if (++slot == L.n_slots) { slot = 0; par ^= 1; }
        if (++pos == L.gpt) { pos = 0; tau += (int)gridDim.x; }
      }
      if (VQT_TIMING_ON && A.dbg && blockIdx.x == 0 && lane == 0) { A.dbg[20] = VQT_CLOCK() - t_begin_ld; A.dbg[21] = t_wait_ld; }
    }
  } else if (warp < VQT_LOAD_WARP0 + VQT_LOAD_WARPS) {
    // =================================== loaders of level 0: VQT_LOAD_WARPS / n_slots warps per ring slot ===================================
    const int lw = warp - VQT_LOAD_WARP0;
    const int wps = VQT_LOAD_WARPS / L.n_slots;              // warps per slot
    const int my_slot = lw % L.n_slots, sub = lw / L.n_slots; // this warp takes every wps-th block of 32 chunks of its slot's groups
    const int pg_shift = L.pg == 8 ? 3 : (L.pg == 4 ? 2 : 0);
    const int n_rows = 128 + L.hb + L.ha;
    const int n_chunks = n_rows * L.pg;
    const int n_groups = n_my_tiles * L.gpt;
    {
      const uint32_t s1 = smem_u32(sRing) + (uint32_t)my_slot * slot_bytes;
      const uint32_t s2 = s1 + (uint32_t)L.slot_term_bytes;
      const uint32_t bfull = smem_u32(&bar_full[my_slot]), bempty = smem_u32(&bar_empty[my_slot]);
      uint32_t par = 1;                           // first use of a slot passes immediately
      long long t_wait_ld = 0, t_p1 = 0, t_cpw = 0, t_p2 = 0;
      const long long t_begin_ld = VQT_CLOCK();
      // lane -> (row, plane) of its first chunk; each further chunk of this warp is (32 / pg) * wps rows down
      const int i_first = lane + 32 * sub;
      const int r_first = i_first >> pg_shift, c_lane = i_first & (L.pg - 1);
      const int r_step = (32 >> pg_shift) * wps;
      const uint32_t off0 = (uint32_t)(c_lane * L.a_lbo + 16 * r_first), off_step = 16u * (uint32_t)r_step;
      const int n_it = (n_chunks > i_first ? (n_chunks - i_first + 32 * wps - 1) / (32 * wps) : 0);
#pragma unroll 1
      for (int gi = my_slot; gi < n_groups; gi += L.n_slots, par ^= 1) {
        const int ts = gi / L.gpt, pos = gi - ts * L.gpt;
        const int tau = (int)blockIdx.x + ts * (int)gridDim.x;
        const int clip = tau / A.tiles_per_clip;
        const long long row0 = (long long)(tau - clip * A.tiles_per_clip) * 128;
        const float* yb = A.y32 + (size_t)clip * A.src_stride;
        const bool vec_ok = (reinterpret_cast<uintptr_t>(yb) & 15) == 0;
        const int plane0 = L.g_order[pos] * L.pg;
        const long long base_s = (row0 - L.hb) * R + 8LL * plane0;     // sample of chunk (row 0, plane 0 of the group)
        {
          const long long t0 = VQT_CLOCK();
          // EVERY lane polls: `if (lane == 0) wait(); __syncwarp();` left lane 0 and lanes 1..31 as two convergence groups that
          // ran the fill loop (and, in the epilogue, the whole drain) one after the other -- per-lane clock stamps showed
          // lane 0 finishing a fill 6..10 K clocks apart from the rest (0.635 -> 0.540 ms at cfg2 once removed).  No single-lane
          // code in front of heavy warp-wide work.
          vqt_wait(bempty, par);
          t_wait_ld += VQT_CLOCK() - t0;
        }
        VQT_TL(ts >= 4 && ts < 8 && lane == 0, ((ts - 4) * 8 + lw) * 2);
        if (VQT_KO(8)) {
          mbar_arrive(bfull);
          continue;
        }
        if (vec_ok) {
          // global -> registers -> two-term fp16 split -> shared memory, VQT_LD_U chunks of 32 bytes per lane in flight: the
          // shared-memory image is written once (a cp.async staging pass costs two more trips through shared memory, and
          // the in-place conversion behind it was one dependent LDS -> convert -> STS chain per chunk)
          const long long tp0 = VQT_CLOCK();
          if (VQT_KO(1)) {
          } else {
            // One path for interior tiles and for the first / last tile of a clip (2 of 15 tiles at cfg2: a scalar loop for them
            // cost as much as five interior fills): a chunk that lies outside the signal is zero, a chunk that straddles its end
            // (at most one per row) is read element by element.
            // (sample indices fit 32 bits: a clip holds fewer than 2^31 samples)
            int s_it = (int)base_s + r_first * R + 8 * c_lane;       // first sample of this lane's chunk of the current iteration
            const int s_step = r_step * R;
            uint32_t d1 = s1 + off0, d2 = s2 + off0;
            const int n_sig_i = A.n_sig;
            // warp-uniform trip count (lanes own 16 or 17 chunks): a lane-dependent bound splits the warp for the whole loop
            // under independent thread scheduling -- measured: lanes 24..31 ran the loop AFTER lanes 0..23, doubling the fill
            // time of every slot -- so the bound is the maximum and the chunks beyond a lane's share are predicated off
            const int n_it_max = (n_chunks + 32 * wps - 1) / (32 * wps);
            // VQT_LD_U chunks (32 bytes each) per lane in flight: the loop is one load -> wait -> convert -> store chain per
            // warp, so its rate is (bytes in flight) / (memory latency + convert time)
#pragma unroll 1
            for (int it = 0; it < n_it_max; it += VQT_LD_U) {
              float4 va[VQT_LD_U], vb[VQT_LD_U];
              uint32_t ragged = 0;
#pragma unroll
              for (int u = 0; u < VQT_LD_U; ++u) {
                const int s0 = s_it + u * s_step;
                const bool in = it + u < n_it && s0 >= 0 && s0 + 8 <= n_sig_i;
                const float4* sp = reinterpret_cast<const float4*>(yb + s0);
                va[u] = in ? __ldg(sp) : make_float4(0.f, 0.f, 0.f, 0.f);
                vb[u] = in ? __ldg(sp + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (!in && it + u < n_it && s0 < n_sig_i && s0 + 8 > 0) ragged |= 1u << u;
              }
#pragma unroll
              for (int u = 0; u < VQT_LD_U; ++u) {
                const float v[8] = {va[u].x, va[u].y, va[u].z, va[u].w, vb[u].x, vb[u].y, vb[u].z, vb[u].w};
                uint4 c1, c2;
                split8(v, c1, c2);
                if (it + u < n_it) {
                  sts128(d1 + u * off_step, c1);
                  sts128(d2 + u * off_step, c2);
                }
              }
              if (ragged) {                 // rare: overwrite the straddling chunks (stored as zeros above)
#pragma unroll 1
                for (int u = 0; u < VQT_LD_U; ++u) {
                  if (ragged & (1u << u)) {
                    const int s0 = s_it + u * s_step;
                    float e[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) e[k] = (s0 + k >= 0 && s0 + k < n_sig_i) ? __ldg(yb + s0 + k) : 0.f;
                    uint4 c1, c2;
                    split8(e, c1, c2);
                    sts128(d1 + u * off_step, c1);
                    sts128(d2 + u * off_step, c2);
                  }
                }
              }
              d1 += VQT_LD_U * off_step; d2 += VQT_LD_U * off_step; s_it += VQT_LD_U * s_step;
            }
          }
          t_p1 += VQT_CLOCK() - tp0;
        } else {
          // unaligned fp32 clips (n_samples not a multiple of 4): element-wise loads
          for (int it = 0; it < n_it; ++it) {
            const int r = r_first + it * r_step;
            const long long s0 = base_s + (long long)r * R + 8 * c_lane;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (s0 + e >= 0 && s0 + e < A.n_sig) ? __ldg(yb + s0 + e) : 0.f;
            uint4 c1, c2;
            split8(v, c1, c2);
            const uint32_t off = off0 + (uint32_t)it * off_step;
            sts128(s1 + off, c1);
            sts128(s2 + off, c2);
          }
        }
        VQT_TL(ts == 5 && lw == 1, 384 + lane);                                         // every lane of loader warp 1, tile 5: work done
        VQT_TL(ts >= 4 && ts < 8 && lane == 0, 280 + ((ts - 4) * 8 + lw) * 3);        // lane 0: work done
        VQT_TL(ts >= 4 && ts < 8 && lane == 31, 280 + ((ts - 4) * 8 + lw) * 3 + 1);   // lane 31: work done
        if (!VQT_KO(32)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        VQT_TL(ts >= 4 && ts < 8 && lane == 0, 280 + ((ts - 4) * 8 + lw) * 3 + 2);    // lane 0: fenced
        mbar_arrive(bfull);                        // every lane (its own stores are fenced)
#ifndef ZNS_VQT_NO_L2_PREFETCH
        // (single-lane block: kept BEHIND the warp-wide fill, see above)
        if (SRC_F32 && pos == 0 && sub == 0 && lane == 0) {
          // pull the samples of the tile VQT_PF_DIST tiles ahead (one contiguous span of the clip) into L2: the ring holds a
          // single tile, so its loads cannot be issued early -- but their DRAM latency can be taken now.  A distance of one
          // tile leaves the demand loads barely behind the prefetch (28 % L2 hits), yet two or three tiles measured SLOWER.
#pragma unroll 1
          for (int d = (ts == 0 ? 1 : VQT_PF_DIST); d <= VQT_PF_DIST; ++d) {
            if (ts + d >= n_my_tiles) break;
            const int tau_n = tau + d * (int)gridDim.x;
            const int clip_n = tau_n / A.tiles_per_clip;
            long long s_lo = ((long long)(tau_n - clip_n * A.tiles_per_clip) * 128 - L.hb) * R;
            long long s_hi = s_lo + (long long)n_rows * R;
            s_lo = max(s_lo, 0LL); s_hi = min(s_hi, (long long)A.n_sig);
            const float* pn = A.y32 + (size_t)clip_n * A.src_stride + s_lo;
            const uint32_t bytes = (uint32_t)((s_hi - s_lo) * 4) & ~15u;
            if (s_hi > s_lo && bytes > 0 && (reinterpret_cast<uintptr_t>(pn) & 15) == 0)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pn), "r"(bytes) : "memory");
          }
        }
#endif
        VQT_TL(ts >= 4 && ts < 8 && lane == 0, ((ts - 4) * 8 + lw) * 2 + 1);
      }
      if (VQT_TIMING_ON && A.dbg && blockIdx.x == 0 && lw == 0 && lane == 0) { A.dbg[20] = VQT_CLOCK() - t_begin_ld; A.dbg[21] = t_wait_ld;
                                                              A.dbg[22] = t_p1; A.dbg[23] = t_cpw; A.dbg[24] = t_p2; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == VQT_ISSUE_WARP0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- frames whose window crosses a clip edge: reflect padding, direct fp32 evaluation -------------
struct VqtEdgeParams {
  int n_oct, bpo, n_bins, n_frames;
  int n_fft[ZNS_VQT_MAX_OCT], hop[ZNS_VQT_MAX_OCT], n_sig[ZNS_VQT_MAX_OCT];
  int item0[ZNS_VQT_MAX_OCT + 1];          // first (frame, filter) item of each octave
  int n_left[ZNS_VQT_MAX_OCT], t_right[ZNS_VQT_MAX_OCT];
  long long stride[ZNS_VQT_MAX_OCT];
  int q[ZNS_VQT_MAX_OCT], rtot[ZNS_VQT_MAX_OCT], hb[ZNS_VQT_MAX_OCT];   // layout of the level buffers (level_index)
  const float* coef[ZNS_VQT_MAX_OCT];      // [n][2][bpo/2][2]
  const uint16_t* hi[ZNS_VQT_MAX_OCT];
  const uint16_t* lo[ZNS_VQT_MAX_OCT];
};

__device__ __forceinline__ int reflect_idx32(int qq, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  qq %= period;
  if (qq < 0) qq += period;
  return qq < n ? qq : period - qq;
}

// one warp per edge frame: the lanes gather the frame's window ONCE into shared memory (reflect index + level layout are the
// expensive part), then lane j < 2 bpo accumulates output component j (filter j / 2, re / im) over the taps -- coefficient
// reads are coalesced ([tap][filter](re, im)), the sample is a shared-memory broadcast, no cross-lane reduction (the first
// version reduced 24 partial sums with 120 shuffles and needed 62 registers: three waves of blocks at cfg2)
#define VQT_EDGE_MAX_BPO 12
#define VQT_EDGE_MAX_NFFT 256
__global__ void __launch_bounds__(256)
vqt_edge_kernel(const __grid_constant__ VqtEdgeParams P, const float* __restrict__ y32, long long y_stride,
                const float* __restrict__ inv_sqrt_len, float* __restrict__ out) {
  __shared__ float win[8][VQT_EDGE_MAX_NFFT];
  const int clip = blockIdx.z;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 8 + wib;
  pdl_wait();                              // launched behind the last level kernel with programmatic serialization
  if (item >= P.item0[P.n_oct]) return;
  int oct = 0;
  while (item >= P.item0[oct + 1]) ++oct;
  const int e = item - P.item0[oct];
  const int nf = P.n_fft[oct], hop = P.hop[oct], n = P.n_sig[oct], F = P.n_frames, bpo = P.bpo;
  const int t = e < P.n_left[oct] ? e : P.t_right[oct] + (e - P.n_left[oct]);
  for (int i = lane; i < nf; i += 32) {
    const int idx = reflect_idx32(t * hop + i - nf / 2, n);
    float s;
    if (oct == 0) s = __ldg(y32 + (size_t)clip * y_stride + idx);
    else {
      const size_t o = (size_t)clip * P.stride[oct] + (size_t)level_index(idx, P.q[oct], P.rtot[oct], P.hb[oct]);
      s = __half2float(__ushort_as_half(P.hi[oct][o])) + __half2float(__ushort_as_half(P.lo[oct][o])) * (1.f / 2048.f);
    }
    win[wib][i] = s;
  }
  __syncwarp();
  const int nc = 2 * bpo;                  // output components of a frame
  float acc = 0.f;
  if (lane < nc) {
    const float* cj = P.coef[oct] + lane;  // [tap][filter](re, im)
#pragma unroll 8
    for (int i = 0; i < nf; ++i) acc = fmaf(__ldg(cj + (size_t)i * nc), win[wib][i], acc);
  }
  const float other = __shfl_xor_sync(0xffffffffu, acc, 1);
  if (lane < nc && !(lane & 1)) {
    const int bin = P.n_bins - bpo * (oct + 1) + (lane >> 1);
    out[((size_t)clip * P.n_bins + bin) * F + t] = logf(sqrtf(acc * acc + other * other) * __ldg(inv_sqrt_len + bin) + 1e-9f);
  }
}

// ---------------------------------------------------------------------------------------------
// host: forward
// ---------------------------------------------------------------------------------------------
static long long* g_vqt_dbg = nullptr;
// Diagnostic: device buffer of 32 int64 per level that CTA 0 of every level kernel fills with cycle counters:
// [4 i + 0..3] issuer i: total, waiting for operands, waiting for an accumulator stage, MMAs issued;
// [16],[17] epilogue (first half) total / waiting, [18],[19] second half, [20],[21] loader warp 0 total / waiting for its slot
extern "C" int zns_dbg_vqt_timing(long long* buf) { g_vqt_dbg = buf; return ZNS_OK; }

int vqt_umma_forward(zns_vqt_plan* p, const float* y, int batch, int n_samples, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int n_frames = zns_vqt_num_frames(n_samples, p->hop);
  static bool attr_set[64] = {};
  static int n_sm[64] = {};
  int dev = 0;
  ZNS_CHECK_CUDA(cudaGetDevice(&dev));
  ZNS_REQUIRE(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
  if (!attr_set[dev]) {
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(vqt_level_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(vqt_level_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(vqt_level_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(vqt_level_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(vqt_level_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(vqt_level_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    ZNS_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev));
    attr_set[dev] = true;
  }
  if (n_samples < p->umma_last_n) {
    // a shorter signal than last time: the level buffers must read as zero beyond its end
    for (int i = 1; i < p->n_oct; ++i) {
      const size_t bytes = (size_t)p->max_batch * p->sig_stride[i] * sizeof(uint16_t);
      ZNS_CHECK_CUDA(cudaMemsetAsync(p->d_hi[i], 0, bytes, st));
      ZNS_CHECK_CUDA(cudaMemsetAsync(p->d_lo[i], 0, bytes, st));
    }
  }
  p->umma_last_n = n_samples;
  static const bool pdl = !(getenv("ZNS_VQT_PDL") && atoi(getenv("ZNS_VQT_PDL")) == 0);
  cudaLaunchAttribute pdl_attr = {};
  pdl_attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr.val.programmaticStreamSerializationAllowed = 1;
  VqtEdgeParams E;
  memset(&E, 0, sizeof(E));
  E.n_oct = p->n_oct; E.bpo = p->bpo; E.n_bins = p->n_bins; E.n_frames = n_frames;
  int n_cur = n_samples;
  for (int i = 0; i < p->n_oct; ++i) {
    const VqtLevelDev& L = p->level[i];
    const int R = 8 * L.q;
    const int n_next = (n_cur + 1) / 2;
    int rows = (n_frames + L.fpr - 1) / L.fpr;
    if (L.dec_w > 0) rows = std::max(rows, (n_cur + R - 1) / R);
    const bool last = (i == p->n_oct - 1);
    VqtLevelArgs a;
    memset(&a, 0, sizeof(a));
    a.y32 = (i == 0) ? y : nullptr;
    a.src_hi = p->d_hi[i]; a.src_lo = p->d_lo[i];
    a.n_sig = n_cur;
    a.src_stride = (i == 0) ? (long long)n_samples : p->sig_stride[i];
    a.bimg = p->d_bimg[i];
    a.inv_sqrt_len = p->d_inv_sqrt_len;
    a.out = out; a.n_frames = n_frames; a.n_bins = p->n_bins;
    a.dst_hi = last ? nullptr : p->d_hi[i + 1];
    a.dst_lo = last ? nullptr : p->d_lo[i + 1];
    a.n_valid = n_cur / 2;
    a.dst_stride = last ? 0 : p->sig_stride[i + 1];
    a.dst_q = last ? 1 : p->level[i + 1].q;
    a.dst_rtot = last ? 0 : p->level[i + 1].rtot; a.dst_hb = last ? 0 : p->level[i + 1].hb; a.dst_ha = last ? 0 : p->level[i + 1].ha;
    a.dst_cap = last ? 0 : p->sig_cap[i + 1];
    a.tiles_per_clip = (rows + 127) / 128;
    a.n_tiles = a.tiles_per_clip * batch;
    a.dbg = g_vqt_dbg ? g_vqt_dbg + 32 * i : nullptr;
#ifdef ZNS_VQT_TIMING
    a.mode = getenv("ZNS_VQT_DBG_MODE") ? atoi(getenv("ZNS_VQT_DBG_MODE")) : 0;
#endif
    const int grid = std::min(a.n_tiles, n_sm[dev]);
    const size_t smem = level_smem(L);
    // levels >= 1 (and the edge-frame kernel) are programmatic dependents of the launch in front of them
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(VQT_LEVEL_THREADS(i == 0)); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = &pdl_attr; cfg.numAttrs = (pdl && i > 0) ? 1 : 0;
    // two epilogue groups need two stages in every accumulator ring in use and one decimator pass per tile (stage == tile parity)
    // (measured at cfg2, profiles/r02_vqt_epilogue_groups_ab.txt: multi-frame levels 0.440 -> 0.434 ms, one-frame levels -- which are
    // bound by their MMAs -- 0.440 -> 0.441: only the former by default)
    static const int grp_mask = getenv("ZNS_VQT_GROUPS") ? atoi(getenv("ZNS_VQT_GROUPS")) : 6;   // bit 0: one-frame levels, bit 1 / 2: two- / four-frame levels
    const bool can_group = i > 0 && L.ring_stages[0] == 2 && (L.n_jobs == 1 || (L.ring_stages[1] == 2 && L.n_pass == 1));
    const bool grouped = can_group && (grp_mask & (L.fpr == 1 ? 1 : (L.fpr == 2 ? 2 : 4)));
    if (i == 0 && L.fpr == 1) ZNS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, vqt_level_kernel<true, true, false>, L, a));
    else if (i == 0) ZNS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, vqt_level_kernel<true, false, false>, L, a));
    else if (L.fpr == 1 && grouped) ZNS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, vqt_level_kernel<false, true, true>, L, a));
    else if (L.fpr == 1) ZNS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, vqt_level_kernel<false, true, false>, L, a));
    else if (grouped) ZNS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, vqt_level_kernel<false, false, true>, L, a));
    else ZNS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, vqt_level_kernel<false, false, false>, L, a));
    // edge frames: left t*hop < nf/2 ; right t*hop + nf/2 > n
    const int nl = std::min(n_frames, (L.n_fft / 2 + L.hop - 1) / L.hop);
    int tr = (n_cur >= L.n_fft / 2) ? (n_cur - L.n_fft / 2) / L.hop + 1 : 0;
    tr = std::max(tr, nl);
    E.n_fft[i] = L.n_fft; E.hop[i] = L.hop; E.n_sig[i] = n_cur; E.stride[i] = p->sig_stride[i]; E.q[i] = L.q; E.rtot[i] = L.rtot; E.hb[i] = L.hb;
    E.coef[i] = p->d_coef[i]; E.hi[i] = p->d_hi[i]; E.lo[i] = p->d_lo[i];
    E.n_left[i] = nl; E.t_right[i] = tr;
    E.item0[i + 1] = E.item0[i] + (nl + std::max(0, n_frames - tr));          // one warp per edge frame
    n_cur = n_next;
  }
  const int n_items = E.item0[p->n_oct];
  if (n_items > 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((n_items + 7) / 8, 1, batch); cfg.blockDim = dim3(256); cfg.stream = st;
    cfg.attrs = &pdl_attr; cfg.numAttrs = pdl ? 1 : 0;
    ZNS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, vqt_edge_kernel, E, y, (long long)n_samples, (const float*)p->d_inv_sqrt_len, out));
  }
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
