// cv2..cv8 of the encoders as implicit GEMMs on tcgen05 / TMEM with TMA-staged operands.
//
// Reference semantics: nn.Conv2d "same" convolutions of _CNN
// (/root/reference/zeroNoteSamba/models/models.py:17-23,41-70), their data gradient (the same
// kernel on flipped/transposed packed weights) and their weight gradient.
//
// Layout trick (see common.cuh): activations are bf16 [G][H][W][8][C].  A TMA box
// (64 ch, 8 clips, w-span, 1, 1) lands in shared memory as one 1024-byte, 128B-swizzled atom per
// time frame w (8 clip rows x 64 channels).  Consequences:
//   * forward / dgrad: M = 128 output positions = 16 frames x 8 clips of one frequency row.  The
//     A operand of filter tap (r, s) is the SAME shared-memory row buffer offset by s atoms, so
//     one halo row (16 + kw - 1 atoms) is loaded once per 64-channel chunk and reused by all kw
//     taps and by up to HT output rows (HT accumulators in TMEM share every weight tile).
//     Zero "same" padding along W is TMA out-of-bounds fill; padding rows along H are skipped.
//   * wgrad: the reduction runs over positions; both operands are MN-major views of the same
//     kind of tile (x halo row shifted by the tap, dy tile), accumulators are per-tap
//     [c_in x c_out] blocks in TMEM, reduced across position slices with fp32 atomics.
//   * CTA pairs: the default kernels run as clusters of two CTAs on one TPC issuing M = 256
//     tcgen05.mma.cta_group::2 -- each CTA stages its own M half (input rows / x rows) and half of the
//     shared N operand (weight tile / dy tile), which takes the operand reads per SM below the
//     128 B/clk shared-memory limit of a single-CTA N = 128 MMA and halves the shared operand's L2 traffic.
//     ZNS_CONV_PAIR=0..3 selects how many kernel families use pairs (A/B switch, default 3 = all).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <queue>
#include <tuple>
#include <vector>

#include "common.cuh"

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// host: tensor-map encoding through the driver entry point (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int get_encode() {
  if (g_encode) return ZNS_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  ZNS_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return zns_set_error(ZNS_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  g_encode = (PFN_encodeTiled)fn;
  return ZNS_OK;
}

// act bf16 [G][H][W][8][C] -> 5-D map (C, 8, W, H, G), box (64, 8, wbox, 1, 1), 128B swizzle
static int make_act_map(CUtensorMap* m, const void* base, int G, int H, int W, int C, int wbox) {
  int rc = get_encode();
  if (rc) return rc;
  cuuint64_t dims[5] = {(cuuint64_t)C, 8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)G};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 16, (cuuint64_t)W * C * 16, (cuuint64_t)H * W * C * 16};
  cuuint32_t box[5] = {64, 8, (cuuint32_t)wbox, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return zns_set_error(ZNS_ERR_CUDA, "cuTensorMapEncodeTiled(act) failed: %d", (int)r);
  return ZNS_OK;
}

// packed weights bf16 [taps][rows][K] -> 3-D map (K, rows, taps), box (64, nrows, 1)
static int make_w_map(CUtensorMap* m, const void* base, int taps, int rows, int K, int nrows) {
  int rc = get_encode();
  if (rc) return rc;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)rows * K * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)nrows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return zns_set_error(ZNS_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return ZNS_OK;
}

#define ZNS_SMEM_LIMIT 232448  // 227 KB opt-in maximum per CTA
#define WT 16                  // frames per M tile (x 8 clips = 128 rows)
#define MAX_RING 8
#define ZNS_NUM_SMS 148
// Tensor-core conv kernels: warp 0 = TMA producer, warp 1 = MMA issuer, the rest = epilogue.  A warp can read
// only the TMEM lane quadrant (warp % 4), so ZNS_EPI epilogue warps share a quadrant and take its 32-column
// blocks round-robin (nothing else runs on the SM while the accumulators drain, so this time is not hidden).
#ifndef ZNS_EPI
#define ZNS_EPI 4   // measured on the training step: 1 -> 2 warps per quadrant +4.7 %, 2 -> 4 another +1.4 % (profiles/r01_epilogue_warps.txt)
#endif
#define ZNS_CONV_THREADS (64 + 128 * ZNS_EPI)

// ---------------------------------------------------------------------------------------------
// Tile plan.  Every conv CTA owns an SM (its shared memory does not leave room for a second), and the
// layers offer only a few hundred equal tiles, so with one tile size the last wave is mostly empty
// (640 tiles on 148 SMs = 4.3 waves).  A column (group, frame tile, branch) of `units` output rows is
// therefore cut into `nb` big tiles of `hb` rows followed by `ns` small tiles of `hs` rows; big tiles
// come first in block order, so the hardware's in-order dispatch behaves like longest-first list
// scheduling.  plan_tiles() picks (hb, nb, hs, ns) by simulating that schedule.
// ---------------------------------------------------------------------------------------------
struct TilePlan {
  int n_cols;   // G * n_wtiles * n_br
  int hb, nb, hs, ns;
  int n_big;    // n_cols * nb
  int n_total;  // n_cols * (nb + ns)
};

__device__ __forceinline__ void tile_decode(const TilePlan& tp, int n_wtiles, int G, int t, int& br, int& g, int& wt,
                                            int& u0, int& un) {
  int j, col;
  if (t < tp.n_big) {
    j = t / tp.n_cols; col = t - j * tp.n_cols; u0 = j * tp.hb; un = tp.hb;
  } else {
    t -= tp.n_big;
    j = t / tp.n_cols; col = t - j * tp.n_cols; u0 = tp.nb * tp.hb + j * tp.hs; un = tp.hs;
  }
  wt = col % n_wtiles; col /= n_wtiles;
  g = col % G;
  br = col / G;
}

// units: output rows (or stacked row pairs) per column; u_max: accumulators that fit TMEM / shared memory;
// ovh: per-CTA prologue + epilogue cost in units of one row's MMA time.
// ctas: CTAs sharing one weight tile (a pair loads half a tile each, so L2 feeds it twice as many tile rows).
static TilePlan plan_tiles(int units, int n_cols, int u_max, double ovh, int ctas) {
  // plans are pure functions of their arguments: cache them (the simulation costs ~1 ms of host time)
  static std::mutex mu;
  static std::map<std::tuple<int, int, int, long, int>, TilePlan> cache;
  const auto key = std::make_tuple(units, n_cols, u_max, lround(ovh * 4096.0), ctas);   // (before env scaling: env is process-wide)
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  // weight-tile bytes per MMA clock fall as 64/h B/clk; ZNS_PLAN_L2 (default 38 B/clk per SM, per CTA of a pair) is what L2 is
  // assumed to deliver, ZNS_PLAN_OVH scales the per-CTA overhead (both only for tuning experiments)
  static const double l2_env = getenv("ZNS_PLAN_L2") ? atof(getenv("ZNS_PLAN_L2")) : 0.0;
  const double l2_rate = l2_env > 0.0 ? l2_env : 38.0 * ctas;
  static const double ovh_scale = getenv("ZNS_PLAN_OVH") ? atof(getenv("ZNS_PLAN_OVH")) : 1.0;
  ovh *= ovh_scale;
  auto cost = [&](int h) {
    const double mult = std::max(1.0, (64.0 / h) / l2_rate);
    return h * mult + ovh;
  };
  auto simulate = [&](int hb, int nb, int hs, int ns) {
    std::priority_queue<double, std::vector<double>, std::greater<double>> sm;   // SM finish times, earliest first
    for (int i = 0; i < ZNS_NUM_SMS; ++i) sm.push(0.0);
    double last = 0.0;
    auto push = [&](double d) {
      const double t = sm.top() + d;
      sm.pop();
      sm.push(t);
      last = std::max(last, t);
    };
    for (long i = 0; i < (long)n_cols * nb; ++i) push(cost(hb));
    for (long i = 0; i < (long)n_cols * ns; ++i) push(cost(hs));
    return last;
  };
  TilePlan best;
  memset(&best, 0, sizeof(best));
  double best_t = 1e30;
  u_max = std::max(1, std::min(u_max, units));
  for (int hb = u_max; hb >= 1; --hb)
    for (int nb = units / hb; nb >= 0; --nb) {
      const int rem = units - nb * hb;
      for (int hs = std::min(hb, std::max(rem, 1)); hs >= 1; --hs) {
        if (rem == 0 && hs != hb) continue;
        if (rem > 0 && rem % hs != 0) continue;
        if (nb == 0 && hs != hb) continue;      // all-small plans are covered by a smaller hb
        const int ns = rem / hs;
        const double t = simulate(hb, nb, hs, ns);
        if (t < best_t * 0.995) {
          best_t = t;
          best.hb = hb; best.nb = nb; best.hs = hs; best.ns = ns;
        }
      }
    }
  best.n_cols = n_cols;
  best.n_big = n_cols * best.nb;
  best.n_total = n_cols * (best.nb + best.ns);
  {
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = best;
  }
  return best;
}

// Host-side launch geometry, reported instead of launching when a launcher is given a non-NULL `dry`
// (zns_dbg_conv_*_plan: lets the CPU tests check tile plans and work-item tables without a GPU).
struct PlanDump {
  int kernel;               // 0 = conv_fwd_umma, 1 = conv_fwd_stack_umma, 2 = conv_fwdT_umma, 3 = conv_wgrad_umma
  int n, ctas;              // MMA N (NB for the weight gradient), CTAs per cluster
  TilePlan tiles;           // forward kernels
  int n_slots, n_stages;    // forward: A-row ring / weight ring; weight gradient: -, stage ring
  int grid_x, grid_z;
  size_t smem;
  // weight gradient
  int n_slices, n_acc, n_sgroups, grp_base, grp_rem, n_rows, stack_dy, fold, n_cin_blocks, n_cout_blocks, n_item_pairs;
  uint32_t items[256];
};

// Optional outputs / epilogue modes of the forward kernels beyond the plain act store.
struct FwdExtra {
  void* const* out2;   // per branch: bf16 copy of the output (NULL: none)
  int pool;            // > 0: fused MaxPool2d((pool, 1)) -> ReLU -> Dropout; out / out2 are the pooled tensors
  void* const* arg;    // per branch: first-arg-max row of every pooled element (NULL: not wanted)
};

// Launch with clusters of two CTAs: blocks (2i, 2i+1) form a pair -- same rows and branch, adjacent frame
// tiles (callers check that the columns per branch are even, so a pair never straddles a row block or a branch).
template <typename Kern, typename... Args>
static cudaError_t launch_pair(Kern kern, dim3 grid, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(ZNS_CONV_THREADS, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// ---------------------------------------------------------------------------------------------
// forward / data-gradient kernel
// ---------------------------------------------------------------------------------------------
struct FwdParams {
  int G, H, W, batch;
  int kh, kw, ph, pw;
  int n_chunks;             // input channels / 64
  int n_wtiles;
  TilePlan tiles;           // rows per CTA (units = output rows)
  int n_slots, n_bstages;   // A row ring, B tile ring
  uint32_t slot_bytes;      // (WT + kw - 1) * 1024
  int relu;
  float drop_p, scale;
  uint32_t seed, stream_id;
  const uint32_t* seed_dev;
  const float* bias[2];
  const bf16* mask[2];
  bf16* out[2];
  bf16* out2[2];               // optional second copy of the output, always bf16 (x operand of the next weight gradient)
  int a_f16, b_f16, out_f16;   // element types: input activations / weights / output (0 = bf16, 1 = fp16)
  // fused MaxPool2d((pool, 1)) -> ReLU -> Dropout epilogue (models.py:50-53): every tile holds exactly one pool window
  // (`pool` accumulators = `pool` output rows); out / out2 are the POOLED tensors [G][H/pool][W][8][N], arg the row of the
  // first maximum inside each window (one byte per pooled element, the routing table of the backward pass)
  int pool;
  uint8_t* arg[2];
};

struct FwdBarriers {
  uint64_t a_full[MAX_RING], a_empty[MAX_RING], b_full[MAX_RING], b_empty[MAX_RING], acc_full;
  uint32_t tmem_base;
};

// CTAS = 2: two CTAs of a cluster (one TPC) work on adjacent frame tiles of the same rows as ONE
// M = 256 MMA (tcgen05 cta_group::2).  Each CTA stages its own input rows and HALF of every weight
// tile, so the operand reads per SM fall from 8 KB to 6 KB per 64-clock MMA at N = 128 -- below the
// 128 B/clk shared-memory limit that the single-CTA kernel sits on -- and the weight traffic from L2
// halves.  The even CTA (leader) issues the MMAs and owns the "full" barriers, which count the bytes
// of both CTAs' TMA loads; the "empty" and accumulator barriers are signalled in both CTAs by
// multicast commits.
template <int N, int HT, int CTAS>
__global__ void __launch_bounds__(ZNS_CONV_THREADS, 1)
conv_fwd_umma_kernel(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_in1,
                     const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1,
                     const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + p.n_slots * p.slot_bytes;
  constexpr uint32_t kBTile = N * 128 / CTAS;
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  FwdBarriers* bars = reinterpret_cast<FwdBarriers*>(smem_raw + (b_base + p.n_bstages * kBTile - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  int br, g, wt, h0, ht_eff;
  tile_decode(p.tiles, p.n_wtiles, p.G, blockIdx.x, br, g, wt, h0, ht_eff);
  const CUtensorMap* tm_in = br ? &tm_in1 : &tm_in0;
  const CUtensorMap* tm_w = br ? &tm_w1 : &tm_w0;
  const int w0 = wt * WT;
  // input rows rr (relative): hh = h0 - ph + rr, valid when 0 <= hh < H
  const int rr_lo = max(0, p.ph - h0);
  const int rr_hi = min(ht_eff + p.kh - 1, p.H + p.ph - h0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_slots; ++i) { mbar_init(smem_u32(&bars->a_full[i]), 1); mbar_init(smem_u32(&bars->a_empty[i]), 1); }
    for (int i = 0; i < p.n_bstages; ++i) { mbar_init(smem_u32(&bars->b_full[i]), 1); mbar_init(smem_u32(&bars->b_empty[i]), 1); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    mbar_fence_init();
    tma_prefetch_desc(tm_in);
    tma_prefetch_desc(tm_w);
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_pair(smem_u32(&bars->tmem_base), 512); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(&bars->tmem_base), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();   // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  const uint32_t bar_a_full = smem_u32(&bars->a_full[0]), bar_a_empty = smem_u32(&bars->a_empty[0]);
  const uint32_t bar_b_full = smem_u32(&bars->b_full[0]), bar_b_empty = smem_u32(&bars->b_empty[0]);
  const uint32_t n_slots = p.n_slots, n_bst = p.n_bstages;
  // Producer and MMA issuer are each ONE elected thread running the whole loop: tcgen05.mma is
  // asynchronous, so the issuing instruction stream only has to be shorter than the MMAs it feeds
  // (measured: ~80 clk of scalar work per MMA and ~330 clk per elect block in a naive loop, which is
  // why nothing below divides, takes a modulo or re-elects).
  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      uint32_t a_slot = 0, a_par = 1, b_st = 0, b_par = 1;
      for (int c = 0; c < p.n_chunks; ++c) {
        int next_row = rr_lo;
        for (int r = 0; r < p.kh; ++r) {
          const int need_hi = min(r + ht_eff, rr_hi);
          while (next_row < need_hi) {
            mbar_wait(bar_a_empty + 8 * a_slot, a_par);
            if (leader) mbar_expect_tx(bar_a_full + 8 * a_slot, CTAS * p.slot_bytes);
            if (CTAS == 2)
              tma_load_5d_pair(a_base + a_slot * p.slot_bytes, tm_in, (bar_a_full + 8 * a_slot) & ZNS_PEER_MASK, c * 64, 0,
                               w0 - p.pw, h0 - p.ph + next_row, g);
            else
              tma_load_5d(a_base + a_slot * p.slot_bytes, tm_in, bar_a_full + 8 * a_slot, c * 64, 0, w0 - p.pw,
                          h0 - p.ph + next_row, g);
            if (++a_slot == n_slots) { a_slot = 0; a_par ^= 1; }
            ++next_row;
          }
          if (max(r, rr_lo) >= need_hi) continue;  // tap row touches no valid input row
          const int tap0 = r * p.kw;
          for (int s = 0; s < p.kw; ++s) {
            mbar_wait(bar_b_empty + 8 * b_st, b_par);
            if (leader) mbar_expect_tx(bar_b_full + 8 * b_st, CTAS * kBTile);
            if (CTAS == 2)
              tma_load_3d_pair(b_base + b_st * kBTile, tm_w, (bar_b_full + 8 * b_st) & ZNS_PEER_MASK, c * 64,
                               (int)cta_rank * (N / 2), tap0 + s);
            else
              tma_load_3d(b_base + b_st * kBTile, tm_w, bar_b_full + 8 * b_st, c * 64, 0, tap0 + s);
            if (++b_st == n_bst) { b_st = 0; b_par ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc_16(128 * CTAS, N, 0, 0, p.a_f16, p.b_f16);
      auto commit = [](uint32_t bar) { if (CTAS == 2) umma_commit_pair(bar); else umma_commit(bar); };
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version 1, SWIZZLE_128B
      uint32_t a_slot = 0, a_par = 0, b_st = 0, b_par = 0;
      uint32_t rel_slot = 0;      // slot of the oldest unreleased row (rel_row)
      uint32_t started = 0;       // bit h: accumulator h holds data
      for (int c = 0; c < p.n_chunks; ++c) {
        int next_row = rr_lo, rel_row = rr_lo;
        for (int r = 0; r < p.kh; ++r) {
          const int need_hi = min(r + ht_eff, rr_hi);
          while (next_row < need_hi) {
            mbar_wait(bar_a_full + 8 * a_slot, a_par);
            if (++a_slot == n_slots) { a_slot = 0; a_par ^= 1; }
            ++next_row;
          }
          const int lo = max(r, rr_lo), hi = need_hi;
          if (lo < hi) {
            // rows lo..hi-1 are live, rel_row <= lo: slot(lo) = rel_slot + (lo - rel_row) (mod n_slots)
            uint32_t lo_slot = rel_slot + (uint32_t)(lo - rel_row);
            if (lo_slot >= n_slots) lo_slot -= n_slots;
            // descriptors carry the 18-bit offset inside the CTA's shared window (the same in both CTAs of a pair)
            const uint32_t lo_addr = (a_base & 0x3FFFFu) + lo_slot * p.slot_bytes;
            const uint32_t wrap_addr = (a_base & 0x3FFFFu) + n_slots * p.slot_bytes;
            for (int s = 0; s < p.kw; ++s) {
              mbar_wait(bar_b_full + 8 * b_st, b_par);
              tc_fence_after();
              const uint32_t b_lo = ((((b_base & 0x3FFFFu) + b_st * kBTile)) >> 4) | (1u << 16);
              uint32_t row_addr = lo_addr + s * 1024;
              for (int rr = lo; rr < hi; ++rr) {
                const int h = rr - r;
                const uint32_t a_lo = (row_addr >> 4) | (1u << 16);
                const uint32_t acc = ((started >> h) & 1) | (s > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (CTAS == 2)
                    umma_bf16_pair(tmem + h * N, ((uint64_t)kDescHi << 32) | (a_lo + 2 * k),
                                   ((uint64_t)kDescHi << 32) | (b_lo + 2 * k), idesc, acc | (k > 0));
                  else
                    umma_bf16(tmem + h * N, ((uint64_t)kDescHi << 32) | (a_lo + 2 * k),
                              ((uint64_t)kDescHi << 32) | (b_lo + 2 * k), idesc, acc | (k > 0));
                }
                row_addr += p.slot_bytes;
                if (row_addr >= wrap_addr) row_addr -= n_slots * p.slot_bytes;
              }
              commit(bar_b_empty + 8 * b_st);
              if (++b_st == n_bst) { b_st = 0; b_par ^= 1; }
            }
            started |= ((1u << (hi - lo)) - 1u) << (lo - r);
          }
          while (rel_row <= r && rel_row < rr_hi) {   // row r has had its last use
            commit(bar_a_empty + 8 * rel_slot);
            if (++rel_slot == n_slots) rel_slot = 0;
            ++rel_row;
          }
        }
        while (rel_row < rr_hi) {
          commit(bar_a_empty + 8 * rel_slot);
          if (++rel_slot == n_slots) rel_slot = 0;
          ++rel_row;
        }
      }
      commit(smem_u32(&bars->acc_full));
    }
    __syncwarp();
  } else {
    // ===== epilogue: TMEM -> registers -> bias / ReLU / dropout / mask -> bf16 act =====
    mbar_wait(smem_u32(&bars->acc_full), 0);
    tc_fence_after();
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int w = w0 + (m >> 3), b8 = m & 7;
    const bool valid = (w < p.W);
    const float* bias = p.bias[br];
    const bf16* mask = p.mask[br];
    bf16* out = p.out[br];
    bf16* out2 = p.out2[br];
    uint32_t seed = p.seed;
    if (p.seed_dev) seed ^= __ldg(p.seed_dev) * 0x9E3779B9u;
    const bool do_drop = p.drop_p > 0.f;
    const double thr_d = (double)p.drop_p * 4294967296.0;
    const uint32_t thr = thr_d >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)thr_d;
    const float keep = do_drop ? 1.f / (1.f - p.drop_p) : 1.f;
    if (p.pool) {
      // pooled epilogue: the warps of a quadrant take the 32-column blocks; a thread keeps a running maximum (and the row of
      // the first maximum) over the window's accumulators, then bias -> ReLU -> dropout -> fp16 / bf16 stores of the POOLED
      // tensor.  max and "+ bias" commute (the bias is per channel), rounding is monotone, so this is the unfused result.
      const int hp = h0 / p.pool, Hp = p.H / p.pool;
      const size_t e0 = zns_act_index(g, hp, valid ? w : 0, b8, 0, Hp, p.W, N);
      uint8_t* argp = p.arg[br];
#pragma unroll 1
      for (int nb = half; nb < N / 32; nb += ZNS_EPI) {
        uint32_t v[32];
        float m[32];
        uint32_t am[8] = {0, 0, 0, 0, 0, 0, 0, 0};         // four 8-bit row indices per word
        tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + nb * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) m[j] = __uint_as_float(v[j]);
        for (int h = 1; h < p.pool; ++h) {
          tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + h * N + nb * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(v[j]);
            if (x > m[j]) { m[j] = x; am[j >> 2] = (am[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | ((uint32_t)h << (8 * (j & 3))); }
          }
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = m[j];
            if (bias) x += __ldg(bias + nb * 32 + j);
            x = fmaxf(x, 0.f);
            if (do_drop) x = (zns_hash32(e0 + nb * 32 + j, seed, p.stream_id + br) >= thr) ? x * keep : 0.f;
            m[j] = x;
          }
          uint4* dst = reinterpret_cast<uint4*>(out + e0 + nb * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(pack_act2(m[q * 8], m[q * 8 + 1], p.out_f16), pack_act2(m[q * 8 + 2], m[q * 8 + 3], p.out_f16),
                                pack_act2(m[q * 8 + 4], m[q * 8 + 5], p.out_f16), pack_act2(m[q * 8 + 6], m[q * 8 + 7], p.out_f16));
          if (out2) {
            uint4* dst2 = reinterpret_cast<uint4*>(out2 + e0 + nb * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              dst2[q] = make_uint4(pack_bf16x2(m[q * 8], m[q * 8 + 1]), pack_bf16x2(m[q * 8 + 2], m[q * 8 + 3]),
                                   pack_bf16x2(m[q * 8 + 4], m[q * 8 + 5]), pack_bf16x2(m[q * 8 + 6], m[q * 8 + 7]));
          }
          if (argp) {
            uint4* da = reinterpret_cast<uint4*>(argp + e0 + nb * 32);
            da[0] = make_uint4(am[0], am[1], am[2], am[3]);
            da[1] = make_uint4(am[4], am[5], am[6], am[7]);
          }
        }
      }
    }
    int nb = half;   // blocks of all rows are dealt round-robin: (h, nb) -> warp (h * N/32 + nb) % ZNS_EPI of the quadrant
    for (int h = 0; h < (p.pool ? 0 : ht_eff); ++h, nb -= N / 32) {
      const size_t e0 = zns_act_index(g, h0 + h, valid ? w : 0, b8, 0, p.H, p.W, N);
#pragma unroll 1
      for (; nb < N / 32; nb += ZNS_EPI) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + h * N + nb * 32, v);
        tmem_ld_wait();
        if (valid) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(v[j]);
            if (bias) x += __ldg(bias + nb * 32 + j);
            if (p.relu) x = fmaxf(x, 0.f);
            if (do_drop) x = (zns_hash32(e0 + nb * 32 + j, seed, p.stream_id + br) >= thr) ? x * keep : 0.f;
            f[j] = x;
          }
          if (mask) {
            const uint4* mp = reinterpret_cast<const uint4*>(mask + e0 + nb * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 mv = __ldg(mp + q);
              const uint32_t* m2 = reinterpret_cast<const uint32_t*>(&mv);   // fp16 or bf16 forward activations: sign / zero test
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (!act_bits_positive(m2[j] & 0xFFFFu)) f[q * 8 + 2 * j] = 0.f;
                if (!act_bits_positive(m2[j] >> 16)) f[q * 8 + 2 * j + 1] = 0.f;
              }
            }
          }
          uint4* dst = reinterpret_cast<uint4*>(out + e0 + nb * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(pack_act2(f[q * 8] * p.scale, f[q * 8 + 1] * p.scale, p.out_f16),
                                pack_act2(f[q * 8 + 2] * p.scale, f[q * 8 + 3] * p.scale, p.out_f16),
                                pack_act2(f[q * 8 + 4] * p.scale, f[q * 8 + 5] * p.scale, p.out_f16),
                                pack_act2(f[q * 8 + 6] * p.scale, f[q * 8 + 7] * p.scale, p.out_f16));
          if (out2) {
            uint4* dst2 = reinterpret_cast<uint4*>(out2 + e0 + nb * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              dst2[q] = make_uint4(pack_bf16x2(f[q * 8] * p.scale, f[q * 8 + 1] * p.scale),
                                   pack_bf16x2(f[q * 8 + 2] * p.scale, f[q * 8 + 3] * p.scale),
                                   pack_bf16x2(f[q * 8 + 4] * p.scale, f[q * 8 + 5] * p.scale),
                                   pack_bf16x2(f[q * 8 + 6] * p.scale, f[q * 8 + 7] * p.scale));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the pair's MMAs / signals are in flight
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// transposed forward / data-gradient kernel for C_out <= 128
//   The 128 x N x 16 MMA reads a 4 KB A tile however small N is, and shared-memory operand
//   bandwidth (~90 B/clk measured) bounds the N = 64 / 128 kernels above at 38 % / 63 % of peak.
//   Here the roles are swapped: A (M = 128) = weight tile, B (N = 256) = 32 frames x 8 clips of an
//   input row -> 12 KB per 128-cycle MMA, like the N = 256 kernel.  D[c_out, position] lives in
//   TMEM with channels on lanes; the epilogue writes 2-byte values, 32 consecutive channels per
//   warp store.  For c_out = 64 two adjacent output rows are stacked on M: rows (h, h-1) use
//   weight tap rows (r, r+1) on the same input row, so the tap-row loop has kh + 1 steps.
// ---------------------------------------------------------------------------------------------
#define WTT 32  // frames per tile in the transposed kernel

struct FwdTParams {
  int G, H, W, batch;
  int kh, kw, ph, pw;
  int n_chunks;
  int n_wtiles, n_htiles;
  TilePlan tiles;    // stacked kernel only: accumulators (row pairs) per CTA
  int cout;          // 64 or 128
  int stack;         // output rows stacked on M: 128 / cout
  int n_acc;         // accumulators (of 256 TMEM columns) per CTA
  int n_slots, n_wstages;
  uint32_t slot_bytes;  // (WTT + kw - 1) * 1024
  int relu;
  float drop_p, scale;
  uint32_t seed, stream_id;
  const uint32_t* seed_dev;
  const float* bias[2];
  const bf16* mask[2];
  bf16* out[2];
  int a_f16, b_f16, out_f16;   // element types: input activations / weights / output (0 = bf16, 1 = fp16)
  // stacked kernel only: fused MaxPool2d((pool, 1)) -> ReLU -> Dropout epilogue (models.py:41-44); a tile holds whole pool
  // windows (2 * n_acc rows, a multiple of pool); out / out2 are the pooled tensors, arg the first-arg-max row per element
  bf16* out2[2];
  int pool;
  uint8_t* arg[2];
};

struct FwdTBarriers {
  uint64_t x_full[MAX_RING], x_empty[MAX_RING], w_full[MAX_RING], w_empty[MAX_RING], acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 1)
conv_fwdT_umma_kernel(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_in1,
                      const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1,
                      const FwdTParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_base = smem_base;
  const uint32_t w_base = x_base + p.n_slots * p.slot_bytes;
  constexpr uint32_t kWTile = 128 * 128;  // 128 rows x 64 channels bf16
  FwdTBarriers* bars = reinterpret_cast<FwdTBarriers*>(smem_raw + (w_base + p.n_wstages * kWTile - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int br = blockIdx.z;
  const CUtensorMap* tm_in = br ? &tm_in1 : &tm_in0;
  const CUtensorMap* tm_w = br ? &tm_w1 : &tm_w0;

  int t = blockIdx.x;
  const int wt = t % p.n_wtiles; t /= p.n_wtiles;
  const int htile = t % p.n_htiles; t /= p.n_htiles;
  const int g = t;
  const int rows_per_cta = p.n_acc * p.stack;
  const int w0 = wt * WTT, h0 = htile * rows_per_cta;
  const int acc_eff = min(p.n_acc, (p.H - h0 + p.stack - 1) / p.stack);
  const int n_iter = p.kh + p.stack - 1;                     // tap-row steps r' = 0 .. kh + stack - 2
  // relative input rows rr: hh = h0 - ph + rr; accumulator a needs row r' + a*stack at step r'
  const int rr_lo = max(0, p.ph - h0);
  const int rr_hi = min(acc_eff * p.stack + p.kh - 1, p.H + p.ph - h0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_slots; ++i) { mbar_init(smem_u32(&bars->x_full[i]), 1); mbar_init(smem_u32(&bars->x_empty[i]), 1); }
    for (int i = 0; i < p.n_wstages; ++i) { mbar_init(smem_u32(&bars->w_full[i]), 1); mbar_init(smem_u32(&bars->w_empty[i]), 1); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    mbar_fence_init();
    tma_prefetch_desc(tm_in);
    tma_prefetch_desc(tm_w);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // does step r' touch a valid input row for some accumulator?
  auto step_active = [&](int rp) {
    for (int a = 0; a < acc_eff; ++a) {
      const int rr = rp + a * p.stack;
      if (rr >= rr_lo && rr < rr_hi) return true;
    }
    return false;
  };

  const uint32_t bar_x_full = smem_u32(&bars->x_full[0]), bar_x_empty = smem_u32(&bars->x_empty[0]);
  const uint32_t bar_w_full = smem_u32(&bars->w_full[0]), bar_w_empty = smem_u32(&bars->w_empty[0]);
  const uint32_t n_slots = p.n_slots, n_wst = p.n_wstages;
  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer (one elected thread, no div/mod) =====
      uint32_t x_slot = 0, x_par = 1, w_st = 0, w_par = 1;
      for (int c = 0; c < p.n_chunks; ++c) {
        int next_row = rr_lo;
        for (int rp = 0; rp < n_iter; ++rp) {
          const int need_hi = min(rp + (acc_eff - 1) * p.stack + 1, rr_hi);
          while (next_row < need_hi) {
            mbar_wait(bar_x_empty + 8 * x_slot, x_par);
            mbar_expect_tx(bar_x_full + 8 * x_slot, p.slot_bytes);
            tma_load_5d(x_base + x_slot * p.slot_bytes, tm_in, bar_x_full + 8 * x_slot, c * 64, 0, w0 - p.pw,
                        h0 - p.ph + next_row, g);
            if (++x_slot == n_slots) { x_slot = 0; x_par ^= 1; }
            ++next_row;
          }
          if (!step_active(rp)) continue;
          for (int s = 0; s < p.kw; ++s) {
            mbar_wait(bar_w_empty + 8 * w_st, w_par);
            mbar_expect_tx(bar_w_full + 8 * w_st, kWTile);
            for (int j = 0; j < p.stack; ++j) {
              const int r = rp - (p.stack - 1) + j;      // weight tap row of M block j (out of range -> TMA zero fill)
              const int tap = (r < 0 || r >= p.kh) ? -1 : r * p.kw + s;
              tma_load_3d(w_base + w_st * kWTile + j * p.cout * 128, tm_w, bar_w_full + 8 * w_st, c * 64, 0, tap);
            }
            if (++w_st == n_wst) { w_st = 0; w_par ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      uint32_t x_slot = 0, x_par = 0, w_st = 0, w_par = 0;
      uint32_t rel_slot = 0;
      uint32_t started = 0;
      const uint32_t ring_bytes = n_slots * p.slot_bytes;
      for (int c = 0; c < p.n_chunks; ++c) {
        int next_row = rr_lo, rel_row = rr_lo;
        for (int rp = 0; rp < n_iter; ++rp) {
          const int need_hi = min(rp + (acc_eff - 1) * p.stack + 1, rr_hi);
          while (next_row < need_hi) {
            mbar_wait(bar_x_full + 8 * x_slot, x_par);
            if (++x_slot == n_slots) { x_slot = 0; x_par ^= 1; }
            ++next_row;
          }
          if (step_active(rp)) {
            // live rows are >= rel_row and fewer than n_slots: address of row rr by offset from rel_slot
            uint32_t row_off[2];
            uint32_t use = 0;
            for (int a = 0; a < acc_eff; ++a) {
              const int rr = rp + a * p.stack;
              if (rr >= rr_lo && rr < rr_hi) {
                uint32_t off = (rel_slot + (uint32_t)(rr - rel_row)) * p.slot_bytes;
                if (off >= ring_bytes) off -= ring_bytes;
                row_off[a] = off;
                use |= 1u << a;
              }
            }
            for (int s = 0; s < p.kw; ++s) {
              mbar_wait(bar_w_full + 8 * w_st, w_par);
              tc_fence_after();
              const uint32_t a_lo = ((w_base + w_st * kWTile) >> 4) | (1u << 16);
#pragma unroll
              for (int a = 0; a < 2; ++a) {
                if ((use >> a) & 1) {
                  const uint32_t b_lo = ((x_base + row_off[a] + s * 1024) >> 4) | (1u << 16);
                  const uint32_t acc = ((started >> a) & 1) | (s > 0);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    umma_bf16(tmem + a * 256, ((uint64_t)kDescHi << 32) | (a_lo + 2 * k), ((uint64_t)kDescHi << 32) | (b_lo + 2 * k),
                              idesc, acc | (k > 0));
                  }
                }
              }
              umma_commit(bar_w_empty + 8 * w_st);
              if (++w_st == n_wst) { w_st = 0; w_par ^= 1; }
            }
            started |= use;
          }
          while (rel_row <= rp && rel_row < rr_hi) {   // row rp has had its last use (accumulator 0)
            umma_commit(bar_x_empty + 8 * rel_slot);
            if (++rel_slot == n_slots) rel_slot = 0;
            ++rel_row;
          }
        }
        while (rel_row < rr_hi) {
          umma_commit(bar_x_empty + 8 * rel_slot);
          if (++rel_slot == n_slots) rel_slot = 0;
          ++rel_row;
        }
      }
      umma_commit(smem_u32(&bars->acc_full));
    }
    __syncwarp();
  } else {
    // ===== epilogue: lane = output channel (and stacked row), columns = positions =====
    mbar_wait(smem_u32(&bars->acc_full), 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int j = m / p.cout, co = m - j * p.cout;
    const float bias = p.bias[br] ? __ldg(p.bias[br] + co) : 0.f;
    const bf16* mask = p.mask[br];
    bf16* out = p.out[br];
    uint32_t seed = p.seed;
    if (p.seed_dev) seed ^= __ldg(p.seed_dev) * 0x9E3779B9u;
    const bool do_drop = p.drop_p > 0.f;
    const double thr_d = (double)p.drop_p * 4294967296.0;
    const uint32_t thr = thr_d >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)thr_d;
    const float keep = do_drop ? 1.f / (1.f - p.drop_p) : 1.f;
    const int n_valid_cols = min(WTT, p.W - w0) * 8;
    for (int a = 0; a < acc_eff; ++a) {
      const int h = h0 + a * p.stack + (p.stack - 1 - j);
      const bool row_ok = h < p.H;
      const size_t e0 = zns_act_index(g, row_ok ? h : 0, w0, 0, co, p.H, p.W, p.cout);
#pragma unroll 1
      for (int nb = 0; nb < 8; ++nb) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + a * 256 + nb * 32, v);
        tmem_ld_wait();
        if (!row_ok) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int n = nb * 32 + i;
          if (n < n_valid_cols) {
            const size_t e = e0 + (size_t)n * p.cout;
            float x = __uint_as_float(v[i]) + bias;
            if (p.relu) x = fmaxf(x, 0.f);
            if (do_drop) x = (zns_hash32(e, seed, p.stream_id + br) >= thr) ? x * keep : 0.f;
            if (mask && !(__bfloat162float(mask[e]) > 0.f)) x = 0.f;
            out[e] = __float2bfloat16(x * p.scale);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// C_out = 64 forward / data gradient with two output rows stacked on N.
//   A 128 x 64 x 16 MMA is bound by its 4 KB A-tile read (48 clk instead of 32, measured), so the
//   N = 64 kernel cannot pass 67 % of peak.  Here the weight tile is [W(r-1,s) ; W(r,s)] (N = 128):
//   on input row hh one MMA feeds output row h+1 with tap row r-1 and output row h with tap row r,
//   i.e. an accumulator holds two adjacent output rows and the tap-row loop has kh + 1 steps (tap
//   rows -1 and kh are TMA out-of-bounds zero fill).  Three accumulators = six output rows per CTA.
// ---------------------------------------------------------------------------------------------
// CTAS = 2: CTA pair as in conv_fwd_umma_kernel; the leader stages the upper weight block (tap row r-1), its
// peer the lower one (tap row r) -- together the N = 128 tile of one M = 256 MMA.
template <int CTAS>
__global__ void __launch_bounds__(ZNS_CONV_THREADS, 1)
conv_fwd_stack_umma_kernel(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_in1,
                      const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1,
                      const FwdTParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_base = smem_base;
  const uint32_t w_base = x_base + p.n_slots * p.slot_bytes;
  constexpr uint32_t kWTile = 128 * 128 / CTAS;  // 2 stacked tap rows x 64 output channels x 64 input channels bf16 (one tap row per CTA of a pair)
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  FwdTBarriers* bars = reinterpret_cast<FwdTBarriers*>(smem_raw + (w_base + p.n_wstages * kWTile - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int br, g, wt, u0, acc_eff;                       // tile = acc_eff stacked row pairs starting at pair u0
  tile_decode(p.tiles, p.n_wtiles, p.G, blockIdx.x, br, g, wt, u0, acc_eff);
  const CUtensorMap* tm_in = br ? &tm_in1 : &tm_in0;
  const CUtensorMap* tm_w = br ? &tm_w1 : &tm_w0;
  const int w0 = wt * WT, h0 = u0 * p.stack;
  const int n_iter = p.kh + p.stack - 1;                     // tap-row steps r' = 0 .. kh + stack - 2
  // relative input rows rr: hh = h0 - ph + rr; accumulator a needs row r' + a*stack at step r'
  const int rr_lo = max(0, p.ph - h0);
  const int rr_hi = min(acc_eff * p.stack + p.kh - 1, p.H + p.ph - h0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_slots; ++i) { mbar_init(smem_u32(&bars->x_full[i]), 1); mbar_init(smem_u32(&bars->x_empty[i]), 1); }
    for (int i = 0; i < p.n_wstages; ++i) { mbar_init(smem_u32(&bars->w_full[i]), 1); mbar_init(smem_u32(&bars->w_empty[i]), 1); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    mbar_fence_init();
    tma_prefetch_desc(tm_in);
    tma_prefetch_desc(tm_w);
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_pair(smem_u32(&bars->tmem_base), 512); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(&bars->tmem_base), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // does step r' touch a valid input row for some accumulator?
  auto step_active = [&](int rp) {
    for (int a = 0; a < acc_eff; ++a) {
      const int rr = rp + a * p.stack;
      if (rr >= rr_lo && rr < rr_hi) return true;
    }
    return false;
  };

  const uint32_t bar_x_full = smem_u32(&bars->x_full[0]), bar_x_empty = smem_u32(&bars->x_empty[0]);
  const uint32_t bar_w_full = smem_u32(&bars->w_full[0]), bar_w_empty = smem_u32(&bars->w_empty[0]);
  const uint32_t n_slots = p.n_slots, n_wst = p.n_wstages;
  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer (one elected thread, no div/mod) =====
      uint32_t x_slot = 0, x_par = 1, w_st = 0, w_par = 1;
      for (int c = 0; c < p.n_chunks; ++c) {
        int next_row = rr_lo;
        for (int rp = 0; rp < n_iter; ++rp) {
          const int need_hi = min(rp + (acc_eff - 1) * p.stack + 1, rr_hi);
          while (next_row < need_hi) {
            mbar_wait(bar_x_empty + 8 * x_slot, x_par);
            if (leader) mbar_expect_tx(bar_x_full + 8 * x_slot, CTAS * p.slot_bytes);
            if (CTAS == 2)
              tma_load_5d_pair(x_base + x_slot * p.slot_bytes, tm_in, (bar_x_full + 8 * x_slot) & ZNS_PEER_MASK, c * 64, 0,
                               w0 - p.pw, h0 - p.ph + next_row, g);
            else
              tma_load_5d(x_base + x_slot * p.slot_bytes, tm_in, bar_x_full + 8 * x_slot, c * 64, 0, w0 - p.pw,
                          h0 - p.ph + next_row, g);
            if (++x_slot == n_slots) { x_slot = 0; x_par ^= 1; }
            ++next_row;
          }
          if (!step_active(rp)) continue;
          for (int s = 0; s < p.kw; ++s) {
            mbar_wait(bar_w_empty + 8 * w_st, w_par);
            if (leader) mbar_expect_tx(bar_w_full + 8 * w_st, CTAS * kWTile);
            if (CTAS == 2) {
              const int r = rp - 1 + (int)cta_rank;      // this CTA's half of the stacked tile
              const int tap = (r < 0 || r >= p.kh) ? -1 : r * p.kw + s;
              tma_load_3d_pair(w_base + w_st * kWTile, tm_w, (bar_w_full + 8 * w_st) & ZNS_PEER_MASK, c * 64, 0, tap);
            } else {
              for (int j = 0; j < p.stack; ++j) {
                const int r = rp - (p.stack - 1) + j;      // weight tap row of M block j (out of range -> TMA zero fill)
                const int tap = (r < 0 || r >= p.kh) ? -1 : r * p.kw + s;
                tma_load_3d(w_base + w_st * kWTile + j * p.cout * 128, tm_w, bar_w_full + 8 * w_st, c * 64, 0, tap);
              }
            }
            if (++w_st == n_wst) { w_st = 0; w_par ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc_16(128 * CTAS, 128, 0, 0, p.a_f16, p.b_f16);
      auto commit = [](uint32_t bar) { if (CTAS == 2) umma_commit_pair(bar); else umma_commit(bar); };
      const uint32_t x_dbase = x_base & 0x3FFFFu, w_dbase = w_base & 0x3FFFFu;   // descriptor offsets (same in both CTAs)
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      uint32_t x_slot = 0, x_par = 0, w_st = 0, w_par = 0;
      uint32_t rel_slot = 0;
      uint32_t started = 0;
      const uint32_t ring_bytes = n_slots * p.slot_bytes;
      for (int c = 0; c < p.n_chunks; ++c) {
        int next_row = rr_lo, rel_row = rr_lo;
        for (int rp = 0; rp < n_iter; ++rp) {
          const int need_hi = min(rp + (acc_eff - 1) * p.stack + 1, rr_hi);
          while (next_row < need_hi) {
            mbar_wait(bar_x_full + 8 * x_slot, x_par);
            if (++x_slot == n_slots) { x_slot = 0; x_par ^= 1; }
            ++next_row;
          }
          if (step_active(rp)) {
            // live rows are >= rel_row and fewer than n_slots: address of row rr by offset from rel_slot
            uint32_t row_off[4];
            uint32_t use = 0;
            for (int a = 0; a < acc_eff; ++a) {
              const int rr = rp + a * p.stack;
              if (rr >= rr_lo && rr < rr_hi) {
                uint32_t off = (rel_slot + (uint32_t)(rr - rel_row)) * p.slot_bytes;
                if (off >= ring_bytes) off -= ring_bytes;
                row_off[a] = off;
                use |= 1u << a;
              }
            }
            for (int s = 0; s < p.kw; ++s) {
              mbar_wait(bar_w_full + 8 * w_st, w_par);
              tc_fence_after();
              const uint32_t b_lo = ((w_dbase + w_st * kWTile) >> 4) | (1u << 16);
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                if ((use >> a) & 1) {
                  const uint32_t a_lo = ((x_dbase + row_off[a] + s * 1024) >> 4) | (1u << 16);
                  const uint32_t acc = ((started >> a) & 1) | (s > 0);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    if (CTAS == 2)
                      umma_bf16_pair(tmem + a * 128, ((uint64_t)kDescHi << 32) | (a_lo + 2 * k),
                                     ((uint64_t)kDescHi << 32) | (b_lo + 2 * k), idesc, acc | (k > 0));
                    else
                      umma_bf16(tmem + a * 128, ((uint64_t)kDescHi << 32) | (a_lo + 2 * k),
                                ((uint64_t)kDescHi << 32) | (b_lo + 2 * k), idesc, acc | (k > 0));
                  }
                }
              }
              commit(bar_w_empty + 8 * w_st);
              if (++w_st == n_wst) { w_st = 0; w_par ^= 1; }
            }
            started |= use;
          }
          while (rel_row <= rp && rel_row < rr_hi) {   // row rp has had its last use (accumulator 0)
            commit(bar_x_empty + 8 * rel_slot);
            if (++rel_slot == n_slots) rel_slot = 0;
            ++rel_row;
          }
        }
        while (rel_row < rr_hi) {
          commit(bar_x_empty + 8 * rel_slot);
          if (++rel_slot == n_slots) rel_slot = 0;
          ++rel_row;
        }
      }
      commit(smem_u32(&bars->acc_full));
    }
    __syncwarp();
  } else {
    // ===== epilogue: thread = position row; accumulator a holds rows (h0+2a+1 | h0+2a) x 64 channels =====
    mbar_wait(smem_u32(&bars->acc_full), 0);
    tc_fence_after();
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int w = w0 + (m >> 3), b8 = m & 7;
    const bool valid = (w < p.W);
    const float* bias = p.bias[br];
    const bf16* mask = p.mask[br];
    bf16* out = p.out[br];
    uint32_t seed = p.seed;
    if (p.seed_dev) seed ^= __ldg(p.seed_dev) * 0x9E3779B9u;
    const bool do_drop = p.drop_p > 0.f;
    const double thr_d = (double)p.drop_p * 4294967296.0;
    const uint32_t thr = thr_d >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)thr_d;
    const float keep = do_drop ? 1.f / (1.f - p.drop_p) : 1.f;
    if (p.pool) {
      // pooled epilogue: work items = (pool window, 32-channel block); row r of the tile lives in accumulator r / 2, columns
      // [0, 64) when r is odd (upper row of the stacked pair) and [64, 128) when it is even
      const int n_win = 2 * acc_eff / p.pool, Hp = p.H / p.pool;
      bf16* out2 = p.out2[br];
      uint8_t* argp = p.arg[br];
#pragma unroll 1
      for (int item = half; item < 2 * n_win; item += ZNS_EPI) {
        const int wi = item >> 1, c0 = (item & 1) * 32;
        uint32_t v[32];
        float mx[32];
        uint32_t am[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < p.pool; ++k) {
          const int r = wi * p.pool + k;
          tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + (r >> 1) * 128 + ((r & 1) ? 0 : 64) + c0, v);
          tmem_ld_wait();
          if (k == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx[i] = __uint_as_float(v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = __uint_as_float(v[i]);
              if (x > mx[i]) { mx[i] = x; am[i >> 2] = (am[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | ((uint32_t)k << (8 * (i & 3))); }
            }
          }
        }
        if (valid) {
          const size_t e0 = zns_act_index(g, h0 / p.pool + wi, w, b8, c0, Hp, p.W, 64);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = mx[i];
            if (bias) x += __ldg(bias + c0 + i);
            x = fmaxf(x, 0.f);
            if (do_drop) x = (zns_hash32(e0 + i, seed, p.stream_id + br) >= thr) ? x * keep : 0.f;
            mx[i] = x;
          }
          uint4* dst = reinterpret_cast<uint4*>(out + e0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(pack_act2(mx[q * 8], mx[q * 8 + 1], p.out_f16), pack_act2(mx[q * 8 + 2], mx[q * 8 + 3], p.out_f16),
                                pack_act2(mx[q * 8 + 4], mx[q * 8 + 5], p.out_f16), pack_act2(mx[q * 8 + 6], mx[q * 8 + 7], p.out_f16));
          if (out2) {
            uint4* dst2 = reinterpret_cast<uint4*>(out2 + e0);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              dst2[q] = make_uint4(pack_bf16x2(mx[q * 8], mx[q * 8 + 1]), pack_bf16x2(mx[q * 8 + 2], mx[q * 8 + 3]),
                                   pack_bf16x2(mx[q * 8 + 4], mx[q * 8 + 5]), pack_bf16x2(mx[q * 8 + 6], mx[q * 8 + 7]));
          }
          if (argp) {
            uint4* da = reinterpret_cast<uint4*>(argp + e0);
            da[0] = make_uint4(am[0], am[1], am[2], am[3]);
            da[1] = make_uint4(am[4], am[5], am[6], am[7]);
          }
        }
      }
    }
    int nb = half;
    for (int a = 0; a < (p.pool ? 0 : acc_eff); ++a, nb -= 4) {
#pragma unroll 1
      for (; nb < 4; nb += ZNS_EPI) {
        const int j = nb >> 1;                              // stacked block: 0 -> upper row, 1 -> lower row
        const int h = h0 + 2 * a + (1 - j);
        const int c0 = (nb & 1) * 32;
        uint32_t v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + a * 128 + nb * 32, v);
        tmem_ld_wait();
        if (valid && h < p.H) {
          const size_t e0 = zns_act_index(g, h, w, b8, c0, p.H, p.W, 64);
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(v[i]);
            if (bias) x += __ldg(bias + c0 + i);
            if (p.relu) x = fmaxf(x, 0.f);
            if (do_drop) x = (zns_hash32(e0 + i, seed, p.stream_id + br) >= thr) ? x * keep : 0.f;
            f[i] = x;
          }
          if (mask) {
            const uint4* mp = reinterpret_cast<const uint4*>(mask + e0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 mv = __ldg(mp + q);
              const uint32_t* m2 = reinterpret_cast<const uint32_t*>(&mv);   // fp16 or bf16 forward activations: sign / zero test
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (!act_bits_positive(m2[i] & 0xFFFFu)) f[q * 8 + 2 * i] = 0.f;
                if (!act_bits_positive(m2[i] >> 16)) f[q * 8 + 2 * i + 1] = 0.f;
              }
            }
          }
          uint4* dst = reinterpret_cast<uint4*>(out + e0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(pack_act2(f[q * 8] * p.scale, f[q * 8 + 1] * p.scale, p.out_f16),
                                pack_act2(f[q * 8 + 2] * p.scale, f[q * 8 + 3] * p.scale, p.out_f16),
                                pack_act2(f[q * 8 + 4] * p.scale, f[q * 8 + 5] * p.scale, p.out_f16),
                                pack_act2(f[q * 8 + 6] * p.scale, f[q * 8 + 7] * p.scale, p.out_f16));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}


// ctas = 2: CTA-pair variant (half a weight tile per CTA, up to four accumulators)
static bool fwd_stack_config(const zns_conv_desc* d, FwdTParams* p, int ctas) {
  if (d->c_out != 64 || (d->H & 1)) return false;
  const uint32_t slot = (uint32_t)(WT + d->kw - 1) * 1024u;
  const uint32_t wtile = 16384u / ctas;
  const uint32_t budget = ZNS_SMEM_LIMIT - 1024 - (uint32_t)sizeof(FwdTBarriers) - 64;
  int n_acc = std::min(ctas == 2 ? 4 : 3, d->H / 2);
  while (n_acc > 1 && (uint64_t)((n_acc - 1) * 2 + 2) * slot + 2ull * wtile > budget) --n_acc;
  int slots = (n_acc - 1) * 2 + 2;
  if (slots > MAX_RING || (uint64_t)slots * slot + 2ull * wtile > budget) return false;
  int wst = 2;
  while (wst < MAX_RING && (uint64_t)slots * slot + (uint64_t)(wst + 1) * wtile <= budget) ++wst;
  p->cout = 64; p->stack = 2; p->n_acc = n_acc; p->n_slots = slots; p->n_wstages = wst; p->slot_bytes = slot;
  return true;
}

template <int CTAS>
static int launch_fwd_stack(const zns_conv_desc* d, const FwdTParams& cfg, int n_br, const void* const* in,
                            const void* const* wpk, const float* const* bias, const void* const* mask, void* const* out,
                            const FwdExtra& ex, cudaStream_t st, PlanDump* dry = nullptr) {
  const int G = zns_groups(d->batch);
  FwdTParams p = cfg;
  p.G = G; p.H = d->H; p.W = d->W; p.batch = d->batch;
  p.kh = d->kh; p.kw = d->kw; p.ph = d->kh / 2; p.pw = d->kw / 2;
  p.n_chunks = d->c_in / 64;
  p.n_wtiles = (d->W + WT - 1) / WT;
  if (ex.pool > 0) {
    // whole pool windows per tile: the smallest number of stacked pairs whose rows are a multiple of the pool (3 pairs = 6
    // rows = two windows of 3; 2 pairs = one window of 4)
    int pairs = (ex.pool % 2 == 0) ? ex.pool / 2 : ex.pool;
    ZNS_REQUIRE(pairs <= p.n_acc && d->H % (2 * pairs) == 0 && !mask && d->relu == 0,
                "fused pooling: pool %d does not fit %d stacked accumulators / H %d", ex.pool, p.n_acc, d->H);
    while (2 * pairs <= p.n_acc && d->H % (4 * pairs) == 0 && ex.pool % 2 == 0) pairs *= 2;
    memset(&p.tiles, 0, sizeof(p.tiles));
    p.tiles.n_cols = G * p.n_wtiles * n_br;
    p.tiles.hb = pairs; p.tiles.nb = d->H / (2 * pairs); p.tiles.hs = pairs; p.tiles.ns = 0;
    p.tiles.n_big = p.tiles.n_cols * p.tiles.nb; p.tiles.n_total = p.tiles.n_big;
    p.n_acc = pairs;
    p.n_slots = std::min(p.n_slots, std::min(MAX_RING, (p.n_acc - 1) * 2 + 3));
    p.pool = ex.pool;
  } else {
    const double pair_clk = (double)p.n_chunks * (d->kh + 1) * d->kw * 4.0 * 64.0;   // N = 128 MMAs per stacked pair
    p.tiles = plan_tiles(d->H / 2, G * p.n_wtiles * n_br, p.n_acc, 8000.0 / pair_clk, CTAS);
    p.n_acc = p.tiles.hb;
    p.n_slots = std::min(p.n_slots, std::min(MAX_RING, (p.n_acc - 1) * 2 + 3));
  }
  p.relu = d->relu; p.drop_p = d->dropout_p; p.scale = d->out_scale == 0.f ? 1.f : d->out_scale;
  p.a_f16 = (d->fmt & ZNS_FMT_IN_F16) ? 1 : 0; p.b_f16 = (d->fmt & ZNS_FMT_W_F16) ? 1 : 0; p.out_f16 = (d->fmt & ZNS_FMT_OUT_F16) ? 1 : 0;
  p.seed = d->seed; p.stream_id = d->rng_stream; p.seed_dev = d->seed_dev;
  const size_t smem = 1024 + (size_t)p.n_slots * p.slot_bytes + (size_t)p.n_wstages * (16384 / CTAS) + sizeof(FwdTBarriers) + 64;
  if (dry) {
    dry->kernel = 1; dry->n = 128; dry->ctas = CTAS; dry->tiles = p.tiles; dry->n_slots = p.n_slots; dry->n_stages = p.n_wstages;
    dry->grid_x = p.tiles.n_total; dry->grid_z = 1; dry->smem = smem;
    return ZNS_OK;
  }
  CUtensorMap tm_in[2], tm_w[2];
  for (int b = 0; b < 2; ++b) {
    const int s = b < n_br ? b : 0;
    int rc = make_act_map(&tm_in[b], in[s], G, d->H, d->W, d->c_in, WT + d->kw - 1);
    if (rc) return rc;
    rc = make_w_map(&tm_w[b], wpk[s], d->kh * d->kw, d->c_out, d->c_in, d->c_out);
    if (rc) return rc;
    p.bias[b] = bias ? bias[s] : nullptr;
    p.mask[b] = mask ? (const bf16*)mask[s] : nullptr;
    p.out[b] = (bf16*)out[s];
    p.out2[b] = ex.out2 ? (bf16*)ex.out2[s] : nullptr;
    p.arg[b] = ex.arg ? (uint8_t*)ex.arg[s] : nullptr;
  }
  auto kern = conv_fwd_stack_umma_kernel<CTAS>;
  static unsigned long long attr_devs = 0;          // one bit per device ordinal: function attributes are per device
  if (zns_first_use_on_device(&attr_devs)) {
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ZNS_SMEM_LIMIT));
  }
  dim3 grid(p.tiles.n_total, 1, 1);
  if (CTAS == 2) {
    ZNS_CHECK_CUDA(launch_pair(kern, grid, smem, st, tm_in[0], tm_in[1], tm_w[0], tm_w[1], p));
  } else {
    kern<<<grid, ZNS_CONV_THREADS, smem, st>>>(tm_in[0], tm_in[1], tm_w[0], tm_w[1], p);
  }
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

static bool fwdT_config(const zns_conv_desc* d, FwdTParams* p) {
  if (d->c_out != 64 && d->c_out != 128) return false;
  const int stack = 128 / d->c_out;
  if (stack == 2 && (d->H & 1)) return false;
  const uint32_t slot = (uint32_t)(WTT + d->kw - 1) * 1024u;
  const uint32_t budget = ZNS_SMEM_LIMIT - 1024 - (uint32_t)sizeof(FwdTBarriers) - 64;
  int n_acc = 2;
  if (d->H <= stack) n_acc = 1;
  int slots = (n_acc - 1) * stack + 2;
  if ((uint64_t)slots * slot + 2ull * 16384 > budget) return false;
  int wst = 2;
  while (wst < 4 && (uint64_t)slots * slot + (uint64_t)(wst + 1) * 16384 <= budget) ++wst;
  if (slots < MAX_RING && (uint64_t)(slots + 1) * slot + (uint64_t)wst * 16384 <= budget && d->H > stack) ++slots;
  while (wst < MAX_RING && (uint64_t)slots * slot + (uint64_t)(wst + 1) * 16384 <= budget) ++wst;
  p->cout = d->c_out; p->stack = stack; p->n_acc = n_acc; p->n_slots = slots; p->n_wstages = wst; p->slot_bytes = slot;
  return true;
}

static int launch_fwdT(const zns_conv_desc* d, const FwdTParams& cfg, int n_br, const void* const* in,
                       const void* const* wpk, const float* const* bias, const void* const* mask, void* const* out,
                       cudaStream_t st) {
  const int G = zns_groups(d->batch);
  FwdTParams p = cfg;
  p.G = G; p.H = d->H; p.W = d->W; p.batch = d->batch;
  p.kh = d->kh; p.kw = d->kw; p.ph = d->kh / 2; p.pw = d->kw / 2;
  p.n_chunks = d->c_in / 64;
  p.n_wtiles = (d->W + WTT - 1) / WTT;
  const int rows_per_cta = p.n_acc * p.stack;
  p.n_htiles = (d->H + rows_per_cta - 1) / rows_per_cta;
  p.relu = d->relu; p.drop_p = d->dropout_p; p.scale = d->out_scale == 0.f ? 1.f : d->out_scale;
  p.a_f16 = (d->fmt & ZNS_FMT_IN_F16) ? 1 : 0; p.b_f16 = (d->fmt & ZNS_FMT_W_F16) ? 1 : 0; p.out_f16 = (d->fmt & ZNS_FMT_OUT_F16) ? 1 : 0;
  p.seed = d->seed; p.stream_id = d->rng_stream; p.seed_dev = d->seed_dev;
  const size_t smem = 1024 + (size_t)p.n_slots * p.slot_bytes + (size_t)p.n_wstages * 16384 + sizeof(FwdTBarriers) + 64;
  CUtensorMap tm_in[2], tm_w[2];
  for (int b = 0; b < 2; ++b) {
    const int s = b < n_br ? b : 0;
    int rc = make_act_map(&tm_in[b], in[s], G, d->H, d->W, d->c_in, WTT + d->kw - 1);
    if (rc) return rc;
    rc = make_w_map(&tm_w[b], wpk[s], d->kh * d->kw, d->c_out, d->c_in, d->c_out);
    if (rc) return rc;
    p.bias[b] = bias ? bias[s] : nullptr;
    p.mask[b] = mask ? (const bf16*)mask[s] : nullptr;
    p.out[b] = (bf16*)out[s];
  }
  static unsigned long long attr_devs = 0;          // one bit per device ordinal: function attributes are per device
  if (zns_first_use_on_device(&attr_devs)) {
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(conv_fwdT_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ZNS_SMEM_LIMIT));
  }
  dim3 grid(p.n_wtiles * p.n_htiles * G, 1, n_br);
  conv_fwdT_umma_kernel<<<grid, 192, smem, st>>>(tm_in[0], tm_in[1], tm_w[0], tm_w[1], p);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

template <int N, int HT, int CTAS = 1>
static int launch_fwd(const zns_conv_desc* d, int n_br, const void* const* in, const void* const* wpk,
                      const float* const* bias, const void* const* mask, void* const* out, const FwdExtra& ex,
                      cudaStream_t st, PlanDump* dry = nullptr) {
  void* const* out2 = ex.out2;
  const int G = zns_groups(d->batch);
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.G = G; p.H = d->H; p.W = d->W; p.batch = d->batch;
  p.kh = d->kh; p.kw = d->kw; p.ph = d->kh / 2; p.pw = d->kw / 2;
  p.n_chunks = d->c_in / 64;
  p.n_wtiles = (d->W + WT - 1) / WT;
  p.slot_bytes = (uint32_t)(WT + d->kw - 1) * 1024u;
  p.relu = d->relu; p.drop_p = d->dropout_p; p.scale = d->out_scale == 0.f ? 1.f : d->out_scale;
  p.a_f16 = (d->fmt & ZNS_FMT_IN_F16) ? 1 : 0; p.b_f16 = (d->fmt & ZNS_FMT_W_F16) ? 1 : 0; p.out_f16 = (d->fmt & ZNS_FMT_OUT_F16) ? 1 : 0;
  p.seed = d->seed; p.stream_id = d->rng_stream; p.seed_dev = d->seed_dev;
  const uint32_t btile = N * 128 / CTAS;   // a CTA of a pair stages half of each weight tile
  const uint32_t budget = ZNS_SMEM_LIMIT - 1024 - (uint32_t)sizeof(FwdBarriers) - 64;
  int u_max = std::min(HT, d->H);
  while (u_max > 1 && (uint64_t)(u_max + 1) * p.slot_bytes + 2ull * btile > budget) --u_max;
  if (ex.pool > 0) {
    // one pool window per tile: `pool` rows = `pool` accumulators, windows aligned to multiples of `pool`
    ZNS_REQUIRE(ex.pool <= u_max && d->H % ex.pool == 0, "fused pooling needs pool %d <= %d accumulators and H %% pool == 0", ex.pool, u_max);
    ZNS_REQUIRE(!mask && d->relu == 0, "fused pooling is a forward-pass epilogue (no mask, ReLU comes after the pool)");
    memset(&p.tiles, 0, sizeof(p.tiles));
    p.tiles.n_cols = G * p.n_wtiles * n_br;
    p.tiles.hb = ex.pool; p.tiles.nb = d->H / ex.pool; p.tiles.hs = ex.pool; p.tiles.ns = 0;
    p.tiles.n_big = p.tiles.n_cols * p.tiles.nb; p.tiles.n_total = p.tiles.n_big;
    p.pool = ex.pool;
  } else {
    const double row_clk = (double)p.n_chunks * d->kh * d->kw * 4.0 * (N / 2);
    p.tiles = plan_tiles(d->H, G * p.n_wtiles * n_br, u_max, 8000.0 / row_clk, CTAS);
  }
  const int ht_max = p.tiles.hb;
  int slots = ht_max + 1, bst = 2;
  ZNS_REQUIRE((uint64_t)slots * p.slot_bytes + (uint64_t)bst * btile <= budget,
              "conv tile does not fit shared memory (kw %d, c_out %d)", d->kw, N);
  while (bst < 4 && (uint64_t)slots * p.slot_bytes + (uint64_t)(bst + 1) * btile <= budget) ++bst;
  if (slots < MAX_RING && slots < ht_max + d->kh - 1 &&
      (uint64_t)(slots + 1) * p.slot_bytes + (uint64_t)bst * btile <= budget)
    ++slots;
  while (bst < MAX_RING && (uint64_t)slots * p.slot_bytes + (uint64_t)(bst + 1) * btile <= budget) ++bst;
  p.n_slots = slots; p.n_bstages = bst;
  const size_t smem = 1024 + (size_t)slots * p.slot_bytes + (size_t)bst * btile + sizeof(FwdBarriers) + 64;
  if (dry) {
    dry->kernel = 0; dry->n = N; dry->ctas = CTAS; dry->tiles = p.tiles; dry->n_slots = slots; dry->n_stages = bst;
    dry->grid_x = p.tiles.n_total; dry->grid_z = 1; dry->smem = smem;
    return ZNS_OK;
  }

  CUtensorMap tm_in[2], tm_w[2];
  for (int b = 0; b < 2; ++b) {
    const int s = b < n_br ? b : 0;
    int rc = make_act_map(&tm_in[b], in[s], G, d->H, d->W, d->c_in, WT + d->kw - 1);
    if (rc) return rc;
    rc = make_w_map(&tm_w[b], wpk[s], d->kh * d->kw, d->c_out, d->c_in, N / CTAS);
    if (rc) return rc;
    p.bias[b] = bias ? bias[s] : nullptr;
    p.mask[b] = mask ? (const bf16*)mask[s] : nullptr;
    p.out[b] = (bf16*)out[s];
    p.out2[b] = out2 ? (bf16*)out2[s] : nullptr;
    p.arg[b] = ex.arg ? (uint8_t*)ex.arg[s] : nullptr;
  }
  auto kern = conv_fwd_umma_kernel<N, HT, CTAS>;
  static unsigned long long attr_devs = 0;          // one bit per device ordinal: function attributes are per device
  if (zns_first_use_on_device(&attr_devs)) {
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ZNS_SMEM_LIMIT));
  }
  dim3 grid(p.tiles.n_total, 1, 1);
  if (CTAS == 2) {
    ZNS_CHECK_CUDA(launch_pair(kern, grid, smem, st, tm_in[0], tm_in[1], tm_w[0], tm_w[1], p));
  } else {
    kern<<<grid, ZNS_CONV_THREADS, smem, st>>>(tm_in[0], tm_in[1], tm_w[0], tm_w[1], p);
  }
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

// Kernel choice and launch of the forward / data-gradient convolution; with `dry` only the geometry is reported.
static int conv_fwd_dispatch(const zns_conv_desc* d, int n_br, const void* const* in, const void* const* wpk,
                             const float* const* bias, const void* const* mask, void* const* out, const FwdExtra& ex,
                             cudaStream_t st, PlanDump* dry) {
  void* const* out2 = ex.out2;
  ZNS_REQUIRE(d != nullptr, "NULL argument");
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(d->c_in % 64 == 0 && d->c_in >= 64, "c_in must be a multiple of 64 (got %d)", d->c_in);
  ZNS_REQUIRE((d->kh & 1) && (d->kw & 1) && d->kw <= 41 && d->kh <= 15, "filter %dx%d not supported", d->kh, d->kw);
  ZNS_REQUIRE(d->batch > 0 && d->H > 0 && d->W > 0, "bad geometry");
  ZNS_REQUIRE(d->dropout_p >= 0.f && d->dropout_p < 1.f, "dropout_p out of range");
  // CTA-pair (cta_group::2) kernels need an even number of frame-tile columns per branch.
  // ZNS_CONV_PAIR (A/B switch): 0 = single-CTA kernels, 1 = pairs for N = 128 and the stacked kernel,
  // 2 = also N = 256, 3 (default) = also the weight gradient
  static const int pair_mode = getenv("ZNS_CONV_PAIR") ? atoi(getenv("ZNS_CONV_PAIR")) : 3;
  const bool use_pair = pair_mode != 0;
  const bool can_pair = ((zns_groups(d->batch) * ((d->W + WT - 1) / WT)) % 2) == 0;
  {
    // Experimental transposed kernel (weights on M, 256 positions on N) for C_out <= 128: measured
    // slower than the direct kernel once the issue loops were fixed (profiles/README.md); opt-in only.
    static const bool use_t = getenv("ZNS_CONV_TRANSPOSED") != nullptr;
    FwdTParams cfg;
    memset(&cfg, 0, sizeof(cfg));
    if (use_t && !dry && d->fmt == 0 && !out2 && !ex.pool && fwdT_config(d, &cfg)) return launch_fwdT(d, cfg, n_br, in, wpk, bias, mask, out, st);
  }
  {
    static const bool no_stack = getenv("ZNS_CONV_NO_STACK") != nullptr;   // A/B switch
    FwdTParams cfg;
    memset(&cfg, 0, sizeof(cfg));
    // (without pooling a second bf16 output is written by the direct kernels only)
    const bool stack_ok = !no_stack && (!out2 || ex.pool > 0);
    auto pool_fits = [&](const FwdTParams& c) {
      if (ex.pool <= 0) return true;
      const int pairs = (ex.pool % 2 == 0) ? ex.pool / 2 : ex.pool;
      return pairs <= c.n_acc && d->H % (2 * pairs) == 0;
    };
    if (stack_ok && use_pair && can_pair && fwd_stack_config(d, &cfg, 2) && pool_fits(cfg))
      return launch_fwd_stack<2>(d, cfg, n_br, in, wpk, bias, mask, out, ex, st, dry);
    if (stack_ok && fwd_stack_config(d, &cfg, 1) && pool_fits(cfg)) return launch_fwd_stack<1>(d, cfg, n_br, in, wpk, bias, mask, out, ex, st, dry);
  }
  switch (d->c_out) {
    case 64: return launch_fwd<64, 4>(d, n_br, in, wpk, bias, mask, out, ex, st, dry);
    case 128:
      if (use_pair && can_pair) return launch_fwd<128, 4, 2>(d, n_br, in, wpk, bias, mask, out, ex, st, dry);
      return launch_fwd<128, 4>(d, n_br, in, wpk, bias, mask, out, ex, st, dry);
    case 256:
      if (pair_mode >= 2 && can_pair) return launch_fwd<256, 2, 2>(d, n_br, in, wpk, bias, mask, out, ex, st, dry);
      return launch_fwd<256, 2>(d, n_br, in, wpk, bias, mask, out, ex, st, dry);
    default: return zns_set_error(ZNS_ERR_INVALID, "c_out must be 64, 128 or 256 (got %d)", d->c_out);
  }
}

extern "C" int zns_conv_fwd(const zns_conv_desc* d, int n_br, const void* const* in, const void* const* wpk,
                            const float* const* bias, const void* const* mask, void* const* out, void* const* out_bf16,
                            void* stream) {
  ZNS_REQUIRE(d && in && wpk && out, "NULL argument");
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  for (int b = 0; b < n_br; ++b) ZNS_REQUIRE(in[b] && wpk[b] && out[b] && (!out_bf16 || out_bf16[b]), "NULL tensor for branch %d", b);
  const FwdExtra ex = {out_bf16, 0, nullptr};
  return conv_fwd_dispatch(d, n_br, in, wpk, bias, mask, out, ex, (cudaStream_t)stream, nullptr);
}

// Convolution with the pooling block of models.py:41-44,50-53 fused into its epilogue: conv + bias -> MaxPool2d((pool, 1))
// -> ReLU -> Dropout(d->dropout_p), written as the pooled act tensor (+ optional bf16 copy + arg-max bytes).
extern "C" int zns_conv_pool_fwd(const zns_conv_desc* d, int pool, int n_br, const void* const* in, const void* const* wpk,
                                 const float* const* bias, void* const* out_pooled, void* const* out_pooled_bf16,
                                 void* const* argmax, void* stream) {
  ZNS_REQUIRE(d && in && wpk && out_pooled, "NULL argument");
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(pool >= 2, "pool must be >= 2");
  for (int b = 0; b < n_br; ++b)
    ZNS_REQUIRE(in[b] && wpk[b] && out_pooled[b] && (!out_pooled_bf16 || out_pooled_bf16[b]) && (!argmax || argmax[b]),
                "NULL tensor for branch %d", b);
  const FwdExtra ex = {out_pooled_bf16, pool, argmax};
  return conv_fwd_dispatch(d, n_br, in, wpk, bias, nullptr, out_pooled, ex, (cudaStream_t)stream, nullptr);
}

// Host-only: the launch geometry zns_conv_fwd would use for this layer (no CUDA call is made).
//   out[0..13] = kernel (0 direct, 1 stacked), N, CTAs per cluster, hb, nb, hs, ns, n_cols, n_total, A-row slots,
//                weight stages, grid x, dynamic shared memory bytes, units per column (rows, or row pairs when stacked)
extern "C" int zns_dbg_conv_fwd_plan(const zns_conv_desc* d, int n_br, int* out) {
  ZNS_REQUIRE(d && out, "NULL argument");
  PlanDump pd;
  memset(&pd, 0, sizeof(pd));
  const FwdExtra ex0 = {nullptr, 0, nullptr};
  int rc = conv_fwd_dispatch(d, n_br, nullptr, nullptr, nullptr, nullptr, nullptr, ex0, nullptr, &pd);
  if (rc) return rc;
  const int vals[14] = {pd.kernel, pd.n, pd.ctas, pd.tiles.hb, pd.tiles.nb, pd.tiles.hs, pd.tiles.ns, pd.tiles.n_cols,
                        pd.tiles.n_total, pd.n_slots, pd.n_stages, pd.grid_x, (int)pd.smem, pd.kernel == 1 ? d->H / 2 : d->H};
  for (int i = 0; i < 14; ++i) out[i] = vals[i];
  return ZNS_OK;
}

// ---------------------------------------------------------------------------------------------
// weight-gradient kernel
//   D_tap[cin][cout] = sum_p x[p + tap][cin] * dy[p][cout]
//   A (M side) = x halo row, MN-major; M = 128 = two 64-channel chunks (LBO = chunk stride) or, for
//   c_in == 64, two adjacent taps (LBO = one atom).  B (N side) = dy tile, MN-major, N = NB.
// ---------------------------------------------------------------------------------------------
#define WG_MAX_ITEMS 192
struct WgParams {
  int G, H, W;
  int Cin, Cout;
  int kh, kw, ph, pw;
  int n_wtiles;
  int stack_dy;             // 1: c_out == 64, N = 128 = dy rows (h, h + 1); row items 0..kh, every second row is a step
  int n_rows;               // row items: kh (+ 1 with stack_dy)
  int fold;                 // 1: c_in == 64, M = 2 taps x 64 channels
  int n_acc;                // accumulators (TMEM) per CTA
  int taps_per_cta;         // n_acc * (fold ? 2 : 1)
  int n_sgroups;            // tap groups per filter row
  int grp_base, grp_rem;    // group i covers grp_base + (i < grp_rem) accumulators (balanced split of the row)
  int n_cin_blocks;         // fold ? 1 : Cin / 128
  int n_cout_blocks;        // Cout / NB
  int n_slices;             // position slices
  int n_stages;
  uint32_t x_chunk_bytes;   // (WT + taps_per_cta - 1) * 1024
  uint32_t stage_bytes;
  float* dw[2];
  // CTA-pair kernel only: (tap row, tap group, cin block) work items, two consecutive entries per pair
  int n_item_pairs;
  uint32_t items[WG_MAX_ITEMS];   // r | s0 << 4 | n_acc << 10 | cib << 14 | valid << 18
};

struct WgBarriers {
  uint64_t full[MAX_RING], empty[MAX_RING], acc_full;
  uint32_t tmem_base;
  uint32_t any_step;
};

// CTAS = 2 (NB = 128): a CTA pair shares one dy tile -- each CTA stages 64 of its 128 channels -- and each CTA
// brings the x rows of its own work item (tap row, tap group, cin block) as its half of an M = 256 MMA, so the
// operand reads per SM fall from 8 KB to 6 KB per 64-clock MMA and the dy traffic halves.  Both items of a pair
// have the same accumulator count; a step runs when either item's input row is inside the image (the other's
// TMA load is then zero-filled), a dummy item (odd class size) repeats its partner and skips the epilogue.
template <int NB, int CTAS>
__global__ void __launch_bounds__(ZNS_CONV_THREADS, 1)
conv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
                       const __grid_constant__ CUtensorMap tm_dy0, const __grid_constant__ CUtensorMap tm_dy1,
                       const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  WgBarriers* bars = reinterpret_cast<WgBarriers*>(smem_raw + (smem_base + p.n_stages * p.stage_bytes - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int br = blockIdx.z;
  const CUtensorMap* tm_x = br ? &tm_x1 : &tm_x0;
  const CUtensorMap* tm_dy = br ? &tm_dy1 : &tm_dy0;

  // work item: (tap row r, s-group, cin block, cout block) x position slice
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  int t = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  // the tap group is the slowest index: groups with one more accumulator (sg < grp_rem) are dispatched
  // first, so the in-order block scheduler behaves like longest-first list scheduling
  const int slice = t % p.n_slices; t /= p.n_slices;
  const int cob = t % p.n_cout_blocks; t /= p.n_cout_blocks;
  const int unit = p.fold ? 2 : 1;                       // taps per accumulator
  int cib, r, r_peer, s0, n_acc_eff;
  bool item_valid = true;
  if (CTAS == 2) {
    const uint32_t me = p.items[2 * t + cta_rank], peer = p.items[2 * t + (cta_rank ^ 1u)];
    r = me & 15; s0 = (me >> 4) & 63; n_acc_eff = (me >> 10) & 15; cib = (me >> 14) & 15; item_valid = (me >> 18) & 1;
    r_peer = peer & 15;
  } else {
    cib = t % p.n_cin_blocks; t /= p.n_cin_blocks;
    r = t % p.n_rows; t /= p.n_rows;
    const int sg = t;
    const int acc0 = sg * p.grp_base + min(sg, p.grp_rem); // first accumulator (in row order) of this group
    s0 = acc0 * unit;
    n_acc_eff = p.grp_base + (sg < p.grp_rem ? 1 : 0);
    r_peer = r;
  }
  const int n_xchunks = p.fold ? 1 : 2;
  // stack_dy: only every second output row is a step (its dy tile carries rows h and h + 1)
  const int hmul = p.stack_dy ? 2 : 1;
  const int Hs = (p.H + hmul - 1) / hmul;
  const int n_steps_total = p.G * Hs * p.n_wtiles;
  const int q0 = (int)((long long)n_steps_total * slice / p.n_slices);
  const int q1 = (int)((long long)n_steps_total * (slice + 1) / p.n_slices);
  constexpr uint32_t kDyChunk = WT * 1024;  // 128 positions x 64 channels
  // a step (output row h) is live when the input row of this item or of its pair partner exists
  auto live = [&](int h) {
    const int hh = h + r - p.ph, hp = h + r_peer - p.ph;
    return (hh >= 0 && hh < p.H) || (hp >= 0 && hp < p.H);
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_stages; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->acc_full), 1);
    bars->any_step = 0;
    mbar_fence_init();
    tma_prefetch_desc(tm_x);
    tma_prefetch_desc(tm_dy);
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_pair(smem_u32(&bars->tmem_base), 512); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(&bars->tmem_base), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t dy_off = n_xchunks * p.x_chunk_bytes;  // dy tiles follow the x chunks inside a stage

  const uint32_t bar_full = smem_u32(&bars->full[0]), bar_empty = smem_u32(&bars->empty[0]);
  const uint32_t n_st = p.n_stages;
  // position-step cursor (wt, h, g) of step q0, advanced incrementally (no div/mod in the loops)
  const int wt_0 = q0 % p.n_wtiles, gh_0 = q0 / p.n_wtiles;
  const int h_0 = (gh_0 % Hs) * hmul, g_0 = gh_0 / Hs;
  if (warp == 0) {
    if (elect_one()) {
      uint32_t st = 0, par = 1;
      int wt = wt_0, h = h_0, g = g_0;
      for (int q = q0; q < q1; ++q) {
        const int hh = h + r - p.ph;
        if (live(h)) {
          mbar_wait(bar_empty + 8 * st, par);
          const uint32_t base = smem_base + st * p.stage_bytes;
          if (leader) mbar_expect_tx(bar_full + 8 * st, CTAS * p.stage_bytes);
          if (CTAS == 2) {
            const uint32_t bar = (bar_full + 8 * st) & ZNS_PEER_MASK;
            for (int xc = 0; xc < n_xchunks; ++xc)      // a row outside the image is zero-filled by TMA
              tma_load_5d_pair(base + xc * p.x_chunk_bytes, tm_x, bar, (cib * n_xchunks + xc) * 64, 0,
                               wt * WT - p.pw + s0, hh, g);
            if (p.stack_dy) tma_load_5d_pair(base + dy_off, tm_dy, bar, 0, 0, wt * WT, h + (int)cta_rank, g);
            else tma_load_5d_pair(base + dy_off, tm_dy, bar, cob * NB + (int)cta_rank * 64, 0, wt * WT, h, g);
          } else {
            for (int xc = 0; xc < n_xchunks; ++xc)
              tma_load_5d(base + xc * p.x_chunk_bytes, tm_x, bar_full + 8 * st, (cib * n_xchunks + xc) * 64, 0,
                          wt * WT - p.pw + s0, hh, g);
            for (int j = 0; j < NB / 64; ++j) {
              if (p.stack_dy) tma_load_5d(base + dy_off + j * kDyChunk, tm_dy, bar_full + 8 * st, 0, 0, wt * WT, h + j, g);
              else tma_load_5d(base + dy_off + j * kDyChunk, tm_dy, bar_full + 8 * st, cob * NB + j * 64, 0, wt * WT, h, g);
            }
          }
          if (++st == n_st) { st = 0; par ^= 1; }
        }
        if (++wt == p.n_wtiles) { wt = 0; h += hmul; if (h >= p.H) { h = 0; ++g; } }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128 * CTAS, NB, 1, 1);
      const uint32_t dbase = smem_base & 0x3FFFFu;   // descriptor offsets (the same in both CTAs of a pair)
      const uint32_t a_lbo = p.fold ? 1024u : p.x_chunk_bytes;
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version 1, SWIZZLE_128B
      const uint32_t a_hi16 = ((a_lbo >> 4) & 0x3FFF) << 16, b_hi16 = (kDyChunk >> 4) << 16;
      const uint32_t a_step = p.fold ? 2048u : 1024u;
      uint32_t st = 0, par = 0, any = 0;
      int wt = wt_0, h = h_0;
      for (int q = q0; q < q1; ++q) {
        if (live(h)) {
          mbar_wait(bar_full + 8 * st, par);
          tc_fence_after();
          const uint32_t base = dbase + st * p.stage_bytes;
          const uint32_t b_lo = ((base + dy_off) >> 4) | b_hi16;
          uint32_t xa = base;
          for (int a = 0; a < n_acc_eff; ++a) {
            const uint32_t a_lo = (xa >> 4) | a_hi16;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (CTAS == 2)
                umma_bf16_pair(tmem + a * NB, ((uint64_t)kDescHi << 32) | (a_lo + 128 * k),
                               ((uint64_t)kDescHi << 32) | (b_lo + 128 * k), idesc, any | (k > 0));
              else
                umma_bf16(tmem + a * NB, ((uint64_t)kDescHi << 32) | (a_lo + 128 * k),
                          ((uint64_t)kDescHi << 32) | (b_lo + 128 * k), idesc, any | (k > 0));
            }
            xa += a_step;
          }
          if (CTAS == 2) umma_commit_pair(bar_empty + 8 * st); else umma_commit(bar_empty + 8 * st);
          if (++st == n_st) { st = 0; par ^= 1; }
          any = 1;
        }
        if (++wt == p.n_wtiles) { wt = 0; h += hmul; if (h >= p.H) h = 0; }
      }
      if (CTAS == 2) {
        if (any) umma_commit_pair(smem_u32(&bars->acc_full));   // (no live step: the epilogues of both CTAs do not wait)
      } else if (any) {
        *reinterpret_cast<volatile uint32_t*>(&bars->any_step) = 1;
        umma_commit(smem_u32(&bars->acc_full));
      } else {
        mbar_arrive(smem_u32(&bars->acc_full));
      }
    }
    __syncwarp();
  } else {
    bool any;
    if (CTAS == 2) {
      // both CTAs decide locally whether the slice has a live step (at most H rows to look at)
      any = false;
      const int gh1 = (q1 - 1) / p.n_wtiles;
      if (q1 > q0) {
        const int gh_end = min(gh1, gh_0 + Hs - 1);   // every row step is covered after Hs of them
        for (int gh = gh_0; gh <= gh_end; ++gh) any = any || live((gh % Hs) * hmul);
      }
      if (any) { mbar_wait(smem_u32(&bars->acc_full), 0); tc_fence_after(); }
    } else {
      mbar_wait(smem_u32(&bars->acc_full), 0);
      tc_fence_after();
      any = *reinterpret_cast<volatile uint32_t*>(&bars->any_step) != 0;
    }
    if (any && item_valid) {
      const int quad = warp & 3, half = (warp - 2) >> 2;
      const int m = quad * 32 + lane;
      float* dw = p.dw[br];
      int nb = half;
      for (int a = 0; a < n_acc_eff; ++a, nb -= NB / 32) {
        int tap_s, cin;
        if (p.fold) { tap_s = s0 + 2 * a + (m >> 6); cin = m & 63; }
        else        { tap_s = s0 + a; cin = cib * 128 + m; }
        const bool valid = tap_s < p.kw;
#pragma unroll 1
        for (; nb < NB / 32; nb += ZNS_EPI) {
          uint32_t v[32];
          tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + a * NB + nb * 32, v);
          tmem_ld_wait();
          // stack_dy: columns 0-63 pair x row hh with dy row h (tap row r), columns 64-127 with dy row h + 1 (tap row r - 1)
          const int tap_r = p.stack_dy ? r - (nb >> 1) : r;
          const int cout0 = p.stack_dy ? (nb & 1) * 32 : cob * NB + nb * 32;
          if (valid && tap_r >= 0 && tap_r < p.kh) {
            const size_t o0 = ((size_t)(tap_r * p.kw + tap_s) * p.Cout + cout0) * p.Cin + cin;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dw + o0 + (size_t)j * p.Cin, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

// stack_dy (c_out == 64 with NB = 128): a 128 x 64 x 16 MMA is bound by its A-tile read (48 clk instead of 32), so the
// dy tiles of output rows h and h + 1 are stacked on N.  On input row hh = h + rho - ph the first 64 columns then
// accumulate tap row rho and the last 64 tap row rho - 1; with steps on even h only, tap row r gets its even rows from
// item rho = r and its odd rows from item rho = r + 1: kh + 1 row items of H/2 steps each, every MMA at N = 128.
template <int NB, int CTAS = 1>
static int launch_wgrad(const zns_conv_desc* d, int n_br, const void* const* x, const void* const* dy, float* const* dwpk,
                        cudaStream_t st, bool stack_dy = false, PlanDump* dry = nullptr) {
  static_assert(CTAS == 1 || NB == 128, "the CTA-pair weight-gradient kernel splits a 128-channel dy tile");
  if (stack_dy && (NB != 128 || d->c_out != 64)) return zns_set_error(ZNS_ERR_INVALID, "stack_dy needs c_out == 64 and NB == 128");
  const int G = zns_groups(d->batch);
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.G = G; p.H = d->H; p.W = d->W; p.Cin = d->c_in; p.Cout = d->c_out;
  p.kh = d->kh; p.kw = d->kw; p.ph = d->kh / 2; p.pw = d->kw / 2;
  p.n_wtiles = (d->W + WT - 1) / WT;
  p.fold = d->c_in == 64;
  p.stack_dy = stack_dy ? 1 : 0;
  p.n_rows = d->kh + p.stack_dy;
  p.n_acc = 512 / NB;
  if (p.fold) p.n_acc = std::min(p.n_acc, (d->kw + 1) / 2);
  else p.n_acc = std::min(p.n_acc, d->kw);
  // balanced split of a filter row over tap groups: e.g. 17 taps with 4 accumulators -> 5 groups of
  // 4,4,3,3,3 accumulators (a 4,4,4,4,1 split leaves one CTA with a quarter of the MMAs per loaded byte)
  {
    const int unit = p.fold ? 2 : 1;
    const int row_acc = (d->kw + unit - 1) / unit;           // accumulators needed for one filter row
    const int groups = (row_acc + p.n_acc - 1) / p.n_acc;
    p.n_sgroups = groups;
    p.grp_base = row_acc / groups;
    p.grp_rem = row_acc % groups;
    p.n_acc = p.grp_base + (p.grp_rem ? 1 : 0);
    p.taps_per_cta = p.n_acc * unit;
  }
  p.n_cin_blocks = p.fold ? 1 : d->c_in / 128;
  p.n_cout_blocks = stack_dy ? 1 : d->c_out / NB;
  p.x_chunk_bytes = (uint32_t)(WT + p.taps_per_cta - 1) * 1024u;
  p.stage_bytes = (p.fold ? 1 : 2) * p.x_chunk_bytes + (NB / 64 / CTAS) * WT * 1024u;   // per CTA
  const uint32_t budget = ZNS_SMEM_LIMIT - 1024 - (uint32_t)sizeof(WgBarriers) - 64;
  p.n_stages = std::min<int>(MAX_RING, budget / p.stage_bytes);
  ZNS_REQUIRE(p.n_stages >= 2, "wgrad stage does not fit shared memory twice");
  int items = p.n_rows * p.n_sgroups * p.n_cin_blocks * p.n_cout_blocks;
  int n_big_items = 0, n_small_items = 0;   // pair kernel: CTAs (dummies included) per class, cout block and slice
  if (CTAS == 2) {
    // work items by class (accumulator count), bigger class first; an odd class is padded with a dummy
    // that repeats its partner (valid bit clear)
    const int unit = p.fold ? 2 : 1;
    int n = 0;
    for (int cls = 0; cls < 2; ++cls) {
      const int first = n;
      for (int sg = 0; sg < p.n_sgroups; ++sg) {
        const bool big = p.grp_rem == 0 || sg < p.grp_rem;
        if (big != (cls == 0)) continue;
        const int acc0 = sg * p.grp_base + std::min(sg, p.grp_rem);
        const int n_acc = p.grp_base + (sg < p.grp_rem ? 1 : 0);
        for (int r = 0; r < p.n_rows; ++r)
          for (int cib = 0; cib < p.n_cin_blocks; ++cib) {
            ZNS_REQUIRE(n < WG_MAX_ITEMS - 1, "too many weight-gradient work items for the pair kernel");
            p.items[n++] = (uint32_t)r | ((uint32_t)(acc0 * unit) << 4) | ((uint32_t)n_acc << 10) | ((uint32_t)cib << 14) | (1u << 18);
          }
      }
      if ((n - first) & 1) { p.items[n] = p.items[n - 1] & ~(1u << 18); ++n; }
      (cls == 0 ? n_big_items : n_small_items) = n - first;
    }
    p.n_item_pairs = n / 2;
    items = n * p.n_cout_blocks;   // CTAs per slice and branch
  }
  const int n_steps_total = G * (stack_dy ? (d->H + 1) / 2 : d->H) * p.n_wtiles;
  // position slices: simulate the in-order dispatch of the two CTA classes (groups with grp_base + 1 and
  // with grp_base accumulators) for every slice count and keep the shortest makespan; more slices also
  // mean more atomics in the epilogue, hence the small per-CTA overhead term
  int best = 1;
  {
    long per_group = (long)p.n_rows * p.n_cin_blocks * p.n_cout_blocks * n_br;   // CTAs per tap group and slice
    const int acc_big = p.grp_base + (p.grp_rem ? 1 : 0), acc_small = p.grp_base;
    int n_big_groups = p.grp_rem ? p.grp_rem : p.n_sgroups, n_small_groups = p.grp_rem ? p.n_sgroups - p.grp_rem : 0;
    if (CTAS == 2) { per_group = (long)p.n_cout_blocks * n_br; n_big_groups = n_big_items; n_small_groups = n_small_items; }
    double best_t = 1e30;
    for (int sl = 1; sl <= 64; ++sl) {
      const double steps = (double)n_steps_total / sl;
      if (steps < 16 && sl > 1) break;
      const double ovh = 24.0;   // prologue + TMEM drain + atomics, in units of one accumulator-step (8 MMAs)
      std::priority_queue<double, std::vector<double>, std::greater<double>> sm;
      for (int i = 0; i < ZNS_NUM_SMS; ++i) sm.push(0.0);
      double last = 0.0;
      auto push = [&](double dur) { const double t = sm.top() + dur; sm.pop(); sm.push(t); last = std::max(last, t); };
      for (long i = 0; i < per_group * n_big_groups * sl; ++i) push(steps * acc_big + ovh);
      for (long i = 0; i < per_group * n_small_groups * sl; ++i) push(steps * acc_small + ovh);
      if (last < best_t * 0.99) { best_t = last; best = sl; }
    }
  }
  p.n_slices = best;
  const size_t smem = 1024 + (size_t)p.n_stages * p.stage_bytes + sizeof(WgBarriers) + 64;
  if (dry) {
    dry->kernel = 3; dry->n = NB; dry->ctas = CTAS; dry->n_stages = p.n_stages; dry->smem = smem;
    dry->grid_x = items * p.n_slices; dry->grid_z = n_br;
    dry->n_slices = p.n_slices; dry->n_acc = p.n_acc; dry->n_sgroups = p.n_sgroups; dry->grp_base = p.grp_base;
    dry->grp_rem = p.grp_rem; dry->n_rows = p.n_rows; dry->stack_dy = p.stack_dy; dry->fold = p.fold;
    dry->n_cin_blocks = p.n_cin_blocks; dry->n_cout_blocks = p.n_cout_blocks; dry->n_item_pairs = p.n_item_pairs;
    static_assert(sizeof(dry->items) >= sizeof(p.items), "PlanDump::items too small");
    memcpy(dry->items, p.items, sizeof(p.items));
    return ZNS_OK;
  }

  CUtensorMap tm_x[2], tm_dy[2];
  for (int b = 0; b < 2; ++b) {
    const int s = b < n_br ? b : 0;
    int rc = make_act_map(&tm_x[b], x[s], G, d->H, d->W, d->c_in, WT + p.taps_per_cta - 1);
    if (rc) return rc;
    rc = make_act_map(&tm_dy[b], dy[s], G, d->H, d->W, d->c_out, WT);
    if (rc) return rc;
    p.dw[b] = dwpk[s];
  }
  auto kern = conv_wgrad_umma_kernel<NB, CTAS>;
  static unsigned long long attr_devs = 0;          // one bit per device ordinal: function attributes are per device
  if (zns_first_use_on_device(&attr_devs)) {
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ZNS_SMEM_LIMIT));
  }
  dim3 grid(items * p.n_slices, 1, n_br);
  if (CTAS == 2) {
    ZNS_CHECK_CUDA(launch_pair(kern, grid, smem, st, tm_x[0], tm_x[1], tm_dy[0], tm_dy[1], p));
  } else {
    kern<<<grid, ZNS_CONV_THREADS, smem, st>>>(tm_x[0], tm_x[1], tm_dy[0], tm_dy[1], p);
  }
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

// Kernel choice and launch of the weight gradient; with `dry` only the geometry is reported.
static int conv_wgrad_dispatch(const zns_conv_desc* d, int n_br, const void* const* x, const void* const* dy,
                               float* const* dwpk, cudaStream_t st, PlanDump* dry) {
  ZNS_REQUIRE(d != nullptr, "NULL argument");
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  ZNS_REQUIRE(d->c_in == 64 || d->c_in % 128 == 0, "wgrad needs c_in == 64 or a multiple of 128 (got %d)", d->c_in);
  ZNS_REQUIRE(d->c_out % 64 == 0, "c_out must be a multiple of 64");
  ZNS_REQUIRE((d->kh & 1) && (d->kw & 1) && d->kw <= 41 && d->kh <= 15, "filter %dx%d not supported", d->kh, d->kw);
  ZNS_REQUIRE(d->batch > 0 && d->H > 0 && d->W > 0, "bad geometry");
  // CTA-pair kernel when the (tap row, tap group, cin block) items fit its table
  static const int pair_mode = getenv("ZNS_CONV_PAIR") ? atoi(getenv("ZNS_CONV_PAIR")) : 3;
  const int unit = d->c_in == 64 ? 2 : 1;
  const int row_acc = (d->kw + unit - 1) / unit;
  const int n_items = (d->kh + 1) * ((row_acc + 3) / 4) * (d->c_in == 64 ? 1 : d->c_in / 128);
  const bool pair_ok = pair_mode >= 3 && n_items + 2 <= WG_MAX_ITEMS && d->c_in / 128 <= 15;
  if (d->c_out == 64) {
    // dy rows (h, h + 1) stacked on N (N = 128 MMAs, CTA pairs) instead of the A-read-bound N = 64 kernel: +1.0 .. 1.5 % on the
    // training step in an alternating A/B on one box (profiles/r02_wgrad_stack_ab.txt); ZNS_WGRAD_STACK=0 restores N = 64
    static const bool stack = getenv("ZNS_WGRAD_STACK") == nullptr || atoi(getenv("ZNS_WGRAD_STACK")) != 0;
    if (stack && d->kh < 15)
      return pair_ok ? launch_wgrad<128, 2>(d, n_br, x, dy, dwpk, st, true, dry) : launch_wgrad<128, 1>(d, n_br, x, dy, dwpk, st, true, dry);
    return launch_wgrad<64>(d, n_br, x, dy, dwpk, st, false, dry);
  }
  if (pair_ok) return launch_wgrad<128, 2>(d, n_br, x, dy, dwpk, st, false, dry);
  return launch_wgrad<128>(d, n_br, x, dy, dwpk, st, false, dry);
}

extern "C" int zns_conv_wgrad(const zns_conv_desc* d, int n_br, const void* const* x, const void* const* dy,
                              float* const* dwpk, void* stream) {
  ZNS_REQUIRE(d && x && dy && dwpk, "NULL argument");
  ZNS_REQUIRE(n_br == 1 || n_br == 2, "n_br must be 1 or 2");
  for (int b = 0; b < n_br; ++b) ZNS_REQUIRE(x[b] && dy[b] && dwpk[b], "NULL tensor for branch %d", b);
  return conv_wgrad_dispatch(d, n_br, x, dy, dwpk, (cudaStream_t)stream, nullptr);
}

// Host-only: the launch geometry zns_conv_wgrad would use for this layer (no CUDA call is made).
//   out[0..15] = NB, CTAs per cluster, position slices, accumulators per CTA, tap groups per row, grp_base, grp_rem,
//                row items, stack_dy, fold, cin blocks, cout blocks, item pairs, stages, grid x, dynamic smem bytes
//   items[0..2*out[12]) = the CTA-pair work-item table (r | s0 << 4 | n_acc << 10 | cib << 14 | valid << 18); may be NULL
extern "C" int zns_dbg_conv_wgrad_plan(const zns_conv_desc* d, int n_br, int* out, unsigned int* items) {
  ZNS_REQUIRE(d && out, "NULL argument");
  PlanDump pd;
  memset(&pd, 0, sizeof(pd));
  int rc = conv_wgrad_dispatch(d, n_br, nullptr, nullptr, nullptr, nullptr, &pd);
  if (rc) return rc;
  const int vals[16] = {pd.n, pd.ctas, pd.n_slices, pd.n_acc, pd.n_sgroups, pd.grp_base, pd.grp_rem, pd.n_rows, pd.stack_dy,
                        pd.fold, pd.n_cin_blocks, pd.n_cout_blocks, pd.n_item_pairs, pd.n_stages, pd.grid_x, (int)pd.smem};
  for (int i = 0; i < 16; ++i) out[i] = vals[i];
  if (items) for (int i = 0; i < 2 * pd.n_item_pairs; ++i) items[i] = pd.items[i];
  return ZNS_OK;
}

// ---------------------------------------------------------------------------------------------
// descriptor probe: one CTA, one accumulator, operands written to shared memory by plain stores
// with the 128B swizzle applied by hand, so that the UMMA descriptor conventions are tested
// independently of TMA.   variant 0: A, B K-major ([128][k], [n][k]);  variant 1: A, B MN-major
// ([k][128], [k][n]).  D fp32 [128][n].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(int variant, const bf16* __restrict__ A, const bf16* __restrict__ B, float* __restrict__ D, int n, int k) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int kblocks = k / 64;
  const uint32_t a_bytes = variant == 0 ? (uint32_t)kblocks * 128 * 128 : (uint32_t)2 * k * 128;
  uint8_t* sa = sm;
  uint8_t* sb = sm + a_bytes;
  if (variant == 0) {
    // K-major: tile per 64-wide k block: [rows][128 B], atoms of 8 rows
    for (int i = threadIdx.x; i < 128 * k; i += 128) {
      const int row = i / k, kk = i % k;
      const int kb = kk / 64, kl = kk % 64;
      const uint32_t off = kb * (128 * 128) + (row / 8) * 1024 + (row % 8) * 128 + (((kl / 8) ^ (row % 8)) * 16) + (kl % 8) * 2;
      *reinterpret_cast<bf16*>(sa + off) = A[i];
    }
    for (int i = threadIdx.x; i < n * k; i += 128) {
      const int row = i / k, kk = i % k;
      const int kb = kk / 64, kl = kk % 64;
      const uint32_t off = kb * (n * 128) + (row / 8) * 1024 + (row % 8) * 128 + (((kl / 8) ^ (row % 8)) * 16) + (kl % 8) * 2;
      *reinterpret_cast<bf16*>(sb + off) = B[i];
    }
  } else {
    // MN-major: per 64-wide mn group: [k rows][128 B]; atoms of 8 k rows; groups at LBO = k*128
    for (int i = threadIdx.x; i < 128 * k; i += 128) {
      const int kk = i / 128, mm = i % 128;
      const int mg = mm / 64, ml = mm % 64;
      const uint32_t off = mg * (k * 128) + (kk / 8) * 1024 + (kk % 8) * 128 + (((ml / 8) ^ (kk % 8)) * 16) + (ml % 8) * 2;
      *reinterpret_cast<bf16*>(sa + off) = A[i];
    }
    for (int i = threadIdx.x; i < n * k; i += 128) {
      const int kk = i / n, nn = i % n;
      const int ng = nn / 64, nl = nn % 64;
      const uint32_t off = ng * (k * 128) + (kk / 8) * 1024 + (kk % 8) * 128 + (((nl / 8) ^ (kk % 8)) * 16) + (nl % 8) * 2;
      *reinterpret_cast<bf16*>(sb + off) = B[i];
    }
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 256); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async proxy (UMMA) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n, variant, variant);
    const uint32_t sa_u = base, sb_u = base + a_bytes;
    for (int ks = 0; ks < k / 16; ++ks) {
      uint64_t da, db;
      if (variant == 0) {
        const int kb = ks / 4, kq = ks % 4;
        da = umma_desc_sw128(sa_u + kb * (128 * 128) + kq * 32, 16, 1024);
        db = umma_desc_sw128(sb_u + kb * (n * 128) + kq * 32, 16, 1024);
      } else {
        da = umma_desc_sw128(sa_u + ks * 2048, (uint32_t)k * 128, 1024);
        db = umma_desc_sw128(sb_u + ks * 2048, (uint32_t)k * 128, 1024);
      }
      umma_bf16(tmem, da, db, idesc, ks > 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int nb = 0; nb < n / 32; ++nb) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + nb * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * n + nb * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

extern "C" int zns_dbg_umma_probe(int variant, const void* a, const void* b, float* d, int n, int k, void* stream) {
  ZNS_REQUIRE(a && b && d, "NULL argument");
  ZNS_REQUIRE((variant == 0 || variant == 1) && n % 64 == 0 && n >= 64 && n <= 256 && k % 64 == 0 && k >= 64 && k <= 256,
              "probe supports n,k in multiples of 64 up to 256");
  const size_t smem = 1024 + (size_t)(128 + n) * k * 2 + 1024;
  static unsigned long long attr_devs = 0;          // one bit per device ordinal: function attributes are per device
  if (zns_first_use_on_device(&attr_devs)) {
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(variant, (const bf16*)a, (const bf16*)b, d, n, k);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 rate microbenchmark (test/diagnostic only): one CTA per SM issues `iters` groups of
// `per_group` MMAs (128 x n x 16, bf16) on shared-memory-resident garbage and reports cycles.
//   mode bit0: rotate over `n_acc` accumulators instead of one
//   mode bit1: commit + wait on an mbarrier after every group (pipeline round trip)
//   mode bit2: B operand address changes per MMA (different smem tiles)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_rate_kernel(int n, int iters, int per_group, int mode, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc_bf16(128, n, 0, 0);
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const int n_acc = (mode & 1) ? 512 / n : 1;
    uint32_t phase = 0;
    const long long t0 = clock64();
    if (mode & 8) {
      // tight loop: one elected thread, descriptors precomputed, 4 MMAs unrolled (like the conv kernels)
      if (elect_one()) {
        const uint32_t a_lo = (base >> 4) | (1u << 16), b_lo = ((base + 64 * 1024) >> 4) | (1u << 16);
        const int n_acc8 = (mode & 1) ? 512 / n : 1;
        uint32_t acc = 0, a_off = 0;
        for (int it = 0; it < iters * per_group / 4; ++it) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + acc * n, ((uint64_t)kDescHi << 32) | (a_lo + a_off + 2 * k), ((uint64_t)kDescHi << 32) | (b_lo + 2 * k), idesc, 1);
          if (++acc == (uint32_t)n_acc8) acc = 0;
          a_off += 64; if (a_off >= 1024) a_off = 0;
        }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), 0);
    } else
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        for (int g = 0; g < per_group; ++g) {
          const int q = it * per_group + g;
          const uint32_t a_addr = base + (q % 16) * 1024 + (g & 3) * 32;
          const uint32_t b_addr = base + 64 * 1024 + ((mode & 4) ? (q % 3) * 32768 : 0) + (g & 3) * 32;
          umma_bf16(tmem + (q % n_acc) * n, ((uint64_t)kDescHi << 32) | ((a_addr & 0x3FFFF) >> 4) | (1u << 16),
                    ((uint64_t)kDescHi << 32) | ((b_addr & 0x3FFFF) >> 4) | (1u << 16), idesc, 1);
        }
        if (mode & 2) umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      if (mode & 2) { mbar_wait(smem_u32(&bar), phase); phase ^= 1; }
    }
    if (!(mode & 2) && !(mode & 8)) {
      if (elect_one()) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), 0);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

extern "C" int zns_dbg_umma_rate(int n, int iters, int per_group, int mode, int n_ctas, long long* cycles, void* stream) {
  ZNS_REQUIRE(cycles && n >= 16 && n <= 256 && n % 16 == 0 && iters > 0 && per_group > 0 && n_ctas > 0, "bad arguments");
  static unsigned long long attr_devs = 0;          // one bit per device ordinal: function attributes are per device
  if (zns_first_use_on_device(&attr_devs)) {
    ZNS_CHECK_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  umma_rate_kernel<<<n_ctas, 128, 170 * 1024, (cudaStream_t)stream>>>(n, iters, per_group, mode, cycles);
  ZNS_CHECK_LAUNCH();
  return ZNS_OK;
}
