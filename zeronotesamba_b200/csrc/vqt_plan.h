// Plan of the VQT / CQT front-end (shared by vqt.cu and vqt_umma.cu).
#pragma once
#include <stdint.h>
#include <vector>

#define ZNS_VQT_MAX_OCT 10
#define ZNS_VQT_MAX_MMA 120

// One tcgen05.mma of a level tile (vqt_umma.cu).  Operand offsets are bytes from the start of the shared-memory
// slot that holds the MMA's plane group (A; the term selects the x1 / x2 half of the slot) and from the level's
// coefficient image (B); both operands use the no-swizzle K-major layout with rows at 16-byte pitch (SBO = 128).
struct VqtMma {
  uint32_t a_off;
  uint32_t b_off;
  uint16_t n;       // MMA N (multiple of 16)
  uint16_t d_col;   // TMEM column relative to the job's accumulator stage
  uint8_t term;     // 0: x1 (leading fp16 term of the signal), 1: x2 (residual * 2048)
  uint8_t job;      // 0: filterbank, 1 + p: decimator pass p
  uint8_t part;     // accumulator part: 0 = columns fed by x1 . g1 (filterbank: x1 . [g1; g2]), 1 = the 1/2048-scaled columns
  uint8_t b_rows8;  // rows / 8 of the coefficient tile this window lives in (B LBO = 128 * b_rows8)
};
// The same MMA as an issuing thread consumes it (16 bytes, read from the kernel parameter bank): descriptor low words
// without the shared-memory base, the instruction descriptor, the accumulator column.
struct VqtMmaPacked {
  uint32_t a_lo;    // (a_off + term * slot_term_bytes) >> 4 | (LBO >> 4) << 16
  uint32_t b_lo;    // b_off >> 4 | (LBO >> 4) << 16
  uint32_t idesc;
  uint32_t col;     // TMEM column of the window relative to stage 0 of its accumulator ring
};
// A run of MMAs of one (job, part) unit inside one plane group, issued by one of the issuer warps; the inner loop over a
// segment is branch free.
struct VqtSeg {
  uint16_t begin, count;   // MMAs [begin, begin + count) of mma[] / pk[]
  uint8_t job, part;
  uint8_t flags;           // 1: the unit's first segment in the tile (wait for the accumulator stage, clear the unit's
                           //    columns), 2: its last (commit to the stage's "full" barrier)
  uint8_t pad;
};
#define ZNS_VQT_MAX_SEGS 64
#define ZNS_VQT_ISSUERS 4
#define ZNS_VQT_MAX_GROUPS 8
#define ZNS_VQT_MAX_JOBS 3

// Geometry + MMA program of one pyramid level, passed to the kernel by value.
struct VqtLevelDev {
  int q;            // 16-byte chunk planes per row (row = 8 q samples)
  int hb, ha;       // halo rows before / after the 128 tile rows
  int rtot;         // rows per plane in shared memory (128 + hb + ha, padded against bank conflicts)
  int a_lbo;        // A descriptor leading-dimension byte offset (plane stride; 16 when q == 1)
  int fpr;          // frames per row
  int hop, n_fft;   // at this level's rate
  int bin0;         // first output bin of this octave
  int dec_w;        // decimator outputs per row (4 q); 0 on the last level
  int dec_wp;       // accumulator columns per decimator pass and part (<= 64)
  int n_pass;       // decimator passes per tile (column blocks of 64 outputs)
  int fb_n1, fb_n2; // filterbank accumulator columns: x1 . [g1; g2] per frame (48 each), x2 . g1 (24 each, >= 32)
  int pg;           // planes per group (shared-memory slot)
  int gpt;          // groups per tile
  int n_slots;      // slots in the shared-memory ring
  int slot_term_bytes;  // bytes of one term (x1 or x2) of a slot
  int g_order[ZNS_VQT_MAX_GROUPS];          // group processed at position 0 .. gpt-1
  int ring_base[2], ring_width[2], ring_stages[2];   // TMEM accumulator rings: [0] filterbank, [1] decimator
  int n_jobs;                               // jobs per tile
  int ep_job[ZNS_VQT_MAX_JOBS];             // job ids in completion order (epilogue order)
  int n_mma;
  int b_bytes;      // coefficient image size
  float dec_scale;  // sqrt(2) / (tap scale)
  float fb_scale;   // 1 / (coefficient scale)
  int n_seg;
  int seg_begin[ZNS_VQT_ISSUERS][ZNS_VQT_MAX_GROUPS + 1];  // segment range of each (issuer, group position)
  VqtSeg seg[ZNS_VQT_MAX_SEGS];
  VqtMma mma[ZNS_VQT_MAX_MMA];              // MMA list (host / tests)
  // the same list as the issuing threads read it -- from the kernel parameter (constant) bank, NOT from shared memory:
  // a shared-memory load of an issuer queues behind the loader warps' traffic in the SM's in-order load/store pipe
  VqtMmaPacked pk[ZNS_VQT_MAX_MMA];
};

struct zns_vqt_plan {
  int sr, hop, n_bins, bpo, n_oct;
  int max_batch, max_samples;
  int n_fft[ZNS_VQT_MAX_OCT];
  float* d_coef[ZNS_VQT_MAX_OCT];  // [n_fft][2][bpo/2][2] interleaved (SIMT filterbank kernel, edge frames)
  uint16_t* d_coef_bf[ZNS_VQT_MAX_OCT];  // [2 terms][2*bpo columns][n_fft] fp16 (mma.sync filterbank)
  float coef_inv_scale[ZNS_VQT_MAX_OCT]; // 1 / (power-of-two scale applied to the fp16 coefficients)
  uint16_t* d_coef_umma[ZNS_VQT_MAX_OCT]; // round-1 tcgen05 filterbank: stacked B operand image, 128B-swizzled
  float* d_inv_sqrt_len;           // [n_bins]
  float* d_scratch[ZNS_VQT_MAX_OCT];  // decimated signals, octave >= 1 (legacy path)
  float* d_stage_in;               // for *_host: [max_batch][max_samples]
  float* d_stage_out;              // [max_batch][n_bins][frames]
  // ---- tcgen05 pyramid (vqt_umma.cu) ----
  bool umma_ok;                    // configuration supported by the level kernels
  VqtLevelDev level[ZNS_VQT_MAX_OCT];
  std::vector<uint16_t>* h_bimg[ZNS_VQT_MAX_OCT];  // host copy of each level's coefficient image (tests)
  uint16_t* d_bimg[ZNS_VQT_MAX_OCT];
  uint16_t* d_hi[ZNS_VQT_MAX_OCT];  // level >= 1 signal, leading fp16 term  [max_batch][stride]
  uint16_t* d_lo[ZNS_VQT_MAX_OCT];  // residual * 2048
  long long sig_stride[ZNS_VQT_MAX_OCT];   // halfwords per clip of a level buffer (tiled layout incl. the two zero pad tiles)
  long long sig_cap[ZNS_VQT_MAX_OCT];      // samples per clip a level buffer holds
  int umma_last_n;                 // n_samples of the previous forward (a shorter signal needs the buffer tails cleared)
};

// vqt_umma.cu
float vqt_coef_scale(const float* re, const float* im, int n);
int vqt_umma_build(zns_vqt_plan* p, double fmin, double gamma_in);   // after the basis has been computed
void vqt_umma_free(zns_vqt_plan* p);
int vqt_umma_forward(zns_vqt_plan* p, const float* y, int batch, int n_samples, float* out, void* stream);
