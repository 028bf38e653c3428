// Plan of the VQT / CQT front-end (shared by vqt.cu and vqt_umma.cu).
#pragma once
#include <stdint.h>
#include <vector>

#define ZNS_VQT_MAX_OCT 10
#define ZNS_VQT_MAX_MMA 112

// One tcgen05.mma of a level tile (vqt_umma.cu): operand offsets are bytes from the term's plane base
// (A) and from the level's coefficient image (B); both operands use the no-swizzle K-major layout with
// rows at 16-byte pitch (SBO = 128).
struct VqtMma {
  uint32_t a_off;
  uint32_t b_off;
  uint16_t n;       // MMA N (multiple of 16)
  uint16_t d_col;   // TMEM column of the accumulator window
  uint16_t term;    // 0: x1 (leading fp16 term of the signal), 1: x2 (residual * 2048)
  uint16_t b_rows;  // rows of the coefficient tile this window lives in (B LBO = 16 * b_rows)
};

// Geometry + MMA list of one pyramid level, passed to the kernel by value.
struct VqtLevelDev {
  int q;            // 16-byte chunk planes per row (row = 8 q samples)
  int hb, ha;       // halo rows before / after the 128 tile rows
  int rtot;         // rows per plane in shared memory (128 + hb + ha, padded against bank conflicts)
  int a_lbo;        // A descriptor leading-dimension byte offset (plane stride; 16 when q == 1)
  int fpr;          // frames per row
  int hop, n_fft;   // at this level's rate
  int bin0;         // first output bin of this octave
  int dec_w;        // decimator outputs per row (4 q); 0 on the last level
  int wacc;         // accumulator columns per decimator accumulator
  int dec_a_col, dec_b_col, fb_a_col, fb_b_col, fb_b_stride, tmem_cols;
  int n_mma;
  int b_bytes;      // coefficient image size
  float dec_scale;  // sqrt(2) / (tap scale)
  float fb_scale;   // 1 / (coefficient scale)
  VqtMma mma[ZNS_VQT_MAX_MMA];
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
  long long sig_stride[ZNS_VQT_MAX_OCT];
};

// vqt_umma.cu
float vqt_coef_scale(const float* re, const float* im, int n);
int vqt_umma_build(zns_vqt_plan* p, double fmin, double gamma_in);   // after the basis has been computed
void vqt_umma_free(zns_vqt_plan* p);
int vqt_umma_forward(zns_vqt_plan* p, const float* y, int batch, int n_samples, float* out, void* stream);
