/*
 * kk_oracle.c — C twin of oracle/oracle.py's value arithmetic plus a multi-threaded CPU loader.
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; never by the product.
 *
 * PARITY UNPINNED AGAINST THE REFERENCE (eminwux/kukeon has no weight loader — SURVEY.md §0); pinned
 * against safetensors 0.7.0 / gguf 0.19.0 (gguf/quants.py:220-572) / torch RNE through oracle.py, which
 * tests/test_oracle_values.py checks this file against bit for bit.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -ffp-contract=off).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint16_t f32bits_to_bf16(uint32_t u) {
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0x7FFF; /* NaN -> canonical (PTX cvt.rn.bf16.f32) */
  uint32_t lsb = (u >> 16) & 1u;
  return (uint16_t)((u + 0x7FFFu + lsb) >> 16);
}
static inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return f32bits_to_bf16(u);
}
static inline float f16bits_to_f32(uint16_t h) {
  _Float16 x;
  memcpy(&x, &h, 2);
  return (float)x;
}

void orc_f32_to_bf16(const uint32_t* src, uint16_t* dst, uint64_t n) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; ++i) dst[i] = f32bits_to_bf16(src[i]);
}

void orc_f16_to_bf16(const uint16_t* src, uint16_t* dst, uint64_t n) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; ++i) dst[i] = f32_to_bf16(f16bits_to_f32(src[i]));
}

/* FP8 -> bf16 (exact widenings; opt-in KK_LOAD_F8_TO_BF16).  E4M3 "fn": bias 7, no infinities, S.1111.111 = NaN. */
static inline uint16_t f8e4m3_to_bf16(uint8_t b) {
  const uint16_t s = (uint16_t)((b >> 7) << 15);
  const unsigned e = (b >> 3) & 15u, m = b & 7u;
  if (e == 15 && m == 7) return 0x7FFF;
  if (e == 0) return (uint16_t)(s | f32_to_bf16((float)m * 0.001953125f)); /* subnormal: m * 2^-9 */
  return (uint16_t)(s | ((e + 120u) << 7) | (m << 4));
}
static inline uint16_t f8e5m2_to_bf16(uint8_t b) { return f32_to_bf16(f16bits_to_f32((uint16_t)((uint16_t)b << 8))); }

/* kind: 0 = E4M3, 1 = E5M2 */
void orc_f8_to_bf16(int kind, const uint8_t* src, uint16_t* dst, uint64_t n) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; ++i) dst[i] = kind ? f8e5m2_to_bf16(src[i]) : f8e4m3_to_bf16(src[i]);
}

/* One Q4_K super-block: d f16 | dmin f16 | scales[12] | qs[128] -> 256 bf16.
 * y = (d*sc_j)*q - (dmin*m_j); every operation rounds to fp32 (no FMA: -ffp-contract=off). */
static void q4k_block(const uint8_t* b, uint16_t* out) {
  uint16_t hd, hm;
  memcpy(&hd, b, 2);
  memcpy(&hm, b + 2, 2);
  const float d = f16bits_to_f32(hd), dmin = f16bits_to_f32(hm);
  const uint8_t* s = b + 4;
  const uint8_t* qs = b + 16;
  for (int j = 0; j < 8; ++j) {
    uint8_t sc, m;
    if (j < 4) {
      sc = s[j] & 63;
      m = s[j + 4] & 63;
    } else {
      sc = (uint8_t)((s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4));
      m = (uint8_t)((s[j + 4] >> 4) | ((s[j] >> 6) << 4));
    }
    volatile float dsc = d * (float)sc;
    volatile float dmn = dmin * (float)m;
    const uint8_t* q = qs + 32 * (j >> 1);
    const int sh = (j & 1) * 4;
    for (int i = 0; i < 32; ++i) {
      volatile float p = dsc * (float)((q[i] >> sh) & 0x0F);
      float y = p - dmn;
      out[32 * j + i] = f32_to_bf16(y);
    }
  }
}

void orc_q4k_to_bf16(const uint8_t* blocks, uint16_t* dst, uint64_t nblocks) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < nblocks; ++i) q4k_block(blocks + 144 * i, dst + 256 * i);
}

/* Q8_0 block (34 B): d f16 | 32 x int8; y = q * d (gguf/quants.py:395-401). */
static void q8_0_block(const uint8_t* b, uint16_t* out) {
  uint16_t hd;
  memcpy(&hd, b, 2);
  const float d = f16bits_to_f32(hd);
  for (int i = 0; i < 32; ++i) out[i] = f32_to_bf16((float)(int8_t)b[2 + i] * d);
}
void orc_q8_0_to_bf16(const uint8_t* blocks, uint16_t* dst, uint64_t nblocks) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < nblocks; ++i) q8_0_block(blocks + 34 * i, dst + 32 * i);
}

/* Q6_K super-block (210 B): ql[128] | qh[64] | scales[16] int8 | d f16 (gguf/quants.py:552-572). */
static void q6k_block(const uint8_t* b, uint16_t* out) {
  const uint8_t* ql = b;
  const uint8_t* qh = b + 128;
  const int8_t* sc = (const int8_t*)(b + 192);
  uint16_t hd;
  memcpy(&hd, b + 208, 2);
  const float d = f16bits_to_f32(hd);
  for (int g = 0; g < 8; ++g)
    for (int i = 0; i < 32; ++i) {
      const int e = 32 * g + i;
      const int lo = (ql[64 * (g / 4) + 32 * (g % 2) + i] >> (4 * ((g % 4) / 2))) & 0x0F;
      const int hi = (qh[32 * (g / 4) + i] >> (2 * (g % 4))) & 0x03;
      const int q = (lo | (hi << 4)) - 32;
      volatile float dsc = d * (float)sc[e / 16];
      out[e] = f32_to_bf16(dsc * (float)q);
    }
}
void orc_q6k_to_bf16(const uint8_t* blocks, uint16_t* dst, uint64_t nblocks) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < nblocks; ++i) q6k_block(blocks + 210 * i, dst + 256 * i);
}

/* ---- the remaining legacy and K quants; element formulas as in oracle/oracle.py ---------- */
static inline float ld_f16(const uint8_t* p) {
  uint16_t h;
  memcpy(&h, p, 2);
  return f16bits_to_f32(h);
}
/* 32-weight legacy blocks: element e < 16 is the low nibble of qs[e], e >= 16 the high nibble of qs[e-16]. */
static inline int legacy_nibble(const uint8_t* qs, int e) { return e < 16 ? (qs[e] & 0x0F) : (qs[e - 16] >> 4); }

/* Q4_0 (18 B): d f16 | qs[16]; y = d * (q - 8)  (gguf/quants.py:220-231). */
static void q4_0_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b);
  for (int e = 0; e < 32; ++e) out[e] = f32_to_bf16(d * (float)(legacy_nibble(b + 2, e) - 8));
}
/* Q4_1 (20 B): d f16 | m f16 | qs[16]; y = (d*q) + m  (gguf/quants.py:254-267). */
static void q4_1_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b), m = ld_f16(b + 2);
  for (int e = 0; e < 32; ++e) {
    volatile float p = d * (float)legacy_nibble(b + 4, e);
    out[e] = f32_to_bf16(p + m);
  }
}
/* Q5_0 (22 B): d f16 | qh u32 | qs[16]; bit 4 of element e is bit e of qh; y = d * (q - 16)  (gguf/quants.py:291-308). */
static void q5_0_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b);
  uint32_t qh;
  memcpy(&qh, b + 2, 4);
  for (int e = 0; e < 32; ++e) {
    const int q = (legacy_nibble(b + 6, e) | (int)(((qh >> e) & 1u) << 4)) - 16;
    out[e] = f32_to_bf16(d * (float)q);
  }
}
/* Q5_1 (24 B): d f16 | m f16 | qh u32 | qs[16]; y = (d*q) + m  (gguf/quants.py:333-352). */
static void q5_1_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b), m = ld_f16(b + 2);
  uint32_t qh;
  memcpy(&qh, b + 4, 4);
  for (int e = 0; e < 32; ++e) {
    const int q = legacy_nibble(b + 8, e) | (int)(((qh >> e) & 1u) << 4);
    volatile float p = d * (float)q;
    out[e] = f32_to_bf16(p + m);
  }
}
/* Q2_K (84 B): scales[16] | qs[64] | d f16 | dmin f16; e = 128h + 32s + i: q = (qs[32h+i] >> 2s) & 3;
 * y = (d*(scales[e/16]&15))*q - dmin*(scales[e/16]>>4)  (gguf/quants.py:404-428). */
static void q2k_block(const uint8_t* b, uint16_t* out) {
  const uint8_t* sc = b;
  const uint8_t* qs = b + 16;
  const float d = ld_f16(b + 80), dmin = ld_f16(b + 82);
  for (int e = 0; e < 256; ++e) {
    const int h = e >> 7, s = (e >> 5) & 3, i = e & 31;
    const int q = (qs[32 * h + i] >> (2 * s)) & 3;
    volatile float dl = d * (float)(sc[e >> 4] & 0x0F);
    volatile float ml = dmin * (float)(sc[e >> 4] >> 4);
    volatile float p = dl * (float)q;
    out[e] = f32_to_bf16(p - ml);
  }
}
/* Q3_K (110 B): hmask[32] | qs[64] | scales[12] | d f16  (gguf/quants.py:431-472). */
static void q3k_block(const uint8_t* b, uint16_t* out) {
  const uint8_t* hm = b;
  const uint8_t* qs = b + 32;
  const uint8_t* s = b + 96;
  const float d = ld_f16(b + 108);
  for (int e = 0; e < 256; ++e) {
    const int k = e >> 4, g = e >> 5, i = e & 31;
    const int lo4 = k < 8 ? (s[k] & 0x0F) : (s[k - 8] >> 4);
    const int hi2 = (s[8 + (k & 3)] >> (2 * (k >> 2))) & 3;
    const int scale = (lo4 | (hi2 << 4)) - 32;
    const int lo = (qs[32 * (g >> 2) + i] >> (2 * (g & 3))) & 3;
    const int q = lo - ((((hm[i] >> g) & 1) ^ 1) << 2);
    volatile float dl = d * (float)scale;
    out[e] = f32_to_bf16(dl * (float)q);
  }
}
/* Q5_K (176 B): d f16 | dmin f16 | scales[12] | qh[32] | qs[128]; scales packed as in Q4_K  (gguf/quants.py:525-548). */
static void q5k_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b), dmin = ld_f16(b + 2);
  const uint8_t* s = b + 4;
  const uint8_t* qh = b + 16;
  const uint8_t* qs = b + 48;
  for (int j = 0; j < 8; ++j) {
    uint8_t sc, m;
    if (j < 4) {
      sc = s[j] & 63;
      m = s[j + 4] & 63;
    } else {
      sc = (uint8_t)((s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4));
      m = (uint8_t)((s[j + 4] >> 4) | ((s[j] >> 6) << 4));
    }
    volatile float dsc = d * (float)sc;
    volatile float dmn = dmin * (float)m;
    for (int i = 0; i < 32; ++i) {
      const int q = ((qs[32 * (j >> 1) + i] >> (4 * (j & 1))) & 0x0F) | (((qh[i] >> j) & 1) << 4);
      volatile float p = dsc * (float)q;
      out[32 * j + i] = f32_to_bf16(p - dmn);
    }
  }
}

/* IQ4_NL (18 B): Q4_0's layout with a 16-entry codebook; IQ4_XS (136 B): d f16 | scales_h u16 | scales_l[4] | qs[128], eight
 * 32-weight sub-blocks with 6-bit scales; MXFP4 (17 B): E8M0 scale byte | qs[16] (gguf/quants.py:1330-1380, 656-708). */
static const int8_t kIQ4NL[16] = {-127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113};
static const int8_t kMXFP4[16] = {0, 1, 2, 3, 4, 6, 8, 12, 0, -1, -2, -3, -4, -6, -8, -12};
static void iq4nl_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b);
  for (int e = 0; e < 32; ++e) out[e] = f32_to_bf16(d * (float)kIQ4NL[legacy_nibble(b + 2, e)]);
}
static void iq4xs_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b);
  uint16_t sh;
  memcpy(&sh, b + 2, 2);
  const uint8_t* sl = b + 4;
  const uint8_t* qs = b + 8;
  for (int j = 0; j < 8; ++j) {
    const int ls = ((sl[j >> 1] >> (4 * (j & 1))) & 0x0F) | (((sh >> (2 * j)) & 3) << 4);
    volatile float dl = d * (float)(ls - 32);
    for (int i = 0; i < 32; ++i) out[32 * j + i] = f32_to_bf16(dl * (float)kIQ4NL[legacy_nibble(qs + 16 * j, i)]);
  }
}
static void mxfp4_block(const uint8_t* b, uint16_t* out) {
  const uint32_t e = b[0];
  const uint32_t bits = e < 2 ? (0x00200000u << e) : ((e - 1u) << 23); /* half the E8M0 scale */
  float d;
  memcpy(&d, &bits, 4);
  for (int i = 0; i < 32; ++i) out[i] = f32_to_bf16(d * (float)kMXFP4[legacy_nibble(b + 1, i)]);
}

/* ---- lattice i-quants, ternary types, NVFP4 (element formulas as in oracle/oracle.py; codebooks: oracle/iq_grids_c.h, generated) ---- */
#include "iq_grids_c.h"
static const uint64_t g_iq2xxs[KK_GRID_IQ2XXS_SIZE] = {KK_GRID_IQ2XXS_VALUES};
static const uint64_t g_iq2xs[KK_GRID_IQ2XS_SIZE] = {KK_GRID_IQ2XS_VALUES};
static const uint64_t g_iq2s[KK_GRID_IQ2S_SIZE] = {KK_GRID_IQ2S_VALUES};
static const uint32_t g_iq3xxs[KK_GRID_IQ3XXS_SIZE] = {KK_GRID_IQ3XXS_VALUES};
static const uint32_t g_iq3s[KK_GRID_IQ3S_SIZE] = {KK_GRID_IQ3S_VALUES};
static const uint64_t g_iq1s[KK_GRID_IQ1S_SIZE] = {KK_GRID_IQ1S_VALUES};

static inline uint32_t ld_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t ld_u16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
/* 7 stored sign bits, the 8th is their parity (ggml ksigns_iq2xs) */
static inline uint32_t ksigns(uint32_t i7) { return i7 | ((uint32_t)(__builtin_popcount(i7) & 1) << 7); }
/* eight weights: (db * grid byte k) * (+1 | -1 by bit k of signs) */
static inline void put8(uint16_t* out, float db, uint64_t grid, uint32_t signs) {
  for (int k = 0; k < 8; ++k) {
    volatile float p = db * (float)((grid >> (8 * k)) & 0xFF);
    out[k] = f32_to_bf16((signs >> k) & 1 ? -p : p);
  }
}
static void iq2xxs_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b);
  for (int g = 0; g < 8; ++g) {
    const uint32_t q0 = ld_u32(b + 2 + 8 * g), q1 = ld_u32(b + 6 + 8 * g);
    volatile float t = d * (0.5f + (float)(q1 >> 28));
    const float db = t * 0.25f;
    for (int k = 0; k < 4; ++k) put8(out + 32 * g + 8 * k, db, g_iq2xxs[(q0 >> (8 * k)) & 0xFF], ksigns((q1 >> (7 * k)) & 0x7F));
  }
}
static void iq2xs_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b);
  for (int j = 0; j < 32; ++j) {
    const uint32_t q = ld_u16(b + 2 + 2 * j);
    const uint32_t s = (b[66 + (j >> 2)] >> (4 * ((j >> 1) & 1))) & 0x0F;
    volatile float t = d * (0.5f + (float)s);
    put8(out + 8 * j, t * 0.25f, g_iq2xs[q & 511], ksigns(q >> 9));
  }
}
static void iq2s_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b);
  for (int j = 0; j < 32; ++j) {
    const uint32_t idx = b[2 + j] | (((b[66 + (j >> 2)] >> (2 * (j & 3))) & 3u) << 8);
    const uint32_t s = (b[74 + (j >> 2)] >> (4 * ((j >> 1) & 1))) & 0x0F;
    volatile float t = d * (0.5f + (float)s);
    put8(out + 8 * j, t * 0.25f, g_iq2s[idx], b[34 + j]);
  }
}
static void iq3xxs_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b);
  for (int g = 0; g < 8; ++g) {
    const uint32_t w = ld_u32(b + 66 + 4 * g);
    volatile float t = d * (0.5f + (float)(w >> 28));
    const float db = t * 0.5f;
    for (int k = 0; k < 4; ++k) {
      const uint64_t grid = (uint64_t)g_iq3xxs[b[2 + 8 * g + 2 * k]] | ((uint64_t)g_iq3xxs[b[2 + 8 * g + 2 * k + 1]] << 32);
      put8(out + 32 * g + 8 * k, db, grid, ksigns((w >> (7 * k)) & 0x7F));
    }
  }
}
static void iq3s_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b);
  for (int m = 0; m < 32; ++m) { /* 8 weights = two 4-value grid entries */
    const int j0 = 2 * m, j1 = 2 * m + 1, g = m >> 2;
    const uint32_t i0 = b[2 + j0] | (((b[66 + (j0 >> 3)] >> (j0 & 7)) & 1u) << 8);
    const uint32_t i1 = b[2 + j1] | (((b[66 + (j1 >> 3)] >> (j1 & 7)) & 1u) << 8);
    const uint32_t s = (b[106 + (g >> 1)] >> (4 * (g & 1))) & 0x0F;
    const float db = d * (float)(1 + 2 * (int)s);
    put8(out + 8 * m, db, (uint64_t)g_iq3s[i0] | ((uint64_t)g_iq3s[i1] << 32), b[74 + m]);
  }
}
/* IQ1: grid bytes are value + 1; y = dl * ((value) + delta) */
static inline void put8_iq1(uint16_t* out, float dl, uint64_t grid, float delta) {
  for (int k = 0; k < 8; ++k) {
    volatile float v = (float)((int)((grid >> (8 * k)) & 0xFF) - 1) + delta;
    out[k] = f32_to_bf16(dl * v);
  }
}
static void iq1s_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b);
  for (int g = 0; g < 8; ++g) {
    const uint32_t qh = ld_u16(b + 34 + 2 * g);
    const float dl = d * (float)(2 * (int)((qh >> 12) & 7) + 1);
    const float delta = (qh & 0x8000) ? -0.125f : 0.125f;
    for (int k = 0; k < 4; ++k) put8_iq1(out + 32 * g + 8 * k, dl, g_iq1s[b[2 + 4 * g + k] | (((qh >> (3 * k)) & 7u) << 8)], delta);
  }
}
static void iq1m_block(const uint8_t* b, uint16_t* out) {
  uint16_t sc[4];
  for (int i = 0; i < 4; ++i) sc[i] = ld_u16(b + 48 + 2 * i);
  const uint16_t dbits = (uint16_t)((sc[0] >> 12) | ((sc[1] >> 12) << 4) | ((sc[2] >> 12) << 8) | ((sc[3] >> 12) << 12));
  const float d = f16bits_to_f32(dbits);
  for (int j = 0; j < 32; ++j) {
    const int k = j >> 1;
    const uint32_t s3 = (sc[k >> 2] >> (3 * (k & 3))) & 7u;
    const float dl = d * (float)(2 * (int)s3 + 1);
    const uint32_t nib = (b[32 + (j >> 1)] >> (4 * (j & 1))) & 0x0F;
    put8_iq1(out + 8 * j, dl, g_iq1s[b[j] | ((nib & 7u) << 8)], (nib & 8) ? -0.125f : 0.125f);
  }
}
static void tq2_0_block(const uint8_t* b, uint16_t* out) {
  const float d = ld_f16(b + 64);
  for (int e = 0; e < 256; ++e) {
    const int h = e >> 7, s = (e >> 5) & 3, i = e & 31;
    out[e] = f32_to_bf16(d * (float)((int)((b[32 * h + i] >> (2 * s)) & 3) - 1));
  }
}
static void tq1_0_block(const uint8_t* b, uint16_t* out) {
  static const uint32_t pow3[5] = {1, 3, 9, 27, 81};
  const float d = ld_f16(b + 52);
  for (int e = 0; e < 256; ++e) {
    int byte, p;
    if (e < 160) { byte = e % 32; p = e / 32; }
    else if (e < 240) { byte = 32 + (e - 160) % 16; p = (e - 160) / 16; }
    else { byte = 48 + (e - 240) % 4; p = (e - 240) / 4; }
    const uint32_t v = (b[byte] * pow3[p]) & 0xFF;
    out[e] = f32_to_bf16(d * (float)((int)((v * 3) >> 8) - 1));
  }
}
static void nvfp4_block(const uint8_t* b, uint16_t* out) {
  for (int s = 0; s < 4; ++s) {
    const uint32_t x = b[s], e = (x >> 3) & 0xF, m = x & 7;
    float d; /* HALF the unsigned-E4M3 scale; 0x00 and 0x7F decode to 0 */
    if (x == 0 || x == 0x7F) d = 0.0f;
    else if (e == 0) d = (float)m * 0.0009765625f; /* m * 2^-9 * 0.5 */
    else { const uint32_t bits = ((e + 119u) << 23) | (m << 20); memcpy(&d, &bits, 4); } /* (1 + m/8) * 2^(e-8) */
    const uint8_t* qs = b + 4 + 8 * s;
    for (int i = 0; i < 16; ++i) out[16 * s + i] = f32_to_bf16(d * (float)kMXFP4[i < 8 ? (qs[i] & 0x0F) : (qs[i - 8] >> 4)]);
  }
}

/* Dispatch by ggml type id (gguf/constants.py GGMLQuantizationType): block bytes / weights per block / function. */
typedef void (*orc_block_fn)(const uint8_t*, uint16_t*);
static orc_block_fn block_fn(uint32_t ggml_type, uint32_t* bytes, uint32_t* elems) {
  switch (ggml_type) {
    case 2: *bytes = 18; *elems = 32; return q4_0_block;
    case 3: *bytes = 20; *elems = 32; return q4_1_block;
    case 6: *bytes = 22; *elems = 32; return q5_0_block;
    case 7: *bytes = 24; *elems = 32; return q5_1_block;
    case 8: *bytes = 34; *elems = 32; return q8_0_block;
    case 10: *bytes = 84; *elems = 256; return q2k_block;
    case 11: *bytes = 110; *elems = 256; return q3k_block;
    case 12: *bytes = 144; *elems = 256; return q4k_block;
    case 13: *bytes = 176; *elems = 256; return q5k_block;
    case 14: *bytes = 210; *elems = 256; return q6k_block;
    case 20: *bytes = 18; *elems = 32; return iq4nl_block;
    case 23: *bytes = 136; *elems = 256; return iq4xs_block;
    case 39: *bytes = 17; *elems = 32; return mxfp4_block;
    case 16: *bytes = 66; *elems = 256; return iq2xxs_block;
    case 17: *bytes = 74; *elems = 256; return iq2xs_block;
    case 18: *bytes = 98; *elems = 256; return iq3xxs_block;
    case 19: *bytes = 50; *elems = 256; return iq1s_block;
    case 21: *bytes = 110; *elems = 256; return iq3s_block;
    case 22: *bytes = 82; *elems = 256; return iq2s_block;
    case 29: *bytes = 56; *elems = 256; return iq1m_block;
    case 34: *bytes = 54; *elems = 256; return tq1_0_block;
    case 35: *bytes = 66; *elems = 256; return tq2_0_block;
    case 40: *bytes = 36; *elems = 64; return nvfp4_block;
    default: return 0;
  }
}
/* Returns 0, or -1 for a type this oracle does not define. */
int orc_dequant_to_bf16(uint32_t ggml_type, const uint8_t* blocks, uint16_t* dst, uint64_t nblocks) {
  uint32_t nb = 0, ne = 0;
  const orc_block_fn fn = block_fn(ggml_type, &nb, &ne);
  if (!fn) return -1;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < nblocks; ++i) fn(blocks + (uint64_t)nb * i, dst + (uint64_t)ne * i);
  return 0;
}

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

uint64_t orc_checksum(const uint8_t* p, uint64_t nbytes) {
  const uint64_t nw = nbytes >> 3;
  uint64_t acc = 0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
  for (uint64_t i = 0; i < nw; ++i) {
    uint64_t w;
    memcpy(&w, p + 8 * i, 8);
    acc += mix64(w + i * 0x9E3779B97F4A7C15ull);
  }
  if (nbytes & 7) {
    uint64_t last = 0;
    memcpy(&last, p + 8 * nw, nbytes & 7);
    acc += mix64(last + nw * 0x9E3779B97F4A7C15ull);
  }
  return acc;
}

/* ---- synthetic content (counter-based, reproducible per (seed, index)) ------------------------ */
static inline uint64_t ctr_rand(uint64_t seed, uint64_t i) { return mix64(seed * 0xD1B54A32D192ED03ull + i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull); }

void orc_fill_bytes(uint8_t* dst, uint64_t n, uint64_t seed) {
  const uint64_t nw = n >> 3;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < nw; ++i) {
    uint64_t r = ctr_rand(seed, i);
    memcpy(dst + 8 * i, &r, 8);
  }
  if (n & 7) {
    uint64_t r = ctr_rand(seed, nw);
    memcpy(dst + 8 * nw, &r, n & 7);
  }
}

/* Uniform 16-bit patterns with the exponent clamped so every value is a finite bf16. */
void orc_fill_bf16_finite(uint16_t* dst, uint64_t n, uint64_t seed) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < (n + 3) / 4; ++i) {
    uint64_t r = ctr_rand(seed, i);
    for (int k = 0; k < 4 && 4 * i + k < n; ++k) {
      uint16_t v = (uint16_t)(r >> (16 * k));
      if ((v & 0x7F80) == 0x7F80) v &= (uint16_t)~0x0080; /* exponent 0xFF -> 0xFE */
      dst[4 * i + k] = v;
    }
  }
}

/* fp32 values ~ U(-0.04, 0.04) plus a sprinkle of exact ties/denormals; all finite. */
void orc_fill_f32(float* dst, uint64_t n, uint64_t seed) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t r = ctr_rand(seed, i);
    uint32_t u = (uint32_t)r;
    float f = ((float)(u >> 8) * (1.0f / 16777216.0f) - 0.5f) * 0.08f;
    if ((r >> 60) == 0) { /* 1/16 of the values: raw finite bit pattern (exercises RNE ties, subnormals) */
      uint32_t b = (uint32_t)(r >> 16);
      if ((b & 0x7F800000u) == 0x7F800000u) b &= ~0x00800000u;
      memcpy(&f, &b, 4);
    }
    dst[i] = f;
  }
}

/* Q4_K blocks: random bytes, d and dmin overwritten with finite fp16 in [2^-10, 2^-4]. */
void orc_fill_q4k(uint8_t* dst, uint64_t nblocks, uint64_t seed) {
#pragma omp parallel for schedule(static)
  for (uint64_t b = 0; b < nblocks; ++b) {
    uint8_t* p = dst + 144 * b;
    for (int w = 0; w < 18; ++w) {
      uint64_t r = ctr_rand(seed, b * 18 + w);
      memcpy(p + 8 * w, &r, 8);
    }
    uint64_t r = ctr_rand(seed ^ 0xABCDEF, b);
    /* fp16: exponent field e in [5, 11] => 2^(e-15) in [2^-10, 2^-4]; random mantissa; positive */
    uint16_t d = (uint16_t)((((r & 0xFF) % 7 + 5) << 10) | ((r >> 8) & 0x3FF));
    uint16_t m = (uint16_t)(((((r >> 20) & 0xFF) % 7 + 5) << 10) | ((r >> 28) & 0x3FF));
    memcpy(p, &d, 2);
    memcpy(p + 2, &m, 2);
  }
}

/* ---- CPU loader ("port" of the hot path for the cpu_baseline legs) ---------------------------- */
enum { ORC_COPY = 0, ORC_F32_BF16 = 1, ORC_F16_BF16 = 2, ORC_Q4K_BF16 = 3, ORC_Q8_0_BF16 = 4, ORC_Q6K_BF16 = 5, ORC_F8E4M3_BF16 = 6, ORC_F8E5M2_BF16 = 7,
       ORC_DEQUANT = 0x100 /* | ggml type id: any block-quantised type block_fn() knows */ };

typedef struct {
  uint32_t shard;
  uint32_t op;
  uint64_t file_off;
  uint64_t nbytes; /* source bytes */
  uint64_t dst_off;
} orc_job;

static int pread_full(int fd, uint8_t* dst, uint64_t len, uint64_t off) {
  while (len) {
    ssize_t r = pread(fd, dst, len, (off_t)off);
    if (r < 0) {
      if (errno == EINTR) continue;
      return -1;
    }
    if (r == 0) return -1;
    dst += r; off += (uint64_t)r; len -= (uint64_t)r;
  }
  return 0;
}

/* Reads every job's bytes from its shard and writes the converted bf16 (or verbatim bytes) into `pool`
 * (host memory).  Jobs must be small enough for a per-thread scratch buffer of `scratch_bytes`.
 * Returns 0, or -1 on I/O error.  threads <= 0: OpenMP default. */
int orc_cpu_load(const char* const* shard_paths, uint32_t n_shards, const orc_job* jobs, uint64_t n_jobs, uint8_t* pool,
                 uint64_t scratch_bytes, int threads) {
  int* fds = (int*)malloc(sizeof(int) * n_shards);
  int err = 0;
  for (uint32_t i = 0; i < n_shards; ++i) {
    fds[i] = open(shard_paths[i], O_RDONLY);
    if (fds[i] < 0) err = -1;
  }
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  if (!err) {
#pragma omp parallel
    {
      uint8_t* scratch = (uint8_t*)malloc(scratch_bytes ? scratch_bytes : 1);
#pragma omp for schedule(static, 1) /* job j always runs on thread j mod T: the pages a thread first-touches are the pages it rewrites in every later pass */
      for (uint64_t j = 0; j < n_jobs; ++j) {
        const orc_job* J = &jobs[j];
        if (J->nbytes > scratch_bytes && J->op != ORC_COPY) { err = -1; continue; }
        if (J->op == ORC_COPY) { /* straight into the pool */
          if (pread_full(fds[J->shard], pool + J->dst_off, J->nbytes, J->file_off)) err = -1;
          continue;
        }
        if (pread_full(fds[J->shard], scratch, J->nbytes, J->file_off)) { err = -1; continue; }
        uint16_t* out = (uint16_t*)(pool + J->dst_off);
        if (J->op == ORC_F32_BF16) {
          for (uint64_t i = 0; i < J->nbytes / 4; ++i) { uint32_t u; memcpy(&u, scratch + 4 * i, 4); out[i] = f32bits_to_bf16(u); }
        } else if (J->op == ORC_F16_BF16) {
          for (uint64_t i = 0; i < J->nbytes / 2; ++i) { uint16_t h; memcpy(&h, scratch + 2 * i, 2); out[i] = f32_to_bf16(f16bits_to_f32(h)); }
        } else if (J->op == ORC_Q4K_BF16) {
          for (uint64_t i = 0; i < J->nbytes / 144; ++i) q4k_block(scratch + 144 * i, out + 256 * i);
        } else if (J->op == ORC_Q8_0_BF16) {
          for (uint64_t i = 0; i < J->nbytes / 34; ++i) q8_0_block(scratch + 34 * i, out + 32 * i);
        } else if (J->op == ORC_Q6K_BF16) {
          for (uint64_t i = 0; i < J->nbytes / 210; ++i) q6k_block(scratch + 210 * i, out + 256 * i);
        } else if (J->op == ORC_F8E4M3_BF16) {
          for (uint64_t i = 0; i < J->nbytes; ++i) out[i] = f8e4m3_to_bf16(scratch[i]);
        } else if (J->op == ORC_F8E5M2_BF16) {
          for (uint64_t i = 0; i < J->nbytes; ++i) out[i] = f8e5m2_to_bf16(scratch[i]);
        } else if (J->op & ORC_DEQUANT) {
          uint32_t nb = 0, ne = 0;
          const orc_block_fn fn = block_fn(J->op & 0xFFu, &nb, &ne);
          if (!fn) { err = -1; continue; }
          for (uint64_t i = 0; i < J->nbytes / nb; ++i) fn(scratch + (uint64_t)nb * i, out + (uint64_t)ne * i);
        } else err = -1;
      }
      free(scratch);
    }
  }
  for (uint32_t i = 0; i < n_shards; ++i)
    if (fds[i] >= 0) close(fds[i]);
  free(fds);
  return err;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
