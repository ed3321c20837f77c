/*
 * synth_fill.c — fast seeded content generator for tools/synth.py (synthetic checkpoints for tests and
 * bench.py).  Not part of the product and not part of the oracle.  Content of byte range
 * [off, off+n) of tensor `idx` depends only on (seed, idx, position), so any piece can be regenerated.
 * Built by tools/Makefile into tools/_build/libkk_synth.so.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { K_BYTES = 0, K_BF16 = 1, K_F16 = 2, K_F32 = 3, K_Q4K = 4, K_Q8_0 = 5, K_Q6K = 6,
       K_Q4_0 = 7, K_Q4_1 = 8, K_Q5_0 = 9, K_Q5_1 = 10, K_Q2K = 11, K_Q3K = 12, K_Q5K = 13,
       K_IQ4NL = 14, K_IQ4XS = 15, K_MXFP4 = 16,
       K_IQ2XXS = 17, K_IQ2XS = 18, K_IQ2S = 19, K_IQ3XXS = 20, K_IQ3S = 21, K_IQ1S = 22, K_IQ1M = 23, K_TQ1_0 = 24, K_TQ2_0 = 25, K_NVFP4 = 26 };

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
static inline uint64_t rnd(uint64_t seed, uint64_t idx, uint64_t i) {
  return mix64((seed + 0x1234567ull) * 0xD1B54A32D192ED03ull + idx * 0xA0761D6478BD642Full + i * 0x9E3779B97F4A7C15ull);
}

/* fill words [w0, w0+nw) (8-byte words of the tensor) into dst */
static void fill_words(uint8_t* dst, uint64_t w0, uint64_t nw, int kind, uint64_t seed, uint64_t idx) {
  for (uint64_t k = 0; k < nw; ++k) {
    uint64_t r = rnd(seed, idx, w0 + k);
    if (kind == K_BF16) { /* clear exponent 0xFF -> 0xFE in each 16-bit lane: all finite */
      uint64_t e = r & 0x7F807F807F807F80ull;
      uint64_t full = (e + 0x0080008000800080ull) & 0x8000800080008000ull; /* lane exponent == 0xFF <=> carry into bit 15 */
      r &= ~(full >> 8);                                                    /* clear bit 7 of those lanes */
    } else if (kind == K_F16) {
      uint64_t e = r & 0x7C007C007C007C00ull;
      uint64_t full = (e + 0x0400040004000400ull) & 0x8000800080008000ull;
      r &= ~(full >> 5);                                                    /* clear bit 10 */
    } else if (kind == K_F32) {
      for (int h = 0; h < 2; ++h) { /* two floats ~ U(-0.04, 0.04) */
        uint32_t u = (uint32_t)(r >> (32 * h));
        float f = ((float)(u >> 8) * (1.0f / 16777216.0f) - 0.5f) * 0.08f;
        memcpy(dst + 8 * k + 4 * h, &f, 4);
      }
      continue;
    }
    memcpy(dst + 8 * k, &r, 8);
  }
}

static void fix_q4k(uint8_t* buf, uint64_t first_block, uint64_t nblocks, uint64_t seed, uint64_t idx) {
  for (uint64_t b = 0; b < nblocks; ++b) {
    uint64_t r = rnd(seed ^ 0xABCDEFull, idx, first_block + b);
    uint16_t d = (uint16_t)((((r & 0xFF) % 7 + 5) << 10) | ((r >> 8) & 0x3FF));          /* 2^-10 .. 2^-4 */
    uint16_t m = (uint16_t)(((((r >> 20) & 0xFF) % 7 + 5) << 10) | ((r >> 28) & 0x3FF));
    memcpy(buf + 144 * b, &d, 2);
    memcpy(buf + 144 * b + 2, &m, 2);
  }
}

/* Block-quantised kinds other than Q4_K: random payload, the fp16 scale `d` (at byte `doff` of every `bsz`-byte block)
 * overwritten with a finite value in [2^-10, 2^-4]. */
static void fix_scale(uint8_t* buf, uint64_t first_block, uint64_t nblocks, uint64_t bsz, uint64_t doff, uint64_t seed, uint64_t idx) {
  for (uint64_t b = 0; b < nblocks; ++b) {
    uint64_t r = rnd(seed ^ 0x5CA1Eull, idx, first_block + b);
    uint16_t d = (uint16_t)((((r & 0xFF) % 7 + 5) << 10) | ((r >> 8) & 0x3FF) | ((r >> 40) & 1 ? 0x8000 : 0));
    memcpy(buf + bsz * b + doff, &d, 2);
  }
}
/* The later kinds: block size, offset of the fp16 scale d and of the second fp16 (min / dmin; -1 when the type has none). */
typedef struct { uint64_t bsz; int d_off, m_off; } blk_geom;
static blk_geom geom_of(int kind) {
  switch (kind) {
    case K_Q4_0: return (blk_geom){18, 0, -1};
    case K_Q4_1: return (blk_geom){20, 0, 2};
    case K_Q5_0: return (blk_geom){22, 0, -1};
    case K_Q5_1: return (blk_geom){24, 0, 2};
    case K_Q2K: return (blk_geom){84, 80, 82};
    case K_Q3K: return (blk_geom){110, 108, -1};
    case K_Q5K: return (blk_geom){176, 0, 2};
    case K_IQ4NL: return (blk_geom){18, 0, -1};
    case K_IQ4XS: return (blk_geom){136, 0, -1};
    case K_IQ2XXS: return (blk_geom){66, 0, -1};
    case K_IQ2XS: return (blk_geom){74, 0, -1};
    case K_IQ2S: return (blk_geom){82, 0, -1};
    case K_IQ3XXS: return (blk_geom){98, 0, -1};
    case K_IQ3S: return (blk_geom){110, 0, -1};
    case K_IQ1S: return (blk_geom){50, 0, -1};
    case K_IQ1M: return (blk_geom){56, -1, -1}; /* fp16 scale spread over four nibbles, fixed up below */
    case K_TQ1_0: return (blk_geom){54, 52, -1};
    case K_TQ2_0: return (blk_geom){66, 64, -1};
    case K_NVFP4: return (blk_geom){36, -1, -1}; /* four UE4M3 scale bytes: any byte is finite */
    case K_MXFP4: return (blk_geom){17, -1, -1}; /* no fp16 scale: byte 0 is an E8M0 exponent, fixed up below */
    default: return (blk_geom){0, 0, -1};
  }
}
static uint64_t block_bytes(int kind) {
  return kind == K_Q4K ? 144 : kind == K_Q8_0 ? 34 : kind == K_Q6K ? 210 : geom_of(kind).bsz;
}
/* Overwrite the fp16 scale(s) of a run of blocks so that every block dequantises to finite values. */
static void fix_blocks(uint8_t* buf, uint64_t first_block, uint64_t nblocks, int kind, uint64_t seed, uint64_t idx) {
  if (kind == K_Q4K) fix_q4k(buf, first_block, nblocks, seed, idx);
  else if (kind == K_Q8_0) fix_scale(buf, first_block, nblocks, 34, 0, seed, idx);
  else if (kind == K_Q6K) fix_scale(buf, first_block, nblocks, 210, 208, seed, idx);
  else {
    const blk_geom g = geom_of(kind);
    if (!g.bsz) return;
    if (kind == K_MXFP4) { /* keep the shared exponent in 2^-20 .. 2^+9 so that every product is a finite, normal fp32 */
      for (uint64_t b = 0; b < nblocks; ++b) buf[g.bsz * b] = (uint8_t)(108 + rnd(seed ^ 0xE8E8ull, idx, first_block + b) % 30);
      return;
    }
    if (kind == K_IQ1M) { /* a finite fp16 in [2^-10, 2^-4], one nibble into the top of each of the four scale words at byte 48 */
      for (uint64_t b = 0; b < nblocks; ++b) {
        uint64_t r = rnd(seed ^ 0x1713ull, idx, first_block + b);
        uint16_t d = (uint16_t)((((r & 0xFF) % 7 + 5) << 10) | ((r >> 8) & 0x3FF));
        for (int i = 0; i < 4; ++i) {
          uint8_t* hi = buf + g.bsz * b + 48 + 2 * i + 1;
          *hi = (uint8_t)((*hi & 0x0F) | (((d >> (4 * i)) & 0xF) << 4));
        }
      }
      return;
    }
    if (g.d_off < 0) return;
    fix_scale(buf, first_block, nblocks, g.bsz, (uint64_t)g.d_off, seed, idx);
    if (g.m_off >= 0) fix_scale(buf, first_block, nblocks, g.bsz, (uint64_t)g.m_off, seed ^ 0x3117ull, idx);
  }
}

/* Generate bytes [0, nbytes) of tensor idx and pwrite them at file_off. nbytes % 8 may be non-zero.
 * Q4_K: nbytes must be a multiple of 144.  Returns 0 or -errno. */
int synth_write(int fd, uint64_t file_off, uint64_t nbytes, int kind, uint64_t seed, uint64_t idx) {
  const uint64_t bsz = block_bytes(kind);
  const uint64_t CH = kind == K_Q4K ? (144ull * 8 * 7168) : kind == K_Q8_0 ? (136ull * 61440) : kind == K_Q6K ? (840ull * 9984)
                      : bsz ? (bsz * 8) * ((8ull << 20) / (bsz * 8))
                            : (8ull << 20); /* multiple of 8 and of the block size */
  const uint64_t nch = (nbytes + CH - 1) / CH;
  int err = 0;
  /* no more threads than chunks: a 128-thread team spun up for a 3 KB bias vector costs far more than the vector (GPT-2's 148 tensors took
   * 14 s on the 128-thread GPU host, 2.5 s on 8 cores) */
  int team = 1;
#ifdef _OPENMP
  team = omp_get_max_threads();
#endif
  if ((uint64_t)team > nch) team = (int)(nch ? nch : 1);
#pragma omp parallel num_threads(team)
  {
    uint8_t* buf = (uint8_t*)malloc(CH + 8);
#pragma omp for schedule(dynamic, 1)
    for (uint64_t c = 0; c < nch; ++c) {
      const uint64_t b0 = c * CH;
      const uint64_t n = nbytes - b0 < CH ? nbytes - b0 : CH;
      fill_words(buf, b0 / 8, (n + 7) / 8, bsz ? K_BYTES : kind, seed, idx);
      if (bsz) fix_blocks(buf, b0 / bsz, n / bsz, kind, seed, idx);
      uint64_t done = 0;
      while (done < n) {
        ssize_t w = pwrite(fd, buf + done, n - done, (off_t)(file_off + b0 + done));
        if (w < 0) {
          if (errno == EINTR) continue;
          err = -errno;
          break;
        }
        done += (uint64_t)w;
      }
    }
    free(buf);
  }
  return err;
}

/* Same content into memory (tests regenerate pieces without a file). */
void synth_fill(uint8_t* dst, uint64_t nbytes, int kind, uint64_t seed, uint64_t idx) {
  uint64_t nw = nbytes / 8;
  const int wk = block_bytes(kind) ? K_BYTES : kind;
  fill_words(dst, 0, nw, wk, seed, idx);
  if (nbytes & 7) {
    uint8_t tmp[8];
    fill_words(tmp, nw, 1, wk, seed, idx);
    memcpy(dst + 8 * nw, tmp, nbytes & 7);
  }
  if (block_bytes(kind)) fix_blocks(dst, 0, nbytes / block_bytes(kind), kind, seed, idx);
}

/* torchrun exports OMP_NUM_THREADS=1; callers that want all cores say so explicitly. */
void synth_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
