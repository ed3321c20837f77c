// Shared between the host planner and the device kernels: the per-launch segment table and the
// tile geometry.  A *segment* is one contiguous run of one tensor inside one staged chunk; a *tile*
// is the unit a CTA pulls through shared memory (<= KK_TILE_SRC_BYTES of source).
#pragma once
#include <stdint.h>

#define KK_TILE_SRC_BYTES 32768u /* payload bytes per smem stage */
#define KK_STAGE_PAD 128u        /* slack so a 16-B-aligned superset of a misaligned tile still fits */
#define KK_Q4K_BLOCK_BYTES 144u
#define KK_Q4K_BLOCK_ELEMS 256u
#define KK_Q4K_TILE_BLOCKS 224u  /* 224*144 = 32256 B in, 224*512 = 114688 B out */
#define KK_Q8_0_BLOCK_BYTES 34u
#define KK_Q8_0_BLOCK_ELEMS 32u
#define KK_Q8_0_TILE_BLOCKS 960u /* 960*34 = 32640 B in (a multiple of 16), 960*64 = 61440 B out */
#define KK_Q6K_BLOCK_BYTES 210u
#define KK_Q6K_TILE_BLOCKS 152u  /* 152*210 = 31920 B in (a multiple of 16), 152*512 = 77824 B out */
#define KK_T_ROWS 32u            /* transpose tile: 32 source rows ... */
#define KK_T_COLS 128u           /* ... x 128 source columns (elements) */
#define KK_T_PITCH_PAD 16u       /* TMA-staged transpose rows sit KK_T_COLS*es + 16 bytes apart (bank spread) */
#define KK_MAX_DST 8

enum KKOp : uint32_t {
  KK_OP_COPY = 0,      // units = bytes
  KK_OP_F32_BF16 = 1,  // units = elements
  KK_OP_F16_BF16 = 2,  // units = elements
  KK_OP_Q4K_BF16 = 3,  // units = 256-weight blocks
  // 2-D transposes: units = source rows in this segment, p0 = source columns (elements),
  // p1 = destination row length in elements (= total source rows of the tensor),
  // p2 = index of this segment's first source row (destination column offset).
  KK_OP_T_F32_BF16 = 4,
  KK_OP_T_F16_BF16 = 5,
  KK_OP_T_B16 = 6,     // 2-byte elements moved verbatim (bf16, i16, ...)
  KK_OP_T_B32 = 7,     // 4-byte elements moved verbatim
  KK_OP_Q8_0_BF16 = 8, // units = 32-weight blocks (34 B: d f16 | 32 x int8)
  KK_OP_Q6K_BF16 = 9,  // units = 256-weight super-blocks (210 B: ql[128] | qh[64] | scales[16] int8 | d f16)
  // SCATTER exchange: units = bytes of whole source rows; every row is cut into N column slices of p1 bytes and slice j
  // goes to pool j (ConvertLaunch::xdst[j]) at dst_off + row * p1.  p0 = row bytes, p1 = slice bytes, p2 = index of the
  // first row of this rank's piece, p3 = bytes of the piece that precede this segment.
  KK_OP_ROWSPLIT = 10,
  KK_OP_COUNT = 11
};

struct KKSeg {
  uint64_t src_off;    // byte offset of the segment's first source byte from the launch's src base
  uint64_t dst_off;    // byte offset into the pool (for transposes: of the destination tensor's origin)
  uint64_t units;      // op-specific, see KKOp
  uint32_t op;
  uint32_t tile_begin; // index of this segment's first tile within the launch
  uint32_t p0, p1, p2, p3;
};

#ifdef __cplusplus
static_assert(sizeof(KKSeg) == 48, "KKSeg layout is shared with the device");

// Units one tile covers, and the number of tiles of a segment (host + device).
static inline
#ifdef __CUDACC__
__host__ __device__
#endif
uint64_t kk_seg_tiles(uint32_t op, uint64_t units, uint32_t p0) {
  switch (op) {
    case KK_OP_COPY:
    case KK_OP_ROWSPLIT: return (units + KK_TILE_SRC_BYTES - 1) / KK_TILE_SRC_BYTES;
    case KK_OP_F32_BF16: return (units + KK_TILE_SRC_BYTES / 4 - 1) / (KK_TILE_SRC_BYTES / 4);
    case KK_OP_F16_BF16: return (units + KK_TILE_SRC_BYTES / 2 - 1) / (KK_TILE_SRC_BYTES / 2);
    case KK_OP_Q4K_BF16: return (units + KK_Q4K_TILE_BLOCKS - 1) / KK_Q4K_TILE_BLOCKS;
    case KK_OP_Q8_0_BF16: return (units + KK_Q8_0_TILE_BLOCKS - 1) / KK_Q8_0_TILE_BLOCKS;
    case KK_OP_Q6K_BF16: return (units + KK_Q6K_TILE_BLOCKS - 1) / KK_Q6K_TILE_BLOCKS;
    case KK_OP_T_F32_BF16:
    case KK_OP_T_B32:
    case KK_OP_T_F16_BF16:
    case KK_OP_T_B16: {
      uint64_t ct = ((uint64_t)p0 + KK_T_COLS - 1) / KK_T_COLS;
      return ((units + KK_T_ROWS - 1) / KK_T_ROWS) * ct;
    }
    default: return 0;
  }
}
#endif
