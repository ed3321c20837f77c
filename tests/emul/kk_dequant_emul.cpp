// Host emulation of kk_convert_kernel's block dequantisers.  TEST INFRASTRUCTURE ONLY.
//
// Compiles kukeon_b200/csrc/kk_dequant.cuh — the exact device source — with g++ by binding the handful of primitives it
// uses (shared-memory loads, single-rounding fp32 arithmetic, bf16 packing, the 16-byte store) to plain C++, then plays
// all 16 consumer warps x 32 lanes of one tile in a loop.  What this checks, on the CPU test tier, is the part of a
// dequantiser that is easy to get wrong and cannot be seen by reading a formula: which lane reads which bytes of which
// block and where its 16 output bytes land.  The bindings are stricter than the hardware: a misaligned 16/32-bit
// shared load, a read outside the staged tile, a store outside the tile's output window or two stores to the same 16 bytes
// all fail the run.  It does not check PTX spelling, memory ordering or anything about the producer warp; the -m gpu
// parity tests do.  The product never links this file and has no CPU path.
#include <cmath>
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../kukeon_b200/csrc/kk_iq_grids.h"
#include "../../kukeon_b200/csrc/kk_tile.h"

namespace {

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

struct Emu {
  const uint8_t* tile = nullptr;  // the stage buffer; shared-memory address a <-> tile[a]
  uint8_t* wtile = nullptr;       // same buffer when the code under test may store into it (sts16 / sts32)
  std::vector<uint32_t>* trace = nullptr;  // when set: every 16/32-bit shared load appends its address (bank-conflict accounting)
  std::vector<uint64_t>* strace = nullptr; // when set: every 16-byte store appends its destination offset (store-transaction accounting)
  uint32_t tile_bytes = 0;
  uint8_t* out = nullptr;         // destination pool 0 (the one hits / masks describe)
  uint8_t* more_out[KK_MAX_DST - 1] = {};  // further destination pools of a fan-out launch: every store is replicated into them
  int n_more = 0;
  uint64_t out_bytes = 0;
  uint8_t* hits = nullptr;  // one counter per 16 output bytes
  uint8_t* out_mask = nullptr;  // per 2 output bytes: written by a scalar (tail) store (per byte for store1_all: see byte_mask)
  uint8_t* byte_mask = nullptr; // per output byte: written by store1_all
  // warp-shuffle emulation (see __shfl_sync below)
  static constexpr int kMaxShfl = 64;
  int shfl_mode = 0, shfl_calls = 0, lane = 0;
  float shfl_table[kMaxShfl][32];
  int err = 0;              // first failure: 1 load out of range, 2 misaligned load, 3 store out of range/misaligned, 4 double store
};
thread_local Emu g;

inline void flag(int e) { if (!g.err) g.err = e; }
inline uint32_t lds8(uint32_t a) {
  if (a >= g.tile_bytes) { flag(1); return 0; }
  return g.tile[a];
}
inline uint32_t lds16(uint32_t a) {
  if (a & 1u) { flag(2); return 0; }
  if (a + 2 > g.tile_bytes) { flag(1); return 0; }
  if (g.trace) g.trace->push_back(a);
  uint16_t v; memcpy(&v, g.tile + a, 2); return v;
}
inline uint32_t lds32(uint32_t a) {
  if (a & 3u) { flag(2); return 0; }
  if (a + 4 > g.tile_bytes) { flag(1); return 0; }
  if (g.trace) g.trace->push_back(a);
  uint32_t v; memcpy(&v, g.tile + a, 4); return v;
}
inline void sts16(uint32_t a, uint32_t v) {
  if ((a & 1u) || a + 2 > g.tile_bytes || !g.wtile) { flag(6); return; }
  uint16_t h = (uint16_t)v; memcpy(g.wtile + a, &h, 2);
}
inline void sts32(uint32_t a, uint32_t v) {
  if ((a & 3u) || a + 4 > g.tile_bytes || !g.wtile) { flag(6); return; }
  memcpy(g.wtile + a, &v, 4);
}
inline uint2 lds64(uint32_t a) {
  uint2 v{0, 0};
  if (a & 7u) { flag(2); return v; }
  if (a + 8 > g.tile_bytes) { flag(1); return v; }
  memcpy(&v, g.tile + a, 8);
  return v;
}
// Warp shuffle by record / replay: the driver runs every lane of a warp twice.  Pass 1 (g.shfl_mode == 1) records the value each lane
// offers at its k-th shuffle and suppresses stores; pass 2 (== 2) hands out the recorded value of the source lane.  Valid for code whose
// shuffled values do not depend on earlier shuffle results and whose lanes all execute the same shuffles (true of q4k_quad).
inline float __shfl_sync(unsigned, float v, int src) {
  if (g.shfl_mode == 1) {
    if (g.shfl_calls >= Emu::kMaxShfl) { flag(7); return v; }
    g.shfl_table[g.shfl_calls++][g.lane] = v;
    return v;
  }
  if (g.shfl_mode == 2) {
    if (g.shfl_calls >= Emu::kMaxShfl) { flag(7); return v; }
    return g.shfl_table[g.shfl_calls++][src & 31];
  }
  flag(7);  // a shuffle outside record/replay cannot be emulated
  return v;
}
// Warp vote, by the same record / replay: pass 1 records every lane's predicate (and returns it: both arms of a vote-guarded branch must
// issue the same shuffles, which they do — the vote only selects the arithmetic), pass 2 returns the AND over the 32 lanes.
inline bool kk_all(bool p) {
  if (g.shfl_mode == 1) {
    if (g.shfl_calls >= Emu::kMaxShfl) { flag(7); return p; }
    g.shfl_table[g.shfl_calls++][g.lane] = p ? 1.0f : 0.0f;
    return p;
  }
  if (g.shfl_mode == 2) {
    if (g.shfl_calls >= Emu::kMaxShfl) { flag(7); return p; }
    bool all = true;
    for (int l = 0; l < 32; ++l) all = all && g.shfl_table[g.shfl_calls][l] != 0.0f;
    g.shfl_calls++;
    return all;
  }
  return p;
}
inline uint32_t kk_f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline uint32_t kk_popc(uint32_t u) { return (uint32_t)__builtin_popcount(u); }
inline uint32_t kk_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  sh &= 31u;
  return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
}
// Aligned word that starts inside the payload and may end up to 3 bytes behind it: on the device those bytes are whatever the stage's
// slack holds, so here they are poison — a result that depends on them fails the comparison with the oracle.
inline uint32_t lds32_slack(uint32_t a) {
  if (a & 3u) { flag(2); return 0; }
  if (a >= g.tile_bytes) { flag(1); return 0; }
  if (g.trace) g.trace->push_back(a);
  uint32_t v = 0;
  for (uint32_t k = 0; k < 4; ++k) v |= (uint32_t)(a + k < g.tile_bytes ? g.tile[a + k] : (uint8_t)(0xA5u ^ (a + k))) << (8 * k);
  return v;
}
const uint64_t kGridIq2xxs[KK_GRID_IQ2XXS_SIZE] = {KK_GRID_IQ2XXS_VALUES};
const uint64_t kGridIq2xs[KK_GRID_IQ2XS_SIZE] = {KK_GRID_IQ2XS_VALUES};
const uint64_t kGridIq2s[KK_GRID_IQ2S_SIZE] = {KK_GRID_IQ2S_VALUES};
const uint32_t kGridIq3xxs[KK_GRID_IQ3XXS_SIZE] = {KK_GRID_IQ3XXS_VALUES};
const uint32_t kGridIq3s[KK_GRID_IQ3S_SIZE] = {KK_GRID_IQ3S_VALUES};
const uint64_t kGridIq1s[KK_GRID_IQ1S_SIZE] = {KK_GRID_IQ1S_VALUES};
// an index outside the table is a bug in the code under test, not something to read through
inline uint64_t kk_grid_iq2xxs(uint32_t i) { if (i >= KK_GRID_IQ2XXS_SIZE) { flag(8); return 0; } return kGridIq2xxs[i]; }
inline uint64_t kk_grid_iq2xs(uint32_t i) { if (i >= KK_GRID_IQ2XS_SIZE) { flag(8); return 0; } return kGridIq2xs[i]; }
inline uint64_t kk_grid_iq2s(uint32_t i) { if (i >= KK_GRID_IQ2S_SIZE) { flag(8); return 0; } return kGridIq2s[i]; }
inline uint32_t kk_grid_iq3xxs(uint32_t i) { if (i >= KK_GRID_IQ3XXS_SIZE) { flag(8); return 0; } return kGridIq3xxs[i]; }
inline uint32_t kk_grid_iq3s(uint32_t i) { if (i >= KK_GRID_IQ3S_SIZE) { flag(8); return 0; } return kGridIq3s[i]; }
inline uint64_t kk_grid_iq1s(uint32_t i) { if (i >= KK_GRID_IQ1S_SIZE) { flag(8); return 0; } return kGridIq1s[i]; }
inline uint32_t kk_ldg8(const uint8_t* p) { return *p; }
// PRMT (default mode) as the code under test uses it: selector nibble k (low 16 bits; the hardware ignores the rest) picks byte
// (nibble & 7) of {b:a}.  Bit 3 of a nibble asks the hardware to replicate the byte's sign bit instead — never what this code wants, so a
// selector that has it set is a failure (9), not something to mask away.
inline uint32_t kk_byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
  if (sel & 0x8888u) flag(9);
  const uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) r |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7u))) & 0xFFu) << (8 * k);
  return r;
}
inline float kk_bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline float kk_fma(float a, float b, float c) { return std::fmaf(a, b, c); }  // correctly rounded single rounding (glibc, -ffp-contract=off)
inline float kk_h2f(uint32_t h) {
  uint16_t b = (uint16_t)h;
  _Float16 x; memcpy(&x, &b, 2);
  return (float)x;
}
inline void kk_h2x2f(uint32_t w, float& x, float& y) { x = kk_h2f(w & 0xFFFFu); y = kk_h2f(w >> 16); }
// one IEEE operation each; -ffp-contract=off in the build line, volatile as a second guard
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline uint32_t bf16_rne(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0x7FFFu;  // cvt.rn.bf16x2.f32: canonical NaN
  return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}
inline uint32_t pack_bf16x2(float a, float b) { return bf16_rne(a) | (bf16_rne(b) << 16); }

struct Dsts { int unused; };
// hits[u] counts the BYTES stored into 16-byte unit u: 16 after exactly one vector store (or eight 2-byte tail stores)
inline void store16_all(const Dsts&, uint64_t off, const uint4& v) {
  if (g.shfl_mode == 1) return;
  if ((off & 15u) || off + 16 > g.out_bytes) { flag(3); return; }
  if (g.hits[off >> 4]) flag(4);
  g.hits[off >> 4] += 16;
  if (g.strace) g.strace->push_back(off);
  memcpy(g.out + off, &v, 16);
  for (int i = 0; i < g.n_more; ++i) memcpy(g.more_out[i] + off, &v, 16);
}
inline void store1_all(const Dsts&, uint64_t off, uint8_t v) {
  if (g.shfl_mode == 1) return;
  if (off + 1 > g.out_bytes) { flag(3); return; }
  if (g.byte_mask[off]) flag(4);
  g.byte_mask[off] = 1;
  g.hits[off >> 4] += 1;
  g.out[off] = v;
  for (int i = 0; i < g.n_more; ++i) g.more_out[i][off] = v;
}
inline void store2_all(const Dsts&, uint64_t off, uint16_t v) {
  if (g.shfl_mode == 1) return;
  if ((off & 1u) || off + 2 > g.out_bytes) { flag(3); return; }
  if (g.out_mask[off >> 1]) flag(4);
  g.out_mask[off >> 1] = 1;
  g.hits[off >> 4] += 2;
  memcpy(g.out + off, &v, 2);
  for (int i = 0; i < g.n_more; ++i) memcpy(g.more_out[i] + off, &v, 2);
}
inline void store8_all(const Dsts&, uint64_t off, uint32_t lo, uint32_t hi) {  // half an output vector: four entries of the 2-byte mask
  if (g.shfl_mode == 1) return;
  if ((off & 7u) || off + 8 > g.out_bytes) { flag(3); return; }
  for (int k = 0; k < 4; ++k) {
    if (g.out_mask[(off >> 1) + k]) flag(4);
    g.out_mask[(off >> 1) + k] = 1;
  }
  g.hits[off >> 4] += 8;
  const uint32_t v[2] = {lo, hi};
  memcpy(g.out + off, v, 8);
  for (int i = 0; i < g.n_more; ++i) memcpy(g.more_out[i] + off, v, 8);
}
inline void store4_all(const Dsts&, uint64_t off, uint32_t v) {  // one 32-bit element (verbatim transposes): two entries of the 2-byte mask
  if (g.shfl_mode == 1) return;
  if ((off & 3u) || off + 4 > g.out_bytes) { flag(3); return; }
  if (g.out_mask[off >> 1] || g.out_mask[(off >> 1) + 1]) flag(4);
  g.out_mask[off >> 1] = g.out_mask[(off >> 1) + 1] = 1;
  g.hits[off >> 4] += 4;
  memcpy(g.out + off, &v, 4);
  for (int i = 0; i < g.n_more; ++i) memcpy(g.more_out[i] + off, &v, 4);
}
inline uint4 lds128(uint32_t a) {
  uint4 v{0, 0, 0, 0};
  if (a & 15u) { flag(2); return v; }
  if (a + 16 > g.tile_bytes) { flag(1); return v; }
  memcpy(&v, g.tile + a, 16);
  return v;
}
// FP8 -> fp16 bit patterns (what cvt.rn.f16x2.e4m3x2 / .e5m2x2 produce; exact)
inline uint32_t f8_to_f16_bits(uint32_t b, bool e5m2) {
  b &= 0xFFu;
  if (e5m2) return b << 8;
  const uint32_t s = (b >> 7) << 15, e = (b >> 3) & 15u, m = b & 7u;
  if (e == 15 && m == 7) return s | 0x7FFFu;                  // NaN
  if (e == 0) {                                              // subnormal m * 2^-9 -> normal fp16
    if (!m) return s;
    int sh = m >= 4 ? 0 : m >= 2 ? 1 : 2;                    // leading one at bit 2 - sh
    return s | (uint32_t)((8 - sh) << 10) | (((m << (sh + 1)) & 7u) << 7);
  }
  return s | ((e + 8u) << 10) | (m << 7);                    // bias 7 -> 15
}
template <bool E5M2>
inline uint32_t kk_f8x2_to_f16x2(uint32_t v) { return f8_to_f16_bits(v, E5M2) | (f8_to_f16_bits(v >> 8, E5M2) << 16); }

constexpr int kConsumerWarps = 16;  // must equal kConsumerWarps of kk_kernels.cu
constexpr int kConsumerThreads = kConsumerWarps * 32;
#define KK_DQ_DEV static inline
#define min(a, b) ((a) < (b) ? (a) : (b))
#include "../../kukeon_b200/csrc/kk_consume_core.cuh"
#include "../../kukeon_b200/csrc/kk_dequant.cuh"
#undef min

// (run_warps, below, plays all warps of a tile; ops that shuffle are run twice per warp)
// One lane of the consumer side of a block / elementwise op — the same switch the kernel has.  false: op not covered here.
inline bool run_op(uint32_t op, const Dsts& D, uint32_t pay, uint32_t n, uint64_t dst_off, int cwarp, int lane) {
  switch (op) {
    case KK_OP_Q8_0_BF16: consume_q8_0(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_Q6K_BF16: consume_q6k(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_Q4_0_BF16: consume_legacy32<KK_Q4_0_BLOCK_BYTES, false, false>(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_Q4_1_BF16: consume_legacy32<KK_Q4_1_BLOCK_BYTES, true, false>(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_Q5_0_BF16: consume_legacy32<KK_Q5_0_BLOCK_BYTES, false, true>(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_Q5_1_BF16: consume_legacy32<KK_Q5_1_BLOCK_BYTES, true, true>(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_Q2K_BF16: consume_q2k(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_Q3K_BF16: consume_q3k(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_Q5K_BF16: consume_q5k(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_IQ4NL_BF16: consume_codebook32<KK_IQ4NL_BLOCK_BYTES, 0>(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_MXFP4_BF16: consume_codebook32<KK_MXFP4_BLOCK_BYTES, 1>(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_IQ4XS_BF16: consume_iq4xs(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_IQ2XXS_BF16: consume_iq2xxs(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_IQ2XS_BF16: consume_iq2xs(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_IQ2S_BF16: consume_iq2s(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_IQ3XXS_BF16: consume_iq3xxs(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_IQ3S_BF16: consume_iq3s(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_IQ1S_BF16: consume_iq1s(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_IQ1M_BF16: consume_iq1m(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_TQ1_0_BF16: consume_tq1_0(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_TQ2_0_BF16: consume_tq2_0(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_NVFP4_BF16: consume_nvfp4(D, pay, n, dst_off, cwarp, lane); return true;
    case KK_OP_Q4K_BF16: consume_q4k(D, pay, n, dst_off, cwarp, lane); return true;
    // elementwise ops: n = elements (COPY: bytes) of the tile, threads indexed 0..511 across the consumer warps
    case KK_OP_COPY: consume_copy(D, pay, n, dst_off, cwarp * 32 + lane); return true;
    case KK_OP_F32_BF16: consume_f32(D, pay, n, dst_off, cwarp * 32 + lane); return true;
    case KK_OP_F16_BF16: consume_f16(D, pay, n, dst_off, cwarp * 32 + lane); return true;
    case KK_OP_F8E4M3_BF16: consume_f8<false>(D, pay, n, dst_off, cwarp * 32 + lane); return true;
    case KK_OP_F8E5M2_BF16: consume_f8<true>(D, pay, n, dst_off, cwarp * 32 + lane); return true;
    default: return false;
  }
}

}  // namespace

inline bool run_warps(uint32_t op, const Dsts& D, uint32_t pay, uint32_t n, uint64_t dst_off) {
  const bool shuffles = op == KK_OP_Q4K_BF16 || op == KK_OP_Q5K_BF16;
  for (int cwarp = 0; cwarp < kConsumerWarps; ++cwarp) {
    for (int pass = shuffles ? 1 : 0; pass <= (shuffles ? 2 : 0); ++pass) {
      g.shfl_mode = pass;
      for (int lane = 0; lane < 32; ++lane) {
        g.lane = lane;
        g.shfl_calls = 0;
        if (!run_op(op, D, pay, n, dst_off, cwarp, lane)) return false;
      }
    }
  }
  g.shfl_mode = 0;
  return true;
}

// Shared-memory wavefronts of per-lane load traces: the k-th recorded load of every lane of a warp is one warp instruction; it needs as
// many wavefronts as the worst bank has distinct 4-byte words (same word = broadcast).  Adds to wavefronts / ideal.
void count_wavefronts(const std::vector<std::vector<uint32_t>>& traces, uint64_t& wavefronts, uint64_t& ideal) {
  for (size_t w = 0; w + 31 < traces.size(); w += 32) {
    size_t longest = 0;
    for (int l = 0; l < 32; ++l) longest = std::max(longest, traces[w + l].size());
    for (size_t k = 0; k < longest; ++k) {
      std::vector<uint32_t> words[32];
      bool any = false;
      for (int l = 0; l < 32; ++l) {
        const auto& tr = traces[w + l];
        if (k >= tr.size()) continue;
        any = true;
        const uint32_t word = tr[k] >> 2;
        auto& v = words[word & 31u];
        if (std::find(v.begin(), v.end(), word) == v.end()) v.push_back(word);
      }
      if (!any) continue;
      size_t worst = 1;
      for (auto& v : words) worst = std::max(worst, v.size());
      wavefronts += worst;
      ideal += 1;
    }
  }
}

// Run the consumer side of one tile of block op `op` (KKOp): `nblk` blocks whose first byte sits at tile[pay_off].
// out receives nblk * out_bytes_per_block bytes; hits (out_bytes/16 counters, zeroed here) how often each 16-byte unit
// was stored.  Returns 0, a positive Emu::err code, or -1 for an op this harness does not cover.
extern "C" int kk_emul_dequant_tile(uint32_t op, const uint8_t* tile, uint32_t tile_bytes, uint32_t pay_off, uint32_t nblk, uint8_t* out,
                                    uint64_t out_bytes, uint8_t* hits) {
  g = Emu{};
  g.tile = tile; g.tile_bytes = tile_bytes; g.out = out; g.out_bytes = out_bytes; g.hits = hits;
  memset(hits, 0, (out_bytes + 15) / 16);
  std::vector<uint8_t> mask(out_bytes / 2 + 1, 0), bmask(out_bytes + 1, 0);
  g.out_mask = mask.data();
  g.byte_mask = bmask.data();
  const Dsts D{0};
  if (!run_warps(op, D, pay_off, nblk, 0)) return -1;
  return g.err;
}

// Same as kk_emul_dequant_tile, additionally reporting the shared-memory wavefronts of the 16/32-bit loads (stats[0]) against one per warp
// load instruction (stats[1]).  Only meaningful when all lanes of a warp execute the same loads (full warps of a full tile); byte loads
// (lds8) and vector loads (lds64 / lds128) are not traced.  Ops that shuffle are traced in their replay pass.
extern "C" int kk_emul_dequant_tile_stats(uint32_t op, const uint8_t* tile, uint32_t tile_bytes, uint32_t pay_off, uint32_t nblk, uint8_t* out,
                                          uint64_t out_bytes, uint8_t* hits, uint64_t* stats) {
  g = Emu{};
  g.tile = tile; g.tile_bytes = tile_bytes; g.out = out; g.out_bytes = out_bytes; g.hits = hits;
  memset(hits, 0, (out_bytes + 15) / 16);
  std::vector<uint8_t> mask(out_bytes / 2 + 1, 0), bmask(out_bytes + 1, 0);
  g.out_mask = mask.data();
  g.byte_mask = bmask.data();
  const Dsts D{0};
  const bool shuffles = op == KK_OP_Q4K_BF16 || op == KK_OP_Q5K_BF16;
  std::vector<std::vector<uint32_t>> traces((size_t)kConsumerWarps * 32);
  for (int cwarp = 0; cwarp < kConsumerWarps; ++cwarp)
    for (int pass = shuffles ? 1 : 0; pass <= (shuffles ? 2 : 0); ++pass) {
      g.shfl_mode = pass;
      for (int lane = 0; lane < 32; ++lane) {
        g.lane = lane;
        g.shfl_calls = 0;
        g.trace = pass == 1 ? nullptr : &traces[(size_t)(cwarp * 32 + lane)];
        if (!run_op(op, D, pay_off, nblk, 0, cwarp, lane)) return -1;
      }
    }
  g.trace = nullptr;
  g.shfl_mode = 0;
  uint64_t wf = 0, ideal = 0;
  count_wavefronts(traces, wf, ideal);
  if (stats) { stats[0] = wf; stats[1] = ideal; }
  return g.err;
}

// A whole segment the way the kernel walks it: the producer's per-tile arithmetic (kk_block_tile, the same function the
// kernel calls), the 16-byte aligned superset its bulk copy brings into the stage (pay_off = source address & 15), then the
// consumers.  `src` holds the segment's blocks; src_misalign (0..15) is the alignment class of its first byte in the staged
// chunk.  out/hits cover units * out_bytes_per_block bytes.  Returns like kk_emul_dequant_tile.
extern "C" int kk_emul_dequant_segment(uint32_t op, const uint8_t* src, uint64_t units, uint32_t src_misalign, uint8_t* out, uint64_t out_bytes,
                                       uint8_t* hits) {
  const KKBlockGeom gm = kk_block_geom(op);
  if (!gm.block_bytes) return -1;
  KKSeg seg{};
  seg.op = op;
  seg.units = units;
  seg.src_off = src_misalign;  // pretend the launch's src base is 16-byte aligned and the segment starts here
  memset(hits, 0, out_bytes / 16);
  static thread_local uint8_t stage[KK_TILE_SRC_BYTES + KK_STAGE_PAD];
  std::vector<uint8_t> mask(out_bytes / 2 + 1, 0), bmask(out_bytes + 1, 0);
  const uint64_t n_tiles = kk_seg_tiles(op, units, 0);
  for (uint64_t t = 0; t < n_tiles; ++t) {
    const KKBlockTile bt = kk_block_tile(seg, (uint32_t)t);
    const uint32_t mis = (uint32_t)(bt.in_off & 15u);
    const uint32_t tx = (mis + bt.in_bytes + 15u) & ~15u;  // what the producer asks TMA for
    if (tx > sizeof stage) return 5;
    memset(stage, 0xEE, sizeof stage);
    // bytes before the payload / after it come from the neighbouring source bytes on the GPU; the consumers must not depend on them
    memcpy(stage + mis, src + (bt.in_off - src_misalign), bt.in_bytes);
    g = Emu{};
    g.tile = stage; g.tile_bytes = mis + bt.in_bytes; g.out = out; g.out_bytes = out_bytes; g.hits = hits;
    g.out_mask = mask.data();
    g.byte_mask = bmask.data();
    const Dsts D{0};
    if (!run_warps(op, D, mis, bt.n_blocks, bt.dst_off)) return -1;
    if (g.err) return g.err;
  }
  return 0;
}

// One transpose tile (KK_OP_T_*: 8 rows, compact rows) the way the
// kernel runs it.  `src` points at source element (r0, c0) of a row-major [*, C] tensor of ES-byte elements; staged != 0 lays the tile out
// as the producer's bulk copies do, staged == 0 runs the consumers' gather fallback first.  `dst` is the destination tensor's origin
// ([C_total, R] row-major, 2-byte elements — 4-byte for KK_OP_T_B32), dst_bytes its size; hits as in kk_emul_dequant_tile over dst.
// stats: [0] shared-memory wavefronts the consumers' loads need (one per distinct 4-byte word per bank per warp instruction), [1] the
// ideal (one per warp load), [2] warp-level 16-byte store instructions, [3] distinct 128-byte lines and [4] distinct 32-byte sectors those
// touch, summed over instructions (what the LSU / L2 see as transactions).
extern "C" int kk_emul_t_tile(uint32_t op, const uint8_t* src, uint32_t C, uint32_t nr, uint32_t nc, uint32_t R, uint32_t col0, uint32_t row0,
                              int staged, uint8_t* dst, uint64_t dst_bytes, uint8_t* hits, uint64_t* stats) {
  if (!kk_is_transpose(op)) return -1;
  const uint32_t es = kk_t_src_es(op);
  if (nr > KK_T_ROWS || nc * es > KK_T_ROW_BYTES) return 5;
  static thread_local uint8_t stage[KK_TILE_SRC_BYTES + KK_STAGE_PAD];
  memset(stage, 0xEE, sizeof stage);
  uint32_t pitch = nc * es;
  std::vector<uint8_t> mask(dst_bytes / 2 + 1, 0), bmask(dst_bytes + 1, 0);
  memset(hits, 0, (dst_bytes + 15) / 16);
  g = Emu{};
  g.tile = stage; g.wtile = stage; g.tile_bytes = sizeof stage; g.out = dst; g.out_bytes = dst_bytes; g.hits = hits; g.out_mask = mask.data();
  g.byte_mask = bmask.data();
  const Dsts D{0};
  const int nthreads = kConsumerWarps * 32;
  if (staged) {
    for (uint32_t r = 0; r < nr; ++r) memcpy(stage + r * pitch, src + (uint64_t)r * C * es, (size_t)nc * es);
  } else {
    pitch = (pitch + 3u) & ~3u;
    for (int t = 0; t < nthreads; ++t) {
      if (es == 4) t_gather<4>(src, 0, pitch, nr, nc, C, t);
      else t_gather<2>(src, 0, pitch, nr, nc, C, t);
    }
    // (the kernel has a named barrier here)
  }
  std::vector<std::vector<uint32_t>> traces((size_t)nthreads);
  std::vector<std::vector<uint64_t>> straces((size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    g.trace = &traces[(size_t)t];
    g.strace = &straces[(size_t)t];
    switch (op) {
      case KK_OP_T_F32_BF16: consume_t<4, 1>(D, 0, pitch, nr, nc, R, col0, row0, 0, t); break;
      case KK_OP_T_F16_BF16: consume_t<2, 2>(D, 0, pitch, nr, nc, R, col0, row0, 0, t); break;
      case KK_OP_T_B16: consume_t<2, 0>(D, 0, pitch, nr, nc, R, col0, row0, 0, t); break;
      default: consume_t<4, 3>(D, 0, pitch, nr, nc, R, col0, row0, 0, t); break;
    }
  }
  g.trace = nullptr;
  g.strace = nullptr;
  uint64_t wavefronts = 0, ideal = 0, st_instr = 0, st_lines = 0, st_sectors = 0;
  for (int w = 0; w < kConsumerWarps; ++w) {
    size_t longest = 0;
    for (int l = 0; l < 32; ++l) longest = std::max(longest, traces[(size_t)(32 * w + l)].size());
    for (size_t k = 0; k < longest; ++k) {  // the k-th shared load of every lane of the warp is one warp instruction
      std::vector<uint32_t> words[32];
      bool any = false;
      for (int l = 0; l < 32; ++l) {
        const auto& tr = traces[(size_t)(32 * w + l)];
        if (k >= tr.size()) continue;
        any = true;
        const uint32_t word = tr[k] >> 2;
        auto& v = words[word & 31u];
        if (std::find(v.begin(), v.end(), word) == v.end()) v.push_back(word);
      }
      if (!any) continue;
      size_t worst = 1;
      for (auto& v : words) worst = std::max(worst, v.size());
      wavefronts += worst;
      ideal += 1;
    }
    longest = 0;
    for (int l = 0; l < 32; ++l) longest = std::max(longest, straces[(size_t)(32 * w + l)].size());
    for (size_t k = 0; k < longest; ++k) {  // likewise the k-th 16-byte store
      std::vector<uint64_t> lines, sectors;
      for (int l = 0; l < 32; ++l) {
        const auto& tr = straces[(size_t)(32 * w + l)];
        if (k >= tr.size()) continue;
        const uint64_t ln = tr[k] >> 7, sc = tr[k] >> 5;
        if (std::find(lines.begin(), lines.end(), ln) == lines.end()) lines.push_back(ln);
        if (std::find(sectors.begin(), sectors.end(), sc) == sectors.end()) sectors.push_back(sc);
      }
      if (lines.empty()) continue;
      st_instr += 1;
      st_lines += lines.size();
      st_sectors += sectors.size();
    }
  }
  if (stats) { stats[0] = wavefronts; stats[1] = ideal; stats[2] = st_instr; stats[3] = st_lines; stats[4] = st_sectors; }
  return g.err;
}

// A whole launch of kk_convert_kernel, tile by tile: resolve the tile's segment from tile_begin[] (as the producer does), kk_make_tile (the
// function the kernel's producer calls when built with KK_PRODUCER_SHARED), perform the bulk copies it asks for into a stage buffer with the
// hardware's constraints checked (16-byte aligned source, destination and size; bytes issued == bytes announced), then the consumer side: the
// aligned-copy bulk stores, or the same per-lane device functions the other entry points run, for all 16 x 32 consumer lanes.
// src: the staged chunk (src base assumed 256-byte aligned like the staging buffers); dst[0..n_dst): destination pools of pool_bytes each;
// hits / masks describe dst[0] and ACCUMULATE across calls (the caller zeroes them once per pool).  Not covered: the scatter row exchange (return -2).
extern "C" int kk_emul_launch(const uint8_t* src, uint64_t src_bytes, const KKSeg* segs, uint32_t n_segs, uint32_t n_tiles, uint32_t flags,
                              uint8_t* const* dst, uint32_t n_dst, uint64_t pool_bytes, uint8_t* hits, uint8_t* mask2, uint8_t* mask1) {
  if (n_dst < 1 || n_dst > KK_MAX_DST) return -1;
  static thread_local uint8_t stage[KK_TILE_SRC_BYTES + KK_STAGE_PAD];
  const uint64_t src_addr = 0x100000;  // only the alignment of the src base matters
  uint32_t cur = 0;
  for (uint32_t tile = 0; tile < n_tiles; ++tile) {
    while (cur + 1 < n_segs && segs[cur + 1].tile_begin <= tile) ++cur;
    const KKSeg& seg = segs[cur];
    if (tile < seg.tile_begin) return 9;  // tile_begin[] not monotone from 0
    KKTileDesc d;
    KKTileLoad ld;
    kk_make_tile(seg, tile - seg.tile_begin, src_addr, flags, d, ld);
    memset(stage, 0xEE, sizeof stage);
    uint32_t staged = 0;
    if (ld.kind == 1) {
      if (((src_addr + ld.g_off) & 15u) || (ld.tx & 15u) || ld.tx > sizeof stage) return 10;
      if (ld.g_off > src_bytes) return 11;
      // the last tile of a chunk may over-read up to 15 bytes of slack behind the staged bytes (the planner leaves room for it)
      const uint64_t avail = src_bytes - ld.g_off;
      if (ld.tx > avail + 15) return 11;
      memcpy(stage, src + ld.g_off, (size_t)std::min<uint64_t>(ld.tx, avail));
      staged = ld.tx;
    } else if (ld.kind == 2) {
      uint32_t sum = 0;
      for (uint32_t r = 0; r < ld.nrows; ++r) {
        const uint64_t go = ld.g_off + (uint64_t)r * ld.gpitch;
        const uint32_t so = r * ld.spitch;
        if (((src_addr + go) & 15u) || (so & 15u) || (ld.row_bytes & 15u) || so + ld.row_bytes > sizeof stage) return 10;
        if (go + ld.row_bytes > src_bytes) return 11;
        memcpy(stage + so, src + go, ld.row_bytes);
        sum += ld.row_bytes;
        staged = std::max(staged, so + ld.row_bytes);
      }
      if (sum != ld.tx) return 12;  // the mbarrier would never complete (or complete early)
    }
    g = Emu{};
    g.tile = stage; g.wtile = stage; g.tile_bytes = ld.kind ? staged : (uint32_t)sizeof stage;
    g.out = dst[0]; g.out_bytes = pool_bytes; g.hits = hits; g.out_mask = mask2; g.byte_mask = mask1;
    g.n_more = (int)n_dst - 1;
    for (int i = 0; i < g.n_more; ++i) g.more_out[i] = dst[i + 1];
    const Dsts D{0};
    const uint32_t nr = d.n_units & 0xFFFFu, nc = d.n_units >> 16;
    if (d.bulk == 1) {  // aligned verbatim copy: one bulk store from the stage to every pool
      if ((d.dst_off & 15u) || (d.n_units & 15u) || (d.pay_off & 15u) || d.dst_off + d.n_units > pool_bytes) return 13;
      for (uint32_t u = 0; u < d.n_units; u += 16) {
        if (hits[(d.dst_off + u) >> 4]) return 4;
        hits[(d.dst_off + u) >> 4] += 16;
      }
      for (uint32_t i = 0; i < n_dst; ++i) memcpy(dst[i] + d.dst_off, stage + d.pay_off, d.n_units);
      continue;
    }
    if (d.bulk == 3 || d.op == KK_OP_ROWSPLIT) return -2;
    switch (d.op) {
      case KK_OP_T_F32_BF16: case KK_OP_T_F16_BF16: case KK_OP_T_B16: case KK_OP_T_B32: {
        const uint32_t es = kk_t_src_es(d.op);
        uint32_t pitch = nc * es;
        if (d.bulk != 4u) {  // run_t of the kernel: gather, (barrier), consume
          pitch = (pitch + 3u) & ~3u;
          g.tile_bytes = (uint32_t)sizeof stage;
          if (d.src_off + ((uint64_t)(nr ? nr - 1 : 0) * d.C + nc) * es > src_bytes) return 11;
          for (int t = 0; t < kConsumerThreads; ++t) {
            if (es == 4) t_gather<4>(src + d.src_off, 0, pitch, nr, nc, d.C, t);
            else t_gather<2>(src + d.src_off, 0, pitch, nr, nc, d.C, t);
          }
        }
        for (int t = 0; t < kConsumerThreads; ++t) {
          switch (d.op) {
            case KK_OP_T_F32_BF16: consume_t<4, 1>(D, 0, pitch, nr, nc, d.R, d.col0, d.row0, d.dst_off, t); break;
            case KK_OP_T_F16_BF16: consume_t<2, 2>(D, 0, pitch, nr, nc, d.R, d.col0, d.row0, d.dst_off, t); break;
            case KK_OP_T_B16: consume_t<2, 0>(D, 0, pitch, nr, nc, d.R, d.col0, d.row0, d.dst_off, t); break;
            default: consume_t<4, 3>(D, 0, pitch, nr, nc, d.R, d.col0, d.row0, d.dst_off, t); break;
          }
        }
        break;
      }
      default:
        if (!run_warps(d.op, D, d.pay_off, d.n_units, d.dst_off)) return -1;
        break;
    }
    if (g.err) return g.err;
  }
  return 0;
}

// Geometry the kernel and the planner share (kk_ops.h), exported so the test can cross-check the Python-side tables.
extern "C" void kk_emul_block_geom(uint32_t op, uint32_t* block_bytes, uint32_t* out_bytes, uint32_t* tile_blocks) {
  const KKBlockGeom gm = kk_block_geom(op);
  *block_bytes = gm.block_bytes; *out_bytes = gm.out_bytes; *tile_blocks = gm.tile_blocks;
}
