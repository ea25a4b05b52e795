// bsg_gram.cuh -- integer Gram tiles between lines of a packed 2-bit matrix on the tensor pipe.
//
//   S[i][j] = sum_k  fA(code(i, k)) * fB(code(j, k))        (exact, int32)
//
// for a 128 x 64 tile of line pairs, where fA / fB select a "plane" of the staged code:
//   PL_A : the genotype with missing -> 0   (0, 1, 2)
//   PL_B : valid indicator                   (1 unless missing)
//   PL_H : [genotype == 2]                   (so that sum a^2 = sum a + 2 sum h)
// Both operands of mma.sync.m16n8k32.u8.u8 come straight from the packed words of their lines: a register of the
// A fragment and a register of the B fragment are both "4 consecutive-k bytes of one line", i.e. a class mask
// ((plane >> 2c) & 0x03030303) of the same word index -- no shared-memory staging, no per-element unpack.
// Used by the windowed correlations (bsg_cor.cu) and, with a per-k weight digit folded into the B bytes, by the
// Gram product of bed_tcrossprodSelf (bsg_la.cu).
#pragma once
#include <stdint.h>

namespace bsg {
namespace gram {

constexpr int TM = 128, TN = 64;    // tile of line pairs per CTA
constexpr int WARPS = 8;            // 4 (M) x 2 (N) warps, 32 x 32 outputs each
constexpr int THREADS = WARPS * 32;
constexpr int CHUNK = 64;           // bytes per line per step (one LDG.128 per lane = 4 words = 64 codes)

enum Plane { PL_A = 0, PL_B = 1, PL_H = 2 };

struct Tile {
  int i0, j0;        // first A line, first B line
  int mode;          // 0: one product (a,a), codes without missing values; 1: the six products of the NA-aware cor
  long long out;     // offset (in int32) of this tile's sums: [nprod][TM][TN]
};

__device__ __forceinline__ void mma_u8u8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg128(const uint8_t *p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// plane word of 16 codes
template <int PL>
__device__ __forceinline__ uint32_t plane_word(uint32_t w) {
  const uint32_t n = w & (w >> 1) & 0x55555555u;  // bit 2p set iff code p is missing
  if (PL == PL_A) return w & ~(n | (n << 1));
  if (PL == PL_B) return ~n & 0x55555555u;
  return ((w & ~(n | (n << 1))) >> 1) & 0x55555555u;  // PL_H
}

}  // namespace gram
}  // namespace bsg
