// distance.cuh -- the two guide distances on the bit-plane layout, usable from host and device.
//
// A guide of L <= 27 bases is held as two 32-bit planes: bit i of `lo` / `hi` is the low / high
// bit of the 2-bit code (A=0 C=1 G=2 T=3) of base i.  Position i of two guides differs iff
// lo-bits differ or hi-bits differ, so the mismatch mask is (qlo^tlo)|(qhi^thi): two LOP3 and one
// POPC per pair for ANY L <= 32 -- this replaces nmslib's bit_hamming over the 4L-bit one-hot
// string (core.py:379-386, :451-455), whose value is exactly twice this count.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GM_HD __host__ __device__ __forceinline__
#else
#define GM_HD static inline
#endif

namespace gm {

GM_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

GM_HD int hamming_planes(uint32_t qlo, uint32_t qhi, uint32_t tlo, uint32_t thi) {
    return popc32((qlo ^ tlo) | (qhi ^ thi));
}

// Unit-cost Levenshtein distance (nmslib space `leven`, core.py:461-465) between two guides of
// the same length L, by Myers' bit-parallel recurrence in Hyyro's formulation for GLOBAL
// alignment: the query is the pattern (its L rows live in one 32-bit word), the target is the
// text.  Pv/Mv are the +1/-1 vertical deltas of the current DP column; row 0 has horizontal delta
// +1, hence the `| 1`.  D[L][L] = L + popc(Pv) - popc(Mv) over the low L bits.  Bits >= L hold
// garbage that never flows downwards (only left shifts and carries), so masking once suffices.
GM_HD int myers_planes(uint32_t qlo, uint32_t qhi, uint32_t tlo, uint32_t thi, int L) {
    uint32_t Pv = 0xFFFFFFFFu, Mv = 0u;
    for (int j = 0; j < L; j++) {
        const uint32_t LO = 0u - ((tlo >> j) & 1u);           // broadcast text base j
        const uint32_t HI = 0u - ((thi >> j) & 1u);
        const uint32_t Eq = ~((qlo ^ LO) | (qhi ^ HI));        // pattern rows equal to text base j
        const uint32_t Xv = Eq | Mv;
        const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        uint32_t Ph = Mv | ~(Xh | Pv);
        uint32_t Mh = Pv & Xh;
        Ph = (Ph << 1) | 1u;
        Mh = Mh << 1;
        Pv = Mh | ~(Xv | Ph);
        Mv = Ph & Xv;
    }
    const uint32_t mask = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
    return L + popc32(Pv & mask) - popc32(Mv & mask);
}

// 2-bit interleaved guide <-> planes (host copies of the device helpers in common.cuh)
GM_HD uint32_t compress_even_bits_hd(uint64_t x) {
    x &= 0x5555555555555555ULL;
    x = (x | (x >> 1)) & 0x3333333333333333ULL;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFULL;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFULL;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFULL;
    return (uint32_t)x;
}

}  // namespace gm
