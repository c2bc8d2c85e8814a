// Shared device helpers: 256-bit Hamming distance on the integer pipe, packed top-2 keys.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace plm {

constexpr uint32_t KEY32_ABSENT = 0xFFFFFFFFu;
constexpr unsigned long long KEY64_ABSENT = 0xFFFFFFFFFFFFFFFFull;

// A 256-bit descriptor held in registers as two 128-bit halves.
struct Desc {
    uint4 lo, hi;
};

__device__ __forceinline__ Desc load_desc(const uint4 *__restrict__ base, long long row) {
    Desc d;
    d.lo = __ldg(base + 2 * row);
    d.hi = __ldg(base + 2 * row + 1);
    return d;
}

// Same through a generic pointer (the row may live in shared memory).
__device__ __forceinline__ Desc load_desc_any(const uint4 *base, long long row) {
    Desc d;
    d.lo = base[2 * row];
    d.hi = base[2 * row + 1];
    return d;
}

// StVO::distance (stvo-pl/src/matching.cpp:93-109): 8 x (xor, popcount).  LOP3 + POPC + IADD3.
__device__ __forceinline__ int hamming256(const Desc &a, const uint4 &blo, const uint4 &bhi) {
    int s0 = __popc(a.lo.x ^ blo.x) + __popc(a.lo.y ^ blo.y) + __popc(a.lo.z ^ blo.z);
    int s1 = __popc(a.lo.w ^ blo.w) + __popc(a.hi.x ^ bhi.x) + __popc(a.hi.y ^ bhi.y);
    int s2 = __popc(a.hi.z ^ bhi.z) + __popc(a.hi.w ^ bhi.w);
    return s0 + s1 + s2;
}

__device__ __forceinline__ int hamming256(const Desc &a, const Desc &b) {
    return hamming256(a, b.lo, b.hi);
}

// Carry-save variant: two LOP3 full adders (sum = a^b^c, carry = maj(a,b,c)) compress six of the
// eight xor words into 2 "ones" + 2 "twos" words, then a third compresses the ones again:
// 5 POPC instead of 8 at the price of 6 LOP3 -- POPC issues at a quarter of the LOP3 rate, so this
// moves work from the saturated pipe to the idle one.  Exactly the same integer result.
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ int hamming256_csa(const Desc &a, const uint4 &blo, const uint4 &bhi) {
    const uint32_t x0 = a.lo.x ^ blo.x, x1 = a.lo.y ^ blo.y, x2 = a.lo.z ^ blo.z, x3 = a.lo.w ^ blo.w;
    const uint32_t x4 = a.hi.x ^ bhi.x, x5 = a.hi.y ^ bhi.y, x6 = a.hi.z ^ bhi.z, x7 = a.hi.w ^ bhi.w;
    const uint32_t sa = lop3_xor3(x0, x1, x2), ca = lop3_maj(x0, x1, x2);
    const uint32_t sb = lop3_xor3(x3, x4, x5), cb = lop3_maj(x3, x4, x5);
    const uint32_t sc = lop3_xor3(sa, sb, x6), cc = lop3_maj(sa, sb, x6);
    const int ones = __popc(sc) + __popc(x7);
    const int twos = __popc(ca) + __popc(cb) + __popc(cc);
    return ones + 2 * twos;
}

// Four-POPC carry-save form: a fourth full adder compresses the three "twos" words again, so the
// distance is popc(ones0) + popc(ones1) + 2*popc(twos) + 4*popc(fours): 4 POPC + 16 LOP3 per pair.
__device__ __forceinline__ int hamming256_csa4(const Desc &a, const uint4 &blo, const uint4 &bhi) {
    const uint32_t x0 = a.lo.x ^ blo.x, x1 = a.lo.y ^ blo.y, x2 = a.lo.z ^ blo.z, x3 = a.lo.w ^ blo.w;
    const uint32_t x4 = a.hi.x ^ bhi.x, x5 = a.hi.y ^ bhi.y, x6 = a.hi.z ^ bhi.z, x7 = a.hi.w ^ bhi.w;
    const uint32_t sa = lop3_xor3(x0, x1, x2), ca = lop3_maj(x0, x1, x2);
    const uint32_t sb = lop3_xor3(x3, x4, x5), cb = lop3_maj(x3, x4, x5);
    const uint32_t sc = lop3_xor3(sa, sb, x6), cc = lop3_maj(sa, sb, x6);
    const uint32_t se = lop3_xor3(ca, cb, cc), ce = lop3_maj(ca, cb, cc);
    // weights folded with integer multiply-adds: IMAD issues on the FMA pipe, which is idle, while
    // the ALU pipe (LOP3 / IADD3 / VIMNMX) is the one that binds
    return static_cast<int>(imad(__popc(ce), 4u, imad(__popc(se), 2u, imad(__popc(sc), 1u, __popc(x7)))));
}

// Thirteen-LOP3 form.  Both descriptors are first mapped through the same invertible GF(2)-linear transform
//     T(w) = (w0, w1, w0^w1^w2, w3, w4, w3^w4^w5, w0^w1^w2^w3^w4^w5^w6, w7)
// (the query once per thread, a train row once per shared-memory stage, i.e. amortised over >= 64 pairs).  With
// x_k = a_k ^ b_k, the eight xors y = T(a) ^ T(b) then ARE x0, x1, sa = x0^x1^x2, x3, x4, sb = x3^x4^x5,
// sc = sa^sb^x6 and x7: the three "sum" outputs of the first carry-save level cost nothing, and each carry is one
// LOP3 of (two addends, their sum): maj(p, q, r) with r = p^q^s is the 3-input function 0xD4 of (p, q, s).
// 8 xor + 3 carries + 1 full adder on the carries = 13 LOP3 + 4 POPC per pair, same integer result as
// StVO::distance (stvo-pl/src/matching.cpp:93-109).
__device__ __forceinline__ uint32_t lop3_carry_from_sum(uint32_t p, uint32_t q, uint32_t s) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(r) : "r"(p), "r"(q), "r"(s));
    return r;
}
__device__ __forceinline__ void desc_transform13(uint4 &lo, uint4 &hi) {
    lo.z = lop3_xor3(lo.x, lo.y, lo.z);
    hi.y = lop3_xor3(lo.w, hi.x, hi.y);
    hi.z = lop3_xor3(lo.z, hi.y, hi.z);
}
// a, blo, bhi are TRANSFORMED descriptors.
__device__ __forceinline__ int hamming256_t13(const Desc &a, const uint4 &blo, const uint4 &bhi) {
    const uint32_t x0 = a.lo.x ^ blo.x, x1 = a.lo.y ^ blo.y, sa = a.lo.z ^ blo.z, x3 = a.lo.w ^ blo.w;
    const uint32_t x4 = a.hi.x ^ bhi.x, sb = a.hi.y ^ bhi.y, sc = a.hi.z ^ bhi.z, x7 = a.hi.w ^ bhi.w;
    const uint32_t ca = lop3_carry_from_sum(x0, x1, sa);
    const uint32_t cb = lop3_carry_from_sum(x3, x4, sb);
    const uint32_t cc = lop3_carry_from_sum(sa, sb, sc);
    const uint32_t se = lop3_xor3(ca, cb, cc), ce = lop3_maj(ca, cb, cc);
    return static_cast<int>(imad(__popc(ce), 4u, imad(__popc(se), 2u, imad(__popc(sc), 1u, __popc(x7)))));
}

// Three-POPC LOWER BOUND (the "B" rows of knn2 variant 4).  Transform as above except the last word, which becomes
// the parity of all eight words: the xor y7 is then s1 = x0 ^ ... ^ x7, the ones digit of the per-column count.
// The count of a column (0..8) is s1 + 2 s2 + 4 s4 + 8 c8; c8 (all eight words differ in that column) is dropped:
//     lb = popc(s1) + 2 popc(s2) + 4 popc(s4)  <=  d,   d - lb = 8 popc(c8)   (mean 1 for random descriptors).
// 16 LOP3 + 3 POPC: more ALU work, less XU work than the 13 / 4 form -- a block mixes both to balance the two pipes.
// The exact distance (update path only) adds 8 popc(ce & se & c1).
__device__ __forceinline__ uint32_t lop3_andn(uint32_t a, uint32_t b) { // a & ~b
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %2, 0x30;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t lop3_xor_and(uint32_t a, uint32_t b, uint32_t c) { // a ^ (b & c)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_and3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ void desc_transform16(uint4 &lo, uint4 &hi) {
    desc_transform13(lo, hi);
    hi.w ^= hi.z;
}
// a, blo, bhi are descriptors in the parity-transformed domain.  EXACT selects the exact distance.
template <bool EXACT>
__device__ __forceinline__ int hamming256_t16(const Desc &a, const uint4 &blo, const uint4 &bhi) {
    const uint32_t x0 = a.lo.x ^ blo.x, x1 = a.lo.y ^ blo.y, sa = a.lo.z ^ blo.z, x3 = a.lo.w ^ blo.w;
    const uint32_t x4 = a.hi.x ^ bhi.x, sb = a.hi.y ^ bhi.y, sc = a.hi.z ^ bhi.z, s1 = a.hi.w ^ bhi.w;
    const uint32_t ca = lop3_carry_from_sum(x0, x1, sa);
    const uint32_t cb = lop3_carry_from_sum(x3, x4, sb);
    const uint32_t cc = lop3_carry_from_sum(sa, sb, sc);
    const uint32_t c1 = lop3_andn(sc, s1);                   // sc & x7 with x7 = sc ^ s1
    const uint32_t se = lop3_xor3(ca, cb, cc), ce = lop3_maj(ca, cb, cc);
    const uint32_t s2 = se ^ c1, s4 = lop3_xor_and(ce, se, c1);
    uint32_t d = imad(__popc(s4), 4u, imad(__popc(s2), 2u, __popc(s1)));
    if (EXACT) d = imad(__popc(lop3_and3(ce, se, c1)), 8u, d);
    return static_cast<int>(d);
}

// Packed 64-bit key: (distance << 32) | global train index.  Unsigned min == (dist, idx) lexicographic.
__device__ __forceinline__ unsigned long long make_key64(uint32_t dist, uint32_t idx) {
    return (static_cast<unsigned long long>(dist) << 32) | idx;
}

// Insert key k into the sorted pair (b0 <= b1).
__device__ __forceinline__ void top2_insert(uint32_t &b0, uint32_t &b1, uint32_t k) {
    b1 = min(b1, max(b0, k));
    b0 = min(b0, k);
}
__device__ __forceinline__ void top2_insert(unsigned long long &b0, unsigned long long &b1,
                                            unsigned long long k) {
    b1 = min(b1, max(b0, k));
    b0 = min(b0, k);
}

} // namespace plm
