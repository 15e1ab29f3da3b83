// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// Restatement of the reference's bit-set helpers:
//   rs-game-utils/src/bit_flag.rs:20-171          (Bitflag<N>, select_by_rank, random_single)
//   rs-doko/src/util/bitflag/bitflag.rs:24-107    (free-function twin)
//   rs-doko/src/util/bitflag/select_rank.rs:2-68  (select_by_rank twin)
#pragma once
#include <cstdint>
#include "rng.hpp"

namespace oracle {

inline uint32_t popcount64(uint64_t v) { return (uint32_t)__builtin_popcountll(v); }

// select_by_rank(v, r): the bit of rank r counted FROM THE MOST SIGNIFICANT set bit
// (rank 0 = highest set bit).  Branch-free ladder as in rs-game-utils/src/bit_flag.rs:104-171
// (Stanford bithacks "SelectPosFromMSBRank").
inline uint64_t select_by_rank(uint64_t v, uint64_t r) {
    uint64_t s, t, a, b, c, d;
    r = r + 1;
    a = v - ((v >> 1) & (~0ull / 3));
    b = (a & (~0ull / 5)) + ((a >> 2) & (~0ull / 5));
    c = (b + (b >> 4)) & (~0ull / 0x11);
    d = (c + (c >> 8)) & (~0ull / 0x101);
    t = (d >> 32) + (d >> 48);
    s = 64;
    s -= ((t - r) & 256) >> 3; r -= (t & ((t - r) >> 8));
    t = (d >> (s - 16)) & 0xff;
    s -= ((t - r) & 256) >> 4; r -= (t & ((t - r) >> 8));
    t = (c >> (s - 8)) & 0xf;
    s -= ((t - r) & 256) >> 5; r -= (t & ((t - r) >> 8));
    t = (b >> (s - 4)) & 0x7;
    s -= ((t - r) & 256) >> 6; r -= (t & ((t - r) >> 8));
    t = (a >> (s - 2)) & 0x3;
    s -= ((t - r) & 256) >> 7; r -= (t & ((t - r) >> 8));
    t = (v >> (s - 1)) & 0x1;
    s -= ((t - r) & 256) >> 8;
    return 1ull << (s - 1);
}

// Bitflag::random_single (bit_flag.rs:86-94) / bitflag_random_single (bitflag.rs:100-107):
// index = gen_range(0..popcount) then MSB-first rank select.
inline uint64_t random_single(uint64_t bits, Rng& rng, Site site) {
    uint32_t index = rng.below(site, popcount64(bits));
    return select_by_rank(bits, index);
}

// Bitflag::to_vec order (ascending bit), bit_flag.rs:67-82.
inline int bitflag_to_vec(uint64_t bits, uint64_t* out) {
    int n = 0;
    uint64_t bit = 1;
    while (bit <= bits && bit != 0) {
        if ((bits & bit) == bit) out[n++] = bit;
        bit <<= 1;
    }
    return n;
}

}  // namespace oracle
