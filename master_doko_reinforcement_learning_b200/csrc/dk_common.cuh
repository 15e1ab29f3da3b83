// dk_common.cuh — shared device helpers for the B200 Doppelkopf kernels: Philox4x32-10, bit select,
// card attribute arithmetic.  Everything is integer / bitwise; no tensor-core path exists or is wanted
// (the hot path is integer-issue bound).
//
// The functions are __host__ __device__ so that tests/hostsim can compile the SAME per-thread logic with
// g++ and run it on the CPU of the build container (which has no GPU).  That simulator is test
// infrastructure; the product library (libdoko_cuda.so) only ever launches the __global__ kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DK_HD __host__ __device__ __forceinline__
#define DK_D __device__ __forceinline__
#else
#define DK_HD inline
#define DK_D inline
#endif

namespace dk {

// ---- intrinsics with host shims ------------------------------------------------------------------
DK_HD uint32_t popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
DK_HD uint32_t popcll(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popcll(x);
#else
    return (uint32_t)__builtin_popcountll(x);
#endif
}
DK_HD uint32_t mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// Chained draw (DESIGN.md "Philox parity contract"): one 32-bit word serves several draws in a row.  (idx, rest) = (high, low) half of
// v * n: idx is the draw over n choices, the low half is the fractional part of v * n / 2^32 — uniform again on a grid of spacing
// n / 2^32 — and feeds the next draw.  After the draws n_0 .. n_{k-1} the word has become word * n_0 * .. * n_{k-1} mod 2^32, so a
// caller that resumes in the middle of a chain needs only that product; the relative bias of a draw is below n_0 * .. * n_k / 2^32.
DK_HD uint32_t draw_chain(uint32_t& v, uint32_t n) {
    const uint64_t p = (uint64_t)v * n;
    v = (uint32_t)p;
    return (uint32_t)(p >> 32);
}
DK_HD uint32_t ffs0(uint32_t x) {  // index of lowest set bit (x != 0)
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffs((int)x) - 1);
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}
DK_HD uint32_t ffs0ll(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffsll((long long)x) - 1);
#else
    return (uint32_t)__builtin_ctzll(x);
#endif
}

DK_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {  // low 32 bits of (hi:lo) >> sh, sh < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
}

// ---- Philox4x32-10 ----------------------------------------------------------------------------------
// Parity stream shared with the oracle (DESIGN.md "Philox parity contract"):
//   counter = (unit_lo, unit_hi, site<<16 | block, epoch), key = (seed_lo, seed_hi);
//   decision k of call-site class `site` inside one unit uses word (k & 3) of block (k >> 2),
//   mapped onto n choices by idx = mulhi(word, n); announcement decisions take one BIT each (fdo_rules.cuh AnnStream);
//   the deal draws three times from a word and a trick's four card draws share the word of the trick (draw_chain below).
enum Site : uint32_t {
    SITE_DEAL = 0, SITE_RESERVATION = 1, SITE_ANNOUNCEMENT = 2, SITE_CARD = 3,
    SITE_MATCH_CARD = 4, SITE_MATCH_RESERVATION = 5, SITE_ASSIGN = 6, SITE_STEP = 7, SITE_KEEP = 8, SITE_EXPAND = 9
};

struct U4 { uint32_t x, y, z, w; };

struct RngKey {
    uint32_t seed_lo, seed_hi, unit_lo, unit_hi, epoch;
};

DK_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}
DK_HD U4 rng_block(const RngKey& k, uint32_t site, uint32_t block) {
    return philox4x32_10(k.unit_lo, k.unit_hi, (site << 16) | block, k.epoch, k.seed_lo, k.seed_hi);
}
DK_HD uint32_t u4_word(const U4& b, uint32_t i) {  // dynamic word pick without local memory
    uint32_t lo = (i & 1) ? b.y : b.x;
    uint32_t hi = (i & 1) ? b.w : b.z;
    return (i & 2) ? hi : lo;
}

// ---- rank select ---------------------------------------------------------------------------------------
// Position of the k-th (0-based, counted from the LSB) set bit of a 32-bit mask; k < popc(x).
// The reference picks MSB-first (rs-game-utils/src/bit_flag.rs:104-171: rank 0 = highest set bit), so
// callers pass k = popc(x) - 1 - idx.
DK_HD uint32_t select_lsb(uint32_t x, uint32_t k) {
    uint32_t pos = 0, c;
    c = popc(x & 0xFFFFu); if (k >= c) { k -= c; x >>= 16; pos += 16; }
    c = popc(x & 0xFFu);   if (k >= c) { k -= c; x >>= 8;  pos += 8; }
    c = popc(x & 0xFu);    if (k >= c) { k -= c; x >>= 4;  pos += 4; }
    c = popc(x & 0x3u);    if (k >= c) { k -= c; x >>= 2;  pos += 2; }
    c = x & 1u;            if (k >= c) { pos += 1; }
    return pos;
}
// 24-bit card masks: three levels of 12/6/3 then two single-bit steps.
DK_HD uint32_t select_lsb24(uint32_t x, uint32_t k) {
    uint32_t pos = 0, c;
    c = popc(x & 0xFFFu); if (k >= c) { k -= c; x >>= 12; pos += 12; }
    c = popc(x & 0x3Fu);  if (k >= c) { k -= c; x >>= 6;  pos += 6; }
    c = popc(x & 0x7u);   if (k >= c) { k -= c; x >>= 3;  pos += 3; }
    c = x & 1u;           if (k >= c) { k -= c; x >>= 1;  pos += 1; }
    c = x & 1u;           if (k >= c) { pos += 1; }
    return pos;
}
// The reference's random pick: index idx counted from the MOST significant set bit.
DK_HD uint32_t pick_msb_rank24(uint32_t mask, uint32_t idx) { return select_lsb24(mask, popc(mask) - 1u - idx); }

// Rank select with a 64-entry table for the last level (the lock-step card loop's version: two popc levels 12 / 6, then one
// shared-memory lookup on the idle LSU pipe instead of three more compare/select levels on the saturated ALU pipe).
// Table word b (b = 6-bit mask): bits 3j..3j+2 = position of the j-th set bit of b counted from the LSB.
DK_HD uint32_t rank_lut6_entry(uint32_t b) {
    uint32_t e = 0, j = 0;
    for (uint32_t pos = 0; pos < 6u; ++pos) if ((b >> pos) & 1u) { e |= pos << (3u * j); j++; }
    return e;
}
// The lookup tables of the playout kernels, one word array: built once on the host (lut_word, fdo_rules.cuh), kept in device memory
// and copied into shared memory by every block (kernels.cuh stage_lut: one bulk copy of a prefix of the image).  Three regions, so that a
// kernel stages only what it reads:
//   CARD   [0,216) 9 game types x 24 cards: what a card means as the FIRST card of a trick (lead_lut_entry) | [216,223)
//          fdo_thr_lut_word | [224,288) rank_lut6_entry | [288,1128) 7 trump sets x 5 lead classes x 24 cards x 32 bit: the card's
//          record for the trick accumulator (pow_lut_entry)
//   SEL12  [1128,9320) the 12-bit rank-select table (4096 x 64 bit, sel12_entry) of the kernels that can afford 32 KB more shared
//          memory per block; the others pick cards through the 64-entry table above
//   ANN    [9320,9512) call thresholds of both teams by (wedding shift, re level, kontra level) (fdo_thr2_lut_word) |
//          [9512,9576) 256 bytes: who calls in a segment of an announcement round, by (eligible seats, decision bits)
//          (fdo_seg_lut_byte) — read only by the announcement replay, i.e. by the kernels that report game steps or play with calls
constexpr uint32_t CARD_LUT_WORDS = 1128u;
constexpr uint32_t LEAD_LUT_BASE = 0u;
constexpr uint32_t THR_LUT_BASE = 216u;
constexpr uint32_t RANK_LUT_BASE = 224u;
constexpr uint32_t POW_LUT_BASE = 288u;
constexpr uint32_t SEL12_LUT_BASE = 1128u;
constexpr uint32_t SEL12_WORDS = 8192u;
constexpr uint32_t ANN_LUT_BASE = SEL12_LUT_BASE + SEL12_WORDS;
constexpr uint32_t ANN_LUT_WORDS = 256u;
constexpr uint32_t THR2_LUT_BASE = ANN_LUT_BASE;
constexpr uint32_t SEG_LUT_BASE = ANN_LUT_BASE + 192u;
constexpr uint32_t FULL_LUT_WORDS = ANN_LUT_BASE + ANN_LUT_WORDS;
DK_HD uint32_t select_lsb24_lut(uint32_t x, uint32_t k, const uint32_t* __restrict__ lut) {
    uint32_t pos = 0, c;
    c = popc(x & 0xFFFu); if (k >= c) { k -= c; x >>= 12; pos = 12u; }
    c = popc(x & 0x3Fu);  if (k >= c) { k -= c; x >>= 6;  pos += 6u; }
    return pos + ((lut[RANK_LUT_BASE + (x & 63u)] >> (3u * k)) & 7u);
}
DK_HD uint32_t pick_msb_rank24_lut(uint32_t mask, uint32_t idx, const uint32_t* __restrict__ lut) {
    return select_lsb24_lut(mask, popc(mask) - 1u - idx, lut);
}

// One-level rank select: entry h (a 12-bit half of the mask), nibble j = position of the j-th set bit counted from the MSB.
// One popc decides the half; the position comes out of ONE 64-bit shared-memory load (the LSU pipe has room, the ALU pipe is the
// limiter: profiles/r01_fdo_playout_v7_attribution.json).  14 instead of 25 instructions per pick.
DK_HD uint64_t sel12_entry(uint32_t h) {
    uint64_t e = 0;
    uint32_t j = 0;
    for (int pos = 11; pos >= 0; --pos) if ((h >> pos) & 1u) { e |= (uint64_t)pos << (4u * j); j++; }
    return e;
}
DK_HD uint32_t pick_msb_rank24_tab(uint32_t mask, uint32_t idx, const uint64_t* __restrict__ sel12) {
    const uint32_t hi = mask >> 12, cu = popc(hi);
    const bool up = idx < cu;
    const uint32_t h = up ? hi : (mask & 0xFFFu), kk = up ? idx : idx - cu;
    const uint32_t p = (uint32_t)(sel12[h] >> (4u * kk)) & 15u;
    return up ? p + 12u : p;
}

// The reference's random pick on a 39-bit action mask: index from the most significant set bit (bit_flag.rs:86-94,104-171).
DK_HD uint32_t pick_msb_rank64(uint64_t mask, uint32_t idx) {
    uint32_t lo = (uint32_t)mask, hi = (uint32_t)(mask >> 32);
    uint32_t k = popc(lo) + popc(hi) - 1u - idx, cl = popc(lo);
    return k < cl ? select_lsb(lo, k) : 32u + select_lsb(hi, k - cl);
}

// ---- card arithmetic ------------------------------------------------------------------------------------
// card id c = suit*6 + rank; suits ♦0 ♥1 ♣2 ♠3; ranks 9,10,J,Q,K,A = 0..5 (rs-full-doko/src/card/cards.rs:7-36)
DK_HD uint32_t card_suit(uint32_t c) { return (c * 43u) >> 8; }           // c / 6 for c < 24
DK_HD uint32_t card_eyes_by_rank(uint32_t rank) { return (0xB432A0u >> (4u * rank)) & 15u; }  // 0,10,2,3,4,11 (card_to_eyes.rs:8-40)

// Trump masks per FdoGameType (rs-full-doko/src/card/card_color_masks.rs:7-246); plain colours are the
// natural suits minus the trumps (equivalent to card_to_color.rs:11-258, checked by tests/test_oracle_rule_tables.py).
DK_HD uint32_t trump_mask_for_game_type(uint32_t gt) {
    const uint32_t JACKS = (1u << 2) | (1u << 8) | (1u << 14) | (1u << 20);
    const uint32_t QUEENS = (1u << 3) | (1u << 9) | (1u << 15) | (1u << 21);
    const uint32_t H10 = 1u << 7;
    const uint32_t JQ10 = JACKS | QUEENS | H10;
    // suit trumps: all six cards of the suit (J/Q already trump; the ♥10 is trump anyway)
    uint32_t m;
    switch (gt) {
        case 0: case 1: case 2: m = JQ10 | 0x00003Fu; break;   // Normal, Wedding, ♦-Solo
        case 3: m = JQ10 | 0x000FC0u; break;                   // ♥-Solo
        case 4: m = JQ10 | 0xFC0000u; break;                   // ♠-Solo (suit 3)
        case 5: m = JQ10 | 0x03F000u; break;                   // ♣-Solo (suit 2)
        case 6: m = 0u; break;                                 // Trumpless
        case 7: m = QUEENS; break;                             // Queens-Solo
        default: m = JACKS; break;                             // Jacks-Solo
    }
    return m;
}
// Mask of the cards that FOLLOW a trick led with card `first`:
// trump lead → all trumps; plain lead → the cards of its suit that are not trump.
DK_HD uint32_t follow_mask(uint32_t first_card, uint32_t trump) {
    uint32_t suit_cards = 0x3Fu << (6u * card_suit(first_card));
    return ((trump >> first_card) & 1u) ? trump : (suit_cards & ~trump);
}
// Strength of a card inside a trick.  `power(b) > power(a)` (strict) <=> is_greater_in_trick(b, a, colour, game_type)
// (rs-full-doko/src/card/card_in_trick_logic.rs:17-141): trumps beat plain cards and are ordered by trump_to_rank
// (:37-77), cards of the led plain colour are ordered by eyes (:135), anything else never wins.
DK_HD uint32_t card_power(uint32_t c, uint32_t trump, uint32_t follow) {
    uint32_t suit = card_suit(c), rank = c - 6u * suit;
    uint32_t bit = 1u << c;
    uint32_t so = suit ^ (suit >> 1);                       // ♦0 ♥1 ♠2 ♣3  (J/Q order ♦<♥<♠<♣)
    uint32_t plain = (0x310020u >> (4u * rank)) & 15u;      // 9:0 10:2 K:1 A:3
    uint32_t tp = rank == 2u ? 4u + so : (rank == 3u ? 8u + so : (c == 7u ? 12u : plain));
    uint32_t cp = 1u + card_eyes_by_rank(rank);
    return (trump & bit) ? 16u + tp : ((follow & bit) ? cp : 0u);
}

// Strength-in-trick table: row (trump set, lead class) x card.  Trump set = game type with Normal / Wedding / ♦-solo folded into one
// (pow_trump_set: they share their trumps), lead class = suit of the first card (0..3) when it is a plain card, 4 when it is a trump.  One 32-bit shared-memory load per card gives everything the trick accumulator needs, laid out so that the
// accumulator is ONE max, ONE add and one multiply-add per card (the ALU pipe limits the playout kernels; the compare / three selects
// / mask of the first form were 5 of its ~28 ALU instructions per card step):
//   bits 0-7 eyes | bit 8 the card is a ♦A | bits 9-10 zero (the caller ORs in 3 - position) | bits 11-15 card id | bits 16-23
//   card_power(card, trump mask of the game type, follow mask of the lead class).
// max over (entry | (3 - k) << 9) of the four cards = the winning card: the strength decides; cards of equal positive strength are
// copies of one card type (every strength > 0 belongs to one type within a row), so their low bits are equal and the position field
// makes the FIRST of them win (strict `>` in the reference); a strength-0 card never beats the first card, whose strength is positive.
constexpr uint32_t POW_K_SHIFT = 9u, POW_CARD_SHIFT = 11u, POW_PW_SHIFT = 16u, POW_FOX_BIT = 0x100u;
DK_HD uint32_t pow_trump_set(uint32_t gt) { return gt <= 2u ? 0u : gt - 2u; }           // 0 = the normal trumps, 1..6 = ♥ ♠ ♣ trumpless queens jacks solo
DK_HD uint32_t pow_lut_entry(uint32_t ts, uint32_t cls, uint32_t c) {
    const uint32_t trump = trump_mask_for_game_type(ts == 0u ? 0u : ts + 2u);
    const uint32_t follow = cls == 4u ? trump : ((0x3Fu << (6u * cls)) & ~trump);
    const uint32_t suit = card_suit(c);
    return card_eyes_by_rank(c - 6u * suit) | (c == 5u ? POW_FOX_BIT : 0u) | (c << POW_CARD_SHIFT) | (card_power(c, trump, follow) << POW_PW_SHIFT);
}
// The same record from a card played earlier (the bridge from a stored record to the playout form).
DK_HD uint32_t pow_entry_of(uint32_t c, uint32_t trump, uint32_t follow) {
    const uint32_t suit = card_suit(c);
    return card_eyes_by_rank(c - 6u * suit) | (c == 5u ? POW_FOX_BIT : 0u) | (c << POW_CARD_SHIFT) | (card_power(c, trump, follow) << POW_PW_SHIFT);
}
DK_HD uint32_t pow_row(uint32_t gt, uint32_t first_card, uint32_t first_suit, uint32_t trump) {
    return (pow_trump_set(gt) * 5u + (((trump >> first_card) & 1u) ? 4u : first_suit)) * 24u;
}
// What card c means as the first card of a trick in game type gt: the follow mask of the trick (bits 0-23) and its row of the
// strength table, as row / 8 in bits 24-31 — one shared-memory load + two decode instructions at the first card of every trick
// instead of the suit / trump select chain (8 instructions on the ALU pipe).
DK_HD uint32_t lead_lut_entry(uint32_t gt, uint32_t c) {
    const uint32_t trump = trump_mask_for_game_type(gt);
    return follow_mask(c, trump) | ((pow_row(gt, c, card_suit(c), trump) / 8u) << 24);
}
DK_HD uint32_t lead_lookup(const uint32_t* __restrict__ lut, uint32_t gt, uint32_t c) { return lut[LEAD_LUT_BASE + gt * 24u + c]; }
DK_HD uint32_t lead_follow(uint32_t e) { return e & 0xFFFFFFu; }
DK_HD uint32_t lead_row(uint32_t e) { return (e >> 24) * 8u; }
DK_HD uint32_t pow_lookup(const uint32_t* __restrict__ lut, uint32_t row, uint32_t c) { return lut[POW_LUT_BASE + row + c]; }
// Trick accumulator shared by both engines: `best` = max of (entry | (3 - k) << 9), `acc` = sum of the entries (its low byte is the
// trick's eyes: at most 44, no carry out of the byte matters), `fox` = sum of (entry & ♦A bit) << k.
DK_HD uint32_t pow_best_k(uint32_t best) { return 3u - ((best >> POW_K_SHIFT) & 3u); }
DK_HD uint32_t pow_best_card(uint32_t best) { return (best >> POW_CARD_SHIFT) & 31u; }

}  // namespace dk
