// encode.cuh — observation encoders on the dk_state record.  Each produces the token VALUES (all < 256) through an
// output functor; the kernels stage them as bytes in shared memory and then widen to the reference's i64 rows with
// coalesced 64-bit stores (the HBM-bound part, SURVEY.md §8d: 2488 B per observation).
#pragma once
#include "state_ops.cuh"

namespace dk {

// 24 five-bit entries packed into two 60-bit constants.
DK_HD uint32_t lut24x5(uint64_t lo, uint64_t hi, uint32_t c) {
    uint64_t w = c < 12u ? lo : hi;
    uint32_t sh = 5u * (c < 12u ? c : c - 12u);
    return (uint32_t)(w >> sh) & 31u;
}
#define DK_PACK12(a, b, c, d, e, f, g, h, i, j, k, l)                                                                              \
    ((uint64_t)(a) | ((uint64_t)(b) << 5) | ((uint64_t)(c) << 10) | ((uint64_t)(d) << 15) | ((uint64_t)(e) << 20) | ((uint64_t)(f) << 25) | \
     ((uint64_t)(g) << 30) | ((uint64_t)(h) << 35) | ((uint64_t)(i) << 40) | ((uint64_t)(j) << 45) | ((uint64_t)(k) << 50) | ((uint64_t)(l) << 55))

// encode_reservation_or_card_or_none_card (rs-doko-networks/src/full_doko/var2/encode_reservation_or_card_or_none.rs:70-107)
DK_HD uint32_t fdo_card_token(uint32_t c) {
    //                 ♦9  ♦10 ♦J ♦Q ♦K  ♦A  ♥9  ♥10 ♥J ♥Q ♥K  ♥A
    const uint64_t lo = DK_PACK12(13, 11, 9, 5, 12, 10, 24, 1, 8, 4, 23, 22);
    //                 ♣9  ♣10 ♣J ♣Q ♣K  ♣A  ♠9  ♠10 ♠J ♠Q ♠K  ♠A
    const uint64_t hi = DK_PACK12(17, 15, 6, 2, 16, 14, 21, 19, 7, 3, 20, 18);
    return lut24x5(lo, hi, c);
}

// encode_state_pi (rs-doko-networks/src/full_doko/var1/encode_pi.rs:27-216) with obs = observation_for_current_player():
// 62 slots x 5 channels (token, position, player, sub-position, team) + phase = 311 values.
// The sink receives out.slot(n, token, position, player, sub-position, team) for the 62 slots and out.phase(v); row element
// indices are n, 62+n, 124+n, 186+n, 248+n and 310.
// Every access to the record uses a compile-time offset (the card / call loops run over all 48 / 10 positions under a predicate, the
// hands are rotated with selects): the 128-byte record stays in registers instead of local memory, which matters because the encode
// kernels configure almost the whole L1 as shared memory.
template <class Out>
DK_HD void fdo_encode_pi(const dk_state& s, Out& out) {
    const uint32_t cur = st_phase(s) == DK_PHASE_FINISHED ? 0u : st_cur(s);   // current_player.unwrap_or(BOTTOM) (:31-33)
    uint32_t n = 0;
    auto push = [&](uint32_t tok, uint32_t pos, uint32_t ply, uint32_t sub, uint32_t team) {
        out.slot(n, tok, pos, ply, sub, team);
        n++;
    };
    const uint32_t start = st_game_start(s), nres = s.n_reservations;
    const uint32_t res4 = (uint32_t)s.reservations[0] | ((uint32_t)s.reservations[1] << 8) | ((uint32_t)s.reservations[2] << 16) | ((uint32_t)s.reservations[3] << 24);
#pragma unroll
    for (uint32_t i = 0; i < 4u; ++i) {                                       // real reservations in play order (:43-81)
        const bool made = i < nres;
        push(made ? 25u + ((res4 >> (8u * i)) & 255u) : 36u, made ? i + 1u : 0u, made ? ((start + i - cur) & 3u) + 1u : 0u, 0u, 0u);
    }
    const uint32_t ci = s.card_index, tricks = s.tricks;
    const uint32_t* cards32 = reinterpret_cast<const uint32_t*>(s.cards);
#pragma unroll
    for (uint32_t t = 0; t < 12u; ++t) {                                      // played cards (:83-99), trick by trick
        if (4u * t < ci) {                                                    // (uniform enough: games of a batch are at similar depths)
            const uint32_t quad = cards32[t], lead = (tricks >> (2u * t)) & 3u;
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) {
                const uint32_t j = 4u * t + k;
                if (j < ci) out.slot(4u + j, fdo_card_token((quad >> (8u * k)) & 255u), j + 5u, ((lead + k - cur) & 3u) + 1u, 0u, 0u);
            }
        }
    }
    n = 4u + ci;
    // the four hands from the seat to move (:102-121): rotate with selects instead of indexing by (cur + i) & 3
    uint64_t h0 = s.hands[0], h1 = s.hands[1], h2 = s.hands[2], h3 = s.hands[3];
    if (cur & 1u) { uint64_t t = h0; h0 = h1; h1 = h2; h2 = h3; h3 = t; }
    if (cur & 2u) { uint64_t t0 = h0, t1 = h1; h0 = h2; h1 = h3; h2 = t0; h3 = t1; }
#pragma unroll
    for (uint32_t i = 0; i < 4u; ++i) {
        const uint64_t h = i == 0u ? h0 : (i == 1u ? h1 : (i == 2u ? h2 : h3));
        uint64_t b = h;
        while (b) {
            uint32_t pos = ffs0ll(b);
            b &= b - 1ull;
            uint32_t c = pos < 24u ? pos : pos - 24u;
            uint32_t second = (pos >= 24u && ((h >> c) & 1ull)) ? 1u : 0u;     // already_encoded_cards.contains(card)
            push(fdo_card_token(c), 53u + i, 0u, 11u + second, 0u);
        }
    }
    const uint32_t n_calls = st_n_calls(s), re = st_re_mask(s);
    const uint32_t* ann32 = reinterpret_cast<const uint32_t*>(s.announcements);
    uint32_t sub = 0, last = 0xFFFFFFFFu;
#pragma unroll
    for (uint32_t a = 0; a < 10u; ++a) {                                      // calls (:139-165); position = raw card_index + 1; then padding (:167-179)
        const uint32_t v = (ann32[a >> 1] >> (16u * (a & 1u))) & 0xFFFFu, cidx = v & 63u, seat = (v >> 6) & 3u, lvl = (v >> 8) & 7u;
        if (a < n_calls) {
            if (cidx != last) { last = cidx; sub = 0; }
            push(lvl == 6u ? 38u : 37u + lvl, cidx + 1u, ((seat - cur) & 3u) + 1u, sub + 1u, ((re >> seat) & 1u) ? 1u : 2u);
            sub++;
        } else push(37u, 0u, 0u, 0u, 0u);
    }
    out.phase(st_phase(s));                                                   // encode_phase (var1/phase.rs:9-18)
}

// encode_state_ipi (rs-doko-networks/src/full_doko/var1/encode_ipi.rs:48-306) with obs = observation_for_current_player(): the
// imperfect-information layout of the autoregressive hand predictor.  Slots: the four VISIBLE reservations in play order (a
// NotRevealed one replaced by the guess, :68-79), played cards, the observer's own hand, then for the other seats from the observer
// on: the cards guessed so far followed by one "unknown" slot per card still missing (:132-174), 10 call slots; the last value is
// the seat the next card is guessed for, relative to the observer (:232).  assumed hands are bitboards by ABSOLUTE seat (the
// observer's entry is ignored), assumed reservations DK_RES_* by absolute seat or DK_RES_NONE.
// Returns 0, or 1 when the reference would panic (a guessed hand larger than the real one ⇒ `hand.len() - assumed.len()` underflows,
// or the slot count is not 62); the row is then zero-padded / truncated.
// Like fdo_encode_pi, every access to the record and to the guesses uses a compile-time offset (seat rotations by selects, the card
// and call loops over all positions under a predicate, the guessed reservations as one packed word): nothing lives in local memory.
template <class Out>
DK_HD uint32_t fdo_encode_ipi(const dk_state& s, const uint64_t assumed[4], const uint8_t assumed_res[4], uint32_t next_player, Out& out) {
    const uint32_t cur = st_phase(s) == DK_PHASE_FINISHED ? 0u : st_cur(s);
    uint32_t n = 0, err = 0;
    auto push = [&](uint32_t tok, uint32_t pos, uint32_t ply, uint32_t sub, uint32_t team) {
        if (n < 62u) out.slot(n, tok, pos, ply, sub, team); else err = 1u;
        n++;
    };
    const uint32_t start = st_game_start(s), nres = s.n_reservations;
    const bool completed = nres == 4u;
    const uint32_t ares4 = (uint32_t)assumed_res[0] | ((uint32_t)assumed_res[1] << 8) | ((uint32_t)assumed_res[2] << 16) | ((uint32_t)assumed_res[3] << 24);
    bool solo_seen = false;
#pragma unroll
    for (uint32_t i = 0; i < 4u; ++i) {                                       // get_visible_reservations(observer) (visible_reservations_logic.rs:7-70)
        const uint32_t seat = (start + i) & 3u;
        uint32_t tok = 35u;                                                   // NoneYet
        if (i < nres) {
            const uint32_t code = s.reservations[i];
            if (code == 0u) tok = 25u;
            else if (code == 1u) tok = (completed || seat == cur) ? 26u : 34u;
            else if (completed && !solo_seen) { tok = 25u + code; solo_seen = true; }
            else tok = 34u;                                                   // NotRevealed
            const uint32_t guess = (ares4 >> (8u * seat)) & 255u;
            if (tok == 34u && guess != 0xFFu) tok = 25u + guess;
        }
        push(tok, i + 1u, ((seat - cur) & 3u) + 1u, 0u, 0u);
    }
    const uint32_t ci = s.card_index, tricks = s.tricks;
#pragma unroll
    for (uint32_t t = 0; t < 12u; ++t) {                                      // played cards (:93-110), trick by trick
        if (4u * t < ci) {
            const uint32_t quad = st_word_of(s.cards, t), lead = (tricks >> (2u * t)) & 3u;
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) {
                const uint32_t j = 4u * t + k;
                if (j < ci) out.slot(4u + j, fdo_card_token((quad >> (8u * k)) & 255u), j + 5u, ((lead + k - cur) & 3u) + 1u, 0u, 0u);
            }
        }
    }
    n = 4u + ci;
    // real and guessed hands from the observer on: rotate with selects instead of indexing by (cur + i) & 3
    uint64_t r0 = s.hands[0], r1 = s.hands[1], r2 = s.hands[2], r3 = s.hands[3];
    uint64_t g0 = assumed[0], g1 = assumed[1], g2 = assumed[2], g3 = assumed[3];
    if (cur & 1u) { uint64_t t = r0; r0 = r1; r1 = r2; r2 = r3; r3 = t; t = g0; g0 = g1; g1 = g2; g2 = g3; g3 = t; }
    if (cur & 2u) { uint64_t t0 = r0, t1 = r1; r0 = r2; r1 = r3; r2 = t0; r3 = t1; t0 = g0; t1 = g1; g0 = g2; g1 = g3; g2 = t0; g3 = t1; }
#pragma unroll
    for (uint32_t i = 0; i < 4u; ++i) {
        const uint64_t real = i == 0u ? r0 : (i == 1u ? r1 : (i == 2u ? r2 : r3));
        const uint64_t h = i == 0u ? real : (i == 1u ? g1 : (i == 2u ? g2 : g3));
        for (uint64_t b = h; b; b &= b - 1ull) {
            uint32_t pos = ffs0ll(b);
            uint32_t c = pos < 24u ? pos : pos - 24u;
            uint32_t second = (pos >= 24u && ((h >> c) & 1ull)) ? 1u : 0u;
            push(fdo_card_token(c), 53u + i, 0u, 11u + second, 0u);
        }
        if (i > 0u) {
            const uint32_t have = popcll(real), guessed = popcll(h);
            if (guessed > have) err = 1u;
            for (uint32_t k = guessed; k < have; ++k) push(0u, 53u + i, 0u, 0u, 0u);   // still unknown (:160-173)
        }
    }
    const uint32_t n_calls = st_n_calls(s), re = st_re_mask(s);
    uint32_t sub = 0, last = 0xFFFFFFFFu;
#pragma unroll
    for (uint32_t a = 0; a < 10u; ++a) {                                      // calls (:191-215), then padding
        const uint32_t v = s.announcements[a], cidx = v & 63u, seat = (v >> 6) & 3u, lvl = (v >> 8) & 7u;
        if (a < n_calls) {
            if (cidx != last) { last = cidx; sub = 0; }
            push(lvl == 6u ? 38u : 37u + lvl, cidx + 1u, ((seat - cur) & 3u) + 1u, sub + 1u, ((re >> seat) & 1u) ? 1u : 2u);
            sub++;
        } else push(37u, 0u, 0u, 0u, 0u);
    }
    if (n != 62u) err = 1u;
    for (; n < 62u; ++n) out.slot(n, 0u, 0u, 0u, 0u, 0u);
    out.phase(((next_player - cur) & 3u) + 1u);
    return err;
}

// encode_state / encode_state_with_reservations (rs-doko-embeddings/src/encode_state.rs:84-317): 110 / 114 values.
template <class Out>
DK_HD void doko_encode(const dk_state& s, bool with_reservations, Out& out) {
    const uint32_t phase = st_phase(s);
    const uint32_t cur = phase == DK_PHASE_FINISHED ? 0u : st_cur(s);
    out(0u, phase == DK_PHASE_RESERVATION ? 0u : (phase == DK_PHASE_PLAY_CARD ? 1u : 2u));   // DoPhase
    out(1u, ((st_game_start(s) - cur) & 3u) + 1u);
    const uint32_t nt = st_n_tricks(s);
    for (uint32_t t = 0; t < 12u; ++t) out(2u + t, t < nt ? ((st_trick_start(s, t) - cur) & 3u) + 1u : 0u);
    const uint32_t ci = s.card_index;
    for (uint32_t j = 0; j < 48u; ++j) out(14u + j, j < ci ? (uint32_t)s.cards[j] + 1u : 0u);
    // hands: descending card_to_rank_in_normal_game (:17-50), doubles adjacent, seats relative to the seat to move
    //   ♥10 ♣Q ♠Q ♥Q ♦Q ♣J ♠J ♥J ♦J ♦A ♦10 ♦K ♦9 ♣A ♣10 ♣K ♣9 ♠A ♠10 ♠K ♠9 ♥A ♥K ♥9
    const uint64_t lo = DK_PACK12(7, 15, 21, 9, 3, 14, 20, 8, 2, 5, 1, 4), hi = DK_PACK12(0, 17, 13, 16, 12, 23, 19, 22, 18, 11, 10, 6);
    for (uint32_t p = 0; p < 4u; ++p) {
        uint32_t any = hand_any24(s.hands[p]), both = hand_both24(s.hands[p]);
        uint32_t base = 62u + 12u * ((p - cur) & 3u), m = 0;
        for (uint32_t r = 0; r < 24u; ++r) {
            uint32_t c = lut24x5(lo, hi, r);
            if ((any >> c) & 1u) { out(base + m, c + 1u); m++; if ((both >> c) & 1u) { out(base + m, c + 1u); m++; } }
        }
        for (; m < 12u; ++m) out(base + m, 0u);
    }
    if (with_reservations) {
        // get_visible_reservations quirk (rs-doko/src/reservation/visible_reservations_logic.rs:6-34): slots are
        // compared with the ABSOLUTE observing seat and stored by slot.
        bool completed = s.n_reservations == 4u;
        for (uint32_t i = 0; i < 4u; ++i) {
            uint32_t v = 0;
            if (i < s.n_reservations) v = s.reservations[i] == 1u ? 2u : ((completed || i == cur) ? 3u : 1u);
            out(110u + i, v);
        }
    }
}

}  // namespace dk
