// state_ops.cuh — operations on the 128-byte dk_state record (include/doko_cuda.h): the batch form of the reference's
// by-value FdoState / DoState seam (legal-action query, apply action, terminal check, rewards) plus the bridges between the
// stored record and the register-resident playout form (fdo_rules.cuh / doko_rules.cuh).
#pragma once
#include "../../include/doko_cuda.h"
#include "dk_common.cuh"
#include "doko_rules.cuh"
#include "fdo_rules.cuh"

namespace dk {

// ---- field accessors -------------------------------------------------------------------------------------------------------
DK_HD uint32_t st_phase(const dk_state& s) { return s.meta & 3u; }
DK_HD uint32_t st_cur(const dk_state& s) { return (s.meta >> 2) & 3u; }
DK_HD uint32_t st_game_start(const dk_state& s) { return (s.meta >> 4) & 3u; }
DK_HD uint32_t st_gt(const dk_state& s) { return (s.meta >> 6) & 15u; }
DK_HD uint32_t st_team_tag(const dk_state& s) { return (s.meta >> 10) & 3u; }
DK_HD uint32_t st_wed_seat(const dk_state& s) { return (s.meta >> 12) & 3u; }
DK_HD uint32_t st_solved_idx(const dk_state& s) { return (s.meta >> 14) & 3u; }
DK_HD uint32_t st_re_mask(const dk_state& s) { return (s.meta >> 16) & 15u; }
DK_HD uint32_t st_re_low(const dk_state& s) { return (s.meta >> 20) & 7u; }
DK_HD uint32_t st_ko_low(const dk_state& s) { return (s.meta >> 23) & 7u; }
DK_HD uint32_t st_turns(const dk_state& s) { return (s.meta >> 26) & 7u; }
DK_HD uint32_t st_ann_start(const dk_state& s) { return (s.meta >> 29) & 3u; }
DK_HD uint32_t st_n_tricks(const dk_state& s) { return (s.tricks >> 24) & 15u; }
DK_HD uint32_t st_n_calls(const dk_state& s) { return (s.tricks >> 28) & 15u; }
DK_HD uint32_t st_trick_start(const dk_state& s, uint32_t t) { return (s.tricks >> (2u * t)) & 3u; }
DK_HD void st_set(uint32_t& word, uint32_t shift, uint32_t width, uint32_t v) {
    uint32_t m = ((1u << width) - 1u) << shift;
    word = (word & ~m) | ((v << shift) & m);
}
DK_HD uint32_t hand_any24(uint64_t h) { return (uint32_t)((h | (h >> 24)) & 0xFFFFFFull); }
DK_HD uint32_t hand_both24(uint64_t h) { return (uint32_t)((h & (h >> 24)) & 0xFFFFFFull); }

// ---- run-time indexed fields through compile-time offsets ----------------------------------------------------------------------
// A seat / card position / call number known only at run time would make the compiler put the whole 128-byte record into local
// memory (a dynamically indexed array cannot live in registers).  The transition functions below therefore reach those fields
// through select chains over compile-time offsets: the record stays in registers in every kernel that only uses these functions
// (0-byte stack frame for dk_apply / dk_legal_mask / the fused step + encode; the kernels configure most of the SM's L1 as shared
// memory, so a local-memory record was served from L2).  The host simulator compiles the same code.
// Every function that needs them takes IDX: false (default) = the select chains, true = plain indexed accesses for the kernels whose
// register budget is better spent elsewhere and that keep the record in local memory anyway (the UCT search: 64 registers at 16
// blocks per SM; the PIMC evaluator: the whole rollout state is live next to the record) — measured, profiles/r01_kernels_v17.json.
template <bool IDX = false>
DK_HD uint64_t st_hand(const dk_state& s, uint32_t p) {
    if constexpr (IDX) return s.hands[p];
    else {
        const uint64_t a = (p & 1u) ? s.hands[1] : s.hands[0], b = (p & 1u) ? s.hands[3] : s.hands[2];
        return (p & 2u) ? b : a;
    }
}
template <bool IDX = false>
DK_HD void st_set_hand(dk_state& s, uint32_t p, uint64_t h) {
    if constexpr (IDX) s.hands[p] = h;
    else {
#pragma unroll
        for (uint32_t q = 0; q < 4u; ++q) s.hands[q] = q == p ? h : s.hands[q];
    }
}
DK_HD uint32_t st_word_of(const uint8_t* bytes, uint32_t w) {            // 32-bit word w of a byte field (little endian), w compile-time
    return (uint32_t)bytes[4u * w] | ((uint32_t)bytes[4u * w + 1u] << 8) | ((uint32_t)bytes[4u * w + 2u] << 16) | ((uint32_t)bytes[4u * w + 3u] << 24);
}
DK_HD void st_set_word_of(uint8_t* bytes, uint32_t w, uint32_t v) {
    bytes[4u * w] = (uint8_t)v; bytes[4u * w + 1u] = (uint8_t)(v >> 8); bytes[4u * w + 2u] = (uint8_t)(v >> 16); bytes[4u * w + 3u] = (uint8_t)(v >> 24);
}
// the four cards of trick t (byte k = k-th card; 0xFF = not played yet)
template <bool IDX = false>
DK_HD uint32_t st_quad(const dk_state& s, uint32_t t) {
    if constexpr (IDX) return (uint32_t)s.cards[4u * t] | ((uint32_t)s.cards[4u * t + 1u] << 8) | ((uint32_t)s.cards[4u * t + 2u] << 16) | ((uint32_t)s.cards[4u * t + 3u] << 24);
    else {
        uint32_t q = 0;
#pragma unroll
        for (uint32_t w = 0; w < 12u; ++w) q = w == t ? st_word_of(s.cards, w) : q;
        return q;
    }
}
template <bool IDX = false>
DK_HD uint32_t st_card(const dk_state& s, uint32_t i) { return IDX ? (uint32_t)s.cards[i] : (st_quad<false>(s, i >> 2) >> (8u * (i & 3u))) & 255u; }
template <bool IDX = false>
DK_HD void st_set_card(dk_state& s, uint32_t i, uint32_t c) {
    if constexpr (IDX) s.cards[i] = (uint8_t)c;
    else {
        const uint32_t sh = 8u * (i & 3u), keep = ~(255u << sh), ins = c << sh, t = i >> 2;
#pragma unroll
        for (uint32_t w = 0; w < 12u; ++w) { const uint32_t old = st_word_of(s.cards, w); st_set_word_of(s.cards, w, w == t ? (old & keep) | ins : old); }
    }
}
template <bool IDX = false>
DK_HD void st_set_call(dk_state& s, uint32_t n, uint32_t v) {            // announcements[n] = v
    if constexpr (IDX) s.announcements[n] = (uint16_t)v;
    else {
#pragma unroll
        for (uint32_t a = 0; a < 12u; ++a) s.announcements[a] = a == n ? (uint16_t)v : s.announcements[a];
    }
}
template <bool IDX = false>
DK_HD void st_push_reservation(dk_state& s, uint32_t code) {             // reservations[n_reservations++] = code
    const uint32_t n = s.n_reservations;
    if constexpr (IDX) s.reservations[n] = (uint8_t)code;
    else {
#pragma unroll
        for (uint32_t i = 0; i < 4u; ++i) s.reservations[i] = i == n ? (uint8_t)code : s.reservations[i];
    }
    s.n_reservations = (uint8_t)(n + 1u);
}
template <bool IDX = false>
DK_HD void st_add_eyes(dk_state& s, uint32_t seat, uint32_t e) {
    if constexpr (IDX) s.eyes[seat] = (uint8_t)(s.eyes[seat] + e);
    else {
#pragma unroll
        for (uint32_t p = 0; p < 4u; ++p) s.eyes[p] = (uint8_t)(s.eyes[p] + (p == seat ? e : 0u));
    }
}

DK_HD void st_clear(dk_state& s) {
    for (int p = 0; p < 4; ++p) { s.hands[p] = 0; s.reservations[p] = 0xFF; s.eyes[p] = 0; s.points[p] = 0; }
    for (int i = 0; i < 48; ++i) s.cards[i] = 0xFF;
    for (int i = 0; i < 12; ++i) s.announcements[i] = 0xFFFF;
    s.tricks = 0; s.num_tricks = 0; s.card_index = 0; s.n_reservations = 0; s.meta = 0;
}
// new_game_from_hand_and_start_player (rs-full-doko/src/state/state.rs:125-166, rs-doko/src/state/state.rs:115-157)
DK_HD void st_new_game(dk_state& s, const uint64_t hands[4], uint32_t start) {
    st_clear(s);
    for (int p = 0; p < 4; ++p) s.hands[p] = hands[p];
    uint32_t m = 0;
    m |= (uint32_t)DK_PHASE_RESERVATION;
    m |= start << 2; m |= start << 4; m |= (uint32_t)DK_GT_NONE << 6; m |= (uint32_t)DK_TEAM_IN_RESERVATIONS << 10;
    s.meta = m;
}

// =====================================================================================================================
// rs-full-doko
// =====================================================================================================================
DK_HD uint32_t fdo_res_code_from_action(uint32_t a) {   // action 24..32 → FdoReservation code (reservation.rs:11-24)
    // 24 Healthy 0, 25 Wedding 1, 26 ♦ 2, 27 ♥ 3, 28 ♠ 4, 29 ♣ 5, 30 Trumpless 8, 31 Queens 6, 32 Jacks 7
    return a <= 29u ? a - 24u : (a == 30u ? 8u : a - 25u);
}
DK_HD uint32_t fdo_action_from_res_code(uint32_t r) { return r <= 5u ? r + 24u : (r == 8u ? 30u : r + 25u); }
DK_HD uint32_t fdo_gt_from_res_code(uint32_t r) {       // solo reservation → FdoGameType (reservation_winning_logic.rs:16-32)
    // ♦2→2 ♥3→3 ♠4→4 ♣5→5 Queens6→7 Jacks7→8 Trumpless8→6
    return r <= 5u ? r : (r == 8u ? 6u : r + 1u);
}
// The call a seat may make in the announcement phase (level 1..5) or 0.
template <bool IDX = false>
DK_HD uint32_t fdo_state_allowed_call(const dk_state& s, uint32_t seat) {
    uint32_t tag = st_team_tag(s);
    if (tag == TEAM_WEDDING_UNSOLVED || tag == TEAM_IN_RESERVATIONS) return 0;
    uint32_t w = tag == TEAM_WEDDING_SOLVED ? st_solved_idx(s) : 0u;
    bool is_re = (st_re_mask(s) >> seat) & 1u;
    uint32_t m = is_re ? st_re_low(s) : st_ko_low(s), e = is_re ? st_ko_low(s) : st_re_low(s);
    return fdo_allowed_call(popcll(st_hand<IDX>(s, seat)), m, e, w);
}
// FdoAllowedActions::calculate_allowed_actions (rs-full-doko/src/action/allowed_actions.rs:68-169); Finished → 0.
template <bool IDX = false>
DK_HD uint64_t fdo_state_legal_mask(const dk_state& s) {
    uint32_t phase = st_phase(s), cur = st_cur(s);
    if (phase == DK_PHASE_RESERVATION) {
        uint64_t m = (1ull << 24) | (0x7Full << 26);
        if ((hand_both24(st_hand<IDX>(s, cur)) >> CARD_CQ) & 1u) m |= 1ull << 25;
        return m;
    }
    if (phase == DK_PHASE_ANNOUNCEMENT) {
        uint32_t call = fdo_state_allowed_call<IDX>(s, cur);
        return (1ull << 38) | (call ? (1ull << (32u + call)) : 0ull);
    }
    if (phase == DK_PHASE_PLAY_CARD) {
        uint32_t single = hand_any24(st_hand<IDX>(s, cur));
        uint32_t ci = s.card_index, k = ci & 3u;
        if (k == 0u || st_n_tricks(s) >= 12u) return single;       // state.rs:360-372: no colour in the 12th trick
        uint32_t first = st_quad<IDX>(s, ci >> 2) & 255u;
        uint32_t f = single & follow_mask(first, trump_mask_for_game_type(st_gt(s)));
        return f ? f : single;
    }
    return 0;
}
// internal_progress (announcement.rs:130-175) from seat p; sets phase / current seat / turns.
// The reference asks the seats one after the other (seats with an empty allowed set are skipped and count as a "no") until one may
// call or four consecutive no's end the round.  Here the four seats are decided at once: a seat may call iff its hand holds at least
// the threshold of its team (fdo_min_cards_to_call is exactly "fdo_allowed_call != 0"), one byte per seat as in the playout kernels
// (fdo_eligible_nibble); the first eligible seat among the next 4 - turns is found with one bit scan.  Straight-line code instead of
// up to four rounds of run-time-indexed hand reads — the loop was the largest single item of the transition's instruction count
// (profiles/r02_apply_sorted_ncu_summary.json).
template <bool IDX = false>
DK_HD void fdo_state_progress(dk_state& s, uint32_t p) {
    uint32_t turns = st_turns(s);
    const uint32_t tag = st_team_tag(s);
    uint32_t elig = 0;
    if (tag == TEAM_WEDDING_SOLVED || tag == TEAM_NO_WEDDING) {
        const uint32_t w = tag == TEAM_WEDDING_SOLVED ? st_solved_idx(s) : 0u, re = st_re_mask(s), rl = st_re_low(s), kl = st_ko_low(s);
        const uint32_t thr_re = fdo_min_cards_to_call(rl, kl, w), thr_ko = fdo_min_cards_to_call(kl, rl, w);
        const uint32_t re_spread = fdo_spread4(re);
        const uint32_t thr4 = thr_re * re_spread + thr_ko * (0x01010101u - re_spread);
        const uint32_t cards4h = 0x80808080u + (popcll(s.hands[0]) | (popcll(s.hands[1]) << 8) | (popcll(s.hands[2]) << 16) | (popcll(s.hands[3]) << 24));
        elig = fdo_eligible_nibble(cards4h, thr4);
    }
    const uint32_t left = 4u - (turns < 4u ? turns : 4u);                 // seats still to ask before the round is over
    const uint32_t win = ((elig * 0x11u) >> p) & ((1u << left) - 1u);     // bit d: seat p + d may call and is reached
    if (win == 0u) {                                                      // RoundIsOver(starting_player)
        turns = 4u;
        st_set(s.meta, 0, 2, DK_PHASE_PLAY_CARD);
        st_set(s.meta, 2, 2, st_ann_start(s));
    } else {
        const uint32_t d = ffs0(win);
        turns += d;
        st_set(s.meta, 0, 2, DK_PHASE_ANNOUNCEMENT);
        st_set(s.meta, 2, 2, (p + d) & 3u);
    }
    st_set(s.meta, 26, 3, turns);
}
// progress_next_card_or_announcement (state.rs:184-206) = start_round(current seat)
template <bool IDX = false>
DK_HD void fdo_state_start_round(dk_state& s) {
    st_set(s.meta, 26, 3, 0u);
    st_set(s.meta, 29, 2, st_cur(s));
    fdo_state_progress<IDX>(s, st_cur(s));
}
// Winner of a completed trick from its four cards (`quad`, byte k = k-th card) and lead seat (trick_winning_player_logic.rs:15-45).
DK_HD uint32_t fdo_quad_winner(uint32_t quad, uint32_t lead, uint32_t trump, uint32_t* win_card, uint32_t* eyes, uint32_t* fox_mask) {
    uint32_t c0 = quad & 255u;
    uint32_t follow = follow_mask(c0, trump);
    uint32_t best = 0, bestk = 0, bestc = c0, e = 0, fm = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k) {
        uint32_t c = (quad >> (8u * k)) & 255u;
        uint32_t pw = card_power(c, trump, follow);
        if (k == 0u || pw > best) { best = pw; bestk = k; bestc = c; }
        e += card_eyes_by_rank(c - 6u * card_suit(c));
        if (c == CARD_DA) fm |= 1u << k;
    }
    if (win_card) *win_card = bestc;
    if (eyes) *eyes = e;
    if (fox_mask) *fox_mask = fm;
    return (lead + bestk) & 3u;
}
template <bool IDX = false>
DK_HD uint32_t fdo_state_trick_winner(const dk_state& s, uint32_t t, uint32_t trump, uint32_t* win_card, uint32_t* eyes, uint32_t* fox_mask) {
    return fdo_quad_winner(st_quad<IDX>(s, t), st_trick_start(s, t), trump, win_card, eyes, fox_mask);
}
// FdoEndOfGameStats::calculate (stats/stats.rs:46-240) from the stored history.
DK_HD void fdo_state_final_points(dk_state& s, uint32_t last_winner, uint32_t last_win_card) {
    uint32_t re = st_re_mask(s), trump = trump_mask_for_game_type(st_gt(s));
    uint32_t re_eyes = 0, re_tricks = 0;
#pragma unroll
    for (uint32_t p = 0; p < 4u; ++p)
        if ((re >> p) & 1u) { re_eyes += s.eyes[p]; re_tricks += (s.num_tricks >> (4u * p)) & 15u; }
    int32_t extras = 0;
    if (popc(re) != 1u) {
#pragma unroll
        for (uint32_t t = 0; t < 12u; ++t) {
            // the winner of trick t leads trick t + 1, so only the eyes and the ♦A positions have to be re-read from the cards
            uint32_t e = 0, fm = 0;
            const uint32_t w = t < 11u ? st_trick_start(s, t + 1u) : last_winner, quad = st_word_of(s.cards, t);
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) {
                const uint32_t c = (quad >> (8u * k)) & 255u;
                e += card_eyes_by_rank(c - 6u * card_suit(c));
                fm |= (c == CARD_DA ? 1u : 0u) << k;
            }
            (void)trump;
            bool won_re = (re >> w) & 1u;
            if (e >= 40u) extras += won_re ? 1 : -1;                                  // doppelkopf.rs:8-26
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k)
                if ((fm >> k) & 1u) {                                                  // fuchs_gefangen.rs:9-60
                    bool played_re = (re >> ((st_trick_start(s, t) + k) & 3u)) & 1u;
                    if (played_re != won_re) extras += won_re ? 1 : -1;
                }
        }
        if (last_win_card == CARD_CJ) extras += ((re >> last_winner) & 1u) ? 1 : -1;  // last_trick_karlchen.rs:6-27
    }
    int32_t ko;
    int32_t rp = fdo_score(re_eyes, re_tricks, popc(re), st_re_low(s), st_ko_low(s), extras, &ko);
#pragma unroll
    for (uint32_t p = 0; p < 4u; ++p) s.points[p] = (int8_t)(((re >> p) & 1u) ? rp : ko);
}
// FdoState::play_action (rs-full-doko/src/state/state.rs:208-358).  Returns 0, or 1 if the action is not legal
// (the reference would panic); the state is unchanged in that case.
template <bool IDX = false>
DK_HD uint32_t fdo_state_apply(dk_state& s, uint32_t action) {
    if (action >= 39u || !((fdo_state_legal_mask<IDX>(s) >> action) & 1ull)) return 1;
    uint32_t phase = st_phase(s), cur = st_cur(s);
    if (phase == DK_PHASE_RESERVATION) {
        st_push_reservation<IDX>(s, fdo_res_code_from_action(action));
        uint32_t next = (cur + 1u) & 3u;
        st_set(s.meta, 2, 2, next);
        if (s.n_reservations == 4u) {
            uint32_t start = st_game_start(s);
            uint32_t solo_i = 4, wed_i = 4, solo_code = 0;
#pragma unroll
            for (uint32_t i = 0; i < 4u; ++i) {
                uint32_t r = s.reservations[i];
                if (r >= 2u && solo_i == 4u) { solo_i = i; solo_code = r; }
                if (r == 1u) wed_i = i;
            }
            uint32_t gt, tag, re = 0, wed = 0;
            if (solo_i < 4u) { gt = fdo_gt_from_res_code(solo_code); tag = TEAM_NO_WEDDING; re = 1u << ((start + solo_i) & 3u); }
            else if (wed_i < 4u) { gt = GT_WEDDING; tag = TEAM_WEDDING_UNSOLVED; wed = (start + wed_i) & 3u; }
            else {
                gt = GT_NORMAL; tag = TEAM_NO_WEDDING;
#pragma unroll
                for (uint32_t p = 0; p < 4u; ++p) if ((hand_any24(s.hands[p]) >> CARD_CQ) & 1u) re |= 1u << p;
            }
            st_set(s.meta, 6, 4, gt); st_set(s.meta, 10, 2, tag); st_set(s.meta, 12, 2, wed); st_set(s.meta, 16, 4, re);
            s.tricks = (s.tricks & 0xF0000000u) | next | (1u << 24);        // trick 0 led by the start seat
            fdo_state_start_round<IDX>(s);
        }
        return 0;
    }
    if (phase == DK_PHASE_ANNOUNCEMENT) {
        uint32_t turns = st_turns(s);
        if (action == 38u) turns++;
        else {
            turns = 0;
            uint32_t level = action - 32u;                                   // 33 → ReContra even for a counter (action.rs:297-299)
            uint32_t n = st_n_calls(s);
            if (n < 12u) { st_set_call<IDX>(s, n, s.card_index | (cur << 6) | (level << 8)); st_set(s.tricks, 28, 4, n + 1u); }
            if ((st_re_mask(s) >> cur) & 1u) st_set(s.meta, 20, 3, level); else st_set(s.meta, 23, 3, level);
        }
        st_set(s.meta, 26, 3, turns);
        fdo_state_progress<IDX>(s, (cur + 1u) & 3u);
        return 0;
    }
    // card
    uint32_t c = action;
    uint64_t h = st_hand<IDX>(s, cur);
    if ((h >> (c + 24u)) & 1ull) h &= ~(1ull << (c + 24u)); else h &= ~(1ull << c);   // hand.remove: copy B first
    st_set_hand<IDX>(s, cur, h);
    uint32_t ci = s.card_index;
    st_set_card<IDX>(s, ci, c);
    s.card_index = (uint8_t)(ci + 1u);
    if (((ci + 1u) & 3u) == 0u) {
        uint32_t t = ci >> 2, trump = trump_mask_for_game_type(st_gt(s)), wc, e;
        uint32_t w = fdo_state_trick_winner<IDX>(s, t, trump, &wc, &e, nullptr);
        st_add_eyes<IDX>(s, w, e);
        s.num_tricks = (uint16_t)(s.num_tricks + (1u << (4u * w)));
        if (st_team_tag(s) == TEAM_WEDDING_UNSOLVED) {                       // team_logic.rs:59-112
            uint32_t wed = st_wed_seat(s);
            if (w != wed) { st_set(s.meta, 10, 2, TEAM_WEDDING_SOLVED); st_set(s.meta, 14, 2, t); st_set(s.meta, 16, 4, (1u << wed) | (1u << w)); }
            else if (t == 2u) { st_set(s.meta, 10, 2, TEAM_WEDDING_SOLVED); st_set(s.meta, 14, 2, 2u); st_set(s.meta, 16, 4, 1u << wed); }
        }
        if (t == 11u) {
            st_set(s.meta, 0, 2, DK_PHASE_FINISHED);
            st_set(s.meta, 2, 2, 0u);
            fdo_state_final_points(s, w, wc);
            return 0;
        }
        st_set(s.meta, 2, 2, w);
        st_set(s.tricks, 2u * (t + 1u), 2, w);
        st_set(s.tricks, 24, 4, t + 2u);
        fdo_state_start_round<IDX>(s);
        return 0;
    }
    st_set(s.meta, 2, 2, (cur + 1u) & 3u);
    fdo_state_start_round<IDX>(s);
    return 0;
}
// FdoAzEnvState::take_action_by_action_index(action, skip_single, _) (rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:121-154):
// after the action, keep playing while exactly one non-call action is legal.
template <bool IDX = false>
DK_HD uint32_t fdo_state_apply_az(dk_state& s, uint32_t action, bool skip_single) {
    uint32_t err = fdo_state_apply<IDX>(s, action);
    if (err || !skip_single) return err;
    const uint64_t calls = 0x1Full << 33;
    for (;;) {
        if (st_phase(s) == DK_PHASE_FINISHED) break;
        uint64_t m = fdo_state_legal_mask<IDX>(s) & ~calls;
        if (popcll(m) != 1u) break;
        fdo_state_apply<IDX>(s, ffs0ll(m));
    }
    return 0;
}

// The chained card draws of a trick (draw_chain): how many card types the k-th play of the running trick chose from — the seat's hand
// (`hand_any`: its card types NOW) with the played card `c` back in it, restricted to the led colour (`follow`, k > 0) when it can
// follow.  A playout that resumes inside a trick continues the chain at word * product of these counts (FdoResume::chain_mul).
DK_HD uint32_t chain_count(uint32_t hand_any, uint32_t c, uint32_t k, uint32_t follow) {
    const uint32_t had = hand_any | (1u << c), f = k ? had & follow : 0u;
    return popc(f ? f : had);
}

// Bridge: stored state → register-resident playout form.  Returns false when the game is already finished.
template <bool IDX = false>
DK_HD bool fdo_state_to_live(const dk_state& s, FdoLive& g, FdoResume& rs) {
    fdo_live_clear(g);
    uint32_t phase = st_phase(s);
    if (phase == DK_PHASE_FINISHED) return false;
    uint32_t dup = 0;
#pragma unroll
    for (uint32_t p = 0; p < 4u; ++p) dup |= hand_both24(s.hands[p]);
    g.dup = dup;
    rs.n_res = s.n_reservations;
    rs.t0 = 0; rs.k0 = 0; rs.starts = 0; rs.ann_ci = 0; rs.ann_p = 0; rs.ann_turns = 0xFFFFFFFFu; rs.chain_mul = 1u;
    trick_acc_clear(rs.acc);
#pragma unroll
    for (uint32_t i = 0; i < 4u; ++i) rs.res_action[i] = i < s.n_reservations ? fdo_action_from_res_code(s.reservations[i]) : 0u;
    uint32_t base;
    if (phase == DK_PHASE_RESERVATION) {
        base = st_game_start(s);
    } else {
        uint32_t ci = s.card_index, t0 = ci >> 2, k0 = ci & 3u;
        g.gt = st_gt(s); g.trump = trump_mask_for_game_type(g.gt);
        g.team_tag = st_team_tag(s); g.re_mask = st_re_mask(s); g.wed_seat = st_wed_seat(s); g.solved_idx = st_solved_idx(s);
        g.re_low = st_re_low(s); g.ko_low = st_ko_low(s);
#pragma unroll
        for (uint32_t p = 0; p < 4u; ++p) { g.eyes |= (uint32_t)s.eyes[p] << (8u * p); }
        g.ntricks = s.num_tricks;
        // trackers of the completed tricks: the winner of trick t leads trick t + 1, so only the eyes and the ♦A positions have to be
        // re-read from the cards
        auto track = [&](uint32_t t, uint32_t quad) {
            uint32_t e = 0, fm = 0, w = st_trick_start(s, t + 1u);
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) {
                const uint32_t c = (quad >> (8u * k)) & 255u;
                e += card_eyes_by_rank(c - 6u * card_suit(c));
                fm |= (c == CARD_DA ? 1u : 0u) << k;
            }
            if (e >= 40u) g.dkc += 1u << (4u * w);
            g.foxes = fdo_fox_record(g.foxes, fm, st_trick_start(s, t), w);
            rs.starts |= st_trick_start(s, t) << (2u * t);
            g.last_winner = w;
        };
        if (IDX) {
#pragma unroll 1
            for (uint32_t t = 0; t < t0; ++t) track(t, st_quad<true>(s, t));
        } else {
#pragma unroll
            for (uint32_t t = 0; t < 12u; ++t) {                             // compile-time offsets
                if (t >= t0) break;
                track(t, st_word_of(s.cards, t));
            }
        }
        base = st_trick_start(s, t0);
        rs.t0 = t0; rs.k0 = k0;
        const uint32_t quad0 = k0 ? st_quad<IDX>(s, t0) : 0u;
#pragma unroll
        for (uint32_t k = 0; k < 3u; ++k) {                                  // partial trick
            if (k >= k0) break;
            uint32_t c = (quad0 >> (8u * k)) & 255u;
            if (k == 0u) { rs.acc.follow = follow_mask(c, g.trump); rs.acc.prow = pow_row(g.gt, c, card_suit(c), g.trump); }
            const uint32_t v = pow_entry_of(c, g.trump, rs.acc.follow), cand = v | ((3u - k) << POW_K_SHIFT);
            if (k == 0u || cand > rs.acc.best) rs.acc.best = cand;
            rs.acc.acc += v;
            rs.acc.fox += (v & POW_FOX_BIT) << k;
            rs.chain_mul *= chain_count(hand_any24(st_hand<IDX>(s, (base + k) & 3u)), c, k, rs.acc.follow);
        }
        if (phase == DK_PHASE_ANNOUNCEMENT) { rs.ann_ci = ci; rs.ann_p = st_cur(s); rs.ann_turns = st_turns(s); }
        else { rs.ann_ci = ci + 1u; rs.ann_turns = 0xFFFFFFFFu; }            // the round before card ci is over
    }
    g.base = 0;
    g.h0 = hand_any24(s.hands[0]); g.h1 = hand_any24(s.hands[1]); g.h2 = hand_any24(s.hands[2]); g.h3 = hand_any24(s.hands[3]);
    fdo_rotate(g, base);                                                     // frame index 0 = seat `base` (selects, no indexed array)
    return true;
}

// clone_with_different_hands_and_reservations (state.rs:96-119) on the PLAYOUT form: (g, rs) is the bridge of the info-state `leaf`
// (fdo_state_to_live, built once per leaf), `hands` a card_matching sample by absolute seat, `res4` its reservations (one byte per
// absolute seat).  Only what the sample changes is replaced: the four hands, the doubled-card mask, in the reservation phase the
// hidden reservations of the seats that have declared, and — inside a trick — the product that positions the trick's chained card
// draws (chain_count over the SAMPLED hands: the oracle positions its stream on the determinized state).
DK_HD void fdo_live_with_sample(FdoLive& g, FdoResume& rs, const dk_state& leaf, const uint64_t hands[4], uint32_t res4) {
    const uint32_t base = g.base;                                                 // frame seat k = absolute seat base + k
    const uint32_t any[4] = {hand_any24(hands[0]), hand_any24(hands[1]), hand_any24(hands[2]), hand_any24(hands[3])};
    g.dup = hand_both24(hands[0]) | hand_both24(hands[1]) | hand_both24(hands[2]) | hand_both24(hands[3]);
    g.h0 = any[base & 3u]; g.h1 = any[(base + 1u) & 3u]; g.h2 = any[(base + 2u) & 3u]; g.h3 = any[(base + 3u) & 3u];
    if (rs.n_res < 4u) {                                                          // reservation phase: frame base = game start seat
#pragma unroll
        for (uint32_t k = 0; k < 3u; ++k) if (k < rs.n_res) rs.res_action[k] = fdo_action_from_res_code((res4 >> (8u * ((base + k) & 3u))) & 255u);
    }
    if (rs.k0 > 0u) {
        const uint32_t quad0 = st_quad<true>(leaf, rs.t0);
        uint32_t m = 1u;
#pragma unroll
        for (uint32_t k = 0; k < 3u; ++k) if (k < rs.k0) m *= chain_count(any[(base + k) & 3u], (quad0 >> (8u * k)) & 255u, k, rs.acc.follow);
        rs.chain_mul = m;
    }
}

// =====================================================================================================================
// rs-doko
// =====================================================================================================================
// calculate_allowed_actions_in_normal_game (rs-doko/src/action/allowed_actions.rs:131-198)
template <bool IDX = false>
DK_HD uint64_t doko_state_legal_mask(const dk_state& s) {
    uint32_t phase = st_phase(s), cur = st_cur(s);
    if (phase == DK_PHASE_RESERVATION) return (1ull << 24) | (((hand_both24(st_hand<IDX>(s, cur)) >> CARD_CQ) & 1u) ? (1ull << 25) : 0ull);
    if (phase == DK_PHASE_PLAY_CARD) {
        uint32_t single = hand_any24(st_hand<IDX>(s, cur));
        uint32_t ci = s.card_index, k = ci & 3u;
        if (k == 0u) return single;
        uint32_t f = single & follow_mask(st_quad<IDX>(s, ci >> 2) & 255u, DOKO_TRUMP_MASK);
        return f ? f : single;
    }
    return 0;
}
template <bool IDX = false>
DK_HD uint32_t doko_state_trick_winner(const dk_state& s, uint32_t t, uint32_t* eyes) {
    const uint32_t quad = st_quad<IDX>(s, t);
    uint32_t follow = follow_mask(quad & 255u, DOKO_TRUMP_MASK);
    uint32_t best = 0, bestk = 0, e = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k) {
        uint32_t c = (quad >> (8u * k)) & 255u;
        uint32_t pw = card_power(c, DOKO_TRUMP_MASK, follow);
        if (k == 0u || pw > best) { best = pw; bestk = k; }
        e += card_eyes_by_rank(c - 6u * card_suit(c));
    }
    *eyes = e;
    return (st_trick_start(s, t) + bestk) & 3u;
}
// DoState::play_action (rs-doko/src/state/state.rs:189-309)
template <bool IDX = false>
DK_HD uint32_t doko_state_apply(dk_state& s, uint32_t action) {
    if (action >= 26u || !((doko_state_legal_mask<IDX>(s) >> action) & 1ull)) return 1;
    uint32_t phase = st_phase(s), cur = st_cur(s);
    if (phase == DK_PHASE_RESERVATION) {
        st_push_reservation<IDX>(s, action == 25u ? 0u : 1u);                           // DoReservation: Wedding 0, Healthy 1
        uint32_t next = (cur + 1u) & 3u;
        st_set(s.meta, 2, 2, next);
        if (s.n_reservations == 4u) {
            uint32_t start = st_game_start(s), wed_i = 4;
#pragma unroll
            for (uint32_t i = 0; i < 4u; ++i) if (s.reservations[i] == 0u) wed_i = i;
            uint32_t re = 0;
            if (wed_i < 4u) { st_set(s.meta, 6, 4, GT_WEDDING); st_set(s.meta, 10, 2, TEAM_WEDDING_UNSOLVED); st_set(s.meta, 12, 2, (start + wed_i) & 3u); }
            else {
#pragma unroll
                for (uint32_t p = 0; p < 4u; ++p) if ((hand_any24(s.hands[p]) >> CARD_CQ) & 1u) re |= 1u << p;
                st_set(s.meta, 6, 4, GT_NORMAL); st_set(s.meta, 10, 2, TEAM_NO_WEDDING); st_set(s.meta, 16, 4, re);
            }
            s.tricks = next | (1u << 24);
            st_set(s.meta, 0, 2, DK_PHASE_PLAY_CARD);
        }
        return 0;
    }
    uint32_t c = action;
    uint64_t h = st_hand<IDX>(s, cur);
    if ((h >> c) & 1ull) h &= ~(1ull << c); else h &= ~(1ull << (c + 24u));       // hand_remove: copy A first (rs-doko/src/hand/hand.rs:38-46)
    st_set_hand<IDX>(s, cur, h);
    uint32_t ci = s.card_index;
    st_set_card<IDX>(s, ci, c);
    s.card_index = (uint8_t)(ci + 1u);
    if (((ci + 1u) & 3u) == 0u) {
        uint32_t t = ci >> 2, e;
        uint32_t w = doko_state_trick_winner<IDX>(s, t, &e);
        st_add_eyes<IDX>(s, w, e);
        s.num_tricks = (uint16_t)(s.num_tricks + (1u << (4u * w)));
        if (st_team_tag(s) == TEAM_WEDDING_UNSOLVED) {
            uint32_t wed = st_wed_seat(s);
            if (w != wed) { st_set(s.meta, 10, 2, TEAM_WEDDING_SOLVED); st_set(s.meta, 14, 2, t); st_set(s.meta, 16, 4, (1u << wed) | (1u << w)); }
            else if (t == 2u) { st_set(s.meta, 10, 2, TEAM_WEDDING_SOLVED); st_set(s.meta, 14, 2, 2u); st_set(s.meta, 16, 4, 1u << wed); }
        }
        if (t == 11u) {
            st_set(s.meta, 0, 2, DK_PHASE_FINISHED);
            st_set(s.meta, 2, 2, 0u);
            DokoLive g; doko_live_clear(g);
            g.re_mask = st_re_mask(s); g.ntricks = s.num_tricks;
#pragma unroll
            for (uint32_t p = 0; p < 4u; ++p) g.eyes |= (uint32_t)s.eyes[p] << (8u * p);
            int32_t pts[4]; doko_final_points(g, pts);
#pragma unroll
            for (uint32_t p = 0; p < 4u; ++p) s.points[p] = (int8_t)pts[p];
            return 0;
        }
        st_set(s.meta, 2, 2, w);
        st_set(s.tricks, 2u * (t + 1u), 2, w);
        st_set(s.tricks, 24, 4, t + 2u);
        return 0;
    }
    st_set(s.meta, 2, 2, (cur + 1u) & 3u);
    return 0;
}
template <bool IDX = false>
DK_HD bool doko_state_to_live(const dk_state& s, DokoLive& g, DokoResume& rs) {
    doko_live_clear(g);
    uint32_t phase = st_phase(s);
    if (phase == DK_PHASE_FINISHED) return false;
    uint32_t dup = 0;
#pragma unroll
    for (uint32_t p = 0; p < 4u; ++p) dup |= hand_both24(s.hands[p]);
    g.dup = dup;
    rs.n_res = s.n_reservations; rs.t0 = 0; rs.k0 = 0; rs.chain_mul = 1u;
    doko_trick_acc_clear(rs.acc);
#pragma unroll
    for (uint32_t i = 0; i < 4u; ++i) rs.res_action[i] = i < s.n_reservations ? (s.reservations[i] == 0u ? 25u : 24u) : 0u;
    uint32_t base;
    if (phase == DK_PHASE_RESERVATION) base = st_game_start(s);
    else {
        uint32_t ci = s.card_index, t0 = ci >> 2, k0 = ci & 3u;
        g.team_tag = st_team_tag(s); g.re_mask = st_re_mask(s); g.wed_seat = st_wed_seat(s); g.solved_idx = st_solved_idx(s);
        g.wedding = st_gt(s) == GT_WEDDING ? 1u : 0u;
#pragma unroll
        for (uint32_t p = 0; p < 4u; ++p) g.eyes |= (uint32_t)s.eyes[p] << (8u * p);
        g.ntricks = s.num_tricks;
        base = st_trick_start(s, t0);
        rs.t0 = t0; rs.k0 = k0;
        const uint32_t quad0 = k0 ? st_quad<IDX>(s, t0) : 0u;
#pragma unroll
        for (uint32_t k = 0; k < 3u; ++k) {
            if (k >= k0) break;
            uint32_t c = (quad0 >> (8u * k)) & 255u;
            if (k == 0u) { rs.acc.follow = follow_mask(c, DOKO_TRUMP_MASK); rs.acc.prow = pow_row(0u, c, card_suit(c), DOKO_TRUMP_MASK); }
            const uint32_t v = pow_entry_of(c, DOKO_TRUMP_MASK, rs.acc.follow), cand = v | ((3u - k) << POW_K_SHIFT);
            if (k == 0u || cand > rs.acc.best) rs.acc.best = cand;
            rs.acc.acc += v;
            rs.chain_mul *= chain_count(hand_any24(st_hand<IDX>(s, (base + k) & 3u)), c, k, rs.acc.follow);
        }
    }
    g.base = 0;
    g.h0 = hand_any24(s.hands[0]); g.h1 = hand_any24(s.hands[1]); g.h2 = hand_any24(s.hands[2]); g.h3 = hand_any24(s.hands[3]);
    doko_rotate(g, base);
    return true;
}

}  // namespace dk
