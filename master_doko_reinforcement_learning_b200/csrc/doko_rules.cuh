// doko_rules.cuh — the simplified engine (rs-doko: normal game + wedding, no solos, no announcements, simple scoring)
// as a register-resident thread-per-game program.  Same rotating-frame layout as fdo_rules.cuh; 52 lock-step steps
// (4 reservations + 48 cards) with no divergent control flow at all.
#pragma once
#include "dk_common.cuh"
#include "fdo_rules.cuh"

namespace dk {

constexpr uint32_t DOKO_TRUMP_MASK = 0x30C3BFu;  // rs-doko/src/card/card_color_masks.rs:3-15 (pinned at :36-41)

struct DokoLive {
    uint32_t h0, h1, h2, h3;   // frame-relative hands ("at least one copy")
    uint32_t dup;              // card types with both copies in one hand
    uint32_t base;             // absolute seat of frame index 0
    uint32_t eyes, ntricks;    // 8 / 4 bits per absolute seat
    uint32_t team_tag, re_mask, wed_seat, solved_idx;
    uint32_t wedding;          // reservation result: 1 = Wedding(wed_seat)
    uint32_t steps;
};

DK_HD void doko_rotate(DokoLive& g, uint32_t r) {
    uint32_t a0 = g.h0, a1 = g.h1, a2 = g.h2, a3 = g.h3;
    if (r & 1u) { uint32_t t = a0; a0 = a1; a1 = a2; a2 = a3; a3 = t; }
    if (r & 2u) { uint32_t t0 = a0, t1 = a1; a0 = a2; a1 = a3; a2 = t0; a3 = t1; }
    g.h0 = a0; g.h1 = a1; g.h2 = a2; g.h3 = a3;
    g.base = (g.base + r) & 3u;
}

// Reservation pick (rs-doko/src/action/allowed_actions.rs:139-151): {Healthy 24, Wedding 25 iff both ♣Q};
// MSB-first ⇒ rank 0 = Wedding when present.  The draw is consumed even for a single legal action.
DK_HD uint32_t doko_pick_reservation(uint32_t h, uint32_t dup, uint32_t word) {
    uint32_t has_wedding = ((h & dup) >> CARD_CQ) & 1u;
    uint32_t idx = mulhi(word, 1u + has_wedding);
    return (has_wedding && idx == 0u) ? 25u : 24u;
}

// Reservation round outcome (rs-doko/src/reservation/reservation_winning_logic.rs:13-38: the LAST wedding wins;
// teams/team_logic.rs:69-143).
DK_HD void doko_finish_reservations(DokoLive& g, const uint32_t res_action[4]) {
    uint32_t wed_i = 4;
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) if (res_action[i] == 25u) wed_i = i;
    if (wed_i < 4u) {
        g.wedding = 1; g.team_tag = TEAM_WEDDING_UNSOLVED; g.wed_seat = (g.base + wed_i) & 3u; g.re_mask = 0;
    } else {
        g.wedding = 0; g.team_tag = TEAM_NO_WEDDING;
        uint32_t m = 0;
        m |= ((g.h0 >> CARD_CQ) & 1u) << (g.base & 3u);
        m |= ((g.h1 >> CARD_CQ) & 1u) << ((g.base + 1u) & 3u);
        m |= ((g.h2 >> CARD_CQ) & 1u) << ((g.base + 2u) & 3u);
        m |= ((g.h3 >> CARD_CQ) & 1u) << ((g.base + 3u) & 3u);
        g.re_mask = m;
    }
}

struct DokoTrickAcc { uint32_t follow, best, acc, prow; };   // best / acc: dk_common.cuh pow_lut_entry; prow: row of that table (game type 0 = the rs-doko trumps)
DK_HD void doko_trick_acc_clear(DokoTrickAcc& a) { a.follow = 0; a.best = 0; a.acc = 0; a.prow = 0; }

// Card step of frame seat K (rs-doko/src/action/allowed_actions.rs:153-192, state/state.rs:194-252).
template <int K, bool SEL12 = false>
DK_HD uint32_t doko_card_step(DokoLive& g, uint32_t& h, DokoTrickAcc& a, uint32_t& word, const uint32_t* __restrict__ lut) {
    uint32_t mask = h;
    if (K > 0) { uint32_t f = h & a.follow; mask = f ? f : h; }
    uint32_t idx = draw_chain(word, popc(mask));         // the trick's word serves its four draws in a row (dk_common.cuh)
    uint32_t c = SEL12 ? pick_msb_rank24_tab(mask, idx, reinterpret_cast<const uint64_t*>(lut + SEL12_LUT_BASE)) : pick_msb_rank24_lut(mask, idx, lut);
    uint32_t bit = 1u << c;
    uint32_t dbl = g.dup & bit;
    g.dup ^= dbl;
    h ^= bit ^ dbl;
    if (K == 0) { const uint32_t e = lead_lookup(lut, 0u, c); a.follow = lead_follow(e); a.prow = lead_row(e); }   // (game type 0 = the rs-doko trumps)
    const uint32_t v = pow_lookup(lut, a.prow, c), cand = v | ((3u - (uint32_t)K) << POW_K_SHIFT);
    a.best = (K == 0 || cand > a.best) ? cand : a.best;
    a.acc += v;
    g.steps++;
    return c;
}

// The twelfth trick played from its first card: every seat holds one card (see fdo_card_step_last).
template <int K>
DK_HD uint32_t doko_card_step_last(DokoLive& g, uint32_t& h, DokoTrickAcc& a, const uint32_t* __restrict__ lut) {
    const uint32_t c = ffs0(h);
    h = 0u;
    if (K == 0) a.prow = lead_row(lead_lookup(lut, 0u, c));
    const uint32_t v = pow_lookup(lut, a.prow, c), cand = v | ((3u - (uint32_t)K) << POW_K_SHIFT);
    a.best = (K == 0 || cand > a.best) ? cand : a.best;
    a.acc += v;
    g.steps++;
    return c;
}

DK_HD void doko_trick_done(DokoLive& g, const DokoTrickAcc& a, uint32_t t) {
    const uint32_t bestk = pow_best_k(a.best);
    uint32_t w = (g.base + bestk) & 3u;
    g.eyes += (a.acc & 255u) << (8u * w);
    g.ntricks += 1u << (4u * w);
    if (g.team_tag == TEAM_WEDDING_UNSOLVED) {
        if (w != g.wed_seat) { g.team_tag = TEAM_WEDDING_SOLVED; g.solved_idx = t; g.re_mask = (1u << g.wed_seat) | (1u << w); }
        else if (t == 2u) { g.team_tag = TEAM_WEDDING_SOLVED; g.solved_idx = 2u; g.re_mask = 1u << g.wed_seat; }
    }
    doko_rotate(g, bestk);
}

// calculate_end_of_game_stats (rs-doko/src/stats/stats.rs:25-135): Re wins iff re_eyes > kontra_eyes (120:120 → Kontra);
// 1 point each for the winner reaching 120/150/180/210 eyes and for all 12 tricks; a lone Re player scores x3.
DK_HD void doko_final_points(const DokoLive& g, int32_t pts[4]) {
    uint32_t re_eyes = 0, re_tricks = 0;
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s)
        if ((g.re_mask >> s) & 1u) { re_eyes += (g.eyes >> (8u * s)) & 255u; re_tricks += (g.ntricks >> (4u * s)) & 15u; }
    uint32_t ko_eyes = 240u - re_eyes;
    bool re_wins = re_eyes > ko_eyes;
    uint32_t we = re_wins ? re_eyes : ko_eyes, wt = re_wins ? re_tricks : 12u - re_tricks;
    int32_t p = (int32_t)((we >= 120u) + (we >= 150u) + (we >= 180u) + (we >= 210u) + (wt == 12u));
    int32_t re = re_wins ? p : -p, ko = -re;
    if (popc(g.re_mask) == 1u) re *= 3;
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) pts[s] = ((g.re_mask >> s) & 1u) ? re : ko;
}

struct DokoResume {
    uint32_t n_res;
    uint32_t res_action[4];
    uint32_t t0, k0;
    DokoTrickAcc acc;
    uint32_t chain_mul;    // product of the legal-card counts of the k0 plays already made in trick t0 (see FdoResume)
};

// trace (optional, FRESH only): 52 action ids in play order.
template <bool FRESH, bool TRACE, bool SEL12 = false>
DK_HD void doko_play_to_end(DokoLive& g, const RngKey& key, const DokoResume* rs, uint8_t* trace, const uint32_t* __restrict__ lut) {
    uint32_t n_res = FRESH ? 0u : rs->n_res;
    if (n_res < 4u) {
        U4 blk = rng_block(key, SITE_RESERVATION, 0);
        uint32_t ra[4];
        ra[0] = (!FRESH && n_res > 0u) ? rs->res_action[0] : doko_pick_reservation(g.h0, g.dup, blk.x);
        ra[1] = (!FRESH && n_res > 1u) ? rs->res_action[1] : doko_pick_reservation(g.h1, g.dup, blk.y);
        ra[2] = (!FRESH && n_res > 2u) ? rs->res_action[2] : doko_pick_reservation(g.h2, g.dup, blk.z);
        ra[3] = doko_pick_reservation(g.h3, g.dup, blk.w);
        g.steps += 4u - n_res;
        if (TRACE) { trace[0] = (uint8_t)ra[0]; trace[1] = (uint8_t)ra[1]; trace[2] = (uint8_t)ra[2]; trace[3] = (uint8_t)ra[3]; }
        doko_finish_reservations(g, ra);
    }
    uint32_t t = FRESH ? 0u : rs->t0;
    // (same shape as fdo_play_to_end: a resumed game first finishes its partial trick; the loop body is unconditional; the twelfth
    // trick, when it starts from its first card, is four forced moves and needs neither its Philox block nor a rank select)
    // (card draws: trick t takes word t & 3 of block t >> 2 of SITE_CARD, chained over the trick's four plays — see fdo_play_to_end)
    U4 cblk;
    cblk.x = cblk.y = cblk.z = cblk.w = 0u;
    if (!FRESH && t < 12u) {
        cblk = rng_block(key, SITE_CARD, t >> 2);
        uint32_t word = u4_word(cblk, t & 3u) * rs->chain_mul;
        DokoTrickAcc a;
        doko_trick_acc_clear(a);
        const uint32_t k0 = rs->k0;
        if (k0 > 0u) a = rs->acc;
        if (k0 <= 0u) doko_card_step<0, SEL12>(g, g.h0, a, word, lut);
        if (k0 <= 1u) doko_card_step<1, SEL12>(g, g.h1, a, word, lut);
        if (k0 <= 2u) doko_card_step<2, SEL12>(g, g.h2, a, word, lut);
        doko_card_step<3, SEL12>(g, g.h3, a, word, lut);
        doko_trick_done(g, a, t);
        ++t;
    }
    for (; t < 11u; ++t) {
        if ((t & 3u) == 0u) cblk = rng_block(key, SITE_CARD, t >> 2);
        uint32_t word = u4_word(cblk, t & 3u);
        DokoTrickAcc a;
        doko_trick_acc_clear(a);
        const uint32_t c0 = doko_card_step<0, SEL12>(g, g.h0, a, word, lut);
        const uint32_t c1 = doko_card_step<1, SEL12>(g, g.h1, a, word, lut);
        const uint32_t c2 = doko_card_step<2, SEL12>(g, g.h2, a, word, lut);
        const uint32_t c3 = doko_card_step<3, SEL12>(g, g.h3, a, word, lut);
        if (TRACE) { trace[4 + 4 * t] = (uint8_t)c0; trace[5 + 4 * t] = (uint8_t)c1; trace[6 + 4 * t] = (uint8_t)c2; trace[7 + 4 * t] = (uint8_t)c3; }
        doko_trick_done(g, a, t);
    }
    if (t == 11u) {
        DokoTrickAcc a;
        doko_trick_acc_clear(a);
        const uint32_t c0 = doko_card_step_last<0>(g, g.h0, a, lut);
        const uint32_t c1 = doko_card_step_last<1>(g, g.h1, a, lut);
        const uint32_t c2 = doko_card_step_last<2>(g, g.h2, a, lut);
        const uint32_t c3 = doko_card_step_last<3>(g, g.h3, a, lut);
        if (TRACE) { trace[48] = (uint8_t)c0; trace[49] = (uint8_t)c1; trace[50] = (uint8_t)c2; trace[51] = (uint8_t)c3; }
        doko_trick_done(g, a, 11u);
    }
}

DK_HD void doko_live_clear(DokoLive& g) {
    g.h0 = g.h1 = g.h2 = g.h3 = g.dup = 0; g.base = 0; g.eyes = g.ntricks = 0; g.team_tag = TEAM_IN_RESERVATIONS;
    g.re_mask = 0; g.wed_seat = 0; g.solved_idx = 0; g.wedding = 0; g.steps = 0;
}

template <bool TRACE, class Deck, bool SEL12 = false, class Ready = NoWait>
DK_HD void doko_playout_fresh(const RngKey& key, Deck& deck, const uint32_t* __restrict__ lut, int32_t pts[4], uint32_t& steps, uint8_t* trace, uint32_t* aux,
                              Ready tables_ready = Ready()) {
    DokoLive g;
    doko_live_clear(g);
    FdoLive dummy; dummy.base = 0;
    uint32_t ah[4], start;
    fdo_deal(dummy, key, deck, ah, g.dup, start);        // same deal contract as the full engine
    tables_ready();
    g.h0 = ah[0]; g.h1 = ah[1]; g.h2 = ah[2]; g.h3 = ah[3];
    doko_rotate(g, start);
    doko_play_to_end<true, TRACE, SEL12>(g, key, nullptr, trace, lut);
    doko_final_points(g, pts);
    steps = g.steps;
    if (TRACE && aux) { aux[0] = g.wedding ? 1u : 0u; aux[1] = g.re_mask; aux[2] = g.eyes; aux[3] = g.ntricks; }
}

}  // namespace dk
