// assignment.cuh — hidden-hand sampler of the simplified engine (rs-doko-assignment/src/assignment.rs:13-581) as a
// thread-per-sample program on count planes instead of Vecs.
//
// The reference keeps ordered Vec<DoCard> multisets and uses position()/remove(); what the ORDER influences is (a) which
// single-owner card is found first (ascending card id, the hand_to_vec order of `remaining`) and (b) which element a random
// index denotes (remaining cards in ascending id with doubles adjacent; eligible seats ascending).  Both are reproduced on
// two 24-bit planes per multiset (a = count >= 1, b = count == 2).
#pragma once
#include "matching.cuh"

namespace dk {

struct AssignPrep {
    Hand2 remaining;
    Hand2 allowed[4];      // by ABSOLUTE seat (the observer's list is empty)
    uint32_t len[4];
    uint32_t observer;
    uint32_t obs_a, obs_b;
    uint32_t valid;
};

DK_HD uint32_t h2_count(const Hand2& h) { return popc(h.a) + popc(h.b); }

// calc_remaining_cards (:13-41) + player_allowed_to_have (:78-282)
DK_HD void doko_assign_prepare(const dk_state& s, AssignPrep& m) {
    const uint32_t obs = st_phase(s) == DK_PHASE_FINISHED ? 0u : st_cur(s);
    m.valid = 1; m.observer = obs;
    const uint32_t own_a = hand_any24(s.hands[obs]), own_b = hand_both24(s.hands[obs]);
    m.obs_a = own_a; m.obs_b = own_b;
    // counts of played cards
    uint32_t p1 = 0, p2 = 0;
    const uint32_t ci = s.card_index;
    for (uint32_t j = 0; j < ci; ++j) { uint32_t bit = 1u << s.cards[j]; p2 |= p1 & bit; p1 |= bit; }
    // remaining = 2 - played - own per card type
    uint32_t gone1 = p1 | own_a, gone2 = p2 | own_b | (p1 & own_a);     // at least one / both copies accounted for
    m.remaining.a = ~gone2 & 0xFFFFFFu;                                  // count >= 1
    m.remaining.b = ~gone1 & 0xFFFFFFu;                                  // count == 2
    uint32_t drop[4] = {0, 0, 0, 0};
    for (uint32_t j = 0; j < ci; ++j) {                                  // seats that did not follow lose the led colour (:226-246)
        uint32_t f = follow_mask(s.cards[j & ~3u], DOKO_TRUMP_MASK);
        if (!((f >> s.cards[j]) & 1u)) drop[(st_trick_start(s, j >> 2) + (j & 3u)) & 3u] |= f;
    }
    const uint32_t tag = st_team_tag(s);
    if (tag == TEAM_WEDDING_UNSOLVED || tag == TEAM_WEDDING_SOLVED)      // wedding announced: nobody else holds a ♣Q (:252-268)
        for (uint32_t p = 0; p < 4u; ++p) if (p != st_wed_seat(s)) drop[p] |= 1u << CARD_CQ;
    for (uint32_t p = 0; p < 4u; ++p) {
        m.len[p] = popcll(s.hands[p]);
        m.allowed[p].a = p == obs ? 0u : (m.remaining.a & ~drop[p]);
        m.allowed[p].b = p == obs ? 0u : (m.remaining.b & ~drop[p]);
    }
}

// The sample's working state.  The reference keeps one "allowed" list per seat and removes a distributed card from every one of
// them (:284-316); the lists are always the REMAINING multiset restricted to the card types the seat may hold, so only the 24-bit
// type mask `allow` is kept per seat (0 for a seat without open slots — the reference clears a full hand's list, and never offers
// anything to a seat whose hand is already complete) together with the list's length `cnt` (parked at ASSIGN_CNT_DONE when the seat is
// full): the eight population counts per iteration of distribute_exactly_as_per_hand's test and the per-seat multiset updates of
// distribute_card become a compare and a decrement (same technique as the full-rules sampler, matching.cuh).
constexpr uint32_t ASSIGN_CNT_DONE = 255u;
// The three HIDDEN seats are kept in absolute seat order in slots 0..2 (seat of slot j = j + (j >= observer), the order every rule
// of the reference walks them in); the observer never receives a card, so carrying it through the loops as a fourth seat (first form)
// cost a quarter of distribute_card and the wider "exactly one of four" mask.
struct AssignState { Hand2 remaining, hand[3]; uint32_t allow[3], len[3], cnt[3], n_remaining; };

// distribute_card (:284-316), branch-free: the receiving seat is selected by a mask
DK_HD void doko_assign_distribute(AssignState& a, uint32_t player, uint32_t c) {
    const uint32_t bit = 1u << c;
#pragma unroll
    for (uint32_t i = 0; i < 3u; ++i) {
        const bool mine = i == player;
        h2_add(a.hand[i], mine ? bit : 0u);
        a.len[i] -= mine ? 1u : 0u;
        a.cnt[i] -= (a.allow[i] & bit) ? 1u : 0u;
        const bool full = mine && a.len[i] == 0u;                    // a full hand can take none of the remaining cards
        a.cnt[i] = full ? ASSIGN_CNT_DONE : a.cnt[i];
        a.allow[i] = full ? 0u : a.allow[i];
    }
    h2_remove_one(a.remaining, bit);
    a.n_remaining -= 1u;
}
DK_HD uint32_t doko_assign_eligible(const AssignState& a, uint32_t bit) {
    uint32_t m = 0;
#pragma unroll
    for (uint32_t i = 0; i < 3u; ++i) if (a.allow[i] & bit) m |= 1u << i;
    return m;
}
// k-th element (0-based) of a multiset listed in ascending card id with doubles adjacent.
DK_HD uint32_t h2_select_adjacent(const Hand2& h, uint32_t k) {
    uint32_t lo = 0, hi = 24;                           // invariant: answer in [lo, hi)
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        uint32_t mid = (lo + hi) >> 1;
        uint32_t below = (1u << mid) - 1u;
        uint32_t cnt = popc(h.a & below) + popc(h.b & below);   // elements with card id < mid
        if (k >= cnt) lo = mid; else hi = mid;
    }
    return lo;
}
// The same with a 64-entry table for the last level: three halvings (12 / 6 / 3 card types, two population counts each), then
// entry (a & 7) | (b & 7) << 3 lists the offsets of the group's elements, two bits each (adj3_entry).
DK_HD uint32_t adj3_entry(uint32_t idx) {
    const uint32_t a3 = idx & 7u, b3 = (idx >> 3) & 7u;
    uint32_t e = 0, j = 0;
    for (uint32_t pos = 0; pos < 3u; ++pos) {
        if ((a3 >> pos) & 1u) { e |= pos << (2u * j); j++; }
        if ((a3 >> pos) & (b3 >> pos) & 1u) { e |= pos << (2u * j); j++; }
    }
    return e;
}
DK_HD uint32_t h2_select_adjacent_lut(const Hand2& h, uint32_t k, const uint32_t* __restrict__ adj3) {
    uint32_t xa = h.a, xb = h.b, pos = 0, c;
    c = popc(xa & 0xFFFu) + popc(xb & 0xFFFu); if (k >= c) { k -= c; xa >>= 12; xb >>= 12; pos = 12u; }
    c = popc(xa & 0x3Fu) + popc(xb & 0x3Fu);   if (k >= c) { k -= c; xa >>= 6;  xb >>= 6;  pos += 6u; }
    c = popc(xa & 7u) + popc(xb & 7u);         if (k >= c) { k -= c; xa >>= 3;  xb >>= 3;  pos += 3u; }
    return pos + ((adj3[(xa & 7u) | ((xb & 7u) << 3)] >> (2u * k)) & 3u);
}
// k-th set bit (0-based, from the LSB) of a 3-bit mask
DK_HD uint32_t select_lsb3(uint32_t e, uint32_t k) {
    uint32_t pos = 0, seen = 0;
#pragma unroll
    for (uint32_t i = 0; i < 3u; ++i) { const uint32_t on = (e >> i) & 1u; pos = (on && seen == k) ? i : pos; seen += on; }
    return pos;
}
// sample_assignment (:493-581).  Returns 0 ok, 1 dead end.  adj3 (optional): the table of h2_select_adjacent_lut in shared memory.
DK_HD uint32_t doko_assign_sample(const AssignPrep& p, const RngKey& key, uint64_t hands_out[4], const uint32_t* __restrict__ adj3 = nullptr) {
    AssignState a;
    a.remaining = p.remaining;
    a.n_remaining = h2_count(p.remaining);
    const uint32_t obs = p.observer;
#pragma unroll
    for (uint32_t j = 0; j < 3u; ++j) {
        // slot j = absolute seat j + (j >= obs): a select over compile-time seats, no indexed access
        const bool up = j >= obs;
        const uint32_t ln = up ? p.len[j + 1u] : p.len[j];
        const Hand2 al = up ? p.allowed[j + 1u] : p.allowed[j];
        a.len[j] = ln; a.hand[j].a = 0; a.hand[j].b = 0;
        a.allow[j] = ln ? al.a : 0u;
        a.cnt[j] = ln ? h2_count(al) : ASSIGN_CNT_DONE;
    }
    U4 blk; blk.x = blk.y = blk.z = blk.w = 0;
    uint32_t blk_id = 0xFFFFFFFFu, ord = 0, status = 0;
    // The body is data dependent and the lanes of a warp diverge in it, so it is kept small: the single-owner rule and the random rule
    // share ONE distribute call at the end of the body, distribute_exactly_as_per_hand has the other (the first version had eleven
    // inlined copies and spent 13 % of its instructions on register moves at the control-flow joins).
    for (;;) {
        uint32_t player, c;
        // distribute_single_cards (:336-377): first remaining card (ascending id) with exactly one eligible seat — bit-parallel:
        // "exactly one of the three hidden seats" = odd parity minus "all three"
        const uint32_t e0 = a.allow[0], e1 = a.allow[1], e2 = a.allow[2];
        const uint32_t one = (e0 ^ e1 ^ e2) & ~(e0 & e1 & e2) & a.remaining.a;
        if (one) {
            // The last open seat takes what is left (see fdo_match_sample): one hidden seat open, every remaining card type allowed
            // for it, counts equal — distribute_single_cards would hand it the remaining multiset card by card and the sample would
            // be complete, without a draw in between.
            const uint32_t al = e0 | e1 | e2;
            const bool one_open = ((e0 | e1) == 0u) | ((e0 | e2) == 0u) | ((e1 | e2) == 0u);
            if (one_open && (a.remaining.a & ~al) == 0u) {
                const uint32_t j = e0 ? 0u : (e1 ? 1u : 2u);
                const uint32_t ln = j == 0u ? a.len[0] : (j == 1u ? a.len[1] : a.len[2]);
                if (a.n_remaining == ln) {
#pragma unroll
                    for (uint32_t o = 0; o < 3u; ++o) {
                        const uint32_t xa = o == j ? a.remaining.a : 0u, xb = o == j ? a.remaining.b : 0u;
                        a.hand[o].b |= xb | (a.hand[o].a & xa);
                        a.hand[o].a |= xa;
                    }
                    break;
                }
            }
            c = ffs0(one);
            const uint32_t bit = 1u << c;
            player = (e0 & bit) ? 0u : ((e1 & bit) ? 1u : 2u);
        } else {
            // distribute_exactly_as_per_hand (:379-417): the first seat (ascending) whose open slots equal its list takes the whole list
            // (a snapshot; ascending card id, the second copy of a double right after the first)
            uint32_t seat = 3u;
#pragma unroll
            for (uint32_t i = 0; i < 3u; ++i)
                if (seat == 3u && a.len[i] == a.cnt[i]) seat = i;
            if (seat < 3u) {
                const uint32_t al = seat == 0u ? a.allow[0] : (seat == 1u ? a.allow[1] : a.allow[2]);
                Hand2 pend; pend.a = a.remaining.a & al; pend.b = a.remaining.b & al;
                while (pend.a) { const uint32_t cc = ffs0(pend.a); h2_remove_one(pend, 1u << cc); doko_assign_distribute(a, seat, cc); }
                continue;
            }
            // distribute_single_card_randomly (:419-456)
            const uint32_t n = a.n_remaining;
            if (n == 0u) break;
            // one word of SITE_ASSIGN per card: the card and then its seat are chained draws from it (draw_chain, dk_common.cuh; relative
            // bias <= 36*3 / 2^32) — the sampler drew two words per card before, 18 of a sample's Philox blocks instead of 9
            uint32_t w0;
            { uint32_t o = ord++; if ((o >> 2) != blk_id) { blk_id = o >> 2; blk = rng_block(key, SITE_ASSIGN, blk_id); } w0 = u4_word(blk, o & 3u); }
            const uint32_t k = draw_chain(w0, n);
            c = adj3 ? h2_select_adjacent_lut(a.remaining, k, adj3) : h2_select_adjacent(a.remaining, k);
            const uint32_t e = doko_assign_eligible(a, 1u << c);
            if (e == 0u) { status = 1u; break; }            // `.choose(rng).unwrap()` on an empty list would panic
            player = select_lsb3(e, draw_chain(w0, popc(e)));
        }
        doko_assign_distribute(a, player, c);
    }
    uint64_t hs[3];
#pragma unroll
    for (uint32_t j = 0; j < 3u; ++j) hs[j] = (uint64_t)a.hand[j].a | ((uint64_t)a.hand[j].b << 24);
    const uint64_t own = (uint64_t)p.obs_a | ((uint64_t)p.obs_b << 24);     // hand_from_vec(hand_to_vec(own)): canonical copy order
#pragma unroll
    for (uint32_t seat = 0; seat < 4u; ++seat)
        hands_out[seat] = seat == obs ? own : (seat == 0u ? hs[0] : (seat == 1u ? (obs == 0u ? hs[0] : hs[1]) : (seat == 2u ? (obs <= 1u ? hs[1] : hs[2]) : hs[2])));
    return status;
}

}  // namespace dk
