// matching.cuh — full-rules determinizer (rs-full-doko/src/matching/card_matching.rs:31-467) as a thread-per-sample program.
//
// A multiset hand is two 24-bit planes: A = "at least one copy", B = "both copies" (B ⊆ A, which is exactly the invariant the
// reference's add/remove keep: add fills copy A first, remove takes copy B first, hand.rs:217-247).  The constraint tables of an
// info-state (available cards, per-seat possible cards, open slots, ♣Q obligations, visible reservations) are computed ONCE per
// block and staged in shared memory; every thread then runs the greedy rule loop for its own sample with its own Philox stream.
#pragma once
#include "state_ops.cuh"

namespace dk {

struct Hand2 { uint32_t a, b; };
DK_HD bool h2_has(const Hand2& h, uint32_t bit) { return (h.a & bit) != 0u; }
// branch-free (three logic operations each; `bit` may be 0 = no-op, and removing a card that is not there is a no-op)
DK_HD void h2_add(Hand2& h, uint32_t bit) { const uint32_t t = h.a & bit; h.b |= t; h.a |= bit; }      // second copy goes to plane B
DK_HD void h2_remove_one(Hand2& h, uint32_t bit) { const uint32_t t = h.b & bit; h.a &= ~(bit ^ t); h.b ^= t; }   // copy B first; also "remove_ignore"
DK_HD uint32_t h2_len(const Hand2& h) { return popc(h.a) + popc(h.b); }

// Constraint tables of one info-state.  The three hidden seats are kept in ABSOLUTE seat order in slots 0..2
// (seat of slot j = j + (j >= observer)), which is the order every rule of the reference walks them in.
struct MatchPrep {
    Hand2 avail;
    Hand2 possible[3];
    uint32_t slots[3];
    uint32_t must_q;        // bit j: slot j must still receive a ♣Q
    uint32_t observer;
    uint32_t played_q;      // 2 bits per ABSOLUTE seat: ♣Q already played by that seat
    uint32_t visible;       // 4 bits per ABSOLUTE seat: 0 NoneYet, 1 NotRevealed, else 2 + FdoReservation code (revealed)
    uint32_t hidden_res;    // != 0: some seat's reservation is NotRevealed, so a sample draws reservations (SITE_MATCH_RESERVATION)
    uint32_t obs_a, obs_b;  // the observer's real hand
    uint32_t valid;         // 0 when the game is finished (card_matching would panic)
};

// card_matching.rs:241-395 — everything before `execute`.
DK_HD void fdo_match_prepare(const dk_state& s, MatchPrep& m) {
    const uint32_t obs = st_cur(s), start = st_game_start(s), nres = s.n_reservations;
    m.valid = st_phase(s) != DK_PHASE_FINISHED;
    m.observer = obs;
    uint32_t any[4], both[4];
    for (uint32_t p = 0; p < 4u; ++p) { any[p] = hand_any24(s.hands[p]); both[p] = hand_both24(s.hands[p]); }
    m.obs_a = any[obs]; m.obs_b = both[obs];
    uint32_t A = 0, B = 0;
    for (uint32_t p = 0; p < 4u; ++p) {                      // plus_hand over the three hidden hands (:258-265)
        if (p == obs) continue;
        B |= both[p] | (A & any[p]);
        A |= any[p];
    }
    m.avail.a = A; m.avail.b = B;
    const bool completed = nres == 4u;
    uint32_t vis = 0, wedding_seats = 0;
    bool solo_seen = false;
    for (uint32_t i = 0; i < nres; ++i) {                    // get_visible_reservations(observer) (visible_reservations_logic.rs:7-70)
        uint32_t seat = (start + i) & 3u, code = s.reservations[i], v;
        if (code == 0u) v = 2u;                              // Healthy
        else if (code == 1u) { v = (completed || seat == obs) ? 3u : 1u; if (v == 3u) wedding_seats |= 1u << seat; }
        else if (completed && !solo_seen) { v = 2u + code; solo_seen = true; }
        else v = 1u;
        vis |= v << (4u * seat);
    }
    m.visible = vis;
    uint32_t hidden = 0;
    for (uint32_t seat = 0; seat < 4u; ++seat) hidden |= ((vis >> (4u * seat)) & 15u) == 1u ? 1u << seat : 0u;
    m.hidden_res = hidden;
    uint32_t pq = 0;
    const uint32_t ci = s.card_index;
    for (uint32_t j = 0; j < ci; ++j)
        if (s.cards[j] == CARD_CQ) pq += 1u << (2u * ((st_trick_start(s, j >> 2) + (j & 3u)) & 3u));
    m.played_q = pq;
    const uint32_t CQ = 1u << CARD_CQ;
    uint32_t must = 0;
    uint32_t drop[4] = {0, 0, 0, 0};                         // cards a seat cannot hold (both planes)
    for (uint32_t p = 0; p < 4u; ++p)                        // a visible wedding: nobody else holds a ♣Q (:289-310)
        if (wedding_seats & ~(1u << p)) drop[p] |= CQ;
    const uint32_t gt = st_gt(s);
    if (gt != (uint32_t)DK_GT_NONE) {
        const uint32_t trump = trump_mask_for_game_type(gt);
        for (uint32_t j = 0; j < ci; ++j) {                  // gather_impossible_colors (gather_impossible_colors.rs:11-43)
            uint32_t f = follow_mask(s.cards[j & ~3u], trump);
            if (!((f >> s.cards[j]) & 1u)) drop[(st_trick_start(s, j >> 2) + (j & 3u)) & 3u] |= f;
        }
        if (gt == GT_NORMAL) {                               // calls reveal the ♣Q (:357-379)
            const uint32_t re = st_re_mask(s), n_calls = st_n_calls(s);
            for (uint32_t a = 0; a < n_calls; ++a) {
                uint32_t seat = (s.announcements[a] >> 6) & 3u;
                if ((re >> seat) & 1u) { if ((pq >> (2u * seat)) & 3u) must &= ~(1u << seat); else must |= 1u << seat; }
                else drop[seat] |= CQ;
            }
        }
    }
    m.must_q = 0;
    for (uint32_t j = 0; j < 3u; ++j) {
        uint32_t seat = j + (j >= obs ? 1u : 0u);
        m.slots[j] = popcll(s.hands[seat]);
        m.possible[j].a = m.slots[j] ? (A & ~drop[seat]) : 0u;
        m.possible[j].b = m.slots[j] ? (B & ~drop[seat]) : 0u;
        if ((must >> seat) & 1u) m.must_q |= 1u << j;
    }
}

// The sample's working state.  The reference keeps a "possible cards" multiset per hidden seat and removes an assigned card from
// every one of them (:49-76).  Those multisets are always the AVAILABLE multiset restricted to the card types the seat may hold —
// both lose one copy of the same card at the same moment — so only the 24-bit type mask `allow` is kept per seat (0 once the seat is
// full) and possible[o] is (avail.a & allow[o], avail.b & allow[o]) wherever a rule reads it.  An assignment then touches the
// available set, one hand and one counter instead of four multisets: the ALU pipe is what bounds this kernel (84 % busy, LSU 2 %;
// profiles/r02_k3_v1_ncu_summary.json: 32 % of the instructions were assign_card).
// `len[o]` = |possible[o]|, kept up to date by assign_card (every seat that could have held the card loses one): rule 2 fires for a
// seat exactly when len == slots, which replaces six population counts per iteration of the rule loop.  A seat without open slots is
// parked at MATCH_LEN_DONE (never equal to a slot count; its `allow` is empty, so nothing is subtracted any more).
constexpr uint32_t MATCH_LEN_DONE = 255u;
struct MatchState {
    Hand2 avail, assigned[3];
    uint32_t allow[3];
    uint32_t len[3];
    uint32_t slots[3], must_q, status;
};

// CardMatchingState::assign_card (:49-76) over compile-time slot numbers (the record of a sample stays in registers: the indexed
// form put it into local memory — 13 % of the kernel's instructions were LDL / STL, profiles/r01_determinize_attribution.json).
// (Tried and rejected: the rules as a one-card-per-iteration state machine with a single assign site, which is what sped up the
// rs-doko sampler by 1.5x — here the lanes of a warp then sit in different rules and every iteration runs all of them: 2x slower.)
template <uint32_t J>
DK_HD void fdo_match_assign_to(MatchState& m, uint32_t c) {
    const uint32_t bit = 1u << c;
    h2_remove_one(m.avail, bit);
    h2_add(m.assigned[J], bit);
    m.slots[J] -= 1u;
    const bool full = m.slots[J] == 0u;
#pragma unroll
    for (uint32_t o = 0; o < 3u; ++o) m.len[o] -= (o == J || (m.allow[o] & bit)) ? 1u : 0u;       // (the receiving slot always could hold the card)
    m.len[J] = full ? MATCH_LEN_DONE : m.len[J];
    m.allow[J] = full ? 0u : m.allow[J];                                   // a full seat can hold nothing more
    // (`must_q` is not touched: rule 3 reads "must still receive a ♣Q" as "had to at the start and holds none yet")
}
// The receiving slot as a compile-time constant: a block works on ONE info-state, so its threads fill the same seats in the same order
// (rule 4 hands a card to the first seat in seat order that can hold it) and `j` is almost always uniform over a warp — three
// specialised bodies of 18 instructions behind a switch instead of one branch-free body of 49 that updates every slot under a mask.
DK_HD void fdo_match_assign(MatchState& m, uint32_t j, uint32_t c) {
    if (j == 0u) fdo_match_assign_to<0>(m, c);
    else if (j == 1u) fdo_match_assign_to<1>(m, c);
    else fdo_match_assign_to<2>(m, c);
}
// rule 1 (:78-113): walk a SNAPSHOT of the available cards (copy-A bits ascending, then copy-B bits) and hand every card that
// exactly one hidden seat can hold to that seat.  The seats' possible sets only change when a card is assigned, so between two
// assignments the "exactly one owner" mask is constant: instead of visiting every card, jump to the next snapshot card whose bit is
// set in that mask (bit-parallel, exact).
DK_HD uint32_t fdo_match_single_owner_mask(const MatchState& m) {
    const uint32_t p0 = m.allow[0], p1 = m.allow[1], p2 = m.allow[2];
    return m.avail.a & (p0 ^ p1 ^ p2) & ~(p0 & p1 & p2);      // odd parity minus "all three" = exactly one
}
DK_HD void fdo_match_rule1(MatchState& m) {
    uint32_t snap[2] = {m.avail.a, m.avail.b};
#pragma unroll
    for (uint32_t plane = 0; plane < 2u; ++plane) {
        uint32_t todo = snap[plane];
        for (;;) {
            uint32_t cand = todo & fdo_match_single_owner_mask(m);
            if (cand == 0u) break;
            uint32_t c = ffs0(cand), bit = 1u << c;
            todo &= ~(bit | (bit - 1u));                          // everything up to and including c has been visited
            uint32_t j = (m.allow[0] & bit) ? 0u : ((m.allow[1] & bit) ? 1u : 2u);
            fdo_match_assign(m, j, c);
        }
    }
}
// rule 2 (:115-145): a seat whose open slots equal its possible cards takes them all.
DK_HD uint32_t fdo_match_possible_len(const MatchState& m, uint32_t j) { return popc(m.avail.a & m.allow[j]) + popc(m.avail.b & m.allow[j]); }
DK_HD bool fdo_match_rule2(MatchState& m) {
    // fast exit (the common case): if no seat qualifies now, none will — nothing is assigned in between
    if (m.len[0] != m.slots[0] && m.len[1] != m.slots[1] && m.len[2] != m.slots[2]) return false;
    bool changed = false;
#pragma unroll
    for (uint32_t j = 0; j < 3u; ++j) {
        if (m.len[j] == m.slots[j]) {
            changed = true;
            uint32_t snap[2] = {m.avail.a & m.allow[j], m.avail.b & m.allow[j]};
            for (uint32_t plane = 0; plane < 2u; ++plane) {
                uint32_t bits = snap[plane];
                while (bits) { uint32_t c = ffs0(bits); bits &= bits - 1u; fdo_match_assign(m, j, c); }
            }
        }
    }
    return changed;
}
// rule 3 (:147-172): a seat that must hold a ♣Q gets one.  The reference clears a seat's obligation when the seat receives a ♣Q by
// whatever rule; here the obligation mask stays as the info-state gave it (block-uniform, almost always empty: one test) and the seat's
// assigned hand says whether it has been met.
DK_HD bool fdo_match_rule3(MatchState& m) {
    if (m.must_q == 0u) return false;                          // the common case
    bool changed = false;
    const uint32_t CQ = 1u << CARD_CQ;
#pragma unroll
    for (uint32_t j = 0; j < 3u; ++j)
        if (((m.must_q >> j) & 1u) && !(m.assigned[j].a & CQ) && (m.avail.a & m.allow[j] & CQ)) { changed = true; fdo_match_assign(m, j, CARD_CQ); }
    return changed;
}

// Rule 4's draws are chained three to a word (draw_chain, dk_common.cuh): use k of SITE_MATCH_CARD takes word k / 3 — a sample makes 13
// of them on average, so one or two Philox blocks instead of four; relative bias of a draw <= 36*35*34 / 2^32 = 1e-5.
struct MatchRng { U4 blk; uint32_t blk_id, words, v, left; };   // cached block, its index, words consumed, what is left of the running word, draws left in it

// One sample: execute (:207-238) + hidden reservations (:418-464).  hands_out / res_out by ABSOLUTE seat.
// rank6 (optional): the 64-entry rank-select table (rank_lut6_entry) in shared memory — the card of rule 4 then comes from two popc
// levels and one lookup on the idle LSU pipe instead of five compare / shift levels on the ALU pipe.
DK_HD uint32_t fdo_match_sample(const MatchPrep& p, const RngKey& key, uint64_t hands_out[4], uint8_t res_out[4], const uint32_t* __restrict__ rank6 = nullptr) {
    MatchState m;
    m.avail = p.avail; m.must_q = p.must_q; m.status = 0;
#pragma unroll
    for (uint32_t j = 0; j < 3u; ++j) {
        m.allow[j] = p.possible[j].a; m.slots[j] = p.slots[j]; m.assigned[j].a = 0; m.assigned[j].b = 0;
        m.len[j] = p.slots[j] ? h2_len(p.possible[j]) : MATCH_LEN_DONE;
    }
    MatchRng r; r.blk_id = 0xFFFFFFFFu; r.words = 0; r.v = 0; r.left = 0; r.blk.x = r.blk.y = r.blk.z = r.blk.w = 0;
    for (;;) {                                                            // (with nothing left to assign no rule fires and the test before rule 4 ends the loop)
        if (fdo_match_single_owner_mask(m) != 0u) {                       // (rule 1 finds nothing while the mask is empty)
            // The LAST open seat takes what is left.  Rule 4 hands a card to the first seat in seat order that can hold it, so the
            // seats fill up one after the other and a third of a sample's cards (7.5 of 21, host-simulator counts) reach the last seat
            // through rule 1, one assign_card each.  When one seat is open, every available card type is allowed for it and the
            // counts agree, rule 1 would assign exactly the available multiset to it and the sample would be complete (no rule
            // draws in between): do that as one multiset addition.  Any other situation (a card nobody may hold: a dead end) takes
            // the general path, so partial results stay identical too.
            const uint32_t a0 = m.allow[0], a1 = m.allow[1], a2 = m.allow[2], al = a0 | a1 | a2;
            const bool one_open = ((a0 | a1) == 0u) | ((a0 | a2) == 0u) | ((a1 | a2) == 0u);
            if (one_open && (m.avail.a & ~al) == 0u) {
                const uint32_t j = a0 ? 0u : (a1 ? 1u : 2u);
                const uint32_t sl = j == 0u ? m.slots[0] : (j == 1u ? m.slots[1] : m.slots[2]);
                if (popc(m.avail.a) + popc(m.avail.b) == sl) {
#pragma unroll
                    for (uint32_t o = 0; o < 3u; ++o) {
                        const uint32_t xa = o == j ? m.avail.a : 0u, xb = o == j ? m.avail.b : 0u;
                        m.assigned[o].b |= xb | (m.assigned[o].a & xa);
                        m.assigned[o].a |= xa;
                    }
                    m.avail.a = 0u; m.avail.b = 0u;
                    break;
                }
            }
            fdo_match_rule1(m);
        }
        if (fdo_match_rule2(m)) continue;
        if (fdo_match_rule3(m)) continue;
        if ((m.avail.a | m.avail.b) == 0u) break;
        // rule 4 (:174-204): a uniformly chosen available card (ascending-bit index over both planes) goes to the FIRST seat
        // in seat order that can hold it
        uint32_t na = popc(m.avail.a), n = na + popc(m.avail.b);
        if (r.left == 0u) {
            const uint32_t wi = r.words++;
            if ((wi >> 2) != r.blk_id) { r.blk_id = wi >> 2; r.blk = rng_block(key, SITE_MATCH_CARD, wi >> 2); }
            r.v = u4_word(r.blk, wi & 3u);
            r.left = 3u;
        }
        r.left--;
        uint32_t idx = draw_chain(r.v, n);
        const bool first_plane = idx < na;
        const uint32_t plane = first_plane ? m.avail.a : m.avail.b, k = first_plane ? idx : idx - na;
        uint32_t c = rank6 ? select_lsb24_lut(plane, k, rank6 - RANK_LUT_BASE) : select_lsb24(plane, k);
        uint32_t bit = 1u << c;
        uint32_t j = (m.allow[0] & bit) ? 0u : ((m.allow[1] & bit) ? 1u : ((m.allow[2] & bit) ? 2u : 3u));
        if (j == 3u) { m.status = 1u; break; }                // `.first().unwrap()` would panic: dead end
        fdo_match_assign(m, j, c);
    }
    const uint32_t obs = p.observer;
    uint32_t oa[4], ob[4];
#pragma unroll
    for (uint32_t j = 0; j < 3u; ++j) { uint32_t seat = j + (j >= obs ? 1u : 0u); oa[seat] = m.assigned[j].a; ob[seat] = m.assigned[j].b; }
    oa[obs] = p.obs_a; ob[obs] = p.obs_b;
    U4 rb; rb.x = rb.y = rb.z = rb.w = 0;
    if (p.hidden_res != 0u) rb = rng_block(key, SITE_MATCH_RESERVATION, 0);   // (block-uniform; in the card phase every reservation is known)
    const uint32_t CQ = 1u << CARD_CQ;
    for (uint32_t seat = 0; seat < 4u; ++seat) {
        hands_out[seat] = (uint64_t)oa[seat] | ((uint64_t)ob[seat] << 24);
        uint32_t v = (p.visible >> (4u * seat)) & 15u, code;
        if (v == 0u) code = 0xFFu;
        else if (v >= 2u) code = v - 2u;
        else {                                                // NotRevealed: draw among [Wedding?] + the seven solos
            uint32_t pq = (p.played_q >> (2u * seat)) & 3u;
            bool wed_ok = (ob[seat] & CQ) || ((oa[seat] & CQ) && pq == 1u) || pq == 2u;
            uint32_t idx = mulhi(u4_word(rb, seat), wed_ok ? 8u : 7u);
            code = wed_ok ? (idx == 0u ? 1u : idx + 1u) : idx + 2u;
        }
        res_out[seat] = (uint8_t)code;
    }
    return m.status;
}

// clone_with_different_hands_and_reservations (state.rs:96-119) on the record.
DK_HD void fdo_state_with_hands_and_reservations(dk_state& s, const uint64_t hands[4], const uint8_t res[4]) {
    const uint32_t start = st_game_start(s);
    uint32_t n = 0;
    for (uint32_t p = 0; p < 4u; ++p) s.hands[p] = hands[p];
    for (uint32_t i = 0; i < 4u; ++i) { uint8_t r = res[(start + i) & 3u]; if (r != 0xFFu) s.reservations[n++] = r; }
    for (uint32_t i = n; i < 4u; ++i) s.reservations[i] = 0xFF;
    s.n_reservations = (uint8_t)n;
}

}  // namespace dk
