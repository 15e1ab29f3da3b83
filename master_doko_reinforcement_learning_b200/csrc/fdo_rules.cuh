// fdo_rules.cuh — full-rules Doppelkopf (rs-full-doko) as a register-resident, thread-per-game program.
//
// Design (B200-first, not a translation of the reference's Vec/enum state):
//   * The four hands live in FOUR 24-bit registers held in a ROTATING FRAME: frame index 0 is always the seat
//     that leads the current trick, so inside the (unrolled) trick the seat to move is a compile-time register —
//     no dynamic register indexing, no local memory.  One extra register `dup` marks card types whose two copies
//     sit in the same hand (then copy B ⊆ copy A, rs-full-doko/src/hand/hand.rs:217-247).
//   * All warp lanes execute card step (t,k) together (48 lock-step card steps, no divergent control flow).  The
//     announcement protocol — the data-dependent part of full rules (SURVEY.md §7 "hard parts") — does not depend on
//     WHICH cards are played, so it is replayed afterwards as a flat per-lane state machine
//     (fdo_replay_announcements): first measured version interleaved it and ran at 17.7/32 active lanes.
//   * Scoring inputs (eyes, tricks, Doppelkopf tricks, caught foxes, Karlchen) are accumulated while playing;
//     the reference re-walks all 12 tricks at the end (rs-full-doko/src/stats/stats.rs:46-240).
//   * Random decisions come from per-site Philox sub-streams so the card draws of trick t are the four words of
//     ONE Philox block with static indices.
#pragma once
#include "dk_common.cuh"

namespace dk {

enum : uint32_t { TEAM_IN_RESERVATIONS = 0, TEAM_WEDDING_UNSOLVED = 1, TEAM_WEDDING_SOLVED = 2, TEAM_NO_WEDDING = 3 };
enum : uint32_t { GT_NORMAL = 0, GT_WEDDING = 1 };
enum : uint32_t { CARD_DA = 5, CARD_CJ = 14, CARD_CQ = 15 };

// ---- announcements ------------------------------------------------------------------------------------------
// Closed form of calc_allowed_announcements (rs-full-doko/src/announcement/calc_announcement.rs:51-228):
// levels 0 none, 1 Re/Kontra, 2 No90, 3 No60, 4 No30, 5 Black; 6 = CounterReContra (only reachable through
// hand-built states: action 33 always records ReContra, action/action.rs:297-299).
// Returns the level the seat may call next (1..5), or 0 if its allowed set is empty (the seat is auto-skipped,
// announcement.rs:156-165).  `m` = own team's lowest call, `e` = enemy team's, `c` = cards on hand,
// `w` = solved_trick_index of a solved wedding else 0.  Both a regular Re/Kontra and the counter map to action 33.
DK_HD uint32_t fdo_allowed_call(uint32_t c, uint32_t m, uint32_t e, uint32_t w) {
    uint32_t ml = m == 6u ? 0u : m;                      // all_higher_than(Counter) = {Counter}: no regular level made
    if (ml < 5u && c + ml + w >= 11u) return ml + 1u;    // exactly the next level (calc_announcement.rs:88-141)
    // counter (:146-170): no regular call possible, enemy called a regular level, own team called nothing at all
    if (e >= 1u && e <= 5u && m == 0u && c + e + w >= 11u) return 1u;
    return 0u;
}
// Smallest hand size with which some seat of a team (own lowest m, enemy lowest e) could still call.
DK_HD uint32_t fdo_min_cards_to_call(uint32_t m, uint32_t e, uint32_t w) {
    // = min(regular: m' < 5 ? 11 - m' - w : none, counter: (1 <= e <= 5 && m == 0) ? 11 - e - w : none) with m' = (m == 6 ? 0 : m):
    // the team has called nothing (m == 0): 11 - w, lowered by the enemy's regular level; otherwise 11 - w - m', none (99) at Black.
    uint32_t sub = m == 0u ? (e - 1u < 5u ? e : 0u) : ((0x0F43210u >> (4u * m)) & 15u);
    return sub == 15u ? 99u : 11u - w - sub;
}

// The same through the shared-memory table (words THR_LUT_BASE + m, nibble e = the amount the base threshold 11 - w is lowered by,
// 15 = the team can never call again): one lookup on the LSU pipe in the replay loop instead of a select chain on the ALU pipe.
DK_HD uint32_t fdo_thr_lut_word(uint32_t m) {
    uint32_t word = 0;
    for (uint32_t e = 0; e < 8u; ++e) {
        uint32_t sub = m == 0u ? ((e >= 1u && e <= 5u) ? e : 0u) : (m < 5u ? m : (m == 5u ? 15u : 0u));
        word |= sub << (4u * e);
    }
    return word;
}
DK_HD uint32_t fdo_min_cards_to_call_lut(uint32_t m, uint32_t e, uint32_t w, const uint32_t* __restrict__ lut) {
    uint32_t sub = (lut[THR_LUT_BASE + m] >> (4u * e)) & 15u;
    return sub == 15u ? 99u : 11u - w - sub;
}

// Both teams' thresholds in one word, indexed by 64 w + 8 re_low + ko_low (w = wedding shift 0..2): thr_re | thr_ko << 8 (99 = never).
DK_HD uint32_t fdo_thr2_lut_word(uint32_t i) {
    const uint32_t w = i >> 6, re = (i >> 3) & 7u, ko = i & 7u;
    return fdo_min_cards_to_call(re, ko, w) | (fdo_min_cards_to_call(ko, re, w) << 8);
}
// One segment of an announcement round: `win` (4 bits, bit d = the seat d places after the next seat to ask is eligible and is
// reached before the count of consecutive passes hits 4), `hit` = the decision bits of those seats in visiting order (bit k = the
// k-th eligible seat calls).  Entry = d | j << 2: the first caller is the (j+1)-th eligible seat, d places ahead.  0 when nobody calls.
DK_HD uint32_t fdo_seg_lut_byte(uint32_t win, uint32_t hit) {
    if (hit == 0u) return 0u;
    uint32_t j = 0;
    while (!((hit >> j) & 1u)) j++;
    uint32_t wj = win;
    for (uint32_t k = 0; k < j; ++k) wj &= wj - 1u;
    if (wj == 0u) return 0u;                             // more decision bits than eligible seats: not a reachable index
    uint32_t d = 0;
    while (!((wj >> d) & 1u)) d++;
    return d | (j << 2);
}
// Word i of the shared lookup table (layout: dk_common.cuh CARD_LUT_WORDS; the SEL12 region is filled from sel12_entry).
DK_HD uint32_t lut_word(uint32_t i) {
    if (i < THR_LUT_BASE) return lead_lut_entry(i / 24u, i % 24u);
    if (i < THR_LUT_BASE + 7u) return fdo_thr_lut_word(i - THR_LUT_BASE);
    if (i < RANK_LUT_BASE) return 0u;
    if (i < POW_LUT_BASE) return rank_lut6_entry(i - RANK_LUT_BASE);
    if (i < CARD_LUT_WORDS) { const uint32_t idx = i - POW_LUT_BASE; return pow_lut_entry(idx / 120u, (idx / 24u) % 5u, idx % 24u); }
    if (i < THR2_LUT_BASE) return 0u;                    // (SEL12 region)
    if (i < SEG_LUT_BASE) return fdo_thr2_lut_word(i - THR2_LUT_BASE);
    uint32_t v = 0;
    for (uint32_t b = 0; b < 4u; ++b) { const uint32_t idx = 4u * (i - SEG_LUT_BASE) + b; v |= fdo_seg_lut_byte(idx >> 4, idx & 15u) << (8u * b); }
    return v;
}

// ---- scoring ----------------------------------------------------------------------------------------------------
// Closed form of FdoEndOfGameStats::calculate (stats/stats.rs:46-240) and its callees re_won (win_conditions/re_won.rs),
// kontra_won, FdoBasicWinningPointsDetails::calculate (basic_points/basic_winning_points.rs:48-284),
// FdoBasicDrawPointsDetails::calculate (basic_points/basic_draw_points.rs:27-174) and
// FdoAdditionalPointsDetails (additional_points/additional_points.rs:57-127).
// rl/kl: lowest calls (0..6); extras = (doko_re - doko_ko) + (fox_re - fox_ko) + (karl_re - karl_ko), ignored in a solo.
// Returns re points; kontra points are written to *kontra_points.
DK_HD int32_t fdo_score(uint32_t re_eyes, uint32_t re_tricks, uint32_t n_re_players, uint32_t rl, uint32_t kl, int32_t extras,
                        int32_t* kontra_points) {
    // Straight-line form (every rollout of every kernel ends here; the case-by-case form it replaced — kept by the test suite as the
    // specification and compared with this one over every input, test_straight_line_score_equals_its_specification — was 190
    // instructions, a tenth of a leaf rollout).  a = eyes / 30 counts the 30-eye steps a side has reached, which turns every "< 90 / < 60 / < 30" and
    // ">= 120 / >= 90 / >= 60 / >= 30 against the other side's No90 .. Black" sum into a clamp.
    const uint32_t ko_eyes = 240u - re_eyes;
    const bool re_all = re_tricks == 12u, ko_all = re_tricks == 0u;
    const uint32_t R = rl == 6u ? 1u : rl, K = kl == 6u ? 1u : kl;   // Counter counts as "Re/Kontra said"
    const int32_t a_re = (int32_t)(re_eyes / 30u), a_ko = (int32_t)(ko_eyes / 30u);
    // won?  an own call of level L >= 2 needs 91 + 30 L eyes (151 / 181 / 211; Black: every trick); otherwise against the other side's
    // level O >= 2, 150 - 30 O eyes are enough (90 / 60 / 30; against Black: one trick); otherwise 121 — 120 for Kontra, and for Re
    // instead when only Kontra was said
    const uint32_t only_kontra = (R == 0u && K == 1u) ? 1u : 0u;
    const uint32_t thr_re = R >= 2u ? 91u + 30u * R : (K >= 2u ? 150u - 30u * K : 121u - only_kontra);
    const uint32_t thr_ko = K >= 2u ? 91u + 30u * K : (R >= 2u ? 150u - 30u * R : 120u + only_kontra);
    const bool re_won = R == 5u ? re_all : ((R < 2u && K == 5u) ? !ko_all : re_eyes >= thr_re);
    const bool ko_won = K == 5u ? ko_all : ((K < 2u && R == 5u) ? !re_all : ko_eyes >= thr_ko);
    const bool solo = n_re_players == 1u;
    // (e)/(f): points for reaching 120/90/60/30 against the other side's No90/No60/No30/Black: the levels j = 2 .. O with a >= 6 - j
    const int32_t lo_re = 6 - a_re > 2 ? 6 - a_re : 2, lo_ko = 6 - a_ko > 2 ? 6 - a_ko : 2;
    const int32_t re_reached = (int32_t)K - lo_re + 1 > 0 ? (int32_t)K - lo_re + 1 : 0;
    const int32_t ko_reached = (int32_t)R - lo_ko + 1 > 0 ? (int32_t)R - lo_ko + 1 : 0;
    const int32_t below_re = 3 - (a_re < 3 ? a_re : 3), below_ko = 3 - (a_ko < 3 ? a_ko : 3);       // (e < 90) + (e < 60) + (e < 30)
    int32_t re_pts;
    if (!re_won && !ko_won) {                            // draw: stats.rs:120-147
        re_pts = below_ko - below_re + re_reached - ko_reached + (solo ? 0 : extras);
    } else {
        const int32_t w = 1 + (re_won ? below_ko : below_re) + ((re_won ? re_all : ko_all) ? 1 : 0) + (R >= 1u ? 2 : 0) + (K >= 1u ? 2 : 0) +
                          (int32_t)(R > 1u ? R : 1u) - 1 + (int32_t)(K > 1u ? K : 1u) - 1 + re_reached + ko_reached;
        // "against the club queens" when Kontra wins (:77-82)
        re_pts = (re_won ? w : -w) + (solo ? 0 : extras - (re_won ? 0 : 1));
    }
    *kontra_points = -re_pts;
    return solo ? 3 * re_pts : re_pts;                   // stats.rs:215-218
}

// ---- live game ------------------------------------------------------------------------------------------------------
struct FdoLive {
    uint32_t h0, h1, h2, h3;  // frame-relative hands ("at least one copy"), 24 bits
    uint32_t dup;             // card types with both copies in one hand
    uint32_t base;            // absolute seat of frame index 0
    uint32_t eyes;            // 8 bits per absolute seat
    uint32_t ntricks;         // 4 bits per absolute seat
    uint32_t dkc;             // 4 bits per absolute seat: tricks with >= 40 eyes (Doppelkopf)
    uint32_t foxes;           // up to two 8-bit records (fdo_fox_record)
    uint32_t trump;           // trump mask of the game type
    uint32_t gt;              // FdoGameType
    uint32_t team_tag, re_mask, wed_seat, solved_idx;
    uint32_t re_low, ko_low;  // lowest call per team (level codes)
    uint32_t karl;            // last trick won with the ♣J
    uint32_t last_winner;     // absolute seat that won the last completed trick
    uint32_t steps;           // play_action calls so far
    uint32_t ann_count;       // announcement decisions so far in this call (ordinal of SITE_ANNOUNCEMENT)
};

DK_HD void fdo_rotate(FdoLive& g, uint32_t r) {  // new frame index i = old index (i + r) & 3
    uint32_t a0 = g.h0, a1 = g.h1, a2 = g.h2, a3 = g.h3;
    if (r & 1u) { uint32_t t = a0; a0 = a1; a1 = a2; a2 = a3; a3 = t; }
    if (r & 2u) { uint32_t t0 = a0, t1 = a1; a0 = a2; a1 = a3; a2 = t0; a3 = t1; }
    g.h0 = a0; g.h1 = a1; g.h2 = a2; g.h3 = a3;
    g.base = (g.base + r) & 3u;
}

DK_HD void fdo_live_clear(FdoLive& g) {
    g.h0 = g.h1 = g.h2 = g.h3 = g.dup = 0; g.base = 0; g.eyes = g.ntricks = g.dkc = g.foxes = 0; g.trump = 0; g.gt = 15;
    g.team_tag = TEAM_IN_RESERVATIONS; g.re_mask = 0; g.wed_seat = 0; g.solved_idx = 0; g.re_low = g.ko_low = 0; g.karl = 0;
    g.last_winner = 0; g.steps = 0; g.ann_count = 0;
}

// Deal: draw 0 is the start seat (draw(4)), draws s = 1..47 are the steps i = 48 - s = 47..1 of a Durstenfeld shuffle (j = draw(i+1),
// swap) of [c0,c0,c1,c1,...]; seat p receives positions 12p..12p+11 (add: copy A first, then copy B)
// (rs-full-doko/src/state/state.rs:169-178, hand/hand.rs:116-188).  Draw s takes word min(s / 3, 11) of SITE_DEAL for s <= 36: three
// chained draws per word (draw_chain; the twelfth word serves the four draws 33..36), so the whole deal is THREE Philox blocks — it was
// ten with one word per draw, 280 of a game's 4700 instructions.  Relative bias of a draw <= 48*47*46 / 2^32 = 2.4e-5.
// `Deck` provides 48 bytes of per-thread scratch, set up a word at a time and then byte-addressed (shared memory on the device,
// word-interleaved across the block: conflict-free).  Position i is final after step i, so its card goes straight into seat i/12's hand.
// Only the steps i = 47..12 are executed: the steps i = 11..1 permute positions 0..11 among themselves — all of them seat 0's — so
// seat 0 simply holds what the other three seats did not get (two copies of every type minus the copies given away).  The draws
// 37..47 (words 12..15 of the site, three per word) are never computed.
template <class Deck>
DK_HD void fdo_deal(FdoLive& g, const RngKey& key, Deck& deck, uint32_t abs_hand[4], uint32_t& dup, uint32_t& start) {
#pragma unroll
    for (uint32_t w = 0; w < 12; ++w) deck.set(w, (2u * w) * 0x0101u + (2u * w + 1u) * 0x01010000u);
    uint32_t h[4] = {0, 0, 0, 0};
    uint32_t d = 0;
    start = 0;
#pragma unroll
    for (uint32_t b = 0; b < 3; ++b) {                    // draws 0 .. 36 <=> start seat, i = 47 .. 12
        U4 blk = rng_block(key, SITE_DEAL, b);
        uint32_t ws[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            const uint32_t word = 4u * b + q;
            uint32_t v = ws[q];
#pragma unroll
            for (uint32_t e = 0; e < (word == 11u ? 4u : 3u); ++e) {
                const uint32_t ord = 3u * word + e;
                if (ord == 0u) { start = draw_chain(v, 4u); continue; }
                uint32_t i = 48u - ord;                   // 47 .. 12
                uint32_t j = draw_chain(v, i + 1u);
                uint32_t ci = deck.get8(i);               // card currently at position i (byte access: no shift / mask arithmetic)
                uint32_t cj = deck.get8(j);               // card that ends up at position i
                deck.set8(j, ci);
                uint32_t bit = 1u << cj;
                uint32_t seat = i / 12u;                  // 3, 2 or 1
                d |= h[seat] & bit;
                h[seat] |= bit;
            }
        }
    }
    {   // seat 0: per card type, the copies nobody else holds
        const uint32_t any = h[1] | h[2] | h[3];
        const uint32_t two = (h[1] & h[2]) | (h[1] & h[3]) | (h[2] & h[3]) | d;      // both copies given away
        h[0] = 0xFFFFFFu & ~two;
        d |= 0xFFFFFFu & ~any;                                                          // both copies stay with seat 0
    }
    abs_hand[0] = h[0]; abs_hand[1] = h[1]; abs_hand[2] = h[2]; abs_hand[3] = h[3];
    dup = d;
    (void)g;
}

// Outcome of the reservation round (reservation/reservation_winning_logic.rs:36-73, team/team_logic.rs:37-122):
// the FIRST solo in seat order from the start seat wins; else the LAST wedding; else a normal game.
// res_action[i] = action index (24..32) chosen by frame seat i (frame base = game start seat).
DK_HD void fdo_finish_reservations(FdoLive& g, const uint32_t res_action[4]) {
    uint32_t solo_i = 4, wed_i = 4;
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        if (res_action[i] >= 26u && solo_i == 4u) solo_i = i;
        if (res_action[i] == 25u) wed_i = i;
    }
    if (solo_i < 4u) {
        uint32_t a = res_action[0];
        a = solo_i == 1u ? res_action[1] : a; a = solo_i == 2u ? res_action[2] : a; a = solo_i == 3u ? res_action[3] : a;
        g.gt = a - 24u;                                   // 26..32 → ♦,♥,♠,♣,Trumpless,Queens,Jacks solo = 2..8
        g.team_tag = TEAM_NO_WEDDING;
        g.re_mask = 1u << ((g.base + solo_i) & 3u);
    } else if (wed_i < 4u) {
        g.gt = GT_WEDDING;
        g.team_tag = TEAM_WEDDING_UNSOLVED;
        g.wed_seat = (g.base + wed_i) & 3u;
        g.re_mask = 0;
    } else {
        g.gt = GT_NORMAL;
        g.team_tag = TEAM_NO_WEDDING;
        uint32_t m = 0;                                   // holders of a ♣Q at the end of the reservation round
        m |= ((g.h0 >> CARD_CQ) & 1u) << (g.base & 3u);
        m |= ((g.h1 >> CARD_CQ) & 1u) << ((g.base + 1u) & 3u);
        m |= ((g.h2 >> CARD_CQ) & 1u) << ((g.base + 2u) & 3u);
        m |= ((g.h3 >> CARD_CQ) & 1u) << ((g.base + 3u) & 3u);
        g.re_mask = m;
    }
    g.trump = trump_mask_for_game_type(g.gt);
}

// One reservation decision of frame seat i with hand `h` (action/allowed_actions.rs:76-96; MSB-first pick over
// Healthy 24, [Wedding 25], solos 26..32).
DK_HD uint32_t fdo_pick_reservation(uint32_t h, uint32_t dup, uint32_t word) {
    uint32_t has_wedding = ((h & dup) >> CARD_CQ) & 1u;
    uint32_t n = 8u + has_wedding;
    uint32_t idx = mulhi(word, n);
    uint32_t a = 32u - idx;                               // rank 0 = bit 32 (Jacks) ... rank 6 = bit 26 (♦-Solo)
    return (!has_wedding && idx == 7u) ? 24u : a;         // without Wedding the 8th choice is Healthy
}

// Announcement protocol (state.rs:184-206, announcement.rs:83-215, calc_announcement.rs:176-228), replayed AFTER the card
// play of the game.  With the random policy the calls never influence which cards are legal, and the cards influence the
// calls only through (a) the hand sizes, which are a function of the card index, and (b) the moment a wedding is solved —
// so the rounds can be run as one per-lane state machine once the tricks are known.  This keeps the 48 card steps of a
// warp in perfect lock-step and makes the cost of the (data-dependent) rounds max-over-lanes of the SUM of visits instead
// of the sum over card positions of the max.
//   ci      card index whose round is running (the round before card ci), p = absolute seat to ask next,
//   turns   consecutive "no" so far (auto-skipped seats count, announcement.rs:156-165)
//   starts  2 bits per trick: absolute seat that leads trick t
// Every decision that actually reaches a seat is one play_action ("game step") and consumes one word of SITE_ANNOUNCEMENT.
// Announcement decisions are two-way draws ({NoAnnouncement, call}; MSB-rank 1 = the call), so the parity contract gives them ONE
// BIT each: decision k of the call is bit (k & 31) of word (k >> 5) of the SITE_ANNOUNCEMENT stream.  Block 0 (128 decisions; the
// longest game seen in 2x10^5 has 47) is computed in lock-step by all lanes before the data-dependent replay loop, so the loop
// contains no Philox code; later blocks are fetched on demand.
// The replay loop reads the stream through a 64-bit shift buffer: peek = one AND, consume = one 64-bit shift; a 32-bit word is
// appended (rare branch, about once per 28 decisions) whenever fewer than 4 bits are left.
struct AnnBits {
    uint64_t buf;      // decisions [ord, ord + avail) of the stream, LSB first
    uint32_t avail;    // valid bits in buf (>= 4 between calls)
    uint32_t next;     // index of the next 32-bit word to append (word k = decisions 32k .. 32k+31 = word k & 3 of block k >> 2)
    U4 w;              // cached Philox block
    uint32_t blk;
};
DK_HD void fdo_ann_refill(AnnBits& st, const RngKey& key) {
    while (st.avail <= 32u) {
        if ((st.next >> 2) != st.blk) { st.blk = st.next >> 2; st.w = rng_block(key, SITE_ANNOUNCEMENT, st.blk); }
        st.buf |= (uint64_t)u4_word(st.w, st.next & 3u) << st.avail;
        st.avail += 32u;
        st.next++;
    }
}
// Position the buffer at decision `ord`; computes the Philox block that holds it (in lock-step, before the data-dependent loop).
DK_HD void fdo_ann_open(AnnBits& st, const RngKey& key, uint32_t ord) {
    st.blk = ord >> 7;
    st.w = rng_block(key, SITE_ANNOUNCEMENT, st.blk);
    const uint32_t wi = ord >> 5;
    st.buf = (uint64_t)(u4_word(st.w, wi & 3u) >> (ord & 31u));
    st.avail = 32u - (ord & 31u);
    st.next = wi + 1u;
    fdo_ann_refill(st, key);
}
DK_HD uint32_t fdo_ann_peek(const AnnBits& st, uint32_t m) { return (uint32_t)st.buf & ((1u << m) - 1u); }   // m <= 4 decisions
DK_HD void fdo_ann_consume(AnnBits& st, const RngKey& key, uint32_t n) {
    st.buf >>= n;
    st.avail -= n;
    if (st.avail < 4u) fdo_ann_refill(st, key);
}

// Byte-parallel "cards on hand >= threshold of the seat's team" for the four seats: `cards4h` holds 0x80 + cards per byte (absolute
// seat), `thr4` the threshold per byte (<= 99, so no borrow crosses a byte); bit 7 of each byte of the difference is the answer and one
// multiplication gathers the four bits into a nibble (bit 8s -> bit 28 + s; the partial products land on distinct bits below 28).
DK_HD uint32_t fdo_spread4(uint32_t nibble) { return (nibble & 1u) | ((nibble & 2u) << 7) | ((nibble & 4u) << 14) | ((nibble & 8u) << 21); }
DK_HD uint32_t fdo_eligible_nibble(uint32_t cards4h, uint32_t thr4) { return ((((cards4h - thr4) >> 7) & 0x01010101u) * 0x10204080u) >> 28; }

// (Measured and not kept, round 2: ONE straight-line body for both kinds of event — a call / the end of a round — selected by
// predicates.  A game makes 10 calls and passes 4.7 rounds on average, so the lanes of a warp sit in different kinds of event in
// nearly every iteration and the two-branch body below executes both branches; the merged body is still 1.6 % slower (3.59 -> 3.64 ms):
// the branches share too little.  A round-synchronous nested loop — all lanes in the same round, calls in an inner loop — costs more as
// well: 12.3 rounds per warp with 38 inner iterations against 22 flat iterations, host-simulator counts.)
// The loop advances by SEGMENTS: from (seat p, `turns` consecutive no's) the next 4 - turns seats are visited unless somebody
// calls; the eligible ones among them each consume one decision bit.  All-zero bits → the round is over (advance to the next
// card); otherwise the first set bit is a call, which changes the levels and restarts the count.  Iterations per lane =
// #calls + #rounds (about 16) instead of #asks + #rounds (about 26), and there is no Philox code inside the loop.
#ifndef DK_REPLAY_ITER
#define DK_REPLAY_ITER()   // host experiments count the loop's iterations per game here (profiles/experiments/replay_balance.cpp)
#endif
template <bool WITH_ANN>
DK_HD void fdo_replay_announcements(FdoLive& g, const RngKey& key, uint32_t starts, uint32_t ci, uint32_t p, uint32_t turns, const uint32_t* __restrict__ lut) {
    const bool wedding = g.team_tag == TEAM_WEDDING_SOLVED;       // (an unsolved wedding cannot survive trick 2)
    const uint32_t w = wedding ? g.solved_idx : 0u;
    const uint32_t first_ci = wedding ? 4u * (g.solved_idx + 1u) : 0u;   // no calls while the wedding is unsolved
    if (ci < first_ci) { ci = first_ci; turns = 0xFFFFFFFFu; }
    if (turns == 0xFFFFFFFFu) {                                   // a fresh round: the seat that plays card ci is asked first
        p = (((starts >> (2u * (ci >> 2))) & 3u) + (ci & 3u)) & 3u;
        turns = 0;
    }
    if (ci >= 48u) return;
    const uint32_t re = g.re_mask & 15u;
    uint32_t re_low = g.re_low, ko_low = g.ko_low, ord = g.ann_count;
    const uint32_t* __restrict__ thr2 = lut + THR2_LUT_BASE + 64u * w;              // both teams' thresholds by 8 re_low + ko_low
    const uint8_t* __restrict__ seg = reinterpret_cast<const uint8_t*>(lut + SEG_LUT_BASE);
    uint32_t tp = thr2[8u * re_low + ko_low], thr_re = tp & 255u, thr_ko = tp >> 8;
    AnnBits st; st.buf = 0; st.avail = 0; st.next = 0; st.blk = 0; st.w.x = st.w.y = st.w.z = st.w.w = 0;
    if (WITH_ANN) fdo_ann_open(st, key, ord);                      // lock-step: every lane fetches its block (block 0 for a fresh game) before the loop
    // Eligibility of the four seats at once, one byte per ABSOLUTE seat:
    //   cards4h  0x80 + cards on hand (the seats that already played in the running trick hold one card less); a seat's byte is
    //            decremented when it plays, so no per-trick recomputation is needed
    //   thr4     the threshold of the seat's team
    // cards >= threshold  <=>  bit 7 of (cards4h - thr4) in that byte (thresholds are <= 99: no borrow crosses a byte); the four bits
    // are gathered into a nibble by one multiplication (bit 8s -> bit 28 + s; the partial products do not collide).
    const uint32_t re_spread = fdo_spread4(re);
    const uint32_t ko_spread = 0x01010101u - re_spread;
    uint32_t thr4 = thr_re * re_spread + thr_ko * ko_spread;
    uint32_t cmax = 12u - (ci >> 2);
    uint32_t q;                                                    // the seat that plays card ci
    uint32_t cards4h;
    {
        const uint32_t base = (starts >> (2u * (ci >> 2))) & 3u;
        q = (base + (ci & 3u)) & 3u;
        uint32_t played = ((1u << (ci & 3u)) - 1u) << base; played = (played | (played >> 4)) & 15u;   // seats that already played in this trick
        cards4h = (0x80u + cmax) * 0x01010101u - fdo_spread4(played);
    }
    uint32_t vis = (1u << (4u - turns)) - 1u;                      // the seats visited before the count reaches 4 (all four after the first segment)
    while (cmax >= (thr_re < thr_ko ? thr_re : thr_ko)) {         // else: monotone, nobody can ever call again
        DK_REPLAY_ITER();
        const uint32_t elig = fdo_eligible_nibble(cards4h, thr4);
        const uint32_t win = ((elig * 0x11u) >> p) & vis;          // bit d: seat p + d is eligible and reached
        const uint32_t m = popc(win);
        const uint32_t hit = WITH_ANN ? fdo_ann_peek(st, m) : 0u;  // (m == 0 reads no bit)
        vis = 15u;
        if (hit == 0u) {                                          // everybody passes: RoundIsOver → card ci is played; next round
            ord += m;
            if (WITH_ANN) fdo_ann_consume(st, key, m);
            cards4h -= 1u << (8u * q);
            ci++;
            if (ci >= 48u) break;
            q = (ci & 3u) ? ((q + 1u) & 3u) : ((starts >> (2u * (ci >> 2))) & 3u);
            p = q;
            cmax = 12u - (ci >> 2);
            continue;
        }
        const uint32_t sv = seg[16u * win + hit];                 // j eligible seats pass, the (j+1)-th calls: it sits d places after p
        const uint32_t j = sv >> 2, d = sv & 3u;
        ord += j + 1u;
        fdo_ann_consume(st, key, j + 1u);
        p = (p + d) & 3u;
        const uint32_t is_re = (re >> p) & 1u;
        const uint32_t c = (cards4h >> (8u * p)) & 0x7Fu;
        uint32_t ml = is_re ? re_low : ko_low;
        ml = ml == 6u ? 0u : ml;
        const uint32_t call = (ml < 5u && c + ml + w >= 11u) ? ml + 1u : 1u;   // next level, else the counter (recorded as Re/Kontra)
        if (is_re) re_low = call; else ko_low = call;             // announcement.rs:203-210
        tp = thr2[8u * re_low + ko_low]; thr_re = tp & 255u; thr_ko = tp >> 8;
        thr4 = thr_re * re_spread + thr_ko * ko_spread;
        p = (p + 1u) & 3u;
    }
    g.re_low = re_low; g.ko_low = ko_low; g.steps += ord - g.ann_count; g.ann_count = ord;
}

struct TrickAcc { uint32_t follow, best, acc, fox, prow; };   // best / acc / fox: dk_common.cuh pow_lut_entry; prow: the trick's row of that table (pow_row)
DK_HD void trick_acc_clear(TrickAcc& a) { a.follow = 0; a.best = 0; a.acc = 0; a.fox = 0; a.prow = 0; }

// Card step of frame seat K (compile-time) with hand register `h` (action/allowed_actions.rs:97-140, state.rs:274-357).
template <int K, bool SEL12 = false>
DK_HD void fdo_card_step(FdoLive& g, uint32_t& h, TrickAcc& a, uint32_t& word, bool last_trick, const uint32_t* __restrict__ lut) {
    uint32_t mask = h;
    if (K > 0 && !last_trick) {                           // state.rs:360-372: no colour is enforced in the 12th trick
        uint32_t f = h & a.follow;
        mask = f ? f : h;
    }
    uint32_t n = popc(mask);
    uint32_t idx = draw_chain(word, n);                   // the trick's word serves its four draws in a row (dk_common.cuh)
    uint32_t c = SEL12 ? pick_msb_rank24_tab(mask, idx, reinterpret_cast<const uint64_t*>(lut + SEL12_LUT_BASE)) : pick_msb_rank24_lut(mask, idx, lut);
    uint32_t bit = 1u << c;
    uint32_t dbl = g.dup & bit;                           // hand.remove: a doubled card stays in the hand once
    g.dup ^= dbl;
    h ^= bit ^ dbl;
    if (K == 0) { const uint32_t e = lead_lookup(lut, g.gt, c); a.follow = lead_follow(e); a.prow = lead_row(e); }
    const uint32_t v = pow_lookup(lut, a.prow, c), cand = v | ((3u - (uint32_t)K) << POW_K_SHIFT);
    a.best = (K == 0 || cand > a.best) ? cand : a.best;               // one max: first of equals wins through the position field
    a.acc += v;
    a.fox += (v & POW_FOX_BIT) * (1u << K);
    g.steps++;
}

// Card step of the TWELFTH trick when it is played from its first card: every seat holds exactly one card and no colour is enforced
// (state.rs:360-372), so the pick has one choice — a draw over one choice is 0 whatever the word is.  The trick's word (word 11 of
// SITE_CARD, never fetched) and the four rank selects are not computed at all; a trick owns its word, so nothing else moves in the
// stream.
template <int K>
DK_HD void fdo_card_step_last(FdoLive& g, uint32_t& h, TrickAcc& a, const uint32_t* __restrict__ lut) {
    const uint32_t c = ffs0(h);
    h = 0u;
    if (K == 0) a.prow = lead_row(lead_lookup(lut, g.gt, c));
    const uint32_t v = pow_lookup(lut, a.prow, c), cand = v | ((3u - (uint32_t)K) << POW_K_SHIFT);
    a.best = (K == 0 || cand > a.best) ? cand : a.best;
    a.acc += v;
    a.fox += (v & POW_FOX_BIT) * (1u << K);
    g.steps++;
}

// ♦A log: one 8-bit record per trick that contained a ♦A (at most two): absolute seats that played one (4 bits) << 2 | winner seat.
// foxm is frame-relative (bit k = k-th card of the trick), lead = absolute seat of the trick's first card.
DK_HD uint32_t fdo_fox_record(uint32_t foxes, uint32_t foxm, uint32_t lead, uint32_t winner) {
    uint32_t abs_mask = ((foxm << lead) | (foxm >> (4u - lead))) & 15u;
    return foxm ? ((foxes << 8) | (abs_mask << 2) | winner) : foxes;
}

// Book-keeping when a trick is complete (state.rs:293-352, team/team_logic.rs:59-112, additional_points/*.rs).
DK_HD void fdo_trick_done(FdoLive& g, const TrickAcc& a, uint32_t t) {
    const uint32_t bestk = pow_best_k(a.best), teyes = a.acc & 255u, foxm = (a.fox >> 8) & 15u;
    uint32_t w = (g.base + bestk) & 3u;
    g.eyes += teyes << (8u * w);
    g.ntricks += 1u << (4u * w);
    if (teyes >= 40u) g.dkc += 1u << (4u * w);
    g.foxes = fdo_fox_record(g.foxes, foxm, g.base, w);
    if (g.team_tag == TEAM_WEDDING_UNSOLVED) {
        if (w != g.wed_seat) { g.team_tag = TEAM_WEDDING_SOLVED; g.solved_idx = t; g.re_mask = (1u << g.wed_seat) | (1u << w); }
        else if (t == 2u) { g.team_tag = TEAM_WEDDING_SOLVED; g.solved_idx = 2u; g.re_mask = 1u << g.wed_seat; }
    }
    if (t == 11u) g.karl = pow_best_card(a.best) == CARD_CJ ? 1u : 0u;
    g.last_winner = w;
    fdo_rotate(g, bestk);
}

// Final scoring from the accumulated trackers → player_points per ABSOLUTE seat.
DK_HD void fdo_final_points(const FdoLive& g, int32_t pts[4]) {
    // Team sums without a loop over the seats: the Re seats' bytes / nibbles are masked and added up by one multiplication each
    // (eyes <= 240 per game, tricks <= 12, Doppelkopf tricks <= 6: no column of the product carries).
    const uint32_t rm = g.re_mask & 15u;
    const uint32_t re8 = fdo_spread4(rm) * 0xFFu;                                   // 0xFF in the bytes of the Re seats
    const uint32_t re4 = ((rm & 1u) | ((rm & 2u) << 3) | ((rm & 4u) << 6) | ((rm & 8u) << 9)) * 0xFu;   // 0xF in their nibbles (bit s -> bit 4 s)
    const uint32_t re_eyes = ((g.eyes & re8) * 0x01010101u) >> 24;
    const uint32_t re_tricks = (((g.ntricks & re4) * 0x1111u) >> 12) & 15u;
    const uint32_t d_re = (((g.dkc & re4) * 0x1111u) >> 12) & 15u, d_all = (((g.dkc & 0xFFFFu) * 0x1111u) >> 12) & 15u;
    int32_t extras = 2 * (int32_t)d_re - (int32_t)d_all;                            // Doppelkopf tricks: Re's minus Kontra's
#pragma unroll
    for (uint32_t f = 0; f < 2; ++f) {                                // caught foxes: ♦A played by the other team than the trick's winner
        uint32_t rec = (g.foxes >> (8u * f)) & 255u, players = rec >> 2;
        bool won_re = (g.re_mask >> (rec & 3u)) & 1u;
        int32_t caught = (int32_t)popc(players & (won_re ? ~g.re_mask : g.re_mask) & 15u);
        extras += won_re ? caught : -caught;
    }
    if (g.karl) extras += ((g.re_mask >> g.last_winner) & 1u) ? 1 : -1;
    int32_t ko;
    int32_t re = fdo_score(re_eyes, re_tricks, popc(g.re_mask), g.re_low, g.ko_low, extras, &ko);
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) pts[s] = ((g.re_mask >> s) & 1u) ? re : ko;
}

// Resume descriptor for playouts that start from a stored state.
struct FdoResume {
    uint32_t n_res;        // reservations already made (0..4); < 4 means we are still in the reservation phase
    uint32_t res_action[4];// actions of the reservations already made, by frame seat (frame base = game start seat)
    uint32_t t0, k0;       // trick index / cards already in that trick
    uint32_t starts;       // lead seats of the tricks 0..t0-1 (2 bits each); trick t0's is g.base
    uint32_t ann_ci, ann_p, ann_turns;   // where the announcement protocol resumes: round before card ann_ci, next ABSOLUTE
                           // seat to ask, turns without call.  Phase PlayCard (round already over) → ann_ci = card_index + 1.
    TrickAcc acc;          // partial trick accumulator (valid when k0 > 0)
    uint32_t chain_mul;    // product of the legal-card counts of the k0 plays already made in trick t0 (1 when k0 == 0): the trick's
                           // word times this product is where the chained card draws of the trick continue (draw_chain)
};

// Plays the game to the end.  FRESH: hands/base already set by the deal, nothing played yet.
// STEPS = false (only with the no-announcement policy): the caller wants the points only, so the announcement rounds — which then
// change nothing but the step count — are not replayed and the ANN region of the table is not needed (leaf rollouts, PIMC, UCT).
template <bool WITH_ANN, bool FRESH, bool SEL12 = false, bool STEPS = true>
DK_HD void fdo_play_to_end(FdoLive& g, const RngKey& key, const FdoResume* rs, const uint32_t* __restrict__ lut) {
    uint32_t n_res = FRESH ? 0u : rs->n_res;
    if (n_res < 4u) {
        U4 blk = rng_block(key, SITE_RESERVATION, 0);
        uint32_t ra[4];
        ra[0] = (!FRESH && n_res > 0u) ? rs->res_action[0] : fdo_pick_reservation(g.h0, g.dup, blk.x);
        ra[1] = (!FRESH && n_res > 1u) ? rs->res_action[1] : fdo_pick_reservation(g.h1, g.dup, blk.y);
        ra[2] = (!FRESH && n_res > 2u) ? rs->res_action[2] : fdo_pick_reservation(g.h2, g.dup, blk.z);
        ra[3] = fdo_pick_reservation(g.h3, g.dup, blk.w);
        g.steps += 4u - n_res;
        fdo_finish_reservations(g, ra);
    }
    uint32_t t = FRESH ? 0u : rs->t0;
    uint32_t starts = FRESH ? 0u : rs->starts;
    // A resumed game first finishes its current (possibly partial) trick; the remaining tricks run through the same unconditional
    // body as a fresh game.  With the `is this position already played` tests inside the one loop, the loop body was divergent code:
    // the shared-memory window base (a uniform register on sm_100) was recomputed at every table access — 4 of ~45 instructions per
    // card step of the kernels that start from stored records (profiles/r02_k4_roll_v1 attribution).
    // Card draws: trick t takes word t & 3 of block t >> 2 of SITE_CARD, its four draws chained through draw_chain — three Philox blocks
    // per game instead of eleven (320 of a game's 4700 instructions).  The block stays in registers over four tricks.
    U4 cblk;
    cblk.x = cblk.y = cblk.z = cblk.w = 0u;
    if (!FRESH && t < 12u) {
        starts |= g.base << (2u * t);
        cblk = rng_block(key, SITE_CARD, t >> 2);
        uint32_t word = u4_word(cblk, t & 3u) * rs->chain_mul;
        TrickAcc a;
        trick_acc_clear(a);
        const uint32_t k0 = rs->k0;
        if (k0 > 0u) a = rs->acc;
        const bool last = t == 11u;
        if (k0 <= 0u) fdo_card_step<0, SEL12>(g, g.h0, a, word, last, lut);
        if (k0 <= 1u) fdo_card_step<1, SEL12>(g, g.h1, a, word, last, lut);
        if (k0 <= 2u) fdo_card_step<2, SEL12>(g, g.h2, a, word, last, lut);
        fdo_card_step<3, SEL12>(g, g.h3, a, word, last, lut);
        fdo_trick_done(g, a, t);
        ++t;
    }
    for (; t < 11u; ++t) {
        starts |= g.base << (2u * t);
        if ((t & 3u) == 0u) cblk = rng_block(key, SITE_CARD, t >> 2);
        uint32_t word = u4_word(cblk, t & 3u);
        TrickAcc a;
        trick_acc_clear(a);
        fdo_card_step<0, SEL12>(g, g.h0, a, word, false, lut);
        fdo_card_step<1, SEL12>(g, g.h1, a, word, false, lut);
        fdo_card_step<2, SEL12>(g, g.h2, a, word, false, lut);
        fdo_card_step<3, SEL12>(g, g.h3, a, word, false, lut);
        fdo_trick_done(g, a, t);
    }
    if (t == 11u) {                                       // the last trick from its first card: forced moves (fdo_card_step_last)
        starts |= g.base << 22;
        TrickAcc a;
        trick_acc_clear(a);
        fdo_card_step_last<0>(g, g.h0, a, lut);
        fdo_card_step_last<1>(g, g.h1, a, lut);
        fdo_card_step_last<2>(g, g.h2, a, lut);
        fdo_card_step_last<3>(g, g.h3, a, lut);
        fdo_trick_done(g, a, 11u);
    }
    // announcement rounds (see fdo_replay_announcements): fresh games start with the round before card 0
    if (!WITH_ANN && !STEPS) return;
    if (FRESH) fdo_replay_announcements<WITH_ANN>(g, key, starts, 0u, starts & 3u, 0u, lut);
    else fdo_replay_announcements<WITH_ANN>(g, key, starts, rs->ann_ci, rs->ann_p, rs->ann_turns, lut);
}

// Fresh game: deal + reservations + 12 tricks + scoring.  `tables_ready` is called between the deal (which reads no table) and the
// first table access: the kernels wait there for the bulk copy that stages the tables, so the copy runs under the deal.
struct NoWait { DK_HD void operator()() const {} };
template <bool WITH_ANN, class Deck, bool SEL12 = false, class Ready = NoWait>
DK_HD void fdo_playout_fresh(const RngKey& key, Deck& deck, const uint32_t* __restrict__ lut, int32_t pts[4], uint32_t& steps, Ready tables_ready = Ready()) {
    FdoLive g;
    fdo_live_clear(g);
    uint32_t ah[4], start;
    fdo_deal(g, key, deck, ah, g.dup, start);
    tables_ready();
    g.h0 = ah[0]; g.h1 = ah[1]; g.h2 = ah[2]; g.h3 = ah[3];
    g.base = 0;
    fdo_rotate(g, start);
    fdo_play_to_end<WITH_ANN, true, SEL12>(g, key, nullptr, lut);
    fdo_final_points(g, pts);
    steps = g.steps;
}

}  // namespace dk
