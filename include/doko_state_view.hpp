// doko_state_view.hpp — the by-value game state of the reference (`FdoState`, rs-full-doko/src/state/state.rs:24-75) as a plain C++
// struct, and the conversion between it and the 128-byte `dk_state` record of the C ABI (doko_cuda.h).
//
// This is the C++ twin of `impl From<&FdoState> for dk_state` / `impl From<&dk_state> for FdoState` in rs-doko-cuda/src/convert.rs: an
// integrator who holds reference states (Rust) or their fields (any language) converts them with exactly these rules.  Host code only, no
// CUDA, header-only; compiled and checked against the oracle's exports by tests/test_state_view.py (CPU) and used by doko_env.hpp.
//
// Field by field (reference file:line → record field):
//   reservations_round.reservations (PlayerOrientedVec, reservation_round.rs:10-12)  → reservations[k] in PLAY order, n_reservations,
//                                                                                      meta bits 4-5 = starting player
//   tricks (heapless::Vec<FdoTrick,12>, trick.rs:11-17)                               → cards[4t+k] (play order), tricks bits 2t..2t+1 = start seat
//                                                                                      of trick t, bits 24-27 = number of tricks started
//   hands (PlayerZeroOrientedArr<FdoHand>, hand.rs:20)                                → hands[seat] (the same 48-bit board)
//   announcements.announcements (announcement.rs:38-43)                               → announcements[k] = card_index | player << 6 | level << 8,
//                                                                                      tricks bits 28-31 = count
//   announcements.{re,contra}_lowest_announcement, number_of_turns_without_announcement, starting_player (announcement.rs:45-58)
//                                                                                    → meta bits 20-22, 23-25, 26-28, 29-30
//   announcements.current_player_allowed_announcements                               → NOT stored: a function of the rest (calc_announcement.rs:176-228)
//   card_index, current_player, current_phase                                        → card_index, meta bits 2-3 (0 when None), meta bits 0-1
//   reservation_result, game_type                                                    → meta bits 6-9 (game type; the result is a function of the round)
//   player_eyes, player_num_tricks                                                   → eyes[seat], num_tricks nibble per seat
//   team_state (team_logic.rs:11-28)                                                 → meta bits 10-11 tag, 12-13 wedding seat, 14-15 solved trick, 16-19 re seats
//   end_of_game_stats.player_points                                                  → points[seat] (the other statistics are recomputed from the history)
//   FdoTrick::{winning_player, winning_card}                                         → NOT stored: winner of trick t = lead of trick t+1; recomputed for the last
#pragma once
#include <cstdint>
#include <cstring>

#include "doko_cuda.h"

namespace doko {

struct TrickView {
    int starting_player = 0;
    int n_cards = 0;
    int cards[4] = {-1, -1, -1, -1};   // play order
    int winning_player = -1;           // Option<FdoPlayer>: set when the trick is complete
    int winning_card = -1;
};
struct AnnouncementOccurrenceView { int card_index = 0, player = 0, announcement = DK_ANN_NONE; };

struct FdoStateView {
    int reservation_starting_player = 0, n_reservations = 0;
    int reservations[4] = {DK_RES_NONE, DK_RES_NONE, DK_RES_NONE, DK_RES_NONE};   // play order from the starting player
    int n_tricks = 0;
    TrickView tricks[12];
    uint64_t hands[4] = {0, 0, 0, 0};
    int n_announcements = 0;
    AnnouncementOccurrenceView announcements[12];
    int re_lowest_announcement = DK_ANN_NONE, contra_lowest_announcement = DK_ANN_NONE;
    int number_of_turns_without_announcement = 0, announcement_starting_player = 0;
    int current_player_allowed_call = 0;   // announcement phase: the level (DK_ANN_RE_CONTRA..DK_ANN_BLACK) the seat to move may call, 0 = none
    int card_index = 0;
    int current_player = -1;               // -1 = None (finished)
    int current_phase = DK_PHASE_RESERVATION;
    int reservation_result = 0;            // 0 not decided yet, 1 NoReservation, 2 Solo(player, reservation), 3 Wedding(player)
    int reservation_result_player = -1, reservation_result_reservation = DK_RES_NONE;
    int game_type = DK_GT_NONE;
    uint32_t player_eyes[4] = {0, 0, 0, 0}, player_num_tricks[4] = {0, 0, 0, 0};
    int team_tag = DK_TEAM_IN_RESERVATIONS, wedding_player = -1, solved_trick_index = 0;
    uint32_t re_players = 0;               // bit per seat; meaningful for WeddingSolved / NoWedding
    int player_points[4] = {0, 0, 0, 0};   // valid when current_phase == DK_PHASE_FINISHED
};

namespace detail {
inline uint32_t trump_mask(int gt) {   // rs-full-doko/src/card/card_color_masks.rs:7-246
    const uint32_t J = (1u << 2) | (1u << 8) | (1u << 14) | (1u << 20), Q = (1u << 3) | (1u << 9) | (1u << 15) | (1u << 21), H10 = 1u << 7;
    switch (gt) {
        case DK_GT_NORMAL: case DK_GT_WEDDING: case DK_GT_DIAMONDS_SOLO: return J | Q | H10 | 0x00003Fu;
        case DK_GT_HEARTS_SOLO: return J | Q | H10 | 0x000FC0u;
        case DK_GT_SPADES_SOLO: return J | Q | H10 | 0xFC0000u;
        case DK_GT_CLUBS_SOLO: return J | Q | H10 | 0x03F000u;
        case DK_GT_TRUMPLESS_SOLO: return 0u;
        case DK_GT_QUEENS_SOLO: return Q;
        default: return J;
    }
}
// strength of card c in a trick led by `first` (card_in_trick_logic.rs:17-141): strict > decides, the first of equals wins
inline uint32_t power(int c, int first, uint32_t trump) {
    const uint32_t suit = (uint32_t)c / 6u, rank = (uint32_t)c % 6u, bit = 1u << c;
    const uint32_t so = suit ^ (suit >> 1);                                       // J / Q order: diamond < heart < spade < club
    static const uint32_t plain[6] = {0, 2, 0, 0, 1, 3}, eyes[6] = {0, 10, 2, 3, 4, 11};
    const uint32_t tp = rank == 2u ? 4u + so : (rank == 3u ? 8u + so : (c == 7 ? 12u : plain[rank]));
    const uint32_t follow = ((trump >> first) & 1u) ? trump : ((0x3Fu << (6u * ((uint32_t)first / 6u))) & ~trump);
    return (trump & bit) ? 16u + tp : ((follow & bit) ? 1u + eyes[rank] : 0u);
}
// calc_allowed_announcements in closed form (calc_announcement.rs:51-228): next level of the own team, else the counter, else nothing
inline int allowed_call(uint32_t cards, int own, int enemy, int wedding_shift) {
    const int ml = own == DK_ANN_COUNTER ? 0 : own;
    if (ml < 5 && (int)cards + ml + wedding_shift >= 11) return ml + 1;
    if (enemy >= 1 && enemy <= 5 && own == 0 && (int)cards + enemy + wedding_shift >= 11) return DK_ANN_RE_CONTRA;
    return 0;
}
inline uint32_t popcount64(uint64_t x) { uint32_t n = 0; while (x) { x &= x - 1; ++n; } return n; }
}  // namespace detail

// `impl From<&FdoState> for dk_state`
inline dk_state to_record(const FdoStateView& v) {
    dk_state r;
    std::memset(&r, 0, sizeof r);
    std::memset(r.cards, 0xFF, sizeof r.cards);
    for (int a = 0; a < 12; ++a) r.announcements[a] = 0xFFFF;
    for (int p = 0; p < 4; ++p) { r.hands[p] = v.hands[p]; r.eyes[p] = (uint8_t)v.player_eyes[p]; r.reservations[p] = (uint8_t)DK_RES_NONE; }
    for (int k = 0; k < v.n_reservations; ++k) r.reservations[k] = (uint8_t)v.reservations[k];
    r.n_reservations = (uint8_t)v.n_reservations;
    uint32_t tr = 0;
    int ci = 0;
    for (int t = 0; t < v.n_tricks; ++t) {
        tr |= (uint32_t)v.tricks[t].starting_player << (2 * t);
        for (int k = 0; k < v.tricks[t].n_cards; ++k) r.cards[ci++] = (uint8_t)v.tricks[t].cards[k];
    }
    tr |= (uint32_t)v.n_tricks << 24;
    tr |= (uint32_t)v.n_announcements << 28;
    r.tricks = tr;
    for (int a = 0; a < v.n_announcements; ++a)
        r.announcements[a] = (uint16_t)(v.announcements[a].card_index | (v.announcements[a].player << 6) | (v.announcements[a].announcement << 8));
    uint16_t nt = 0;
    for (int p = 0; p < 4; ++p) nt |= (uint16_t)(v.player_num_tricks[p] << (4 * p));
    r.num_tricks = nt;
    r.card_index = (uint8_t)v.card_index;
    const bool finished = v.current_phase == DK_PHASE_FINISHED;
    for (int p = 0; p < 4; ++p) r.points[p] = finished ? (int8_t)v.player_points[p] : 0;
    const bool wedding = v.team_tag == DK_TEAM_WEDDING_UNSOLVED || v.team_tag == DK_TEAM_WEDDING_SOLVED;
    const bool teams_final = v.team_tag == DK_TEAM_WEDDING_SOLVED || v.team_tag == DK_TEAM_NO_WEDDING;
    uint32_t m = 0;
    m |= (uint32_t)v.current_phase;
    m |= (uint32_t)(v.current_player < 0 ? 0 : v.current_player) << 2;
    m |= (uint32_t)v.reservation_starting_player << 4;
    m |= (uint32_t)(v.game_type < 0 ? DK_GT_NONE : v.game_type) << 6;
    m |= (uint32_t)v.team_tag << 10;
    m |= (uint32_t)(wedding ? v.wedding_player : 0) << 12;
    m |= (uint32_t)(v.team_tag == DK_TEAM_WEDDING_SOLVED ? v.solved_trick_index : 0) << 14;
    m |= (uint32_t)(teams_final ? v.re_players : 0u) << 16;
    m |= (uint32_t)v.re_lowest_announcement << 20;
    m |= (uint32_t)v.contra_lowest_announcement << 23;
    m |= (uint32_t)v.number_of_turns_without_announcement << 26;
    m |= (uint32_t)v.announcement_starting_player << 29;
    r.meta = m;
    return r;
}

// `impl From<&dk_state> for FdoState`: the stored fields plus the derived ones (trick winners, reservation result, allowed call).
inline FdoStateView from_record(const dk_state& r) {
    FdoStateView v;
    const uint32_t m = r.meta, tr = r.tricks;
    v.current_phase = (int)(m & 3u);
    v.current_player = v.current_phase == DK_PHASE_FINISHED ? -1 : (int)((m >> 2) & 3u);
    v.reservation_starting_player = (int)((m >> 4) & 3u);
    v.game_type = (int)((m >> 6) & 15u);
    v.team_tag = (int)((m >> 10) & 3u);
    const bool wedding = v.team_tag == DK_TEAM_WEDDING_UNSOLVED || v.team_tag == DK_TEAM_WEDDING_SOLVED;
    v.wedding_player = wedding ? (int)((m >> 12) & 3u) : -1;
    v.solved_trick_index = (int)((m >> 14) & 3u);
    v.re_players = (m >> 16) & 15u;
    v.re_lowest_announcement = (int)((m >> 20) & 7u);
    v.contra_lowest_announcement = (int)((m >> 23) & 7u);
    v.number_of_turns_without_announcement = (int)((m >> 26) & 7u);
    v.announcement_starting_player = (int)((m >> 29) & 3u);
    v.n_reservations = r.n_reservations;
    for (int k = 0; k < 4; ++k) v.reservations[k] = k < v.n_reservations ? r.reservations[k] : (int)DK_RES_NONE;
    v.card_index = r.card_index;
    for (int p = 0; p < 4; ++p) {
        v.hands[p] = r.hands[p]; v.player_eyes[p] = r.eyes[p]; v.player_num_tricks[p] = (r.num_tricks >> (4 * p)) & 15u;
        v.player_points[p] = v.current_phase == DK_PHASE_FINISHED ? r.points[p] : 0;
    }
    // reservation result (reservation_winning_logic.rs:36-73): the first solo in play order wins, else the last wedding
    if (v.n_reservations == 4) {
        int solo = -1, wed = -1;
        for (int k = 0; k < 4; ++k) {
            if (v.reservations[k] >= DK_RES_DIAMONDS_SOLO && solo < 0) solo = k;
            if (v.reservations[k] == DK_RES_WEDDING) wed = k;
        }
        if (solo >= 0) { v.reservation_result = 2; v.reservation_result_player = (v.reservation_starting_player + solo) & 3; v.reservation_result_reservation = v.reservations[solo]; }
        else if (wed >= 0) { v.reservation_result = 3; v.reservation_result_player = (v.reservation_starting_player + wed) & 3; }
        else v.reservation_result = 1;
    }
    v.n_tricks = (int)((tr >> 24) & 15u);
    const uint32_t trump = detail::trump_mask(v.game_type);
    int ci = 0;
    for (int t = 0; t < v.n_tricks; ++t) {
        TrickView& k = v.tricks[t];
        k.starting_player = (int)((tr >> (2 * t)) & 3u);
        for (int j = 0; j < 4 && ci < v.card_index; ++j) k.cards[k.n_cards++] = r.cards[ci++];
        if (k.n_cards == 4) {                                                     // trick_winning_player_logic.rs:15-45
            uint32_t best = 0; int bestk = 0;
            for (int j = 0; j < 4; ++j) { const uint32_t pw = detail::power(k.cards[j], k.cards[0], trump); if (j == 0 || pw > best) { best = pw; bestk = j; } }
            k.winning_player = (k.starting_player + bestk) & 3;
            k.winning_card = k.cards[bestk];
        }
    }
    v.n_announcements = (int)((tr >> 28) & 15u);
    for (int a = 0; a < v.n_announcements; ++a) {
        const uint16_t o = r.announcements[a];
        v.announcements[a].card_index = o & 63; v.announcements[a].player = (o >> 6) & 3; v.announcements[a].announcement = (o >> 8) & 7;
    }
    if (v.current_phase == DK_PHASE_ANNOUNCEMENT && (v.team_tag == DK_TEAM_WEDDING_SOLVED || v.team_tag == DK_TEAM_NO_WEDDING)) {
        const bool re = (v.re_players >> v.current_player) & 1u;
        v.current_player_allowed_call = detail::allowed_call(detail::popcount64(v.hands[v.current_player]), re ? v.re_lowest_announcement : v.contra_lowest_announcement,
                                                             re ? v.contra_lowest_announcement : v.re_lowest_announcement,
                                                             v.team_tag == DK_TEAM_WEDDING_SOLVED ? v.solved_trick_index : 0);
    }
    return v;
}

}  // namespace doko
