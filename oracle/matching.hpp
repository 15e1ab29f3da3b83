// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// CPU restatement of the reference's full-rules determinizer and its validity oracle:
//   rs-full-doko/src/matching/card_matching.rs:31-497       (CardMatchingState, card_matching, card_matching_full)
//   rs-full-doko/src/matching/gather_impossible_colors.rs:11-43
//   rs-full-doko/src/matching/is_consistent.rs:38-315
#pragma once
#include <cstdint>
#include "fdo.hpp"

namespace oracle {
namespace fdo {

// gather_impossible_colors (gather_impossible_colors.rs:11-43): bit `color` of out[player].
inline void gather_impossible_colors(const Trick* tricks, int n_tricks, int game_type, uint32_t out[4]) {
    for (int p = 0; p < 4; ++p) out[p] = 0;
    if (game_type == GT_NONE) return;
    for (int t = 0; t < n_tricks; ++t) {
        Color tc = tricks[t].color(game_type);
        if (tc == COLOR_NONE) continue;
        for (int k = 0; k < tricks[t].len; ++k)
            if (card_to_color(tricks[t].cards[k], game_type) != tc) out[tricks[t].player_at(k)] |= 1u << tc;
    }
}

enum MatchStatus : int { MATCH_OK = 0, MATCH_DEAD_END = 1 };  // dead end = the reference would panic

struct CardMatchingState {                                               // card_matching.rs:31-46
    int observing_player;
    Hand available_cards;
    bool must_have_q_club[4];
    uint32_t remaining_card_slots[4];
    Hand card_assignments[4];
    Hand possible_cards[4];
    int status = MATCH_OK;

    void assign_card(int player, int card) {                             // :49-76
        card_assignments[player].add(card);
        remaining_card_slots[player] -= 1;
        available_cards.remove(card);
        possible_cards[player].remove(card);
        if (remaining_card_slots[player] == 0) possible_cards[player] = Hand();
        for (int o = 0; o < 4; ++o) if (o != player) possible_cards[o].remove_ignore(card);
        if (card == CQ) must_have_q_club[player] = false;
    }
    void rule1() {                                                       // :78-113 (iterates a SNAPSHOT of available_cards)
        int cards[48]; int m = available_cards.iter(cards);
        for (int i = 0; i < m; ++i) {
            int card = cards[i];
            int single = -1;
            for (int p = 0; p < 4; ++p) {
                if (p == observing_player) continue;
                if (possible_cards[p].contains(card)) {
                    if (single >= 0) { single = -1; break; }
                    single = p;
                }
            }
            if (single >= 0) assign_card(single, card);
        }
    }
    bool rule2() {                                                       // :115-145
        bool changed = false;
        for (int p = 0; p < 4; ++p) {
            if (p == observing_player) continue;
            if (remaining_card_slots[p] > 0 && remaining_card_slots[p] == possible_cards[p].len()) {
                changed = true;
                int cards[48]; int m = possible_cards[p].iter(cards);
                for (int i = 0; i < m; ++i) assign_card(p, cards[i]);
            }
        }
        return changed;
    }
    bool rule3() {                                                       // :147-172
        bool changed = false;
        for (int p = 0; p < 4; ++p) {
            if (p == observing_player) continue;
            if (must_have_q_club[p] && possible_cards[p].contains(CQ)) { changed = true; assign_card(p, CQ); }
        }
        return changed;
    }
    void rule4(Rng& rng) {                                               // :174-204
        int cards[48]; int m = available_cards.iter(cards);
        int chosen = cards[rng.choose_unsized(SITE_MATCH_CARD, (uint32_t)m)];
        int player = -1;
        for (int p = 0; p < 4; ++p) {                                    // FIRST eligible seat (quirk A.9 (8))
            if (p == observing_player) continue;
            if (possible_cards[p].contains(chosen)) { player = p; break; }
        }
        if (player < 0) { status = MATCH_DEAD_END; return; }             // `.first().unwrap()` would panic
        assign_card(player, chosen);
    }
    void execute(Rng& rng) {                                             // :207-238
        for (;;) {
            if (available_cards.len() == 0) break;
            rule1();
            if (rule2()) continue;
            if (rule3()) continue;
            if (available_cards.len() == 0) break;
            rule4(rng);
            if (status != MATCH_OK) return;
        }
    }
};

// card_matching (card_matching.rs:241-467) with obs = state.observation_for_current_player()
// (CAPSampling::sample, rs-doko-py-bridge/src/compare_impi/compare_impi.rs:64-83).
// out_reservations: per ABSOLUTE seat, R_NONE when the seat has not made a reservation yet.
inline int card_matching(const State& state, Rng& rng, Hand out_hands[4], int out_reservations[4]) {
    if (state.current_player < 0) throw std::runtime_error("Das Spiel ist beendet.");
    int current_player = state.current_player;
    Hand available;
    for (int p = 0; p < 4; ++p) if (p != current_player) available = available.plus_hand(state.hands[p]);  // :258-265
    CardMatchingState m;
    m.observing_player = current_player;
    m.available_cards = available;
    for (int p = 0; p < 4; ++p) {
        m.remaining_card_slots[p] = state.hands[p].len();                // :270-274
        m.possible_cards[p] = available;                                 // :280-285
        m.must_have_q_club[p] = false;
        m.card_assignments[p] = Hand();
    }
    m.remaining_card_slots[current_player] = 0;
    int visible[4]; get_visible_reservations(state.reservations_round, current_player, visible);
    for (int i = 0; i < 4; ++i) {                                        // :289-310 (PlayerOrientedArr order; order-independent)
        int player = player_next(state.reservations_round.starting_player, i);
        if (visible[player] == VR_WEDDING)
            for (int o = 0; o < 4; ++o) if (o != player) { m.possible_cards[o].remove_ignore(CQ); m.possible_cards[o].remove_ignore(CQ); }
    }
    m.possible_cards[current_player] = Hand();                           // :312
    uint32_t played_q_clubs[4] = {0, 0, 0, 0};                           // :314-327
    for (int t = 0; t < state.n_tricks; ++t)
        for (int k = 0; k < state.tricks[t].len; ++k)
            if (state.tricks[t].cards[k] == CQ) played_q_clubs[state.tricks[t].player_at(k)] += 1;
    if (state.game_type != GT_NONE) {                                    // :331-380
        uint32_t impossible[4]; gather_impossible_colors(state.tricks, state.n_tricks, state.game_type, impossible);
        for (int p = 0; p < 4; ++p) {
            if (p == current_player) continue;
            for (int c = 0; c < 5; ++c) if ((impossible[p] >> c) & 1) m.possible_cards[p].remove_color((Color)c, state.game_type);
        }
        if (state.game_type == GT_NORMAL) {
            if (!state.team_state.has_re_players()) throw std::runtime_error("Im Normalspiel stehen die Re-Spieler fest.");
            for (int a = 0; a < state.announcements.n; ++a) {
                int ap = state.announcements.occ[a].player;
                if ((state.team_state.re_players >> ap) & 1) m.must_have_q_club[ap] = !(played_q_clubs[ap] > 0);
                else { m.possible_cards[ap].remove_ignore(CQ); m.possible_cards[ap].remove_ignore(CQ); }
            }
        }
    }
    for (int p = 0; p < 4; ++p) if (m.remaining_card_slots[p] == 0) m.possible_cards[p] = Hand();  // :391-395
    try {
        m.execute(rng);
    } catch (const std::runtime_error&) { m.status = MATCH_DEAD_END; }    // a panic inside assign_card (remove of a missing card)
    for (int p = 0; p < 4; ++p) out_hands[p] = m.card_assignments[p];
    out_hands[current_player] = state.hands[current_player];             // :410-412
    // hidden reservations (:418-464)
    for (int p = 0; p < 4; ++p) {
        int vr = visible[p];
        if (vr == VR_NONE_YET) { out_reservations[p] = R_NONE; continue; }
        // EnumSet iteration order = enum declaration order (reservation.rs:11-24)
        int possible[9]; int np = 0;
        bool wedding_ok = out_hands[p].contains_both(CQ) || (out_hands[p].contains(CQ) && played_q_clubs[p] == 1) || played_q_clubs[p] == 2;
        if (wedding_ok) possible[np++] = R_WEDDING;
        possible[np++] = R_DIAMONDS_SOLO; possible[np++] = R_HEARTS_SOLO; possible[np++] = R_SPADES_SOLO; possible[np++] = R_CLUBS_SOLO;
        possible[np++] = R_QUEENS_SOLO; possible[np++] = R_JACKS_SOLO; possible[np++] = R_TRUMPLESS_SOLO;
        switch (vr) {
            case VR_WEDDING: out_reservations[p] = R_WEDDING; break;
            case VR_HEALTHY: out_reservations[p] = R_HEALTHY; break;
            case VR_NOT_REVEALED: {
                // exact-size iterator ⇒ one index draw.  Philox contract: word = seat index.
                rng.set_ordinal(SITE_MATCH_RESERVATION, (uint32_t)p);
                out_reservations[p] = possible[rng.below(SITE_MATCH_RESERVATION, (uint32_t)np)];
                break;
            }
            case VR_DIAMONDS_SOLO: out_reservations[p] = R_DIAMONDS_SOLO; break;
            case VR_HEARTS_SOLO: out_reservations[p] = R_HEARTS_SOLO; break;
            case VR_SPADES_SOLO: out_reservations[p] = R_SPADES_SOLO; break;
            case VR_CLUBS_SOLO: out_reservations[p] = R_CLUBS_SOLO; break;
            case VR_QUEENS_SOLO: out_reservations[p] = R_QUEENS_SOLO; break;
            case VR_JACKS_SOLO: out_reservations[p] = R_JACKS_SOLO; break;
            case VR_TRUMPLESS_SOLO: out_reservations[p] = R_TRUMPLESS_SOLO; break;
            default: throw std::runtime_error("Unbekannter Vorbehalt.");
        }
    }
    return m.status;
}

// card_matching_full (card_matching.rs:469-497) / clone_with_different_hands_and_reservations (state.rs:96-119)
inline State with_hands_and_reservations(const State& s, const Hand hands[4], const int reservations[4]) {
    State n = s;
    for (int p = 0; p < 4; ++p) n.hands[p] = hands[p];
    ReservationRound rr; rr.starting_player = s.reservations_round.starting_player;
    for (int i = 0; i < 4; ++i) {                                        // rotate_to(start).all_present()
        int r = reservations[player_next(rr.starting_player, i)];
        if (r != R_NONE) rr.play_reservation(r);
    }
    n.reservations_round = rr;
    return n;
}

// _is_consistent (is_consistent.rs:38-305).  Returns 0 when consistent, else 1 + NotConsistentReason index.
enum NotConsistentReason : int {
    NC_OK = 0, NC_HAND_SIZE_MISMATCH = 1, NC_NOT_IN_REMAINING = 2, NC_REMAINING_LEFT = 3, NC_ALREADY_DISCARDED_COLOR = 4,
    NC_CQ_BUT_OTHER_WEDDING = 5, NC_NO_CQ_BUT_RE = 6, NC_CQ_BUT_KONTRA = 7, NC_WRONG_RESERVATION = 8, NC_WRONG_RESERVATION_CQ = 9,
    NC_OBSERVER_HAND_CHANGED = 10,
};
inline int is_consistent(const State& state, const Hand assumed_hands[4], const int assumed_reservations[4] /* by absolute seat */) {
    int observing = state.observing_player();
    if (!(state.hands[observing].bits == assumed_hands[observing].bits)) return NC_OBSERVER_HAND_CHANGED;  // assert! at :54
    for (int p = 0; p < 4; ++p) if (assumed_hands[p].len() != state.hands[p].len()) return NC_HAND_SIZE_MISMATCH;  // :67-73
    Hand remaining = state.hands[0];                                      // :76-80
    remaining = remaining.plus_hand(state.hands[1]); remaining = remaining.plus_hand(state.hands[2]); remaining = remaining.plus_hand(state.hands[3]);
    for (int p = 0; p < 4; ++p) {                                         // :82-92
        int cards[48]; int m = assumed_hands[p].iter(cards);
        for (int i = 0; i < m; ++i) { if (!remaining.contains(cards[i])) return NC_NOT_IN_REMAINING; remaining.remove(cards[i]); }
    }
    if (remaining.len() != 0) return NC_REMAINING_LEFT;                   // :94-98
    if (state.game_type != GT_NONE) {                                     // :101-119
        uint32_t impossible[4]; gather_impossible_colors(state.tricks, state.n_tricks, state.game_type, impossible);
        for (int p = 0; p < 4; ++p)
            for (int c = 0; c < 5; ++c)
                if (((impossible[p] >> c) & 1) && assumed_hands[p].contains_card_of_color((Color)c, state.game_type)) return NC_ALREADY_DISCARDED_COLOR;
    }
    int wedding_player = -1;                                              // :121-142 (play order; the LAST wedding seat wins)
    for (int i = 0; i < 4; ++i) {
        int p = player_next(state.reservations_round.starting_player, i);
        if (assumed_reservations[p] == R_WEDDING) wedding_player = p;
    }
    if (wedding_player >= 0)
        for (int p = 0; p < 4; ++p) if (p != wedding_player && assumed_hands[p].contains(CQ)) return NC_CQ_BUT_OTHER_WEDDING;
    if (state.team_state.has_re_players() && state.game_type == GT_NORMAL) {   // :145-193
        bool played_q[4] = {false, false, false, false};
        for (int t = 0; t < state.n_tricks; ++t)
            for (int k = 0; k < state.tricks[t].len; ++k) if (state.tricks[t].cards[k] == CQ) played_q[state.tricks[t].player_at(k)] = true;
        for (int a = 0; a < state.announcements.n; ++a) {
            int ap = state.announcements.occ[a].player;
            bool in_hand = assumed_hands[ap].contains(CQ);
            if ((state.team_state.re_players >> ap) & 1) { if (!in_hand && !played_q[ap]) return NC_NO_CQ_BUT_RE; }
            else if (in_hand) return NC_CQ_BUT_KONTRA;
        }
    }
    int poa[4] = {0, 0, 0, 0};                                            // :197-228
    for (int t = 0; t < state.n_tricks; ++t)
        for (int k = 0; k < state.tricks[t].len; ++k) if (state.tricks[t].cards[k] == CQ) poa[state.tricks[t].player_at(k)] += 1;
    for (int p = 0; p < 4; ++p) { if (assumed_hands[p].contains_both(CQ)) poa[p] += 2; else if (assumed_hands[p].contains(CQ)) poa[p] += 1; }
    int visible[4]; get_visible_reservations(state.reservations_round, observing, visible);
    for (int p = 0; p < 4; ++p) {                                         // :230-300
        int vr = visible[p], ar = assumed_reservations[p];
        if (vr != VR_NONE_YET && ar == R_NONE) return NC_WRONG_RESERVATION;
        if (vr == VR_NONE_YET && ar != R_NONE) return NC_WRONG_RESERVATION;
        if (vr == VR_HEALTHY && ar != R_HEALTHY) return NC_WRONG_RESERVATION;
        if (vr == VR_WEDDING && ar != R_WEDDING) return NC_WRONG_RESERVATION;
        if (vr == VR_DIAMONDS_SOLO && ar != R_DIAMONDS_SOLO) return NC_WRONG_RESERVATION;
        if (vr == VR_HEARTS_SOLO && ar != R_HEARTS_SOLO) return NC_WRONG_RESERVATION;
        if (vr == VR_SPADES_SOLO && ar != R_SPADES_SOLO) return NC_WRONG_RESERVATION;
        if (vr == VR_CLUBS_SOLO && ar != R_CLUBS_SOLO) return NC_WRONG_RESERVATION;
        if (vr == VR_TRUMPLESS_SOLO && ar != R_TRUMPLESS_SOLO) return NC_WRONG_RESERVATION;
        if (vr == VR_JACKS_SOLO && ar != R_JACKS_SOLO) return NC_WRONG_RESERVATION;
        if (vr == VR_QUEENS_SOLO && ar != R_QUEENS_SOLO) return NC_WRONG_RESERVATION;
        if (ar == R_WEDDING && poa[p] < 2) return NC_WRONG_RESERVATION_CQ;
    }
    return NC_OK;
}

}  // namespace fdo
}  // namespace oracle
