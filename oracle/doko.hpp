// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// CPU restatement of the reference's SIMPLIFIED Doppelkopf engine `rs-doko`
// (normal game + wedding, no solos, no announcements, simple scoring).
#pragma once
#include <cstdint>
#include <stdexcept>
#include "bitflag.hpp"
#include "rng.hpp"

namespace oracle {
namespace doko {

// DoCard (rs-doko/src/card/cards.rs:7-33): same index layout as FdoCard.
enum Card : int {
    D9, D10, DJ, DQ, DK, DA, H9, H10, HJ, HQ, HK, HA, C9, C10, CJ, CQ, CK, CA, S9, S10, SJ, SQ, SK, SA
};
enum Color : int { TRUMP = 0, HEART = 1, SPADE = 2, CLUB = 3, COLOR_NONE = -1 };   // basic/color.rs
enum Phase : int { PH_RESERVATION = 0, PH_PLAYCARD = 1, PH_FINISHED = 2 };         // basic/phase.rs
enum Reservation : int { R_WEDDING = 0, R_HEALTHY = 1, R_NONE = -1 };              // reservation/reservation.rs:9-12
enum VisibleReservation : int { VR_WEDDING = 0, VR_HEALTHY = 1, VR_NOT_REVEALED = 2, VR_NONE = -1 };
enum ActionIndex : int { ACT_RES_HEALTHY = 24, ACT_RES_WEDDING = 25, ACT_COUNT = 26 };  // action/action.rs:7-38
enum TeamTag : int { TS_IN_RESERVATIONS = 0, TS_WEDDING_UNSOLVED = 1, TS_WEDDING_SOLVED = 2, TS_NO_WEDDING = 3 };

inline uint64_t card_bit(int c) { return 1ull << c; }

// card/card_color_masks.rs:3-29 (pinned by the binary literals at :36-41)
constexpr uint64_t TRUMP_MASK = 0x30C3BF, HEART_MASK = 0x000C40, SPADE_MASK = 0xCC0000, CLUB_MASK = 0x033000;

inline uint32_t card_to_eyes(int c) { static const uint32_t e[6] = {0, 10, 2, 3, 4, 11}; return e[c % 6]; }  // card/card_to_eyes.rs

inline Color card_to_color_in_normal_game(int c) {                       // card/card_to_color.rs:9-40
    static const char* t = "TTTTTT" "HTTTHH" "CCTTCC" "SSTTSS";
    switch (t[c]) { case 'T': return TRUMP; case 'H': return HEART; case 'S': return SPADE; default: return CLUB; }
}
inline int trump_to_rank(int c) {                                        // card/card_in_trick_logic.rs:24-44
    switch (c) {
        case D9: return 0; case DK: return 1; case D10: return 2; case DA: return 3;
        case DJ: return 4; case HJ: return 5; case SJ: return 6; case CJ: return 7;
        case DQ: return 8; case HQ: return 9; case SQ: return 10; case CQ: return 11; case H10: return 12;
    }
    throw std::runtime_error("This card is not a trump card");
}
inline bool is_greater_in_trick_in_normal_game(int cur, int prev, Color trick_color) {  // :58-106
    Color cc = card_to_color_in_normal_game(cur), pc = card_to_color_in_normal_game(prev);
    bool ct = cc == TRUMP, pt = pc == TRUMP;
    if (ct && !pt) return true;
    if (!ct && pt) return false;
    if (ct && pt) return trump_to_rank(prev) < trump_to_rank(cur);
    bool c_tc = cc == trick_color, p_tc = pc == trick_color;
    if (c_tc && !p_tc) return true;
    if (c_tc && p_tc) return card_to_eyes(cur) > card_to_eyes(prev);
    return false;
}

// hand/hand.rs:13-46,78 — NOTE: remove is A-first here (B-first in rs-full-doko).
inline bool hand_contains(uint64_t h, int c) { return (h & card_bit(c)) || (h & (card_bit(c) << 24)); }
inline bool hand_contains_both(uint64_t h, int c) { return (h & card_bit(c)) && (h & (card_bit(c) << 24)); }
inline uint64_t hand_add(uint64_t h, int c) { return hand_contains(h, c) ? (h | (card_bit(c) << 24)) : (h | card_bit(c)); }
inline uint64_t hand_remove(uint64_t h, int c) {
    if (h & card_bit(c)) return h & ~card_bit(c);
    if (h & (card_bit(c) << 24)) return h & ~(card_bit(c) << 24);
    throw std::runtime_error("Karte nicht in Hand");
}
inline uint32_t hand_len(uint64_t h) { return popcount64(h); }
// hand_to_vec (hand/hand.rs:135-179): ascending card id, doubles adjacent.
inline int hand_to_vec(uint64_t h, int* out) {
    int n = 0;
    for (int c = 0; c < 24; ++c) {
        if (hand_contains_both(h, c)) { out[n++] = c; out[n++] = c; }
        else if (hand_contains(h, c)) out[n++] = c;
    }
    return n;
}
// distribute_cards (hand/hand_random.rs:44-60)
inline void distribute_cards(Rng& rng, uint64_t hands[4]) {
    uint8_t cards[48];
    for (int i = 0; i < 48; ++i) cards[i] = (uint8_t)(i / 2);
    rng.shuffle48(cards);
    for (int i = 0; i < 4; ++i) { hands[i] = 0; for (int j = 0; j < 12; ++j) hands[i] = hand_add(hands[i], cards[i * 12 + j]); }
}

struct Trick {                                                           // trick/trick.rs:8-13
    int8_t cards[4] = {-1, -1, -1, -1};
    int8_t start_player = 0;
    bool present = false;
    bool is_completed() const { return cards[0] >= 0 && cards[1] >= 0 && cards[2] >= 0 && cards[3] >= 0; }
    Color color() const { return cards[0] < 0 ? COLOR_NONE : card_to_color_in_normal_game(cards[0]); }  // :27-29
    void play_card(int c) { for (int i = 0; i < 4; ++i) if (cards[i] < 0) { cards[i] = (int8_t)c; return; } throw std::runtime_error("trick full"); }
    int winner() const {                                                 // trick_winning_player_logic.rs:11-34
        Color tc = color();
        int wc = cards[0], wi = 0;
        for (int i = 0; i < 4; ++i) if (is_greater_in_trick_in_normal_game(cards[i], wc, tc)) { wc = cards[i]; wi = i; }
        return (wi + start_player) % 4;
    }
    uint32_t eyes() const { uint32_t e = 0; for (int i = 0; i < 4; ++i) if (cards[i] >= 0) e += card_to_eyes(cards[i]); return e; }  // trick_eyes.rs
};

struct ReservationRound {                                               // reservation/reservation_round.rs:8-13
    int8_t r[4] = {-1, -1, -1, -1};  // slot i = i-th reservation made, starting with start_player
    int8_t start_player = 0;
    bool is_completed() const { return r[0] >= 0 && r[1] >= 0 && r[2] >= 0 && r[3] >= 0; }
    void play(int res) { for (int i = 0; i < 4; ++i) if (r[i] < 0) { r[i] = (int8_t)res; return; } throw std::runtime_error("round full"); }
    int len() const { int n = 0; for (int i = 0; i < 4; ++i) n += r[i] >= 0; return n; }
};
// winning_player_in_reservation_round (reservation_winning_logic.rs:13-38): last wedding wins; -1 none
inline int wedding_player_of_round(const ReservationRound& rr) {
    int w = -1;
    for (int i = 0; i < 4; ++i) if (rr.r[i] == R_WEDDING) w = i;
    return w < 0 ? -1 : (w + rr.start_player) % 4;
}
// get_visible_reservations (visible_reservations_logic.rs:6-34).  QUIRK: walks
// reservation SLOTS but compares the slot index with the absolute observing player and stores by slot.
inline void get_visible_reservations(const ReservationRound& rr, int observing_player, int out[4]) {
    bool completed = rr.is_completed();
    for (int i = 0; i < 4; ++i) {
        if (rr.r[i] == R_HEALTHY) out[i] = VR_HEALTHY;
        else if (rr.r[i] == R_WEDDING) out[i] = (completed || i == observing_player) ? VR_WEDDING : VR_NOT_REVEALED;
        else out[i] = VR_NONE;
    }
}

struct TeamState { int tag = TS_IN_RESERVATIONS; int wedding_player = -1; int solved_trick_index = 0; uint32_t re_players = 0;
                   bool is_final() const { return tag == TS_WEDDING_SOLVED || tag == TS_NO_WEDDING; } };
// resolve_team_state (teams/team_logic.rs:69-143); reservation result: wedding_player (-1 = NoReservation)
inline TeamState resolve_team_state(int wedding_player, const Trick tricks[12], const uint64_t hands[4]) {
    TeamState ts;
    if (wedding_player < 0) {
        for (int i = 0; i < 4; ++i) if (hand_contains(hands[i], CQ)) ts.re_players |= 1u << i;
        ts.tag = TS_NO_WEDDING; return ts;
    }
    int partner = -1, partner_trick = -1, completed = 0;
    for (int i = 0; i < 3; ++i) {
        if (tricks[i].present) {
            if (!tricks[i].is_completed()) break;
            completed += 1;
            int w = tricks[i].winner();     // recomputed each time (quirk A.9 (3))
            if (w != wedding_player) { partner = w; partner_trick = i; break; }
        }
    }
    ts.wedding_player = wedding_player;
    if (partner >= 0) { ts.tag = TS_WEDDING_SOLVED; ts.solved_trick_index = partner_trick; ts.re_players = (1u << wedding_player) | (1u << partner); }
    else if (completed == 3) { ts.tag = TS_WEDDING_SOLVED; ts.solved_trick_index = 2; ts.re_players = 1u << wedding_player; }
    else ts.tag = TS_WEDDING_UNSOLVED;
    return ts;
}

struct EndOfGameStats { bool present = false; int winning_team = 0; uint32_t re_players = 0; bool is_solo = false;
                        uint32_t player_eyes[4] = {0,0,0,0}; uint32_t re_eyes = 0, kontra_eyes = 0; int re_points = 0, kontra_points = 0;
                        int player_points[4] = {0,0,0,0}; };
// calculate_end_of_game_stats (stats/stats.rs:25-135)
inline EndOfGameStats calculate_end_of_game_stats(uint32_t re_players, const uint32_t eyes[4], const uint32_t ntricks[4]) {
    uint32_t re_tricks = 0, ko_tricks = 0, re_eyes = 0, ko_eyes = 0;
    for (int p = 0; p < 4; ++p) {
        if ((re_players >> p) & 1) { re_eyes += eyes[p]; re_tricks += ntricks[p]; } else { ko_eyes += eyes[p]; ko_tricks += ntricks[p]; }
    }
    bool re_wins = re_eyes > ko_eyes;                                    // 120:120 → Kontra
    bool is_solo = __builtin_popcount(re_players) == 1;
    uint32_t we = re_wins ? re_eyes : ko_eyes, wt = re_wins ? re_tricks : ko_tricks;
    int pts = 0;
    if (we >= 120) pts++;
    if (we >= 150) pts++;
    if (we >= 180) pts++;
    if (we >= 210) pts++;
    if (wt == 12) pts++;
    int re_points = re_wins ? pts : -pts, ko_points = re_wins ? -pts : pts;
    if (is_solo) re_points *= 3;
    EndOfGameStats s; s.present = true; s.winning_team = re_wins ? 0 : 1; s.re_players = re_players; s.is_solo = is_solo;
    s.re_eyes = re_eyes; s.kontra_eyes = ko_eyes; s.re_points = re_points; s.kontra_points = ko_points;
    for (int p = 0; p < 4; ++p) { s.player_eyes[p] = eyes[p]; s.player_points[p] = ((re_players >> p) & 1) ? re_points : ko_points; }
    return s;
}

// calculate_allowed_actions_in_normal_game (action/allowed_actions.rs:131-198)
inline uint64_t calculate_allowed_actions_in_normal_game(int phase, Color trick_color, uint64_t hand) {
    switch (phase) {
        case PH_RESERVATION: {
            uint64_t a = 1ull << ACT_RES_HEALTHY;
            if (hand_contains_both(hand, CQ)) a |= 1ull << ACT_RES_WEDDING;
            return a;
        }
        case PH_PLAYCARD: {
            uint64_t single = (hand | (hand >> 24)) & 0xFFFFFFull;
            if (trick_color == COLOR_NONE) return single;
            uint64_t mask = trick_color == TRUMP ? TRUMP_MASK : trick_color == HEART ? HEART_MASK : trick_color == SPADE ? SPADE_MASK : CLUB_MASK;
            if (single & mask) return single & mask;
            return single;
        }
        default: return 0;                                               // Finished
    }
}

struct State {                                                           // state/state.rs:29-76
    ReservationRound reservations_round;
    Trick tricks[12];
    uint64_t hands[4] = {0, 0, 0, 0};
    int current_player = -1;
    int current_phase = PH_RESERVATION;
    int current_trick_index = 0;
    bool has_reservation_result = false;
    int wedding_player_result = -1;       // DoReservationResult: -1 = NoReservation, else Wedding(player)
    uint32_t player_eyes[4] = {0, 0, 0, 0};
    uint32_t player_num_tricks[4] = {0, 0, 0, 0};
    TeamState team_state;
    EndOfGameStats end_of_game_stats;
    uint32_t n_play_actions = 0;

    static State new_game_from_hand_and_start_player(const uint64_t hands[4], int start) {  // :115-157
        State s; s.reservations_round.start_player = (int8_t)start; for (int p = 0; p < 4; ++p) s.hands[p] = hands[p];
        s.current_player = start; return s;
    }
    static State new_game(Rng& rng) {                                    // :159-168
        int start = (int)rng.start_player();
        uint64_t hands[4]; distribute_cards(rng, hands);
        return new_game_from_hand_and_start_player(hands, start);
    }
    void play_action(int action) {                                       // :189-309
        if (current_phase == PH_FINISHED) throw std::runtime_error("finished");
        n_play_actions++;
        int cp = current_player;
        if (action < 24) {
            hands[cp] = hand_remove(hands[cp], action);
            Trick& t = tricks[current_trick_index];
            t.play_card(action);
            if (t.is_completed()) {
                current_trick_index += 1;
                int wp = t.winner();
                player_eyes[wp] += t.eyes();
                player_num_tricks[wp] += 1;
                if (!team_state.is_final()) team_state = resolve_team_state(wedding_player_result, tricks, hands);
                if (current_trick_index == 12) {
                    current_phase = PH_FINISHED; current_player = -1;
                    if (!team_state.is_final()) throw std::runtime_error("teams unresolved at end");
                    end_of_game_stats = calculate_end_of_game_stats(team_state.re_players, player_eyes, player_num_tricks);
                    return;
                }
                tricks[current_trick_index] = Trick(); tricks[current_trick_index].present = true;
                tricks[current_trick_index].start_player = (int8_t)wp;
                current_player = wp;
                return;
            }
            current_player = (cp + 1) % 4;
        } else {
            int res = action == ACT_RES_HEALTHY ? R_HEALTHY : R_WEDDING;
            reservations_round.play(res);
            int np = (cp + 1) % 4;
            current_player = np;
            if (reservations_round.is_completed()) {
                current_phase = PH_PLAYCARD;
                wedding_player_result = wedding_player_of_round(reservations_round);
                has_reservation_result = true;
                current_trick_index = 0;
                tricks[0] = Trick(); tricks[0].present = true; tricks[0].start_player = (int8_t)np;
                if (!team_state.is_final()) team_state = resolve_team_state(wedding_player_result, tricks, hands);
            }
        }
    }
    Color current_trick_color() const {                                  // :325-329 / :340-346
        if (current_trick_index < 12 && tricks[current_trick_index].present) return tricks[current_trick_index].color();
        return COLOR_NONE;
    }
    uint64_t allowed_actions() const {
        int obs = current_player < 0 ? 0 : current_player;
        return calculate_allowed_actions_in_normal_game(current_phase, current_trick_color(), hands[obs]);
    }
    int card_index() const { int ci = 0; for (int t = 0; t < 12; ++t) for (int k = 0; k < 4; ++k) ci += tricks[t].present && tricks[t].cards[k] >= 0; return ci; }
    // Chained card draws (rng.hpp PhiloxStream::chain; see fdo::State::card_chain_mul).  rs-doko enforces the colour in every trick.
    uint32_t card_chain_mul() const {
        if (current_phase != PH_PLAYCARD || current_trick_index >= 12 || !tricks[current_trick_index].present) return 1;
        const Trick& t = tricks[current_trick_index];
        uint32_t mul = 1;
        for (int k = 0; k < 4 && t.cards[k] >= 0; ++k) {
            const uint64_t h = hand_add(hands[(t.start_player + k) % 4], t.cards[k]);
            const Color col = k > 0 ? card_to_color_in_normal_game(t.cards[0]) : COLOR_NONE;
            mul *= (uint32_t)__builtin_popcountll(calculate_allowed_actions_in_normal_game(PH_PLAYCARD, col, h));
        }
        return mul;
    }
    void position_streams(Rng& rng) const {
        rng.set_card_position((uint32_t)card_index(), card_chain_mul());
        rng.set_ordinal(SITE_RESERVATION, (uint32_t)reservations_round.len());
    }
    bool random_action_for_current_player(Rng& rng, int* action_out = nullptr) {  // :315-334
        if (current_phase == PH_FINISHED) return true;
        uint64_t allowed = allowed_actions();
        uint64_t bit = random_single(allowed, rng, current_phase == PH_RESERVATION ? SITE_RESERVATION : SITE_CARD);
        int a = __builtin_ctzll(bit);
        if (action_out) *action_out = a;
        play_action(a);
        return false;
    }
};

}  // namespace doko
}  // namespace oracle
