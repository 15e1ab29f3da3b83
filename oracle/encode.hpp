// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// CPU restatement of the reference's observation encoders:
//   rs-doko-embeddings/src/encode_state.rs:84-317            (110 / 114 tokens, rs-doko)
//   rs-doko-networks/src/full_doko/var1/encode_pi.rs:27-216   (311 x i64, rs-full-doko, perfect information)
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>
#include "doko.hpp"
#include "fdo.hpp"

namespace oracle {

// ---------------------------------------------------------------------------------------------
// rs-doko: encode_state / encode_state_with_reservations
// ---------------------------------------------------------------------------------------------
namespace doko {

inline int get_relative_player_index(int current, int target) { return (target + 4 - current) % 4; }  // encode_state.rs:10-15
inline int card_to_rank_in_normal_game(int c) {                          // encode_state.rs:17-50
    switch (c) {
        case H9: return 0; case HK: return 1; case HA: return 2;
        case S9: return 3; case SK: return 4; case S10: return 5; case SA: return 6;
        case C9: return 7; case CK: return 8; case C10: return 9; case CA: return 10;
        case D9: return 11; case DK: return 12; case D10: return 13; case DA: return 14;
        case DJ: return 15; case HJ: return 16; case SJ: return 17; case CJ: return 18;
        case DQ: return 19; case HQ: return 20; case SQ: return 21; case CQ: return 22;
        case H10: return 23;
    }
    return -1;
}
inline int64_t encode_card_or_none(int c) { return c < 0 ? 0 : c + 1; }   // encode_state.rs:52-82

// Writes 110 (with_reservations=false) or 114 tokens.
inline int encode_state(const State& s, int64_t* out, bool with_reservations) {
    int n = 0;
    int current_player = s.current_player < 0 ? 0 : s.current_player;    // :92-94
    out[n++] = s.current_phase;                                           // :90
    out[n++] = get_relative_player_index(current_player, s.reservations_round.start_player) + 1;  // :96-98
    for (int i = 0; i < 12; ++i)                                          // :100-104,151-161
        out[n++] = s.tricks[i].present ? get_relative_player_index(current_player, s.tricks[i].start_player) + 1 : 0;
    int64_t trick_cards[48]; for (int i = 0; i < 48; ++i) trick_cards[i] = 0;
    int ci = 0;
    for (int t = 0; t < 12; ++t) {                                        // :110-129
        if (!s.tricks[t].present) break;
        for (int k = 0; k < 4; ++k) { if (s.tricks[t].cards[k] < 0) break; trick_cards[ci++] = encode_card_or_none(s.tricks[t].cards[k]); }
    }
    for (int i = 0; i < 48; ++i) out[n++] = trick_cards[i];
    int64_t hand_cards[48]; for (int i = 0; i < 48; ++i) hand_cards[i] = 0;
    for (int i = 0; i < 4; ++i) {                                         // :133-147
        int rel = get_relative_player_index(current_player, i);
        int v[24]; int m = hand_to_vec(s.hands[i], v);
        // sort_by_key(|c| -rank): STABLE, descending rank
        std::stable_sort(v, v + m, [](int a, int b) { return -card_to_rank_in_normal_game(a) < -card_to_rank_in_normal_game(b); });
        for (int j = 0; j < 12; ++j) hand_cards[rel * 12 + j] = j < m ? encode_card_or_none(v[j]) : 0;
    }
    for (int i = 0; i < 48; ++i) out[n++] = hand_cards[i];
    if (with_reservations) {                                              // :253-257,266-275
        int vr[4]; get_visible_reservations(s.reservations_round, current_player, vr);
        for (int i = 0; i < 4; ++i)
            out[n++] = vr[i] == VR_NONE ? 0 : (vr[i] == VR_NOT_REVEALED ? 1 : (vr[i] == VR_HEALTHY ? 2 : 3));
    }
    return n;
}

}  // namespace doko

// ---------------------------------------------------------------------------------------------
// rs-full-doko: encode_state_pi → [i64; 311]
// ---------------------------------------------------------------------------------------------
namespace fdo {

inline int64_t player_index_relative(int player, int relative_to) { return ((player - relative_to) % 4 + 4) % 4; }
inline int64_t encode_player_or_none(int player, int relative_to) {      // var1/player.rs:19-30
    return player < 0 ? 0 : player_index_relative(player, relative_to) + 1;
}
inline int64_t encode_position_or_unknown_hand(int player, int relative_to) {  // var2/encode_position_or_unknown.rs:14-24
    return player < 0 ? 0 : player_index_relative(player, relative_to) + 1 + 52;
}
inline int64_t encode_position_or_unknown_int(int position) {            // :26-33
    if (position < 0 || position >= 52) throw std::runtime_error("position out of range");
    return position + 1;
}
inline int64_t encode_pi_announcement(int ann) {                          // var2/encode_reservation_or_card_or_none.rs:6-22
    switch (ann) {
        case A_NONE: return 37; case A_RE_CONTRA: return 38; case A_COUNTER_RE_CONTRA: return 38;
        case A_NO90: return 39; case A_NO60: return 40; case A_NO30: return 41; case A_BLACK: return 42;
    }
    throw std::runtime_error("should not happen");
}
inline int64_t encode_pi_reservation(int res) {                           // :24-44
    switch (res) {
        case R_NONE: return 36; case R_HEALTHY: return 25; case R_WEDDING: return 26; case R_DIAMONDS_SOLO: return 27;
        case R_HEARTS_SOLO: return 28; case R_SPADES_SOLO: return 29; case R_CLUBS_SOLO: return 30;
        case R_QUEENS_SOLO: return 31; case R_JACKS_SOLO: return 32; case R_TRUMPLESS_SOLO: return 33;
    }
    throw std::runtime_error("bad reservation");
}
inline int64_t encode_card_token(int card) {                              // :70-107
    switch (card) {
        case CARD_NONE: return 0;
        case H10: return 1; case CQ: return 2; case SQ: return 3; case HQ: return 4; case DQ: return 5;
        case CJ: return 6; case SJ: return 7; case HJ: return 8; case DJ: return 9;
        case DA: return 10; case D10: return 11; case DK: return 12; case D9: return 13;
        case CA: return 14; case C10: return 15; case CK: return 16; case C9: return 17;
        case SA: return 18; case S10: return 19; case SK: return 20; case S9: return 21;
        case HA: return 22; case HK: return 23; case H9: return 24;
    }
    throw std::runtime_error("bad card");
}
inline int64_t encode_subposition_card(int sub) { return sub < 0 ? 0 : (sub == 0 ? 11 : 12); }  // var2/encode_subposition.rs:3-12
inline int64_t encode_subposition_pos(int pos) { return pos < 0 ? 0 : pos + 1; }                // :14-21
inline int64_t encode_announcement_team(int team) { return team < 0 ? 0 : (team == 0 ? 1 : 2); }  // var2/encode_annoucnement_team.rs
inline int64_t encode_phase(int phase) { return phase; }                                         // var1/phase.rs:9-18

// encode_state_pi (var1/encode_pi.rs:27-216) with obs = state.observation_for_current_player()
// (as called by FdoAzEnvState::encode_into_memory, rs-doko-alpha-zero/.../full_doko.rs:71-78).
inline void encode_state_pi(const State& s, int64_t out[311]) {
    int current_player = s.current_player < 0 ? 0 : s.current_player;    // :31-33
    int64_t tok[62], pos[62], ply[62], sub[62], team[62];
    int n = 0, card_index = 0;
    auto push = [&](int64_t p, int64_t t, int64_t pl, int64_t su, int64_t te) {
        if (n >= 62) throw std::runtime_error("encode_state_pi: more than 62 slots");
        pos[n] = p; tok[n] = t; ply[n] = pl; sub[n] = su; team[n] = te; n++;
    };
    // 4 reservation slots in PLAY order from the round's starting player (:43-81)
    for (int i = 0; i < 4; ++i) {
        int player = player_next(s.reservations_round.starting_player, i);
        if (i >= s.reservations_round.len) push(0, encode_pi_reservation(R_NONE), 0, 0, 0);
        else push(encode_position_or_unknown_int(card_index), encode_pi_reservation(s.reservations_round.r[i]),
                  encode_player_or_none(player, current_player), 0, 0);
        card_index += 1;
    }
    // played cards, trick by trick (:83-99)
    for (int t = 0; t < s.n_tricks; ++t)
        for (int k = 0; k < s.tricks[t].len; ++k) {
            push(encode_position_or_unknown_int(card_index), encode_card_token(s.tricks[t].cards[k]),
                 encode_player_or_none(s.tricks[t].player_at(k), current_player), 0, 0);
            card_index += 1;
        }
    // all four (real) hands, seats starting at the current player, cards in ascending-BIT order (:102-121)
    for (int i = 0; i < 4; ++i) {
        int player = player_next(current_player, i);
        Hand already;
        int cards[48]; int m = s.hands[player].iter(cards);
        for (int j = 0; j < m; ++j) {
            push(encode_position_or_unknown_hand(player, current_player), encode_card_token(cards[j]), 0,
                 encode_subposition_card(already.contains(cards[j]) ? 1 : 0), 0);
            already.add(cards[j]);
        }
    }
    // announcements in order of occurrence (:139-165); position = raw card_index (quirk A.9 (7))
    int subposition_index = 0; int64_t last_position = -1;
    for (int a = 0; a < s.announcements.n; ++a) {
        const AnnouncementOccurrence& o = s.announcements.occ[a];
        if ((int64_t)o.card_index != last_position) { last_position = o.card_index; subposition_index = 0; }
        if (!s.team_state.has_re_players()) throw std::runtime_error("not possible");
        int tm = ((s.team_state.re_players >> o.player) & 1) ? 0 : 1;
        push(encode_position_or_unknown_int(o.card_index), encode_pi_announcement(o.announcement),
             encode_player_or_none(o.player, current_player), encode_subposition_pos(subposition_index), encode_announcement_team(tm));
        subposition_index += 1;
    }
    if (s.announcements.n > 10) throw std::runtime_error("10 - announcements.len() underflows");      // :167
    for (int i = 0; i < 10 - s.announcements.n; ++i) push(0, encode_pi_announcement(A_NONE), 0, 0, 0);  // :167-179
    if (n != 62) throw std::runtime_error("encode_state_pi: slot count != 62");                        // :184-188 try_into
    // concat: tokens, positions, players, subpositions, teams, phase (:208-215)
    for (int i = 0; i < 62; ++i) { out[i] = tok[i]; out[62 + i] = pos[i]; out[124 + i] = ply[i]; out[186 + i] = sub[i]; out[248 + i] = team[i]; }
    out[310] = encode_phase(s.current_phase);
}


// encode_reservation_or_card_or_reservation (var2/encode_reservation_or_card_or_none.rs:48-68): FdoVisibleReservation → token
inline int64_t encode_visible_reservation_token(int vr) {
    switch (vr) {
        case VR_HEALTHY: return 25; case VR_WEDDING: return 26; case VR_DIAMONDS_SOLO: return 27; case VR_HEARTS_SOLO: return 28;
        case VR_SPADES_SOLO: return 29; case VR_CLUBS_SOLO: return 30; case VR_QUEENS_SOLO: return 31; case VR_JACKS_SOLO: return 32;
        case VR_TRUMPLESS_SOLO: return 33; case VR_NOT_REVEALED: return 34; case VR_NONE_YET: return 35;
    }
    throw std::runtime_error("bad visible reservation");
}
// reservation_to_visible_reservation (var1/encode_ipi.rs:29-46)
inline int reservation_to_visible(int r) {
    switch (r) {
        case R_NONE: return VR_NONE_YET; case R_HEALTHY: return VR_HEALTHY; case R_WEDDING: return VR_WEDDING;
        case R_DIAMONDS_SOLO: return VR_DIAMONDS_SOLO; case R_HEARTS_SOLO: return VR_HEARTS_SOLO; case R_SPADES_SOLO: return VR_SPADES_SOLO;
        case R_CLUBS_SOLO: return VR_CLUBS_SOLO; case R_QUEENS_SOLO: return VR_QUEENS_SOLO; case R_JACKS_SOLO: return VR_JACKS_SOLO;
        case R_TRUMPLESS_SOLO: return VR_TRUMPLESS_SOLO;
    }
    throw std::runtime_error("bad reservation");
}

// encode_state_ipi (var1/encode_ipi.rs:48-306) with obs = state.observation_for_current_player(): the imperfect-information layout of
// the autoregressive hand predictor.  assumed_hands / assumed_reservations are indexed by ABSOLUTE seat (PlayerZeroOrientedArr);
// assumed_reservations[p] = R_NONE for "no guess yet".
inline void encode_state_ipi(const State& s, const Hand assumed_hands[4], const int assumed_reservations[4], int next_player_to_chose_card_for,
                             int64_t out[311]) {
    const int current_player = s.current_player < 0 ? 0 : s.current_player;                      // :56-58
    int64_t tok[62], pos[62], ply[62], sub[62], team[62];
    int n = 0, index = 0;
    auto push = [&](int64_t p, int64_t t, int64_t pl, int64_t su, int64_t te) {
        if (n >= 62) throw std::runtime_error("encode_state_ipi: more than 62 slots");
        pos[n] = p; tok[n] = t; ply[n] = pl; sub[n] = su; team[n] = te; n++;
    };
    int visible[4];
    get_visible_reservations(s.reservations_round, current_player, visible);
    for (int i = 0; i < 4; ++i) {                                                                // visible reservations in PLAY order (:68-91)
        int player = player_next(s.reservations_round.starting_player, i);
        int v = visible[player];
        if (v == VR_NOT_REVEALED && assumed_reservations[player] != R_NONE) v = reservation_to_visible(assumed_reservations[player]);
        push(encode_position_or_unknown_int(index), encode_visible_reservation_token(v), encode_player_or_none(player, current_player), 0, 0);
        index += 1;
    }
    for (int t = 0; t < s.n_tricks; ++t)                                                         // played cards (:93-110)
        for (int k = 0; k < s.tricks[t].len; ++k) {
            push(encode_position_or_unknown_int(index), encode_card_token(s.tricks[t].cards[k]),
                 encode_player_or_none(s.tricks[t].player_at(k), current_player), 0, 0);
            index += 1;
        }
    {                                                                                            // the observer's own hand (:113-130)
        Hand already;
        int cards[48]; int m = s.hands[current_player].iter(cards);
        for (int j = 0; j < m; ++j) {
            push(encode_position_or_unknown_hand(current_player, current_player), encode_card_token(cards[j]), 0,
                 encode_subposition_card(already.contains(cards[j]) ? 1 : 0), 0);
            already.add(cards[j]);
        }
    }
    for (int i = 1; i < 4; ++i) {                                                                // the other seats from the observer on (:132-174)
        int player = player_next(current_player, i);
        Hand already;
        int cards[48]; int m = assumed_hands[player].iter(cards);
        for (int j = 0; j < m; ++j) {
            push(encode_position_or_unknown_hand(player, current_player), encode_card_token(cards[j]), 0,
                 encode_subposition_card(already.contains(cards[j]) ? 1 : 0), 0);
            already.add(cards[j]);
        }
        if (s.hands[player].len() < assumed_hands[player].len()) throw std::runtime_error("hand.len() - assumed.len() underflows");   // :158
        uint32_t diff = s.hands[player].len() - assumed_hands[player].len();
        for (uint32_t j = 0; j < diff; ++j)                                                      // still unknown cards
            push(encode_position_or_unknown_hand(player, current_player), encode_card_token(CARD_NONE), 0, 0, 0);
    }
    int subposition_index = 0; int64_t last_position = -1;                                       // announcements (:191-215), as in encode_state_pi
    for (int a = 0; a < s.announcements.n; ++a) {
        const AnnouncementOccurrence& o = s.announcements.occ[a];
        if ((int64_t)o.card_index != last_position) { last_position = o.card_index; subposition_index = 0; }
        if (!s.team_state.has_re_players()) throw std::runtime_error("not possible");
        int tm = ((s.team_state.re_players >> o.player) & 1) ? 0 : 1;
        push(encode_position_or_unknown_int(o.card_index), encode_pi_announcement(o.announcement),
             encode_player_or_none(o.player, current_player), encode_subposition_pos(subposition_index), encode_announcement_team(tm));
        subposition_index += 1;
    }
    if (s.announcements.n > 10) throw std::runtime_error("10 - announcements.len() underflows");      // :217
    for (int i = 0; i < 10 - s.announcements.n; ++i) push(0, encode_pi_announcement(A_NONE), 0, 0, 0);
    if (n != 62) throw std::runtime_error("encode_state_ipi: slot count != 62");                        // try_into().unwrap() (:238-242)
    for (int i = 0; i < 62; ++i) { out[i] = tok[i]; out[62 + i] = pos[i]; out[124 + i] = ply[i]; out[186 + i] = sub[i]; out[248 + i] = team[i]; }
    out[310] = encode_player_or_none(next_player_to_chose_card_for, current_player);                   // :232, :298-305
}

}  // namespace fdo
}  // namespace oracle
