// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// CPU restatement of rs-doko-assignment (the Vec-based hidden-hand sampler of the simplified engine):
//   rs-doko-assignment/src/assignment.rs:13-581.  Vectors keep the reference's element ORDER (remaining cards in hand_to_vec
//   order ♦,♥,♣,♠ with doubles adjacent; candidate lists in all_cards_two_times order ♦,♥,♠,♣), because `position()`/`remove()`
//   and the random index depend on it.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>
#include "doko.hpp"

namespace oracle {
namespace doko {

inline uint64_t calc_remaining_cards(const Trick tricks[12], uint64_t own_hand) {       // assignment.rs:13-41
    uint64_t available = 0xFFFFFFFFFFFFull;
    for (int i = 0; i < 12; ++i) {
        if (!tricks[i].present) break;
        for (int j = 0; j < 4; ++j) { if (tricks[i].cards[j] < 0) break; available = hand_remove(available, tricks[i].cards[j]); }
    }
    int v[24]; int n = hand_to_vec(own_hand, v);
    for (int i = 0; i < n; ++i) available = hand_remove(available, v[i]);
    return available;
}
inline void remove_one_of(std::vector<int>& v, int card) {                             // :62-76
    auto it = std::find(v.begin(), v.end(), card);
    if (it != v.end()) v.erase(it);
}
inline void remove_all_of(std::vector<int>& v, const std::vector<int>& cards) {         // :43-60
    for (int c : cards) for (int i = 0; i < 2; ++i) remove_one_of(v, c);
}
inline void player_allowed_to_have(int player_marriage, const Trick tricks[12], uint64_t own_hand, int observing_player,
                                   std::vector<int> out[4]) {                           // :78-282
    const std::vector<int> trump = {D9, D10, DJ, DQ, DK, DA, DJ, HJ, SJ, CJ, DQ, HQ, SQ, CQ, H10};
    const std::vector<int> heart = {H9, HK, HA}, spade = {S9, S10, SK, SA}, club = {C9, C10, CK, CA};
    std::vector<int> all;                                                             // ♦,♥,♠,♣ (!), doubles adjacent (:131-183)
    for (int c : {D9, D10, DJ, DQ, DK, DA, H9, H10, HJ, HQ, HK, HA, S9, S10, SJ, SQ, SK, SA, C9, C10, CJ, CQ, CK, CA}) { all.push_back(c); all.push_back(c); }
    for (int i = 0; i < 4; ++i) out[i] = all;
    { int v[24]; int n = hand_to_vec(own_hand, v); out[observing_player].assign(v, v + n); }
    for (int j = 0; j < 12; ++j) {
        if (!tricks[j].present) break;
        Color tc = tricks[j].color();
        if (tc == COLOR_NONE) break;
        int cur = tricks[j].start_player;
        for (int k = 0; k < 4; ++k) {
            int card = tricks[j].cards[k];
            if (card < 0) break;
            for (int i = 0; i < 4; ++i) remove_one_of(out[i], card);
            if (card_to_color_in_normal_game(card) != tc)
                remove_all_of(out[cur], tc == TRUMP ? trump : tc == HEART ? heart : tc == SPADE ? spade : club);
            cur = (cur + 1) % 4;
        }
    }
    if (player_marriage >= 0)
        for (int i = 0; i < 4; ++i) if (i != player_marriage) remove_all_of(out[i], {CQ});
    int v[24]; int n = hand_to_vec(own_hand, v);
    for (int i = 0; i < 4; ++i) for (int q = 0; q < n; ++q) remove_one_of(out[i], v[q]);
}
struct AssignState {
    std::vector<int> remaining;
    std::vector<int> hands[4];
    size_t len[4];
    std::vector<int> allowed[4];
};
inline void distribute_card(AssignState& a, int player, int card) {                     // :284-316
    a.hands[player].push_back(card);
    a.len[player] -= 1;
    for (int i = 0; i < 4; ++i) remove_one_of(a.allowed[i], card);
    remove_one_of(a.remaining, card);
    for (int i = 0; i < 4; ++i) if (a.len[i] == 0) remove_all_of(a.allowed[i], a.remaining);
}
inline std::vector<int> players_for(const AssignState& a, int card) {                   // :318-334
    std::vector<int> p;
    for (int i = 0; i < 4; ++i)
        if (std::find(a.allowed[i].begin(), a.allowed[i].end(), card) != a.allowed[i].end() && a.len[i] > 0) p.push_back(i);
    return p;
}
inline bool distribute_single_cards(AssignState& a) {                                   // :336-377
    for (int card : a.remaining) {
        std::vector<int> p = players_for(a, card);
        if (p.size() == 1) { distribute_card(a, p[0], card); return true; }
    }
    return false;
}
inline bool distribute_exactly_as_per_hand(AssignState& a) {                            // :379-417
    for (int i = 0; i < 4; ++i) {
        if (a.len[i] == 0) continue;
        if (a.len[i] == a.allowed[i].size()) {
            std::vector<int> copy = a.allowed[i];
            for (int card : copy) distribute_card(a, i, card);
            return true;
        }
    }
    return false;
}
// returns 0 = nothing left, 1 = distributed, 2 = dead end (`.choose(rng).unwrap()` on an empty list would panic, :440-445)
inline int distribute_single_card_randomly(AssignState& a, Rng& rng) {                  // :419-456
    if (a.remaining.empty()) return 0;
    int card = a.remaining[rng.below(SITE_ASSIGN, (uint32_t)a.remaining.size())];
    std::vector<int> p = players_for(a, card);
    if (p.empty()) return 2;
    int player = p[rng.below_chained(SITE_ASSIGN, (uint32_t)p.size())];   // the seat comes from the card's word (chained draw, rng.hpp)
    distribute_card(a, player, card);
    return 1;
}
// sample_assignment (:493-581).  hands_out as bitboards (hand_from_vec); returns 0 ok / 1 dead end.
inline int sample_assignment(int player_marriage, const Trick tricks[12], uint64_t own_hand, const size_t lens[4], int observing_player,
                             Rng& rng, uint64_t hands_out[4]) {
    AssignState a;
    uint64_t rem = calc_remaining_cards(tricks, own_hand);
    { int v[48]; int n = hand_to_vec(rem, v); a.remaining.assign(v, v + n); }
    player_allowed_to_have(player_marriage, tricks, own_hand, observing_player, a.allowed);
    for (int i = 0; i < 4; ++i) a.len[i] = lens[i];
    int status = 0;
    for (;;) {
        if (distribute_single_cards(a)) continue;
        if (distribute_exactly_as_per_hand(a)) continue;
        int r = distribute_single_card_randomly(a, rng);
        if (r == 1) continue;
        if (r == 2) status = 1;
        break;
    }
    for (int i = 0; i < 4; ++i) {
        uint64_t h = 0;
        for (int c : a.hands[i]) h = hand_add(h, c);
        hands_out[i] = h;
    }
    hands_out[observing_player] = 0;
    { int v[24]; int n = hand_to_vec(own_hand, v); uint64_t h = 0; for (int i = 0; i < n; ++i) h = hand_add(h, v[i]); hands_out[observing_player] = h; }
    return status;
}
// sample_assignment_full (:458-491)
inline int sample_assignment_full(const State& s, Rng& rng, uint64_t hands_out[4]) {
    size_t lens[4];
    for (int i = 0; i < 4; ++i) lens[i] = hand_len(s.hands[i]);
    int obs = s.current_player < 0 ? 0 : s.current_player;
    int marriage = (s.team_state.tag == TS_WEDDING_UNSOLVED || s.team_state.tag == TS_WEDDING_SOLVED) ? s.team_state.wedding_player : -1;
    return sample_assignment(marriage, s.tricks, s.hands[obs], lens, obs, rng, hands_out);
}

}  // namespace doko
}  // namespace oracle
