// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// CPU restatement of the AlphaZero self-play driver around the env-step/encode hot path:
//   rs-doko-alpha-zero/src/alpha_zero/train/self_play.rs:19-207          self_play (ValueTarget::Default branch)
//   rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:43-172         FdoAzEnvState (AzEnvState impl)
//   rs-doko-alpha-zero/src/alpha_zero/utils/rot_arr.rs:15-38              RotArr::new_from_0 (value target rotation)
// The search slot (`mcts_search`, self_play.rs:106-116 — a network-guided tree search, out of scope) is a parameter: it receives
// the state and the epoch-filtered allowed set and returns (action, policy_target).  The stand-in used by tests and benches is
// `uniform_search`: one Philox draw over the allowed set (SITE_STEP) and the uniform distribution as policy target.
//
// Philox contract of the driver: game `unit`, turn t uses epoch = first_epoch + t; the keep draw is word 0 of SITE_KEEP mapped like
// rand's StandardUniform f32: (w >> 8) * 2^-24.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>
#include "encode.hpp"
#include "fdo.hpp"

namespace oracle {
namespace selfplay {

constexpr size_t MIN_EPOCH = 10;                                   // full_doko.rs:23
constexpr uint64_t CALL_ACTIONS = 0x1Full << 33;                   // AnnouncementReContra … AnnouncementBlack (action.rs:49-53)

// allowed_actions_by_action_index(is_secondary, epoch) as a bit mask (full_doko.rs:76-97); number_of_allowed_actions(epoch) (:99-116)
// is the popcount of the is_secondary = false mask.
inline uint64_t az_allowed(const fdo::State& s, bool is_secondary, size_t epoch) {
    uint64_t m = s.allowed_actions();
    if (is_secondary || epoch < MIN_EPOCH) m &= ~CALL_ACTIONS;
    return m;
}
inline void rewards(const fdo::State& s, float out[4]) {           // rewards_or_none (:54-68)
    for (int p = 0; p < 4; ++p) out[p] = (float)s.end_of_game_stats.player_points[p] / 8.0f;
}
inline float keep_draw(uint32_t word) { return (float)(word >> 8) * (1.0f / 16777216.0f); }

struct Row {
    int64_t state[311];
    float policy[39];
    float value[4];
    uint8_t player;
    uint8_t forced;
    uint16_t turn;
};
using SearchFn = std::function<int(const fdo::State&, uint64_t allowed, uint32_t turn, float policy_target[39])>;

// self_play (self_play.rs:56-207), ValueTarget::Default.  Returns the number of turns played.
inline uint32_t self_play(fdo::State s, size_t epoch, float keep_prob, uint64_t seed, uint64_t unit, uint32_t first_epoch,
                          const SearchFn& search, std::vector<Row>& rows) {
    const size_t first_row = rows.size();
    uint32_t turn = 0;
    while (s.current_phase != fdo::PH_FINISHED) {                                       // :70
        const uint64_t allowed = az_allowed(s, false, epoch);
        Row row{};
        row.turn = (uint16_t)turn;
        row.player = (uint8_t)(s.current_player < 0 ? 0 : s.current_player);            // current_player() (:46-48)
        int action;
        bool keep = true;
        if (__builtin_popcountll(allowed) == 1) {                                       // forced move (:76-101)
            action = __builtin_ctzll(allowed);
            row.forced = 1;
            row.policy[action] = 1.0f;
            keep = keep_draw(philox_word(seed, (uint32_t)unit, (uint32_t)(unit >> 32), first_epoch + turn, SITE_KEEP, 0)) < keep_prob;
            if (keep) fdo::encode_state_pi(s, row.state);
        } else {
            action = search(s, allowed, turn, row.policy);                              // :106-143
            fdo::encode_state_pi(s, row.state);                                         // :147-152
        }
        if (keep) rows.push_back(row);
        s.play_action(action);                                                          // take_action_by_action_index(a, false, epoch)
        ++turn;
    }
    float r[4];
    rewards(s, r);
    for (size_t i = first_row; i < rows.size(); ++i)                                    // RotArr::new_from_0(cp, rewards).extract() (:195-205)
        for (int j = 0; j < 4; ++j) rows[i].value[(4 - rows[i].player + j) % 4] = r[j];
    return turn;
}

// Stand-in search: uniform random action over the allowed set (MSB-first rank pick like FdoAllowedActions::random), uniform policy.
inline int uniform_search(uint64_t seed, uint64_t unit, uint32_t first_epoch, const fdo::State&, uint64_t allowed, uint32_t turn, float policy[39]) {
    const uint32_t n = (uint32_t)__builtin_popcountll(allowed);
    const uint32_t w = philox_word(seed, (uint32_t)unit, (uint32_t)(unit >> 32), first_epoch + turn, SITE_STEP, 0);
    const uint32_t idx = (uint32_t)(((uint64_t)w * n) >> 32);
    const uint64_t bit = select_by_rank(allowed, idx);
    for (int a = 0; a < 39; ++a) policy[a] = ((allowed >> a) & 1ull) ? 1.0f / (float)n : 0.0f;
    return __builtin_ctzll(bit);
}

}  // namespace selfplay
}  // namespace oracle
