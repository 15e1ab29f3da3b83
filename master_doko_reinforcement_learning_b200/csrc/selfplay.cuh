// selfplay.cuh — AlphaZero self-play driver around the env-step/encode hot path, lock-step over a batch of games.
//   self_play, ValueTarget::Default (rs-doko-alpha-zero/src/alpha_zero/train/self_play.rs:19-207)
//   FdoAzEnvState (rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:43-172)
// One TURN of all games = two kernels + the caller's search in between:
//   sp_begin   one pass over the states.  Per game: terminal? epoch-filtered allowed set, forced move?, keep-experience draw; a
//              decoupled look-back scan of the per-block row counts behind the rows already recorded gives deterministic row numbers
//              (turn-major, game order inside a turn: the order in which a sequential driver over the games would push them);
//              encode_state_pi of every kept game goes STRAIGHT INTO its row of the experience buffer (the network batcher reads the
//              rows where they lie: no Vec<i64> copy, no host round trip), with player / game / one-hot policy of forced moves
//   [search: caller fills policy[n][39] and action[n] for the games that are not forced]
//   sp_apply   policy target → row, take_action_by_action_index(action, false, epoch)
// and at the end sp_finalize rotates the final rewards into every row's value target (RotArr::new_from_0, rot_arr.rs:31-38).
#pragma once
#include "state_ops.cuh"

namespace dk {

constexpr uint32_t SP_MIN_EPOCH = 10u;                       // full_doko.rs:23
constexpr uint64_t SP_CALL_ACTIONS = 0x1Full << 33;           // AnnouncementReContra … AnnouncementBlack
constexpr uint32_t SP_DONE = 1u, SP_FORCED = 2u, SP_KEPT = 4u, SP_DROPPED = 8u;   // flags_out bits (DROPPED: buffer full)

// allowed_actions_by_action_index(false, epoch) / number_of_allowed_actions(epoch) as a mask (full_doko.rs:76-116)
DK_HD uint64_t sp_az_allowed(const dk_state& s, uint64_t az_epoch) {
    uint64_t m = fdo_state_legal_mask(s);
    if (az_epoch < SP_MIN_EPOCH) m &= ~SP_CALL_ACTIONS;
    return m;
}
// rand's StandardUniform for f32: 24 random bits scaled by 2^-24 (`rng.gen::<f32>()`, self_play.rs:88)
DK_HD float sp_keep_draw(uint32_t word) { return (float)(word >> 8) * (1.0f / 16777216.0f); }
// value target of a row recorded for mover `player`: new_data[(4 - player + i) % 4] = rewards[i], rewards = points / 8 (full_doko.rs:54-68)
DK_HD float sp_value_target(const dk_state& final_state, uint32_t player, uint32_t k) {
    return (float)final_state.points[(player + k) & 3u] / 8.0f;
}

}  // namespace dk
