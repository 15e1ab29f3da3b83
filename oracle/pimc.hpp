// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// CPU restatement of the PIMC move decision around the hot path:
//   rs-doko-py-bridge/src/compare_impi/policy_fusion.rs:18-123   PolicyFusionMaxN / PolicyFusionAverageStrategy
//   rs-doko-py-bridge/src/compare_impi/compare_impi.rs:212-372    DefaultImpiPolicy::execute (sample → per-sample policy → fuse)
// and of the flat Monte-Carlo per-determinization policy that the CUDA evaluator uses in the place of `EvFullDokoPolicy::evaluate`
// (a new design, not a reference function: the reference plugs a UCT search in there; SURVEY.md §8f N2 names the flat evaluator).
//
// Pinned by the reference's own four fusion tests (policy_fusion.rs:131-313 → tests/golden/policy_fusion_cases.json).
#pragma once
#include <cstdint>
#include <cmath>
#include "fdo.hpp"
#include "matching.hpp"

namespace oracle {
namespace pimc {

constexpr int N_ACTIONS = 39;                  // FdoAction::COUNT
constexpr int ACTION_NONE = 0xFF;

// PolicyFusionMaxN::fuse (policy_fusion.rs:23-73).  rows = successful samples only.
inline int fuse_max_n(const uint32_t* visits /*[n_rows][39]*/, size_t n_rows, uint64_t allowed) {
    uint32_t cumulative[N_ACTIONS];
    for (int a = 0; a < N_ACTIONS; ++a) cumulative[a] = ((allowed >> a) & 1u) ? 0u : UINT32_MAX;      // :31-38
    for (size_t s = 0; s < n_rows; ++s) {
        const uint32_t* v = visits + s * N_ACTIONS;
        int idx[N_ACTIONS], n = 0;
        for (int a = 0; a < N_ACTIONS; ++a) if ((allowed >> a) & 1u) idx[n++] = a;                   // filter (:48-52)
        // sort_by_key(Reverse(visits)) is a STABLE sort (:55): insertion sort keeps equal keys in index order
        for (int i = 1; i < n; ++i) {
            int x = idx[i], j = i - 1;
            while (j >= 0 && v[idx[j]] < v[x]) { idx[j + 1] = idx[j]; --j; }
            idx[j + 1] = x;
        }
        for (int r = 0; r < n; ++r) cumulative[idx[r]] += (uint32_t)r + 1u;                          // :59-62
    }
    int best = 0;                                                                                    // min_by_key → FIRST minimum (:65-70)
    for (int a = 1; a < N_ACTIONS; ++a) if (cumulative[a] < cumulative[best]) best = a;
    return best;
}

// PolicyFusionAverageStrategy::fuse (policy_fusion.rs:79-123): f32 arithmetic in the reference's order; the allowed set is ignored.
inline int fuse_average(const uint32_t* visits, size_t n_rows) {
    float sum[N_ACTIONS];
    for (int a = 0; a < N_ACTIONS; ++a) sum[a] = 0.0f;
    for (size_t s = 0; s < n_rows; ++s) {
        const uint32_t* v = visits + s * N_ACTIONS;
        uint64_t total = 0;
        for (int a = 0; a < N_ACTIONS; ++a) total += v[a];
        for (int a = 0; a < N_ACTIONS; ++a) {
            volatile float p = (float)v[a] / (float)total;                                           // 0/0 = NaN propagates as in Rust
            sum[a] = sum[a] + p;
        }
    }
    // max_by(partial_cmp.unwrap_or(Equal)) keeps the LAST element among equals (Iterator::max_by folds with `Greater => x, _ => y`)
    int best = 0;
    for (int a = 1; a < N_ACTIONS; ++a) {
        bool best_greater = sum[best] > sum[a];                                                      // NaN compares false ⇒ "Equal" ⇒ take a
        if (!best_greater) best = a;
    }
    return best;
}

// Flat Monte-Carlo per-determinization policy on the Philox contract:
//   determinization d of info-state `unit` = card_matching on the stream (unit, d)                       [same as dk_determinize]
//   rollout r of determinization d: for EVERY legal action a of the seat to move, clone the determinized state, play a, then
//   random_rollout (the _no_announcement policy, env_state_full_doko.rs:198-220) on the stream (unit, d * n_rollouts + r) —
//   the same stream for all actions (common random numbers).  value_sum[a] += points[mover]; the action with the strictly
//   greatest value (first in action-index order among equals) gets one visit.
// Returns the determinization status (0 ok; otherwise visits/value_sum stay zero).
inline int flat_mc(const fdo::State& root, uint64_t seed, uint64_t unit, uint32_t det, uint32_t n_rollouts, uint32_t epoch,
                   uint32_t visits[N_ACTIONS], int64_t value_sum[N_ACTIONS]) {
    for (int a = 0; a < N_ACTIONS; ++a) { visits[a] = 0; value_sum[a] = 0; }
    if (root.current_player < 0) return 0;
    const int mover = root.current_player;
    PhiloxStream rm(seed, (uint32_t)unit, det, epoch);
    fdo::Hand oh[4]; int ores[4];
    int status = fdo::card_matching(root, rm, oh, ores);
    if (status != 0) return status;
    const fdo::State ds = fdo::with_hands_and_reservations(root, oh, ores);
    const uint64_t allowed = ds.allowed_actions();
    for (uint32_t r = 0; r < n_rollouts; ++r) {
        int best_a = -1; int32_t best_v = 0;
        for (int a = 0; a < N_ACTIONS; ++a) {
            if (!((allowed >> a) & 1ull)) continue;
            fdo::State t = ds;
            t.play_action(a);
            if (t.current_phase != fdo::PH_FINISHED) {
                PhiloxStream rr(seed, (uint32_t)unit, det * n_rollouts + r, epoch);
                t.position_streams(rr);
                for (;;) { if (t.random_action_for_current_player_no_announcement(rr)) break; }
            }
            int32_t v = t.end_of_game_stats.player_points[mover];
            value_sum[a] += v;
            if (best_a < 0 || v > best_v) { best_a = a; best_v = v; }
        }
        if (best_a >= 0) visits[best_a] += 1;
    }
    return 0;
}

}  // namespace pimc
}  // namespace oracle
