// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// CPU restatement of the reference's UCT search over the full-rules game:
//   rs-doko-mcts/src/mcts/node.rs:21-260        McNode (new_root_node, new, min_max_normalized_q, uct, find_best_child)
//   rs-doko-mcts/src/mcts/mcts.rs:45-250        MCTS (select_promising_node, expand_single, backpropagate, monte_carlo_tree_search)
//   rs-doko-mcts/src/env/envs/env_state_full_doko.rs:62-220   McFullDokoEnvState (allowed_actions, by_action, random_rollout)
//   rs-doko-evaluator/src/full_doko/policy/mcts_policy.rs:96-118   Moves → (action, visits[39], values[39])
// Philox contract of a tree (unit, sub): iteration `it` uses the stream (unit, sub * iterations + it): the expansion pick is word 0 of
// SITE_EXPAND, the rollout draws sit at their usual state-derived ordinals (SITE_RESERVATION / SITE_CARD).
// f64 arithmetic in the reference's order; compile with -ffp-contract=off (Rust never contracts a*b+c).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>
#include "fdo.hpp"

namespace oracle {
namespace mcts {

constexpr uint64_t SOLO_AND_WEDDING_RESERVATIONS = 0x1FEull << 24;      // actions 25..32
constexpr uint64_t CALL_ACTIONS = 0x1Full << 33;

// McFullDokoEnvState::allowed_actions(first_expansion) (env_state_full_doko.rs:132-172)
inline uint64_t allowed_actions(const fdo::State& s, bool first_expansion) {
    uint64_t m = s.allowed_actions();
    if (!first_expansion) {
        for (int i = 0; i < s.reservations_round.len; ++i)
            if (s.reservations_round.r[i] >= fdo::R_DIAMONDS_SOLO) m &= ~SOLO_AND_WEDDING_RESERVATIONS;
        m &= ~CALL_ACTIONS;
    }
    return m;
}

struct Node {
    fdo::State state;
    bool is_terminal;
    int current_player;
    int last_action;
    uint64_t visits = 0;
    double win_score = 0.0;
    uint64_t unexpanded;
    std::vector<int> children;
    int parent;
};

struct Tree {
    std::vector<Node> nodes;

    int new_node(const fdo::State& s, int parent, int last_action, bool root) {                  // node.rs:138-200
        Node n;
        n.state = s;
        n.is_terminal = s.current_phase == fdo::PH_FINISHED;
        n.current_player = s.current_player < 0 ? 0 : s.current_player;
        n.last_action = last_action;
        n.unexpanded = allowed_actions(s, root);
        n.parent = parent;
        nodes.push_back(n);
        return (int)nodes.size() - 1;
    }
    double min_max_normalized_q(int self, int child) const {                                      // node.rs:202-236
        const Node& p = nodes[self];
        if (p.children.empty()) return 1.0;
        double min_q = std::numeric_limits<double>::infinity(), max_q = -std::numeric_limits<double>::infinity();
        for (int c : p.children) {
            double q = nodes[c].visits > 0 ? nodes[c].win_score / (double)nodes[c].visits : 0.0;
            if (q < min_q) min_q = q;
            if (q > max_q) max_q = q;
        }
        if (std::fabs(max_q - min_q) < std::numeric_limits<double>::epsilon()) return 1.0;
        double q = nodes[child].visits > 0 ? nodes[child].win_score / (double)nodes[child].visits : 0.0;
        return 2.0 * (q - min_q) / (max_q - min_q) - 1.0;
    }
    double uct(int self, uint64_t total_visits, double c) const {                                 // node.rs:238-256
        const Node& n = nodes[self];
        if (n.parent < 0) return 0.0;
        if (n.visits == 0) return std::numeric_limits<double>::infinity();
        double norm_q = min_max_normalized_q(n.parent, self);
        return norm_q + c * std::sqrt(std::log((double)total_visits) / (double)n.visits);
    }
    int find_best_child(int self, double c) const {                                               // node.rs:258-278
        double best_uct = -std::numeric_limits<double>::infinity();
        int best = -1;
        for (int ch : nodes[self].children) {
            double u = uct(ch, nodes[self].visits, c);
            if (u > best_uct) { best_uct = u; best = ch; }
        }
        return best;
    }
};

struct Move { int action; uint64_t visits; double value; };

// monte_carlo_tree_search (mcts.rs:160-232) on the Philox contract above.
inline std::vector<Move> search(const fdo::State& root_state, double c, uint32_t iterations, uint64_t seed, uint64_t unit, uint32_t sub, uint32_t epoch) {
    Tree t;
    t.nodes.reserve(iterations + 1);
    const int root = t.new_node(root_state, -1, -1, true);
    for (uint32_t it = 0; it < iterations; ++it) {
        PhiloxStream rng(seed, (uint32_t)unit, sub * iterations + it, epoch);
        int node = root;                                                                          // select_promising_node (:45-63)
        while (!t.nodes[node].children.empty()) {
            if (t.nodes[node].unexpanded != 0) break;
            node = t.find_best_child(node, c);
        }
        int explore = node;
        if (t.nodes[node].unexpanded != 0) {                                                      // expand_single (:65-104)
            uint64_t bit = random_single(t.nodes[node].unexpanded, rng, SITE_EXPAND);
            t.nodes[node].unexpanded &= ~bit;
            int action = __builtin_ctzll(bit);
            fdo::State ns = t.nodes[node].state;
            ns.play_action(action);
            explore = t.new_node(ns, node, action, false);
            t.nodes[node].children.push_back(explore);
        }
        fdo::State r = t.nodes[explore].state;                                                    // random_rollout (env_state_full_doko.rs:198-220)
        r.position_streams(rng);
        for (;;) { if (r.random_action_for_current_player_no_announcement(rng)) break; }
        double result[4];
        for (int p = 0; p < 4; ++p) result[p] = (double)r.end_of_game_stats.player_points[p];
        int temp = explore;                                                                       // backpropagate (:138-158)
        for (;;) {
            int parent = t.nodes[temp].parent;
            t.nodes[temp].visits += 1;
            if (parent < 0) break;
            t.nodes[temp].win_score += result[t.nodes[parent].current_player];
            temp = parent;
        }
    }
    std::vector<Move> moves;                                                                      // :220-229
    for (int ch : t.nodes[root].children)
        moves.push_back({t.nodes[ch].last_action, t.nodes[ch].visits, t.nodes[ch].win_score / (double)t.nodes[ch].visits});
    return moves;
}

// mcts_policy.rs:96-118: visits / values per action index and the move with the most visits (max_by_key keeps the LAST maximum).
inline int moves_to_arrays(const std::vector<Move>& moves, uint32_t visits[39], float values[39]) {
    for (int a = 0; a < 39; ++a) { visits[a] = 0; values[a] = 0.0f; }
    int best = -1; uint64_t best_visits = 0;
    for (const Move& m : moves) {
        visits[m.action] = (uint32_t)m.visits;
        values[m.action] = (float)m.value;
        if (best < 0 || m.visits >= best_visits) { best = m.action; best_visits = m.visits; }
    }
    return best;
}

}  // namespace mcts
}  // namespace oracle
