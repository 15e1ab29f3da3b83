// uct.cuh — the reference's UCT search over the full-rules game, one tree per thread.
//   rs-doko-mcts/src/mcts/node.rs:21-260 (McNode), mcts.rs:45-250 (MCTS::monte_carlo_tree_search),
//   rs-doko-mcts/src/env/envs/env_state_full_doko.rs:62-220 (McFullDokoEnvState), rs-doko-evaluator/.../mcts_policy.rs:96-118.
// A tree lives in a private slice of a device pool: iterations + 1 nodes of 208 bytes (the 128-byte state, exact integer win sum,
// visits, parent, the unexpanded-action mask and up to 12 child indices in insertion order — the order find_best_child walks).
// The f64 UCT arithmetic is done in the reference's order with explicitly rounded operations (no FMA contraction); ln(N) comes from
// a table the host fills with libm's log, the function Rust's f64::ln lowers to, so selection is bit-identical to the CPU path.
#pragma once
#include <math.h>
#include "matching.cuh"
#include "state_ops.cuh"

namespace dk {

constexpr uint32_t UCT_MAX_CHILDREN = 12u;       // distinct card types in a hand <= 12, reservations <= 9, announcements <= 2
constexpr uint32_t UCT_NONE = 0xFFFFFFFFu;
constexpr uint64_t UCT_ACTION_MASK = (1ull << 39) - 1ull;

struct alignas(16) UctNode {
    dk_state state;                   // McNode::state
    long long win;                    // win_score: a sum of integer points, exact
    uint32_t visits;
    uint32_t parent;                  // UCT_NONE for the root
    uint64_t info;                    // bits 0-38 unexpanded_actions | 40-41 current_player | 42 is_terminal | 48-53 last_action | 56-59 #children
    uint32_t child[UCT_MAX_CHILDREN];
};
static_assert(sizeof(UctNode) == 208, "UctNode layout");

DK_HD uint32_t uct_n_children(const UctNode& n) { return (uint32_t)(n.info >> 56) & 15u; }
DK_HD uint32_t uct_cur(const UctNode& n) { return (uint32_t)(n.info >> 40) & 3u; }
DK_HD uint32_t uct_last_action(const UctNode& n) { return (uint32_t)(n.info >> 48) & 63u; }

DK_HD double dk_inf() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7FF0000000000000LL);
#else
    return __builtin_inf();
#endif
}
DK_HD double dk_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
DK_HD double dk_dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
DK_HD double dk_ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
DK_HD double dk_dsqrt(double a) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(a);
#else
    return __builtin_sqrt(a);
#endif
}

// McFullDokoEnvState::allowed_actions(first_expansion) (env_state_full_doko.rs:132-172): below the root the solo / wedding
// reservations disappear once any seat has declared a solo, and the announcement calls always.
DK_HD uint64_t uct_allowed(const dk_state& s, bool first_expansion) {
    uint64_t m = fdo_state_legal_mask(s);
    if (!first_expansion) {
        for (uint32_t i = 0; i < s.n_reservations && i < 4u; ++i)
            if (s.reservations[i] >= 2u) m &= ~(0xFFull << 25);
        m &= ~(0x1Full << 33);
    }
    return m;
}
DK_HD void uct_init_node(UctNode& n, const dk_state& s, uint32_t parent, uint32_t last_action, bool root) {   // node.rs:138-200
    n.state = s;
    n.win = 0; n.visits = 0; n.parent = parent;
    const bool terminal = st_phase(s) == DK_PHASE_FINISHED;
    n.info = uct_allowed(s, root) | ((uint64_t)(terminal ? 0u : st_cur(s)) << 40) | ((uint64_t)(terminal ? 1u : 0u) << 42) | ((uint64_t)(last_action & 63u) << 48);
}
DK_HD double uct_q(const UctNode& c) { return c.visits > 0u ? dk_ddiv((double)c.win, (double)c.visits) : 0.0; }

// find_best_child (node.rs:258-278) with uct (:238-256) and min_max_normalized_q (:202-236) of every child.  The children's
// (visits, win) pairs are fetched first, all loads in flight together (one memory round trip per tree level instead of one per child).
DK_HD uint32_t uct_find_best_child(const UctNode* __restrict__ pool, uint32_t self, double c, const double* __restrict__ ln_table) {
    const UctNode& p = pool[self];
    const uint32_t nch = uct_n_children(p);
    uint32_t idx[UCT_MAX_CHILDREN], vis[UCT_MAX_CHILDREN];
    long long win[UCT_MAX_CHILDREN];
#pragma unroll
    for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) idx[k] = k < nch ? p.child[k] : self;
#pragma unroll
    for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) { vis[k] = pool[idx[k]].visits; win[k] = pool[idx[k]].win; }
    double min_q = dk_inf(), max_q = -dk_inf();
#pragma unroll
    for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) {
        if (k < nch) {
            double q = vis[k] > 0u ? dk_ddiv((double)win[k], (double)vis[k]) : 0.0;
            if (q < min_q) min_q = q;
            if (q > max_q) max_q = q;
        }
    }
    const double span = dk_dadd(max_q, -min_q);
    const bool flat = fabs(span) < 2.220446049250313e-16;            // f64::EPSILON
    const double ln_n = ln_table[p.visits];
    double best_uct = -dk_inf();
    uint32_t best = UCT_NONE;
#pragma unroll
    for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) {
        if (k < nch) {
            double u;
            if (vis[k] == 0u) u = dk_inf();
            else {
                double q = dk_ddiv((double)win[k], (double)vis[k]);
                double norm_q = flat ? 1.0 : dk_dadd(dk_ddiv(dk_dmul(2.0, dk_dadd(q, -min_q)), span), -1.0);
                u = dk_dadd(norm_q, dk_dmul(c, dk_dsqrt(dk_ddiv(ln_n, (double)vis[k]))));
            }
            if (u > best_uct) { best_uct = u; best = idx[k]; }
        }
    }
    return best;
}

// One iteration of monte_carlo_tree_search (mcts.rs:176-199): select → expand_single → random_rollout → backpropagate.
// key = the iteration's Philox unit.  Returns 1 when a node would need more than UCT_MAX_CHILDREN children (cannot happen for
// states reachable by the rules; reported instead of overflowing).
DK_HD uint32_t uct_iteration(UctNode* __restrict__ pool, uint32_t& n_nodes, const RngKey& key, double c, const double* __restrict__ ln_table,
                             const uint32_t* __restrict__ lut) {
    uint32_t node = 0;
    for (;;) {                                                                   // select_promising_node (:45-63)
        const UctNode& n = pool[node];
        if (uct_n_children(n) == 0u || (n.info & UCT_ACTION_MASK) != 0ull) break;
        node = uct_find_best_child(pool, node, c, ln_table);
    }
    uint32_t explore = node;
    const uint64_t unexpanded = pool[node].info & UCT_ACTION_MASK;
    alignas(16) dk_state s = pool[node].state;
    if (unexpanded != 0ull) {                                                    // expand_single (:65-104)
        const uint32_t nch = uct_n_children(pool[node]);
        if (nch >= UCT_MAX_CHILDREN) return 1u;
        U4 blk = rng_block(key, SITE_EXPAND, 0);
        const uint32_t a = pick_msb_rank64(unexpanded, mulhi(blk.x, popcll(unexpanded)));
        fdo_state_apply(s, a);                                                   // by_action
        explore = n_nodes++;
        uct_init_node(pool[explore], s, node, a, false);
        pool[node].child[nch] = explore;
        pool[node].info = (pool[node].info & ~(1ull << a) & ~(15ull << 56)) | ((uint64_t)(nch + 1u) << 56);
    }
    int32_t p[4];                                                                // random_rollout (env_state_full_doko.rs:198-220) from `s`
    {
        FdoLive g; FdoResume rs;
        if (fdo_state_to_live(s, g, rs)) { fdo_play_to_end<false, false>(g, key, &rs, lut); fdo_final_points(g, p); }
        else { p[0] = s.points[0]; p[1] = s.points[1]; p[2] = s.points[2]; p[3] = s.points[3]; }
    }
    uint32_t temp = explore;                                                     // backpropagate (:138-158)
    for (;;) {
        const uint32_t parent = pool[temp].parent;
        pool[temp].visits += 1u;
        if (parent == UCT_NONE) break;
        const uint32_t cp = uct_cur(pool[parent]);
        pool[temp].win += (long long)((cp & 2u) ? ((cp & 1u) ? p[3] : p[2]) : ((cp & 1u) ? p[1] : p[0]));
        temp = parent;
    }
    return 0u;
}

// Moves of the root (mcts.rs:220-229) as mcts_policy.rs:96-118 consumes them: visits / values by action index and the move with
// the most visits (max_by_key keeps the LAST maximum in child order); returns ACTION 0xFF when the root has no child.
template <class VisitT>
DK_HD uint32_t uct_moves(const UctNode* __restrict__ pool, VisitT* __restrict__ visits, float* __restrict__ values) {
    const UctNode& root = pool[0];
    uint32_t best = 0xFFu, best_visits = 0;
    for (uint32_t k = 0; k < uct_n_children(root); ++k) {
        const UctNode& ch = pool[root.child[k]];
        const uint32_t a = uct_last_action(ch);
        if (visits) visits[a] = (VisitT)ch.visits;
        if (values) values[a] = (float)dk_ddiv((double)ch.win, (double)ch.visits);
        if (best == 0xFFu || ch.visits >= best_visits) { best = a; best_visits = ch.visits; }
    }
    return best;
}

}  // namespace dk
