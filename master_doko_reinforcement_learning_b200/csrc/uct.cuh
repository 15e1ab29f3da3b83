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
    uint64_t m = fdo_state_legal_mask<true>(s);
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

DK_HD float dk_fdiv_fast(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);          // <= 2 ulp; the filter's error budget below allows far more
#else
    return a / b;
#endif
}

// Exact UCT value of one child (node.rs:238-256 with min_max_normalized_q :202-236) from exact min / max Q of its siblings.
DK_HD double uct_value_exact(long long win, uint32_t vis, double min_q, double span, bool flat, double ln_n, double c) {
    if (vis == 0u) return dk_inf();
    const double q = dk_ddiv((double)win, (double)vis);
    const double norm_q = flat ? 1.0 : dk_dadd(dk_ddiv(dk_dmul(2.0, dk_dadd(q, -min_q)), span), -1.0);
    return dk_dadd(norm_q, dk_dmul(c, dk_dsqrt(dk_ddiv(ln_n, (double)vis))));
}

// find_best_child (node.rs:258-278): the child with the strictly greatest f64 UCT value, first in child order among equals.
//
// The children's (visits, win) pairs are fetched first, all loads in flight together (one memory round trip per tree level).
// The f64 divisions and square roots of the reference are software sequences of ~30 instructions each on the GPU, so the
// decision is taken in two stages that TOGETHER are bit-identical to evaluating every child in f64:
//   1. a cheap f32 evaluation u~ with a proven error bound |u~ - u| <= eps.  Only children with u~ >= max(u~) - 2 eps can be the
//      exact arg-max (and all exact ties are among them).  Almost always exactly one child survives and is returned — no f64 at all.
//      Bound: Q = win / visits with |Q| <= 128 has an absolute f32 error <= 5e-5, hence numerator and span of the normalisation
//      <= 1.2e-4 each and the normalised Q <= 5e-4 / span~ + 1e-6; the exploration term has a relative error <= 5e-7 and a
//      magnitude <= 4.64 c (N < 2^31).  eps = 1e-3 / span~ + 4e-5 c + 4e-5 leaves a factor >= 2 everywhere; the filter is skipped
//      when span~ < 4e-3 or a child has no visits.
//   2. the survivors are evaluated exactly.  The exact min / max Q come from the children that are minimal / maximal as RATIONALS
//      (integer cross-multiplication; IEEE division is monotone, so the rounded extremes are the extremes of the rounded values).
DK_HD uint32_t uct_find_best_child(const UctNode* __restrict__ pool, uint32_t self, double c, double ln_n, bool use_filter = true) {   // ln_n = ln(visits of `self`)
    const UctNode& p = pool[self];
    const uint32_t nch = uct_n_children(p);
    uint32_t idx[UCT_MAX_CHILDREN], vis[UCT_MAX_CHILDREN];
    long long win[UCT_MAX_CHILDREN];
#pragma unroll
    for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) idx[k] = k < nch ? p.child[k] : self;
#pragma unroll
    for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) { vis[k] = pool[idx[k]].visits; win[k] = pool[idx[k]].win; }
    uint32_t cand = (1u << nch) - 1u;
    {   // stage 1: f32 filter
        float qf[UCT_MAX_CHILDREN];
        float minf = 3.0e38f, maxf = -3.0e38f;
        bool any_unvisited = false;
#pragma unroll
        for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) {
            const bool on = k < nch;
            any_unvisited |= on && vis[k] == 0u;
            const float qk = dk_fdiv_fast((float)win[k], (float)(vis[k] ? vis[k] : 1u));
            qf[k] = qk;
            if (on && qk < minf) minf = qk;
            if (on && qk > maxf) maxf = qk;
        }
        const float spanf = maxf - minf;
        if (use_filter && !any_unvisited && spanf >= 4.0e-3f && p.visits < (1u << 24)) {
            const float cf = (float)c, lnf = (float)ln_n;
            const float eps = dk_fdiv_fast(1.0e-3f, spanf) + 4.0e-5f * cf + 4.0e-5f;
            const float scale = dk_fdiv_fast(2.0f, spanf);
            float uf[UCT_MAX_CHILDREN];
            float top = -3.0e38f;
#pragma unroll
            for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) {
                const float u = (qf[k] - minf) * scale - 1.0f + cf * sqrtf(dk_fdiv_fast(lnf, (float)(vis[k] ? vis[k] : 1u)));
                uf[k] = u;
                if (k < nch && u > top) top = u;
            }
            const float thr = top - 2.0f * eps;
            uint32_t m = 0;
#pragma unroll
            for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) m |= (k < nch && uf[k] >= thr) ? (1u << k) : 0u;
            cand = m;
            if ((m & (m - 1u)) == 0u) {                                   // a single survivor: it is the exact arg-max
                uint32_t best = self;
#pragma unroll
                for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) if ((m >> k) & 1u) best = idx[k];
                return best;
            }
        }
    }
    // stage 2: exact evaluation of the survivors.  Extremes of Q as rationals: win_a / vis_a < win_b / vis_b  <=>  win_a vis_b < win_b vis_a
    // (unvisited children count as Q = 0 = 0 / 1).
    long long lo_w = 0, hi_w = 0;
    uint32_t lo_v = 0, hi_v = 0;
#pragma unroll 1
    for (uint32_t k = 0; k < nch; ++k) {
        const long long w = vis[k] ? win[k] : 0;
        const uint32_t v = vis[k] ? vis[k] : 1u;
        if (lo_v == 0u || w * (long long)lo_v < lo_w * (long long)v) { lo_w = w; lo_v = v; }
        if (hi_v == 0u || w * (long long)hi_v > hi_w * (long long)v) { hi_w = w; hi_v = v; }
    }
    const double min_q = dk_ddiv((double)lo_w, (double)lo_v), max_q = dk_ddiv((double)hi_w, (double)hi_v);
    const double span = dk_dadd(max_q, -min_q);
    const bool flat = fabs(span) < 2.220446049250313e-16;            // f64::EPSILON
    double best_uct = -dk_inf();
    uint32_t best = UCT_NONE;
#pragma unroll 1
    for (uint32_t k = 0; k < nch; ++k) {
        if (!((cand >> k) & 1u)) continue;
        const double u = uct_value_exact(win[k], vis[k], min_q, span, flat, ln_n, c);
        if (u > best_uct) { best_uct = u; best = idx[k]; }
    }
    return best;
}

// One iteration of monte_carlo_tree_search (mcts.rs:176-199): select → expand_single → random_rollout → backpropagate.
// key = the iteration's Philox unit.  Returns 1 when a node would need more than UCT_MAX_CHILDREN children (cannot happen for
// states reachable by the rules; reported instead of overflowing).
// `sync()` is called between the four phases by EVERY thread (active or not): the kernel passes a block barrier so that all warps of
// a block run the same phase — and hence the same stretch of this 100+ KB of code — at the same time; the host simulator passes a no-op.
struct UctNoSync { DK_HD void operator()() const {} };
template <class Sync = UctNoSync>
DK_HD uint32_t uct_iteration(UctNode* __restrict__ pool, uint32_t& n_nodes, const RngKey& key, double c, const double* __restrict__ ln_table,
                             const uint32_t* __restrict__ lut, bool active = true, Sync sync = Sync()) {
    uint32_t node = 0, err = 0;
    if (active) {
        for (;;) {                                                               // select_promising_node (:45-63)
            const UctNode& n = pool[node];
            if (uct_n_children(n) == 0u || (n.info & UCT_ACTION_MASK) != 0ull) break;
            node = uct_find_best_child(pool, node, c, ln_table[n.visits]);
        }
    }
    sync();
    uint32_t explore = node;
    alignas(16) dk_state s;
    if (active) {
        const uint64_t unexpanded = pool[node].info & UCT_ACTION_MASK;
        s = pool[node].state;
        if (unexpanded != 0ull) {                                                // expand_single (:65-104)
            const uint32_t nch = uct_n_children(pool[node]);
            if (nch >= UCT_MAX_CHILDREN) err = 1u;
            else {
                U4 blk = rng_block(key, SITE_EXPAND, 0);
                const uint32_t a = pick_msb_rank64(unexpanded, mulhi(blk.x, popcll(unexpanded)));
                fdo_state_apply<true>(s, a);                                     // by_action (record in local memory)
                explore = n_nodes++;
                uct_init_node(pool[explore], s, node, a, false);
                pool[node].child[nch] = explore;
                pool[node].info = (pool[node].info & ~(1ull << a) & ~(15ull << 56)) | ((uint64_t)(nch + 1u) << 56);
            }
        }
    }
    sync();
    int32_t p[4] = {0, 0, 0, 0};                                                 // random_rollout (env_state_full_doko.rs:198-220) from `s`
    if (active && !err) {
        FdoLive g; FdoResume rs;
        if (fdo_state_to_live<true>(s, g, rs)) { fdo_play_to_end<false, false>(g, key, &rs, lut); fdo_final_points(g, p); }
        else { p[0] = s.points[0]; p[1] = s.points[1]; p[2] = s.points[2]; p[3] = s.points[3]; }
    }
    sync();
    if (active && !err) {
        uint32_t temp = explore;                                                 // backpropagate (:138-158)
        for (;;) {
            const uint32_t parent = pool[temp].parent;
            pool[temp].visits += 1u;
            if (parent == UCT_NONE) break;
            const uint32_t cp = uct_cur(pool[parent]);
            pool[temp].win += (long long)((cp & 2u) ? ((cp & 1u) ? p[3] : p[2]) : ((cp & 1u) ? p[1] : p[0]));
            temp = parent;
        }
    }
    return err;
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
