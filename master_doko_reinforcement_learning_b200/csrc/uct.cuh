// uct.cuh — the reference's UCT search over the full-rules game, one tree per thread, one KERNEL PER PHASE.
//   rs-doko-mcts/src/mcts/node.rs:21-260 (McNode), mcts.rs:45-250 (MCTS::monte_carlo_tree_search),
//   rs-doko-mcts/src/env/envs/env_state_full_doko.rs:62-220 (McFullDokoEnvState), rs-doko-evaluator/.../mcts_policy.rs:96-118.
//
// Round 1 ran the whole iteration (select → expand → rollout → backpropagate) as one 100 KB kernel body: ncu showed 15 warps per
// issue slot waiting for instructions (the blocks of an SM sat in different phases and evicted each other's code), 14.5 of 32 lanes
// and 4.35e8 iterations/s (profiles/r01_uct_ncu_summary.json).  Now an iteration of ALL trees is two launches, so every SM
// executes one phase's code at a time and each phase has its own register budget:
//   uct_tree_kernel     select_promising_node (dependent loads, latency bound: the per-level data is ONE block of at most five
//                       sectors) + expand_single (pick an unexpanded action, play it on the node's record, append the child)
//   uct_rollout_kernel  the K2-style _no_announcement playout from the new record (lock-step card loop, full table set), then
//                       backpropagate: the memory-bound walk up of the warps that are done overlaps the playouts still running
// Per-tree hand-over between the launches (selected node, node to roll out from, packed result) is 4 bytes per tree and launch.
// Measured at 131 072 trees x 256 iterations (profiles/r02_uct_history.json): 4.35e8 iterations/s (round 1) -> 8.4e8 (two kernels,
// compact blocks, chain skipping, path buffer) -> 9.9e8 (backpropagation as 64-bit reductions, one memory round trip per level of
// the walk, special-function f32 filter) -> 1.09e9 with the batch cut into three parts on three streams (cabi.cu: dk_uct_search).
//
// Data layout.  A tree owns `iterations + 1` node slots.  A node is split in two arrays:
//   UctHead (160 B = 5 sectors)  sector 0: info | children 0-3;  sector 1: children 4-11;  sectors 2-4: (visits, exact integer
//                                win sum) of the children, IN INSERTION ORDER (the order find_best_child walks)
//   dk_state (128 B)             McNode::state
// The statistics of a node live in ITS PARENT's block: selection at a node reads one contiguous block instead of the node plus twelve
// scattered children (round 1: 13 sectors in 13 lines per level), and backpropagation updates the slot it came through.  A node's own
// visit count — the N of ln N — is 1 + the visits of its children (the iteration that created it plus every later one through it);
// the root's is the iteration number.  A child entry also carries HOW MANY ACTIONS its target had when it was created, so the walk
// knows before it loads a block which sectors it needs (two for <= 4 actions, all five only for wide nodes).
//
// Single-action nodes are skipped.  Below the root every announcement-phase state has exactly one action (NoAnnouncement; the calls
// are removed, env_state_full_doko.rs:132-172), so an early-game path is mostly chains of such nodes between two card nodes: their
// only child is chosen whatever its numbers are and nobody ever reads their statistics.  The child entry of a multi-action node (an
// ANCHOR; the root always is one) therefore points at the END of the chain below that child — the first node that is an anchor
// itself or still has its action to expand — and is moved when the chain grows.  The tree keeps the reference's shape (every chain
// node is a node, created by its own iteration with its own rollout); only the walks skip them.  Measured before this change: paths
// of 15-20 levels, 217 MB of DRAM reads per iteration of 131 072 trees for the walk up alone (profiles/r02_uct_rollout_v3_*.json).
//
// The walk down records its anchors — (node, slot, seat to move) per level, <= 53 levels: root + 4 reservations + 48 cards — in a
// per-tree path buffer ([level][tree], coalesced).  Backpropagation reads that list and updates the slots with INDEPENDENT
// read-modify-writes: no parent pointers, no dependent loads on the way up.
//
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

struct UctStat { uint32_t vis; int32_t win; };   // win_score: a sum of integer points, exact (|points| < 128, < 2^24 iterations)
struct alignas(32) UctHead {
    uint64_t info;                    // bits 0-38 unexpanded_actions | 40-41 current_player | 42 is_terminal | 44-47 #actions at creation |
                                      // 48-53 last_action | 56-59 #children
    uint32_t spare[2];
    uint32_t child[UCT_MAX_CHILDREN]; // bits 0-23 node at the end of the single-action chain below child k | bits 28-31 #actions of that node
    UctStat stat[UCT_MAX_CHILDREN];   // visits / win_score of child k
};
static_assert(sizeof(UctHead) == 160, "UctHead layout");
constexpr uint32_t UCT_NODE_MASK = 0x00FFFFFFu;
constexpr uint64_t UCT_MAX_ITERATIONS = (1ull << 24) - 2ull;   // int32 win sums, 24-bit node indices
constexpr uint32_t UCT_MAX_PATH = 56u;                          // anchors on a path: root + 4 reservations + 48 cards = 53
DK_HD uint32_t uct_path_entry(uint32_t node, uint32_t slot, uint32_t cur) { return node | (slot << 24) | (cur << 28); }

// Per-tree control words (one array each, [n_trees]): what the phases hand to each other.
enum : uint32_t { UCT_CTL_ACTIVE = 1u, UCT_CTL_ROLLOUT = 2u, UCT_CTL_WIDTH_SHIFT = 8u, UCT_CTL_WIDTH_MASK = 15u << 8 };   // bits 8-11: #actions of the root
struct UctPool {
    UctHead* heads;       // [iterations + 1][n_trees]
    dk_state* states;     // [iterations + 1][n_trees]
    uint32_t* path;       // [UCT_MAX_PATH][n_trees]: the anchors of the current iteration's path (uct_path_entry)
    uint32_t* path_len;
    uint8_t* root_action; // [n_trees][16]: action of the root's child k (the root's entries point past single-action chains)
    uint32_t* ctl;        // UCT_CTL_* flags
    uint32_t* explore;    // node the rollout starts from
    uint32_t* result;     // player_points of the rollout, int8 per seat
    uint32_t* n_nodes;
    uint32_t* status;     // 0 ok, determinization status, 3 = a node would need more than UCT_MAX_CHILDREN children / a path more than UCT_MAX_PATH anchors
    uint64_t n_trees;
    uint64_t node_stride, tree_stride;   // element (node, tree) of heads / states: node * node_stride + tree * tree_stride
    DK_HD UctHead& head(uint64_t t, uint32_t node) const { return heads[(uint64_t)node * node_stride + t * tree_stride]; }
    DK_HD dk_state& state(uint64_t t, uint32_t node) const { return states[(uint64_t)node * node_stride + t * tree_stride]; }
    DK_HD uint32_t& path_at(uint64_t t, uint32_t level) const { return path[(uint64_t)level * ((n_trees + 63ull) & ~63ull) + t]; }
};
// Carves the caller's workspace (256-byte aligned inside): heads | states | path buffer | control arrays.
DK_HD uint64_t uct_workspace_bytes(uint64_t n_trees, uint64_t iterations) {
    const uint64_t arr = (n_trees * 4ull + 255ull) & ~255ull;
    return 256ull + n_trees * (iterations + 1ull) * (sizeof(UctHead) + sizeof(dk_state)) + (UCT_MAX_PATH + 6ull) * arr + ((n_trees * 16ull + 255ull) & ~255ull);
}
// tree_major = false: [node][tree] (the node every tree appends in an iteration and the root level are contiguous across a warp);
// tree_major = true: [tree][node] (a tree's nodes share a few pages).
inline UctPool uct_carve(void* workspace, uint64_t n_trees, uint64_t iterations, bool tree_major = false) {
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255u) & ~(uintptr_t)255u);
    UctPool P;
    P.n_trees = n_trees;
    P.node_stride = tree_major ? 1ull : n_trees;
    P.tree_stride = tree_major ? iterations + 1ull : 1ull;
    P.heads = reinterpret_cast<UctHead*>(p); p += n_trees * (iterations + 1ull) * sizeof(UctHead);
    P.states = reinterpret_cast<dk_state*>(p); p += n_trees * (iterations + 1ull) * sizeof(dk_state);
    const uint64_t arr = (n_trees * 4ull + 255ull) & ~255ull;
    P.path = reinterpret_cast<uint32_t*>(p); p += UCT_MAX_PATH * arr;
    P.path_len = reinterpret_cast<uint32_t*>(p); p += arr;
    P.ctl = reinterpret_cast<uint32_t*>(p); p += arr;
    P.explore = reinterpret_cast<uint32_t*>(p); p += arr;
    P.result = reinterpret_cast<uint32_t*>(p); p += arr;
    P.n_nodes = reinterpret_cast<uint32_t*>(p); p += arr;
    P.status = reinterpret_cast<uint32_t*>(p); p += arr;
    P.root_action = reinterpret_cast<uint8_t*>(p);
    return P;
}

DK_HD uint32_t uct_n_children(uint64_t info) { return (uint32_t)(info >> 56) & 15u; }
DK_HD uint32_t uct_n_actions(uint64_t info) { return (uint32_t)(info >> 44) & 15u; }
DK_HD uint32_t uct_cur(uint64_t info) { return (uint32_t)(info >> 40) & 3u; }
DK_HD uint32_t uct_last_action(uint64_t info) { return (uint32_t)(info >> 48) & 63u; }
DK_HD uint32_t uct_pack_points(const int32_t p[4]) {
    return ((uint32_t)p[0] & 255u) | (((uint32_t)p[1] & 255u) << 8) | (((uint32_t)p[2] & 255u) << 16) | (((uint32_t)p[3] & 255u) << 24);
}
DK_HD int32_t uct_unpack_point(uint32_t packed, uint32_t seat) { return (int32_t)(int8_t)((packed >> (8u * seat)) & 255u); }

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
template <bool IDX = true>
DK_HD uint64_t uct_allowed(const dk_state& s, bool first_expansion) {
    uint64_t m = fdo_state_legal_mask<IDX>(s);
    if (!first_expansion) {
        bool solo = false;
#pragma unroll
        for (uint32_t i = 0; i < 4u; ++i) solo |= i < s.n_reservations && s.reservations[i] >= 2u;
        if (solo) m &= ~(0xFFull << 25);
        m &= ~(0x1Full << 33);
    }
    return m;
}
template <bool IDX = true>
DK_HD uint64_t uct_node_info(const dk_state& s, uint32_t last_action, bool root) {                 // node.rs:138-200
    const bool terminal = st_phase(s) == DK_PHASE_FINISHED;
    const uint64_t allowed = uct_allowed<IDX>(s, root);
    return allowed | ((uint64_t)(terminal ? 0u : st_cur(s)) << 40) | ((uint64_t)(terminal ? 1u : 0u) << 42) | ((uint64_t)popcll(allowed) << 44) |
           ((uint64_t)(last_action & 63u) << 48);
}

DK_HD float dk_fdiv_fast(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);          // <= 2 ulp; the filter's error budget below allows far more
#else
    return a / b;
#endif
}
// 1 / sqrt(v) and ln(n) for the f32 filter: one special-function instruction each on the device (MUFU.RSQ: relative error <= 2^-22;
// MUFU.LG2 * ln 2: relative error <= 2^-21.4 for n >= 2, exactly 0 for n = 1); the host simulator uses libm — the filter's outcome
// does not depend on which (see uct_best_slot).
DK_HD float dk_rsqrt_fast(float v) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#else
    return 1.0f / __builtin_sqrtf(v);
#endif
}
DK_HD float dk_logf_fast(float v) {
#if defined(__CUDA_ARCH__)
    return __logf(v);
#else
    return __builtin_logf(v);
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
// Returns the SLOT of the child.  vis / win: the statistics of the nch <= W children in insertion order (entries >= nch are ignored).
//
// The f64 divisions and square roots of the reference are software sequences of ~30 instructions each on the GPU, so the
// decision is taken in two stages that TOGETHER are bit-identical to evaluating every child in f64:
//   1. a cheap f32 evaluation u~ with a proven error bound |u~ - u| <= eps.  Only children with u~ >= max(u~) - 2 eps can be the
//      exact arg-max (and all exact ties are among them).  Almost always exactly one child survives and is returned — no f64 at all
//      (measured on the host simulator over 256-iteration searches: 98 % of the levels; `same` below takes most of the rest).
//      Bound: Q = win * rcp(visits) (win exact in f32, reciprocal <= 1 ulp, product 0.5 ulp) with |Q| <= 128 has an absolute error
//      <= 5e-5, hence numerator and span of the normalisation <= 1.2e-4 each and the normalised Q <= 5e-4 / span~ + 1e-6; the
//      exploration term c * sqrt(ln N) * rsqrt(visits) has a relative error <= 1e-6 (ln N 3e-7 halved by the root, rsqrt 2.4e-7,
//      four roundings of 6e-8) and a magnitude <= 4.64 c (N < 2^31).  eps = 1e-3 / span~ + 4e-5 c + 4e-5 leaves a factor >= 2
//      everywhere; the filter is skipped when span~ < 4e-3 or a child has no visits.
//   2. the survivors are evaluated exactly.  The exact min / max Q come from the children that are minimal / maximal as RATIONALS
//      (integer cross-multiplication; IEEE division is monotone, so the rounded extremes are the extremes of the rounded values).
// Round 2 (first form) looked (1 / v, 1 / sqrt(v)) up in a table: twelve dependent 8-byte loads per level behind the statistics load —
// a second round trip on the walk's critical path and 24 LSU instructions; now the two factors are special-function instructions.
// `ln_tab` (host libm values) is read only by stage 2.
struct UctTables { const double* ln; };
template <uint32_t W = UCT_MAX_CHILDREN>
DK_HD uint32_t uct_best_slot(uint32_t nch, const uint32_t vis[W], const int32_t win[W], uint32_t parent_visits, double c, const double* ln_tab, double ln_host,
                             bool use_filter = true) {
    uint32_t cand = (1u << nch) - 1u;
    {   // stage 1: f32 filter
        float qf[W], rs[W];
        float minf = 3.0e38f, maxf = -3.0e38f;
        bool any_unvisited = false;
#pragma unroll
        for (uint32_t k = 0; k < W; ++k) {
            const bool on = k < nch;
            any_unvisited |= on && vis[k] == 0u;
            const float vf = (float)(vis[k] ? vis[k] : 1u);
            const float qk = dk_fdiv_fast((float)win[k], vf);
            qf[k] = qk; rs[k] = dk_rsqrt_fast(vf);
            minf = on ? fminf(minf, qk) : minf;
            maxf = on ? fmaxf(maxf, qk) : maxf;
        }
        const float spanf = maxf - minf;
        if (use_filter && !any_unvisited && spanf >= 4.0e-3f && parent_visits < (1u << 24)) {
            const float cf = (float)c, expl = cf * sqrtf(dk_logf_fast((float)parent_visits));
            const float eps = dk_fdiv_fast(1.0e-3f, spanf) + 4.0e-5f * cf + 4.0e-5f;
            const float scale = dk_fdiv_fast(2.0f, spanf);
            float uf[W];
            float top = -3.0e38f;
#pragma unroll
            for (uint32_t k = 0; k < W; ++k) {
                const float u = (qf[k] - minf) * scale - 1.0f + expl * rs[k];
                uf[k] = u;
                top = k < nch ? fmaxf(top, u) : top;
            }
            const float thr = top - 2.0f * eps;
            uint32_t m = 0;
#pragma unroll
            for (uint32_t k = 0; k < W; ++k) m |= (k < nch && uf[k] >= thr) ? (1u << k) : 0u;
            cand = m;
            if ((m & (m - 1u)) == 0u) return ffs0(m);                     // a single survivor: it is the exact arg-max
        }
    }
    if (use_filter) {   // Candidates with IDENTICAL statistics have identical exact values (same inputs, same arithmetic), and the first in child order
        // wins among equals: when every candidate carries the first candidate's numbers — the common tie of a young tree, siblings
        // visited equally often with equal results — the decision needs no f64 at all.  Only valid when the candidates are ALL children
        // or survivors of the filter (then no other child can beat them), which is what `cand` holds here.
        const uint32_t first = ffs0(cand);
        uint32_t v0 = 0; int32_t w0 = 0;
        bool same = true;
#pragma unroll
        for (uint32_t k = 0; k < W; ++k) { v0 = k == first ? vis[k] : v0; w0 = k == first ? win[k] : w0; }
#pragma unroll
        for (uint32_t k = 0; k < W; ++k) same &= !((cand >> k) & 1u) || (vis[k] == v0 && win[k] == w0);
        if (same) return first;
    }
    if (use_filter && cand == (1u << nch) - 1u && c >= 1.0e-3 && parent_visits >= 2u) {
        // Every child has the same Q as a RATIONAL (late in a game the result often no longer depends on the move): the rounded
        // quotients are identical, the span is exactly 0 and each value is 1 + c * sqrt(ln N / visits) — evaluated by the same rounded
        // operations for every child, hence non-increasing in `visits`, and STRICTLY decreasing for c >= 1e-3: two visit counts
        // below 2^24 move ln N / visits by a relative 2^-24 at least, its root by 2^-25, and c * root >= 2e-7 keeps that gap far
        // above the half-ulp of 1 + c * root.  So the arg-max is the first child with the fewest visits: no f64 evaluation.
        // (45 % of the levels the filter cannot decide, measured on the host simulator; unvisited children cannot occur on a walk
        // but are excluded anyway.)
        bool all_equal = vis[0] != 0u;
        uint32_t best = 0u, best_v = vis[0];
#pragma unroll
        for (uint32_t k = 1; k < W; ++k) {
            const bool on = k < nch;
            all_equal &= !on || (vis[k] != 0u && (long long)win[k] * (long long)vis[0] == (long long)win[0] * (long long)vis[k]);
            if (on && vis[k] < best_v) { best_v = vis[k]; best = k; }
        }
        if (all_equal) return best;
    }
    // stage 2: exact evaluation of the survivors.  Extremes of Q as rationals: win_a / vis_a < win_b / vis_b  <=>  win_a vis_b < win_b vis_a
    // (unvisited children count as Q = 0 = 0 / 1).
    long long lo_w = 0, hi_w = 0;
    uint32_t lo_v = 0, hi_v = 0;
#pragma unroll 1
    for (uint32_t k = 0; k < nch; ++k) {
        const long long w = vis[k] ? (long long)win[k] : 0;
        const uint32_t v = vis[k] ? vis[k] : 1u;
        if (lo_v == 0u || w * (long long)lo_v < lo_w * (long long)v) { lo_w = w; lo_v = v; }
        if (hi_v == 0u || w * (long long)hi_v > hi_w * (long long)v) { hi_w = w; hi_v = v; }
    }
    const double ln_n = ln_tab ? ln_tab[parent_visits] : ln_host;
    const double min_q = dk_ddiv((double)lo_w, (double)lo_v), max_q = dk_ddiv((double)hi_w, (double)hi_v);
    const double span = dk_dadd(max_q, -min_q);
    const bool flat = fabs(span) < 2.220446049250313e-16;            // f64::EPSILON
    double best_uct = -dk_inf();
    uint32_t best = UCT_NONE;
#pragma unroll 1
    for (uint32_t k = 0; k < nch; ++k) {
        if (!((cand >> k) & 1u)) continue;
        const double u = uct_value_exact((long long)win[k], vis[k], min_q, span, flat, ln_n, c);
        if (u > best_uct) { best_uct = u; best = k; }
    }
    return best;
}

// ---- the phases of one iteration, per tree (the kernels in selfplay_kernels.cuh call these; the host simulator runs them in sequence) ----

// Root of a tree (mcts.rs:160-175): [determinize →] McNode::new_root_node.  Returns the tree's status (0 = searching).
DK_HD uint32_t uct_phase_root(const UctPool& P, uint64_t t, const dk_state& root_state, bool determinize, const RngKey& det_key) {
    alignas(16) dk_state s = root_state;
    uint32_t status = 0;
    if (determinize && st_phase(s) != DK_PHASE_FINISHED) {
        MatchPrep prep;
        fdo_match_prepare(s, prep);
        uint64_t h[4];
        uint8_t res[4];
        status = fdo_match_sample(prep, det_key, h, res);
        if (status == 0u) fdo_state_with_hands_and_reservations(s, h, res);
    }
    P.status[t] = status;
    P.n_nodes[t] = 1u;
    P.explore[t] = 0u; P.result[t] = 0u; P.path_len[t] = 0u;
    uint32_t ctl = 0u;
    if (status == 0u) {
        const uint64_t info = uct_node_info<true>(s, 63u, true);
        P.head(t, 0u).info = info;
        P.state(t, 0u) = s;
        ctl = UCT_CTL_ACTIVE | (uct_n_actions(info) << UCT_CTL_WIDTH_SHIFT);
    }
    P.ctl[t] = ctl;
    return status;
}

// backpropagate (mcts.rs:138-158): every node on the path from `explore` to the root gains one visit and
// result[parent.current_player] — kept in the parent's slot of that node; the parents that keep statistics are the anchors the walk
// down recorded.  On the device a slot is ONE 64-bit reduction (visits in the low word gain 1 and never carry; the exact integer win
// sum in the high word gains the points modulo 2^32): nothing is read back, so a path costs one round trip for its (coalesced) entries
// and then only fire-and-forget traffic.  The read-modify-write form waited for every slot in turn — 60 % of the rollout kernel's
// warp time (profiles/r02_uct_rollout_v4_ncu_summary.json: the compiler cannot prove that path buffer and heads do not alias).
DK_HD void uct_phase_backprop(const UctPool& P, uint64_t t, uint32_t packed) {
    const uint32_t len = P.path_len[t];
#if defined(__CUDA_ARCH__)
    for (uint32_t l0 = 0; l0 < len; l0 += 8u) {
        uint32_t e[8];
#pragma unroll
        for (uint32_t j = 0; j < 8u; ++j) e[j] = l0 + j < len ? P.path_at(t, l0 + j) : 0u;
#pragma unroll
        for (uint32_t j = 0; j < 8u; ++j)
            if (l0 + j < len) {
                UctStat* st = &P.head(t, e[j] & UCT_NODE_MASK).stat[(e[j] >> 24) & 15u];
                const unsigned long long add = 1ull | ((unsigned long long)(uint32_t)uct_unpack_point(packed, e[j] >> 28) << 32);
                atomicAdd(reinterpret_cast<unsigned long long*>(st), add);
            }
    }
#else
    for (uint32_t l = 0; l < len; ++l) {
        const uint32_t e = P.path_at(t, l);
        UctStat& st = P.head(t, e & UCT_NODE_MASK).stat[(e >> 24) & 15u];
        st.vis += 1u;
        st.win += uct_unpack_point(packed, e >> 28);
    }
#endif
}

// Everything the walk needs from one node, fetched as ONE batch: the info word and — by the width the parent's entry announced —
// the statistics sectors and the child entries (the entry of the chosen slot is then picked from a small local array: the dependent
// re-read of `child[slot]` after the decision cost 0.6 us per level even from the cache).  The loads are volatile asm on the device: they are issued BEFORE the info word is examined (a
// plain load would be sunk below the test that ends the walk, which made a level three dependent round trips: info, statistics,
// entry — profiles/r02_uct_tree_v4_ncu_summary.json).  Slots that are not loaded read as zero.
DK_HD void uct_load_level(const UctHead& h, uint32_t width, uint64_t& info, uint32_t vis[UCT_MAX_CHILDREN], int32_t win[UCT_MAX_CHILDREN],
                          uint32_t child[UCT_MAX_CHILDREN]) {
#if defined(__CUDA_ARCH__)
    const char* base = reinterpret_cast<const char*>(&h);
    uint32_t lo, hi;
    asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(base));
#pragma unroll
    for (uint32_t g = 0; g < UCT_MAX_CHILDREN / 2u; ++g) {
        uint32_t a = 0u, b = 0u, c = 0u, d = 0u;
        const bool need = g < 2u || (g < 4u ? width > 4u : width > 8u);
        if (need) asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "l"(base + 64u + 16u * g));
        vis[2u * g] = a; win[2u * g] = (int32_t)b; vis[2u * g + 1u] = c; win[2u * g + 1u] = (int32_t)d;
    }
#pragma unroll
    for (uint32_t g = 0; g < UCT_MAX_CHILDREN / 4u; ++g) {
        uint32_t a = 0u, b = 0u, c = 0u, d = 0u;
        const bool need = g < 1u || (g < 2u ? width > 4u : width > 8u);
        if (need) asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "l"(base + 16u + 16u * g));
        child[4u * g] = a; child[4u * g + 1u] = b; child[4u * g + 2u] = c; child[4u * g + 3u] = d;
    }
    info = (uint64_t)lo | ((uint64_t)hi << 32);
#else
    info = h.info;
    for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) {
        const bool need = k < 4u || (k < 8u ? width > 4u : width > 8u);
        vis[k] = need ? h.stat[k].vis : 0u;
        win[k] = need ? h.stat[k].win : 0;
        child[k] = need ? h.child[k] : 0u;
    }
#endif
}

// One level of select_promising_node at a node with nch expanded children (all of its actions): the slot chosen by
// find_best_child.  ONE code path for every node width — the lanes of a warp sit at nodes of different widths, and separate paths
// per width ran one after the other (5.6 of 32 lanes, profiles/r02_uct_select_v2_ncu_summary.json).
DK_HD uint32_t uct_select_level(const uint32_t vis[UCT_MAX_CHILDREN], const int32_t win[UCT_MAX_CHILDREN], uint32_t nch, bool is_root, uint32_t root_visits,
                                double c, const UctTables& T) {
    uint32_t total = 1u;                                  // the node's own visits: the iteration that created it + those through its children
#pragma unroll
    for (uint32_t k = 0; k < UCT_MAX_CHILDREN; ++k) total += k < nch ? vis[k] : 0u;
    const uint32_t n = is_root ? root_visits : total;
    return uct_best_slot<UCT_MAX_CHILDREN>(nch, vis, win, n, c, T.ln, 0.0);
}

// select_promising_node (mcts.rs:45-63) + expand_single (mcts.rs:65-104) of one iteration.
// Selection: down from the root while the node has children and nothing left to expand; root_visits = number of iterations
// backpropagated so far (every iteration adds one visit to the root).  Expansion at the node reached: a random unexpanded action
// (SITE_EXPAND word 0 of the iteration's stream), by_action on the node's record, the child appended.  Leaves `explore`, the path of
// anchors and the ctl flags (rollout needed, or the result is the terminal record's points) for the rollout / backpropagation phase.
// expand_single (mcts.rs:65-104) at the node the walk ended in (`info` = its info word, `len` anchors on the path, `last_entry` = the
// last of them) + the hand-over to the rollout / backpropagation phase.
template <bool IDX>
DK_HD void uct_phase_expand(const UctPool& P, uint64_t t, uint32_t node, uint64_t info, uint32_t len, uint32_t last_entry, const RngKey& key, uint32_t ctl_keep) {
    const uint32_t n_actions = uct_n_actions(info);
    const uint64_t unexpanded = info & UCT_ACTION_MASK;
    uint32_t explore = node;
    alignas(16) dk_state s = P.state(t, node);
    if (unexpanded != 0ull) {
        const uint32_t nch = uct_n_children(info);
        if (nch >= UCT_MAX_CHILDREN || len >= UCT_MAX_PATH) { P.status[t] = 3u; P.ctl[t] = 0u; return; }     // cannot happen for states reachable by the rules
        const U4 blk = rng_block(key, SITE_EXPAND, 0);
        const uint32_t a = pick_msb_rank64(unexpanded, mulhi(blk.x, popcll(unexpanded)));
        fdo_state_apply<IDX>(s, a);                                                  // by_action
        explore = P.n_nodes[t];
        P.n_nodes[t] = explore + 1u;
        const uint64_t child_info = uct_node_info<IDX>(s, a, false);
        P.head(t, explore).info = child_info;
        P.state(t, explore) = s;
        const uint32_t entry = explore | (uct_n_actions(child_info) << 28);
        if (n_actions > 1u || node == 0u) {                                          // the node is an anchor: the child gets its slot
            UctHead& h = P.head(t, node);
            UctStat zero; zero.vis = 0u; zero.win = 0;
            h.stat[nch] = zero; h.child[nch] = entry;
            h.info = (info & ~(1ull << a) & ~(15ull << 56)) | ((uint64_t)(nch + 1u) << 56);
            P.path_at(t, len++) = uct_path_entry(node, nch, uct_cur(info));
            if (node == 0u) P.root_action[t * 16u + nch] = (uint8_t)a;
        } else {                                                                      // the end of a single-action chain: the chain grows, the
            P.head(t, last_entry & UCT_NODE_MASK).child[(last_entry >> 24) & 15u] = entry;   // anchor's entry moves to its new end
        }
    }
    P.explore[t] = explore;
    P.path_len[t] = len;
    if (st_phase(s) == DK_PHASE_FINISHED) {                                          // nothing to roll out: the rewards are the record's points
        const int32_t p[4] = {s.points[0], s.points[1], s.points[2], s.points[3]};
        P.result[t] = uct_pack_points(p);
        P.ctl[t] = UCT_CTL_ACTIVE | ctl_keep;
    } else {
        P.ctl[t] = UCT_CTL_ACTIVE | UCT_CTL_ROLLOUT | ctl_keep;
    }
}

// The whole tree phase of ONE tree by one thread.
template <bool IDX>
DK_HD void uct_phase_tree(const UctPool& P, uint64_t t, uint32_t root_visits, double c, const UctTables& T, const RngKey& key) {
    uint32_t node = 0u, len = 0u, last_entry = 0u;
    uint32_t width = 15u;                                 // the root's width is not known before its info word: all sectors
    const uint32_t ctl_keep = P.ctl[t] & UCT_CTL_WIDTH_MASK;
    uint64_t info;
    for (;;) {
        const UctHead& h = P.head(t, node);
        uint32_t vis[UCT_MAX_CHILDREN];
        int32_t win[UCT_MAX_CHILDREN];
        uint32_t child[UCT_MAX_CHILDREN];
        uct_load_level(h, width, info, vis, win, child);
        const uint32_t nch = uct_n_children(info);
        if (nch == 0u || (info & UCT_ACTION_MASK) != 0ull) break;
        // (a node with children and nothing to expand is an anchor: entries never point at an expanded single-action node; a
        // single-action ROOT takes slot 0 whatever the numbers are)
        const uint32_t slot = uct_n_actions(info) <= 1u ? 0u : uct_select_level(vis, win, nch, node == 0u, root_visits, c, T);
        if (len >= UCT_MAX_PATH) { P.status[t] = 3u; P.ctl[t] = 0u; return; }
        last_entry = uct_path_entry(node, slot, uct_cur(info));
        P.path_at(t, len++) = last_entry;
        const uint32_t entry = child[slot];
        node = entry & UCT_NODE_MASK;
        width = entry >> 28;
    }
    uct_phase_expand<IDX>(P, t, node, info, len, last_entry, key, ctl_keep);
}

// random_rollout (env_state_full_doko.rs:198-220) from the node chosen by the expansion; SEL12 = the caller's table holds the 12-bit rank select.
template <bool SEL12>
DK_HD uint32_t uct_phase_rollout(const UctPool& P, uint64_t t, const RngKey& key, const uint32_t* __restrict__ lut) {
    alignas(16) dk_state s = P.state(t, P.explore[t]);
    int32_t p[4] = {0, 0, 0, 0};
    FdoLive g; FdoResume rs;
    if (fdo_state_to_live(s, g, rs)) { fdo_play_to_end<false, false, SEL12, false>(g, key, &rs, lut); fdo_final_points(g, p); }
    else { p[0] = s.points[0]; p[1] = s.points[1]; p[2] = s.points[2]; p[3] = s.points[3]; }
    return uct_pack_points(p);
}

// Moves of the root (mcts.rs:220-229) as mcts_policy.rs:96-118 consumes them: visits / values by action index and the move with
// the most visits (max_by_key keeps the LAST maximum in child order); returns ACTION 0xFF when the root has no child.
template <class VisitT>
DK_HD uint32_t uct_moves(const UctPool& P, uint64_t t, VisitT* __restrict__ visits, float* __restrict__ values) {
    const UctHead& root = P.head(t, 0u);
    uint32_t best = 0xFFu, best_visits = 0;
    const uint32_t nch = uct_n_children(root.info);
    for (uint32_t k = 0; k < nch; ++k) {
        const uint32_t a = P.root_action[t * 16u + k], v = root.stat[k].vis;
        if (visits) visits[a] = (VisitT)v;
        if (values) values[a] = (float)dk_ddiv((double)root.stat[k].win, (double)v);
        if (best == 0xFFu || v >= best_visits) { best = a; best_visits = v; }
    }
    return best;
}

}  // namespace dk
