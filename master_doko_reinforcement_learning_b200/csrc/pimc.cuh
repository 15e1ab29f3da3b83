// pimc.cuh — the fuse step of the PIMC move decision, per root, on device.
//   PolicyFusionMaxN::fuse / PolicyFusionAverageStrategy::fuse (rs-doko-py-bridge/src/compare_impi/policy_fusion.rs:18-123)
//   as called from DefaultImpiPolicy::execute (compare_impi.rs:318-357): only the successful samples are fused.
// Rows are uint32 visit counts [n_rows][39] (the reference holds usize; values here are rollout counts < 2^32).
#pragma once
#include "dk_common.cuh"

namespace dk {

constexpr uint32_t N_ACTIONS = 39u;        // FdoAction::COUNT
constexpr uint32_t ACTION_NONE = 0xFFu;
constexpr uint32_t ROOT_STATS = 80u;       // int64 per root: [0,39) MaxN rank sums, [39,78) visit sums, [78] successful samples, [79] 0

DK_HD float f32_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float q = a / b;
    return q;
#endif
}
DK_HD float f32_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float q = a + b;
    return q;
#endif
}

// 1-based rank of allowed action a inside one row under the reference's STABLE sort_by_key(Reverse(visits)) over the allowed
// actions in index order (policy_fusion.rs:44-62): actions with more visits come first, equal visits keep index order.
DK_HD uint32_t fuse_rank(const uint32_t* __restrict__ v, uint64_t allowed, uint32_t a) {
    uint32_t rank = 1u, va = v[a];
    for (uint32_t b = 0; b < N_ACTIONS; ++b) {
        if (!((allowed >> b) & 1ull)) continue;
        uint32_t vb = v[b];
        rank += (vb > va || (vb == va && b < a)) ? 1u : 0u;
    }
    return rank;
}

// PolicyFusionMaxN::fuse.  status (nullable): rows with status != 0 are failed samples and are skipped.
DK_HD uint32_t fuse_max_n(const uint32_t* __restrict__ visits, const uint8_t* __restrict__ status, uint32_t n_rows, uint64_t allowed, uint32_t* n_ok_out) {
    uint32_t best = 0, best_sum = 0xFFFFFFFFu, n_ok = 0;
    for (uint32_t r = 0; r < n_rows; ++r) n_ok += (!status || status[r] == 0u) ? 1u : 0u;
    for (uint32_t a = 0; a < N_ACTIONS; ++a) {
        uint32_t sum = 0xFFFFFFFFu;                                            // not allowed ⇒ u32::MAX (:31-38)
        if ((allowed >> a) & 1ull) {
            sum = 0;
            for (uint32_t r = 0; r < n_rows; ++r)
                if (!status || status[r] == 0u) sum += fuse_rank(visits + (size_t)r * N_ACTIONS, allowed, a);
        }
        if (a == 0u || sum < best_sum) { best = a; best_sum = sum; }           // min_by_key keeps the FIRST minimum (:65-70)
    }
    if (n_ok_out) *n_ok_out = n_ok;
    return best;
}

// PolicyFusionAverageStrategy::fuse: f32 division and accumulation in the reference's order; max_by keeps the LAST of equal
// maxima and treats an unordered (NaN) comparison as Equal (:113-119).
DK_HD uint32_t fuse_average(const uint32_t* __restrict__ visits, const uint8_t* __restrict__ status, uint32_t n_rows, uint32_t* n_ok_out) {
    float sum[N_ACTIONS];
    for (uint32_t a = 0; a < N_ACTIONS; ++a) sum[a] = 0.0f;
    uint32_t n_ok = 0;
    for (uint32_t r = 0; r < n_rows; ++r) {
        if (status && status[r] != 0u) continue;
        n_ok++;
        const uint32_t* v = visits + (size_t)r * N_ACTIONS;
        uint64_t total = 0;
        for (uint32_t a = 0; a < N_ACTIONS; ++a) total += v[a];
        float ft = (float)total;
        for (uint32_t a = 0; a < N_ACTIONS; ++a) sum[a] = f32_add(sum[a], f32_div((float)v[a], ft));
    }
    uint32_t best = 0;
    for (uint32_t a = 1; a < N_ACTIONS; ++a) if (!(sum[best] > sum[a])) best = a;
    if (n_ok_out) *n_ok_out = n_ok;
    return best;
}

// Integer root statistics for the sharded decision: additive over determinizations, hence over ranks.
DK_HD void root_stats_accumulate(const uint32_t* __restrict__ visits, const uint8_t* __restrict__ status, uint32_t n_rows, uint64_t allowed,
                                 long long* __restrict__ stats) {
    for (uint32_t r = 0; r < n_rows; ++r) {
        if (status && status[r] != 0u) continue;
        const uint32_t* v = visits + (size_t)r * N_ACTIONS;
        for (uint32_t a = 0; a < N_ACTIONS; ++a) {
            if ((allowed >> a) & 1ull) stats[a] += (long long)fuse_rank(v, allowed, a);
            stats[N_ACTIONS + a] += (long long)v[a];
        }
        stats[2u * N_ACTIONS] += 1;
    }
}
// Decision from the (all-reduced) statistics.  MaxN: identical to fuse_max_n over the union of the rows.  Average: arg-max of the
// summed visits, last of equal maxima — identical to fuse_average whenever every row has the same total (flat MC: n_rollouts) and
// the f32 sums are exact (totals that are powers of two), and the exact version of it otherwise.
DK_HD uint32_t root_stats_pick(uint32_t strategy, const long long* __restrict__ stats, uint64_t allowed) {
    if (stats[2u * N_ACTIONS] == 0 || allowed == 0ull) return ACTION_NONE;
    uint32_t best = 0;
    if (strategy == 0u) {
        long long best_sum = 0;
        bool have = false;
        for (uint32_t a = 0; a < N_ACTIONS; ++a) {
            if (!((allowed >> a) & 1ull)) continue;
            if (!have || stats[a] < best_sum) { have = true; best = a; best_sum = stats[a]; }
        }
    } else {
        for (uint32_t a = 1; a < N_ACTIONS; ++a) if (!(stats[N_ACTIONS + best] > stats[N_ACTIONS + a])) best = a;
    }
    return best;
}

}  // namespace dk
