// selfplay_kernels.cuh — kernels of the lock-step self-play driver (see selfplay.cuh).
#pragma once
#include "kernels.cuh"
#include "selfplay.cuh"
#include "uct.cuh"

namespace dk {

struct SpBuffers {
    long long* states;    // [capacity][311]
    float* policy;        // [capacity][39]
    float* value;         // [capacity][4]
    uint8_t* player;      // [capacity]
    uint32_t* game;       // [capacity]
    unsigned long long capacity;
};

constexpr int SP_THREADS = ENC_THREADS;   // plan / encode / apply use the same game → block mapping

__global__ void __launch_bounds__(SP_THREADS)
sp_plan_kernel(RngParams rp, uint64_t n, const dk_state* __restrict__ states, uint64_t az_epoch, float keep_prob, uint32_t search_forced,
               uint64_t* __restrict__ allowed_out, uint8_t* __restrict__ flags_out, uint32_t* __restrict__ block_counts) {
    __shared__ uint4 stage[SP_THREADS * 8];
    const uint64_t first = (uint64_t)blockIdx.x * SP_THREADS, i = first + threadIdx.x;
    StateStage<SP_THREADS>::load(states, first, n, stage);
    uint32_t flags = SP_DONE;
    uint64_t allowed = 0;
    if (i < n) {
        alignas(16) dk_state s;
        StateStage<SP_THREADS>::get(stage, s);
        if (st_phase(s) != DK_PHASE_FINISHED) {
            allowed = sp_az_allowed(s, az_epoch);
            flags = 0;
            if (popcll(allowed) == 1u && !search_forced) {
                flags = SP_FORCED;
                RngKey key = make_key(rp, i, 0, false);
                U4 blk = rng_block(key, SITE_KEEP, 0);
                if (sp_keep_draw(blk.x) < keep_prob) flags |= SP_KEPT;
            } else flags = SP_KEPT;
        }
        allowed_out[i] = allowed;
        flags_out[i] = (uint8_t)flags;
    }
    int c = __syncthreads_count((flags & SP_KEPT) != 0u);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)c;
}

// Single block: block_offsets[b] = rows already recorded + sum of counts of the blocks before b; *count advances (saturating at the
// capacity — rows beyond it are dropped by sp_encode and counted in *dropped).
__global__ void __launch_bounds__(1024)
sp_scan_kernel(uint32_t n_blocks, const uint32_t* __restrict__ block_counts, unsigned long long* __restrict__ block_offsets,
               unsigned long long* __restrict__ count, unsigned long long* __restrict__ dropped, unsigned long long capacity) {
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = *count;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_blocks; base += 1024u) {
        uint32_t b = base + threadIdx.x;
        unsigned long long v = b < n_blocks ? block_counts[b] : 0ull, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o); if ((int)lane >= o) x += y; }
        if (lane == 31u) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sum[lane], z = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, z, o); if ((int)lane >= o) z += y; }
            warp_sum[lane] = z - w;                                  // exclusive prefix of the warp totals
        }
        __syncthreads();
        unsigned long long excl = carry + warp_sum[warp] + (x - v);
        if (b < n_blocks) block_offsets[b] = excl;
        __syncthreads();
        if (threadIdx.x == 1023u) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned long long total = carry;                              // rows before this turn (<= capacity) + rows kept in it
        if (total > capacity) { *dropped += total - capacity; total = capacity; }
        *count = total;
    }
}

__global__ void __launch_bounds__(SP_THREADS)
sp_encode_kernel(uint64_t n, const dk_state* __restrict__ states, const uint64_t* __restrict__ allowed, uint8_t* __restrict__ flags_io,
                 const unsigned long long* __restrict__ block_offsets, SpBuffers buf, long long* __restrict__ row_out, bool dense_ok) {
    __shared__ uint32_t tok[SP_THREADS * PI_ROW];
    __shared__ uint32_t warp_count[SP_THREADS / 32];
    const uint64_t first = (uint64_t)blockIdx.x * SP_THREADS;
    const uint64_t i = first + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t flags = i < n ? flags_io[i] : SP_DONE;
    const bool kept = (flags & SP_KEPT) != 0u;
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, kept);
    if (lane == 0) warp_count[warp] = popc(ballot);
    __syncthreads();
    uint32_t before = popc(ballot & ((1u << lane) - 1u));
    for (uint32_t w = 0; w < warp; ++w) before += warp_count[w];
    long long row = -1;
    if (kept) {
        unsigned long long r = block_offsets[blockIdx.x] + before;
        if (r < buf.capacity) row = (long long)r;
        else { flags = (flags & ~SP_KEPT) | SP_DROPPED; flags_io[i] = (uint8_t)flags; }
    }
    if (i < n && row_out) row_out[i] = row;
    if (row >= 0) {
        alignas(16) dk_state s;
        load_state(states + i, s);
        SmemSlotOut o{tok + before * PI_ROW};                         // staged by RANK: the block's kept rows are consecutive in smem and in the buffer
        fdo_encode_pi(s, o);
        buf.player[row] = (uint8_t)(st_phase(s) == DK_PHASE_FINISHED ? 0u : st_cur(s));
        buf.game[row] = (uint32_t)i;
        if (flags & SP_FORCED) {                                      // one-hot policy target of a forced move (self_play.rs:85-86)
            const uint32_t a = ffs0ll(allowed[i]);
            float* p = buf.policy + (size_t)row * N_ACTIONS;
            for (uint32_t k = 0; k < N_ACTIONS; ++k) p[k] = k == a ? 1.0f : 0.0f;
        }
    }
    __syncthreads();
    // rows block_offsets[b] .. + kept-in-block - 1, cut at the capacity
    uint32_t cnt = 0;
    for (uint32_t w = 0; w < SP_THREADS / 32; ++w) cnt += warp_count[w];
    const unsigned long long row0 = block_offsets[blockIdx.x];
    if (row0 >= buf.capacity) return;
    cnt = (uint32_t)min((unsigned long long)cnt, buf.capacity - row0);
    if (dense_ok) write_rows_pi_dense_at(tok, row0, cnt, buf.states);
    else for (uint32_t r = warp; r < cnt; r += SP_THREADS / 32) write_row_pi(tok + r * PI_ROW, buf.states + (row0 + r) * 311u, lane);
}

// Stand-in for the search slot (tests / benches): one draw over the allowed set (MSB-first rank pick like FdoAllowedActions::random,
// SITE_STEP word 0) and the uniform distribution as policy target, for every game that is neither finished nor forced.
__global__ void __launch_bounds__(STATE_THREADS)
sp_uniform_search_kernel(RngParams rp, uint64_t n, const uint64_t* __restrict__ allowed, const uint8_t* __restrict__ flags,
                         float* __restrict__ policy, uint8_t* __restrict__ action) {
    __shared__ uint64_t mask_s[STATE_THREADS];
    __shared__ float prob_s[STATE_THREADS];
    const uint64_t first = (uint64_t)blockIdx.x * STATE_THREADS;
    const uint64_t i = first + threadIdx.x;
    uint64_t m = 0;
    float u = 0.0f;
    if (i < n) {
        const uint32_t f = flags[i];
        if (f & (SP_DONE | SP_FORCED)) action[i] = (uint8_t)ACTION_NONE;
        else {
            m = allowed[i];
            const uint32_t cnt = popcll(m);
            RngKey key = make_key(rp, i, 0, false);
            U4 blk = rng_block(key, SITE_STEP, 0);
            action[i] = (uint8_t)pick_msb_rank64(m, mulhi(blk.x, cnt));
            u = f32_div(1.0f, (float)cnt);
        }
    }
    mask_s[threadIdx.x] = m;
    prob_s[threadIdx.x] = u;
    __syncthreads();
    // the block's 128 x 39 policy values are contiguous: consecutive threads write consecutive floats (rows of finished / forced games stay untouched)
    const uint32_t games = (uint32_t)min((uint64_t)STATE_THREADS, n - first);
    float* dst = policy + first * N_ACTIONS;
    for (uint32_t e = threadIdx.x; e < games * N_ACTIONS; e += STATE_THREADS) {
        const uint32_t g = e / N_ACTIONS, a = e - g * N_ACTIONS;
        const uint64_t mg = mask_s[g];
        if (mg) dst[e] = ((mg >> a) & 1ull) ? prob_s[g] : 0.0f;
    }
}

__global__ void __launch_bounds__(STATE_THREADS)
sp_apply_kernel(uint64_t n, dk_state* __restrict__ states, const uint64_t* __restrict__ allowed, const uint8_t* __restrict__ flags,
                const long long* __restrict__ rows, const float* __restrict__ policy, const uint8_t* __restrict__ action, SpBuffers buf,
                uint8_t* __restrict__ err_out) {
    __shared__ long long row_s[STATE_THREADS];
    __shared__ uint4 stage[STATE_THREADS * 8];
    const uint64_t first = (uint64_t)blockIdx.x * STATE_THREADS;
    const uint64_t i = first + threadIdx.x;
    StateStage<STATE_THREADS>::load(states, first, n, stage);
    long long copy_row = -1;
    if (i < n) {
        const uint32_t f = flags[i];
        uint32_t err = 0;
        if (!(f & SP_DONE)) {
            const uint64_t m = allowed[i];
            const uint32_t a = (f & SP_FORCED) ? ffs0ll(m) : action[i];
            if (a >= N_ACTIONS || !((m >> a) & 1ull)) err = 1;             // the reference would panic in play_action
            else {
                if (!(f & SP_FORCED)) copy_row = rows[i];
                alignas(16) dk_state s;
                StateStage<STATE_THREADS>::get(stage, s);
                fdo_state_apply(s, a);
                StateStage<STATE_THREADS>::put(stage, s);
            }
        }
        if (err_out) err_out[i] = (uint8_t)err;
    }
    row_s[threadIdx.x] = copy_row;
    StateStage<STATE_THREADS>::store(states, first, n, stage);       // (syncs first: row_s is visible afterwards as well)
    // policy targets of the searched games → their rows (rows of one block are consecutive): coalesced reads and writes
    const uint32_t games = (uint32_t)min((uint64_t)STATE_THREADS, n - first);
    const float* src = policy + first * N_ACTIONS;
    for (uint32_t e = threadIdx.x; e < games * N_ACTIONS; e += STATE_THREADS) {
        const uint32_t g = e / N_ACTIONS, a = e - g * N_ACTIONS;
        const long long row = row_s[g];
        if (row >= 0) buf.policy[(size_t)row * N_ACTIONS + a] = src[e];
    }
}

// Value targets of rows [0, count): rewards of the row's (finished) game rotated to the row's mover.  Rows of unfinished games
// are left untouched and reported through *unfinished.
__global__ void __launch_bounds__(STATE_THREADS)
sp_finalize_kernel(const unsigned long long* __restrict__ count, const dk_state* __restrict__ states, SpBuffers buf,
                   unsigned long long* __restrict__ unfinished) {
    const unsigned long long total = *count;
    for (unsigned long long r = (unsigned long long)blockIdx.x * STATE_THREADS + threadIdx.x; r < total; r += (unsigned long long)gridDim.x * STATE_THREADS) {
        const dk_state* s = states + buf.game[r];
        const uint32_t meta = __ldg(&s->meta);
        if ((meta & 3u) != DK_PHASE_FINISHED) { atomicAdd(unfinished, 1ull); continue; }
        const uint32_t packed = __ldg(reinterpret_cast<const uint32_t*>(s->points));
        const uint32_t player = buf.player[r];
        float4 v;
        v.x = (float)(int8_t)(packed >> (8u * ((player + 0u) & 3u))) / 8.0f;
        v.y = (float)(int8_t)(packed >> (8u * ((player + 1u) & 3u))) / 8.0f;
        v.z = (float)(int8_t)(packed >> (8u * ((player + 2u) & 3u))) / 8.0f;
        v.w = (float)(int8_t)(packed >> (8u * ((player + 3u) & 3u))) / 8.0f;
        reinterpret_cast<float4*>(buf.value)[r] = v;
    }
}

// The reference's on-disk experience record (DBRecord through bincode 1.3.3, experience_replay_buffer3.rs:11-20,94-121):
//   u64 311 | 311 x i64 | u64 4 | 4 x f32 | u64 39 | 39 x f32 = 2684 bytes = 671 little-endian 32-bit words, records back to back.
// Pure byte shuffling: one thread per output word, consecutive threads write consecutive words.
constexpr uint32_t REPLAY_WORDS = 671u;
__global__ void __launch_bounds__(256)
pack_replay_records_kernel(unsigned long long n_rows, const long long* __restrict__ states, const float* __restrict__ value,
                           const float* __restrict__ policy, uint32_t* __restrict__ out) {
    const unsigned long long total = n_rows * REPLAY_WORDS;
    for (unsigned long long g = (unsigned long long)blockIdx.x * 256u + threadIdx.x; g < total; g += (unsigned long long)gridDim.x * 256u) {
        const unsigned long long r = g / REPLAY_WORDS;
        const uint32_t w = (uint32_t)(g - r * REPLAY_WORDS);
        uint32_t v;
        if (w < 2u) v = w == 0u ? 311u : 0u;
        else if (w < 624u) {
            const uint32_t e = (w - 2u) >> 1;
            const unsigned long long x = (unsigned long long)__ldg(states + r * 311u + e);
            v = (w & 1u) ? (uint32_t)(x >> 32) : (uint32_t)x;           // w - 2 even ⇔ w even ⇒ low half
        } else if (w < 626u) v = w == 624u ? 4u : 0u;
        else if (w < 630u) v = __float_as_uint(__ldg(value + r * 4u + (w - 626u)));
        else if (w < 632u) v = w == 630u ? 39u : 0u;
        else v = __float_as_uint(__ldg(policy + r * 39u + (w - 632u)));
        out[g] = v;
    }
}

// N3: one UCT tree per thread.  Tree t = root (t / trees_per_root), sub-index d = t % trees_per_root; with `determinize` the root state
// is first replaced by determinization (first_sub + d) of the info-state (the dk_determinize stream).  Iteration `it` runs on the
// Philox unit (first_id + root, (first_sub + d) * iterations + it).
constexpr int UCT_THREADS = 64;
#ifndef DK_UCT_MIN_BLOCKS
#define DK_UCT_MIN_BLOCKS 16
#endif
__global__ void __launch_bounds__(UCT_THREADS, DK_UCT_MIN_BLOCKS)
fdo_uct_kernel(RngParams rp, uint64_t n_trees, uint32_t trees_per_root, uint32_t iterations, double c, const double* __restrict__ ln_table, int determinize,
               const dk_state* __restrict__ states, UctNode* __restrict__ pool_base, uint32_t* __restrict__ visits_out, float* __restrict__ values_out,
               uint8_t* __restrict__ action_out, uint8_t* __restrict__ status_out) {
    __shared__ uint32_t lut[CARD_LUT_WORDS];
    stage_card_lut(lut);
    const uint64_t t = (uint64_t)blockIdx.x * UCT_THREADS + threadIdx.x;
    if (t >= n_trees) return;
    const uint64_t root = t / trees_per_root;
    const uint32_t sub = rp.first_sub + (uint32_t)(t - root * trees_per_root);
    UctNode* pool = pool_base + t * ((uint64_t)iterations + 1ull);
    alignas(16) dk_state s;
    load_state(states + root, s);
    uint32_t status = 0;
    if (determinize && st_phase(s) != DK_PHASE_FINISHED) {
        MatchPrep prep;
        fdo_match_prepare(s, prep);
        uint64_t h[4];
        uint8_t res[4];
        status = fdo_match_sample(prep, make_key(rp, root, sub, true), h, res);
        if (status == 0u) fdo_state_with_hands_and_reservations(s, h, res);
    }
    uint32_t best = 0xFFu;
    if (visits_out) for (uint32_t a = 0; a < N_ACTIONS; ++a) visits_out[t * N_ACTIONS + a] = 0u;
    if (values_out) for (uint32_t a = 0; a < N_ACTIONS; ++a) values_out[t * N_ACTIONS + a] = 0.0f;
    if (status == 0u) {
        uct_init_node(pool[0], s, UCT_NONE, 63u, true);
        uint32_t n_nodes = 1;
        for (uint32_t it = 0; it < iterations; ++it) {
            RngKey key = make_key(rp, root, sub * iterations + it, true);
            if (uct_iteration(pool, n_nodes, key, c, ln_table, lut)) { status = 3u; break; }
        }
        best = uct_moves(pool, visits_out ? visits_out + t * N_ACTIONS : (uint32_t*)nullptr, values_out ? values_out + t * N_ACTIONS : (float*)nullptr);
    }
    if (action_out) action_out[t] = (uint8_t)best;
    if (status_out) status_out[t] = (uint8_t)status;
}

}  // namespace dk
