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

// Begin of a turn in ONE pass over the states (plan + row numbering + encode): every block
//   1. takes the next block number from a ticket (so numbers follow the order in which blocks start — needed by the look-back),
//   2. decides per game: terminal? epoch-filtered allowed set, forced move?, keep-experience draw,
//   3. publishes its number of kept rows and encodes them into shared memory (staged by rank),
//   4. obtains the number of rows kept by all earlier blocks with a decoupled look-back over the published (aggregate | inclusive)
//      words — sums of integers, so the row numbers are the same deterministic turn-major / game-order numbers a sequential driver
//      would produce, whatever the timing —
//   5. writes its rows straight into the experience buffer as sector-aligned streams.
// status[b]: bits 62-63 = 0 nothing yet, 1 the block's own count, 2 count of all blocks up to and including b (plus the rows recorded
// before this turn); zeroed together with the ticket before the launch.
constexpr unsigned long long SP_ST_AGG = 1ull << 62, SP_ST_INCL = 2ull << 62, SP_ST_VALUE = (1ull << 62) - 1ull;
__device__ __forceinline__ unsigned long long sp_ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sp_st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__global__ void __launch_bounds__(SP_THREADS)
sp_begin_kernel(RngParams rp, uint64_t n, const dk_state* __restrict__ states, uint64_t az_epoch, float keep_prob, uint32_t search_forced,
                uint64_t* __restrict__ allowed_out, uint8_t* __restrict__ flags_out, long long* __restrict__ row_out, SpBuffers buf,
                unsigned int* __restrict__ ticket, unsigned long long* __restrict__ status, const unsigned long long* __restrict__ rows_before,
                unsigned long long* __restrict__ rows_after, unsigned long long* __restrict__ count, unsigned long long* __restrict__ dropped,
                bool dense_ok) {
    __shared__ __align__(16) uint32_t tok[SP_THREADS * PI_ROW];           // first the staged state records (16 KB), then the token rows
    __shared__ uint32_t warp_count[SP_THREADS / 32];
    __shared__ uint32_t s_bid;
    __shared__ unsigned long long s_row0;
    __shared__ uint8_t forced_act[SP_THREADS];                            // by rank: the forced move of the row, 0xFF if the row is searched
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint64_t first = (uint64_t)bid * SP_THREADS, i = first + threadIdx.x;
    uint4* stage = reinterpret_cast<uint4*>(tok);
    StateStage<SP_THREADS>::load(states, first, n, stage);
    alignas(16) dk_state s;
    if (i < n) StateStage<SP_THREADS>::get(stage, s);
    __syncthreads();                                                       // the stage is dead: tok may be overwritten from here on
    uint32_t flags = SP_DONE;
    uint64_t allowed = 0;
    if (i < n) {
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
    }
    const bool kept = (flags & SP_KEPT) != 0u;
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, kept);
    if (lane == 0) warp_count[warp] = popc(ballot);
    __syncthreads();
    uint32_t before = popc(ballot & ((1u << lane) - 1u)), cnt = 0;
#pragma unroll
    for (uint32_t w = 0; w < SP_THREADS / 32; ++w) { before += w < warp ? warp_count[w] : 0u; cnt += warp_count[w]; }
    if (threadIdx.x == 0) sp_st_status(status + bid, SP_ST_AGG | cnt);
    if (kept) {
        SmemSlotOut o{tok + before * PI_ROW};                             // staged by RANK: the block's kept rows are consecutive in smem and in the buffer
        fdo_encode_pi(s, o);
        forced_act[before] = (uint8_t)((flags & SP_FORCED) ? ffs0ll(allowed) : 0xFFu);
    }
    if (warp == 0) {                                                       // look-back: 32 predecessors per step, nearest first
        unsigned long long excl = 0;
        long long base = (long long)bid - 1;
        for (;;) {
            const long long idx = base - (long long)lane;
            unsigned long long st;
            if (idx >= 0) { do { st = sp_ld_status(status + idx); } while ((st >> 62) == 0ull); }
            else st = SP_ST_INCL | (idx == -1 ? *rows_before : 0ull);     // virtual block -1: the rows recorded before this turn (a slot
                                                                          //   no block of this launch writes: the last block may be done already)
            const unsigned incl = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2ull);
            const uint32_t stop = incl ? (uint32_t)(__ffs((int)incl) - 1) : 31u;      // lanes 0..stop contribute
            unsigned long long v = lane <= stop ? (st & SP_ST_VALUE) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            excl += v;
            if (incl) break;
            base -= 32;
        }
        if (lane == 0) {
            sp_st_status(status + bid, SP_ST_INCL | (excl + cnt));
            s_row0 = excl;
            if (first + SP_THREADS >= n) {                                 // the last block closes the turn: rows before + rows kept in it
                unsigned long long total = excl + cnt;
                if (total > buf.capacity) { *dropped += total - buf.capacity; total = buf.capacity; }
                *count = total;
                *rows_after = total;
            }
        }
    }
    __syncthreads();
    const unsigned long long row0 = s_row0;
    long long row = -1;
    if (kept) {
        const unsigned long long r = row0 + before;
        if (r < buf.capacity) row = (long long)r;
        else flags = (flags & ~SP_KEPT) | SP_DROPPED;
    }
    if (i < n) { flags_out[i] = (uint8_t)flags; if (row_out) row_out[i] = row; }
    if (row >= 0) {
        buf.player[row] = (uint8_t)(st_phase(s) == DK_PHASE_FINISHED ? 0u : st_cur(s));
        buf.game[row] = (uint32_t)i;
    }
    if (row0 >= buf.capacity) return;
    cnt = (uint32_t)min((unsigned long long)cnt, buf.capacity - row0);
    // one-hot policy targets of the forced moves (self_play.rs:85-86): the block's rows are consecutive, so consecutive threads write
    // consecutive floats (the rows of searched games are filled by sp_apply)
    float* pol = buf.policy + (size_t)row0 * N_ACTIONS;
    for (uint32_t e = threadIdx.x; e < cnt * N_ACTIONS; e += SP_THREADS) {
        const uint32_t r = e / N_ACTIONS, a = e - r * N_ACTIONS, fa = forced_act[r];
        if (fa != 0xFFu) pol[e] = a == fa ? 1.0f : 0.0f;
    }
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
// Pure byte shuffling.  Four consecutive records are exactly 671 x 16 bytes, so a block writes groups of four records as 671 aligned
// 16-byte stores (one per thread and pass; 224 threads = three nearly full passes); the rows left over at the end go word by word.
constexpr uint32_t REPLAY_WORDS = 671u;
constexpr int REPLAY_THREADS = 224;
__device__ __forceinline__ uint32_t replay_word(unsigned long long r, uint32_t w, const long long* __restrict__ states, const float* __restrict__ value,
                                                const float* __restrict__ policy) {
    if (w < 2u) return w == 0u ? 311u : 0u;
    if (w < 624u) {
        const unsigned long long x = (unsigned long long)__ldg(states + r * 311u + ((w - 2u) >> 1));
        return (w & 1u) ? (uint32_t)(x >> 32) : (uint32_t)x;             // w - 2 even ⇔ w even ⇒ low half
    }
    if (w < 626u) return w == 624u ? 4u : 0u;
    if (w < 630u) return __float_as_uint(__ldg(value + r * 4u + (w - 626u)));
    if (w < 632u) return w == 630u ? 39u : 0u;
    return __float_as_uint(__ldg(policy + r * 39u + (w - 632u)));
}
// Groups of four records: the TMA engine brings the group's source words into shared memory — three bulk copies (cp.async.bulk:
// 4 observation rows = 9952 B, 4 value rows = 64 B, 4 policy rows = 624 B; all 16-byte multiples at 16-byte aligned addresses)
// that complete on an mbarrier, two staging buffers so the next group is in flight while the current one is written.  The record
// image is the same for every group, so each thread computes ONCE where in the staging buffer the twelve words of its three
// 16-byte output chunks come from (the six length words of a record are four constant slots behind the data); the loop body is
// twelve shared-memory loads and three aligned 16-byte stores — no global loads on the LSU path at all.
// History (2^21 rows, profiles/r01_n4_bench.json): every chunk from two or three overlapping 8-byte global loads 2.35 ms; records
// assembled in shared memory by 4-byte cp.async (LDGSTS) copies 2.24 ms — ncu: 84 LDGSTS warp instructions per group at 8 cycles
// each saturate the MIO queue (mio_throttle 11.4 warps per issue), 40 % excess L2->L1 sectors from sector-misaligned 128-byte requests.
constexpr uint32_t REPLAY_STAGE_WORDS = 2664u;               // 2488 observation + 16 value + 156 policy + 4 constants (311, 0, 4, 39)
__device__ __forceinline__ uint32_t replay_source_slot(uint32_t q) {      // word q of the four-record image -> staging slot
    const uint32_t r = q / REPLAY_WORDS, w = q - r * REPLAY_WORDS;
    if (w < 2u) return w == 0u ? 2660u : 2661u;
    if (w < 624u) return r * 622u + (w - 2u);
    if (w < 626u) return w == 624u ? 2662u : 2661u;
    if (w < 630u) return 2488u + 4u * r + (w - 626u);
    if (w < 632u) return w == 630u ? 2663u : 2661u;
    return 2504u + 39u * r + (w - 632u);
}
__device__ __forceinline__ void replay_tma_fetch(unsigned long long g, uint32_t stage, uint32_t bar, const long long* __restrict__ states,
                                                 const float* __restrict__ value, const float* __restrict__ policy) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(9952u + 64u + 624u) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(stage), "l"(states + g * 4ull * 311ull), "r"(9952u), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(stage + 9952u), "l"(value + g * 16ull), "r"(64u), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(stage + 10016u), "l"(policy + g * 156ull), "r"(624u), "r"(bar) : "memory");
}
__device__ __forceinline__ bool replay_mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0u;
}
__global__ void __launch_bounds__(REPLAY_THREADS)
pack_replay_records_kernel(unsigned long long n_rows, const long long* __restrict__ states, const float* __restrict__ value,
                           const float* __restrict__ policy, uint32_t* __restrict__ out, bool aligned16) {
    __shared__ __align__(128) uint32_t stage[2][REPLAY_STAGE_WORDS];
    __shared__ __align__(8) unsigned long long full[2];
    const unsigned long long n_groups = aligned16 ? n_rows / 4ull : 0ull;
    if (blockIdx.x < n_groups) {
        const uint32_t st0 = (uint32_t)__cvta_generic_to_shared(stage[0]), st1 = (uint32_t)__cvta_generic_to_shared(stage[1]);
        const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&full[0]), bar1 = (uint32_t)__cvta_generic_to_shared(&full[1]);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (threadIdx.x < 8u) {                                  // the constant slots of both buffers
            const uint32_t k = threadIdx.x & 3u;
            stage[threadIdx.x >> 2][2660u + k] = k == 0u ? 311u : (k == 1u ? 0u : (k == 2u ? 4u : 39u));
        }
        uint32_t slot[3][4];
#pragma unroll
        for (uint32_t c = 0; c < 3u; ++c) {
            const uint32_t e = min(threadIdx.x + c * REPLAY_THREADS, REPLAY_WORDS - 1u);
#pragma unroll
            for (uint32_t k = 0; k < 4u; ++k) slot[c][k] = replay_source_slot(4u * e + k);
        }
        __syncthreads();                                         // barriers initialised, constants written
        unsigned long long g = blockIdx.x;
        if (threadIdx.x == 0) replay_tma_fetch(g, st0, bar0, states, value, policy);
        for (uint32_t it = 0; g < n_groups; g += gridDim.x, ++it) {
            const uint32_t b = it & 1u, parity = (it >> 1) & 1u;
            __syncthreads();                                     // nobody still reads the other buffer (previous iteration)
            const unsigned long long next = g + gridDim.x;
            if (threadIdx.x == 0 && next < n_groups) replay_tma_fetch(next, b ? st0 : st1, b ? bar0 : bar1, states, value, policy);
            while (!replay_mbar_try_wait(b ? bar1 : bar0, parity)) {}
            const uint32_t* __restrict__ src = stage[b];
            uint4* dst = reinterpret_cast<uint4*>(out + g * 4ull * REPLAY_WORDS);
#pragma unroll
            for (uint32_t c = 0; c < 3u; ++c) {
                const uint32_t e = threadIdx.x + c * REPLAY_THREADS;
                if (e < REPLAY_WORDS) dst[e] = make_uint4(src[slot[c][0]], src[slot[c][1]], src[slot[c][2]], src[slot[c][3]]);
            }
        }
    }
    // tail rows (and everything when the output is not 16-byte aligned): one word per thread
    const unsigned long long row0 = n_groups * 4ull, tail_words = (n_rows - row0) * REPLAY_WORDS;
    for (unsigned long long t = (unsigned long long)blockIdx.x * REPLAY_THREADS + threadIdx.x; t < tail_words; t += (unsigned long long)gridDim.x * REPLAY_THREADS) {
        const unsigned long long r = t / REPLAY_WORDS;
        out[row0 * REPLAY_WORDS + t] = replay_word(row0 + r, (uint32_t)(t - r * REPLAY_WORDS), states, value, policy);
    }
}

// N3: one UCT tree per thread, one kernel per phase (uct.cuh); the per-iteration kernels work on a range [t_begin, t_end) of the trees, so
// that the host can run parts of a batch on different streams (the tree phase of one part overlaps the rollouts of another).  Tree t = root (t / trees_per_root), sub-index d = t % trees_per_root;
// with `determinize` the root state is first replaced by determinization (first_sub + d) of the info-state (the dk_determinize stream).
// Iteration `it` runs on the Philox unit (first_id + root, (first_sub + d) * iterations + it).
#ifndef DK_UCT_TREE_BLOCKS
#define DK_UCT_TREE_BLOCKS 6        // 80 registers, no spills: since the batch runs in parts (cabi.cu) a part still fits one wave (+2-4 %)
#endif
#ifndef DK_UCT_EXPAND_IDX
#define DK_UCT_EXPAND_IDX true      // the expansion's record in local memory (128 B, no spills): the register form spilled 700 B at 64 registers
#endif
#ifndef DK_UCT_ROLLOUT_BLOCKS
#define DK_UCT_ROLLOUT_BLOCKS 3      // 80 registers, no spills (as one launch over 131 072 trees 4 blocks of 64 registers were 7 % faster: 1.15 waves; in parts every launch is one wave)
#endif
constexpr int UCT_THREADS = 128;
__device__ __forceinline__ RngKey uct_iteration_key(const RngParams& rp, uint64_t t, uint32_t trees_per_root, uint32_t iterations, uint32_t it) {
    const uint64_t root = t / trees_per_root;
    const uint32_t sub = rp.first_sub + (uint32_t)(t - root * trees_per_root);
    return make_key(rp, root, sub * iterations + it, true);
}
__global__ void __launch_bounds__(UCT_THREADS)
uct_root_kernel(RngParams rp, UctPool P, uint32_t trees_per_root, int determinize, const dk_state* __restrict__ states, uint32_t* __restrict__ visits_out,
                float* __restrict__ values_out) {
    const uint64_t t = (uint64_t)blockIdx.x * UCT_THREADS + threadIdx.x;
    if (t >= P.n_trees) return;
    const uint64_t root = t / trees_per_root;
    const uint32_t sub = rp.first_sub + (uint32_t)(t - root * trees_per_root);
    alignas(16) dk_state s;
    load_state(states + root, s);
    uct_phase_root(P, t, s, determinize != 0, make_key(rp, root, sub, true));
    if (visits_out) for (uint32_t a = 0; a < N_ACTIONS; ++a) visits_out[t * N_ACTIONS + a] = 0u;
    if (values_out) for (uint32_t a = 0; a < N_ACTIONS; ++a) values_out[t * N_ACTIONS + a] = 0.0f;
}
// select(it) + expand(it): thread = tree for the whole phase (uct_phase_tree).
// Tried and rejected (profiles/r02_uct_tree_v8_packed_levels_ncu_summary.json): a LEVEL-SYNCHRONOUS walk inside the block — after
// every level the trees still walking are packed densely onto the first threads, ordered by the width class of their node (4 / 8 / 12
// child slots, evaluated by class-specialised code), per-tree walk state in shared memory.  It does what it promises for the lanes
// (13.5 -> 18.6 of 32) and the instruction count (29.9 M -> 26.4 M per launch of 131 072 trees), but every level of a block now ends
// in two barriers behind a DRAM round trip: 8.1 warps per issue slot wait at the barrier, issue utilisation drops from 40 % to 25 %,
// the kernel goes from 85 to 110 us and the search from 1.09e9 to 0.92e9 iterations/s.  Sorting the trees once per launch by their
// previous path length instead was measured on the host simulator: 0.63 -> 0.67 of the lanes (a path's length says little about the
// next one's; a perfect sort would give 0.84).
__global__ void __launch_bounds__(UCT_THREADS, DK_UCT_TREE_BLOCKS)
uct_tree_kernel(RngParams rp, UctPool P, uint32_t trees_per_root, uint32_t iterations, uint32_t it, double c, UctTables T, uint64_t t_begin, uint64_t t_end) {
    const uint64_t t = t_begin + (uint64_t)blockIdx.x * UCT_THREADS + threadIdx.x;
    if (t >= t_end) return;
    if (!(P.ctl[t] & UCT_CTL_ACTIVE)) return;
    uct_phase_tree<DK_UCT_EXPAND_IDX>(P, t, it, c, T, uct_iteration_key(rp, t, trees_per_root, iterations, it));
}
// rollout(it) + backpropagate(it)
constexpr int UCT_ROLLOUT_THREADS = 256;
__global__ void __launch_bounds__(UCT_ROLLOUT_THREADS, DK_UCT_ROLLOUT_BLOCKS)
uct_rollout_kernel(RngParams rp, UctPool P, uint32_t trees_per_root, uint32_t iterations, uint32_t it, uint64_t t_begin, uint64_t t_end) {
    __shared__ __align__(16) uint32_t lut[CARD_LUT_WORDS + SEL12_WORDS];
    stage_lut(lut, CARD_LUT_WORDS + SEL12_WORDS);
    __syncthreads();
    const uint64_t t = t_begin + (uint64_t)blockIdx.x * UCT_ROLLOUT_THREADS + threadIdx.x;
    if (t >= t_end) return;
    const uint32_t ctl = P.ctl[t];
    if (!(ctl & UCT_CTL_ACTIVE)) return;
    uint32_t packed;
    if (ctl & UCT_CTL_ROLLOUT) packed = uct_phase_rollout<true>(P, t, uct_iteration_key(rp, t, trees_per_root, iterations, it), lut);
    else packed = P.result[t];                                    // a terminal record: its points, left by the tree phase
    uct_phase_backprop(P, t, packed);
}
__global__ void __launch_bounds__(UCT_THREADS)
uct_moves_kernel(UctPool P, uint32_t* __restrict__ visits_out, float* __restrict__ values_out, uint8_t* __restrict__ action_out, uint8_t* __restrict__ status_out) {
    const uint64_t t = (uint64_t)blockIdx.x * UCT_THREADS + threadIdx.x;
    if (t >= P.n_trees) return;
    const uint32_t status = P.status[t];
    uint32_t best = 0xFFu;
    if (status == 0u) best = uct_moves(P, t, visits_out ? visits_out + t * N_ACTIONS : (uint32_t*)nullptr, values_out ? values_out + t * N_ACTIONS : (float*)nullptr);
    if (action_out) action_out[t] = (uint8_t)best;
    if (status_out) status_out[t] = (uint8_t)status;
}

}  // namespace dk
