// kernels.cuh — __global__ entry points (sm_100a).  One thread = one game / sample / rollout.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>

#include "assignment.cuh"
#include "dk_common.cuh"
#include "doko_rules.cuh"
#include "encode.cuh"
#include "fdo_rules.cuh"
#include "matching.cuh"
#include "pimc.cuh"
#include "state_ops.cuh"

namespace dk {

constexpr int PLAYOUT_THREADS = 256;

struct RngParams {
    uint32_t seed_lo, seed_hi;
    uint64_t first_id;
    uint32_t epoch;
    uint32_t first_sub;
};

// Per-thread 12-word scratch in shared memory, word-interleaved across the block so that every access of a
// warp hits 32 distinct banks regardless of the (data-dependent) word index.
template <int THREADS>
struct SharedDeckT {
    uint32_t* base;  // &smem[threadIdx.x]
    __device__ __forceinline__ uint32_t get(uint32_t w) const { return base[w * THREADS]; }
    __device__ __forceinline__ void set(uint32_t w, uint32_t v) { base[w * THREADS] = v; }
    // byte j of the thread's 48: byte (j & 3) of word (j >> 2); every lane stays in its own bank for any j
    __device__ __forceinline__ uint32_t get8(uint32_t j) const { return reinterpret_cast<const uint8_t*>(base)[(j >> 2) * (THREADS * 4) + (j & 3u)]; }
    __device__ __forceinline__ void set8(uint32_t j, uint32_t v) { reinterpret_cast<uint8_t*>(base)[(j >> 2) * (THREADS * 4) + (j & 3u)] = (uint8_t)v; }
};
using SharedDeck = SharedDeckT<PLAYOUT_THREADS>;

// The lookup tables (dk_common.cuh: CARD 4.7 KB | SEL12 32 KB | ANN 1 KB) live in device memory as ONE image, written once per context
// by dk_init from the host evaluation of lut_word / sel12_entry.  A block stages the prefix of the image it reads with ONE bulk copy of
// the TMA engine (cp.async.bulk global -> shared, completion on an mbarrier) issued by thread 0.  The cooperative form — every thread
// copying its share with LDG / STS — was 2.5 % of the fresh-playout kernel's instructions (25 words per thread) and 5 % of the UCT
// rollout kernel's (37 words per thread for ONE rollout), profiles/r02_k2_v5 attribution.
__device__ __align__(16) uint32_t g_lut[FULL_LUT_WORDS];
// stage_lut_begin issues the copy (and ends with a block barrier); the returned handle is polled by stage_lut_wait before the first
// table access — the fresh-game kernels deal their cards in between.
__device__ __forceinline__ uint32_t stage_lut_begin(uint32_t* lut, uint32_t words) {   // lut: 16-byte aligned; words * 4 a multiple of 16
    __shared__ __align__(8) unsigned long long lut_bar;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&lut_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(words * 4u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(lut)), "l"(g_lut), "r"(words * 4u), "r"(bar) : "memory");
    }
    __syncthreads();                                                             // the barrier is initialised before anybody polls it
    return bar;
}
__device__ __forceinline__ void stage_lut_wait(uint32_t bar) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
}
__device__ __forceinline__ void stage_lut(uint32_t* lut, uint32_t words) { stage_lut_wait(stage_lut_begin(lut, words)); }
struct LutReady { uint32_t bar; __device__ __forceinline__ void operator()() const { stage_lut_wait(bar); } };
__device__ __forceinline__ void stage_card_lut(uint32_t* lut) { stage_lut(lut, CARD_LUT_WORDS); }

__device__ __forceinline__ RngKey make_key(const RngParams& rp, uint64_t index, uint32_t unit_hi_override, bool use_override) {
    uint64_t unit = rp.first_id + index;
    RngKey k;
    k.seed_lo = rp.seed_lo; k.seed_hi = rp.seed_hi;
    k.unit_lo = (uint32_t)unit;
    k.unit_hi = use_override ? unit_hi_override : (uint32_t)(unit >> 32);
    k.epoch = rp.epoch;
    return k;
}

// Playout results, selected at run time (uniform branch) by `mode`:
//   OUT_INT32    int32[4] points + uint32 steps (20 B per game)
//   OUT_COMPACT  int8[4] + uint8 — lossless: |points| < 128, steps < 256 (5 B per game over PCIe)
//   OUT_PACKED   uint16 points + uint8 steps (3 B per game; 2 B when the caller does not want the step counts).  The four points of a
//                game take two values (Re seats / Kontra seats, FdoEndOfGameStats::calculate, stats.rs:215-231) and sum to zero, so seat 0's
//                points (int8) plus "which of seats 1..3 score like seat 0" (3 bits) determine all four: decode with dk_unpack_points.
enum : uint32_t { OUT_INT32 = 0u, OUT_COMPACT = 1u, OUT_PACKED = 2u };
__device__ __forceinline__ uint32_t pack_points(const int32_t p[4]) {
    const uint32_t same = (p[1] == p[0] ? 1u : 0u) | (p[2] == p[0] ? 2u : 0u) | (p[3] == p[0] ? 4u : 0u);
    return ((uint32_t)p[0] & 255u) | (same << 8);
}
__device__ __forceinline__ void store_result(void* __restrict__ points, void* __restrict__ steps, uint64_t i, const int32_t p[4], uint32_t s, uint32_t mode) {
    if (mode == OUT_INT32) {
        if (points) reinterpret_cast<int4*>(points)[i] = make_int4(p[0], p[1], p[2], p[3]);
        if (steps) reinterpret_cast<uint32_t*>(steps)[i] = s;
    } else {
        if (points) {
            if (mode == OUT_COMPACT) reinterpret_cast<char4*>(points)[i] = make_char4((signed char)p[0], (signed char)p[1], (signed char)p[2], (signed char)p[3]);
            else reinterpret_cast<uint16_t*>(points)[i] = (uint16_t)pack_points(p);
        }
        if (steps) reinterpret_cast<uint8_t*>(steps)[i] = (uint8_t)s;
    }
}

// Device-side reduction of a playout batch (dk_playout_summary): what an evaluator keeps of a batch of games — counts, sums, sums of
// squares, wins and the histogram of the number of actions — so that the device->host result is 2 KB per call instead of bytes per game
// (the host-side ingest of per-game results is what stops scaling at 8 GPUs per host).  Integer sums: order independent, bit-reproducible.
// Word layout of dk_playout_stats (include/doko_cuda.h): [0] games | [1] game steps | [2,6) point sums | [6,10) sums of squares |
// [10,14) wins | [14,270) step histogram.
constexpr uint32_t STATS_WORDS = 270u;
struct BlockStats { uint32_t w[STATS_WORDS]; };   // per-block partial sums (32 bit: <= 640 games per block; point sums as two's complement)
__device__ __forceinline__ void block_stats_clear(BlockStats& sm) {                 // caller syncs
    for (uint32_t k = threadIdx.x; k < STATS_WORDS; k += blockDim.x) sm.w[k] = 0u;
}
// Every thread of the block calls this (valid = the thread holds a game of the batch); ends with the block's flush to `stats`.
__device__ __forceinline__ void block_stats_add(BlockStats& sm, unsigned long long* __restrict__ stats, const int32_t p[4], uint32_t s, bool valid) {
    const unsigned full = 0xFFFFFFFFu;
    const bool lead = (threadIdx.x & 31u) == 0u;
    const uint32_t g = (uint32_t)__popc(__ballot_sync(full, valid));
    const uint32_t st = __reduce_add_sync(full, valid ? s : 0u);
    if (lead) { atomicAdd(&sm.w[0], g); atomicAdd(&sm.w[1], st); }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int v = valid ? p[q] : 0;
        const int sum = __reduce_add_sync(full, v);
        const uint32_t sq = __reduce_add_sync(full, (uint32_t)(v * v));
        const uint32_t wins = (uint32_t)__popc(__ballot_sync(full, v > 0));
        if (lead) { atomicAdd(&sm.w[2 + q], (uint32_t)sum); atomicAdd(&sm.w[6 + q], sq); atomicAdd(&sm.w[10 + q], wins); }
    }
    if (valid) atomicAdd(&sm.w[14u + (s < 255u ? s : 255u)], 1u);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < STATS_WORDS; k += blockDim.x) {
        const uint32_t v = sm.w[k];
        if (v) atomicAdd(stats + k, (k >= 2u && k < 6u) ? (unsigned long long)(long long)(int32_t)v : (unsigned long long)v);   // point sums are signed
    }
}

// K2: fresh full-rules playouts.  Replaces FdoState::new_game + the random_action loop
// (rs-full-doko/src/state/state.rs:169-178,378-431).  HBM traffic: 0 B in, 16 B points + 4 B steps out per game.
#ifndef DK_FDO_FRESH_THREADS
#define DK_FDO_FRESH_THREADS 384
#endif
// Fresh-game playout kernels (K1, K2): 3 blocks x 384 threads per SM, 51.6 KB dynamic shared memory per block (tables 33.4 KB + shuffle
// scratch 18 KB).  320 x 4, 256 x 4 and 640 x 2 measure within 1 % (profiles/r01_k2_sel12_experiment.txt): the kernels are ALU-pipe bound.
constexpr int FDO_FRESH_THREADS = DK_FDO_FRESH_THREADS;
#ifndef DK_FDO_FRESH_BLOCKS
#define DK_FDO_FRESH_BLOCKS 3
#endif
constexpr uint32_t FDO_FRESH_SMEM_BYTES = 4u * (FULL_LUT_WORDS + 12u * FDO_FRESH_THREADS);   // dynamic: above the 48 KB static limit
template <bool WITH_ANN>
__global__ void __launch_bounds__(FDO_FRESH_THREADS, DK_FDO_FRESH_BLOCKS)
fdo_playout_fresh_kernel(RngParams rp, uint64_t n, void* __restrict__ points, void* __restrict__ steps, uint32_t mode, unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(16) uint32_t fresh_smem[];
    __shared__ BlockStats bstats;
    uint32_t* lut = fresh_smem;                                        // FULL_LUT_WORDS words (16-byte aligned, SEL12 part 8-byte aligned)
    uint32_t* smem = fresh_smem + FULL_LUT_WORDS;                      // the shuffle scratch: 12 words per thread, word-interleaved
    if (stats) block_stats_clear(bstats);
    const LutReady ready{stage_lut_begin(lut, FULL_LUT_WORDS)};        // (ends with a block barrier; the copy runs under the deal)
    uint64_t i = (uint64_t)blockIdx.x * FDO_FRESH_THREADS + threadIdx.x;
    // Out-of-range lanes play game n-1 again (keeps the warp converged); they just do not store.
    uint64_t gi = i < n ? i : n - 1;
    SharedDeckT<FDO_FRESH_THREADS> deck;
    deck.base = smem + threadIdx.x;
    RngKey key = make_key(rp, gi, 0, false);
    int32_t p[4];
    uint32_t s;
    fdo_playout_fresh<WITH_ANN, SharedDeckT<FDO_FRESH_THREADS>, true, LutReady>(key, deck, lut, p, s, ready);
    if (i < n) store_result(points, steps, i, p, s, mode);
    if (stats) block_stats_add(bstats, stats, p, s, i < n);          // uniform branch
}


// K1: fresh simplified-rules playouts (DoState::new_game + 52 random actions, rs-doko/src/state/state.rs:159-168,315-334).
// TRACE additionally writes the 52 action ids and (wedding flag, re mask, packed eyes, packed tricks) per game.
template <bool TRACE>
__global__ void __launch_bounds__(FDO_FRESH_THREADS, DK_FDO_FRESH_BLOCKS)
doko_playout_fresh_kernel(RngParams rp, uint64_t n, void* __restrict__ points, void* __restrict__ steps, uint32_t mode, unsigned long long* __restrict__ stats,
                          uint8_t* __restrict__ trace, uint4* __restrict__ aux) {
    extern __shared__ __align__(16) uint32_t fresh_smem[];
    __shared__ BlockStats bstats;
    uint32_t* lut = fresh_smem;
    uint32_t* smem = fresh_smem + CARD_LUT_WORDS + SEL12_WORDS;
    if (stats) block_stats_clear(bstats);
    const LutReady ready{stage_lut_begin(lut, CARD_LUT_WORDS + SEL12_WORDS)};
    uint64_t i = (uint64_t)blockIdx.x * FDO_FRESH_THREADS + threadIdx.x;
    uint64_t gi = i < n ? i : n - 1;
    SharedDeckT<FDO_FRESH_THREADS> deck;
    deck.base = smem + threadIdx.x;
    RngKey key = make_key(rp, gi, 0, false);
    int32_t p[4];
    uint32_t s, ax[4];
    uint8_t tr[52];
    doko_playout_fresh<TRACE, SharedDeckT<FDO_FRESH_THREADS>, true, LutReady>(key, deck, lut, p, s, tr, ax, ready);
    if (i < n) {
        store_result(points, steps, i, p, s, mode);
        if (TRACE) {
            if (trace) for (int k = 0; k < 52; ++k) trace[i * 52 + k] = tr[k];
            if (aux) aux[i] = make_uint4(ax[0], ax[1], ax[2], ax[3]);
        }
    }
    if (stats) block_stats_add(bstats, stats, p, s, i < n);          // uniform branch
}

// ---- state record I/O --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_state(const dk_state* __restrict__ src, dk_state& dst) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(&dst);
#pragma unroll
    for (int q = 0; q < 8; ++q) d4[q] = __ldg(s4 + q);
}
__device__ __forceinline__ void store_state(dk_state* __restrict__ dst, const dk_state& src) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    const uint4* s4 = reinterpret_cast<const uint4*>(&src);
#pragma unroll
    for (int q = 0; q < 8; ++q) d4[q] = s4[q];
}

constexpr int STATE_THREADS = 128;

// Block-cooperative state I/O: the BLOCK records of a block are contiguous in HBM (BLOCK x 128 B), so the block moves them with
// fully coalesced 16-byte accesses (a warp instruction covers 512 contiguous bytes instead of 32 separate 128-byte lines) and
// every thread then picks its own record out of shared memory.  16-byte chunk j of record t sits at t*8 + (j ^ (t & 7)): both the
// cooperative phase and the per-thread phase are bank-conflict free.  All threads of the block must call load / store.
template <int BLOCK>
struct StateStage {
    static __device__ __forceinline__ void load(const dk_state* __restrict__ states, uint64_t first, uint64_t n, uint4* sm) {
        const uint4* g = reinterpret_cast<const uint4*>(states + first);
        const uint32_t avail = (uint32_t)min((uint64_t)BLOCK, n - first) * 8u;
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k) {
            const uint32_t e = k * BLOCK + threadIdx.x;
            if (e < avail) { const uint32_t t = e >> 3, j = e & 7u; sm[t * 8u + (j ^ (t & 7u))] = __ldg(g + e); }
        }
        __syncthreads();
    }
    static __device__ __forceinline__ void get(const uint4* sm, dk_state& s) { get_row(sm, threadIdx.x, s); }
    static __device__ __forceinline__ void put(uint4* sm, const dk_state& s) { put_row(sm, threadIdx.x, s); }
    // record `t` of the tile (any thread may take any row; rows taken out of order cost a few bank conflicts on the idle LSU pipe)
    static __device__ __forceinline__ void get_row(const uint4* sm, uint32_t t, dk_state& s) {
        uint4* d4 = reinterpret_cast<uint4*>(&s);
#pragma unroll
        for (uint32_t j = 0; j < 8u; ++j) d4[j] = sm[t * 8u + (j ^ (t & 7u))];
    }
    static __device__ __forceinline__ void put_row(uint4* sm, uint32_t t, const dk_state& s) {
        const uint4* s4 = reinterpret_cast<const uint4*>(&s);
#pragma unroll
        for (uint32_t j = 0; j < 8u; ++j) sm[t * 8u + (j ^ (t & 7u))] = s4[j];
    }
    static __device__ __forceinline__ void store(dk_state* __restrict__ states, uint64_t first, uint64_t n, const uint4* sm) {
        __syncthreads();
        uint4* g = reinterpret_cast<uint4*>(states + first);
        const uint32_t avail = (uint32_t)min((uint64_t)BLOCK, n - first) * 8u;
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k) {
            const uint32_t e = k * BLOCK + threadIdx.x;
            if (e < avail) { const uint32_t t = e >> 3, j = e & 7u; g[e] = sm[t * 8u + (j ^ (t & 7u))]; }
        }
    }
};

// dk_new_games: FdoState::new_game / DoState::new_game (deal from the stream) → records.
__global__ void __launch_bounds__(PLAYOUT_THREADS)
new_games_kernel(RngParams rp, uint64_t n, dk_state* __restrict__ out) {
    __shared__ uint32_t smem[12 * PLAYOUT_THREADS];
    uint64_t i = (uint64_t)blockIdx.x * PLAYOUT_THREADS + threadIdx.x;
    uint64_t gi = i < n ? i : n - 1;
    SharedDeck deck;
    deck.base = smem + threadIdx.x;
    RngKey key = make_key(rp, gi, 0, false);
    FdoLive dummy;
    uint32_t ah[4], dup, start;
    fdo_deal(dummy, key, deck, ah, dup, start);
    if (i < n) {
        uint64_t hands[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) hands[p] = (uint64_t)ah[p] | ((uint64_t)(ah[p] & dup) << 24);   // copy B = doubled cards
        alignas(16) dk_state s;
        st_new_game(s, hands, start);
        store_state(out + i, s);
    }
}
__global__ void __launch_bounds__(STATE_THREADS)
from_deals_kernel(uint64_t n, const uint64_t* __restrict__ hands, const uint8_t* __restrict__ start, dk_state* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * STATE_THREADS + threadIdx.x;
    if (i >= n) return;
    uint64_t h[4] = {hands[4 * i], hands[4 * i + 1], hands[4 * i + 2], hands[4 * i + 3]};
    alignas(16) dk_state s;
    st_new_game(s, h, start[i] & 3u);
    store_state(out + i, s);
}
// dk_legal_mask / dk_legal_mask_az: the mask written is legal & ~drop_mask (AzEnvState::allowed_actions_by_action_index removes the calls
// when is_secondary || epoch < MIN_EPOCH), the count is popc(legal & ~drop_count) (number_of_allowed_actions(epoch) looks at the epoch only;
// rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:80-119).
__device__ __forceinline__ void store_legal(uint64_t legal, uint64_t i, uint64_t* __restrict__ mask_out, uint64_t drop_mask, uint64_t drop_count,
                                            uint8_t* __restrict__ count_out) {
    if (mask_out) mask_out[i] = legal & ~drop_mask;
    if (count_out) count_out[i] = (uint8_t)popcll(legal & ~drop_count);
}
template <int ENGINE>
__global__ void __launch_bounds__(STATE_THREADS)
legal_mask_kernel(uint64_t n, const dk_state* __restrict__ states, uint64_t* __restrict__ mask_out, uint64_t drop_mask, uint64_t drop_count,
                  uint8_t* __restrict__ count_out) {
    __shared__ uint4 stage[STATE_THREADS * 8];
    const uint64_t first = (uint64_t)blockIdx.x * STATE_THREADS, i = first + threadIdx.x;
    StateStage<STATE_THREADS>::load(states, first, n, stage);
    if (i >= n) return;
    alignas(16) dk_state s;
    StateStage<STATE_THREADS>::get(stage, s);
    store_legal(ENGINE == DK_FDO ? fdo_state_legal_mask(s) : doko_state_legal_mask(s), i, mask_out, drop_mask, drop_count, count_out);
}
// 8 blocks of 128 threads per SM (64 registers, 8 bytes of spill in the full-rules instance): 0.283 -> 0.243 ms per 2^22 records.
// A two-buffer cp.async pipeline over tiles (bytes in flight independent of the resident thread count) was measured and is
// SLOWER (0.294 ms): the kernel is bound by the transition's instructions, not by memory latency
// (profiles/r01_apply_occupancy_experiment.txt).
#ifndef DK_APPLY_IDX
#define DK_APPLY_IDX false
#endif
#ifndef DK_APPLY_BLOCKS
#define DK_APPLY_BLOCKS 8
#endif
template <int ENGINE>
__global__ void __launch_bounds__(STATE_THREADS, DK_APPLY_BLOCKS)
apply_kernel(uint64_t n, dk_state* __restrict__ states, const uint8_t* __restrict__ action, uint32_t flags, uint8_t* __restrict__ err_out) {
    __shared__ uint4 stage[STATE_THREADS * 8];
    const uint64_t first = (uint64_t)blockIdx.x * STATE_THREADS, i = first + threadIdx.x;
    StateStage<STATE_THREADS>::load(states, first, n, stage);
    if (i < n) {
        alignas(16) dk_state s;
        StateStage<STATE_THREADS>::get(stage, s);
        uint32_t err = ENGINE == DK_FDO ? fdo_state_apply_az(s, action[i], (flags & DK_APPLY_SKIP_SINGLE) != 0) : doko_state_apply(s, action[i]);
        if (!err) StateStage<STATE_THREADS>::put(stage, s);              // an illegal action leaves the record as it was
        if (err_out) err_out[i] = (uint8_t)err;
    }
    StateStage<STATE_THREADS>::store(states, first, n, stage);
}
// ---- record tiles through the TMA engine -------------------------------------------------------------------------------------
// The record array is a [n][128 B] tensor; a tile of 128 records is moved by ONE cp.async.bulk.tensor instruction in each
// direction.  The 128-byte swizzle mode of the tensor map places 16-byte chunk j of record t at t*8 + (j ^ (t & 7)) — exactly
// the layout StateStage builds by hand — so the per-thread get / put stay conflict-free, and the 16 global + 16 shared-memory
// instructions per thread of the cooperative copy disappear from the LSU / MIO path (what limited the record packer as well,
// selfplay_kernels.cuh).  Rows past n are clipped by the engine (zero-filled on load, dropped on store).
struct TmaTile {
    static __device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
    static __device__ __forceinline__ void init(unsigned long long* bar) {          // one thread, before a __syncthreads
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    static __device__ __forceinline__ void load(const CUtensorMap* tmap, uint4* stage, unsigned long long* bar, uint64_t first, uint32_t bytes) {   // one thread
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(saddr(stage)), "l"(tmap), "r"(0), "r"((int)first), "r"(saddr(bar)) : "memory");
    }
    static __device__ __forceinline__ void wait(unsigned long long* bar, uint32_t parity) {                                                      // all threads
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(saddr(bar)), "r"(parity) : "memory");
    }
    static __device__ __forceinline__ void publish() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }    // writers, before the __syncthreads
    static __device__ __forceinline__ void store(const CUtensorMap* tmap, const uint4* stage, uint64_t first) {                                   // one thread, after it
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(0), "r"((int)first), "r"(saddr(stage)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");                // the block's shared memory must outlive the engine's read
    }
};
// The transition's cost is its DIVERGENCE: a tile holds records in every phase, and a warp whose lanes sit in the reservation, the
// announcement and the card phase (with or without a trick to close) runs all of those paths one after the other — 12.7 of 32 lanes,
// ALU pipe 85 %, 0.75 of the HBM copy peak (profiles/r01_final_apply_tma_ncu_summary.json).  The tile is in shared memory anyway, so
// the block SORTS its records by the path they will take (a counting sort over six classes: ballots + one prefix over the warps) and
// thread j applies record perm[j]: warps become (nearly) uniform and the instruction stream shrinks accordingly.
#ifndef DK_APPLY_SORT
#define DK_APPLY_SORT 1
#endif
// Tiles per block.  With the instruction count halved the kernel waited for memory (long_scoreboard + barrier stalls,
// profiles/r02_apply_v2_ncu_summary.json).  What removed the wait was taking the action bytes into shared memory while the tile is in
// flight (after the sort the action of ANOTHER row is needed: a dependent global load on the critical path before): 0.874 -> 0.974 of
// the HBM copy peak.  Two tiles in flight per block (DK_APPLY_TILES=2: 33 KB of shared memory, 6 blocks per SM) measured SLOWER, 0.865.
#ifndef DK_APPLY_TILES
#define DK_APPLY_TILES 1
#endif
template <int ENGINE>
__global__ void __launch_bounds__(STATE_THREADS, DK_APPLY_BLOCKS)
apply_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint64_t n, const uint8_t* __restrict__ action, uint32_t flags, uint8_t* __restrict__ err_out) {
    __shared__ __align__(1024) uint4 stage_all[DK_APPLY_TILES][STATE_THREADS * 8];
    __shared__ __align__(8) unsigned long long bar[DK_APPLY_TILES];
    __shared__ uint32_t cls_count[32];
    __shared__ uint8_t perm[STATE_THREADS];
    __shared__ uint8_t act_s[DK_APPLY_TILES][STATE_THREADS];
    const uint64_t first0 = (uint64_t)blockIdx.x * (STATE_THREADS * DK_APPLY_TILES);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < DK_APPLY_TILES; ++b) TmaTile::init(&bar[b]);
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < DK_APPLY_TILES; ++b) {
        const uint64_t first = first0 + (uint64_t)b * STATE_THREADS;
        if (first >= n) break;
        if (threadIdx.x == 0) TmaTile::load(&tmap, stage_all[b], &bar[b], first, (uint32_t)min((uint64_t)STATE_THREADS, n) * 128u);
        act_s[b][threadIdx.x] = first + threadIdx.x < n ? action[first + threadIdx.x] : (uint8_t)0;   // (in flight with the tile; read after the sort's barriers)
    }
#pragma unroll
    for (int b = 0; b < DK_APPLY_TILES; ++b) {
        const uint64_t first = first0 + (uint64_t)b * STATE_THREADS;
        if (first >= n) break;                                           // uniform
        uint4* stage = stage_all[b];
        TmaTile::wait(&bar[b], 0u);
        uint32_t row = threadIdx.x;
        if (DK_APPLY_SORT) {
            // class of the own row from its last chunk (bytes 112..127: eyes, trick counts, card_index, n_reservations, points, meta)
            const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
            const uint4 last = stage[t * 8u + (7u ^ (t & 7u))];
            const uint32_t phase = last.w & 3u, ci = (last.y >> 16) & 255u;
            uint32_t cls = phase == DK_PHASE_RESERVATION ? 0u : (phase == DK_PHASE_ANNOUNCEMENT ? 1u : (phase == DK_PHASE_PLAY_CARD ? ((ci & 3u) == 3u ? (ci == 47u ? 4u : 3u) : 2u) : 5u));
            if (first + t >= n) cls = 6u;
            constexpr uint32_t NW = STATE_THREADS / 32;                      // 7 classes x 4 warps = 28 counters: one warp-wide scan
            static_assert(7u * NW <= 32u, "class counters must fit one warp");
            uint32_t mine = 0;
#pragma unroll
            for (uint32_t c = 0; c < 7u; ++c) {
                const uint32_t bl = __ballot_sync(0xFFFFFFFFu, cls == c);
                if (lane == 0) cls_count[c * NW + warp] = (uint32_t)__popc(bl);
                if (cls == c) mine = (uint32_t)__popc(bl & ((1u << lane) - 1u));
            }
            __syncthreads();
            // exclusive prefix over (class, warp) in class-major order, computed by every warp for itself (no second barrier)
            const uint32_t v = lane < 7u * NW ? cls_count[lane] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (uint32_t d = 1; d < 32u; d <<= 1) { const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += up; }
            const uint32_t before = __shfl_sync(0xFFFFFFFFu, incl - v, cls * NW + warp);
            perm[before + mine] = (uint8_t)t;
            __syncthreads();
            row = perm[t];
        } else __syncthreads();                                          // act_s is read across threads either way
        const uint64_t i = first + row;
        if (i < n) {
            const uint32_t a = act_s[b][row];
            alignas(16) dk_state s;
            StateStage<STATE_THREADS>::get_row(stage, row, s);
            uint32_t err = ENGINE == DK_FDO ? fdo_state_apply_az<DK_APPLY_IDX>(s, a, (flags & DK_APPLY_SKIP_SINGLE) != 0) : doko_state_apply(s, a);
            if (!err) StateStage<STATE_THREADS>::put_row(stage, row, s);     // an illegal action leaves the record as it was
            if (err_out) err_out[i] = (uint8_t)err;
        }
        TmaTile::publish();
        __syncthreads();                                                 // (also: perm / cls_count are free for the next tile)
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tmap), "r"(0), "r"((int)first), "r"(TmaTile::saddr(stage)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the block's shared memory must outlive the engine's reads
}
// dk_new_games with the records leaving as two TMA tile stores per block (the per-thread form writes every record with eight 16-byte
// stores at a 128-byte stride: 256 sector accesses per warp).
__global__ void __launch_bounds__(PLAYOUT_THREADS)
new_games_tma_kernel(const __grid_constant__ CUtensorMap tmap, RngParams rp, uint64_t n) {
    __shared__ __align__(1024) uint4 stage[PLAYOUT_THREADS * 8];       // PLAYOUT_THREADS / STATE_THREADS tiles of 16 KB
    uint32_t* smem = reinterpret_cast<uint32_t*>(stage);               // the shuffle scratch (12 words per thread) lives in the tiles until the deal is done
    const uint64_t first = (uint64_t)blockIdx.x * PLAYOUT_THREADS, i = first + threadIdx.x;
    const uint64_t gi = i < n ? i : n - 1;
    SharedDeck deck;
    deck.base = smem + threadIdx.x;
    RngKey key = make_key(rp, gi, 0, false);
    FdoLive dummy;
    uint32_t ah[4], dup, start;
    fdo_deal(dummy, key, deck, ah, dup, start);
    __syncthreads();                                                   // every deck is dealt: the scratch becomes the record tiles
    {
        uint64_t hands[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) hands[p] = (uint64_t)ah[p] | ((uint64_t)(ah[p] & dup) << 24);   // copy B = doubled cards
        alignas(16) dk_state s;
        st_new_game(s, hands, start);
        StateStage<PLAYOUT_THREADS>::put(stage, s);      // tile threadIdx.x / 128, row threadIdx.x % 128, chunk j at j ^ (row & 7)
    }
    TmaTile::publish();
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (uint32_t k = 0; k < (uint32_t)(PLAYOUT_THREADS / STATE_THREADS); ++k) {
            const uint64_t f = first + k * STATE_THREADS;
            if (f < n)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(&tmap), "r"(0), "r"((int)f), "r"(TmaTile::saddr(stage + k * STATE_THREADS * 8)) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}
__global__ void __launch_bounds__(STATE_THREADS)
from_deals_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint64_t n, const uint64_t* __restrict__ hands, const uint8_t* __restrict__ start) {
    __shared__ __align__(1024) uint4 stage[STATE_THREADS * 8];
    const uint64_t first = (uint64_t)blockIdx.x * STATE_THREADS, i = first + threadIdx.x;
    if (i < n) {
        const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(hands + 4 * i);
        const ulonglong2 h01 = __ldg(hp), h23 = __ldg(hp + 1);
        uint64_t h[4] = {h01.x, h01.y, h23.x, h23.y};
        alignas(16) dk_state s;
        st_new_game(s, h, start[i] & 3u);
        StateStage<STATE_THREADS>::put(stage, s);
    }
    TmaTile::publish();
    __syncthreads();
    if (threadIdx.x == 0) TmaTile::store(&tmap, stage, first);
}
template <int ENGINE>
__global__ void __launch_bounds__(STATE_THREADS)
legal_mask_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint64_t n, uint64_t* __restrict__ mask_out, uint64_t drop_mask, uint64_t drop_count,
                      uint8_t* __restrict__ count_out) {
    __shared__ __align__(1024) uint4 stage[STATE_THREADS * 8];
    __shared__ __align__(8) unsigned long long bar;
    const uint64_t first = (uint64_t)blockIdx.x * STATE_THREADS, i = first + threadIdx.x;
    if (threadIdx.x == 0) TmaTile::init(&bar);
    __syncthreads();
    if (threadIdx.x == 0) TmaTile::load(&tmap, stage, &bar, first, (uint32_t)min((uint64_t)STATE_THREADS, n) * 128u);
    TmaTile::wait(&bar, 0u);
    if (i >= n) return;
    alignas(16) dk_state s;
    StateStage<STATE_THREADS>::get(stage, s);
    store_legal(ENGINE == DK_FDO ? fdo_state_legal_mask(s) : doko_state_legal_mask(s), i, mask_out, drop_mask, drop_count, count_out);
}
// dk_state_id: AzEnvState::id() (rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:156-167) = FxHasher64 over the state and the last
// action.  The hash function is fxhash 0.2.1's (hash = (hash.rotate_left(5) ^ word) * 0x517cc1b727220a95 per 64-bit word), fed with the
// sixteen little-endian 64-bit words of the record followed by the last action as one more word (DK_ACTION_NONE = None); the reference
// feeds its own in-memory layout of FdoState + FdoObservation, so the VALUES differ by construction — equal states give equal ids and
// the record is canonical (one byte pattern per state), which is the property an id() caller can rely on.  (No caller exists in the
// reference; the author marks the function as wrong.)
__device__ __forceinline__ uint64_t fx_add(uint64_t h, uint64_t w) { return (((h << 5) | (h >> 59)) ^ w) * 0x517cc1b727220a95ull; }
__global__ void __launch_bounds__(STATE_THREADS)
state_id_kernel(uint64_t n, const dk_state* __restrict__ states, const uint8_t* __restrict__ last_action, uint64_t* __restrict__ id_out) {
    __shared__ uint4 stage[STATE_THREADS * 8];
    const uint64_t first = (uint64_t)blockIdx.x * STATE_THREADS, i = first + threadIdx.x;
    StateStage<STATE_THREADS>::load(states, first, n, stage);
    if (i >= n) return;
    uint64_t h = 0;
    const uint32_t t = threadIdx.x;
#pragma unroll
    for (uint32_t j = 0; j < 8u; ++j) {
        const uint4 q = stage[t * 8u + (j ^ (t & 7u))];
        h = fx_add(h, (uint64_t)q.x | ((uint64_t)q.y << 32));
        h = fx_add(h, (uint64_t)q.z | ((uint64_t)q.w << 32));
    }
    h = fx_add(h, last_action ? (uint64_t)last_action[i] : (uint64_t)ACTION_NONE);
    id_out[i] = h;
}
// dk_random_action: FdoAllowedActions::random (rs-game-utils/src/bit_flag.rs:86-94) over the legal set of the seat to move, with or
// without the announcement calls — the draw of the random policies (FdoState::random_action_for_current_player[_no_announcement],
// state.rs:378-431) and of DefaultImpiPolicy's fallback (compare_impi.rs:357-368) WITHOUT playing it.  Same stream position as the
// lock-step env step (SITE_STEP word 0 of the unit): dk_step_random_encode with the same rng plays exactly this action.
template <int ENGINE>
__global__ void __launch_bounds__(STATE_THREADS)
random_action_kernel(RngParams rp, uint64_t n, const dk_state* __restrict__ states, uint32_t flags, uint8_t* __restrict__ action_out) {
    __shared__ uint4 stage[STATE_THREADS * 8];
    const uint64_t first = (uint64_t)blockIdx.x * STATE_THREADS, i = first + threadIdx.x;
    StateStage<STATE_THREADS>::load(states, first, n, stage);
    if (i >= n) return;
    alignas(16) dk_state s;
    StateStage<STATE_THREADS>::get(stage, s);
    uint64_t legal = ENGINE == DK_FDO ? fdo_state_legal_mask(s) : doko_state_legal_mask(s);
    if (!(flags & DK_PLAYOUT_WITH_ANNOUNCEMENTS)) legal &= ~(0x1Full << 33);
    uint32_t a = ACTION_NONE;
    if (legal) {
        const U4 blk = rng_block(make_key(rp, i, 0, false), SITE_STEP, 0);
        a = pick_msb_rank64(legal, mulhi(blk.x, popcll(legal)));
    }
    action_out[i] = (uint8_t)a;
}
__global__ void __launch_bounds__(STATE_THREADS)
terminal_kernel(uint64_t n, const dk_state* __restrict__ states, uint8_t* __restrict__ done_out, int4* __restrict__ points_out) {
    uint64_t i = (uint64_t)blockIdx.x * STATE_THREADS + threadIdx.x;
    if (i >= n) return;
    uint4 last = __ldg(reinterpret_cast<const uint4*>(states + i) + 7);       // bytes 112..127: eyes, counters, points, meta
    bool done = (last.w & 3u) == DK_PHASE_FINISHED;
    if (done_out) done_out[i] = done ? 1 : 0;
    if (points_out) {
        uint32_t p = last.z;
        points_out[i] = done ? make_int4((int8_t)(p & 255u), (int8_t)((p >> 8) & 255u), (int8_t)((p >> 16) & 255u), (int8_t)(p >> 24)) : make_int4(0, 0, 0, 0);
    }
}

// ---- observation encode ----------------------------------------------------------------------------------------------------
// Token values are staged as bytes in shared memory (one padded row per game, odd word pitch → conflict-free when the 32
// lanes of a warp write the same column), then every warp widens whole rows to i64 with 256-byte-contiguous stores.
constexpr int ENC_THREADS = 128;
constexpr int ENC_ROW = 116;   // bytes; 29 words (odd) — rs-doko layouts (110 / 114 values)
struct SmemRowOut {
    uint8_t* row;
    __device__ __forceinline__ void operator()(uint32_t i, uint32_t v) const { row[i] = (uint8_t)v; }
};
__device__ __forceinline__ void write_rows(const uint8_t* __restrict__ tok, int len, uint64_t first, uint64_t n, int64_t* __restrict__ out, size_t row_stride) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < ENC_THREADS; r += ENC_THREADS / 32) {
        uint64_t g = first + r;
        if (g >= n) break;
        const uint8_t* src = tok + r * ENC_ROW;
        int64_t* dst = out + g * row_stride;
        for (int i = lane; i < len; i += 32) dst[i] = (int64_t)src[i];
    }
}
// Dense rows (row_stride == LEN, 32-byte aligned output): two consecutive rows are a whole number of sectors (2 x 880 B = 55,
// 2 x 912 B = 57 sectors) and a block's first row starts on a sector boundary, so a warp writes its rows in PAIRS as one
// sector-aligned stream of 2 x LEN i64 in aligned 256-byte windows (same finding as for the 311-token rows below: windows that start
// mid-sector, which every second row does in the row-by-row form, cost a fifth of the store bandwidth).
template <int LEN>
__device__ __forceinline__ void write_rows_dense(const uint8_t* __restrict__ tok, uint64_t first, uint64_t n, int64_t* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int q = warp; q < ENC_THREADS / 2; q += ENC_THREADS / 32) {
        const uint64_t g = first + 2u * q;
        if (g >= n) break;
        const int valid = n - g >= 2u ? 2 * LEN : LEN;
        const uint8_t* src = tok + 2 * q * ENC_ROW;
        int64_t* dst = out + g * LEN;
#pragma unroll
        for (int p0 = 0; p0 < 2 * LEN; p0 += 32) {
            const int p = p0 + lane;
            if (p < valid) dst[p] = (int64_t)src[p >= LEN ? p + (ENC_ROW - LEN) : p];
        }
    }
}
// encode_state_pi staging: one packed word per slot (token 6 b | position 6 b | player 3 b | sub-position 4 b | team 2 b) plus the phase;
// 63 words per game (odd pitch → conflict-free), then each warp expands whole rows channel by channel: every store instruction writes
// 32 (resp. 30) consecutive i64 = 256 contiguous bytes.
constexpr int PI_ROW = 63;
struct SmemSlotOut {
    uint32_t* row;
    __device__ __forceinline__ void slot(uint32_t n, uint32_t tok, uint32_t pos, uint32_t ply, uint32_t sub, uint32_t team) const {
        row[n] = tok | (pos << 6) | (ply << 12) | (sub << 15) | (team << 19);
    }
    __device__ __forceinline__ void phase(uint32_t v) const { row[62] = v; }
};
// one staged row → one i64 row, channel by channel (the calling warp's 32 lanes)
__device__ __forceinline__ void write_row_pi(const uint32_t* __restrict__ src, long long* __restrict__ dst, uint32_t lane) {
    uint32_t w0 = src[lane], w1 = lane < 30 ? src[32 + lane] : 0u;
#define DK_PI_CH(CH, SH, MASK)                                                    \
    dst[(CH) * 62 + lane] = (long long)((w0 >> (SH)) & (MASK));                       \
    if (lane < 30) dst[(CH) * 62 + 32 + lane] = (long long)((w1 >> (SH)) & (MASK));
    DK_PI_CH(0, 0, 63u)
    DK_PI_CH(1, 6, 63u)
    DK_PI_CH(2, 12, 7u)
    DK_PI_CH(3, 15, 15u)
    DK_PI_CH(4, 19, 3u)
#undef DK_PI_CH
    if (lane == 0) dst[310] = (long long)src[62];
}
__device__ __forceinline__ void write_rows_pi(const uint32_t* __restrict__ tok, uint64_t first, uint64_t n, int64_t* __restrict__ out, size_t row_stride) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < ENC_THREADS; r += ENC_THREADS / 32) {
        uint64_t g = first + r;
        if (g >= n) break;
        write_row_pi(tok + r * PI_ROW, reinterpret_cast<long long*>(out) + g * row_stride, (uint32_t)lane);
    }
}
// Dense rows (row_stride == 311): four consecutive rows are exactly 311 sectors (4 x 2488 B = 311 x 32 B), so a warp writes its rows
// in groups of four as ONE sector-aligned stream of 1244 i64: window K (K = 0..38) is the 32 elements [32K, 32K + 32), one aligned
// 256-byte store instruction.  Measured with stores alone (profiles/experiments/store_patterns.cu): this pattern sustains 7.2 TB/s,
// the row-by-row pattern above (256-byte stores that start at 8-byte-aligned row offsets, partial sectors at both ends) 5.7 TB/s.
// Which (row, channel, slot) an element of a window belongs to changes at most twice inside a window, at lane numbers known at
// compile time, so the decode is a couple of selects between immediates, not a division.
template <int K, typename T = long long>
__device__ __forceinline__ void pi_dense_window(const uint32_t* __restrict__ tok4, T* __restrict__ dst4, uint32_t lane, uint32_t rows_valid) {
    constexpr int P0 = 32 * K, R0 = P0 / 311, C0 = P0 % 311;             // first element of the window: row R0 of the group, column C0
    constexpr int ROWCUT = 311 - C0;                                     // lanes >= ROWCUT are in row R0 + 1 (columns from 0)
    constexpr int CH0 = C0 / 62;                                         // channel of the first element (5 = the phase value, column 310)
    constexpr int CHCUT = 62 * (CH0 + 1) - C0;                           // lanes >= CHCUT (and < ROWCUT) are in channel CH0 + 1
    constexpr int SH[6] = {0, 6, 12, 15, 19, 0};
    constexpr uint32_t MK[6] = {63u, 63u, 7u, 15u, 3u, 0xFFFFFFFFu};
    // word index inside the 4-row staging area = BASE + lane
    constexpr int BASE_A = R0 * PI_ROW + (CH0 < 5 ? C0 - 62 * CH0 : 62);                       // segment A: row R0, channel CH0
    constexpr int CH1 = CH0 + 1 < 5 ? CH0 + 1 : 5;
    constexpr int BASE_B = R0 * PI_ROW + (CH0 + 1 < 5 ? C0 - 62 * (CH0 + 1) : 62 - (310 - C0));   // segment B: row R0, channel CH0 + 1
    constexpr int BASE_C = (R0 + 1) * PI_ROW - ROWCUT;                                          // segment C: row R0 + 1, channel 0
    constexpr bool HAS_B = CH0 < 5 && CHCUT < 32 && CHCUT < ROWCUT;
    constexpr bool HAS_C = ROWCUT < 32 && R0 + 1 < 4;
    if (P0 + (int)lane >= 1244) return;
    int base = BASE_A, sh = SH[CH0 < 5 ? CH0 : 5];
    uint32_t mk = MK[CH0 < 5 ? CH0 : 5];
    uint32_t row = R0;
    if (HAS_B && (int)lane >= CHCUT) { base = BASE_B; sh = SH[CH1]; mk = MK[CH1]; }
    if (HAS_C && (int)lane >= ROWCUT) { base = BASE_C; sh = 0; mk = 63u; row = R0 + 1; }
    if (row >= rows_valid) return;
    const uint32_t w = tok4[base + (int)lane];
    dst4[P0 + lane] = (T)((w >> sh) & mk);
}
template <typename T, int... KS>
__device__ __forceinline__ void pi_dense_windows(const uint32_t* __restrict__ tok4, T* __restrict__ dst4, uint32_t lane, uint32_t rows_valid,
                                                 std::integer_sequence<int, KS...>) {
    (pi_dense_window<KS, T>(tok4, dst4, lane, rows_valid), ...);
}
__device__ __forceinline__ void write_rows_pi_dense(const uint32_t* __restrict__ tok, uint64_t first, uint64_t n, int64_t* __restrict__ out) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t q = warp; q < ENC_THREADS / 4; q += ENC_THREADS / 32) {
        const uint64_t g = first + 4u * q;
        if (g >= n) break;
        const uint32_t rows_valid = (uint32_t)min((uint64_t)4, n - g);
        pi_dense_windows(tok + 4u * q * PI_ROW, reinterpret_cast<long long*>(out) + g * 311u, lane, rows_valid, std::make_integer_sequence<int, 39>{});
    }
}
// `count` staged rows (smem rows 0..count-1) → the consecutive dense rows row0 .. row0+count-1 of `out` (row0 arbitrary): rows up to the
// next multiple of four and the tail go row by row, everything between as sector-aligned 4-row streams.  `out` must be 32-byte aligned.
__device__ __forceinline__ void write_rows_pi_dense_at(const uint32_t* __restrict__ tok, uint64_t row0, uint32_t count, long long* __restrict__ out) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const uint32_t head = min(count, (uint32_t)((4u - (uint32_t)(row0 & 3ull)) & 3u));
    const uint32_t groups = (count - head) / 4u, tail0 = head + 4u * groups;
    for (uint32_t q = warp; q < groups; q += n_warps)
        pi_dense_windows(tok + (head + 4u * q) * PI_ROW, out + (row0 + head + 4u * q) * 311u, lane, 4u, std::make_integer_sequence<int, 39>{});
    for (uint32_t r = warp; r < head + (count - tail0); r += n_warps) {
        const uint32_t sr = r < head ? r : tail0 + (r - head);
        write_row_pi(tok + sr * PI_ROW, out + (row0 + sr) * 311u, lane);
    }
}
__global__ void __launch_bounds__(ENC_THREADS)
encode_pi_kernel(uint64_t n, const dk_state* __restrict__ states, int64_t* __restrict__ out, size_t row_stride) {
    __shared__ uint32_t tok[ENC_THREADS * PI_ROW];
    uint64_t first = (uint64_t)blockIdx.x * ENC_THREADS;
    uint64_t i = first + threadIdx.x;
    if (i < n) {
        alignas(16) dk_state s;
        load_state(states + i, s);
        SmemSlotOut o{tok + threadIdx.x * PI_ROW};
        fdo_encode_pi(s, o);
    }
    __syncthreads();
    if (row_stride == 311u && !(reinterpret_cast<uintptr_t>(out) & 31u)) write_rows_pi_dense(tok, first, n, out);   // uniform branch
    else write_rows_pi(tok, first, n, out, row_stride);
}

// encode_state_ipi for a batch: per game the guessed hands (u64[4] by absolute seat), guessed reservations (u8[4]) and the seat to guess for.
__global__ void __launch_bounds__(ENC_THREADS)
encode_ipi_kernel(uint64_t n, const dk_state* __restrict__ states, const uint64_t* __restrict__ assumed_hands, const uint8_t* __restrict__ assumed_res,
                  const uint8_t* __restrict__ next_player, int64_t* __restrict__ out, size_t row_stride, uint8_t* __restrict__ err_out) {
    __shared__ uint32_t tok[ENC_THREADS * PI_ROW];
    uint64_t first = (uint64_t)blockIdx.x * ENC_THREADS;
    uint64_t i = first + threadIdx.x;
    if (i < n) {
        alignas(16) dk_state s;
        load_state(states + i, s);
        const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(assumed_hands + 4 * i);
        ulonglong2 h01 = __ldg(hp), h23 = __ldg(hp + 1);
        uint64_t ah[4] = {h01.x, h01.y, h23.x, h23.y};
        uint32_t packed = __ldg(reinterpret_cast<const uint32_t*>(assumed_res) + i);
        uint8_t ar[4] = {(uint8_t)packed, (uint8_t)(packed >> 8), (uint8_t)(packed >> 16), (uint8_t)(packed >> 24)};
        SmemSlotOut o{tok + threadIdx.x * PI_ROW};
        uint32_t err = fdo_encode_ipi(s, ah, ar, next_player[i] & 3u, o);
        if (err_out) err_out[i] = (uint8_t)err;
    }
    __syncthreads();
    if (row_stride == 311u && !(reinterpret_cast<uintptr_t>(out) & 31u)) write_rows_pi_dense(tok, first, n, out);   // uniform branch
    else write_rows_pi(tok, first, n, out, row_stride);
}

template <int LAYOUT>
__global__ void __launch_bounds__(ENC_THREADS)
encode_kernel(uint64_t n, const dk_state* __restrict__ states, int64_t* __restrict__ out, size_t row_stride) {
    __shared__ __align__(16) uint8_t tok[ENC_THREADS * ENC_ROW];
    uint64_t first = (uint64_t)blockIdx.x * ENC_THREADS;
    uint64_t i = first + threadIdx.x;
    if (i < n) {
        alignas(16) dk_state s;
        load_state(states + i, s);
        SmemRowOut o{tok + threadIdx.x * ENC_ROW};
        doko_encode(s, LAYOUT == DK_LAYOUT_DO114, o);
    }
    __syncthreads();
    constexpr int LEN = LAYOUT == DK_LAYOUT_DO114 ? 114 : 110;
    if (row_stride == (size_t)LEN && !(reinterpret_cast<uintptr_t>(out) & 31u)) write_rows_dense<LEN>(tok, first, n, out);   // uniform branch
    else write_rows(tok, LEN, first, n, out, row_stride);
}

// K5: one lock-step self-play env step + observation: legal mask → one draw (SITE_STEP word 0, unit = game id,
// epoch = caller's step counter) → play_action [→ skip forced moves] → encode_state_pi of the new state.
// Algorithmic HBM bytes per game: 128 read + 128 written + 2488 written.
__global__ void __launch_bounds__(ENC_THREADS)
fdo_step_encode_kernel(RngParams rp, uint64_t n, dk_state* __restrict__ states, uint32_t flags, int64_t* __restrict__ obs, size_t row_stride,
                       uint8_t* __restrict__ action_out) {
    __shared__ uint32_t tok[ENC_THREADS * PI_ROW];
    uint64_t first = (uint64_t)blockIdx.x * ENC_THREADS;
    uint64_t i = first + threadIdx.x;
    if (i < n) {
        alignas(16) dk_state s;
        load_state(states + i, s);
        uint64_t legal = fdo_state_legal_mask(s);
        if (!(flags & DK_PLAYOUT_WITH_ANNOUNCEMENTS)) legal &= ~(0x1Full << 33);
        uint32_t a = 0xFF;
        if (legal) {
            RngKey key = make_key(rp, i, 0, false);
            U4 blk = rng_block(key, SITE_STEP, 0);
            a = pick_msb_rank64(legal, mulhi(blk.x, popcll(legal)));
            fdo_state_apply_az(s, a, (flags & DK_STEP_SKIP_SINGLE) != 0);
            store_state(states + i, s);
        }
        if (action_out) action_out[i] = (uint8_t)a;
        if (obs) { SmemSlotOut o{tok + threadIdx.x * PI_ROW}; fdo_encode_pi(s, o); }
    }
    __syncthreads();
    if (obs) {
        if (row_stride == 311u && !(reinterpret_cast<uintptr_t>(obs) & 31u)) write_rows_pi_dense(tok, first, n, obs);   // uniform branch
        else write_rows_pi(tok, first, n, obs, row_stride);
    }
}

// K5 with the record tile moved by the TMA engine: the first 16 KB of the token staging area double as the tile — the records are
// back in HBM (the engine has read the tile) before the tokens overwrite it.  1.805 -> 1.793 ms per 2^22 games against the
// per-thread record loads / stores above (profiles/r01_k5_tma.txt; DOKO_CUDA_NO_TMA=1 selects those).
__global__ void __launch_bounds__(ENC_THREADS)
fdo_step_encode_tma_kernel(const __grid_constant__ CUtensorMap tmap, RngParams rp, uint64_t n, uint32_t flags, int64_t* __restrict__ obs, size_t row_stride,
                           uint8_t* __restrict__ action_out) {
    __shared__ __align__(1024) uint32_t tok[ENC_THREADS * PI_ROW];
    __shared__ __align__(8) unsigned long long bar;
    uint4* stage = reinterpret_cast<uint4*>(tok);
    const uint64_t first = (uint64_t)blockIdx.x * ENC_THREADS, i = first + threadIdx.x;
    if (threadIdx.x == 0) TmaTile::init(&bar);
    __syncthreads();
    if (threadIdx.x == 0) TmaTile::load(&tmap, stage, &bar, first, (uint32_t)min((uint64_t)ENC_THREADS, n) * 128u);
    TmaTile::wait(&bar, 0u);
    alignas(16) dk_state s;
    if (i < n) {
        StateStage<ENC_THREADS>::get(stage, s);
        uint64_t legal = fdo_state_legal_mask(s);
        if (!(flags & DK_PLAYOUT_WITH_ANNOUNCEMENTS)) legal &= ~(0x1Full << 33);
        uint32_t a = 0xFF;
        if (legal) {
            RngKey key = make_key(rp, i, 0, false);
            U4 blk = rng_block(key, SITE_STEP, 0);
            a = pick_msb_rank64(legal, mulhi(blk.x, popcll(legal)));
            fdo_state_apply_az(s, a, (flags & DK_STEP_SKIP_SINGLE) != 0);
            StateStage<ENC_THREADS>::put(stage, s);
        }
        if (action_out) action_out[i] = (uint8_t)a;
    }
    TmaTile::publish();
    __syncthreads();
    if (threadIdx.x == 0) TmaTile::store(&tmap, stage, first);
    __syncthreads();                                              // the engine has read the tile: the staging area is free for the tokens
    if (obs) {
        if (i < n) { SmemSlotOut o{tok + threadIdx.x * PI_ROW}; fdo_encode_pi(s, o); }
        __syncthreads();
        if (row_stride == 311u && !(reinterpret_cast<uintptr_t>(obs) & 31u)) write_rows_pi_dense(tok, first, n, obs);   // uniform branch
        else write_rows_pi(tok, first, n, obs, row_stride);
    }
}

// ---- narrow observation rows (int32 / uint8; dk_encode_narrow, dk_step_random_encode_narrow) ------------------------------------------
// The reference's rows are i64 (`Vec<i64>`), and its Python side narrows them to int32 at once (rs-doko-py-bridge: az_doko.py:369); a
// caller that keeps the tokens on the device can ask for that row directly: 1244 B (int32) or 311 B (uint8) per observation instead
// of 2488 B.  Every token value is < 256.  Staging is the i64 kernels' (one packed word per slot, 63 words per game, conflict-free);
// a warp then expands whole rows channel by channel with lane = slot: every store instruction writes 32 (30) consecutive elements of
// one channel — 128 contiguous bytes as int32, 32 as uint8 — ten store instructions per row like the i64 form, a half or an eighth of
// its bytes.  (Two byte-image forms were measured first, profiles/r02_narrow_rows.json: one staged byte per element costs five
// shared-memory stores per slot — 4-way bank conflicts at the natural pitch of 311 bytes, none at 316, and either way ~1.0 ms per 2^22
// games of staging alone, more than the int32 rows' HBM time.)
template <typename T>
__device__ __forceinline__ void write_row_pi_narrow(const uint32_t* __restrict__ src, T* __restrict__ dst, uint32_t lane) {
    const uint32_t w0 = src[lane], w1 = lane < 30 ? src[32 + lane] : 0u;
#define DK_PI_CH(CH, SH, MASK)                                                   \
    dst[(CH) * 62 + lane] = (T)((w0 >> (SH)) & (MASK));                              \
    if (lane < 30) dst[(CH) * 62 + 32 + lane] = (T)((w1 >> (SH)) & (MASK));
    DK_PI_CH(0, 0, 63u)
    DK_PI_CH(1, 6, 63u)
    DK_PI_CH(2, 12, 7u)
    DK_PI_CH(3, 15, 15u)
    DK_PI_CH(4, 19, 3u)
#undef DK_PI_CH
    if (lane == 0) dst[310] = (T)src[62];
}
// int32: rows in groups of FOUR as one stream of 1244 elements through the i64 form's compile-time windows (pi_dense_window): a warp's
// 39 store instructions of 128 bytes continue each other, so only the two ends of a group (4976 bytes) are partial sectors shared with
// another warp instead of the two ends of every row: the bare encoder 1.03 -> 0.915 ms per 2^22 rows = 0.96 of the HBM copy peak.
// uint8: row by row (32-byte stores; the grouped form measures 8 % slower: 0.69 vs 0.75 ms — this size is bound by the encode itself).
template <typename T>
__device__ __forceinline__ void write_rows_pi_narrow(const uint32_t* __restrict__ tok, uint64_t first, uint64_t n, T* __restrict__ out) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (sizeof(T) == 4) {
        for (uint32_t q = warp; q < ENC_THREADS / 4; q += ENC_THREADS / 32) {
            const uint64_t g = first + 4u * q;
            if (g >= n) break;
            const uint32_t rows_valid = (uint32_t)min((uint64_t)4, n - g);
            pi_dense_windows<T>(tok + 4u * q * PI_ROW, out + g * 311u, lane, rows_valid, std::make_integer_sequence<int, 39>{});
        }
    } else {
        for (uint32_t r = warp; r < ENC_THREADS; r += ENC_THREADS / 32) {
            const uint64_t g = first + r;
            if (g >= n) break;
            write_row_pi_narrow<T>(tok + r * PI_ROW, out + g * 311u, lane);
        }
    }
}
template <typename T>
__global__ void __launch_bounds__(ENC_THREADS)
encode_pi_narrow_kernel(uint64_t n, const dk_state* __restrict__ states, T* __restrict__ out) {
    __shared__ uint32_t tok[ENC_THREADS * PI_ROW];
    const uint64_t first = (uint64_t)blockIdx.x * ENC_THREADS, i = first + threadIdx.x;
    if (i < n) {
        alignas(16) dk_state s;
        load_state(states + i, s);
        SmemSlotOut o{tok + threadIdx.x * PI_ROW};
        fdo_encode_pi(s, o);
    }
    __syncthreads();
    write_rows_pi_narrow<T>(tok, first, n, out);
}
template <int LAYOUT, typename T>
__global__ void __launch_bounds__(ENC_THREADS)
encode_narrow_kernel(uint64_t n, const dk_state* __restrict__ states, T* __restrict__ out) {
    constexpr int LEN = LAYOUT == DK_LAYOUT_DO114 ? 114 : 110;
    __shared__ __align__(16) uint8_t tok[ENC_THREADS * ENC_ROW];
    const uint64_t first = (uint64_t)blockIdx.x * ENC_THREADS, i = first + threadIdx.x;
    if (i < n) {
        alignas(16) dk_state s;
        load_state(states + i, s);
        SmemRowOut o{tok + threadIdx.x * ENC_ROW};
        doko_encode(s, LAYOUT == DK_LAYOUT_DO114, o);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < ENC_THREADS; r += ENC_THREADS / 32) {
        const uint64_t g = first + r;
        if (g >= n) break;
        const uint8_t* src = tok + r * ENC_ROW;
        T* dst = out + g * LEN;
        for (int k = lane; k < LEN; k += 32) dst[k] = (T)src[k];
    }
}
// K5 with narrow rows: the same env step (legal mask -> SITE_STEP draw -> play_action [-> forced moves]) as fdo_step_encode_kernel.
// Algorithmic HBM bytes per game: 128 read + 128 written + 311 x sizeof(T) written (1500 B int32, 567 B uint8; i64: 2744 B).
template <typename T>
__global__ void __launch_bounds__(ENC_THREADS)
fdo_step_encode_narrow_kernel(RngParams rp, uint64_t n, dk_state* __restrict__ states, uint32_t flags, T* __restrict__ obs, uint8_t* __restrict__ action_out) {
    __shared__ uint32_t tok[ENC_THREADS * PI_ROW];
    const uint64_t first = (uint64_t)blockIdx.x * ENC_THREADS, i = first + threadIdx.x;
    if (i < n) {
        alignas(16) dk_state s;
        load_state(states + i, s);
        uint64_t legal = fdo_state_legal_mask(s);
        if (!(flags & DK_PLAYOUT_WITH_ANNOUNCEMENTS)) legal &= ~(0x1Full << 33);
        uint32_t a = 0xFF;
        if (legal) {
            RngKey key = make_key(rp, i, 0, false);
            U4 blk = rng_block(key, SITE_STEP, 0);
            a = pick_msb_rank64(legal, mulhi(blk.x, popcll(legal)));
            fdo_state_apply_az(s, a, (flags & DK_STEP_SKIP_SINGLE) != 0);
            store_state(states + i, s);
        }
        if (action_out) action_out[i] = (uint8_t)a;
        SmemSlotOut o{tok + threadIdx.x * PI_ROW};
        fdo_encode_pi(s, o);
    }
    __syncthreads();
    write_rows_pi_narrow<T>(tok, first, n, obs);
}

// The same with the record tile moved by the TMA engine (see fdo_step_encode_tma_kernel: the first 16 KB of the token staging area
// double as the tile).  With half or an eighth of the row bytes the per-thread record loads / stores of the kernel above are no longer
// hidden under the row stores: int32 step + encode 1.12 ms per 2^22 games against 0.92 ms for the bare encoder.
template <typename T>
__global__ void __launch_bounds__(ENC_THREADS)
fdo_step_encode_narrow_tma_kernel(const __grid_constant__ CUtensorMap tmap, RngParams rp, uint64_t n, uint32_t flags, T* __restrict__ obs,
                                  uint8_t* __restrict__ action_out) {
    __shared__ __align__(1024) uint32_t tok[ENC_THREADS * PI_ROW];
    __shared__ __align__(8) unsigned long long bar;
    uint4* stage = reinterpret_cast<uint4*>(tok);
    const uint64_t first = (uint64_t)blockIdx.x * ENC_THREADS, i = first + threadIdx.x;
    if (threadIdx.x == 0) TmaTile::init(&bar);
    __syncthreads();
    if (threadIdx.x == 0) TmaTile::load(&tmap, stage, &bar, first, (uint32_t)min((uint64_t)ENC_THREADS, n) * 128u);
    TmaTile::wait(&bar, 0u);
    alignas(16) dk_state s;
    if (i < n) {
        StateStage<ENC_THREADS>::get(stage, s);
        uint64_t legal = fdo_state_legal_mask(s);
        if (!(flags & DK_PLAYOUT_WITH_ANNOUNCEMENTS)) legal &= ~(0x1Full << 33);
        uint32_t a = 0xFF;
        if (legal) {
            RngKey key = make_key(rp, i, 0, false);
            U4 blk = rng_block(key, SITE_STEP, 0);
            a = pick_msb_rank64(legal, mulhi(blk.x, popcll(legal)));
            fdo_state_apply_az(s, a, (flags & DK_STEP_SKIP_SINGLE) != 0);
            StateStage<ENC_THREADS>::put(stage, s);
        }
        if (action_out) action_out[i] = (uint8_t)a;
    }
    TmaTile::publish();
    __syncthreads();
    if (threadIdx.x == 0) TmaTile::store(&tmap, stage, first);
    __syncthreads();                                              // the engine has read the tile: the staging area is free for the tokens
    if (i < n) { SmemSlotOut o{tok + threadIdx.x * PI_ROW}; fdo_encode_pi(s, o); }
    __syncthreads();
    write_rows_pi_narrow<T>(tok, first, n, obs);
}

// K2/K4 from stored states: McEnvState::random_rollout (rs-doko-mcts/src/env/envs/env_state_full_doko.rs:198-220) and the
// with-announcement loop.  unit = first_id + (i / per_unit), unit_hi = i % per_unit when per_unit > 1 (leaf rollouts).
// 256-thread blocks, each with the full table set (12-bit rank select included: a quarter of this kernel's instructions were the
// two-level rank select, profiles/r01_playout_state_attribution.json).
constexpr int PLAYOUT_STATE_THREADS = 256;
template <int ENGINE, bool WITH_ANN>
__global__ void __launch_bounds__(PLAYOUT_STATE_THREADS)
playout_state_kernel(RngParams rp, uint64_t n, const dk_state* __restrict__ states, uint32_t per_unit, void* __restrict__ points,
                     void* __restrict__ steps, uint32_t mode, unsigned long long* __restrict__ stats) {
    __shared__ __align__(16) uint32_t lut[FULL_LUT_WORDS];
    __shared__ BlockStats bstats;
    stage_lut(lut, FULL_LUT_WORDS);
    if (stats) block_stats_clear(bstats);
    __syncthreads();
    const uint64_t i_raw = (uint64_t)blockIdx.x * PLAYOUT_STATE_THREADS + threadIdx.x;
    const bool valid = i_raw < n;
    if (!valid && !stats) return;
    const uint64_t i = valid ? i_raw : n - 1;                          // with a summary every thread stays for the block reduction
    uint64_t unit = per_unit > 1u ? i / per_unit : i;
    RngKey key = make_key(rp, unit, per_unit > 1u ? (uint32_t)(i % per_unit) : 0u, per_unit > 1u);
    alignas(16) dk_state s;
    load_state(states + unit, s);
    int32_t p[4];
    uint32_t st = 0;
    if (ENGINE == DK_FDO) {
        FdoLive g; FdoResume rs;
        if (fdo_state_to_live(s, g, rs)) { fdo_play_to_end<WITH_ANN, false, true>(g, key, &rs, lut); fdo_final_points(g, p); st = g.steps; }
        else { p[0] = s.points[0]; p[1] = s.points[1]; p[2] = s.points[2]; p[3] = s.points[3]; }
    } else {
        DokoLive g; DokoResume rs;
        if (doko_state_to_live(s, g, rs)) { doko_play_to_end<false, false, true>(g, key, &rs, nullptr, lut); doko_final_points(g, p); st = g.steps; }
        else { p[0] = s.points[0]; p[1] = s.points[1]; p[2] = s.points[2]; p[3] = s.points[3]; }
    }
    if (valid) store_result(points, steps, i, p, st, mode);
    if (stats) block_stats_add(bstats, stats, p, st, valid);          // uniform branch
}


// K3: determinization.  One block per info-state; the constraint tables are built once (thread 0) and staged in shared
// memory, then each thread draws samples s = tid, tid + blockDim, ...  (card_matching, rs-full-doko/src/matching/card_matching.rs:241-467).
// HBM per sample: 32 B hands + 4 B reservations + 1 B status written; 128 B read per info-state.
constexpr int MATCH_THREADS = 128;
__global__ void __launch_bounds__(MATCH_THREADS)
fdo_determinize_kernel(RngParams rp, uint64_t n_info, uint32_t samples, uint32_t splits, const dk_state* __restrict__ states, uint64_t* __restrict__ hands_out,
                       uint8_t* __restrict__ res_out, uint8_t* __restrict__ status_out) {
    __shared__ MatchPrep prep;
    __shared__ uint32_t rank6[64];                                 // rank-select table of rule 4 (dk_common.cuh rank_lut6_entry)
    // `splits` blocks share one info-state (sample smp belongs to block smp / MATCH_THREADS % splits): a single decision's 4096 samples
    // spread over 32 SMs instead of one, and small batches do not end in a mostly empty last wave.  Sample ids, hence results, do not change.
    const uint64_t i = blockIdx.x / splits;
    const uint32_t part = blockIdx.x - (uint32_t)i * splits;
    if (i >= n_info) return;
    if (threadIdx.x >= 64) rank6[threadIdx.x - 64] = rank_lut6_entry(threadIdx.x - 64);
    if (threadIdx.x == 0) {
        alignas(16) dk_state s;
        load_state(states + i, s);
        fdo_match_prepare(s, prep);
    }
    __syncthreads();
    for (uint32_t smp = part * MATCH_THREADS + threadIdx.x; smp < samples; smp += splits * MATCH_THREADS) {
        RngKey key = make_key(rp, i, rp.first_sub + smp, true);
        uint64_t h[4];
        uint8_t r[4];
        uint32_t st = prep.valid ? fdo_match_sample(prep, key, h, r, rank6) : 2u;
        if (!prep.valid) { h[0] = h[1] = h[2] = h[3] = 0; r[0] = r[1] = r[2] = r[3] = 0xFF; }
        uint64_t o = i * samples + smp;
        if (hands_out) {
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(hands_out + 4 * o);
            dst[0] = make_ulonglong2(h[0], h[1]);
            dst[1] = make_ulonglong2(h[2], h[3]);
        }
        if (res_out) reinterpret_cast<uint32_t*>(res_out)[o] = (uint32_t)r[0] | ((uint32_t)r[1] << 8) | ((uint32_t)r[2] << 16) | ((uint32_t)r[3] << 24);
        if (status_out) status_out[o] = (uint8_t)st;
    }
}

// K3 (rs-doko): sample_assignment (rs-doko-assignment/src/assignment.rs:493-581).  Same block = info-state / thread = sample shape.
// reservations_out repeats the real reservations by absolute seat (the reference only replaces the hands).
__global__ void __launch_bounds__(MATCH_THREADS)
doko_assign_kernel(RngParams rp, uint64_t n_info, uint32_t samples, uint32_t splits, const dk_state* __restrict__ states, uint64_t* __restrict__ hands_out,
                   uint8_t* __restrict__ res_out, uint8_t* __restrict__ status_out) {
    __shared__ AssignPrep prep;
    __shared__ uint32_t res_word;
    __shared__ uint32_t adj3[64];                                  // last level of the multiset rank select (assignment.cuh adj3_entry)
    const uint64_t i = blockIdx.x / splits;                        // `splits` blocks per info-state, see fdo_determinize_kernel
    const uint32_t part = blockIdx.x - (uint32_t)i * splits;
    if (i >= n_info) return;
    if (threadIdx.x >= 64) adj3[threadIdx.x - 64] = adj3_entry(threadIdx.x - 64);
    if (threadIdx.x == 0) {
        alignas(16) dk_state s;
        load_state(states + i, s);
        doko_assign_prepare(s, prep);
        uint32_t r = 0xFFFFFFFFu, start = st_game_start(s);
        for (uint32_t k = 0; k < s.n_reservations; ++k) { uint32_t seat = (start + k) & 3u; r = (r & ~(0xFFu << (8u * seat))) | ((uint32_t)s.reservations[k] << (8u * seat)); }
        res_word = r;
    }
    __syncthreads();
    for (uint32_t smp = part * MATCH_THREADS + threadIdx.x; smp < samples; smp += splits * MATCH_THREADS) {
        RngKey key = make_key(rp, i, rp.first_sub + smp, true);
        uint64_t h[4];
        uint32_t st = doko_assign_sample(prep, key, h, adj3);
        uint64_t o = i * samples + smp;
        if (hands_out) {
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(hands_out + 4 * o);
            dst[0] = make_ulonglong2(h[0], h[1]);
            dst[1] = make_ulonglong2(h[2], h[3]);
        }
        if (res_out) reinterpret_cast<uint32_t*>(res_out)[o] = res_word;
        if (status_out) status_out[o] = (uint8_t)st;
    }
}

// K4: leaf-parallel rollouts, [determinize →] rollout → block reduction.  One block per leaf; rollout r uses the Philox
// unit (leaf id, r).  point_sum[leaf][seat] = exact integer sum of player_points over the rollouts (dead-end samples add 0).
// The bridge from the stored record to the register-resident playout form (trick history trackers, partial trick, announcement
// resume point: fdo_state_to_live) is the same for every rollout of the leaf, so it is built ONCE per block in shared memory.
// Determinized rollouts (DET) run as TWO kernels: K3 (fdo_determinize_kernel) writes the samples of a chunk of leaves to a scratch
// buffer at its own occupancy (54 registers), this kernel reads rollout r's sample (37 bytes, coalesced) and only replaces what
// card_matching changes — the four hands, the doubled-card mask and, in the reservation phase, the hidden reservations of the seats
// that have declared.  History: sampler + record clone + bridge + playout in one loop body, 3.1e9 rollouts/s (72 registers, 400 B
// of local memory per thread); bridge hoisted out of the loop, sampler still inside: 3.2e9 (the sampler's ~3200 instructions per
// sample and its register file decide); two kernels: see DESIGN.md §3.
template <bool DET>
#ifndef DK_LEAF_BLOCKS
#define DK_LEAF_BLOCKS 6
#endif
__global__ void __launch_bounds__(MATCH_THREADS, DK_LEAF_BLOCKS)
fdo_leaf_rollouts_kernel(RngParams rp, uint64_t leaf0, uint64_t n_leaves, uint32_t rollouts, uint32_t splits, const dk_state* __restrict__ states,
                         const uint64_t* __restrict__ smp_hands, const uint8_t* __restrict__ smp_res, const uint8_t* __restrict__ smp_status,
                         long long* __restrict__ point_sum) {
    __shared__ __align__(16) dk_state leaf;
    __shared__ FdoLive live0;
    __shared__ FdoResume resume0;
    __shared__ int live_ok;
    __shared__ int red[4];
    __shared__ __align__(16) uint32_t lut[CARD_LUT_WORDS + SEL12_WORDS];   // card tables + 12-bit rank select; no ANN region: only the points are wanted
    stage_lut(lut, CARD_LUT_WORDS + SEL12_WORDS);
    // `splits` blocks share one leaf when there are fewer leaves than one wave of blocks (rollout r belongs to block
    // r / MATCH_THREADS % splits); their integer sums meet in point_sum by atomics (zeroed by the host), so the result does not change.
    const uint64_t li = blockIdx.x / splits;                       // leaf inside this launch
    const uint32_t part = blockIdx.x - (uint32_t)li * splits;
    if (li >= n_leaves) return;
    const uint64_t i = leaf0 + li;
    if (threadIdx.x < 8) reinterpret_cast<uint4*>(&leaf)[threadIdx.x] = __ldg(reinterpret_cast<const uint4*>(states + i) + threadIdx.x);
    if (threadIdx.x < 4) red[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        FdoLive g; FdoResume rs;
        live_ok = fdo_state_to_live<true>(leaf, g, rs) ? 1 : 0;
        live0 = g; resume0 = rs;
    }
    __syncthreads();
    int acc[4] = {0, 0, 0, 0};
    for (uint32_t r = part * MATCH_THREADS + threadIdx.x; r < rollouts; r += splits * MATCH_THREADS) {
        RngKey key = make_key(rp, i, rp.first_sub + r, true);
        int32_t p[4];
        if (!live_ok) { p[0] = leaf.points[0]; p[1] = leaf.points[1]; p[2] = leaf.points[2]; p[3] = leaf.points[3]; }
        else if (DET) {
            const uint64_t o = li * rollouts + r;                                     // sample of (leaf, rollout) in the scratch buffer
            if (smp_status[o] != 0u) continue;                                        // a dead end adds nothing
            const ulonglong2 h01 = __ldg(reinterpret_cast<const ulonglong2*>(smp_hands + 4 * o)), h23 = __ldg(reinterpret_cast<const ulonglong2*>(smp_hands + 4 * o) + 1);
            const uint32_t res4 = __ldg(reinterpret_cast<const uint32_t*>(smp_res) + o);
            FdoLive g = live0;
            FdoResume rs = resume0;
            const uint64_t sh[4] = {h01.x, h01.y, h23.x, h23.y};
            fdo_live_with_sample(g, rs, leaf, sh, res4);
            fdo_play_to_end<false, false, true, false>(g, key, &rs, lut);
            fdo_final_points(g, p);
        } else {
            FdoLive g = live0;
            fdo_play_to_end<false, false, true, false>(g, key, &resume0, lut);
            fdo_final_points(g, p);
        }
        acc[0] += p[0]; acc[1] += p[1]; acc[2] += p[2]; acc[3] += p[3];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&red[q], v);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        if (splits == 1u) point_sum[4 * i + threadIdx.x] = (long long)red[threadIdx.x];
        else atomicAdd(reinterpret_cast<unsigned long long*>(point_sum) + 4 * i + threadIdx.x, (unsigned long long)(long long)red[threadIdx.x]);
    }
}

// N2: flat Monte-Carlo PIMC evaluator (the per-sample policy slot of DefaultImpiPolicy::execute,
// rs-doko-py-bridge/src/compare_impi/compare_impi.rs:223-262, filled with "every legal action × R rollouts" instead of a UCT search).
// Determinization d of a root is sampled on the stream (root, d) — the dk_determinize stream; rollout r of it plays EVERY legal action
// of the seat to move followed by a _no_announcement rollout on the stream (root, d·R + r) shared by all actions (common random
// numbers), adds points[mover] to value_sum[d][a] and one visit to the best action (first among equals).
// A first version did all of it in one kernel (block = root × a few determinizations): with few rollouts per determinization a block
// spent most of its life in the prologue (one thread building the constraint tables, a handful sampling, the rest waiting) —
// 2.7e9 rollouts/s at 64 x 32 against 8e9 for the bare rollout kernel.  The work is split so that every phase runs at full width:
//   pimc_prepare_kernel   block = root x 64 determinizations, thread = determinization: sample (the dk_determinize stream), legal set,
//                         and for every legal action the state after it in playout form → workspace entry ((root * n_det + d) * 12 + k)
//   pimc_rollout_kernel   thread = (root, determinization, rollout): one _no_announcement rollout per legal action from the workspace
//                         entries (common random numbers), integer sums / winner counts added to the outputs with atomics (order
//                         independent, bit-reproducible)
struct PimcEntry {
    FdoLive g;
    FdoResume rs;
    signed char pts[4];      // final points when the action ends the game (live == 0)
    uint32_t live;
};
constexpr uint32_t PIMC_MAX_LEGAL = 12u;        // <= 12 card types in a hand, 9 reservations, 2 announcement actions
constexpr int PIMC_PREP_THREADS = 64;
__global__ void __launch_bounds__(PIMC_PREP_THREADS)
pimc_prepare_kernel(RngParams rp, uint64_t root0, uint64_t n_roots, uint32_t n_det, const dk_state* __restrict__ states, PimcEntry* __restrict__ ws,
                    uint64_t* __restrict__ masks, uint8_t* __restrict__ status_out) {
    __shared__ MatchPrep prep;
    __shared__ __align__(16) dk_state rootst_s;
    const uint32_t chunks = (n_det + PIMC_PREP_THREADS - 1) / PIMC_PREP_THREADS;
    const uint64_t rl = blockIdx.x / chunks;                                  // root inside this launch
    const uint32_t d = (uint32_t)(blockIdx.x % chunks) * PIMC_PREP_THREADS + threadIdx.x;
    const uint64_t root = root0 + rl;
    if (threadIdx.x < 8) reinterpret_cast<uint4*>(&rootst_s)[threadIdx.x] = __ldg(reinterpret_cast<const uint4*>(states + root) + threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) fdo_match_prepare(rootst_s, prep);
    __syncthreads();
    if (d >= n_det) return;
    const bool finished = !prep.valid;                                        // finished root: no legal action, rows stay zero
    alignas(16) dk_state s = rootst_s;
    uint32_t st = 0;
    if (!finished) {
        RngKey key = make_key(rp, root, rp.first_sub + d, true);
        uint64_t h[4];
        uint8_t res[4];
        st = fdo_match_sample(prep, key, h, res);
        if (st == 0u) fdo_state_with_hands_and_reservations(s, h, res);
    }
    const uint64_t row = rl * n_det + d;
    if (status_out) status_out[root * n_det + d] = (uint8_t)st;
    const uint64_t m = (finished || st != 0u) ? 0ull : fdo_state_legal_mask<true>(s);
    masks[row] = m;
    uint32_t k = 0;
    for (uint64_t mm = m; mm && k < PIMC_MAX_LEGAL; mm &= mm - 1ull, ++k) {
        alignas(16) dk_state t = s;
        fdo_state_apply<true>(t, ffs0ll(mm));
        PimcEntry& e = ws[row * PIMC_MAX_LEGAL + k];
        FdoLive g; FdoResume rs;
        const bool lv = fdo_state_to_live<true>(t, g, rs);
        e.g = g; e.rs = rs; e.live = lv ? 1u : 0u;
        e.pts[0] = t.points[0]; e.pts[1] = t.points[1]; e.pts[2] = t.points[2]; e.pts[3] = t.points[3];
    }
}
constexpr int PIMC_ROLL_THREADS = 128;
__global__ void __launch_bounds__(PIMC_ROLL_THREADS, 6)
pimc_rollout_kernel(RngParams rp, uint64_t root0, uint32_t n_det, uint32_t n_rollouts, uint32_t blocks_per_root, const dk_state* __restrict__ states,
                    const PimcEntry* __restrict__ ws, const uint64_t* __restrict__ masks, uint32_t* __restrict__ visits_out,
                    unsigned long long* __restrict__ value_out) {
    __shared__ __align__(16) uint32_t lut[CARD_LUT_WORDS];
    stage_lut(lut, CARD_LUT_WORDS);
    __syncthreads();
    const uint64_t rl = blockIdx.x / blocks_per_root, root = root0 + rl;
    const uint32_t item = (uint32_t)(blockIdx.x % blocks_per_root) * PIMC_ROLL_THREADS + threadIdx.x;
    const uint32_t n_items = n_det * n_rollouts;
    const bool on = item < n_items;
    const uint32_t d = on ? item / n_rollouts : 0u, r = on ? item - d * n_rollouts : 0u;
    const uint64_t row = rl * n_det + d;
    const uint64_t m = on ? masks[row] : 0ull;
    const bool live = m != 0ull;
    const unsigned active = __ballot_sync(0xFFFFFFFFu, live);
    if (!live) return;
    const uint32_t mover = (__ldg(&states[root].meta) >> 2) & 3u;
    const unsigned peers = __match_any_sync(active, d);                       // lanes of this warp that roll out the same determinization
    const bool leader = (uint32_t)(__ffs((int)peers) - 1) == (threadIdx.x & 31u);
    const RngKey key = make_key(rp, root, (rp.first_sub + d) * n_rollouts + r, true);
    const uint64_t out_row = (root * n_det + d) * N_ACTIONS;
    int best_v = 0;
    uint32_t best_a = ACTION_NONE, k = 0;
    for (uint64_t mm = m; mm && k < PIMC_MAX_LEGAL; mm &= mm - 1ull, ++k) {
        const uint32_t a = ffs0ll(mm);
        const PimcEntry& e = ws[row * PIMC_MAX_LEGAL + k];
        int32_t p[4];
        if (e.live) { FdoLive g = e.g; fdo_play_to_end<false, false, false, false>(g, key, &e.rs, lut); fdo_final_points(g, p); }
        else { p[0] = e.pts[0]; p[1] = e.pts[1]; p[2] = e.pts[2]; p[3] = e.pts[3]; }
        const int v = (mover & 2u) ? ((mover & 1u) ? p[3] : p[2]) : ((mover & 1u) ? p[1] : p[0]);
        const int tot = __reduce_add_sync(peers, v);
        if (leader && value_out) atomicAdd(value_out + out_row + a, (unsigned long long)(long long)tot);
        if (best_a == ACTION_NONE || v > best_v) { best_a = a; best_v = v; }   // first among equals
    }
    if (visits_out && best_a != ACTION_NONE) {
        const unsigned same = __match_any_sync(peers, best_a);                // peers that voted for the same action: one atomic for all
        if ((uint32_t)(__ffs((int)same) - 1) == (threadIdx.x & 31u)) atomicAdd(visits_out + out_row + best_a, (uint32_t)__popc(same));
    }
}

// Fuse per root: one block of 64 threads per root, rows staged 64 at a time in shared memory.
//   MaxN     the rank of every (row, allowed action) pair is computed by its own thread and added to the action's integer rank sum
//            (shared-memory atomics; integers, so the sums are those of the sequential fuse_max_n), thread 0 takes the first minimum
//   Average  thread a owns action a and adds the rows' f32 quotients in ROW ORDER (the reference's accumulation order, bit for bit);
//            the row totals come from the staging pass
// The statistics kernel of the sharded decision (rank sums, visit sums, successful rows) uses the same staging.
constexpr int FUSE_THREADS = 64;
struct FuseStage {
    uint32_t v[FUSE_THREADS][N_ACTIONS];     // 64 rows
    unsigned long long tot[FUSE_THREADS];
    uint8_t ok[FUSE_THREADS];
};
// loads rows [r0, r0 + 64) of the root (coalesced), their totals and success flags; returns the number of rows staged
__device__ __forceinline__ uint32_t fuse_stage_rows(FuseStage& sm, const uint32_t* __restrict__ v, const uint8_t* __restrict__ st, uint32_t r0, uint32_t n_rows) {
    const uint32_t cnt = min((uint32_t)FUSE_THREADS, n_rows - r0);
    __syncthreads();                                                          // the previous chunk has been consumed
    for (uint32_t e = threadIdx.x; e < cnt * N_ACTIONS; e += FUSE_THREADS) (&sm.v[0][0])[e] = v[(size_t)r0 * N_ACTIONS + e];
    __syncthreads();
    if (threadIdx.x < cnt) {
        unsigned long long t = 0;
        for (uint32_t a = 0; a < N_ACTIONS; ++a) t += sm.v[threadIdx.x][a];
        sm.tot[threadIdx.x] = t;
        sm.ok[threadIdx.x] = (!st || st[r0 + threadIdx.x] == 0u) ? 1 : 0;
    }
    __syncthreads();
    return cnt;
}
__global__ void __launch_bounds__(FUSE_THREADS)
fuse_kernel(uint32_t strategy, uint64_t n_roots, uint32_t n_det, const uint32_t* __restrict__ visits, const uint8_t* __restrict__ status,
            const uint64_t* __restrict__ allowed, uint8_t* __restrict__ action_out, uint32_t* __restrict__ n_ok_out) {
    __shared__ FuseStage sm;
    __shared__ uint32_t rank_sum[N_ACTIONS];
    __shared__ float avg_sum[N_ACTIONS];
    __shared__ uint32_t n_ok_s;
    const uint64_t i = blockIdx.x;
    if (i >= n_roots) return;
    const uint32_t* v = visits + i * n_det * N_ACTIONS;
    const uint8_t* st = status ? status + i * n_det : nullptr;
    const uint64_t al = allowed[i];
    if (threadIdx.x < N_ACTIONS) { rank_sum[threadIdx.x] = 0u; avg_sum[threadIdx.x] = 0.0f; }
    if (threadIdx.x == 0) n_ok_s = 0u;
    float my_sum = 0.0f;                                                      // Average: thread a < 39 owns action a
    for (uint32_t r0 = 0; r0 < n_det; r0 += FUSE_THREADS) {
        const uint32_t cnt = fuse_stage_rows(sm, v, st, r0, n_det);
        if (threadIdx.x < cnt && sm.ok[threadIdx.x]) atomicAdd(&n_ok_s, 1u);
        if (strategy == 0u) {
            for (uint32_t e = threadIdx.x; e < cnt * N_ACTIONS; e += FUSE_THREADS) {
                const uint32_t r = e / N_ACTIONS, a = e - r * N_ACTIONS;
                if (sm.ok[r] && ((al >> a) & 1ull)) atomicAdd(&rank_sum[a], fuse_rank(sm.v[r], al, a));
            }
        } else if (threadIdx.x < N_ACTIONS) {
            for (uint32_t r = 0; r < cnt; ++r)
                if (sm.ok[r]) my_sum = f32_add(my_sum, f32_div((float)sm.v[r][threadIdx.x], (float)sm.tot[r]));
        }
    }
    if (strategy != 0u && threadIdx.x < N_ACTIONS) avg_sum[threadIdx.x] = my_sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t best = 0;
        if (strategy == 0u) {
            uint32_t best_sum = 0xFFFFFFFFu;
            for (uint32_t a = 0; a < N_ACTIONS; ++a) {
                const uint32_t sum = ((al >> a) & 1ull) ? rank_sum[a] : 0xFFFFFFFFu;   // not allowed ⇒ u32::MAX (:31-38)
                if (a == 0u || sum < best_sum) { best = a; best_sum = sum; }           // min_by_key keeps the FIRST minimum (:65-70)
            }
        } else {
            for (uint32_t a = 1; a < N_ACTIONS; ++a) if (!(avg_sum[best] > avg_sum[a])) best = a;   // LAST of equal maxima, NaN as Equal (:113-119)
        }
        action_out[i] = (uint8_t)(n_ok_s == 0u ? ACTION_NONE : best);
        if (n_ok_out) n_ok_out[i] = n_ok_s;
    }
}
__global__ void __launch_bounds__(FUSE_THREADS)
root_stats_kernel(uint64_t n_roots, uint32_t n_det, const uint32_t* __restrict__ visits, const uint8_t* __restrict__ status,
                  const uint64_t* __restrict__ allowed, long long* __restrict__ stats, int accumulate) {
    __shared__ FuseStage sm;
    __shared__ unsigned long long acc[ROOT_STATS];
    const uint64_t i = blockIdx.x;
    if (i >= n_roots) return;
    long long* out = stats + i * ROOT_STATS;
    const uint32_t* v = visits + i * n_det * N_ACTIONS;
    const uint8_t* st = status ? status + i * n_det : nullptr;
    const uint64_t al = allowed[i];
    for (uint32_t k = threadIdx.x; k < ROOT_STATS; k += FUSE_THREADS) acc[k] = 0ull;
    for (uint32_t r0 = 0; r0 < n_det; r0 += FUSE_THREADS) {
        const uint32_t cnt = fuse_stage_rows(sm, v, st, r0, n_det);
        if (threadIdx.x < cnt && sm.ok[threadIdx.x]) atomicAdd(&acc[2u * N_ACTIONS], 1ull);
        for (uint32_t e = threadIdx.x; e < cnt * N_ACTIONS; e += FUSE_THREADS) {
            const uint32_t r = e / N_ACTIONS, a = e - r * N_ACTIONS;
            if (!sm.ok[r]) continue;
            if ((al >> a) & 1ull) atomicAdd(&acc[a], (unsigned long long)fuse_rank(sm.v[r], al, a));
            atomicAdd(&acc[N_ACTIONS + a], (unsigned long long)sm.v[r][a]);
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < ROOT_STATS; k += FUSE_THREADS) out[k] = (accumulate ? out[k] : 0ll) + (long long)acc[k];
}
__global__ void __launch_bounds__(STATE_THREADS)
root_pick_kernel(uint32_t strategy, uint64_t n_roots, const long long* __restrict__ stats, const uint64_t* __restrict__ allowed, uint8_t* __restrict__ action_out) {
    uint64_t i = (uint64_t)blockIdx.x * STATE_THREADS + threadIdx.x;
    if (i >= n_roots) return;
    action_out[i] = (uint8_t)root_stats_pick(strategy, stats + i * ROOT_STATS, allowed[i]);
}

}  // namespace dk
