/*
 * doko_cuda.h — C ABI of libdoko_cuda.so, the B200 (sm_100a) batched Doppelkopf simulator.
 *
 * This is the drop-in boundary for ONE hot path of theodm/master_doko_reinforcement_learning:
 * game transition -> determinization -> random rollouts -> observation encode.
 * The reference has no FFI boundary; its seam is a set of Rust traits over by-value state structs.
 * Each entry point below names the reference interface it replaces (paths relative to the
 * reference checkout).  Everything is batch-first, works on caller-owned buffers, takes an explicit
 * CUDA stream and reports errors through integer codes (it never aborts across the FFI).
 *
 * There is NO CPU fallback: every compute entry point launches sm_100a kernels and returns
 * DK_ERR_NO_DEVICE / DK_ERR_CUDA if that is impossible.
 *
 * Memory spaces: pointers marked [dev] are device pointers on the context's GPU; the *_host entry
 * points take [host] pointers and perform the host<->device copies themselves (they are what a
 * Rust/ctypes caller with ordinary slices binds).
 */
#ifndef DOKO_CUDA_H
#define DOKO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DK_API __attribute__((visibility("default")))
#else
#define DK_API
#endif

#define DK_VERSION_MAJOR 0
#define DK_VERSION_MINOR 3   /* 0.3: narrow observation rows (dk_encode_narrow, dk_step_random_encode_narrow); the parity stream draws the
                              * deal, a trick's card picks and card_matching's rule 4 as chained draws — same distributions,
                              * different games per seed than 0.2 */

typedef struct dk_ctx dk_ctx;
typedef int32_t dk_status;
typedef void* dk_stream; /* cudaStream_t; NULL = the context's own stream */

enum {
    DK_OK = 0,
    DK_ERR_INVALID_ARGUMENT = 1,
    DK_ERR_CUDA = 2,       /* a CUDA runtime call or kernel failed; see dk_last_error() */
    DK_ERR_NO_DEVICE = 3,  /* no usable sm_100 GPU: the library has no CPU path */
    DK_ERR_NCCL = 4,
    DK_ERR_UNSUPPORTED = 5
};

/* Rule engines. */
enum {
    DK_DOKO = 0, /* rs-doko: normal game + wedding, no solos, no announcements (rs-doko/src/state/state.rs) */
    DK_FDO = 1   /* rs-full-doko: full DKV rules (rs-full-doko/src/state/state.rs) */
};

/* ------------------------------------------------------------------------------------------------
 * Domain encoding (identical to the reference's).
 *
 * Card id c = suit*6 + rank; suits Diamond 0, Heart 1, Club 2, Spade 3; ranks 9,10,J,Q,K,A = 0..5
 * (rs-full-doko/src/card/cards.rs:7-36).  A hand is a 48-bit board: copy A of card c is bit c,
 * copy B is bit c+24 (rs-full-doko/src/hand/hand.rs:20,107).  Seats: BOTTOM 0, LEFT 1, TOP 2,
 * RIGHT 3, play proceeds +1 mod 4 (rs-full-doko/src/player/player.rs:11-16,67-72).
 *
 * Action index = bit index of the reference's FdoAction / DoAction (rs-full-doko/src/action/action.rs:9-54,
 * rs-doko/src/action/action.rs:7-38): 0..23 play card c; 24 Healthy; 25 Wedding; 26 Diamonds-, 27 Hearts-,
 * 28 Spades-, 29 Clubs-, 30 Trumpless-, 31 Queens-, 32 Jacks-Solo; 33 Re/Kontra; 34 No90; 35 No60; 36 No30;
 * 37 Black; 38 NoAnnouncement.  DK_DOKO uses 0..25 only.  A legal mask has bit i set iff action i is legal.
 * ------------------------------------------------------------------------------------------------ */
#define DK_NUM_ACTIONS_FDO 39
#define DK_NUM_ACTIONS_DOKO 26
#define DK_ACTION_HEALTHY 24
#define DK_ACTION_WEDDING 25
#define DK_ACTION_RE_CONTRA 33
#define DK_ACTION_NO_ANNOUNCEMENT 38

/* FdoPhase (rs-full-doko/src/basic/phase.rs:6-11).  DK_DOKO states use the same codes (its own enum
 * is Reservation 0, PlayCard 1, Finished 2 — rs-doko/src/basic/phase.rs; dk_encode emits that). */
enum { DK_PHASE_RESERVATION = 0, DK_PHASE_ANNOUNCEMENT = 1, DK_PHASE_PLAY_CARD = 2, DK_PHASE_FINISHED = 3 };

/* FdoGameType (rs-full-doko/src/game_type/game_type.rs:6-20). */
enum {
    DK_GT_NORMAL = 0, DK_GT_WEDDING = 1, DK_GT_DIAMONDS_SOLO = 2, DK_GT_HEARTS_SOLO = 3, DK_GT_SPADES_SOLO = 4,
    DK_GT_CLUBS_SOLO = 5, DK_GT_TRUMPLESS_SOLO = 6, DK_GT_QUEENS_SOLO = 7, DK_GT_JACKS_SOLO = 8, DK_GT_NONE = 15
};

/* FdoReservation (rs-full-doko/src/reservation/reservation.rs:11-24) — note: NOT the action order. */
enum {
    DK_RES_HEALTHY = 0, DK_RES_WEDDING = 1, DK_RES_DIAMONDS_SOLO = 2, DK_RES_HEARTS_SOLO = 3, DK_RES_SPADES_SOLO = 4,
    DK_RES_CLUBS_SOLO = 5, DK_RES_QUEENS_SOLO = 6, DK_RES_JACKS_SOLO = 7, DK_RES_TRUMPLESS_SOLO = 8, DK_RES_NONE = 0xFF
};

/* Announcement level codes (FdoAnnouncement, rs-full-doko/src/announcement/announcement.rs:12-22). */
enum { DK_ANN_NONE = 0, DK_ANN_RE_CONTRA = 1, DK_ANN_NO90 = 2, DK_ANN_NO60 = 3, DK_ANN_NO30 = 4, DK_ANN_BLACK = 5, DK_ANN_COUNTER = 6 };

/* FdoTeamState tag (rs-full-doko/src/team/team_logic.rs:11-28). */
enum { DK_TEAM_IN_RESERVATIONS = 0, DK_TEAM_WEDDING_UNSOLVED = 1, DK_TEAM_WEDDING_SOLVED = 2, DK_TEAM_NO_WEDDING = 3 };

/* ------------------------------------------------------------------------------------------------
 * dk_state — one game, 128 bytes, 16-byte aligned POD.  Replaces the by-value `FdoState`
 * (rs-full-doko/src/state/state.rs:24-75, ~1.2 KB) and `DoState` (rs-doko/src/state/state.rs:29-76).
 * It holds the complete history needed by the encoders and the determinizer.
 * ------------------------------------------------------------------------------------------------ */
typedef struct dk_state {
    uint64_t hands[4];          /*   0: current hands by absolute seat (48-bit boards) */
    uint8_t cards[48];          /*  32: played cards in play order, card id 0..23; 0xFF = not played yet */
    uint16_t announcements[12]; /*  80: k-th call: bits 0-5 card_index, 6-7 seat, 8-10 level code; 0xFFFF = unused
                                 *      (FdoAnnouncementOccurrence, announcement.rs:38-43) */
    uint8_t reservations[4];    /* 104: reservations in PLAY order from the game's start seat; DK_RES_NONE = not made.
                                 *      DK_FDO: DK_RES_* codes; DK_DOKO: DoReservation (Wedding 0, Healthy 1) */
    uint32_t tricks;            /* 108: bits 2t..2t+1 start seat of trick t (t<12); bits 24-27 #tricks started; bits 28-31 #calls */
    uint8_t eyes[4];            /* 112: eyes captured per absolute seat (completed tricks only) */
    uint16_t num_tricks;        /* 116: bits 4p..4p+3 tricks won by seat p */
    uint8_t card_index;         /* 118: number of cards played (0..48) */
    uint8_t n_reservations;     /* 119: number of reservations made (0..4) */
    int8_t points[4];           /* 120: player_points per absolute seat; valid iff phase == DK_PHASE_FINISHED */
    uint32_t meta;              /* 124: bits 0-1 phase | 2-3 current seat | 4-5 game start seat | 6-9 game type |
                                 *      10-11 team tag | 12-13 wedding seat | 14-15 solved_trick_index | 16-19 re seats mask |
                                 *      20-22 re lowest call | 23-25 kontra lowest call | 26-28 turns without call |
                                 *      29-30 start seat of the running announcement round | 31 reserved (0) */
} dk_state;

/* Observation layouts for dk_encode. */
enum {
    DK_LAYOUT_DO110 = 0,    /* rs-doko-embeddings/src/encode_state.rs:84-187  (110 tokens, DK_DOKO) */
    DK_LAYOUT_DO114 = 1,    /* rs-doko-embeddings/src/encode_state.rs:189-317 (114 tokens, DK_DOKO) */
    DK_LAYOUT_FDO_PI311 = 2 /* rs-doko-networks/src/full_doko/var1/encode_pi.rs:27-216 ([i64;311], DK_FDO) */
};
#define DK_OBS_LEN_DO110 110
#define DK_OBS_LEN_DO114 114
#define DK_OBS_LEN_FDO_PI311 311

/* Flags. */
#define DK_PLAYOUT_WITH_ANNOUNCEMENTS 1u /* FdoState::random_action_for_current_player (state.rs:378-399); without it the
                                            _no_announcement variant (state.rs:401-431) as used by random_rollout */
#define DK_STEP_SKIP_SINGLE 0x100u       /* dk_step_random_encode: apply with skip_single (same meaning as DK_APPLY_SKIP_SINGLE) */
#define DK_APPLY_SKIP_SINGLE 1u          /* FdoAzEnvState::take_action_by_action_index(.., skip_single=true, ..)
                                            (rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:121-154) */

/* Random stream: Philox4x32-10, key = seed, counter = (unit_lo, unit_hi, site<<16 | block, epoch).
 * unit = first_id + index of the game / info-state / leaf in the batch, so results do not depend on
 * how a batch is sharded across GPUs.  See DESIGN.md "Philox parity contract". */
typedef struct dk_rng {
    uint64_t seed;
    uint64_t first_id;
    uint32_t epoch;
    uint32_t first_sub; /* offset added to the sample / rollout number (unit_hi) of dk_determinize and dk_leaf_rollouts, so that the
                           samples of one unit can be split over several calls or GPUs; 0 elsewhere */
} dk_rng;

/* ---- context -------------------------------------------------------------------------------
 * One context per (process, GPU).  A context owns its scratch buffers (host entry points, PIMC workspace, UCT ln table): keep ONE call in
 * flight per context — calls on different streams of the same context must be ordered by the caller; use one context per host thread
 * for concurrent callers.  dk_destroy releases everything the context allocated, including an NCCL communicator left open.
 * Alignment: record arrays (dk_state*) and every 16-byte vector output named below must be 16-byte aligned, 32-bit outputs 4-byte aligned;
 * a misaligned pointer is refused with DK_ERR_INVALID_ARGUMENT (it would fault on the device). */
DK_API dk_status dk_init(int device, dk_ctx** out);
DK_API dk_status dk_destroy(dk_ctx* ctx);
DK_API const char* dk_last_error(const dk_ctx* ctx); /* message of the last failing call on this context */
DK_API const char* dk_version(void);
DK_API dk_status dk_device_info(const dk_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
DK_API dk_status dk_synchronize(dk_ctx* ctx, dk_stream stream);
/* Number of kernels this library has launched on this context since dk_init (bench.py's gpu_launches). */
DK_API uint64_t dk_launch_count(const dk_ctx* ctx);

/* ---- game construction -----------------------------------------------------------------------
 * replaces FdoState::new_game (rs-full-doko/src/state/state.rs:169-178) /
 *          DoState::new_game  (rs-doko/src/state/state.rs:159-168): start seat, then a 48-card shuffle. */
DK_API dk_status dk_new_games(dk_ctx* ctx, int engine, size_t n, const dk_rng* rng, dk_state* out /*[dev] n*/, dk_stream stream);
/* replaces new_game_from_hand_and_start_player (rs-full-doko/src/state/state.rs:125-166) */
DK_API dk_status dk_from_deals(dk_ctx* ctx, int engine, size_t n, const uint64_t* hands /*[dev] n*4*/,
                        const uint8_t* start /*[dev] n*/, dk_state* out /*[dev] n*/, dk_stream stream);

/* ---- transition --------------------------------------------------------------------------------
 * replaces FdoAllowedActions::calculate_allowed_actions (rs-full-doko/src/action/allowed_actions.rs:68-169) as
 *          surfaced by McEnvState::allowed_actions(first_expansion=true) (rs-doko-mcts/src/env/env_state.rs:15)
 *          and AzEnvState::allowed_actions_by_action_index (rs-doko-alpha-zero/src/env/env_state.rs:20-24);
 *          rs-doko: calculate_allowed_actions_in_normal_game (rs-doko/src/action/allowed_actions.rs:131-198).
 * A finished game yields mask 0. */
DK_API dk_status dk_legal_mask(dk_ctx* ctx, int engine, size_t n, const dk_state* states /*[dev]*/, uint64_t* mask_out /*[dev] n*/,
                        dk_stream stream);
/* AzEnvState::allowed_actions_by_action_index(is_secondary, epoch) and AzEnvState::number_of_allowed_actions(epoch) of FdoAzEnvState
 * (rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:80-119; DK_FDO records): mask_out[i] (nullable) = the legal mask without the
 * five announcement calls (actions 33..37) when is_secondary != 0 || az_epoch < DK_AZ_MIN_EPOCH; n_allowed_out[i] (nullable) = the number
 * of allowed actions under the EPOCH filter alone (the reference's number_of_allowed_actions ignores is_secondary). */
#define DK_AZ_MIN_EPOCH 10u
DK_API dk_status dk_legal_mask_az(dk_ctx* ctx, size_t n, const dk_state* states /*[dev]*/, int is_secondary, uint64_t az_epoch,
                                  uint64_t* mask_out /*[dev] n or NULL*/, uint8_t* n_allowed_out /*[dev] n or NULL*/, dk_stream stream);
/* AzEnvState::id() (full_doko.rs:156-167): FxHasher64 (fxhash 0.2.1: hash = (rotl(hash, 5) ^ word) * 0x517cc1b727220a95 per 64-bit word)
 * over the sixteen little-endian 64-bit words of the record and then the last action as one more word (last_action == NULL or
 * DK_ACTION_NONE = None).  The record is canonical (one byte pattern per game state), so equal states <=> equal hash input, which is
 * what id() promises; the values differ from the reference's, which hashes the in-memory layout of its own structs. */
DK_API dk_status dk_state_id(dk_ctx* ctx, size_t n, const dk_state* states /*[dev]*/, const uint8_t* last_action /*[dev] n or NULL*/,
                             uint64_t* id_out /*[dev] n*/, dk_stream stream);
/* FdoAllowedActions::random (rs-game-utils/src/bit_flag.rs:86-94) over the legal set of the seat to move WITHOUT playing it: the draw of
 * the random policies (FdoState::random_action_for_current_player[_no_announcement], rs-full-doko/src/state/state.rs:378-431; flags =
 * DK_PLAYOUT_WITH_ANNOUNCEMENTS keeps the calls in the set) and of DefaultImpiPolicy's fallback when no determinization succeeded
 * (rs-doko-py-bridge/src/compare_impi/compare_impi.rs:357-368; flags = 0).  Stream position: SITE_STEP word 0 of unit first_id + i, i.e.
 * dk_step_random_encode with the same rng plays exactly this action.  A finished game yields DK_ACTION_NONE. */
DK_API dk_status dk_random_action(dk_ctx* ctx, int engine, size_t n, const dk_state* states /*[dev]*/, const dk_rng* rng, uint32_t flags,
                                  uint8_t* action_out /*[dev] n*/, dk_stream stream);
/* replaces FdoState::play_action (rs-full-doko/src/state/state.rs:208-358) / McEnvState::by_action /
 *          AzEnvState::take_action_by_action_index; rs-doko: DoState::play_action (rs-doko/src/state/state.rs:189-309).
 * An illegal action (the reference would panic) sets err_out[i] != 0 and leaves state i unchanged. */
DK_API dk_status dk_apply(dk_ctx* ctx, int engine, size_t n, dk_state* states /*[dev] in/out*/, const uint8_t* action_idx /*[dev] n*/,
                   uint32_t flags, uint8_t* err_out /*[dev] n, may be NULL*/, dk_stream stream);
/* replaces McEnvState::{is_terminal, rewards_or_none} (rs-doko-mcts/src/env/envs/env_state_full_doko.rs:75-79,191-196):
 * done_out[i] = phase == Finished; points_out[i] = player_points (0 when not finished). */
DK_API dk_status dk_terminal(dk_ctx* ctx, int engine, size_t n, const dk_state* states /*[dev]*/, uint8_t* done_out /*[dev] n*/,
                      int32_t* points_out /*[dev] n*4*/, dk_stream stream);

/* ---- observation encode --------------------------------------------------------------------------
 * replaces AzEnvState::encode_into_memory (rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:71-78) =
 *          encode_state_pi(state, state.observation_for_current_player()); rs-doko: encode_state[_with_reservations].
 * Row i is written at out + i*row_stride (in elements, >= layout length); batch-major like the reference's
 * flattened Vec<i64> (rs-doko-alpha-zero/src/alpha_zero/batch_processor/network_batch_processor.rs:85-91). */
DK_API dk_status dk_encode(dk_ctx* ctx, int layout, size_t n, const dk_state* states /*[dev]*/, int64_t* out /*[dev]*/,
                    size_t row_stride, dk_stream stream);
/* One lock-step self-play env step: legal mask -> one Philox draw -> play_action -> encode_state_pi of the new
 * state (self_play.rs:76-190 without the NN).  action_out may be NULL. */
DK_API dk_status dk_step_random_encode(dk_ctx* ctx, size_t n, dk_state* states /*[dev] in/out*/, const dk_rng* rng, uint32_t flags,
                                int64_t* obs_out /*[dev]*/, size_t row_stride, uint8_t* action_out /*[dev] n*/,
                                dk_stream stream);

/* Narrow observation rows: the same token VALUES as dk_encode / dk_step_random_encode (every value is < 256), written as int32
 * (elem_bytes = 4) or uint8 (elem_bytes = 1) instead of the reference's i64.  The reference's Python side narrows its Vec<i64> rows to
 * int32 before the network sees them (rs-doko-py-bridge/.../az_doko.py:369); a caller that keeps the tokens on the device asks for
 * that row directly and the HBM-bound encoders move 1244 / 311 B per observation instead of 2488 B.  Dense rows only (row stride =
 * layout length); `out` must be 32-byte aligned.  Not a drop-in for encode_into_memory (that is dk_encode): an addition beside it. */
DK_API dk_status dk_encode_narrow(dk_ctx* ctx, int layout, int elem_bytes, size_t n, const dk_state* states /*[dev]*/,
                           void* out /*[dev] n * len * elem_bytes*/, dk_stream stream);
DK_API dk_status dk_step_random_encode_narrow(dk_ctx* ctx, size_t n, dk_state* states /*[dev] in/out*/, const dk_rng* rng, uint32_t flags,
                                       int elem_bytes, void* obs_out /*[dev] n * 311 * elem_bytes*/, uint8_t* action_out /*[dev] n, may be NULL*/,
                                       dk_stream stream);

/* ---- playouts --------------------------------------------------------------------------------------
 * replaces the loop `while !state.random_action_for_current_player[_no_announcement](rng) {}` and
 *          McEnvState::random_rollout (rs-doko-mcts/src/env/envs/env_state_full_doko.rs:198-220,
 *          env_state_doko.rs:150-169).  states == NULL plays fresh games dealt from the stream
 *          (new_game + playout fused on-chip).  points_out[i] = player_points; steps_out[i] = number of
 *          play_action calls ("game steps"); either may be NULL. */
DK_API dk_status dk_playout(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states /*[dev] or NULL*/,
                     const dk_rng* rng, int32_t* points_out /*[dev] n*4*/, uint32_t* steps_out /*[dev] n*/,
                     dk_stream stream);
/* Verification entry (DK_DOKO): fresh playouts that also record the 52 action ids per game (trace_out[n*52]) and
 * aux_out[n*4] = (wedding flag, re seats mask, eyes packed 8 bits/seat, tricks packed 4 bits/seat) — the per-game outputs
 * BASELINE config 1 is checked on. */
DK_API dk_status dk_playout_trace(dk_ctx* ctx, int engine, size_t n, const dk_rng* rng, int32_t* points_out /*[dev] n*4*/,
                                  uint8_t* trace_out /*[dev] n*52*/, uint32_t* aux_out /*[dev] n*4*/, dk_stream stream);
/* Same, results copied to HOST buffers inside the call (what a plain Rust slice caller binds). */
DK_API dk_status dk_playout_host(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states_host /*[host] or NULL*/,
                          const dk_rng* rng, int32_t* points_out_host /*[host] n*4*/, uint32_t* steps_out_host /*[host] n*/);

/* Same with the compact, lossless host-transfer form: int8 points (|points| < 128) and uint8 step counts (< 256) — 5 bytes per game
 * instead of 20 over PCIe. */
DK_API dk_status dk_playout_host_compact(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states_host /*[host] or NULL*/,
                                         const dk_rng* rng, int8_t* points_out_host /*[host] n*4*/, uint8_t* steps_out_host /*[host] n*/);

/* Packed host-transfer form: 2 bytes of points per game (+ 1 byte of steps when steps_out_host != NULL) instead of 5 / 20.  The four
 * player_points take two values (Re seats / Kontra seats) and sum to zero (FdoEndOfGameStats::calculate, rs-full-doko/src/stats/stats.rs:215-231;
 * rs-doko/src/stats/stats.rs:25-135), so v = (uint8)points[0] | same << 8, same bit j-1 = (points[j] == points[0]) for j = 1..3, is lossless:
 * dk_unpack_points restores the four values. */
DK_API dk_status dk_playout_host_packed(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states_host /*[host] or NULL*/,
                                        const dk_rng* rng, uint16_t* points_packed_out_host /*[host] n*/, uint8_t* steps_out_host /*[host] n or NULL*/);
static inline void dk_unpack_points(uint16_t v, int32_t points[4]) {
    const int32_t a = (int8_t)(v & 0xFFu);
    const uint32_t same = (v >> 8) & 7u;
    const int32_t k = 1 + (int32_t)(same & 1u) + (int32_t)((same >> 1) & 1u) + (int32_t)((same >> 2) & 1u);   /* seats that score like seat 0 */
    const int32_t b = k == 4 ? a : -(k * a) / (4 - k);                                                        /* zero sum */
    points[0] = a;
    points[1] = (same & 1u) ? a : b;
    points[2] = (same & 2u) ? a : b;
    points[3] = (same & 4u) ? a : b;
}

/* Device-reduced playouts for evaluator-style callers (what EvFullDokoSingleGameEvaluationResult rows are aggregated to,
 * rs-doko-evaluator/src/full_doko/evaluate_single_game.rs:50-75): the games are played like dk_playout and only their statistics leave the
 * GPU — 2160 bytes per call instead of bytes per game.  All integer, order independent, identical for any sharding of the batch:
 * summing the structs of several calls / ranks gives the statistics of the union (accumulate != 0 adds to *stats instead of overwriting). */
typedef struct dk_playout_stats {
    uint64_t games;            /* games played */
    uint64_t game_steps;       /* sum of the number of play_action calls (number_of_actions) */
    int64_t point_sum[4];      /* sum of player_points per absolute seat */
    uint64_t point_sq_sum[4];  /* sum of player_points^2 per seat */
    uint64_t wins[4];          /* games with player_points > 0 per seat */
    uint64_t step_hist[256];   /* histogram of the number of actions per game (bin 255 = 255 and more) */
} dk_playout_stats;
DK_API dk_status dk_playout_summary(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states /*[dev] or NULL*/, const dk_rng* rng,
                                    dk_playout_stats* stats /*[dev]*/, int accumulate, dk_stream stream);
/* Same with host buffers: states_host (or NULL = fresh games) is copied in, *stats_out_host is written before the call returns. */
DK_API dk_status dk_playout_summary_host(dk_ctx* ctx, int engine, uint32_t flags, size_t n, const dk_state* states_host /*[host] or NULL*/,
                                         const dk_rng* rng, dk_playout_stats* stats_out_host /*[host]*/);

/* ---- determinization ---------------------------------------------------------------------------------
 * DK_DOKO replaces sample_assignment_full (rs-doko-assignment/src/assignment.rs:458-581): hands only; reservations_out repeats
 *          the real reservations (DoReservation codes by absolute seat).
 * DK_FDO replaces card_matching (rs-full-doko/src/matching/card_matching.rs:241-467) as called by CAPSampling::sample
 *          (rs-doko-py-bridge/src/compare_impi/compare_impi.rs:64-83); observer = seat to move.
 * For info-state i and sample s (unit = first_id + i, unit_hi = s):
 *   hands_out[(i*S+s)*4 + seat], reservations_out[(i*S+s)*4 + seat] (DK_RES_* by ABSOLUTE seat, DK_RES_NONE if the
 *   seat has not declared yet), status_out[i*S+s] != 0 for a dead end (the reference would panic). */
DK_API dk_status dk_determinize(dk_ctx* ctx, int engine, size_t n_info, size_t samples_per_info, const dk_state* states /*[dev]*/,
                         const dk_rng* rng, uint64_t* hands_out /*[dev]*/, uint8_t* reservations_out /*[dev]*/,
                         uint8_t* status_out /*[dev]*/, dk_stream stream);
/* Leaf-parallel rollouts (new design over random_rollout): for leaf i run R rollouts (unit = first_id + i,
 * unit_hi = r), each first determinized when `determinize` != 0, then played with the _no_announcement policy.
 * point_sum_out[i*4 + seat] = exact integer sum of player_points over the R rollouts. */
DK_API dk_status dk_leaf_rollouts(dk_ctx* ctx, size_t n_leaves, size_t rollouts_per_leaf, int determinize,
                           const dk_state* states /*[dev]*/, const dk_rng* rng, int64_t* point_sum_out /*[dev] n_leaves*4*/,
                           dk_stream stream);

/* encode_state_ipi (rs-doko-networks/src/full_doko/var1/encode_ipi.rs:48-306): the imperfect-information 311-token
 * layout of the autoregressive hand predictor, observer = seat to move.  Per game: assumed_hands[i*4 + seat] = cards guessed so far for
 * that seat (48-bit board; the observer's entry is ignored), assumed_reservations[i*4 + seat] = guessed DK_RES_* or DK_RES_NONE
 * (used only where the visible reservation is NotRevealed), next_player[i] = seat the next card is guessed for.  Same row layout as
 * DK_LAYOUT_FDO_PI311 (5 channels x 62 slots + 1).  err_out[i] != 0 (nullable) where the reference would panic: a guessed hand larger
 * than the real one (`hand.len() - assumed.len()`, :158) — the row is then truncated / zero-padded. */
DK_API dk_status dk_encode_ipi(dk_ctx* ctx, size_t n, const dk_state* states /*[dev]*/, const uint64_t* assumed_hands /*[dev] n*4*/,
                               const uint8_t* assumed_reservations /*[dev] n*4*/, const uint8_t* next_player /*[dev] n*/,
                               int64_t* out /*[dev] n*row_stride*/, size_t row_stride, uint8_t* err_out /*[dev] n or NULL*/, dk_stream stream);

/* ---- PIMC move decision ------------------------------------------------------------------
 * DefaultImpiPolicy::execute (rs-doko-py-bridge/src/compare_impi/compare_impi.rs:212-372): num_samples determinizations of the
 * info-state (CAPSampling = card_matching, :64-83), a per-sample policy that returns visit counts per action, and a PolicyFusionFn.
 *
 * dk_pimc_evaluate fills the per-sample policy slot (EvFullDokoPolicy::evaluate, rs-doko-evaluator/src/full_doko/policy/policy.rs:22-29)
 * with a flat Monte-Carlo search (new design; the reference plugs a UCT search in there): for root i, determinization d
 * (stream (first_id + i, first_sub + d) — exactly dk_determinize's sample) and rollout r, EVERY legal action a of the seat to move is
 * played on the determinized state and followed by a _no_announcement random_rollout on the stream
 * (first_id + i, (first_sub + d) * n_rollouts + r), the same for all actions (common random numbers);
 *   value_sum_out[(i*n_det + d)*39 + a] = sum over r of player_points[mover]   (exact integer),
 *   visits_out   [(i*n_det + d)*39 + a] = number of rollouts r in which a had the strictly greatest value (first in action order
 *                                         among equals) — each successful row sums to n_rollouts, like the visit counts of a search,
 *   status_out   [i*n_det + d] != 0 for a determinization dead end (a failed sample: ForwardPredResult::NotConsistent; row = 0).
 * A finished root yields zero rows with status 0. */
#define DK_N_ACTIONS 39u    /* FdoAction::COUNT (rs-full-doko/src/action/action.rs:13-54) */
#define DK_ACTION_NONE 0xFFu
DK_API dk_status dk_pimc_evaluate(dk_ctx* ctx, size_t n_roots, size_t n_det, size_t n_rollouts, const dk_state* states /*[dev]*/,
                                  const dk_rng* rng, uint32_t* visits_out /*[dev] or NULL*/, int64_t* value_sum_out /*[dev] or NULL*/,
                                  uint8_t* status_out /*[dev] or NULL*/, dk_stream stream);

/* PolicyFusionFn::fuse (policy_fusion.rs:9-16) per root: rows = the samples of one root, visits[(i*n_rows + s)*39 + a]; rows with
 * status != 0 (status may be NULL) are failed samples and are left out, as execute does (compare_impi.rs:318-330).
 *   DK_FUSE_MAX_N    PolicyFusionMaxN (:23-73): rank the allowed actions per row by visits (stable, descending), sum the ranks,
 *                    smallest sum wins, first index among equals.
 *   DK_FUSE_AVERAGE  PolicyFusionAverageStrategy (:79-123): sum of visits/total per row in f32, largest wins, LAST index among equals;
 *                    `allowed` is ignored, as in the reference.
 * action_out[i] = DK_ACTION_NONE when no sample of root i succeeded (the caller then falls back to a random non-announcement action,
 * compare_impi.rs:357-368); n_success_out[i] (nullable) = number of rows fused. */
#define DK_FUSE_MAX_N 0
#define DK_FUSE_AVERAGE 1
DK_API dk_status dk_fuse(dk_ctx* ctx, int strategy, size_t n_roots, size_t n_rows, const uint32_t* visits /*[dev]*/,
                         const uint8_t* status /*[dev] or NULL*/, const uint64_t* allowed /*[dev] n_roots, bit a = action a*/,
                         uint8_t* action_out /*[dev]*/, uint32_t* n_success_out /*[dev] or NULL*/, dk_stream stream);

/* Sharded decision: every rank evaluates its own determinizations of every root, reduces them to DK_ROOT_STATS int64 per root
 *   [0,39) MaxN rank sums of the allowed actions | [39,78) visit sums | [78] successful samples | [79] 0
 * (accumulate != 0 adds to `stats` instead of overwriting), the ranks sum them with dk_allreduce_root_stats, and dk_pimc_pick decides.
 * MaxN from the statistics is identical to dk_fuse over the union of the rows; Average is the arg-max of the summed visits (last among
 * equals), which is PolicyFusionAverageStrategy in exact arithmetic when every row has the same total.  KNOWN DIVERGENCE: with rows of
 * unequal totals (UCT trees that hit a terminal root early, mixed rollout counts) or when the reference's f32 sums tie or round differently,
 * the sharded Average pick can differ from dk_fuse / the reference on near-ties; gather the rows and call dk_fuse when bit parity with the
 * single-GPU decision matters (tests/test_gpu_pimc.py pins both behaviours).  dk_pimc_pick returns DK_ACTION_NONE for a root without a
 * successful sample: the reference then plays a random non-announcement action (compare_impi.rs:357-368) = dk_random_action(flags = 0),
 * which master_doko_reinforcement_learning_b200.sharding.pimc_decide applies. */
#define DK_ROOT_STATS 80u
DK_API dk_status dk_pimc_root_stats(dk_ctx* ctx, size_t n_roots, size_t n_rows, const uint32_t* visits /*[dev]*/, const uint8_t* status /*[dev] or NULL*/,
                                    const uint64_t* allowed /*[dev]*/, int64_t* stats /*[dev] n_roots*DK_ROOT_STATS*/, int accumulate, dk_stream stream);
DK_API dk_status dk_pimc_pick(dk_ctx* ctx, int strategy, size_t n_roots, const int64_t* stats /*[dev]*/, const uint64_t* allowed /*[dev]*/,
                              uint8_t* action_out /*[dev]*/, dk_stream stream);

/* ---- AlphaZero self-play driver -----------------------------------------------------------
 * self_play with ValueTarget::Default (rs-doko-alpha-zero/src/alpha_zero/train/self_play.rs:19-207) over FdoAzEnvState
 * (rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:43-172), for a batch of games in lock-step.  The experience buffers are
 * caller-owned device memory and take the place of states_buffer / policy_targets_buffer / value_targets_buffer (self_play.rs:51-53);
 * observations are encoded straight into their row, where the network batcher reads them (no Vec<i64>, no host round trip).
 * Row numbers are deterministic: turn-major, game order inside a turn.
 *
 * One turn:   dk_sp_begin_turn  → per game: is_terminal, allowed_actions_by_action_index(false, az_epoch), forced move
 *                                 (number_of_allowed_actions(az_epoch) == 1, :76) with the keep draw `gen::<f32>() < keep_prob` (:88;
 *                                 Philox SITE_KEEP word 0 of unit first_id + game, epoch rng->epoch), row assignment,
 *                                 encode_into_memory into the row, current_player, one-hot policy target of forced moves
 *             [search]          → the caller fills policy[n][39] (visits / visit_sum, :118-137) and action[n] for the games whose flags
 *                                 have neither DK_SP_DONE nor DK_SP_FORCED; dk_sp_turn_view exposes allowed / flags / rows (device
 *                                 pointers owned by the driver, valid until the next begin_turn).  dk_sp_uniform_search is the stand-in
 *                                 used by tests and benches: one draw over the allowed set (SITE_STEP word 0), uniform policy target.
 *             dk_sp_end_turn    → policy target into the row, take_action_by_action_index(action, false, az_epoch) (:101,:188);
 *                                 err_out[i] != 0 for an action outside the allowed set (state unchanged; the reference would panic).
 * After the last turn dk_sp_finalize writes every row's value target: the final rewards (player_points / 8 as f32) rotated to the
 * row's mover (RotArr::new_from_0(cp, rewards).extract(), :195-205). */
typedef struct dk_selfplay dk_selfplay;
typedef struct dk_sp_buffers {
    int64_t* states;   /* [dev] capacity * 311 */
    float* policy;     /* [dev] capacity * 39  */
    float* value;      /* [dev] capacity * 4   */
    uint8_t* player;   /* [dev] capacity       current_player of the row (current_player_vec, self_play.rs:58) */
    uint32_t* game;    /* [dev] capacity       game index of the row */
    size_t capacity;   /* rows; further rows are dropped (flag DK_SP_DROPPED) and counted */
} dk_sp_buffers;
#define DK_SP_DONE 1u
#define DK_SP_FORCED 2u
#define DK_SP_KEPT 4u
#define DK_SP_DROPPED 8u
#define DK_SP_SEARCH_FORCED 1u /* dk_sp_begin_turn flag: no forced-move shortcut (ValueTarget::Greedy / Avg search every turn, :76) */
DK_API dk_status dk_sp_create(dk_ctx* ctx, size_t max_games, const dk_sp_buffers* bufs, dk_selfplay** out);
DK_API dk_status dk_sp_destroy(dk_selfplay* sp);
DK_API dk_status dk_sp_reset(dk_selfplay* sp, dk_stream stream);
DK_API dk_status dk_sp_begin_turn(dk_selfplay* sp, size_t n, const dk_state* states /*[dev]*/, uint64_t az_epoch, float keep_prob, uint32_t flags,
                                  const dk_rng* rng, dk_stream stream);
DK_API dk_status dk_sp_turn_view(dk_selfplay* sp, const uint64_t** allowed /*[dev] n*/, const uint8_t** flags /*[dev] n*/, const int64_t** rows /*[dev] n, -1 = none*/);
DK_API dk_status dk_sp_uniform_search(dk_selfplay* sp, const dk_rng* rng, float* policy_out /*[dev] n*39*/, uint8_t* action_out /*[dev] n*/, dk_stream stream);
DK_API dk_status dk_sp_end_turn(dk_selfplay* sp, dk_state* states /*[dev] in/out*/, const float* policy /*[dev] n*39*/, const uint8_t* action /*[dev] n*/,
                                uint8_t* err_out /*[dev] n or NULL*/, dk_stream stream);
DK_API dk_status dk_sp_finalize(dk_selfplay* sp, const dk_state* states /*[dev]*/, dk_stream stream);
/* Synchronises the stream: rows recorded so far, rows dropped for lack of capacity, rows whose game was unfinished at the last finalize. */
DK_API dk_status dk_sp_counts(dk_selfplay* sp, uint64_t* rows, uint64_t* dropped, uint64_t* unfinished, dk_stream stream);

/* The reference's on-disk experience record for rows of the experience buffers: DBRecord {state, value, policy} through
 * bincode::serialize (rs-doko-alpha-zero/src/alpha_zero/net/experience_replay_buffer3.rs:11-20,94-121; bincode 1.3.3, heapless 0.8.0):
 *   u64 311 | 311 x i64 | u64 4 | 4 x f32 | u64 39 | 39 x f32, little endian, DK_REPLAY_RECORD_BYTES each, back to back in `out`
 * (4-byte aligned).  One D2H copy of `out` then holds exactly the values `append_slice` inserts into sled (the random u64 keys and the
 * store itself stay on the host side).
 * PINNING: the reference holds no golden bytes of this record and cannot be run here, so the layout is pinned to an independent encoder of
 * the bincode 1.x specification and its committed output (tests/golden/bincode_v1.py, dbrecord_golden.json) — not to bytes produced by the
 * Rust crate itself; treat the entry point as EXPERIMENTAL until a maintainer has round-tripped one record through bincode::deserialize. */
#define DK_REPLAY_RECORD_BYTES 2684u
DK_API dk_status dk_pack_replay_records(dk_ctx* ctx, size_t n_rows, const int64_t* states /*[dev] n_rows*311*/, const float* value /*[dev] n_rows*4*/,
                                        const float* policy /*[dev] n_rows*39*/, uint8_t* out /*[dev] n_rows*DK_REPLAY_RECORD_BYTES*/, dk_stream stream);

/* ---- UCT search ----------------------------------------------------------------------------
 * CachedMCTS::monte_carlo_tree_search (rs-doko-mcts/src/mcts/mcts.rs:160-232, node.rs:138-278) over McFullDokoEnvState
 * (rs-doko-mcts/src/env/envs/env_state_full_doko.rs:62-220) as EvFullDokoMCTSPolicy runs it (rs-doko-evaluator/src/full_doko/policy/
 * mcts_policy.rs:76-118): min-max-normalised Q + c*sqrt(ln N / n) selection, expand_single with a random unexpanded action, a
 * _no_announcement random_rollout from the new node, backpropagation of result[parent.current_player].
 * One tree per (root i, d), d < trees_per_root; determinize != 0 first replaces root i by its determinization first_sub + d (the
 * dk_determinize sample) — trees_per_root is then DefaultImpiPolicy's num_samples and visits_out feeds dk_fuse directly; with
 * determinize == 0 the given states are searched as they are (perfect information, mcts_pi_policy).  Iteration `it` of a tree uses
 * the Philox unit (first_id + i, (first_sub + d) * iterations + it): expansion pick = word 0 of site 9, rollout draws at their
 * state-derived ordinals.  Selection is bit-identical to the f64 arithmetic of the reference (ln from the host libm).
 *   visits_out[t*39 + a], values_out[t*39 + a] = win_score / visits as f32 (0 for actions without a child), t = i*trees_per_root + d
 *   action_out[t] = move with the most visits (last among equals in child order), DK_ACTION_NONE if the root has no child
 *   status_out[t] != 0: determinization dead end (row = 0).
 * workspace: caller-owned device memory, >= dk_uct_workspace_bytes(n_roots * trees_per_root, iterations) bytes (288 bytes per node —
 * a 160-byte statistics block and the 128-byte record —, iterations + 1 nodes per tree, plus 24 bytes of control words per tree);
 * iterations <= 2^24.  One iteration of all trees is three kernel launches (select + backpropagation, expansion, rollout). */
DK_API size_t dk_uct_workspace_bytes(size_t n_trees, size_t iterations);
DK_API dk_status dk_uct_search(dk_ctx* ctx, size_t n_roots, size_t trees_per_root, int determinize, size_t iterations, float uct_exploration_constant,
                               const dk_state* states /*[dev]*/, const dk_rng* rng, void* workspace /*[dev]*/, size_t workspace_bytes,
                               uint32_t* visits_out /*[dev] or NULL*/, float* values_out /*[dev] or NULL*/, uint8_t* action_out /*[dev] or NULL*/,
                               uint8_t* status_out /*[dev] or NULL*/, dk_stream stream);

/* ---- multi-GPU root statistics (the only exchange step) --------------------------------------
 * replaces the per-determinization fuse of PolicyFusion* (rs-doko-py-bridge/src/compare_impi/policy_fusion.rs:18-123):
 * integer sums over ranks, order-independent and bit-reproducible.  NCCL is loaded lazily (dlopen). */
typedef struct dk_nccl_id { char bytes[128]; } dk_nccl_id;
DK_API dk_status dk_comm_unique_id(dk_ctx* ctx, dk_nccl_id* out);
DK_API dk_status dk_comm_init(dk_ctx* ctx, int n_ranks, int rank, const dk_nccl_id* id);
DK_API dk_status dk_comm_destroy(dk_ctx* ctx);
DK_API dk_status dk_allreduce_root_stats(dk_ctx* ctx, size_t n_values, int64_t* values /*[dev] in/out, sum over ranks*/,
                                  dk_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* DOKO_CUDA_H */
