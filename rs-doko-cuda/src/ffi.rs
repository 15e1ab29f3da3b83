//! GENERATED from include/doko_cuda.h by rs-doko-cuda/tools/gen_ffi.py — do not edit; `gen_ffi.py --check` runs in the test suite.
//! Raw FFI of libdoko_cuda.so: every DK_API entry point, the POD structs and the constants of the C ABI.
#![allow(non_camel_case_types, non_upper_case_globals, dead_code)]
use std::ffi::{c_char, c_int, c_void};

pub type dk_status = i32;
pub type dk_stream = *mut c_void;

#[repr(C)]
pub struct dk_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct dk_selfplay {
    _private: [u8; 0],
}

#[repr(C, align(16))]
#[derive(Clone, Copy)]
pub struct dk_state {
    pub hands: [u64; 4],
    pub cards: [u8; 48],
    pub announcements: [u16; 12],
    pub reservations: [u8; 4],
    pub tricks: u32,
    pub eyes: [u8; 4],
    pub num_tricks: u16,
    pub card_index: u8,
    pub n_reservations: u8,
    pub points: [i8; 4],
    pub meta: u32,
}
const _: () = assert!(std::mem::size_of::<dk_state>() == 128);

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dk_rng {
    pub seed: u64,
    pub first_id: u64,
    pub epoch: u32,
    pub first_sub: u32,
}
const _: () = assert!(std::mem::size_of::<dk_rng>() == 24);

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dk_playout_stats {
    pub games: u64,
    pub game_steps: u64,
    pub point_sum: [i64; 4],
    pub point_sq_sum: [u64; 4],
    pub wins: [u64; 4],
    pub step_hist: [u64; 256],
}
const _: () = assert!(std::mem::size_of::<dk_playout_stats>() == 2160);

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dk_sp_buffers {
    pub states: *mut i64,
    pub policy: *mut f32,
    pub value: *mut f32,
    pub player: *mut u8,
    pub game: *mut u32,
    pub capacity: usize,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dk_nccl_id {
    pub bytes: [c_char; 128],
}
const _: () = assert!(std::mem::size_of::<dk_nccl_id>() == 128);

pub const DK_VERSION_MAJOR: i32 = 0;
pub const DK_VERSION_MINOR: i32 = 3;
pub const DK_NUM_ACTIONS_FDO: i32 = 39;
pub const DK_NUM_ACTIONS_DOKO: i32 = 26;
pub const DK_ACTION_HEALTHY: i32 = 24;
pub const DK_ACTION_WEDDING: i32 = 25;
pub const DK_ACTION_RE_CONTRA: i32 = 33;
pub const DK_ACTION_NO_ANNOUNCEMENT: i32 = 38;
pub const DK_OBS_LEN_DO110: i32 = 110;
pub const DK_OBS_LEN_DO114: i32 = 114;
pub const DK_OBS_LEN_FDO_PI311: i32 = 311;
pub const DK_PLAYOUT_WITH_ANNOUNCEMENTS: u32 = 1;
pub const DK_STEP_SKIP_SINGLE: u32 = 0x100;
pub const DK_APPLY_SKIP_SINGLE: u32 = 1;
pub const DK_AZ_MIN_EPOCH: u32 = 10;
pub const DK_N_ACTIONS: u32 = 39;
pub const DK_ACTION_NONE: u32 = 0xFF;
pub const DK_FUSE_MAX_N: i32 = 0;
pub const DK_FUSE_AVERAGE: i32 = 1;
pub const DK_ROOT_STATS: u32 = 80;
pub const DK_SP_DONE: u32 = 1;
pub const DK_SP_FORCED: u32 = 2;
pub const DK_SP_KEPT: u32 = 4;
pub const DK_SP_DROPPED: u32 = 8;
pub const DK_SP_SEARCH_FORCED: u32 = 1;
pub const DK_REPLAY_RECORD_BYTES: u32 = 2684;
pub const DK_OK: i32 = 0;
pub const DK_ERR_INVALID_ARGUMENT: i32 = 1;
pub const DK_ERR_CUDA: i32 = 2;
pub const DK_ERR_NO_DEVICE: i32 = 3;
pub const DK_ERR_NCCL: i32 = 4;
pub const DK_ERR_UNSUPPORTED: i32 = 5;
pub const DK_DOKO: i32 = 0;
pub const DK_FDO: i32 = 1;
pub const DK_PHASE_RESERVATION: i32 = 0;
pub const DK_PHASE_ANNOUNCEMENT: i32 = 1;
pub const DK_PHASE_PLAY_CARD: i32 = 2;
pub const DK_PHASE_FINISHED: i32 = 3;
pub const DK_GT_NORMAL: i32 = 0;
pub const DK_GT_WEDDING: i32 = 1;
pub const DK_GT_DIAMONDS_SOLO: i32 = 2;
pub const DK_GT_HEARTS_SOLO: i32 = 3;
pub const DK_GT_SPADES_SOLO: i32 = 4;
pub const DK_GT_CLUBS_SOLO: i32 = 5;
pub const DK_GT_TRUMPLESS_SOLO: i32 = 6;
pub const DK_GT_QUEENS_SOLO: i32 = 7;
pub const DK_GT_JACKS_SOLO: i32 = 8;
pub const DK_GT_NONE: i32 = 0xf;
pub const DK_RES_HEALTHY: i32 = 0;
pub const DK_RES_WEDDING: i32 = 1;
pub const DK_RES_DIAMONDS_SOLO: i32 = 2;
pub const DK_RES_HEARTS_SOLO: i32 = 3;
pub const DK_RES_SPADES_SOLO: i32 = 4;
pub const DK_RES_CLUBS_SOLO: i32 = 5;
pub const DK_RES_QUEENS_SOLO: i32 = 6;
pub const DK_RES_JACKS_SOLO: i32 = 7;
pub const DK_RES_TRUMPLESS_SOLO: i32 = 8;
pub const DK_RES_NONE: i32 = 0xff;
pub const DK_ANN_NONE: i32 = 0;
pub const DK_ANN_RE_CONTRA: i32 = 1;
pub const DK_ANN_NO90: i32 = 2;
pub const DK_ANN_NO60: i32 = 3;
pub const DK_ANN_NO30: i32 = 4;
pub const DK_ANN_BLACK: i32 = 5;
pub const DK_ANN_COUNTER: i32 = 6;
pub const DK_TEAM_IN_RESERVATIONS: i32 = 0;
pub const DK_TEAM_WEDDING_UNSOLVED: i32 = 1;
pub const DK_TEAM_WEDDING_SOLVED: i32 = 2;
pub const DK_TEAM_NO_WEDDING: i32 = 3;
pub const DK_LAYOUT_DO110: i32 = 0;
pub const DK_LAYOUT_DO114: i32 = 1;
pub const DK_LAYOUT_FDO_PI311: i32 = 2;

extern "C" {
    pub fn dk_init(device: c_int, out: *mut *mut dk_ctx) -> dk_status;
    pub fn dk_destroy(ctx: *mut dk_ctx) -> dk_status;
    pub fn dk_last_error(ctx: *const dk_ctx) -> *const c_char;
    pub fn dk_version() -> *const c_char;
    pub fn dk_device_info(ctx: *const dk_ctx, sm_count: *mut c_int, cc_major: *mut c_int, cc_minor: *mut c_int, total_mem: *mut usize) -> dk_status;
    pub fn dk_synchronize(ctx: *mut dk_ctx, stream: dk_stream) -> dk_status;
    pub fn dk_launch_count(ctx: *const dk_ctx) -> u64;
    pub fn dk_new_games(ctx: *mut dk_ctx, engine: c_int, n: usize, rng: *const dk_rng, out: *mut dk_state, stream: dk_stream) -> dk_status;
    pub fn dk_from_deals(ctx: *mut dk_ctx, engine: c_int, n: usize, hands: *const u64, start: *const u8, out: *mut dk_state, stream: dk_stream) -> dk_status;
    pub fn dk_legal_mask(ctx: *mut dk_ctx, engine: c_int, n: usize, states: *const dk_state, mask_out: *mut u64, stream: dk_stream) -> dk_status;
    pub fn dk_legal_mask_az(ctx: *mut dk_ctx, n: usize, states: *const dk_state, is_secondary: c_int, az_epoch: u64, mask_out: *mut u64, n_allowed_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_state_id(ctx: *mut dk_ctx, n: usize, states: *const dk_state, last_action: *const u8, id_out: *mut u64, stream: dk_stream) -> dk_status;
    pub fn dk_random_action(ctx: *mut dk_ctx, engine: c_int, n: usize, states: *const dk_state, rng: *const dk_rng, flags: u32, action_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_apply(ctx: *mut dk_ctx, engine: c_int, n: usize, states: *mut dk_state, action_idx: *const u8, flags: u32, err_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_terminal(ctx: *mut dk_ctx, engine: c_int, n: usize, states: *const dk_state, done_out: *mut u8, points_out: *mut i32, stream: dk_stream) -> dk_status;
    pub fn dk_encode(ctx: *mut dk_ctx, layout: c_int, n: usize, states: *const dk_state, out: *mut i64, row_stride: usize, stream: dk_stream) -> dk_status;
    pub fn dk_step_random_encode(ctx: *mut dk_ctx, n: usize, states: *mut dk_state, rng: *const dk_rng, flags: u32, obs_out: *mut i64, row_stride: usize, action_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_encode_narrow(ctx: *mut dk_ctx, layout: c_int, elem_bytes: c_int, n: usize, states: *const dk_state, out: *mut c_void, stream: dk_stream) -> dk_status;
    pub fn dk_step_random_encode_narrow(ctx: *mut dk_ctx, n: usize, states: *mut dk_state, rng: *const dk_rng, flags: u32, elem_bytes: c_int, obs_out: *mut c_void, action_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_playout(ctx: *mut dk_ctx, engine: c_int, flags: u32, n: usize, states: *const dk_state, rng: *const dk_rng, points_out: *mut i32, steps_out: *mut u32, stream: dk_stream) -> dk_status;
    pub fn dk_playout_trace(ctx: *mut dk_ctx, engine: c_int, n: usize, rng: *const dk_rng, points_out: *mut i32, trace_out: *mut u8, aux_out: *mut u32, stream: dk_stream) -> dk_status;
    pub fn dk_playout_host(ctx: *mut dk_ctx, engine: c_int, flags: u32, n: usize, states_host: *const dk_state, rng: *const dk_rng, points_out_host: *mut i32, steps_out_host: *mut u32) -> dk_status;
    pub fn dk_playout_host_compact(ctx: *mut dk_ctx, engine: c_int, flags: u32, n: usize, states_host: *const dk_state, rng: *const dk_rng, points_out_host: *mut i8, steps_out_host: *mut u8) -> dk_status;
    pub fn dk_playout_host_packed(ctx: *mut dk_ctx, engine: c_int, flags: u32, n: usize, states_host: *const dk_state, rng: *const dk_rng, points_packed_out_host: *mut u16, steps_out_host: *mut u8) -> dk_status;
    pub fn dk_playout_summary(ctx: *mut dk_ctx, engine: c_int, flags: u32, n: usize, states: *const dk_state, rng: *const dk_rng, stats: *mut dk_playout_stats, accumulate: c_int, stream: dk_stream) -> dk_status;
    pub fn dk_playout_summary_host(ctx: *mut dk_ctx, engine: c_int, flags: u32, n: usize, states_host: *const dk_state, rng: *const dk_rng, stats_out_host: *mut dk_playout_stats) -> dk_status;
    pub fn dk_determinize(ctx: *mut dk_ctx, engine: c_int, n_info: usize, samples_per_info: usize, states: *const dk_state, rng: *const dk_rng, hands_out: *mut u64, reservations_out: *mut u8, status_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_leaf_rollouts(ctx: *mut dk_ctx, n_leaves: usize, rollouts_per_leaf: usize, determinize: c_int, states: *const dk_state, rng: *const dk_rng, point_sum_out: *mut i64, stream: dk_stream) -> dk_status;
    pub fn dk_encode_ipi(ctx: *mut dk_ctx, n: usize, states: *const dk_state, assumed_hands: *const u64, assumed_reservations: *const u8, next_player: *const u8, out: *mut i64, row_stride: usize, err_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_pimc_evaluate(ctx: *mut dk_ctx, n_roots: usize, n_det: usize, n_rollouts: usize, states: *const dk_state, rng: *const dk_rng, visits_out: *mut u32, value_sum_out: *mut i64, status_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_fuse(ctx: *mut dk_ctx, strategy: c_int, n_roots: usize, n_rows: usize, visits: *const u32, status: *const u8, allowed: *const u64, action_out: *mut u8, n_success_out: *mut u32, stream: dk_stream) -> dk_status;
    pub fn dk_pimc_root_stats(ctx: *mut dk_ctx, n_roots: usize, n_rows: usize, visits: *const u32, status: *const u8, allowed: *const u64, stats: *mut i64, accumulate: c_int, stream: dk_stream) -> dk_status;
    pub fn dk_pimc_pick(ctx: *mut dk_ctx, strategy: c_int, n_roots: usize, stats: *const i64, allowed: *const u64, action_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_sp_create(ctx: *mut dk_ctx, max_games: usize, bufs: *const dk_sp_buffers, out: *mut *mut dk_selfplay) -> dk_status;
    pub fn dk_sp_destroy(sp: *mut dk_selfplay) -> dk_status;
    pub fn dk_sp_reset(sp: *mut dk_selfplay, stream: dk_stream) -> dk_status;
    pub fn dk_sp_begin_turn(sp: *mut dk_selfplay, n: usize, states: *const dk_state, az_epoch: u64, keep_prob: f32, flags: u32, rng: *const dk_rng, stream: dk_stream) -> dk_status;
    pub fn dk_sp_turn_view(sp: *mut dk_selfplay, allowed: *mut *const u64, flags: *mut *const u8, rows: *mut *const i64) -> dk_status;
    pub fn dk_sp_uniform_search(sp: *mut dk_selfplay, rng: *const dk_rng, policy_out: *mut f32, action_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_sp_end_turn(sp: *mut dk_selfplay, states: *mut dk_state, policy: *const f32, action: *const u8, err_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_sp_finalize(sp: *mut dk_selfplay, states: *const dk_state, stream: dk_stream) -> dk_status;
    pub fn dk_sp_counts(sp: *mut dk_selfplay, rows: *mut u64, dropped: *mut u64, unfinished: *mut u64, stream: dk_stream) -> dk_status;
    pub fn dk_pack_replay_records(ctx: *mut dk_ctx, n_rows: usize, states: *const i64, value: *const f32, policy: *const f32, out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_uct_workspace_bytes(n_trees: usize, iterations: usize) -> usize;
    pub fn dk_uct_search(ctx: *mut dk_ctx, n_roots: usize, trees_per_root: usize, determinize: c_int, iterations: usize, uct_exploration_constant: f32, states: *const dk_state, rng: *const dk_rng, workspace: *mut c_void, workspace_bytes: usize, visits_out: *mut u32, values_out: *mut f32, action_out: *mut u8, status_out: *mut u8, stream: dk_stream) -> dk_status;
    pub fn dk_comm_unique_id(ctx: *mut dk_ctx, out: *mut dk_nccl_id) -> dk_status;
    pub fn dk_comm_init(ctx: *mut dk_ctx, n_ranks: c_int, rank: c_int, id: *const dk_nccl_id) -> dk_status;
    pub fn dk_comm_destroy(ctx: *mut dk_ctx) -> dk_status;
    pub fn dk_allreduce_root_stats(ctx: *mut dk_ctx, n_values: usize, values: *mut i64, stream: dk_stream) -> dk_status;
}
