//! Thin FFI over `include/doko_cuda.h` plus a batch wrapper with the reference's env vocabulary.
//! Untested here (no Rust toolchain in the build image); the C ABI it binds is exercised by the Python/ctypes tests.
#![allow(non_camel_case_types)]
use std::ffi::{c_char, c_int, c_void, CStr};

#[repr(C)]
pub struct dk_ctx {
    _private: [u8; 0],
}

/// include/doko_cuda.h: dk_state (128 bytes)
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

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dk_rng {
    pub seed: u64,
    pub first_id: u64,
    pub epoch: u32,
    pub first_sub: u32,
}

pub const DK_DOKO: c_int = 0;
pub const DK_FDO: c_int = 1;
pub const DK_PLAYOUT_WITH_ANNOUNCEMENTS: u32 = 1;
pub const DK_APPLY_SKIP_SINGLE: u32 = 1;
pub const DK_LAYOUT_FDO_PI311: c_int = 2;

extern "C" {
    pub fn dk_init(device: c_int, out: *mut *mut dk_ctx) -> i32;
    pub fn dk_destroy(ctx: *mut dk_ctx) -> i32;
    pub fn dk_last_error(ctx: *const dk_ctx) -> *const c_char;
    pub fn dk_new_games(ctx: *mut dk_ctx, engine: c_int, n: usize, rng: *const dk_rng, out: *mut dk_state, stream: *mut c_void) -> i32;
    pub fn dk_legal_mask(ctx: *mut dk_ctx, engine: c_int, n: usize, states: *const dk_state, mask_out: *mut u64, stream: *mut c_void) -> i32;
    pub fn dk_apply(ctx: *mut dk_ctx, engine: c_int, n: usize, states: *mut dk_state, action_idx: *const u8, flags: u32, err_out: *mut u8, stream: *mut c_void) -> i32;
    pub fn dk_terminal(ctx: *mut dk_ctx, engine: c_int, n: usize, states: *const dk_state, done_out: *mut u8, points_out: *mut i32, stream: *mut c_void) -> i32;
    pub fn dk_encode(ctx: *mut dk_ctx, layout: c_int, n: usize, states: *const dk_state, out: *mut i64, row_stride: usize, stream: *mut c_void) -> i32;
    pub fn dk_playout_host(ctx: *mut dk_ctx, engine: c_int, flags: u32, n: usize, states_host: *const dk_state, rng: *const dk_rng,
                           points_out_host: *mut i32, steps_out_host: *mut u32) -> i32;
    pub fn dk_determinize(ctx: *mut dk_ctx, engine: c_int, n_info: usize, samples_per_info: usize, states: *const dk_state, rng: *const dk_rng,
                          hands_out: *mut u64, reservations_out: *mut u8, status_out: *mut u8, stream: *mut c_void) -> i32;
    pub fn dk_leaf_rollouts(ctx: *mut dk_ctx, n_leaves: usize, rollouts_per_leaf: usize, determinize: c_int, states: *const dk_state,
                            rng: *const dk_rng, point_sum_out: *mut i64, stream: *mut c_void) -> i32;
    // PIMC move decision (DefaultImpiPolicy::execute, PolicyFusionFn::fuse)
    pub fn dk_pimc_evaluate(ctx: *mut dk_ctx, n_roots: usize, n_det: usize, n_rollouts: usize, states: *const dk_state, rng: *const dk_rng,
                            visits_out: *mut u32, value_sum_out: *mut i64, status_out: *mut u8, stream: *mut c_void) -> i32;
    pub fn dk_fuse(ctx: *mut dk_ctx, strategy: c_int, n_roots: usize, n_rows: usize, visits: *const u32, status: *const u8, allowed: *const u64,
                   action_out: *mut u8, n_success_out: *mut u32, stream: *mut c_void) -> i32;
    pub fn dk_pimc_root_stats(ctx: *mut dk_ctx, n_roots: usize, n_rows: usize, visits: *const u32, status: *const u8, allowed: *const u64,
                              stats: *mut i64, accumulate: c_int, stream: *mut c_void) -> i32;
    pub fn dk_pimc_pick(ctx: *mut dk_ctx, strategy: c_int, n_roots: usize, stats: *const i64, allowed: *const u64, action_out: *mut u8,
                        stream: *mut c_void) -> i32;
    // CachedMCTS::monte_carlo_tree_search per (root, sample)
    pub fn dk_uct_workspace_bytes(n_trees: usize, iterations: usize) -> usize;
    pub fn dk_uct_search(ctx: *mut dk_ctx, n_roots: usize, trees_per_root: usize, determinize: c_int, iterations: usize, uct_exploration_constant: f32,
                         states: *const dk_state, rng: *const dk_rng, workspace: *mut c_void, workspace_bytes: usize, visits_out: *mut u32,
                         values_out: *mut f32, action_out: *mut u8, status_out: *mut u8, stream: *mut c_void) -> i32;
    // encode_state_ipi, replay records
    pub fn dk_encode_ipi(ctx: *mut dk_ctx, n: usize, states: *const dk_state, assumed_hands: *const u64, assumed_reservations: *const u8,
                         next_player: *const u8, out: *mut i64, row_stride: usize, err_out: *mut u8, stream: *mut c_void) -> i32;
    pub fn dk_pack_replay_records(ctx: *mut dk_ctx, n_rows: usize, states: *const i64, value: *const f32, policy: *const f32, out: *mut u8,
                                  stream: *mut c_void) -> i32;
    // lock-step self_play driver
    pub fn dk_sp_create(ctx: *mut dk_ctx, max_games: usize, bufs: *const dk_sp_buffers, out: *mut *mut dk_selfplay) -> i32;
    pub fn dk_sp_destroy(sp: *mut dk_selfplay) -> i32;
    pub fn dk_sp_reset(sp: *mut dk_selfplay, stream: *mut c_void) -> i32;
    pub fn dk_sp_begin_turn(sp: *mut dk_selfplay, n: usize, states: *const dk_state, az_epoch: u64, keep_prob: f32, flags: u32, rng: *const dk_rng,
                            stream: *mut c_void) -> i32;
    pub fn dk_sp_turn_view(sp: *mut dk_selfplay, allowed: *mut *const u64, flags: *mut *const u8, rows: *mut *const i64) -> i32;
    pub fn dk_sp_end_turn(sp: *mut dk_selfplay, states: *mut dk_state, policy: *const f32, action: *const u8, err_out: *mut u8, stream: *mut c_void) -> i32;
    pub fn dk_sp_finalize(sp: *mut dk_selfplay, states: *const dk_state, stream: *mut c_void) -> i32;
    pub fn dk_sp_counts(sp: *mut dk_selfplay, rows: *mut u64, dropped: *mut u64, unfinished: *mut u64, stream: *mut c_void) -> i32;
}

#[repr(C)]
pub struct dk_selfplay {
    _private: [u8; 0],
}
/// include/doko_cuda.h: dk_sp_buffers — caller-owned device memory standing in for states_buffer / policy_targets_buffer / value_targets_buffer
#[repr(C)]
pub struct dk_sp_buffers {
    pub states: *mut i64,
    pub policy: *mut f32,
    pub value: *mut f32,
    pub player: *mut u8,
    pub game: *mut u32,
    pub capacity: usize,
}
pub const DK_FUSE_MAX_N: c_int = 0;
pub const DK_FUSE_AVERAGE: c_int = 1;
pub const DK_ACTION_NONE: u8 = 0xFF;

pub struct DokoCuda {
    ctx: *mut dk_ctx,
}
unsafe impl Send for DokoCuda {}

#[derive(Debug)]
pub struct DokoError(pub i32, pub String);

impl DokoCuda {
    pub fn new(device: i32) -> Result<Self, DokoError> {
        let mut ctx = std::ptr::null_mut();
        let st = unsafe { dk_init(device, &mut ctx) };
        if st != 0 {
            return Err(DokoError(st, "dk_init failed: an sm_100 GPU is required (no CPU fallback)".into()));
        }
        Ok(DokoCuda { ctx })
    }
    fn check(&self, st: i32) -> Result<(), DokoError> {
        if st == 0 {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(dk_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(DokoError(st, msg))
    }
    /// The batched replacement of `McFullDokoEnvState::random_rollout` for `n` fresh games
    /// (rs-doko-mcts/src/env/envs/env_state_full_doko.rs:198-220): rewards = player_points as f64.
    pub fn random_playouts(&self, n: usize, seed: u64, first_id: u64, with_announcements: bool) -> Result<Vec<[f64; 4]>, DokoError> {
        let rng = dk_rng { seed, first_id, epoch: 0, first_sub: 0 };
        let mut pts = vec![0i32; n * 4];
        let flags = if with_announcements { DK_PLAYOUT_WITH_ANNOUNCEMENTS } else { 0 };
        self.check(unsafe { dk_playout_host(self.ctx, DK_FDO, flags, n, std::ptr::null(), &rng, pts.as_mut_ptr(), std::ptr::null_mut()) })?;
        Ok(pts.chunks_exact(4).map(|p| [p[0] as f64, p[1] as f64, p[2] as f64, p[3] as f64]).collect())
    }
}

impl Drop for DokoCuda {
    fn drop(&mut self) {
        unsafe { dk_destroy(self.ctx) };
    }
}
