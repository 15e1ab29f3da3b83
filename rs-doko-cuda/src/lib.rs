//! rs-doko-cuda — the Rust host side of `libdoko_cuda.so`, the B200 batched Doppelkopf simulator (C ABI: `include/doko_cuda.h`).
//!
//! * [`ffi`]        every `DK_API` entry point, struct and constant — GENERATED from the header (`tools/gen_ffi.py`).
//! * [`DokoCuda`]   one context per (process, GPU); [`DeviceBuf`] caller-owned device memory; [`Batch`] a batch of games in device
//!                  memory with the reference's env vocabulary (one C-ABI call per method, whole batch at once).
//! * [`convert`]    `impl From<&FdoState> for dk_state` and back (feature `reference`).
//! * [`env`]        `FdoCudaState`: ONE game by value implementing `McEnvState` (rs-doko-mcts/src/env/env_state.rs:7-28) and
//!                  `AzEnvState` (rs-doko-alpha-zero/src/env/env_state.rs:4-42), so `rs-doko-mcts`, `rs-doko-alpha-zero` and
//!                  `rs-doko-evaluator` can drive the simulator unchanged (feature `reference`).
//! * [`philox_rng`] the parity stream of DESIGN.md as a `rand::RngCore`, for running the reference's own harness on the stream the
//!                  kernels use (feature `reference`).
//!
//! There is no CPU fallback: every method launches sm_100a kernels or returns a [`DokoError`].
#![allow(non_camel_case_types)]
pub mod ffi;
#[cfg(feature = "reference")]
pub mod convert;
#[cfg(feature = "reference")]
pub mod env;
#[cfg(feature = "reference")]
pub mod philox_rng;

use std::ffi::{c_int, c_void, CStr};
use std::marker::PhantomData;
use std::sync::Arc;

pub use ffi::{dk_playout_stats, dk_rng, dk_state};

#[derive(Debug, Clone)]
pub struct DokoError(pub i32, pub String);
impl std::fmt::Display for DokoError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "doko_cuda status {}: {}", self.0, self.1)
    }
}
impl std::error::Error for DokoError {}

// the four CUDA runtime calls the wrapper needs for caller-owned device memory
extern "C" {
    fn cudaMalloc(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    fn cudaFree(ptr: *mut c_void) -> c_int;
    fn cudaMemcpy(dst: *mut c_void, src: *const c_void, bytes: usize, kind: c_int) -> c_int;
    fn cudaSetDevice(device: c_int) -> c_int;
}
const MEMCPY_H2D: c_int = 1;
const MEMCPY_D2H: c_int = 2;
/// `cudaStreamLegacy`: the stream `cudaMemcpy` is ordered with; the wrapper launches on it so that copies and kernels need no extra
/// synchronisation (`NULL` would select the context's own non-blocking stream).
pub const STREAM_LEGACY: ffi::dk_stream = 1 as ffi::dk_stream;

struct CtxHandle(*mut ffi::dk_ctx, i32);
unsafe impl Send for CtxHandle {}
unsafe impl Sync for CtxHandle {}
impl Drop for CtxHandle {
    fn drop(&mut self) {
        unsafe { ffi::dk_destroy(self.0) };
    }
}

/// One simulator context (`dk_init`).  Cheap to clone (shared handle).  Keep one call in flight per context (see the header).
#[derive(Clone)]
pub struct DokoCuda {
    h: Arc<CtxHandle>,
}

impl DokoCuda {
    pub fn new(device: i32) -> Result<Self, DokoError> {
        let mut ctx = std::ptr::null_mut();
        let st = unsafe { ffi::dk_init(device, &mut ctx) };
        if st != ffi::DK_OK {
            return Err(DokoError(st, "dk_init failed: an sm_100 GPU is required (no CPU fallback)".into()));
        }
        Ok(DokoCuda { h: Arc::new(CtxHandle(ctx, device)) })
    }
    pub fn raw(&self) -> *mut ffi::dk_ctx {
        self.h.0
    }
    pub fn device(&self) -> i32 {
        self.h.1
    }
    pub fn check(&self, st: i32) -> Result<(), DokoError> {
        if st == ffi::DK_OK {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(ffi::dk_last_error(self.h.0)) }.to_string_lossy().into_owned();
        Err(DokoError(st, msg))
    }
    pub fn synchronize(&self) -> Result<(), DokoError> {
        self.check(unsafe { ffi::dk_synchronize(self.h.0, STREAM_LEGACY) })
    }
    pub fn rng(seed: u64, first_id: u64, epoch: u32) -> dk_rng {
        dk_rng { seed, first_id, epoch, first_sub: 0 }
    }

    /// `n` fresh full-rules games played to the end by the random policy; per-game points (`dk_playout_host_packed`: 2 bytes per game
    /// cross PCIe) decoded to `[i32; 4]`.  The batched `McEnvState::random_rollout` / `FdoState::random_action_for_current_player` loop.
    pub fn random_playouts(&self, n: usize, rng: &dk_rng, with_announcements: bool) -> Result<Vec<[i32; 4]>, DokoError> {
        let mut packed = vec![0u16; n];
        let flags = if with_announcements { ffi::DK_PLAYOUT_WITH_ANNOUNCEMENTS } else { 0 };
        self.check(unsafe {
            ffi::dk_playout_host_packed(self.raw(), ffi::DK_FDO, flags, n, std::ptr::null(), rng, packed.as_mut_ptr(), std::ptr::null_mut())
        })?;
        Ok(packed.iter().map(|&v| unpack_points(v)).collect())
    }
    /// The same games, only their statistics (`dk_playout_summary_host`: 2160 bytes cross PCIe) — what an evaluator aggregates.
    pub fn random_playouts_summary(&self, n: usize, rng: &dk_rng, with_announcements: bool) -> Result<dk_playout_stats, DokoError> {
        let mut stats: dk_playout_stats = unsafe { std::mem::zeroed() };
        let flags = if with_announcements { ffi::DK_PLAYOUT_WITH_ANNOUNCEMENTS } else { 0 };
        self.check(unsafe { ffi::dk_playout_summary_host(self.raw(), ffi::DK_FDO, flags, n, std::ptr::null(), rng, &mut stats) })?;
        Ok(stats)
    }
}

/// `dk_unpack_points` of the header: the four points from the packed 16-bit form.
pub fn unpack_points(v: u16) -> [i32; 4] {
    let a = (v & 0xFF) as u8 as i8 as i32;
    let same = (v >> 8) & 7;
    let k = 1 + (same & 1) as i32 + ((same >> 1) & 1) as i32 + ((same >> 2) & 1) as i32;
    let b = if k == 4 { a } else { -(k * a) / (4 - k) };
    [a, if same & 1 != 0 { a } else { b }, if same & 2 != 0 { a } else { b }, if same & 4 != 0 { a } else { b }]
}

/// Caller-owned device memory (`cudaMalloc`), 256-byte aligned as the C ABI's vector accesses need.
pub struct DeviceBuf<T: Copy> {
    ptr: *mut c_void,
    len: usize,
    device: i32,
    _t: PhantomData<T>,
}
unsafe impl<T: Copy> Send for DeviceBuf<T> {}
impl<T: Copy> DeviceBuf<T> {
    pub fn new(dk: &DokoCuda, len: usize) -> Result<Self, DokoError> {
        let mut ptr = std::ptr::null_mut();
        unsafe {
            if cudaSetDevice(dk.device()) != 0 || (len > 0 && cudaMalloc(&mut ptr, len * std::mem::size_of::<T>()) != 0) {
                return Err(DokoError(ffi::DK_ERR_CUDA, "cudaMalloc failed".into()));
            }
        }
        Ok(DeviceBuf { ptr, len, device: dk.device(), _t: PhantomData })
    }
    pub fn from_host(dk: &DokoCuda, host: &[T]) -> Result<Self, DokoError> {
        let b = Self::new(dk, host.len())?;
        b.upload(host)?;
        Ok(b)
    }
    pub fn upload(&self, host: &[T]) -> Result<(), DokoError> {
        assert!(host.len() <= self.len);
        let rc = unsafe { cudaMemcpy(self.ptr, host.as_ptr() as *const c_void, host.len() * std::mem::size_of::<T>(), MEMCPY_H2D) };
        if rc != 0 { Err(DokoError(ffi::DK_ERR_CUDA, "cudaMemcpy H2D failed".into())) } else { Ok(()) }
    }
    pub fn to_host(&self) -> Result<Vec<T>, DokoError> {
        let mut v: Vec<T> = Vec::with_capacity(self.len);
        let rc = unsafe { cudaMemcpy(v.as_mut_ptr() as *mut c_void, self.ptr, self.len * std::mem::size_of::<T>(), MEMCPY_D2H) };
        if rc != 0 {
            return Err(DokoError(ffi::DK_ERR_CUDA, "cudaMemcpy D2H failed".into()));
        }
        unsafe { v.set_len(self.len) };
        Ok(v)
    }
    pub fn as_ptr(&self) -> *const T {
        self.ptr as *const T
    }
    pub fn as_mut_ptr(&self) -> *mut T {
        self.ptr as *mut T
    }
    pub fn len(&self) -> usize {
        self.len
    }
    pub fn is_empty(&self) -> bool {
        self.len == 0
    }
}
impl<T: Copy> Drop for DeviceBuf<T> {
    fn drop(&mut self) {
        unsafe {
            cudaSetDevice(self.device);
            cudaFree(self.ptr);
        }
    }
}

/// A batch of full-rules games in device memory: the batched counterpart of `McFullDokoEnvState` / `FdoAzEnvState`.  Every method is
/// one C-ABI call over the whole batch; results stay on the device unless a `*_host` variant is used.
pub struct Batch {
    pub dk: DokoCuda,
    pub states: DeviceBuf<dk_state>,
}

impl Batch {
    /// `FdoState::new_game` for `n` games: game `i` is dealt from the stream (`rng.seed`, `rng.first_id + i`, `rng.epoch`).
    pub fn new_games(dk: &DokoCuda, n: usize, rng: &dk_rng) -> Result<Self, DokoError> {
        let states = DeviceBuf::new(dk, n)?;
        dk.check(unsafe { ffi::dk_new_games(dk.raw(), ffi::DK_FDO, n, rng, states.as_mut_ptr(), STREAM_LEGACY) })?;
        Ok(Batch { dk: dk.clone(), states })
    }
    /// A batch from host records (e.g. converted `FdoState`s, see [`convert`]).
    pub fn from_records(dk: &DokoCuda, records: &[dk_state]) -> Result<Self, DokoError> {
        Ok(Batch { dk: dk.clone(), states: DeviceBuf::from_host(dk, records)? })
    }
    pub fn len(&self) -> usize {
        self.states.len()
    }
    pub fn is_empty(&self) -> bool {
        self.states.is_empty()
    }
    pub fn records(&self) -> Result<Vec<dk_state>, DokoError> {
        self.states.to_host()
    }
    /// `AzEnvState::allowed_actions_by_action_index(is_secondary, epoch)` as bit masks + `number_of_allowed_actions(epoch)`.
    pub fn allowed_actions(&self, is_secondary: bool, epoch: u64) -> Result<(Vec<u64>, Vec<u8>), DokoError> {
        let (m, c) = (DeviceBuf::<u64>::new(&self.dk, self.len())?, DeviceBuf::<u8>::new(&self.dk, self.len())?);
        self.dk.check(unsafe {
            ffi::dk_legal_mask_az(self.dk.raw(), self.len(), self.states.as_ptr(), is_secondary as c_int, epoch, m.as_mut_ptr(), c.as_mut_ptr(), STREAM_LEGACY)
        })?;
        Ok((m.to_host()?, c.to_host()?))
    }
    /// `take_action_by_action_index(action[i], skip_single, _)` for every game, in place; `Err` flags per game (1 = the reference would panic).
    pub fn take_actions(&mut self, actions: &[u8], skip_single: bool) -> Result<Vec<u8>, DokoError> {
        assert_eq!(actions.len(), self.len());
        let (a, e) = (DeviceBuf::from_host(&self.dk, actions)?, DeviceBuf::<u8>::new(&self.dk, self.len())?);
        let flags = if skip_single { ffi::DK_APPLY_SKIP_SINGLE } else { 0 };
        self.dk.check(unsafe { ffi::dk_apply(self.dk.raw(), ffi::DK_FDO, self.len(), self.states.as_mut_ptr(), a.as_ptr(), flags, e.as_mut_ptr(), STREAM_LEGACY) })?;
        e.to_host()
    }
    /// `is_terminal` + `rewards_or_none` (player_points; the MCTS env reads them as f64, the AlphaZero env divides by 8).
    pub fn terminal(&self) -> Result<(Vec<u8>, Vec<[i32; 4]>), DokoError> {
        let (d, p) = (DeviceBuf::<u8>::new(&self.dk, self.len())?, DeviceBuf::<[i32; 4]>::new(&self.dk, self.len())?);
        self.dk.check(unsafe {
            ffi::dk_terminal(self.dk.raw(), ffi::DK_FDO, self.len(), self.states.as_ptr(), d.as_mut_ptr(), p.as_mut_ptr() as *mut i32, STREAM_LEGACY)
        })?;
        Ok((d.to_host()?, p.to_host()?))
    }
    /// `encode_into_memory` for every game into a caller-owned device buffer of `len() * 311` i64 (where the network batcher reads it).
    pub fn encode_into(&self, out: &DeviceBuf<i64>) -> Result<(), DokoError> {
        assert!(out.len() >= self.len() * ffi::DK_OBS_LEN_FDO_PI311 as usize);
        self.dk.check(unsafe {
            ffi::dk_encode(self.dk.raw(), ffi::DK_LAYOUT_FDO_PI311, self.len(), self.states.as_ptr(), out.as_mut_ptr(), ffi::DK_OBS_LEN_FDO_PI311 as usize, STREAM_LEGACY)
        })
    }
    /// The same rows as int32 — what the Python side narrows `encode_into_memory`'s `Vec<i64>` to before the network sees it
    /// (`az_doko.py:369`) — into a dense device buffer of `len() * 311` i32 (`dk_encode_narrow`: half the bytes of an HBM-bound kernel).
    pub fn encode_into_i32(&self, out: &DeviceBuf<i32>) -> Result<(), DokoError> {
        assert!(out.len() >= self.len() * ffi::DK_OBS_LEN_FDO_PI311 as usize);
        self.dk.check(unsafe {
            ffi::dk_encode_narrow(self.dk.raw(), ffi::DK_LAYOUT_FDO_PI311, 4, self.len(), self.states.as_ptr(), out.as_mut_ptr() as *mut std::ffi::c_void, STREAM_LEGACY)
        })
    }
    /// `random_rollout` of every game (`_no_announcement` policy unless `with_announcements`): per-game points.
    pub fn random_rollouts(&self, rng: &dk_rng, with_announcements: bool) -> Result<Vec<[i32; 4]>, DokoError> {
        let p = DeviceBuf::<[i32; 4]>::new(&self.dk, self.len())?;
        let flags = if with_announcements { ffi::DK_PLAYOUT_WITH_ANNOUNCEMENTS } else { 0 };
        self.dk.check(unsafe {
            ffi::dk_playout(self.dk.raw(), ffi::DK_FDO, flags, self.len(), self.states.as_ptr(), rng, p.as_mut_ptr() as *mut i32, std::ptr::null_mut(), STREAM_LEGACY)
        })?;
        p.to_host()
    }
    /// `CAPSampling::sample` (card_matching): `samples` hidden-hand samples per game: (hands `[u64;4]`, reservations `[u8;4]`, status).
    pub fn determinize(&self, rng: &dk_rng, samples: usize) -> Result<(Vec<[u64; 4]>, Vec<[u8; 4]>, Vec<u8>), DokoError> {
        let n = self.len() * samples;
        let (h, r, s) = (DeviceBuf::<[u64; 4]>::new(&self.dk, n)?, DeviceBuf::<[u8; 4]>::new(&self.dk, n)?, DeviceBuf::<u8>::new(&self.dk, n)?);
        self.dk.check(unsafe {
            ffi::dk_determinize(self.dk.raw(), ffi::DK_FDO, self.len(), samples, self.states.as_ptr(), rng, h.as_mut_ptr() as *mut u64, r.as_mut_ptr() as *mut u8,
                                s.as_mut_ptr(), STREAM_LEGACY)
        })?;
        Ok((h.to_host()?, r.to_host()?, s.to_host()?))
    }
    /// Leaf-parallel rollouts: exact integer point sums over `rollouts` rollouts per game (each first determinized when `determinize`).
    pub fn leaf_rollouts(&self, rng: &dk_rng, rollouts: usize, determinize: bool) -> Result<Vec<[i64; 4]>, DokoError> {
        let out = DeviceBuf::<[i64; 4]>::new(&self.dk, self.len())?;
        self.dk.check(unsafe {
            ffi::dk_leaf_rollouts(self.dk.raw(), self.len(), rollouts, determinize as c_int, self.states.as_ptr(), rng, out.as_mut_ptr() as *mut i64, STREAM_LEGACY)
        })?;
        out.to_host()
    }
    /// `AzEnvState::id()` of every game.
    pub fn ids(&self, last_actions: Option<&[u8]>) -> Result<Vec<u64>, DokoError> {
        let out = DeviceBuf::<u64>::new(&self.dk, self.len())?;
        let la = match last_actions { Some(a) => Some(DeviceBuf::from_host(&self.dk, a)?), None => None };
        let la_ptr = la.as_ref().map(|b| b.as_ptr()).unwrap_or(std::ptr::null());
        self.dk.check(unsafe { ffi::dk_state_id(self.dk.raw(), self.len(), self.states.as_ptr(), la_ptr, out.as_mut_ptr(), STREAM_LEGACY) })?;
        out.to_host()
    }
}
