//! ONE game by value over the CUDA simulator, implementing the reference's env traits — the drop-in shape:
//!
//! * `McEnvState<FdoAction, 4, 39>` (rs-doko-mcts/src/env/env_state.rs:7-28), in place of `McFullDokoEnvState`
//!   (rs-doko-mcts/src/env/envs/env_state_full_doko.rs:62-220);
//! * `AzEnvState<FdoAction, 4, 39>` (rs-doko-alpha-zero/src/env/env_state.rs:4-42), in place of `FdoAzEnvState`
//!   (rs-doko-alpha-zero/src/env/envs/full_doko/full_doko.rs:43-172).
//!
//! A state is the 128-byte record on the host plus the last action; every trait method is one C-ABI call on a one-record batch
//! (record up, result down: latency bound, about 10 µs).  That makes `MCTS::monte_carlo_tree_search`, `self_play` and the evaluator
//! run unchanged on the simulator's rules; code that wants the simulator's throughput holds its games in a [`crate::Batch`] and calls
//! the same operations once per lock-step (INTEGRATION.md shows both).
//!
//! Errors: the reference panics on an illegal action / on `rewards` of an unfinished game; so do these adaptors (the C ABI reports the
//! condition, the adaptor turns it into the panic the callers expect).
use std::fmt::{Display, Formatter};
use std::hash::{Hash, Hasher};

use rand::prelude::SmallRng;
use rand::RngCore;
use rs_doko_alpha_zero::env::env_state::AzEnvState;
use rs_doko_mcts::env::env_state::McEnvState;
use rs_full_doko::action::action::FdoAction;
use rs_full_doko::action::allowed_actions::FdoAllowedActions;
use rs_full_doko::display::display::display_game;
use rs_full_doko::state::state::FdoState;
use rs_game_utils::bit_flag::Bitflag;

use crate::ffi::*;
use crate::{Batch, DeviceBuf, DokoCuda, STREAM_LEGACY};

const CALLS: u64 = 0x1F << 33; // AnnouncementReContra .. AnnouncementBlack
const SOLOS_AND_WEDDING: u64 = 0xFF << 25;

#[derive(Clone)]
pub struct FdoCudaState {
    pub dk: DokoCuda,
    pub record: dk_state,
    pub last_played_action: Option<FdoAction>,
}

impl FdoCudaState {
    pub fn new(dk: &DokoCuda, record: dk_state, last_played_action: Option<FdoAction>) -> Self {
        FdoCudaState { dk: dk.clone(), record, last_played_action }
    }
    /// From a reference state (e.g. `FdoState::new_game(&mut rng)`).
    pub fn from_state(dk: &DokoCuda, state: &FdoState, last_played_action: Option<FdoAction>) -> Self {
        Self::new(dk, dk_state::from(state), last_played_action)
    }
    pub fn to_state(&self) -> FdoState {
        FdoState::from(&self.record)
    }
    fn bytes(&self) -> [u8; 128] {
        unsafe { std::mem::transmute::<dk_state, [u8; 128]>(self.record) }
    }
    fn one(&self) -> Batch {
        Batch::from_records(&self.dk, std::slice::from_ref(&self.record)).expect("device memory for one record")
    }
    fn phase(&self) -> i32 {
        (self.record.meta & 3) as i32
    }
    /// legal mask with the epoch / is_secondary filter of FdoAzEnvState (dk_legal_mask_az)
    fn legal(&self, is_secondary: bool, epoch: u64) -> u64 {
        self.one().allowed_actions(is_secondary, epoch).expect("dk_legal_mask_az").0[0]
    }
    /// McFullDokoEnvState's filter below the root (env_state_full_doko.rs:132-172)
    fn mc_allowed(&self, first_expansion: bool) -> u64 {
        let mut m = self.legal(false, u64::MAX);
        if !first_expansion {
            let solo_declared = (0..self.record.n_reservations as usize).any(|k| self.record.reservations[k] as i32 >= DK_RES_DIAMONDS_SOLO);
            if solo_declared {
                m &= !SOLOS_AND_WEDDING;
            }
            m &= !CALLS;
        }
        m
    }
    fn apply(&self, action: usize, skip_single: bool) -> Self {
        let mut b = self.one();
        let err = b.take_actions(&[action as u8], skip_single).expect("dk_apply");
        assert!(err[0] == 0, "action {} is not allowed in this state", action); // the reference panics (state.rs:209-218)
        FdoCudaState { dk: self.dk.clone(), record: b.records().expect("records")[0], last_played_action: Some(FdoAction::from_index(action)) }
    }
    fn points(&self) -> Option<[i32; 4]> {
        if self.phase() != DK_PHASE_FINISHED {
            return None;
        }
        let p = self.record.points;
        Some([p[0] as i32, p[1] as i32, p[2] as i32, p[3] as i32])
    }
}

impl PartialEq for FdoCudaState {
    fn eq(&self, o: &Self) -> bool {
        self.bytes() == o.bytes() && self.last_played_action == o.last_played_action
    }
}
impl Eq for FdoCudaState {}
impl Hash for FdoCudaState {
    fn hash<H: Hasher>(&self, h: &mut H) {
        self.bytes().hash(h);
        self.last_played_action.map(|a| a.to_index()).hash(h);
    }
}
impl Display for FdoCudaState {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result {
        write!(f, "{:?}", self.last_played_action)
    }
}

impl McEnvState<FdoAction, 4, 39> for FdoCudaState {
    fn current_player(&self) -> usize {
        ((self.record.meta >> 2) & 3) as usize // BOTTOM (0) when the game is over, as in the reference
    }
    fn is_terminal(&self) -> bool {
        self.phase() == DK_PHASE_FINISHED
    }
    fn last_action(&self) -> Option<FdoAction> {
        self.last_played_action
    }
    fn possible_states(&self, first_expansion: bool) -> heapless::Vec<Self, 39> {
        // one batched call for all children: the record replicated, one action each
        let mask = self.mc_allowed(first_expansion);
        let actions: Vec<u8> = (0..39u8).filter(|a| mask >> a & 1 != 0).collect();
        let mut out = heapless::Vec::new();
        if actions.is_empty() {
            return out;
        }
        let mut b = Batch::from_records(&self.dk, &vec![self.record; actions.len()]).expect("device memory");
        let err = b.take_actions(&actions, false).expect("dk_apply");
        debug_assert!(err.iter().all(|e| *e == 0));
        for (rec, a) in b.records().expect("records").into_iter().zip(actions) {
            let _ = out.push(FdoCudaState { dk: self.dk.clone(), record: rec, last_played_action: Some(FdoAction::from_index(a as usize)) });
        }
        out
    }
    fn allowed_actions(&self, first_expansion: bool) -> FdoAllowedActions {
        FdoAllowedActions(Bitflag(self.mc_allowed(first_expansion)))
    }
    fn by_action(&self, action: FdoAction) -> Self {
        self.apply(action.to_index(), false)
    }
    fn rewards_or_none(&self) -> Option<[f64; 4]> {
        self.points().map(|p| [p[0] as f64, p[1] as f64, p[2] as f64, p[3] as f64])
    }
    /// `_no_announcement` rollout to the end of the game.  The caller's SmallRng supplies the 64-bit unit id of the Philox stream the
    /// kernel draws from (one `next_u64` per rollout), so rollouts stay reproducible from the caller's seed.
    fn random_rollout(&self, rng: &mut SmallRng) -> [f64; 4] {
        let stream = dk_rng { seed: 0xD0C0_5EED, first_id: rng.next_u64(), epoch: 0, first_sub: 0 };
        let mut pts = [0i32; 4];
        self.dk
            .check(unsafe { dk_playout_host(self.dk.raw(), DK_FDO, 0, 1, &self.record, &stream, pts.as_mut_ptr(), std::ptr::null_mut()) })
            .expect("dk_playout_host");
        [pts[0] as f64, pts[1] as f64, pts[2] as f64, pts[3] as f64]
    }
}

impl AzEnvState<FdoAction, 4, 39> for FdoCudaState {
    const GAME_NAME: &'static str = "full_doko";

    fn current_player(&self) -> usize {
        ((self.record.meta >> 2) & 3) as usize
    }
    fn is_terminal(&self) -> bool {
        self.phase() == DK_PHASE_FINISHED
    }
    fn rewards_or_none(&self) -> Option<[f32; 4]> {
        self.points().map(|p| [p[0] as f32 / 8f32, p[1] as f32 / 8f32, p[2] as f32 / 8f32, p[3] as f32 / 8f32])
    }
    fn encode_into_memory(&self, memory: &mut [i64]) {
        let out = DeviceBuf::<i64>::new(&self.dk, DK_OBS_LEN_FDO_PI311 as usize).expect("device memory");
        self.one().encode_into(&out).expect("dk_encode");
        memory.copy_from_slice(&out.to_host().expect("tokens"));
    }
    fn allowed_actions_by_action_index(&self, is_secondary: bool, epoch: usize) -> heapless::Vec<usize, 39> {
        let mask = self.legal(is_secondary, epoch as u64);
        (0..39usize).filter(|a| mask >> a & 1 != 0).collect()
    }
    fn number_of_allowed_actions(&self, epoch: usize) -> usize {
        self.one().allowed_actions(false, epoch as u64).expect("dk_legal_mask_az").1[0] as usize
    }
    fn take_action_by_action_index(&self, action: usize, skip_single: bool, _epoch: usize) -> Self {
        self.apply(action, skip_single)
    }
    fn id(&self) -> u64 {
        let la = self.last_played_action.map(|a| a.to_index() as u8).unwrap_or(DK_ACTION_NONE as u8);
        self.one().ids(Some(&[la])).expect("dk_state_id")[0]
    }
    fn last_action(&self) -> Option<FdoAction> {
        self.last_played_action
    }
    fn display_game(&self) -> String {
        display_game(self.to_state().observation_for_current_player())
    }
}

/// Keeps the legacy-stream constant referenced (the batch wrapper launches on it; see lib.rs).
#[allow(dead_code)]
const _STREAM: dk_stream = STREAM_LEGACY;
