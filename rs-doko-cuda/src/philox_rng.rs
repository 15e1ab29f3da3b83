//! The parity stream of the simulator (DESIGN.md "Philox parity contract") on the Rust side.
//!
//! The kernels draw every random decision from Philox4x32-10 with
//!     key = 64-bit seed,   counter = (unit_lo, unit_hi, site << 16 | block, epoch)
//! where a *site* is a class of reference call sites (deal, reservation pick, announcement decision, card pick, card_matching's rule 4,
//! hidden reservations, ...) and the k-th use of a site inside one unit takes word `k & 3` of block `k >> 2` (announcement decisions:
//! bit `k & 31` of word `k >> 5`), mapped onto `n` choices by `(word as u64 * n as u64) >> 32`.
//!
//! "The reference's harness fed the same Philox stream" therefore needs two things the stock `rand` plumbing cannot give:
//!  1. the draw must know its SITE — `Bitflag::random_single` (rs-game-utils/src/bit_flag.rs:86-94) and `FdoHandIter::choose`
//!     (rs-full-doko/src/matching/card_matching.rs:182-186) receive a bare `&mut SmallRng`;
//!  2. `rand 0.9`'s `random_range` maps a word with a widening multiply AND a rejection step that may draw a second word, and its
//!     `shuffle` batches several indices into one word (`IncreasingUniform`); the contract uses one word per decision, no rejection.
//! So the reference is run on the stream through [`SiteRng`] and the three-line patch of `rs-doko-cuda/PARITY_HARNESS.md`: the
//! harness sets the site from the state's phase before every `random_action_for_current_player` call, and the patched
//! `random_single` / `choose` / `randomly_distributed` ask the generator for `draw(n)` instead of `random_range(0..n)`.
//! With the patch applied, `cargo test -p rs-doko-cuda --features parity-harness` plays the reference's `FdoState` against
//! `dk_playout_trace`-style outputs of the library and compares game by game (what tests/test_gpu_parity_at_size.py does with the
//! C++ restatement in this repository, where no Rust toolchain exists).
use rand::RngCore;

pub const SITE_DEAL: u32 = 0;
pub const SITE_RESERVATION: u32 = 1;
pub const SITE_ANNOUNCEMENT: u32 = 2;
pub const SITE_CARD: u32 = 3;
pub const SITE_MATCH_CARD: u32 = 4;
pub const SITE_MATCH_RESERVATION: u32 = 5;
pub const SITE_ASSIGN: u32 = 6;
pub const SITE_STEP: u32 = 7;
pub const SITE_KEEP: u32 = 8;
pub const SITE_EXPAND: u32 = 9;
const N_SITES: usize = 10;

/// Philox4x32-10 (Salmon et al., Random123): one block of four words.
pub fn philox4x32_10(mut c: [u32; 4], mut k: [u32; 2]) -> [u32; 4] {
    const M0: u64 = 0xD251_1F53;
    const M1: u64 = 0xCD9E_8D57;
    for _ in 0..10 {
        let p0 = M0 * c[0] as u64;
        let p1 = M1 * c[2] as u64;
        c = [(p1 >> 32) as u32 ^ c[1] ^ k[0], p1 as u32, (p0 >> 32) as u32 ^ c[3] ^ k[1], p0 as u32];
        k = [k[0].wrapping_add(0x9E37_79B9), k[1].wrapping_add(0xBB67_AE85)];
    }
    c
}

/// The stream of ONE unit (game / info-state sample / rollout) with a per-site ordinal, as the kernels consume it.
#[derive(Clone, Debug)]
pub struct SiteRng {
    seed: u64,
    unit_lo: u32,
    unit_hi: u32,
    epoch: u32,
    site: u32,
    ordinal: [u32; N_SITES],
}

impl SiteRng {
    /// unit = `first_id + index` of the batch; `unit_hi` = bits 32..63 of the unit for games, the sample / rollout number for
    /// determinizations and leaf rollouts.
    pub fn new(seed: u64, unit_lo: u32, unit_hi: u32, epoch: u32) -> Self {
        SiteRng { seed, unit_lo, unit_hi, epoch, site: SITE_DEAL, ordinal: [0; N_SITES] }
    }
    /// Selects the call-site class of the NEXT draws (the harness derives it from `state.current_phase`).
    pub fn set_site(&mut self, site: u32) {
        self.site = site;
    }
    /// Positions a site's ordinal (rollouts that start mid-game: card picks continue at `card_index`, reservation picks at the number
    /// of reservations made — the ordinals are state-derived, not call-count-derived).
    pub fn set_ordinal(&mut self, site: u32, ordinal: u32) {
        self.ordinal[site as usize] = ordinal;
    }
    fn word(&self, site: u32, k: u32) -> u32 {
        let b = philox4x32_10([self.unit_lo, self.unit_hi, (site << 16) | (k >> 2), self.epoch], [self.seed as u32, (self.seed >> 32) as u32]);
        b[(k & 3) as usize]
    }
    /// One decision among `n` choices at the current site: `(word * n) >> 32`; consumed even for `n == 1`.
    pub fn draw(&mut self, n: u32) -> u32 {
        let site = self.site;
        let k = self.ordinal[site as usize];
        self.ordinal[site as usize] += 1;
        if site == SITE_ANNOUNCEMENT {
            // two-way decisions take one BIT each: decision k = bit k & 31 of word k >> 5; rank 1 (the call) when the bit is set
            debug_assert!(n <= 2);
            let bit = (self.word(site, k >> 5) >> (k & 31)) & 1;
            return if n == 2 { bit } else { 0 };
        }
        ((self.word(site, k) as u64 * n as u64) >> 32) as u32
    }
}

/// `RngCore` over the current site, for code paths that only need words (e.g. `rng.random::<f32>()` of the keep-experience draw):
/// `next_u32` = the next word of the current site.
impl RngCore for SiteRng {
    fn next_u32(&mut self) -> u32 {
        let site = self.site;
        let k = self.ordinal[site as usize];
        self.ordinal[site as usize] += 1;
        self.word(site, k)
    }
    fn next_u64(&mut self) -> u64 {
        let lo = self.next_u32() as u64;
        lo | ((self.next_u32() as u64) << 32)
    }
    fn fill_bytes(&mut self, dst: &mut [u8]) {
        for chunk in dst.chunks_mut(4) {
            let w = self.next_u32().to_le_bytes();
            chunk.copy_from_slice(&w[..chunk.len()]);
        }
    }
}

#[cfg(test)]
mod tests {
    use super::*;

    /// Random123 known-answer vectors for Philox4x32-10 (the ones tests/test_oracle_rng.py checks the C++ and CUDA cores against).
    #[test]
    fn philox_known_answers() {
        assert_eq!(philox4x32_10([0, 0, 0, 0], [0, 0]), [0x6627_e8d5, 0xe169_c58d, 0xbc57_ac4c, 0x9b00_dbd8]);
        assert_eq!(philox4x32_10([0xffff_ffff; 4], [0xffff_ffff; 2]), [0x408f_276d, 0x41c8_3b0e, 0xa20b_c7c6, 0x6d54_51fd]);
        assert_eq!(
            philox4x32_10([0x243f_6a88, 0x85a3_08d3, 0x1319_8a2e, 0x0370_7344], [0xa409_3822, 0x299f_31d0]),
            [0xd16c_fe09, 0x94fd_cceb, 0x5001_e420, 0x2412_6ea1]
        );
    }
}
