//! The parity stream of the simulator (DESIGN.md "Philox parity contract") on the Rust side.
//!
//! The kernels draw every random decision from Philox4x32-10 with
//!     key = 64-bit seed,   counter = (unit_lo, unit_hi, site << 16 | block, epoch)
//! where a *site* is a class of reference call sites (deal, reservation pick, announcement decision, card pick, card_matching's rule 4,
//! hidden reservations, ...) and the k-th use of a site inside one unit takes word `k & 3` of block `k >> 2` (announcement decisions:
//! bit `k & 31` of word `k >> 5`), mapped onto `n` choices by `(word as u64 * n as u64) >> 32`.
//! Two sites draw several times from one word ("chained draws": the high half of `v * n` is the draw, the low half — the fractional part
//! of `v * n / 2^32`, uniform again — feeds the next draw; relative bias below the product of the counts / 2^32): the deal takes word
//! `min(s / 3, 11)` for its draw `s` (0 = start seat, then the shuffle steps), and the card pick at `card_index` takes word
//! `card_index / 4` — the four picks of a trick share the trick's word.
//!
//! "The reference's harness fed the same Philox stream" therefore needs two things the stock `rand` plumbing cannot give:
//!  1. the draw must know its SITE — `Bitflag::random_single` (rs-game-utils/src/bit_flag.rs:86-94) and `FdoHandIter::choose`
//!     (rs-full-doko/src/matching/card_matching.rs:182-186) receive a bare `&mut SmallRng`;
//!  2. `rand 0.9`'s `random_range` maps a word with a widening multiply AND a rejection step that may draw a second word, and its
//!     `shuffle` batches several indices into one word (`IncreasingUniform`); the contract uses one word per decision (or its own chained
//!     draws, above), no rejection.
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
    /// chained draws: (word index of the running chain, what is left of that word) of the deal and of the card picks
    deal_chain: (u32, u32),
    card_chain: (u32, u32),
    /// factor for the first card pick after `set_card_position` inside a trick
    card_mul: u32,
}

impl SiteRng {
    /// unit = `first_id + index` of the batch; `unit_hi` = bits 32..63 of the unit for games, the sample / rollout number for
    /// determinizations and leaf rollouts.
    pub fn new(seed: u64, unit_lo: u32, unit_hi: u32, epoch: u32) -> Self {
        SiteRng { seed, unit_lo, unit_hi, epoch, site: SITE_DEAL, ordinal: [0; N_SITES], deal_chain: (u32::MAX, 0), card_chain: (u32::MAX, 0), card_mul: 1 }
    }
    /// Selects the call-site class of the NEXT draws (the harness derives it from `state.current_phase`).
    pub fn set_site(&mut self, site: u32) {
        self.site = site;
    }
    /// Positions a site's ordinal (rollouts that start mid-game: card picks continue at `card_index`, reservation picks at the number
    /// of reservations made — the ordinals are state-derived, not call-count-derived).
    pub fn set_ordinal(&mut self, site: u32, ordinal: u32) {
        debug_assert!(site != SITE_CARD || ordinal & 3 == 0, "inside a trick use set_card_position");
        self.ordinal[site as usize] = ordinal;
        if site == SITE_DEAL {
            self.deal_chain.0 = u32::MAX;
        }
        if site == SITE_CARD {
            self.card_chain.0 = u32::MAX;
            self.card_mul = 1;
        }
    }
    /// Positions the card picks at `card_index`.  Inside a trick `chain_mul` is the product of the numbers of legal card types the plays
    /// already made in that trick chose from (each: the seat's hand with its card back in it, restricted to the led colour when it can
    /// follow) — the trick's word times this product is where its chained draws continue.
    pub fn set_card_position(&mut self, card_index: u32, chain_mul: u32) {
        self.ordinal[SITE_CARD as usize] = card_index;
        self.card_chain.0 = u32::MAX;
        self.card_mul = if card_index & 3 != 0 { chain_mul } else { 1 };
    }
    fn chain(v: &mut u32, n: u32) -> u32 {
        let p = *v as u64 * n as u64;
        *v = p as u32;
        (p >> 32) as u32
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
        if site == SITE_DEAL {
            // draw k of the deal: three chained draws per word, the twelfth word serves the draws 33..36
            let w = if k < 33 { k / 3 } else if k < 37 { 11 } else { 12 + (k - 37) / 3 };
            if w != self.deal_chain.0 {
                self.deal_chain = (w, self.word(site, w));
            }
            return Self::chain(&mut self.deal_chain.1, n);
        }
        if site == SITE_CARD {
            // the pick at card_index k belongs to trick k / 4, which owns word k / 4; the trick's picks are chained
            if k >> 2 != self.card_chain.0 {
                self.card_chain = (k >> 2, self.word(site, k >> 2).wrapping_mul(self.card_mul));
            }
            self.card_mul = 1;
            return Self::chain(&mut self.card_chain.1, n);
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

    /// Chained draws against the closed form of the contract (tests/test_oracle_rng.py::test_chained_draws_closed_form).
    #[test]
    fn chained_draws_closed_form() {
        let (seed, ul, uh, ep) = (0x0123_4567_89AB_CDEFu64, 4711u32, 3u32, 6u32);
        let word = |site: u32, k: u32| philox4x32_10([ul, uh, (site << 16) | (k >> 2), ep], [seed as u32, (seed >> 32) as u32])[(k & 3) as usize];
        let mut r = SiteRng::new(seed, ul, uh, ep);
        r.set_site(SITE_DEAL);
        let (mut cur, mut mul) = (u32::MAX, 1u32);
        for s in 0..37u32 {
            let n = if s == 0 { 4 } else { 49 - s };
            let w = (s / 3).min(11);
            if w != cur {
                cur = w;
                mul = 1;
            }
            assert_eq!(r.draw(n), ((word(SITE_DEAL, w).wrapping_mul(mul) as u64 * n as u64) >> 32) as u32);
            mul = mul.wrapping_mul(n);
        }
        let ns = [12u32, 7, 7, 3, 11, 11, 2, 9, 10, 1, 4, 6];
        r.set_site(SITE_CARD);
        let mut all = Vec::new();
        for (ci, &n) in ns.iter().enumerate() {
            let mul = ns[ci & !3..ci].iter().fold(1u32, |a, &b| a.wrapping_mul(b));
            let d = r.draw(n);
            assert_eq!(d, ((word(SITE_CARD, ci as u32 >> 2).wrapping_mul(mul) as u64 * n as u64) >> 32) as u32);
            all.push(d);
        }
        for first in 0..ns.len() {
            let mut q = SiteRng::new(seed, ul, uh, ep);
            q.set_site(SITE_CARD);
            q.set_card_position(first as u32, ns[first & !3..first].iter().product());
            for ci in first..ns.len() {
                assert_eq!(q.draw(ns[ci]), all[ci]);
            }
        }
    }

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
