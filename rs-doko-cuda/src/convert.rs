//! `FdoState` (rs-full-doko/src/state/state.rs:24-75) <-> `dk_state` (include/doko_cuda.h), field by field.
//!
//! The rules are those of `include/doko_state_view.hpp` (`to_record` / `from_record`), which is compiled and checked against the
//! oracle's exports in this repository (tests/test_state_view.py); this file is the same mapping written against the reference's types.
//! Stored: reservations (play order), the 48 played cards (play order), hands, announcement occurrences, lowest calls, round
//! counters, card index, seat / phase, game type, eyes, trick counts, team state, final points.
//! Derived on the way back: trick winners (winner of trick t leads trick t+1; the last one through `FdoTrick::play_card`), the
//! reservation result (`winning_player_in_reservation_round`), the allowed calls of the seat to move (`calc_allowed_announcements`),
//! the end-of-game statistics (`FdoEndOfGameStats::calculate`).
use rs_full_doko::announcement::announcement::{FdoAnnouncement, FdoAnnouncementOccurrence, FdoAnnouncements};
use rs_full_doko::announcement::announcement_set::FdoAnnouncementSet;
use rs_full_doko::announcement::calc_announcement::calc_allowed_announcements;
use rs_full_doko::basic::phase::FdoPhase;
use rs_full_doko::card::cards::FdoCard;
use rs_full_doko::game_type::game_type::FdoGameType;
use rs_full_doko::hand::hand::FdoHand;
use rs_full_doko::player::player::FdoPlayer;
use rs_full_doko::player::player_set::FdoPlayerSet;
use rs_full_doko::reservation::reservation::FdoReservation;
use rs_full_doko::reservation::reservation_round::FdoReservationRound;
use rs_full_doko::reservation::reservation_winning_logic::winning_player_in_reservation_round;
use rs_full_doko::state::state::FdoState;
use rs_full_doko::stats::stats::FdoEndOfGameStats;
use rs_full_doko::team::team_logic::FdoTeamState;
use rs_full_doko::trick::trick::FdoTrick;
use rs_full_doko::util::po_zero_arr::PlayerZeroOrientedArr;
use strum::IntoEnumIterator;

use crate::ffi::*;

/// card id 0..23 (suit * 6 + rank) of an `FdoCard` (`1 << id`, cards.rs:7-36)
pub fn card_id(card: FdoCard) -> u8 {
    (card as usize).trailing_zeros() as u8
}
pub fn card_from_id(id: u8) -> FdoCard {
    FdoCard::iter().nth(id as usize).expect("card id 0..23")
}
/// level code DK_ANN_* of an `Option<FdoAnnouncement>` (announcement.rs:12-22: the enum values are `1 << k`)
pub fn ann_code(a: Option<FdoAnnouncement>) -> u32 {
    match a {
        None | Some(FdoAnnouncement::NoAnnouncement) => DK_ANN_NONE as u32,
        Some(x) => (x as usize).trailing_zeros() + 1, // ReContra 1, No90 2, No60 3, No30 4, Black 5, CounterReContra 6
    }
}
pub fn ann_from_code(code: u32) -> Option<FdoAnnouncement> {
    match code {
        1 => Some(FdoAnnouncement::ReContra),
        2 => Some(FdoAnnouncement::No90),
        3 => Some(FdoAnnouncement::No60),
        4 => Some(FdoAnnouncement::No30),
        5 => Some(FdoAnnouncement::Black),
        6 => Some(FdoAnnouncement::CounterReContra),
        _ => None,
    }
}
/// DK_RES_* = declaration order of `FdoReservation` (reservation.rs:11-24)
pub fn res_code(r: FdoReservation) -> u8 {
    r as usize as u8
}
pub fn res_from_code(code: u8) -> FdoReservation {
    match code {
        0 => FdoReservation::Healthy,
        1 => FdoReservation::Wedding,
        2 => FdoReservation::DiamondsSolo,
        3 => FdoReservation::HeartsSolo,
        4 => FdoReservation::SpadesSolo,
        5 => FdoReservation::ClubsSolo,
        6 => FdoReservation::QueensSolo,
        7 => FdoReservation::JacksSolo,
        8 => FdoReservation::TrumplessSolo,
        _ => panic!("reservation code {}", code),
    }
}
/// DK_GT_* = declaration order of `FdoGameType` (game_type.rs:6-20)
pub fn gt_code(g: Option<FdoGameType>) -> u32 {
    match g {
        None => DK_GT_NONE as u32,
        Some(FdoGameType::Normal) => 0,
        Some(FdoGameType::Wedding) => 1,
        Some(FdoGameType::DiamondsSolo) => 2,
        Some(FdoGameType::HeartsSolo) => 3,
        Some(FdoGameType::SpadesSolo) => 4,
        Some(FdoGameType::ClubsSolo) => 5,
        Some(FdoGameType::TrumplessSolo) => 6,
        Some(FdoGameType::QueensSolo) => 7,
        Some(FdoGameType::JacksSolo) => 8,
    }
}
pub fn gt_from_code(code: u32) -> Option<FdoGameType> {
    FdoGameType::iter().nth(code as usize) // 15 (DK_GT_NONE) -> None
}
/// The 48-bit board of a hand: copy A of card c = bit c, copy B = bit c + 24 (hand.rs:20,107), rebuilt through the public accessors.
pub fn hand_bits(h: &FdoHand) -> u64 {
    let mut bits = 0u64;
    for card in FdoCard::iter() {
        let c = card_id(card) as u64;
        if h.contains(card) {
            bits |= 1 << c;
        }
        if h.contains_both(card) {
            bits |= 1 << (c + 24);
        }
    }
    bits
}
pub fn hand_from_bits(bits: u64) -> FdoHand {
    let mut h = FdoHand::empty();
    for c in 0..24u8 {
        if bits >> c & 1 != 0 {
            h.add(card_from_id(c)); // copy A first ...
        }
        if bits >> (c + 24) & 1 != 0 {
            h.add(card_from_id(c)); // ... then copy B (hand.rs:217-230)
        }
    }
    h
}
fn player_set_bits(s: &FdoPlayerSet) -> u32 {
    s.iter().fold(0u32, |m, p| m | 1 << p.index())
}
fn player_set_from_bits(m: u32) -> FdoPlayerSet {
    FdoPlayerSet::from_vec((0..4).filter(|p| m >> p & 1 != 0).map(FdoPlayer::from_index).collect())
}

impl From<&FdoState> for dk_state {
    fn from(s: &FdoState) -> dk_state {
        let mut r = dk_state {
            hands: [0; 4], cards: [0xFF; 48], announcements: [0xFFFF; 12], reservations: [DK_RES_NONE as u8; 4], tricks: 0, eyes: [0; 4], num_tricks: 0,
            card_index: s.card_index as u8, n_reservations: 0, points: [0; 4], meta: 0,
        };
        for p in 0..4 {
            let pl = FdoPlayer::from_index(p);
            r.hands[p] = hand_bits(&s.hands[pl]);
            r.eyes[p] = s.player_eyes[pl] as u8;
            r.num_tricks |= (s.player_num_tricks[pl] as u16) << (4 * p);
        }
        for (k, res) in s.reservations_round.reservations.iter().enumerate() {
            r.reservations[k] = res_code(*res); // play order from the starting player
        }
        r.n_reservations = s.reservations_round.reservations.len() as u8;
        let mut ci = 0usize;
        for (t, trick) in s.tricks.iter().enumerate() {
            r.tricks |= (trick.cards.starting_player.index() as u32) << (2 * t);
            for card in trick.cards.iter() {
                r.cards[ci] = card_id(*card);
                ci += 1;
            }
        }
        r.tricks |= (s.tricks.len() as u32) << 24;
        r.tricks |= (s.announcements.announcements.len() as u32) << 28;
        for (k, occ) in s.announcements.announcements.iter().enumerate() {
            r.announcements[k] = (occ.card_index as u16) | ((occ.player.index() as u16) << 6) | ((ann_code(Some(occ.announcement)) as u16) << 8);
        }
        let finished = s.current_phase == FdoPhase::Finished;
        if let (true, Some(stats)) = (finished, s.end_of_game_stats.as_ref()) {
            for p in 0..4 {
                r.points[p] = stats.player_points[FdoPlayer::from_index(p)] as i8;
            }
        }
        let (tag, wed, solved, re) = match s.team_state {
            FdoTeamState::InReservations => (DK_TEAM_IN_RESERVATIONS, 0, 0, 0),
            FdoTeamState::WeddingUnsolved { wedding_player } => (DK_TEAM_WEDDING_UNSOLVED, wedding_player.index() as u32, 0, 0),
            FdoTeamState::WeddingSolved { wedding_player, solved_trick_index, re_players } => {
                (DK_TEAM_WEDDING_SOLVED, wedding_player.index() as u32, solved_trick_index as u32, player_set_bits(&re_players))
            }
            FdoTeamState::NoWedding { re_players } => (DK_TEAM_NO_WEDDING, 0, 0, player_set_bits(&re_players)),
        };
        r.meta = (s.current_phase as usize as u32)
            | (s.current_player.map(|p| p.index() as u32).unwrap_or(0) << 2)
            | ((s.reservations_round.reservations.starting_player.index() as u32) << 4)
            | (gt_code(s.game_type) << 6)
            | ((tag as u32) << 10)
            | (wed << 12)
            | (solved << 14)
            | (re << 16)
            | (ann_code(s.announcements.re_lowest_announcement) << 20)
            | (ann_code(s.announcements.contra_lowest_announcement) << 23)
            | ((s.announcements.number_of_turns_without_announcement as u32) << 26)
            | ((s.announcements.starting_player.index() as u32) << 29);
        r
    }
}

impl From<&dk_state> for FdoState {
    fn from(r: &dk_state) -> FdoState {
        let (m, tr) = (r.meta, r.tricks);
        let phase = match m & 3 {
            0 => FdoPhase::Reservation,
            1 => FdoPhase::Announcement,
            2 => FdoPhase::PlayCard,
            _ => FdoPhase::Finished,
        };
        let start = FdoPlayer::from_index(((m >> 4) & 3) as usize);
        let reservations_round =
            FdoReservationRound::existing(start, (0..r.n_reservations as usize).map(|k| res_from_code(r.reservations[k])).collect());
        let game_type = gt_from_code((m >> 6) & 15);
        let hands = PlayerZeroOrientedArr::from_full([hand_from_bits(r.hands[0]), hand_from_bits(r.hands[1]), hand_from_bits(r.hands[2]), hand_from_bits(r.hands[3])]);
        let player_eyes = PlayerZeroOrientedArr::from_full([r.eyes[0] as u32, r.eyes[1] as u32, r.eyes[2] as u32, r.eyes[3] as u32]);
        let nt = r.num_tricks as u32;
        let player_num_tricks = PlayerZeroOrientedArr::from_full([nt & 15, (nt >> 4) & 15, (nt >> 8) & 15, (nt >> 12) & 15]);
        // tricks: FdoTrick::play_card recomputes winning_player / winning_card of complete tricks (trick.rs:84-104)
        let mut tricks: heapless::Vec<FdoTrick, 12> = heapless::Vec::new();
        let n_tricks = ((tr >> 24) & 15) as usize;
        let mut ci = 0usize;
        for t in 0..n_tricks {
            let mut trick = FdoTrick::empty(FdoPlayer::from_index(((tr >> (2 * t)) & 3) as usize));
            for _ in 0..4 {
                if ci < r.card_index as usize {
                    trick.play_card(card_from_id(r.cards[ci]), game_type.expect("cards are played after the reservations"));
                    ci += 1;
                }
            }
            tricks.push(trick).expect("<= 12 tricks");
        }
        let re_players = player_set_from_bits((m >> 16) & 15);
        let wedding_player = FdoPlayer::from_index(((m >> 12) & 3) as usize);
        let team_state = match (m >> 10) & 3 {
            0 => FdoTeamState::InReservations,
            1 => FdoTeamState::WeddingUnsolved { wedding_player },
            2 => FdoTeamState::WeddingSolved { wedding_player, solved_trick_index: ((m >> 14) & 3) as usize, re_players },
            _ => FdoTeamState::NoWedding { re_players },
        };
        let current_player = if phase == FdoPhase::Finished { None } else { Some(FdoPlayer::from_index(((m >> 2) & 3) as usize)) };
        let mut announcements = FdoAnnouncements::new();
        for k in 0..((tr >> 28) & 15) as usize {
            let o = r.announcements[k];
            announcements
                .announcements
                .push(FdoAnnouncementOccurrence {
                    card_index: (o & 63) as usize,
                    player: FdoPlayer::from_index(((o >> 6) & 3) as usize),
                    announcement: ann_from_code(((o >> 8) & 7) as u32).expect("a recorded call has a level"),
                })
                .expect("<= 12 calls");
        }
        announcements.re_lowest_announcement = ann_from_code((m >> 20) & 7);
        announcements.contra_lowest_announcement = ann_from_code((m >> 23) & 7);
        announcements.number_of_turns_without_announcement = ((m >> 26) & 7) as usize;
        announcements.starting_player = FdoPlayer::from_index(((m >> 29) & 3) as usize);
        announcements.current_player_allowed_announcements = match (phase, current_player) {
            (FdoPhase::Announcement, Some(p)) => calc_allowed_announcements(
                p,
                hands[p].len(),
                team_state,
                announcements.re_lowest_announcement,
                announcements.contra_lowest_announcement,
            ),
            _ => FdoAnnouncementSet::new(),
        };
        let reservation_result = if reservations_round.is_completed() { Some(winning_player_in_reservation_round(&reservations_round)) } else { None };
        let end_of_game_stats = if phase == FdoPhase::Finished {
            Some(FdoEndOfGameStats::calculate(
                player_eyes,
                player_num_tricks,
                re_players,
                announcements.re_lowest_announcement,
                announcements.contra_lowest_announcement,
                &tricks,
            ))
        } else {
            None
        };
        FdoState {
            reservations_round,
            tricks,
            hands,
            announcements,
            card_index: r.card_index as usize,
            current_player,
            current_phase: phase,
            reservation_result,
            game_type,
            player_eyes,
            player_num_tricks,
            team_state,
            end_of_game_stats,
        }
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use rand::prelude::SmallRng;
    use rand::SeedableRng;

    /// record -> FdoState -> record is the identity on every state of random games, and FdoState -> record -> FdoState reproduces the
    /// reference's own state (`FdoState: PartialEq`).
    #[test]
    fn round_trip_on_random_games() {
        let mut rng = SmallRng::seed_from_u64(1711);
        for _ in 0..200 {
            let mut state = FdoState::new_game(&mut rng);
            loop {
                let rec = dk_state::from(&state);
                let back = FdoState::from(&rec);
                assert_eq!(back, state);
                let rec2 = dk_state::from(&back);
                assert_eq!(unsafe { std::mem::transmute::<dk_state, [u8; 128]>(rec) }, unsafe { std::mem::transmute::<dk_state, [u8; 128]>(rec2) });
                if state.random_action_for_current_player(&mut rng) {
                    break;
                }
            }
        }
    }
}
