#!/usr/bin/env python3
"""Extract the reference's hand-built encode_state_pi known-answer test into tests/golden/encode_pi_vector.json.

Source: /root/reference/rs-doko-networks/src/full_doko/var1/encode_pi.rs:237-793 (fn test_encode_state): the FdoState struct
literal and the 311 `assert_eq_inc!` lines.  The tiny helper encoders named on those lines are evaluated here from their
definitions (var2/encode_reservation_or_card_or_none.rs, var2/encode_position_or_unknown.rs, var1/player.rs, ...).
Run in the build container only.
"""
import json
import os
import re

REF = "/root/reference/rs-doko-networks/src/full_doko/var1/encode_pi.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "encode_pi_vector.json")
CARDS = [s + r for s in ("Diamond", "Heart", "Club", "Spade") for r in ("Nine", "Ten", "Jack", "Queen", "King", "Ace")]
CARD_ID = {n: i for i, n in enumerate(CARDS)}
PL = {"BOTTOM": 0, "LEFT": 1, "TOP": 2, "RIGHT": 3}
RES = ["Healthy", "Wedding", "DiamondsSolo", "HeartsSolo", "SpadesSolo", "ClubsSolo", "QueensSolo", "JacksSolo", "TrumplessSolo"]
GT = ["Normal", "Wedding", "DiamondsSolo", "HeartsSolo", "SpadesSolo", "ClubsSolo", "TrumplessSolo", "QueensSolo", "JacksSolo"]
ANN = {"ReContra": 1, "No90": 2, "No60": 3, "No30": 4, "Black": 5}
PHASE = {"Reservation": 0, "Announcement": 1, "PlayCard": 2, "Finished": 3}
CARD_TOKEN = {"HeartTen": 1, "ClubQueen": 2, "SpadeQueen": 3, "HeartQueen": 4, "DiamondQueen": 5, "ClubJack": 6, "SpadeJack": 7, "HeartJack": 8,
              "DiamondJack": 9, "DiamondAce": 10, "DiamondTen": 11, "DiamondKing": 12, "DiamondNine": 13, "ClubAce": 14, "ClubTen": 15,
              "ClubKing": 16, "ClubNine": 17, "SpadeAce": 18, "SpadeTen": 19, "SpadeKing": 20, "SpadeNine": 21, "HeartAce": 22, "HeartKing": 23,
              "HeartNine": 24}
VIS_TOKEN = {"Healthy": 25, "Wedding": 26, "DiamondsSolo": 27, "HeartsSolo": 28, "SpadesSolo": 29, "ClubsSolo": 30, "QueensSolo": 31,
             "JacksSolo": 32, "TrumplessSolo": 33, "NotRevealed": 34, "NoneYet": 35}


def evaluate(expr):
    m = re.match(r"(\w+)\((.*)\)$", expr)
    fn, arg = m.group(1), m.group(2)
    if fn == "encode_announcement_team":
        return 0 if arg == "None" else (1 if "Re" in arg else 2)
    if fn == "encode_phase":
        return PHASE[arg.split("::")[1]]
    if fn == "encode_player_or_none":
        a, rel = [x.strip() for x in arg.rsplit(",", 1)]
        return 0 if a == "None" else (PL[re.search(r"FdoPlayer::(\w+)", a).group(1)] - PL[rel.split("::")[1]]) % 4 + 1
    if fn == "encode_position_or_unknown_hand":
        a, rel = [x.strip() for x in arg.rsplit(",", 1)]
        return 0 if a == "None" else (PL[re.search(r"FdoPlayer::(\w+)", a).group(1)] - PL[rel.split("::")[1]]) % 4 + 1 + 52
    if fn == "encode_position_or_unknown_int":
        return int(arg) + 1
    if fn == "encode_reservation_or_card_or_none_card":
        return 0 if arg == "None" else CARD_TOKEN[re.search(r"FdoCard::(\w+)", arg).group(1)]
    if fn == "encode_reservation_or_card_or_pi_announcement":
        return 37 if arg == "None" else 37 + ANN[re.search(r"Some\((\w+)\)", arg).group(1)]
    if fn == "encode_reservation_or_card_or_reservation":
        return VIS_TOKEN[re.search(r"FdoVisibleReservation::(\w+)", arg).group(1)]
    if fn == "encode_subposition_card":
        return 0 if arg == "None" else 11 + int(re.search(r"Some\((\d+)\)", arg).group(1))
    if fn == "encode_subposition_pos":
        return 0 if arg == "None" else int(re.search(r"Some\((\d+)\)", arg).group(1)) + 1
    raise ValueError(expr)


def parse_state(lit):
    """The FdoState struct literal of a hand-built reference test → plain dict."""
    m = re.search(r"reservations: PlayerOrientedVec::from_full\(FdoPlayer::(\w+), vec!\[(.*?)\]\)", lit, re.S)
    start = PL[m.group(1)]
    reservations = [RES.index(x) for x in re.findall(r"FdoReservation::(\w+)", m.group(2))]
    tricks = []
    for tm in re.finditer(r"cards: PlayerOrientedVec::from_full\(FdoPlayer::(\w+), vec!\[(.*?)\]\)", lit, re.S):
        tricks.append({"start": PL[tm.group(1)], "cards": [CARD_ID[c] for c in re.findall(r"(\w+)", tm.group(2)) if c in CARD_ID]})
    hands = [[CARD_ID[c] for c in re.findall(r"(\w+)", h) if c in CARD_ID] for h in re.findall(r"FdoHand::from_vec\(vec!\[(.*?)\]\)", lit)]
    calls = [{"card_index": int(a), "player": PL[b], "level": ANN[c]} for a, b, c in
             re.findall(r"card_index: (\d+),\s*player: FdoPlayer::(\w+),\s*announcement: (\w+)", lit)]
    state = {
        "start": start, "reservations": reservations, "tricks": tricks, "hands": hands, "calls": calls,
        "re_lowest": ANN[re.search(r"re_lowest_announcement: Some\((\w+)\)", lit).group(1)],
        "contra_lowest": ANN[re.search(r"contra_lowest_announcement: Some\((\w+)\)", lit).group(1)],
        "turns_without": int(re.search(r"number_of_turns_without_announcement: (\d+)", lit).group(1)),
        "ann_start": PL[re.search(r"starting_player: FdoPlayer::(\w+)", lit).group(1)],
        "card_index": int(re.search(r"\n\s*card_index: (\d+),\n\s*current_player", lit).group(1)),
        "current_player": PL[re.search(r"current_player: Some\(FdoPlayer::(\w+)\)", lit).group(1)],
        "phase": PHASE[re.search(r"current_phase: FdoPhase::(\w+)", lit).group(1)],
        "game_type": GT.index(re.search(r"game_type: Some\((\w+)\)", lit).group(1)),
        "eyes": [int(x) for x in re.search(r"player_eyes: PlayerZeroOrientedArr::from_full\(\[(.*?)\]\)", lit).group(1).split(",")],
        "num_tricks": [int(x) for x in re.search(r"player_num_tricks: PlayerZeroOrientedArr::from_full\(\[(.*?)\]\)", lit).group(1).split(",")],
        "re_players": [PL[p] for p in re.findall(r"FdoPlayer::(\w+)", re.search(r"re_players: FdoPlayerSet::from_vec\(vec!\[(.*?)\]\)", lit).group(1))],
    }
    return state


def main():
    src = open(REF, encoding="utf-8").read()
    body = src[src.index("fn test_encode_state()"):]
    state = parse_state(body[:body.index("let result = encode_state_pi(")])
    expected = [evaluate(x.strip()) for x in re.findall(r"assert_eq_inc!\(result\[i\.\.i\+1\], (.*)\);", body)]
    assert len(expected) == 311, len(expected)
    json.dump({"source": "rs-doko-networks/src/full_doko/var1/encode_pi.rs:237-793", "state": state, "expected": expected}, open(OUT, "w"))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
