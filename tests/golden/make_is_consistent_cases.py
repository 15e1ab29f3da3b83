#!/usr/bin/env python3
"""Extract the is_consistent scripts of the reference (rs-full-doko/src/matching/is_consistent.rs:330-888: two real games, 21
checks) into tests/golden/is_consistent_cases.json as an event list per test:
  ["new", hands[4] (card lists), start] | ["play", action] | ["check", hands[4], res_start, reservations[4] (-1 = None), expected]
Run in the build container only."""
import json
import os
import re

REF = "/root/reference/rs-full-doko/src/matching/is_consistent.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "is_consistent_cases.json")
CARDS = [s + r for s in ("Diamond", "Heart", "Club", "Spade") for r in ("Nine", "Ten", "Jack", "Queen", "King", "Ace")]
CARD_ID = {n: i for i, n in enumerate(CARDS)}
PL = {"BOTTOM": 0, "LEFT": 1, "TOP": 2, "RIGHT": 3}
RES = ["Healthy", "Wedding", "DiamondsSolo", "HeartsSolo", "SpadesSolo", "ClubsSolo", "QueensSolo", "JacksSolo", "TrumplessSolo"]
ACTIONS = {"Card" + n: i for i, n in enumerate(CARDS)}
ACTIONS.update({"ReservationHealthy": 24, "ReservationWedding": 25, "ReservationDiamondsSolo": 26, "ReservationHeartsSolo": 27, "ReservationSpadesSolo": 28,
                "ReservationClubsSolo": 29, "ReservationTrumplessSolo": 30, "ReservationQueensSolo": 31, "ReservationJacksSolo": 32, "AnnouncementReContra": 33,
                "AnnouncementNo90": 34, "AnnouncementNo60": 35, "AnnouncementNo30": 36, "AnnouncementBlack": 37, "NoAnnouncement": 38})


def hands_of(txt):
    return [[CARD_ID[c] for c in re.findall(r"\b(" + "|".join(CARDS) + r")\b", h)] for h in re.findall(r"FdoHand::from_vec\(vec!\[(.*?)\]\)", txt, re.S)]


def main():
    src = re.sub(r"//[^\n]*", "", open(REF, encoding="utf-8").read())
    src = src[src.index("mod tests"):]
    tests = []
    for m in re.finditer(r"pub fn (test_\w+)\(\) \{", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        body = src[m.end():i - 1]
        events = []
        pat = re.compile(r"(FdoState::new_game_from_hand_and_start_player\((.*?)\);)|(state\.play_action\(FdoAction::(\w+)\);)|"
                         r"(let hands = PlayerZeroOrientedArr::from_full\(\[(.*?)\]\);)|(let reservations = PlayerOrientedArr::from_full\(FdoPlayer::(\w+), \[(.*?)\]\);)|"
                         r"(assert_eq!\(super::is_consistent\(.*?\), (true|false)\);)", re.S)
        hands = res = None
        for e in pat.finditer(body):
            if e.group(1):
                events.append(["new", hands_of(e.group(2)), PL[re.findall(r"FdoPlayer::(\w+)", e.group(2))[-1]]])
            elif e.group(3):
                events.append(["play", ACTIONS[e.group(4)]])
            elif e.group(5):
                hands = hands_of(e.group(6))
            elif e.group(7):
                opts = re.findall(r"(None|Some\(FdoReservation::(\w+)\))", e.group(9))
                res = [PL[e.group(8)], [RES.index(o[1]) if o[1] else -1 for o in opts]]
            else:
                events.append(["check", hands, res[0], res[1], e.group(11) == "true"])
        tests.append({"name": m.group(1), "events": events})
    json.dump({"source": "rs-full-doko/src/matching/is_consistent.rs:330-888", "tests": tests}, open(OUT, "w"), separators=(",", ":"))
    print("wrote", OUT, [(t["name"], sum(e[0] == "check" for e in t["events"]), sum(e[0] == "play" for e in t["events"])) for t in tests])


if __name__ == "__main__":
    main()
