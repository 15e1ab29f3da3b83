"""The reference's is_consistent scripts (two real games, 21 checks; rs-full-doko/src/matching/is_consistent.rs:330-888) replayed on
the oracle — this pins the validity oracle that every determinizer sample is checked with."""
import json
import os

from oracle_lib import Fdo, hand_from_cards

T = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "is_consistent_cases.json")))["tests"]


def test_is_consistent_reference_scripts(orc):
    n = 0
    for t in T:
        s = None
        for ev in t["events"]:
            if ev[0] == "new":
                s = Fdo.from_hands(orc, [hand_from_cards(h) for h in ev[1]], ev[2])
            elif ev[0] == "play":
                # test_w predates the "4 consecutive no's after a call" rule (announcement.rs:141-146): answer NoAnnouncement
                # whenever the engine still asks and the script's next action is a card (same adaptation as the game traces)
                while s.info()["phase"] == 1 and ev[1] < 33:
                    s.play(38)
                s.play(ev[1])
            else:
                _, hands, res_start, res, expected = ev
                by_seat = [0xFF] * 4
                for i, r in enumerate(res):                 # PlayerOrientedArr::from_full(start, [..]): slot i belongs to seat start + i
                    by_seat[(res_start + i) % 4] = 0xFF if r < 0 else r
                rc = s.is_consistent([hand_from_cards(h) for h in hands], by_seat)
                assert (rc == 0) == expected, (t["name"], n, rc)
                n += 1
    assert n == 21
