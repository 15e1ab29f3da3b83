"""Reservation / team / announcement-protocol / rs-doko legal-action known answers of the reference
(fixture tests/golden/protocol_tables.json, made by tests/golden/make_protocol_tables.py) replayed on the oracle."""
import ctypes as C
import json
import os

import numpy as np

from oracle_lib import hand_from_cards

T = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "protocol_tables.json")))
GT_OF_RES = {2: 2, 3: 3, 4: 4, 5: 5, 6: 7, 7: 8, 8: 6}


def test_reservation_winner(orc):
    for r in T["reservation_winner"]:
        out = (C.c_int32 * 4)()
        orc.orc_fdo_reservation_result(r["start"], len(r["res"]), (C.c_int32 * 4)(*r["res"]), out)
        assert [out[0], out[1], out[2]] == [r["kind"], r["player"], r["reservation"]]
        assert out[3] == (0 if r["kind"] == 0 else 1 if r["kind"] == 2 else GT_OF_RES[r["reservation"]])
    assert len(T["reservation_winner"]) == 3


def test_visible_reservations(orc):
    for r in T["visible_reservations"]:
        out = (C.c_int32 * 4)()
        res = r["res"] + [0] * (4 - len(r["res"]))
        orc.orc_fdo_visible(r["start"], len(r["res"]), (C.c_int32 * 4)(*res), r["observer"], out)
        assert list(out) == r["expected"], r
    assert len(T["visible_reservations"]) >= 5


def test_team_resolve(orc):
    for r in T["team_resolve"]:
        tricks = np.full(60, -1, dtype=np.int32)
        for t, tr in enumerate(r["tricks"]):
            tricks[t * 5:t * 5 + len(tr["cards"])] = tr["cards"]
            tricks[t * 5 + 4] = tr["start"]
        out = (C.c_int32 * 4)()
        orc.orc_fdo_team_resolve((C.c_int32 * 3)(*r["rr"]), len(r["tricks"]), tricks.ctypes.data_as(C.c_void_p),
                                 (C.c_uint64 * 4)(*[hand_from_cards(h) for h in r["hands"]]), out)
        assert out[0] == r["tag"], r["name"]
        for p in r["re_contains"]:
            assert (out[3] >> p) & 1
        if r["re_len"] is not None:
            assert bin(out[3]).count("1") == r["re_len"]
        if r["wedding_player"] is not None:
            assert out[1] == r["wedding_player"]
        if r["solved_idx"] is not None:
            assert out[2] == r["solved_idx"]
        assert (out[0] in (2, 3)) == r["is_final"]
    assert len(T["team_resolve"]) >= 6


def test_announcement_protocol_scripts(orc):
    """announcement.rs:228-573: start_round / play_announcement walk-throughs incl. auto-skip and counter."""
    orc.orc_ann_new.restype = C.c_void_p
    n_steps = 0
    for sc in T["announcement_scripts"]:
        a = C.c_void_p(orc.orc_ann_new())
        t = sc["team"]
        for st in sc["steps"]:
            out = (C.c_int32 * 24)()
            rc = orc.orc_ann_step(a, st["op"], st["player"], st["ann"], sc["card_index"], (C.c_uint32 * 4)(*sc["lens"]), t["tag"], t["wedding_player"],
                                  t["solved_idx"], t["re_players"], out)
            assert rc == 0
            assert [out[0], out[1]] == st["result"], (sc["name"], st)
            for key, idx in (("n", 2), ("starting_player", 3), ("turns", 4), ("re_lowest", 5), ("contra_lowest", 6), ("allowed", 7)):
                if st[key] is not None:
                    assert out[idx] == st[key], (sc["name"], key, st)
            n_steps += 1
        orc.orc_ann_free(a)
    assert n_steps >= 14


def test_rs_doko_allowed_actions(orc):
    for r in T["doko_allowed_actions"]:
        assert orc.orc_doko_allowed_actions(r["phase"], r["color"], hand_from_cards(r["hand"])) == r["expected"], r
    assert len(T["doko_allowed_actions"]) >= 6
