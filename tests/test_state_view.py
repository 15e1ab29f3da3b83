"""include/doko_state_view.hpp (the C++ twin of rs-doko-cuda's `From<&FdoState> for dk_state` and back): compiled with g++ and checked on
records exported by the oracle — the round trip record → view → record is the identity, and the view's derived fields (trick winners,
reservation result, the call the seat to move may make) equal the oracle's state objects."""
import os
import subprocess

import numpy as np

from oracle_lib import Bulk, Fdo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0xD0C05EED


def build(tmp_path):
    exe = str(tmp_path / "state_view_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "state_view_check.cpp"), "-o", exe])
    return exe


def test_record_view_round_trip_and_derived_fields(orc, tmp_path):
    exe = build(tmp_path)
    b = Bulk(orc, 1, 3000, SEED, first_id=50, epoch=3, mode=1)
    finished = Bulk(orc, 1, 200, SEED, first_id=9000, epoch=3, mode=0)       # + some finished games
    for k in range(120):
        finished.step(500 + k)
    _, recs_f, _ = finished.step(999, want_recs=True)
    assert int((recs_f["meta"] & 3 == 3).sum()) == 200
    recs = np.concatenate([b.recs, recs_f])
    rp, vp = tmp_path / "recs.bin", tmp_path / "views.bin"
    recs.tofile(rp)
    out = subprocess.run([exe, str(rp), str(vp)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split() == [str(len(recs)), "0"]
    views = np.fromfile(vp, dtype=np.int32).reshape(len(recs), 18 + 48)
    phases = set()
    for i in range(0, len(recs), 7):
        o = Fdo.from_dk_state(orc, recs[i:i + 1])
        info, v = o.info(), views[i]
        phases.add(info["phase"])
        for k, name in enumerate(("phase", "current_player", "game_type", "card_index", "n_tricks", "team_tag", "wedding_player", "solved_idx")):
            exp = info[name]
            if name == "game_type" and exp < 0:
                exp = 15
            if name == "solved_idx" and info["team_tag"] != 2:
                continue
            assert v[k] == exp, (i, name, v[k], exp)
        if info["team_tag"] >= 2:
            assert v[8] == info["re_players"]
        assert (v[9], v[10], v[11], v[13]) == (info["re_lowest"], info["contra_lowest"], info["turns_without"], info["n_announcements"])
        if info["phase"] == 1:                                              # announcement phase: the allowed set without NoAnnouncement
            bits = info["current_allowed"] & ~64
            exp_call = {0: 0, 1: 1, 2: 2, 4: 3, 8: 4, 16: 5, 32: 1}[bits]    # the counter maps to Re/Kontra (action 33)
            assert v[14] == exp_call and v[12] == info["ann_start"]
        tr = o.tricks()
        for t in range(info["n_tricks"]):
            n_cards = int((tr[t, :4] >= 0).sum())
            assert v[18 + 4 * t] == tr[t, 4] and v[18 + 4 * t + 3] == n_cards
            if n_cards == 4:
                assert v[18 + 4 * t + 1] == tr[t, 5]
                assert v[18 + 4 * t + 2] == tr[t, (tr[t, 5] - tr[t, 4]) & 3]
            else:
                assert v[18 + 4 * t + 1] == -1
    assert phases == {0, 1, 2, 3}
