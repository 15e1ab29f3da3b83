#!/usr/bin/env python3
"""Extract the four action-by-action real-game traces from the reference's (commented-out) tests.

Source: /root/reference/rs-full-doko/src/state/state.rs:525-1811 (tests full_normal_game ... at :551, :955,
:1350, :1558).  Run in the build container only (the reference is not on the GPU box):

    python tests/golden/make_fdo_traces.py        # writes tests/golden/fdo_traces.json

Per game it records: hands (as lists of card ids in the order the test adds them), start seat, the
`play_action` sequence, and after every action the phase / current player / game type / allowed actions
the reference test asserts, plus the final observation's tricks, eyes, points, re players, calls.
"""
import json
import os
import re

REF = "/root/reference/rs-full-doko/src/state/state.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fdo_traces.json")

CARDS = [s + r for s in ("Diamond", "Heart", "Club", "Spade") for r in ("Nine", "Ten", "Jack", "Queen", "King", "Ace")]
CARD_ID = {n: i for i, n in enumerate(CARDS)}
ACTIONS = {"Card" + n: i for i, n in enumerate(CARDS)}
ACTIONS.update({"ReservationHealthy": 24, "ReservationWedding": 25, "ReservationDiamondsSolo": 26, "ReservationHeartsSolo": 27,
                "ReservationSpadesSolo": 28, "ReservationClubsSolo": 29, "ReservationTrumplessSolo": 30,
                "ReservationQueensSolo": 31, "ReservationJacksSolo": 32, "AnnouncementReContra": 33, "AnnouncementNo90": 34,
                "AnnouncementNo60": 35, "AnnouncementNo30": 36, "AnnouncementBlack": 37, "NoAnnouncement": 38})
PLAYERS = {"BOTTOM": 0, "LEFT": 1, "TOP": 2, "RIGHT": 3}
PHASES = {"Reservation": 0, "Announcement": 1, "PlayCard": 2, "Finished": 3}
GAME_TYPES = ["Normal", "Wedding", "DiamondsSolo", "HeartsSolo", "SpadesSolo", "ClubsSolo", "TrumplessSolo", "QueensSolo", "JacksSolo"]
ANN = {"ReContra": 1, "No90": 2, "No60": 3, "No30": 4, "Black": 5, "CounterReContra": 6}


def parse_obs(line):
    o = {}
    m = re.search(r"game_type: (None|Some\(FdoGameType::(\w+)\))", line)
    o["game_type"] = GAME_TYPES.index(m.group(2)) if m.group(2) else -1
    o["phase"] = PHASES[re.search(r"phase: FdoPhase::(\w+)", line).group(1)]
    m = re.search(r"current_player: (None|Some\(FdoPlayer::(\w+)\))", line)
    o["current_player"] = PLAYERS[m.group(2)] if m.group(2) else -1
    m = re.search(r"allowed_actions_current_player: FdoAllowedActions::from_vec\(vec!\[(.*?)\]\)", line)
    acts = [a.strip() for a in m.group(1).split(",") if a.strip()]
    mask = 0
    for a in acts:
        mask |= 1 << ACTIONS[a.split("::")[-1]]
    o["allowed"] = mask
    m = re.search(r"player_eyes: \[(\d+), (\d+), (\d+), (\d+)\]", line)
    if m:
        o["eyes"] = [int(x) for x in m.groups()]
    return o


def parse_final(line):
    f = {}
    tricks = []
    for m in re.finditer(r"FdoTrick \{ cards: \[(.*?)\], start_player: FdoPlayer::(\w+), winning_player: Some\(FdoPlayer::(\w+)\), "
                         r"winning_card: Some\((\w+)\) \}", line):
        cards = [CARD_ID[c] for c in re.findall(r"Some\((\w+)\)", m.group(1))]
        tricks.append({"cards": cards, "start": PLAYERS[m.group(2)], "winner": PLAYERS[m.group(3)], "winning_card": CARD_ID[m.group(4)]})
    f["tricks"] = tricks
    fs = line[line.index("finished_stats:"):]
    f["re_players"] = sorted(PLAYERS[p] for p in re.findall(r"FdoPlayer::(\w+)", re.search(r"re_players: FdoPlayerSet::from_vec\(vec!\[(.*?)\]\)", fs).group(1)))
    f["is_solo"] = re.search(r"is_solo: (\w+)", fs).group(1) == "true"
    f["player_eyes"] = [int(x) for x in re.search(r"player_eyes: \[(\d+), (\d+), (\d+), (\d+)\]", fs).groups()]
    f["re_eyes"] = int(re.search(r"re_eyes: (\d+)", fs).group(1))
    f["kontra_eyes"] = int(re.search(r"kontra_eyes: (\d+)", fs).group(1))
    f["re_points"] = int(re.search(r"re_points: (-?\d+)", fs).group(1))
    f["kontra_points"] = int(re.search(r"kontra_points: (-?\d+)", fs).group(1))
    f["player_points"] = [int(x) for x in re.search(r"player_points: \[(-?\d+), (-?\d+), (-?\d+), (-?\d+)\]", fs).groups()]
    m = re.search(r"additional_points_details: (None|Some\(FdoAdditionalPointsDetails \{(.*?)\}\))", fs)
    if m.group(2):
        d = m.group(2)
        f["additional"] = {k: (v == "true" if v in ("true", "false") else int(v)) for k, v in re.findall(r"(\w+): (true|false|-?\d+)", d)}
    else:
        f["additional"] = None
    anns = []
    am = re.search(r"announcements: heapless::Vec::from_slice\(&\[(.*?)\]\)\.unwrap\(\), player_eyes", line)
    for m in re.finditer(r"card_index: (\d+), player: FdoPlayer::(\w+), announcement: (?:FdoAnnouncement::)?(\w+)", am.group(1)):
        anns.append({"card_index": int(m.group(1)), "player": PLAYERS[m.group(2)], "level": ANN[m.group(3)]})
    f["announcements"] = anns
    for key in ("re_lowest_announcement", "contra_lowest_announcement"):
        m = re.search(key + r": (None|Some\((?:FdoAnnouncement::)?(\w+)\))", line)
        f[key] = ANN[m.group(2)] if m.group(2) else 0
    return f


def main():
    lines = [l[2:].strip() if l.startswith("//") else l.strip() for l in open(REF, encoding="utf-8")]
    games = []
    cur = None
    for ln, line in enumerate(lines, 1):
        if re.match(r"(pub )?fn \w+\(\) \{", line) and cur is None and ln > 525:
            cur = {"name": re.match(r"(?:pub )?fn (\w+)", line).group(1), "line": ln, "hands": [], "actions": [], "obs": []}
            continue
        if cur is None:
            continue
        if line.startswith("FdoHand::from_vec(vec![") and len(cur["hands"]) < 4:
            cur["hands"].append([CARD_ID[c.strip()] for c in re.search(r"vec!\[(.*?)\]", line).group(1).split(",")])
        elif line.startswith("FdoPlayer::") and "start" not in cur:
            cur["start"] = PLAYERS[re.match(r"FdoPlayer::(\w+)", line).group(1)]
        elif line.startswith("state.play_action("):
            cur["actions"].append(ACTIONS[re.match(r"state\.play_action\((\w+)\)", line).group(1)])
        elif line.startswith("assert_eq!(state.observation_for_current_player()"):
            cur["obs"].append(parse_obs(line))
            if "phase: FdoPhase::Finished" in line:
                cur["final"] = parse_final(line)
                cur["final_line"] = ln
                games.append(cur)
                cur = None
    for g in games:
        assert len(g["hands"]) == 4 and len(g["obs"]) == len(g["actions"]) + 1, (g["name"], len(g["obs"]), len(g["actions"]))
    json.dump({"source": "rs-full-doko/src/state/state.rs:525-1811 (commented-out tests)", "games": games}, open(OUT, "w"), indent=None, separators=(",", ":"))
    print("wrote", OUT, [(g["name"], g["line"], len(g["actions"])) for g in games])


if __name__ == "__main__":
    main()
