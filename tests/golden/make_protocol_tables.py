#!/usr/bin/env python3
"""Extract the reference's reservation / team / announcement-protocol / rs-doko legal-action tests into
tests/golden/protocol_tables.json.  Sources:
  rs-full-doko/src/reservation/reservation_winning_logic.rs:79-140, visible_reservations_logic.rs:79-174
  rs-full-doko/src/team/team_logic.rs:160-385
  rs-full-doko/src/announcement/announcement.rs:228-573
  rs-doko/src/action/allowed_actions.rs:210-312, rs-doko/src/reservation/reservation_winning_logic.rs:45-76
Run in the build container only."""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "protocol_tables.json")
CARDS = [s + r for s in ("Diamond", "Heart", "Club", "Spade") for r in ("Nine", "Ten", "Jack", "Queen", "King", "Ace")]
CARD_ID = {n: i for i, n in enumerate(CARDS)}
PL = {"BOTTOM": 0, "LEFT": 1, "TOP": 2, "RIGHT": 3, "PLAYER_BOTTOM": 0, "PLAYER_LEFT": 1, "PLAYER_TOP": 2, "PLAYER_RIGHT": 3}
RES = ["Healthy", "Wedding", "DiamondsSolo", "HeartsSolo", "SpadesSolo", "ClubsSolo", "QueensSolo", "JacksSolo", "TrumplessSolo"]
VIS = ["Wedding", "Healthy", "NotRevealed", "DiamondsSolo", "HeartsSolo", "SpadesSolo", "ClubsSolo", "QueensSolo", "JacksSolo", "TrumplessSolo", "NoneYet"]
ANN = {"ReContra": 1, "No90": 2, "No60": 4, "No30": 8, "Black": 16, "CounterReContra": 32, "NoAnnouncement": 64}


def read(*p):
    return re.sub(r"//[^\n]*", "", open(os.path.join(REF, *p), encoding="utf-8").read())


def test_fns(src):
    """(name, body) of every #[test] fn."""
    out = []
    for m in re.finditer(r"#\[test\]\s*(?:pub )?fn (\w+)\(\) \{", src):
        i = m.end()
        depth = 1
        while depth:
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        out.append((m.group(1), src[m.end():i - 1]))
    return out


def ann_set(txt):
    return sum(ANN[a] for a in re.findall(r"FdoAnnouncement::(\w+)", txt))


def opt_ann(txt):
    m = re.search(r"FdoAnnouncement::(\w+)", txt)
    return ANN[m.group(1)] if m else 0


def team_state(txt):
    re_players = sum(1 << PL[p] for p in re.findall(r"FdoPlayer::(\w+)", txt.split("re_players")[-1])) if "re_players" in txt else 0
    if "NoWedding" in txt:
        return {"tag": 3, "wedding_player": -1, "solved_idx": 0, "re_players": re_players}
    if "WeddingSolved" in txt:
        return {"tag": 2, "wedding_player": PL[re.search(r"wedding_player: FdoPlayer::(\w+)", txt).group(1)],
                "solved_idx": int(re.search(r"solved_trick_index: (\d+)", txt).group(1)), "re_players": re_players}
    if "WeddingUnsolved" in txt:
        return {"tag": 1, "wedding_player": PL[re.search(r"wedding_player: FdoPlayer::(\w+)", txt).group(1)], "solved_idx": 0, "re_players": 0}
    raise ValueError(txt)


def main():
    T = {}
    # reservation winner
    src = read("rs-full-doko", "src", "reservation", "reservation_winning_logic.rs")
    rows = []
    for name, body in test_fns(src):
        m = re.search(r"FdoReservationRound::existing\(\s*FdoPlayer::(\w+),\s*vec!\[(.*?)\]", body, re.S)
        e = re.search(r"winning_player_in_reservation_round\(&reservation_round\),\s*(FdoReservationResult::\w+(?:\(.*?\))?)", body, re.S)
        if not (m and e):
            continue
        exp = e.group(1)
        kind = 0 if "NoReservation" in exp else (1 if "Solo" in exp else 2)
        pm = re.search(r"FdoPlayer::(\w+)", exp)
        rm = re.search(r"FdoReservation::(\w+)", exp)
        rows.append({"start": PL[m.group(1)], "res": [RES.index(x) for x in re.findall(r"FdoReservation::(\w+)", m.group(2))], "kind": kind,
                     "player": PL[pm.group(1)] if pm else -1, "reservation": RES.index(rm.group(1)) if rm else -1})
    T["reservation_winner"] = rows
    # visible reservations
    src = read("rs-full-doko", "src", "reservation", "visible_reservations_logic.rs")
    rows = []
    for name, body in test_fns(src):
        m = re.search(r"FdoReservationRound::existing\(\s*FdoPlayer::(\w+),\s*vec!\[(.*?)\]", body, re.S)
        o = re.search(r"get_visible_reservations\(FdoPlayer::(\w+)\)", body)
        exp = {PL[p]: VIS.index(v) for p, v in re.findall(r"visible_reservations\[FdoPlayer::(\w+)\], FdoVisibleReservation::(\w+)", body)}
        rows.append({"start": PL[m.group(1)], "res": [RES.index(x) for x in re.findall(r"FdoReservation::(\w+)", m.group(2))], "observer": PL[o.group(1)],
                     "expected": [exp[p] for p in range(4)]})
    T["visible_reservations"] = rows
    # team resolve
    src = read("rs-full-doko", "src", "team", "team_logic.rs")
    rows = []
    for name, body in test_fns(src):
        rr = re.search(r"let reservation_results = (FdoReservationResult::\w+(?:\(.*?\))?);", body)
        if not rr:
            continue
        r = rr.group(1)
        kind = 0 if "NoReservation" in r else (1 if "Solo" in r else 2)
        pm = re.search(r"FdoPlayer::(\w+)", r)
        rm = re.search(r"FdoReservation::(\w+)", r)
        tricks = [{"start": PL[a], "cards": [CARD_ID[c] for c in re.findall(r"FdoCard::(\w+)", b)]}
                  for a, b in re.findall(r"FdoTrick::existing\(FdoPlayer::(\w+), vec!\[(.*?)\]\)", body, re.S)]
        hands = [[CARD_ID[c] for c in re.findall(r"FdoCard::(\w+)", h)] for h in re.findall(r"FdoHand::from_vec\(vec!\[(.*?)\]\)", body, re.S)][:4]
        arm = re.search(r"match team_state \{\s*FdoTeamState::(\w+) \{(.*?)\} => \{(.*?)\}\s*_ =>", body, re.S)
        tag = {"NoWedding": 3, "WeddingSolved": 2, "WeddingUnsolved": 1}[arm.group(1)]
        chk = arm.group(3)
        row = {"name": name, "rr": [kind, PL[pm.group(1)] if pm else -1, RES.index(rm.group(1)) if rm else -1], "tricks": tricks, "hands": hands, "tag": tag,
               "re_contains": [PL[p] for p in re.findall(r"re_players\.contains\(FdoPlayer::(\w+)\)", chk)],
               "re_len": int(re.search(r"re_players\.len\(\), (\d+)", chk).group(1)) if "re_players.len()" in chk else None,
               "wedding_player": PL[re.search(r"wedding_player, FdoPlayer::(\w+)", chk).group(1)] if "wedding_player, FdoPlayer" in chk else None,
               "solved_idx": int(re.search(r"solved_trick_index, (\d+)", chk).group(1)) if "solved_trick_index," in chk else None,
               "is_final": "assert!(team_state.is_final())" in body}
        rows.append(row)
    T["team_resolve"] = rows
    # announcement protocol scripts
    src = read("rs-full-doko", "src", "announcement", "announcement.rs")
    scripts = []
    for name, body in test_fns(src):
        lens = re.search(r"from_full\(\[(\d+), (\d+), (\d+), (\d+)\]\)", body)
        tsm = re.search(r"let team_state = (FdoTeamState::.*?);", body, re.S)
        ci = re.search(r"let card_index = (\d+);", body)
        if not (lens and tsm):
            continue
        steps = []
        parts = re.split(r"let result = announcements\.", body)[1:]
        for part in parts:
            call = part[:part.index(");") + 1]
            if call.startswith("start_round"):
                pm = re.search(r"FdoPlayer::(\w+)", call)
                st = {"op": 0, "player": PL[pm.group(1)], "ann": 0}
            else:
                pm = re.search(r"FdoPlayer::(\w+)", call)
                st = {"op": 1, "player": PL[pm.group(1)], "ann": opt_ann(call)}
            rest = part[len(call):]
            r = re.search(r"assert_eq!\(result, FdoAnnouncementProgressResult::(\w+)\(FdoPlayer::(\w+)\)\)", rest)
            st["result"] = [1 if r.group(1) == "RoundIsOver" else 0, PL[r.group(2)]]
            g = lambda pat: re.search(pat, rest, re.S)
            m = g(r"announcements\.announcements\.len\(\), (\d+)")
            st["n"] = int(m.group(1)) if m else None
            m = g(r"announcements\.starting_player, FdoPlayer::(\w+)")
            st["starting_player"] = PL[m.group(1)] if m else None
            m = g(r"number_of_turns_without_announcement, (\d+)")
            st["turns"] = int(m.group(1)) if m else None
            m = g(r"announcements\.re_lowest_announcement, (None|Some\(.*?\))\)")
            st["re_lowest"] = opt_ann(m.group(1)) if m else None
            m = g(r"announcements\.contra_lowest_announcement, (None|Some\(.*?\))\)")
            st["contra_lowest"] = opt_ann(m.group(1)) if m else None
            m = g(r"current_player_allowed_announcements, (FdoAnnouncementSet::new\(\)|FdoAnnouncementSet::from_vec\(\s*vec!\[.*?\]\s*\))")
            st["allowed"] = ann_set(m.group(1)) if m else None
            steps.append(st)
        scripts.append({"name": name, "lens": [int(x) for x in lens.groups()], "team": team_state(tsm.group(1)), "card_index": int(ci.group(1)) if ci else 0, "steps": steps})
    T["announcement_scripts"] = scripts
    # rs-doko legal actions
    src = read("rs-doko", "src", "action", "allowed_actions.rs")
    DCOL = {"Trump": 0, "Heart": 1, "Spade": 2, "Club": 3}
    DPH = {"Reservation": 0, "PlayCard": 1, "Finished": 2}
    rows = []
    tests = src[src.index("mod tests"):]
    for m in re.finditer(r"calculate_allowed_actions_in_normal_game\(\s*DoPhase::(\w+),\s*(None|Some\(DoColor::(\w+)\)),\s*hand_from_vec\(vec!\[(.*?)\]\)\s*\), (0|allowed_actions_from_vec\(vec!\[(.*?)\]\))\)",
                         tests, re.S):
        hand = [CARD_ID[c] for c in re.findall(r"DoCard::(\w+)", m.group(4))]
        exp = 0
        if m.group(6):
            for a in re.findall(r"DoAction::(\w+)", m.group(6)):
                exp |= (1 << 24) if a == "ReservationHealthy" else (1 << 25) if a == "ReservationWedding" else (1 << CARD_ID[a[4:]])
        rows.append({"phase": DPH[m.group(1)], "color": DCOL[m.group(3)] if m.group(3) else -1, "hand": hand, "expected": exp})
    T["doko_allowed_actions"] = rows
    json.dump(T, open(OUT, "w"), separators=(",", ":"))
    print("wrote", OUT, {k: len(v) for k, v in T.items()}, "announcement steps:", sum(len(s["steps"]) for s in scripts))


if __name__ == "__main__":
    main()
