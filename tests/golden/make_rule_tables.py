#!/usr/bin/env python3
"""Extract the reference's table-style known-answer tests into tests/golden/rule_tables.json.

Sources (all under /root/reference/rs-full-doko/src unless noted):
  card/card_in_trick_logic.rs:226-1023        is_greater / is_smaller rows for every game type
  card/card_to_color.rs:260-530               card_to_color(card, game type) == colour
  card/card_to_eyes.rs:44-77                  eyes
  stats/win_conditions/re_won.rs:119-346      re_won truth table
  stats/win_conditions/kontra_won.rs:116-620  kontra_won truth table
  stats/basic_points/basic_winning_points.rs:287-1032, basic_draw_points.rs:177-280
  announcement/calc_announcement.rs:236-516   calc_allowed_announcements / internal_calc_allowed_annoucements
  stats/stats.rs:258-778                      four end-of-game cases from real games
  rs-doko/src/card/card_in_trick_logic.rs:116-141, rs-game-utils/src/bit_flag.rs:186-192
Run in the build container only (the reference is not on the GPU box).
"""
import json
import os
import re

REF = "/root/reference"
FD = os.path.join(REF, "rs-full-doko", "src")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rule_tables.json")

SUIT = {"♦": 0, "♥": 1, "♣": 2, "♠": 3}
RANK = {"9": 0, "10": 1, "J": 2, "Q": 3, "K": 4, "A": 5}
COLOR = {"T": 0, "♦": 1, "♥": 2, "♠": 3, "♣": 4, "Trump": 0, "Diamond": 1, "Heart": 2, "Spade": 3, "Club": 4}
CARDS = [s + r for s in ("Diamond", "Heart", "Club", "Spade") for r in ("Nine", "Ten", "Jack", "Queen", "King", "Ace")]
CARD_ID = {n: i for i, n in enumerate(CARDS)}
GT = ["Normal", "Wedding", "DiamondsSolo", "HeartsSolo", "SpadesSolo", "ClubsSolo", "TrumplessSolo", "QueensSolo", "JacksSolo"]
ANN_BIT = {"ReContra": 1, "No90": 2, "No60": 4, "No30": 8, "Black": 16, "CounterReContra": 32, "NoAnnouncement": 64}
PL = {"BOTTOM": 0, "LEFT": 1, "TOP": 2, "RIGHT": 3}


def sym_card(s):
    return SUIT[s[0]] * 6 + RANK[s[1:]]


def ann_set(txt):
    return sum(ANN_BIT[a] for a in re.findall(r"FdoAnnouncement::(\w+)", txt))


def ann_set_or_all_higher(txt):
    """FdoAnnouncementSet::from_vec(..) / ::new() / ::all_higher_than(Option) (announcement_set.rs:25-64)."""
    if "all_higher_than" not in txt:
        return ann_set(txt)
    m = re.search(r"FdoAnnouncement::(\w+)", txt)
    if not m:
        return 0
    order = ["ReContra", "No90", "No60", "No30", "Black"]
    name = m.group(1)
    if name == "CounterReContra":
        return 32
    return sum(ANN_BIT[x] for x in order[:order.index(name) + 1])


def opt_ann(txt):
    m = re.search(r"FdoAnnouncement::(\w+)", txt)
    return ANN_BIT[m.group(1)] if m else 0


def read(*p):
    """Source text with `//` comments removed (they sit between call arguments in several tests)."""
    return re.sub(r"//[^\n]*", "", open(os.path.join(*p), encoding="utf-8").read())


def split_args(s):
    """Split a Rust argument list on top-level commas."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def call_bodies(src, name):
    """Yield the argument text of every call `name(...)` (balanced parentheses)."""
    i = 0
    while True:
        i = src.find(name + "(", i)
        if i < 0:
            return
        j = i + len(name) + 1
        depth, k = 1, j
        while depth:
            if src[k] == "(":
                depth += 1
            elif src[k] == ")":
                depth -= 1
            k += 1
        yield src[j:k - 1], k
        i = k


def main():
    T = {}
    # ---- is_greater / is_smaller -------------------------------------------------------------------------------
    src = read(FD, "card", "card_in_trick_logic.rs")
    tests = src[src.index("#[cfg(test)]"):]
    gts = {}
    for m in re.finditer(r"let (\w+) = \[(.*?)\];", tests, re.S):
        gts[m.group(1)] = [GT.index(x.split("::")[-1]) for x in re.findall(r"[\w:]+", m.group(2)) if x.split("::")[-1] in GT]
    rows = []
    for m in re.finditer(r"(is_greater|is_smaller)\(\"([^\"]*)\", \"([^\"]*)\", \"([^\"]*)\", &(\w+)\);", tests):
        kind, cur, colors, cards, var = m.groups()
        rows.append({"greater": kind == "is_greater", "current": sym_card(cur), "colors": [COLOR[c] for c in colors],
                     "previous": [sym_card(c) for c in cards.split()], "game_types": gts[var]})
    T["is_greater_in_trick"] = rows
    # ---- card_to_color -----------------------------------------------------------------------------------------------
    src = read(FD, "card", "card_to_color.rs")
    T["card_to_color"] = [[CARD_ID[c], GT.index(g), COLOR[col]] for c, g, col in
                          re.findall(r"card_to_color\(FdoCard::(\w+), FdoGameType::(\w+)\)\s*(?:==|,)\s*FdoColor::(\w+)", src)]
    # ---- eyes ------------------------------------------------------------------------------------------------------------
    src = read(FD, "card", "card_to_eyes.rs")
    T["eyes"] = [[CARD_ID[c], int(e)] for c, e in re.findall(r"FdoCard::(\w+)\.eyes\(\),\s*(\d+)", src)]
    # ---- re_won / kontra_won ------------------------------------------------------------------------------------------------
    for name, fn in (("re_won", "re_won.rs"), ("kontra_won", "kontra_won.rs")):
        src = read(FD, "stats", "win_conditions", fn)
        tests = src[src.index("#[test]"):]
        rows = []
        for body, end in call_bodies(tests, name):
            a = split_args(body)
            if len(a) != 5 or not a[0].isdigit():
                continue
            exp = re.match(r"\s*,\s*(true|false)", tests[end:])
            rows.append([int(a[0]), ann_set(a[1]), ann_set(a[2]), a[3] == "true", a[4] == "true", exp.group(1) == "true"])
        T[name] = rows
    # ---- basic winning / draw points ----------------------------------------------------------------------------------------
    src = read(FD, "stats", "basic_points", "basic_winning_points.rs")
    tests = src[src.index("#[cfg(test)]"):]
    rows = []
    for body, end in call_bodies(tests, "FdoBasicWinningPointsDetails::calculate"):
        a = split_args(body)
        exp = re.match(r"\s*,\s*\((-?\d+), (-?\d+), FdoBasicWinningPointsDetails \{(.*?)\}\)", tests[end:], re.S)
        det = [int(v) for v in re.findall(r"\w+: (-?\d+)", exp.group(3))]
        assert len(det) == 23
        rows.append({"args": [int(a[0]), int(a[1]), a[2] == "true", ann_set(a[3]), ann_set(a[4]), int(a[5]), int(a[6])],
                     "winner": int(exp.group(1)), "loser": int(exp.group(2)), "details": det})
    T["basic_winning_points"] = rows
    src = read(FD, "stats", "basic_points", "basic_draw_points.rs")
    tests = src[src.index("#[cfg(test)]"):] if "#[cfg(test)]" in src else src[src.index("mod tests"):]
    rows = []
    for body, end in call_bodies(tests, "FdoBasicDrawPointsDetails::calculate"):
        a = split_args(body)
        exp = re.match(r"\s*,\s*\((-?\d+), (-?\d+), FdoBasicDrawPointsDetails \{(.*?)\}\)", tests[end:], re.S)
        if not exp:
            continue
        det = [int(v) for v in re.findall(r"\w+: (-?\d+)", exp.group(3))]
        rows.append({"args": [ann_set(a[0]), ann_set(a[1]), int(a[2]), int(a[3])], "re": int(exp.group(1)), "kontra": int(exp.group(2)), "details": det})
    T["basic_draw_points"] = rows
    # ---- calc_allowed_announcements ---------------------------------------------------------------------------------------------
    src = read(FD, "announcement", "calc_announcement.rs")
    tests = src[src.index("#[cfg(test)]"):]
    rows = []
    for body, end in call_bodies(tests, "calc_allowed_announcements"):
        a = split_args(body)
        if len(a) != 5:
            continue
        ts = a[2]
        re_players = sum(1 << PL[p] for p in re.findall(r"FdoPlayer::(\w+)", ts.split("re_players")[-1])) if "re_players" in ts else 0
        if "NoWedding" in ts:
            tag, wp, si = 3, -1, 0
        elif "WeddingSolved" in ts:
            tag = 2
            wp = PL[re.search(r"wedding_player: FdoPlayer::(\w+)", ts).group(1)]
            si = int(re.search(r"solved_trick_index: (\d+)", ts).group(1))
        elif "WeddingUnsolved" in ts:
            tag, si = 1, 0
            wp = PL[re.search(r"wedding_player: FdoPlayer::(\w+)", ts).group(1)]
        else:
            continue
        exp = re.match(r"\s*,\s*(FdoAnnouncementSet::new\(\)|FdoAnnouncementSet::from_vec\(vec!\[.*?\]\))", tests[end:], re.S)
        rows.append({"player": PL[a[0].split("::")[1]], "n_cards": int(a[1]), "tag": tag, "wedding_player": wp, "solved_idx": si,
                     "re_players": re_players, "re_lowest": opt_ann(a[3]), "contra_lowest": opt_ann(a[4]), "expected": ann_set(exp.group(1))})
    T["calc_allowed_announcements"] = rows
    rows = []
    for body, end in call_bodies(tests, "internal_calc_allowed_annoucements"):
        a = split_args(body)
        if len(a) != 4 or not a[0].isdigit():
            continue
        ws = re.search(r"Some\((\d+)\)", a[2])
        en = re.search(r"Some\((\d+)\)", a[3])
        exp = re.match(r"\s*,\s*(FdoAnnouncementSet::new\(\)|FdoAnnouncementSet::from_vec\(vec!\[.*?\]\))", tests[end:], re.S)
        if not exp:
            continue
        rows.append({"n_cards": int(a[0]), "prev": ann_set_or_all_higher(a[1]), "wedding_solved": int(ws.group(1)) if ws else -1,
                     "enemy_possible": int(en.group(1)) if en else -1, "expected": ann_set(exp.group(1))})
    T["internal_calc_allowed_announcements"] = rows
    # ---- end-of-game stats from four real games --------------------------------------------------------------------------------
    src = read(FD, "stats", "stats.rs")
    tests = src[src.index("#[cfg(test)]"):]
    rows = []
    for body, end in call_bodies(tests, "FdoEndOfGameStats::calculate"):
        a = split_args(body)
        eyes = [int(x) for x in re.findall(r"\d+", a[0])]
        ntr = [int(x) for x in re.findall(r"\d+", a[1])]
        re_players = sum(1 << PL[p] for p in re.findall(r"FdoPlayer::(\w+)", a[2]))
        tricks = []
        for tm in re.finditer(r"FdoTrick::existing\(\s*FdoPlayer::(\w+),\s*vec!\[(.*?)\]", a[5], re.S):
            tricks.append({"start": PL[tm.group(1)], "cards": [CARD_ID[c] for c in re.findall(r"FdoCard::(\w+)", tm.group(2))]})
        rest = tests[end:]
        exp = rest[:rest.index("assert_eq!(actual, expected)")]
        row = {"eyes": eyes, "num_tricks": ntr, "re_players": re_players, "re_lowest": opt_ann(a[3]), "contra_lowest": opt_ann(a[4]), "tricks": tricks,
               "is_solo": re.search(r"is_solo: (\w+)", exp).group(1) == "true",
               "re_eyes": int(re.search(r"re_eyes: (\d+)", exp).group(1)), "kontra_eyes": int(re.search(r"kontra_eyes: (\d+)", exp).group(1)),
               "re_points": int(re.search(r"re_points: (-?\d+)", exp).group(1)), "kontra_points": int(re.search(r"kontra_points: (-?\d+)", exp).group(1)),
               "player_points": [int(x) for x in re.search(r"player_points: PlayerZeroOrientedArr::from_full\(\[(.*?)\]\)", exp).group(1).split(",")]}
        m = re.search(r"additional_points_details: (None|Some\(FdoAdditionalPointsDetails \{(.*?)\}\))", exp, re.S)
        row["additional"] = None if not m.group(2) else {k: (v == "true" if v in ("true", "false") else int(v)) for k, v in re.findall(r"(\w+): (true|false|-?\d+)", m.group(2))}
        rows.append(row)
    T["end_of_game_stats"] = rows
    # ---- rs-doko is_greater rows + select_by_rank -------------------------------------------------------------------------------
    src = read(REF, "rs-doko", "src", "card", "card_in_trick_logic.rs")
    DCOL = {"Trump": 0, "Heart": 1, "Spade": 2, "Club": 3}
    T["doko_is_greater"] = [[CARD_ID[a], CARD_ID[b], DCOL[c], neg == ""] for neg, a, b, c in
                            re.findall(r"assert!\((!?)is_greater_in_trick_in_normal_game\(DoCard::(\w+), DoCard::(\w+), DoColor::(\w+)\)\)", src)]
    src = read(REF, "rs-game-utils", "src", "bit_flag.rs")
    T["select_by_rank"] = [[int(v, 2), int(r), int(e, 2)] for v, r, e in re.findall(r"select_by_rank\(0b(\d+), (\d+)\), 0b(\d+)\)", src)]
    json.dump(T, open(OUT, "w"), separators=(",", ":"))
    print("wrote", OUT, {k: len(v) for k, v in T.items()})


if __name__ == "__main__":
    main()
