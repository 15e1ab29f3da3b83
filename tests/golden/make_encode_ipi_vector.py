#!/usr/bin/env python3
"""Extract the reference's hand-built encode_state_ipi known-answer test into tests/golden/encode_ipi_vector.json.

Source: /root/reference/rs-doko-networks/src/full_doko/var1/encode_ipi.rs:322-889 (fn test_encode_state): the FdoState struct literal,
the assumed hands / assumed reservations / next player passed to encode_state_ipi, and the 311 `assert_eq_inc!` lines (the helper
encoders named on those lines are evaluated by make_encode_pi_vector.evaluate from their definitions).  Build container only.
"""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_encode_pi_vector as pi  # noqa: E402

REF = "/root/reference/rs-doko-networks/src/full_doko/var1/encode_ipi.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "encode_ipi_vector.json")


def main():
    src = open(REF, encoding="utf-8").read()
    body = src[src.index("fn test_encode_state()"):]
    cut = body.index("let result = encode_state_ipi(")
    lit, call = body[:cut], body[cut:body.index("let mut i = 0;")]
    state = pi.parse_state(lit)
    assumed_hands = [[pi.CARD_ID[c] for c in re.findall(r"(\w+)", h) if c in pi.CARD_ID] for h in re.findall(r"FdoHand::from_vec\(vec!\[(.*?)\]\)", call)]
    assert len(assumed_hands) == 4
    res_src = call[call.index("PlayerZeroOrientedArr::from_full(\n                ["):]
    assumed_res = []
    for tok in re.findall(r"Some\(FdoReservation::(\w+)\)|(None)", res_src):
        assumed_res.append(pi.RES.index(tok[0]) if tok[0] else None)
    assumed_res = assumed_res[:4]
    next_player = pi.PL[re.findall(r"FdoPlayer::(\w+),\s*\);", call)[-1]]
    expected = [pi.evaluate(x.strip()) for x in re.findall(r"assert_eq_inc!\(result\[i\.\.i\+1\], (.*)\);", body)]
    assert len(expected) == 311, len(expected)
    json.dump({"source": "rs-doko-networks/src/full_doko/var1/encode_ipi.rs:322-889", "state": state, "assumed_hands": assumed_hands,
               "assumed_reservations": assumed_res, "next_player": next_player, "expected": expected}, open(OUT, "w"))
    print("wrote", OUT, assumed_hands, assumed_res, next_player)


if __name__ == "__main__":
    main()
