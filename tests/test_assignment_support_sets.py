"""rs-doko-assignment's own tests: 1000 samples, de-duplicated, must equal the enumerated set of 9 resp. 12 assignments
(rs-doko-assignment/src/assignment.rs:908-1109; RNG independent).  Checked on the oracle and on the device logic."""
import ctypes as C

import numpy as np
import pytest

import hostsim_lib
from assignment_cases import CASES, canonical, case_record, expected_set


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_support_set_oracle_and_device(orc, case):
    sim = hostsim_lib.load()
    rec = case_record(case)
    tricks = np.full(60, -1, dtype=np.int32)
    for t, tr in enumerate(case["tricks"]):
        tricks[t * 5:t * 5 + len(tr["cards"])] = tr["cards"]
        tricks[t * 5 + 4] = tr["start"]
    from oracle_lib import hand_from_cards

    seen_o, seen_d = set(), set()
    for s in range(1000):
        ho, hd = (C.c_uint64 * 4)(), (C.c_uint64 * 4)()
        st = orc.orc_doko_sample_assignment_raw(case["marriage"], len(case["tricks"]), tricks.ctypes.data_as(C.c_void_p), C.c_uint64(hand_from_cards(case["hand"])),
                                                (C.c_uint32 * 4)(*case["lens"]), case["observer"], C.c_uint64(42), C.c_uint64(0), s, 0, ho)
        assert st == 0
        assert sim.sim_doko_assign(hostsim_lib.ptr(rec), 42, 0, s, 0, hd) == 0
        assert list(ho) == list(hd)
        seen_o.add(canonical([int(x) for x in ho]))
        seen_d.add(canonical([int(x) for x in hd]))
    assert len(seen_o) == case["n_unique"]
    # the reference asserts `len == n` and `contains` for each listed assignment (one list repeats an entry), i.e. a superset check
    assert expected_set(case) <= seen_o and seen_d == seen_o
