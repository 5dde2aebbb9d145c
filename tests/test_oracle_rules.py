"""Hand-derived known-answer tests of the oracle's rule functions, one per reference function
(disease.rs:47-71, 131-154; citizen.rs:47-49, 221-248; interventions.rs:110-184)."""
import ctypes as C

import numpy as np

from epidemicsimulator_b200 import _abi
from oracle import oracle_py
from oracle.oracle_py import default_config

S, E, I, R, V = range(5)


def dstep(kind, t, cfg=None):
    cfg = cfg or default_config()
    x = oracle_py.lib().oracle_disease_step(C.byref(cfg), kind, t)
    return x >> 16, x & 0xFFFF


def test_disease_status_execute_time_step():
    assert dstep(S, 0) == (S, 0)
    assert dstep(R, 0) == (R, 0)
    assert dstep(V, 0) == (V, 0)
    assert dstep(E, 0) == (E, 1)
    assert dstep(E, 95) == (E, 96)
    assert dstep(E, 96) == (I, 0)       # exposed_time(96) <= 96
    assert dstep(I, 0) == (I, 1)
    assert dstep(I, 335) == (I, 336)
    assert dstep(I, 336) == (R, 0)      # infected_time(336) <= 336


def test_exposed_to_recovered_takes_97_plus_337_hours():
    kind, t, hours = E, 0, 0
    became_infected = None
    while kind != R:
        kind, t = dstep(kind, t)
        hours += 1
        if kind == I and became_infected is None:
            became_infected = hours
    assert became_infected == 97 and hours == 97 + 337


def test_binomial_casts_the_count_to_u8():
    L = oracle_py.lib()
    p = 0.00055
    assert L.oracle_binomial(p, 0) == 0.0
    assert L.oracle_binomial(p, 1) == 1.0 - (1.0 - p) ** 1.0
    assert L.oracle_binomial(p, 3) == 1.0 - np.power(1.0 - p, 3.0)
    assert L.oracle_binomial(p, 256) == 0.0                          # 256 as u8 == 0
    assert L.oracle_binomial(p, 257) == L.oracle_binomial(p, 1)
    assert L.oracle_binomial(p, 300) == L.oracle_binomial(p, 44)


def test_get_exposure_chance_and_inverted_mask_logic():
    L = oracle_py.lib()
    cfg = default_config()
    p, eff = 0.00055, 0.70
    ch = lambda vac, mask, ptc: L.oracle_exposure_chance(C.byref(cfg), vac, mask, ptc)
    assert ch(0, _abi.MASK_NONE, 0) == p
    assert ch(0, _abi.MASK_PUBLIC_TRANSPORT, 0) == p
    assert ch(0, _abi.MASK_PUBLIC_TRANSPORT, 1) == p - p * eff
    assert ch(0, _abi.MASK_EVERYWHERE, 0) == p - p * eff
    assert ch(1, _abi.MASK_NONE, 0) == 0.0                            # p - 1 is negative -> clamped
    # Citizen::expose: a *compliant* citizen is evaluated with MaskStatus::None, a non-compliant one with the global status
    pr = lambda compliant, mask, on_pt, n: L.oracle_expose_probability(C.byref(cfg), compliant, mask, on_pt, n)
    one = lambda c: 1.0 - (1.0 - c) ** 1.0
    assert pr(1, _abi.MASK_EVERYWHERE, 0, 1) == one(p)
    assert pr(0, _abi.MASK_EVERYWHERE, 0, 1) == one(p - p * eff)
    assert pr(1, _abi.MASK_PUBLIC_TRANSPORT, 1, 1) == one(p)          # the PublicTransport level never changes anything
    assert pr(0, _abi.MASK_PUBLIC_TRANSPORT, 1, 1) == one(p)
    assert pr(0, _abi.MASK_NONE, 0, 256) == 0.0


def upd(state, p, cfg=None):
    cfg = cfg or default_config()
    arr = (C.c_uint32 * 6)(*state)
    ev = oracle_py.lib().oracle_update_interventions(C.byref(cfg), arr, p)
    return list(arr), ev


def test_intervention_state_machine():
    # state = [lockdown_some, lockdown, vaccination_some, vaccination, mask kind, mask hours]
    st = [0, 0, 0, 0, _abi.MASK_NONE, 0]
    st, ev = upd(st, 0.0005)
    assert st == [0, 0, 0, 0, _abi.MASK_NONE, 1] and ev == 0
    st, ev = upd(st, 0.001)                      # strict '<': exactly on the threshold does nothing
    assert st == [0, 0, 0, 0, _abi.MASK_NONE, 2] and ev == 0
    st, ev = upd(st, 0.0011)
    assert st == [0, 0, 0, 0, _abi.MASK_PUBLIC_TRANSPORT, 0] and ev == 4
    st, ev = upd(st, 0.0023)                     # PT -> Everywhere, lockdown still off
    assert st == [0, 0, 0, 0, _abi.MASK_EVERYWHERE, 0] and ev == 4
    st, ev = upd(st, 0.0035)                     # lockdown starts with hour 0
    assert st == [1, 0, 0, 0, _abi.MASK_EVERYWHERE, 1] and ev == 1
    st, ev = upd(st, 0.0051)                     # vaccination latches, lockdown counts up
    assert st == [1, 1, 1, 0, _abi.MASK_EVERYWHERE, 2] and ev == 2
    st, ev = upd(st, 0.0034)                     # not above the lockdown threshold any more -> removed
    assert st[0] == 0 and st[2] == 1 and st[3] == 0 and ev == 0   # vaccination counter only advances above its threshold
    st, ev = upd(st, 0.0021)                     # Everywhere -> PublicTransport
    assert st[4:] == [_abi.MASK_PUBLIC_TRANSPORT, 0] and ev == 4
    st, ev = upd(st, 0.0009)                     # PublicTransport -> None
    assert st[4:] == [_abi.MASK_NONE, 0] and st[2] == 1
    st, ev = upd(st, 0.5)                        # None can only go to PublicTransport in one update
    assert st[4:] == [_abi.MASK_PUBLIC_TRANSPORT, 0] and st[0] == 1 and st[3] == 1


def test_default_config_matches_reference_constants():
    c = default_config()
    assert (c.exposure_chance, c.mask_effectiveness) == (0.00055, 0.70)
    assert (c.exposed_time, c.infected_time, c.max_time_step, c.vaccination_rate, c.bus_capacity) == (96, 336, 5000, 1530, 20)
    assert (c.lockdown_threshold, c.vaccination_threshold, c.mask_pt_threshold, c.mask_everywhere_threshold) == (0.0034, 0.005, 0.001, 0.0022)
