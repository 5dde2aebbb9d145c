"""Host logic of the import that the public-transport kernel relies on (csrc/pt_spans.h through libesim_host.so): whole
consecutive routes packed into spans of at most 128 riders, and the per-rider segment word."""
import ctypes as C

import numpy as np
import pytest

from epidemicsimulator_b200 import _abi
from epidemicsimulator_b200._lib import host_lib


def pack(lengths, max_riders=128):
    lengths = np.asarray(lengths, np.uint32)
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint32)
    n_routes, n_riders = len(lengths), int(off[-1])
    spans = np.zeros(4 * max(n_routes, 1), np.uint32)
    seg = np.zeros(max(n_riders, 1), np.uint16)
    n = host_lib().esim_pt_pack_spans(off.ctypes.data_as(_abi.u32p), n_routes, max_riders, spans.ctypes.data_as(_abi.u32p),
                                      seg.ctypes.data_as(C.POINTER(C.c_uint16)))
    assert n >= 0
    return off, spans[:4 * n].reshape(n, 4), seg[:n_riders]


def check(lengths, max_riders=128):
    off, spans, seg = pack(lengths, max_riders)
    n_routes = len(lengths)
    # the spans tile the routes and the riders, in order, without gaps
    assert spans[:, 3].sum() == n_routes and spans[:, 1].sum() == off[-1]
    route = rider = 0
    for first_rider, riders, first_route, routes in spans.tolist():
        assert first_route == route and first_rider == rider == off[first_route] and routes >= 1
        assert riders == off[first_route + routes] - off[first_route]
        assert riders <= max_riders or routes == 1                    # only a single over-long route exceeds the limit
        if first_route + routes < n_routes and riders <= max_riders:  # greedy: the next route would not have fitted
            assert riders + lengths[first_route + routes] > max_riders
        for q in range(first_route, first_route + routes):
            want = 0 if riders > max_riders else (off[q] - first_rider) | (lengths[q] << 8)
            assert (seg[off[q]:off[q + 1]] == want).all(), (q, want)
        route += routes
        rider += riders
    return spans


def test_small_routes_share_a_span():
    spans = check([1, 2, 1, 3, 50, 70, 9, 128, 1, 127, 2])
    assert spans[:, 3].tolist() == [6, 1, 1, 2, 1] and spans[:, 1].tolist() == [127, 9, 128, 128, 2]


def test_over_long_route_is_its_own_span():
    spans = check([5, 600, 5, 129, 128])
    assert spans[:, 1].tolist() == [5, 600, 5, 129, 128] and (spans[:, 3] == 1).all()


def test_empty_and_single():
    assert pack([])[1].shape == (0, 4)
    assert check([7])[:, 1].tolist() == [7]
    check([0, 0, 3, 0])   # empty routes cannot occur in the import, but the packing must not break on them


@pytest.mark.parametrize("seed", range(5))
def test_random_route_lengths(seed):
    rng = np.random.default_rng(seed)
    dense = rng.integers(1, 4, size=5000)                                       # cross-area workplaces: one or two riders
    census = np.where(rng.random(3000) < 0.5, rng.integers(20, 121, 3000), rng.integers(3, 20, 3000))  # (A,A) and school routes
    check(dense)
    check(census)
    check(np.concatenate([dense[:100], [400], census[:100]]), max_riders=64)


def test_invalid_arguments():
    off = np.zeros(2, np.uint32)
    out = np.zeros(4, np.uint32)
    seg = np.zeros(1, np.uint16)
    lib = host_lib()
    assert lib.esim_pt_pack_spans(off.ctypes.data_as(_abi.u32p), 1, 129, out.ctypes.data_as(_abi.u32p), seg.ctypes.data_as(C.POINTER(C.c_uint16))) < 0
    assert lib.esim_pt_pack_spans(None, 1, 128, out.ctypes.data_as(_abi.u32p), seg.ctypes.data_as(C.POINTER(C.c_uint16))) < 0
