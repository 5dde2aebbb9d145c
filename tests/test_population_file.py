"""The binary population file (include/esim_popgen.h, csrc/population_io.cpp): round trip of every array, the optional
tables, sharded populations, and rejection of truncated / corrupted / inconsistent files."""
import os

import numpy as np
import pytest

from epidemicsimulator_b200 import SimError, load_population, save_population, shard_population, synthetic_population

ARRAYS = ("home_bldg", "work_bldg", "room", "age", "occupation", "flags", "status", "timer", "bldg_area", "bldg_type", "room_bldg")


@pytest.fixture(scope="module")
def pop():
    return synthetic_population(n_areas=40, areas_per_school=10, cross_area_fraction=0.4)


def same(a, b):
    assert a.n_areas == b.n_areas and a.n_citizens == b.n_citizens and a.n_buildings == b.n_buildings and a.n_rooms == b.n_rooms
    for name in ARRAYS:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_round_trip_with_offsets_and_codes(pop, tmp_path):
    path = tmp_path / "pop.esimpop"
    codes = ["E%08d" % (100 + 7 * a) for a in range(pop.n_areas)]   # OutputAreaID::code (output_area.rs:42-45)
    save_population(pop, path, area_codes=codes)
    back, got_codes = load_population(path)
    same(pop, back)
    assert np.array_equal(back.area_offsets, pop.area_offsets)
    assert got_codes == codes
    assert os.path.getsize(path) % 8 == 0


def test_round_trip_without_optional_tables(pop, tmp_path):
    path = tmp_path / "bare.esimpop"
    bare = pop.copy()
    bare.area_offsets = None
    save_population(bare, path)
    back, codes = load_population(path)
    same(pop, back)
    assert back.area_offsets is None and codes is None


def test_shard_round_trip_keeps_the_shard_header(pop, tmp_path):
    shard = shard_population(pop, 1, 2)
    path = tmp_path / "shard.esimpop"
    save_population(shard, path)
    back, _ = load_population(path)
    same(shard, back)
    assert back.n_global_citizens == pop.n_citizens and back.n_shards == 2
    assert back.n_shared_bldgs == shard.n_shared_bldgs and back.n_shared_rooms == shard.n_shared_rooms
    assert np.array_equal(back.global_id, shard.global_id)


def test_a_reloaded_population_shards_like_the_original(pop, tmp_path):
    path = tmp_path / "pop.esimpop"
    save_population(pop, path)
    back, _ = load_population(path)
    a, b = shard_population(pop, 0, 3), shard_population(back, 0, 3)
    same(a, b)


@pytest.mark.parametrize("damage", ["truncate", "flip", "magic", "index"])
def test_damaged_files_are_rejected(pop, tmp_path, damage):
    path = tmp_path / "pop.esimpop"
    save_population(pop, path)
    blob = bytearray(path.read_bytes())
    if damage == "truncate":
        blob = blob[:-72]
    elif damage == "flip":
        blob[128 + 4 * 17] ^= 0x40            # inside home_bldg: caught by the checksum
    elif damage == "magic":
        blob[0] = ord("X")
    else:
        # a building index past the table, with the checksum recomputed: caught by the index check
        bad = np.frombuffer(bytes(blob[128:132]), np.uint32).copy()
        bad[0] = pop.n_buildings + 5
        blob[128:132] = bad.tobytes()
        h = 14695981039346656037
        for byte in blob[:-8]:
            h = ((h ^ byte) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        blob[-8:] = h.to_bytes(8, "little")
    path.write_bytes(bytes(blob))
    with pytest.raises(SimError):
        load_population(path)


def test_missing_file_is_an_io_error(tmp_path):
    with pytest.raises(SimError) as e:
        load_population(tmp_path / "nope.esimpop")
    assert e.value.code == -11


def test_file_from_an_independent_writer_loads(pop, tmp_path):
    """The layout as documented (and as integration/rust/export_b200.rs writes it: status + timer + area offsets + area codes,
    no age / occupation / global ids), produced here without the library's writer."""
    def padded(a):
        b = np.ascontiguousarray(a).tobytes()
        return b + b"\0" * ((64 - len(b) % 64) % 64)
    codes = ["E%05d" % a for a in range(pop.n_areas)]
    blob = "".join(codes).encode()
    code_off = np.cumsum([0] + [len(c) for c in codes]).astype(np.uint32)
    arrays = [pop.home_bldg, pop.work_bldg, pop.room, pop.flags, pop.status, pop.timer, pop.bldg_area, pop.bldg_type,
              pop.room_bldg, pop.area_offsets.astype(np.uint32), code_off, np.frombuffer(blob, np.uint8)]
    payload = b"".join(padded(a) for a in arrays)
    header = b"ESIMPOP\x01" + np.array([1, 128, pop.n_citizens, pop.n_areas, pop.n_buildings, pop.n_rooms, 0, 0, 0, 0,
                                         4 | 8 | 32 | 64, len(blob)], np.uint32).tobytes() + np.uint64(len(payload)).tobytes()
    header += b"\0" * (128 - len(header))
    body = header + payload
    h = 14695981039346656037
    for chunk in (body,):
        for byte in chunk:
            h = ((h ^ byte) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    path = tmp_path / "independent.esimpop"
    path.write_bytes(body + h.to_bytes(8, "little"))
    back, got = load_population(path)
    for name in ("home_bldg", "work_bldg", "room", "flags", "status", "timer", "bldg_area", "bldg_type", "room_bldg"):
        assert np.array_equal(getattr(back, name), getattr(pop, name)), name
    assert got == codes and np.array_equal(back.area_offsets, pop.area_offsets)
    assert back.age.max() == 0 and back.global_id is None      # optional arrays that were not stored
