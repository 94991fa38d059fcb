"""The lane-batched TranscriptRng service (csrc/host_rng_service.h, host_keccak_lanes.cpp) must produce exactly the byte
stream of the scalar merlin TranscriptRng (host_merlin.h, itself pinned to the oracle / Merlin vectors in test_abi_and_host),
whatever the number of concurrent streams, their lengths and their arrival order.  Host only: no device needed."""
import ctypes as C
import threading

import pytest

from bulletproofs_gadgets_b200 import _lib


def _draw(lib, label, ext, warm, count, svc):
    out = C.create_string_buffer(64 * (count + 1))
    assert lib.bpg_host_rng_draw64(label, len(label), ext, warm, count, svc, out) == 0
    return out.raw


def test_lane_width_reported():
    lib = _lib.load()
    assert lib.bpg_host_rng_lanes() in (1, 4, 8)


@pytest.mark.parametrize("warm", [0, 1, 3])
@pytest.mark.parametrize("count", [0, 1, 5, 511, 512, 513, 3000])
def test_single_stream_equals_scalar(warm, count):
    lib = _lib.load()
    ext = bytes(range(32))
    assert _draw(lib, b"stream", ext, warm, count, 1) == _draw(lib, b"stream", ext, warm, count, 0)


def test_first_scalar_draw_matches_oracle_rng():
    """anchor: the scalar definition equals the Python oracle's TranscriptRng (Merlin spec restatement)"""
    from oracle import pyref
    lib = _lib.load()
    ext = bytes(range(32))
    rng = pyref.Transcript(b"anchor").build_rng([], ext)
    want = b"".join(rng.fill_bytes(64) for _ in range(4))
    assert _draw(lib, b"anchor", ext, 0, 3, 1) == want


def test_concurrent_streams_of_different_lengths():
    lib = _lib.load()
    ext = bytes(range(32, 64))
    res = {}

    def work(i):
        res[i] = _draw(lib, b"L%d" % i, ext, 1, 700 + 977 * i, 1)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(13)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for i in range(13):
        assert res[i] == _draw(lib, b"L%d" % i, ext, 1, 700 + 977 * i, 0), i


def test_stress_random_arrivals():
    """32 streams arriving at random times with lengths from 1 to 20 000 draws, three waves: leaders take over from each other
    (hand-back of unfinished streams), late streams are adopted into free lanes, extra leaders appear when all lanes are busy"""
    import random
    import time
    lib = _lib.load()
    ext = bytes(range(64, 96))
    rnd = random.Random(99)
    for wave in range(3):
        jobs = [(b"S%d-%d" % (wave, i), rnd.choice([1, 7, 300, 513, 4000, 20000]), rnd.random() * 0.02) for i in range(32)]
        res = {}

        def work(k):
            label, count, delay = jobs[k]
            time.sleep(delay)
            res[k] = _draw(lib, label, ext, 1, count, 1)

        ths = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
        for t in ths:
            t.start()
        for t in ths:
            t.join(timeout=120)
            assert not t.is_alive(), "a stream was never served"
        for k, (label, count, _) in enumerate(jobs):
            assert res[k] == _draw(lib, label, ext, 1, count, 0), (wave, k)
