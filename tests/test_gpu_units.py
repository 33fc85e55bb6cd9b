"""GPU unit parity through the C ABI test hooks: Philox block, deterministic log, and Profile::predict
(indels + substitutions + qualities) on explicit draw tapes, each against the CPU oracle bit for bit."""
import ctypes as C

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from scssim_b200 import api
    g = api.GenReads(device=0)
    yield g
    g.close()


def test_philox_block_known_answers(gpu):
    assert gpu.test_philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert gpu.test_philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert gpu.test_philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_det_log_bitwise_equal_to_oracle(gpu):
    L = H.oracle_lib()
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.integers(0, 2 ** 32, 20000).astype(np.float64) / 4294967296.0, rng.random(5000), [0.0, 1.0, 2.0 ** -32]])
    got = gpu.test_det_log(xs)
    want = np.array([L.orc_det_log(float(x)) for x in xs])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("profile,layout", [("Illumina_HiSeq2500", "PE"), ("Illumina_HiSeqXTen", "PE"), ("Illumina_HiSeq2000", "SE")])
@pytest.mark.parametrize("is_read1", [True, False])
def test_predict_matches_oracle_on_explicit_tapes(profile, layout, is_read1):
    from scssim_b200 import api
    L = H.oracle_lib()
    path = H.profile_path(profile)
    err = C.create_string_buffer(256)
    op = L.orc_profile_load(path.encode(), int(layout == "PE"), 260, err, 256)
    assert op
    rng = np.random.default_rng(11)
    with api.GenReads(device=0, layout=layout) as g:
        g.load_profile(path)
        RL = g.read_length
        n, stride = 600, 4 * RL + 512
        src = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(n, RL))
        real = rng.integers(0, 2 ** 32, size=(n, stride), dtype=np.uint64).astype(np.uint32)
        ints = rng.integers(0, 2 ** 32, size=(n, stride), dtype=np.uint64).astype(np.uint32)
        # force indel events (tiny draws) at an increasing rate, including runs that hit the "< 50 bases" guard
        for i in range(n):
            rate = [0.0, 0.01, 0.04, 0.3][i % 4]
            m = rng.random(2 * RL) < rate
            real[i, :2 * RL][m] = rng.integers(0, 1000, m.sum())
        oseq, oqual, olen = g.test_predict(src, is_read1, real, ints, out_stride=384)
    seq = C.create_string_buffer(4096); qual = C.create_string_buffer(4096); used = (C.c_uint64 * 2)()
    checked = 0
    for i in range(n):
        m = L.orc_predict(op, src[i].tobytes(), RL, int(is_read1), real[i].ctypes.data, stride, ints[i].ctypes.data, stride, seq, qual, used)
        if m < 0 or m > 384 or olen[i] == -2:   # tape ran dry / past the kernel's 384-base cap (-1) / > 32 indel events (-2)
            assert olen[i] < 0 or m < 0
            continue
        assert olen[i] == m, (i, olen[i], m)
        assert oseq[i, :m].tobytes() == seq.raw[:m], i
        assert oqual[i, :m].tobytes() == qual.raw[:m], i
        checked += 1
    assert checked > n // 2
    L.orc_profile_free(op)
