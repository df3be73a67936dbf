"""CPU: the nn sub-stream codec of liblbdrn_b200 (N2 of SURVEY.md 8f; host code, no GPU).  Replaces fpzip at encode.py:129 /
decode.py:113 when the package is absent.  Pinned against (a) the published PCmap value map -- the part of fpzip the decoder's
arithmetic depends on --, (b) a pure-Python restatement of the same algorithm (identical bytes), (c) exact round trips for
every precision, (d) the reference-minted nn sub-streams' VALUES (the fixtures were written through the test shim, whose
value map is the same truncation).  Byte compatibility with the real fpzip payload is unverified: the library is not
available to this build."""
import os

import numpy as np
import pytest

from conftest import GOLD, load_case, split_stream
import fpz_oracle as FO
import fpzip as shim          # oracle/shims/fpzip.py
import lbdrn_fpzip as Z


def _weights(n=3000, seed=0):
    rng = np.random.default_rng(seed)
    w = np.concatenate([rng.uniform(-0.01, 0.01, n), rng.uniform(-0.3, 0.3, n // 10), rng.standard_normal(n // 10) * 30])
    return w.astype(np.float32)


@pytest.mark.parametrize("prec", [0, 32, 24, 16, 12, 9, 8, 5, 2])
def test_round_trip_and_value_map(prec):
    w = np.concatenate([_weights(), np.float32([0.0, -0.0, 1e-38, -1e-38, 3.4e38, -3.4e38, np.inf, -np.inf, 1.0, -1.0])])
    blob = Z.compress(w, precision=prec, order="C")
    out = Z.decompress(blob, order="C")
    assert out.shape == (1, 1, 1, w.size) and out.dtype == np.float32
    got = out[0][0][0]
    want = shim.truncate(w, prec)                     # input with the low 32-prec bits cleared (PCmap inverse of forward)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # a second pass is lossless: the values are already representable
    again = Z.decompress(Z.compress(got, precision=prec))[0][0][0]
    assert np.array_equal(again.view(np.uint32), got.view(np.uint32))
    assert len(blob) < 44 + w.size * ((prec or 32) + 8) / 8


def test_pcmap_is_order_preserving_and_matches_the_published_transform():
    """PCmap<float, bits>::forward: r = ~bits(x); r >>= 32-bits; r ^= -(r >> (bits-1)) >> (33-bits); monotone in x."""
    xs = np.sort(np.concatenate([_weights(500, 3), np.float32([0.0, 1e-30, -1e-30, 5.0, -5.0])]))
    for bits in (32, 16, 8):
        m = [FO.map_forward(float(x), bits) for x in xs]
        assert all(a <= b for a, b in zip(m, m[1:])), bits
        for x, r in zip(xs, m):
            assert 0 <= r < (1 << bits)
            back = np.float32(FO.map_inverse(r, bits))
            assert back.view(np.uint32) == shim.truncate(np.float32([x]), bits).view(np.uint32)[0]


@pytest.mark.parametrize("prec", [16, 32, 8, 5])
def test_cpp_codec_equals_the_python_restatement_byte_for_byte(prec):
    w = _weights(400, seed=prec)
    assert Z.compress(w, precision=prec) == FO.compress([float(v) for v in w], prec)


def test_reference_minted_streams_decode_to_the_same_weights_and_similar_size():
    """The nn sub-streams of the reference-minted fixtures (written through the shim) hold the truncated weights; coding
    those weights with this codec round-trips them and lands near the shim's size (an adaptive range coder against zlib on
    byte planes: same order of magnitude; the real fpzip's size for this network is ~19 KB)."""
    for case in ("k5d2_train", "d3_bc256", "k3d1_u16"):
        meta, _, blob, _ = load_case(case)
        nn = split_stream(blob)[1][0][0]
        w = np.asarray(shim.decompress(nn)[0][0][0], np.float32)
        mine = Z.compress(w, precision=16)
        assert np.array_equal(Z.decompress(mine)[0][0][0].view(np.uint32), w.view(np.uint32))
        assert 0.6 * len(nn) < len(mine) < 1.4 * len(nn), (case, len(nn), len(mine))


def test_corrupt_and_foreign_streams_are_rejected():
    import lbdrn_cabi as cabi
    with pytest.raises(cabi.LbdrnError):
        Z.decompress(b"not a stream at all....")
    with pytest.raises(cabi.LbdrnError):
        Z.decompress(shim.compress(_weights(50), precision=16))          # the shim's container is not this format
    good = Z.compress(_weights(200), precision=16)
    with pytest.raises(cabi.LbdrnError):
        Z.decompress(good[: len(good) // 2])                              # truncated payload
    with pytest.raises(TypeError):
        Z.compress(np.zeros(4, np.float64))
