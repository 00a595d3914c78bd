"""The oracle (oracle/oracle_annexb.c) pinned against the reference's known
answers and against the compiled reference itself (CPU only)."""
import numpy as np
import pytest

import known_answers as KA
import support as S

needs_ref = pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref/libh264_ref.so not built")


@pytest.mark.parametrize("case", KA.SCAN)
def test_scan_known_answers(case):
    hexs, exp, off = case
    b = KA.hx(hexs)
    s, e, o, _ = S.oracle_scan(b)
    assert list(zip(s.tolist(), e.tolist())) == exp and o == off
    if S.have_ref():
        rs, re, ro = S.ref_scan(b)
        assert list(zip(rs.tolist(), re.tolist())) == exp and ro == off


def test_strip_insert_known_answers():
    assert bytes(S.oracle_strip(KA.hx(KA.STRIP_IN))) == bytes(KA.hx(KA.STRIP_OUT))
    assert bytes(S.oracle_insert(KA.hx(KA.INSERT_IN))) == bytes(KA.hx(KA.INSERT_OUT))
    if S.have_ref():
        r, off = S.ref_strip(KA.hx(KA.STRIP_IN))
        assert bytes(r) == bytes(KA.hx(KA.STRIP_OUT)) and off == KA.STRIP_FAIL_OFF
        assert bytes(S.ref_insert(KA.hx(KA.INSERT_IN))) == bytes(KA.hx(KA.INSERT_OUT))


@needs_ref
def test_oracle_matches_reference_random():
    rng = np.random.default_rng(7)
    alpha = np.array([0, 0, 0, 1, 2, 3, 4, 0xFF, 0x65], np.uint8)
    for it in range(60):
        n = int(rng.integers(0, 5000))
        b = rng.choice(alpha, n) if it % 2 else S.gen_annexb(rng, 12, 1, 600)
        s, e, o, _ = S.oracle_scan(b)
        rs, re, ro = S.ref_scan(b)
        assert np.array_equal(s, rs) and np.array_equal(e, re) and o == ro
        for k in range(min(len(s), 8)):
            nal = b[int(s[k]):int(e[k])]
            assert np.array_equal(S.oracle_strip(nal), S.ref_strip(nal)[0])
        p = rng.choice(alpha, int(rng.integers(0, 3000)))
        assert np.array_equal(S.oracle_insert(p), S.ref_insert(p))


def test_insert_strip_round_trip():
    rng = np.random.default_rng(9)
    for it in range(30):
        p = rng.choice(np.array([0, 0, 1, 2, 3, 9, 0xFF], np.uint8), int(rng.integers(1, 4000)))
        p[-1] = 0x80
        esc = S.oracle_insert(p)
        assert np.array_equal(S.oracle_strip(esc), p)
        # an escaped payload never contains a start-code-like sequence
        s, _, _, _ = S.oracle_scan(esc)
        assert len(s) == 0
