"""CPU tier: the oracle against the reference's own fixtures (tests/golden/, extracted by
tools/make_golden.py) and the two oracle implementations (big-int Python, C) against each other."""
import json
import os
import random

import numpy as np
import pytest

from oracle.py import curve as C, field as Fd, msm as M, ntt as N, poseidon as Ps, state as St
from conftest import fr_ints, fr_arr

G = os.path.join(os.path.dirname(__file__), "golden")


def test_poseidon_reference_kats_py():
    kats = json.load(open(f"{G}/poseidon_kats.json"))["expected_decimal"]
    for n, want in enumerate(kats, 1):
        assert Ps.poseidon(list(range(n))) == int(want), n


def test_poseidon_reference_kats_c(cref):
    kats = json.load(open(f"{G}/poseidon_kats.json"))["expected_decimal"]
    for n, want in enumerate(kats, 1):
        got = fr_ints(cref.poseidon(fr_arr(list(range(n))).reshape(1, n, 4)))[0]
        assert got == int(want), n


def test_empty_mpn_root_kat():
    want = [h for h in json.load(open(f"{G}/empty_root.json"))["hex_scalars_in_vector"] if int(h, 16)][0]
    assert St.compress_default(St.mpn_state_model(30, 1)) == int(want, 16)


def test_production_vks_decode_on_curve(cref):
    """pins the wire layout: x|y Montgomery limbs, bool flag, u64 length prefix (bincode)."""
    vks = json.load(open(f"{G}/mpn_vks.json"))["vks"]
    first = None
    for name, hx in vks.items():
        b = bytes.fromhex(hx)
        assert len(b) == 1460
        off = 0
        pts = []
        for kind in ("g1", "g1", "g2", "g2", "g1", "g2"):
            sz = 96 if kind == "g1" else 192
            img = b[off:off + sz] + bytes([b[off + sz]]) + bytes(7)
            off += sz + 1
            pts.append((kind, img))
        (cnt,) = np.frombuffer(b[off:off + 8], dtype=np.uint64)
        off += 8
        assert cnt == 6
        for _ in range(int(cnt)):
            pts.append(("g1", b[off:off + 96] + bytes([b[off + 96]]) + bytes(7)))
            off += 97
        assert off == 1460
        for kind, img in pts:
            arr = np.frombuffer(img, dtype=np.uint8)
            if kind == "g1":
                p = C.g1_from_bytes(img)
                assert p is not None and C.on_curve(C.FP, p)
                assert cref.g1_on_curve(arr)
                assert C.g1_to_bytes(p) == img
            else:
                p = C.g2_from_bytes(img)
                assert p is not None and C.on_curve(C.FP2, p)
                assert cref.g2_on_curve(arr)
        head = [img for _, img in pts[:6]]
        if first is None:
            first = head
        assert head == first  # alpha, beta, gamma, delta are shared by the three keys


def test_generators_and_pairing():
    assert C.on_curve(C.FP, C.G1_GEN) and C.on_curve(C.FP2, C.G2_GEN)
    assert C.mul(C.FP, C.G1_GEN, Fd.R_MOD) is None and C.mul(C.FP2, C.G2_GEN, Fd.R_MOD) is None
    a, b = 1234567, 7654321
    P1, Q1 = C.mul(C.FP, C.G1_GEN, a), C.mul(C.FP2, C.G2_GEN, b)
    assert C.pairing_product_is_one([(P1, Q1), (C.neg(C.FP, C.mul(C.FP, C.G1_GEN, a * b % Fd.R_MOD)), C.G2_GEN)])
    assert not C.pairing_product_is_one([(P1, Q1), (C.neg(C.FP, C.mul(C.FP, C.G1_GEN, a * b + 1)), C.G2_GEN)])


def test_c_fields_vs_bigint(cref):
    a, b = cref.fr_random(1, 300), cref.fr_random(2, 300)
    A, B = fr_ints(a), fr_ints(b)
    g = Fd.SplitMix64(1)
    assert A == [g.fr() for _ in range(300)]
    assert fr_ints(cref.fr_mul(a, b)) == [x * y % Fd.R_MOD for x, y in zip(A, B)]
    assert fr_ints(cref.fr_add(a, b)) == [(x + y) % Fd.R_MOD for x, y in zip(A, B)]
    assert fr_ints(cref.fr_sub(a, b)) == [(x - y) % Fd.R_MOD for x, y in zip(A, B)]
    assert fr_ints(cref.fr_inv(a[:20])) == [pow(x, -1, Fd.R_MOD) for x in A[:20]]
    rnd = random.Random(5)
    xs = [rnd.randrange(Fd.P_MOD) for _ in range(296)] + [0, 1, Fd.P_MOD - 1, Fd.P_MOD - 2]
    ys = xs[::-1]
    enc = lambda v: np.frombuffer(b"".join(Fd.fp_to_mont_bytes(x) for x in v), dtype=np.uint64).reshape(-1, 6)
    dec = lambda arr: [Fd.fp_from_mont_bytes(x.tobytes()) for x in arr]
    assert dec(cref.fp_mul(enc(xs), enc(ys))) == [x * y % Fd.P_MOD for x, y in zip(xs, ys)]
    assert dec(cref.fp_sub(enc(xs), enc(ys))) == [(x - y) % Fd.P_MOD for x, y in zip(xs, ys)]
    assert dec(cref.fp_add(enc(xs), enc(ys))) == [(x + y) % Fd.P_MOD for x, y in zip(xs, ys)]


def test_c_curves_and_msm_vs_bigint(cref):
    assert C.g1_from_bytes(cref.g1_generator().tobytes()) == C.G1_GEN
    assert C.g2_from_bytes(cref.g2_generator().tobytes()) == C.G2_GEN
    n = 40
    bs = cref.g1_random_bases(2, n)
    g = Fd.SplitMix64(2)
    pyb = [C.mul(C.FP, C.G1_GEN, g.fr()) for _ in range(n)]
    assert [C.g1_from_bytes(x.tobytes()) for x in bs] == pyb
    scl = fr_ints(cref.fr_random(1, n))
    scl[3], scl[4], scl[5] = 0, 1, Fd.R_MOD - 1
    sc = fr_arr(scl)
    want = C.msm_naive(C.FP, pyb, scl)
    assert M.multiexp(C.FP, pyb, scl) == want
    assert C.g1_from_bytes(cref.msm_g1(bs, sc).tobytes()) == want
    assert C.g1_from_bytes(cref.msm_g1_naive(bs, sc).tobytes()) == want
    b2 = cref.g2_random_bases(7, 8)
    g = Fd.SplitMix64(7)
    pyb2 = [C.mul(C.FP2, C.G2_GEN, g.fr()) for _ in range(8)]
    assert [C.g2_from_bytes(x.tobytes()) for x in b2] == pyb2
    assert C.g2_from_bytes(cref.msm_g2(b2, sc[:8]).tobytes()) == C.msm_naive(C.FP2, pyb2, scl[:8])


def test_c_msm_pippenger_vs_naive_medium(cref):
    n = 600  # window c = ceil(ln 600) = 7
    bs = cref.g1_random_bases(11, n)
    sc = cref.fr_random(12, n)
    assert (cref.msm_g1(bs, sc) == cref.msm_g1_naive(bs, sc)).all()


def test_c_poseidon_vs_bigint(cref):
    for ar in (1, 2, 3, 4, 5, 6, 7, 11, 16):
        inp = cref.fr_random(100 + ar, 3 * ar).reshape(3, ar, 4)
        assert fr_ints(cref.poseidon(inp)) == [Ps.poseidon(fr_ints(inp[i])) for i in range(3)], ar


def test_ntt_bigint_definition():
    g = Fd.SplitMix64(3)
    for log_n in (0, 1, 2, 5):
        a = [g.fr() for _ in range(1 << log_n)]
        assert N.fft(a, log_n) == N.dft_naive(a, log_n)
        assert N.ifft(N.fft(a, log_n), log_n) == a
        assert N.icoset_fft(N.coset_fft(a, log_n), log_n) == a


def test_c_ntt_vs_bigint(cref):
    for log_n in (0, 1, 3, 8):
        a = cref.fr_random(3 + log_n, 1 << log_n)
        A = fr_ints(a)
        assert fr_ints(cref.ntt(a, 0)) == N.fft(A, log_n)
        assert fr_ints(cref.ntt(a, 1)) == N.ifft(A, log_n)
        assert fr_ints(cref.ntt(a, 2)) == N.coset_fft(A, log_n)
        assert fr_ints(cref.ntt(a, 3)) == N.icoset_fft(A, log_n)
        assert fr_ints(cref.divide_by_z_on_coset(a)) == N.divide_by_z_on_coset(A, log_n)
    a = cref.fr_random(9, 1 << 14)
    assert (cref.ntt(cref.ntt(a, 0), 1) == a).all()
    assert (cref.ntt(cref.ntt(a, 2), 3) == a).all()
