"""CPU tier: csrc/pairing.cuh (the verifier's pairing, compiled for the host) against the big-integer oracle.

The product computes e(P,Q)^3 — Miller loop with Jacobian line steps scaled by Fp2 factors, final exponentiation through
3 (p^4-p^2+1)/r = (x-1)^2 (x+p) (x^2+p^2-1) + 3 — so: its value equals the oracle's pairing cubed; its field inversion and
Frobenius are what they claim; and the prepared / batched verifier entry points accept and reject like the oracle's
`verify` (/root/reference/src/zk/groth16/mod.rs:67-121 is the function replaced)."""
import ctypes as ct
import time

import numpy as np
import pytest

from conftest import fr_arr
from test_groth16_cpu import tiny_circuit, to_csr

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
RINV = pow(1 << 384, -1, P)


def _mont(x):
    return np.frombuffer(((x << 384) % P).to_bytes(48, "little"), dtype=np.uint32)


def _from_mont(words):
    return int.from_bytes(np.ascontiguousarray(words, dtype=np.uint32).tobytes(), "little") * RINV % P


def _f12_to_oracle(words):
    """6 Fp2 coefficients of w^k (u = w^6 - 1) -> the oracle's degree-12 polynomial in w"""
    v = [0] * 12
    for k in range(6):
        a, b = _from_mont(words[24 * k:24 * k + 12]), _from_mont(words[24 * k + 12:24 * k + 24])
        v[k] = (v[k] + a - b) % P
        v[k + 6] = (v[k + 6] + b) % P
    return v


def _f12_from_oracle(v):
    out = np.zeros(144, dtype=np.uint32)
    for k in range(6):
        b = v[k + 6]
        a = (v[k] + b) % P
        out[24 * k:24 * k + 12] = _mont(a)
        out[24 * k + 12:24 * k + 24] = _mont(b)
    return out


def _g1_words(pt):
    return np.concatenate([_mont(pt[0]), _mont(pt[1])])


def _g2_words(pt):
    return np.concatenate([_mont(pt[0][0]), _mont(pt[0][1]), _mont(pt[1][0]), _mont(pt[1][1])])


def _p(a):
    return a.ctypes.data_as(ct.c_void_p)


def test_pairing_value_is_the_oracle_pairing_cubed(hostshim):
    from oracle.py import curve as C
    for a, b in ((1, 1), (5, 7), (0x1234567, 0xABCDEF0123)):
        p1, q2 = C.mul(C.FP, C.G1_GEN, a), C.mul(C.FP2, C.G2_GEN, b)
        out = np.zeros(144, dtype=np.uint32)
        g1, g2 = _g1_words(p1), _g2_words(q2)
        hostshim.shim_pairing_cubed(_p(g1), _p(g2), _p(out))
        want = C.f12_pow(C.pairing(q2, p1), 3)
        assert _f12_to_oracle(out) == want, (a, b)
    # bilinearity on the product's own values: e(aP, bQ)^3 == (e(P,Q)^3)^(ab)
    base = np.zeros(144, dtype=np.uint32)
    hostshim.shim_pairing_cubed(_p(_g1_words(C.G1_GEN)), _p(_g2_words(C.G2_GEN)), _p(base))
    assert C.f12_pow(_f12_to_oracle(base), 35) == want or True  # (5*7 case checked below explicitly)
    out = np.zeros(144, dtype=np.uint32)
    hostshim.shim_pairing_cubed(_p(_g1_words(C.mul(C.FP, C.G1_GEN, 5))), _p(_g2_words(C.mul(C.FP2, C.G2_GEN, 7))), _p(out))
    assert _f12_to_oracle(out) == C.f12_pow(_f12_to_oracle(base), 35)


def test_fp12_inverse_frobenius_and_final_exponentiation(hostshim):
    import random
    from oracle.py import curve as C
    rnd = random.Random(3)
    v = [rnd.randrange(P) for _ in range(12)]
    words = _f12_from_oracle(v)
    assert _f12_to_oracle(words) == v
    inv, frob, fe = (np.zeros(144, dtype=np.uint32) for _ in range(3))
    hostshim.shim_f12_inv_frob(_p(words), _p(inv), _p(frob))
    assert C.f12_mul(_f12_to_oracle(inv), v) == C.F12_ONE
    assert _f12_to_oracle(frob) == C.f12_pow(v, P)
    hostshim.shim_final_exp(_p(words), _p(fe))
    assert _f12_to_oracle(fe) == C.f12_pow(C.final_exponentiation(v), 3)


def _proofs(n_proofs):
    from oracle import groth16_c as GC
    from oracle.py import field as Fd
    cs, z = tiny_circuit()
    mats = to_csr(cs)
    g = Fd.SplitMix64(11)
    pk = GC.setup(cs.num_inputs, cs.num_aux, mats, fr_arr([g.fr() for _ in range(5)]))
    zz = fr_arr(z)
    out = []
    for _ in range(n_proofs):
        r, s = fr_arr([g.fr()])[0], fr_arr([g.fr()])[0]
        out.append(GC.proof_bytes(*GC.prove(cs.num_inputs, cs.num_aux, mats, pk, zz[:2], zz[2:], r, s)))
    return pk, zz[1:2], np.stack(out)


def test_prepared_key_and_batch_verifier(cref):
    from bazuka_b200 import groth16 as BG
    pk, pub, proofs = _proofs(7)
    pvk = BG.PreparedVerifyingKey(pk["vk"])
    assert all(pvk.verify(pub, p) for p in proofs)
    pubs = np.repeat(pub[None], len(proofs), axis=0)
    ok, each = pvk.verify_batch(pubs, proofs, seed=12345, threads=3)
    assert ok and each.all()
    # one bad proof (C replaced by another proof's A: still a curve point) and one wrong public input
    bad = proofs.copy()
    bad[2, 290:387] = proofs[3, 0:97]
    ok, each = pvk.verify_batch(pubs, bad, seed=999, threads=2)
    assert not ok and each.tolist() == [True, True, False, True, True, True, True]
    wrong = pubs.copy()
    wrong[5, 0] = fr_arr([123])[0]
    ok, each = pvk.verify_batch(wrong, proofs, seed=7)
    assert not ok and each.tolist() == [True] * 5 + [False, True]
    # not a curve point at all
    junk = proofs.copy()
    junk[0, 5] ^= 1
    ok, each = pvk.verify_batch(pubs, junk, seed=1)
    assert not ok and not each[0] and each[1:].all()
    assert pvk.verify_batch(pubs[:0], proofs[:0])[0]
    pvk.free()


def test_verifier_is_milliseconds(cref):
    """the round-1 verifier took 65 ms (slower than the CPU path it replaces); bound the new one loosely so that a
    regression to that regime fails even on a loaded CI core."""
    from bazuka_b200 import groth16 as BG
    pk, pub, proofs = _proofs(1)
    blob = BG.vk_to_bincode(pk["vk"])
    assert BG.verify_bytes(blob, pub, proofs[0])         # prepares + caches the key
    t0 = time.perf_counter()
    for _ in range(5):
        assert BG.verify_bytes(blob, pub, proofs[0])
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"verify_bytes (cached prepared key): {ms:.2f} ms")
    assert ms < 25
