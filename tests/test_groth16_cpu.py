"""CPU tier: Groth16 oracle self-consistency — big-integer prover vs C prover on a tiny circuit,
pairing verification (accept / reject), and the synthetic MPN-like constraint generator."""
import numpy as np
import pytest

from oracle import groth16_c as GC
from oracle.py import curve as C, field as Fd, groth16 as G
from conftest import fr_arr, fr_ints


def tiny_circuit():
    """x = w^3 + w + 5 with public x: the shape of the reference's gadget tests (tiny circuit,
    setup -> prove -> verify; /root/reference/src/zk/groth16/gadgets/common/test.rs:46-63)."""
    cs = G.R1CS(num_inputs=2, num_aux=3)
    X, W, U, Y = 1, 2, 3, 4
    cs.enforce([(W, 1)], [(W, 1)], [(U, 1)])
    cs.enforce([(U, 1)], [(W, 1)], [(Y, 1)])
    cs.enforce([(Y, 1), (W, 1), (0, 5)], [(0, 1)], [(X, 1)])
    w = 3
    z = [1, w ** 3 + w + 5, w, w * w, w ** 3]
    return cs, z


def to_csr(cs):
    mats = []
    for k in range(3):
        rp, col, val = [0], [], []
        for row in cs.rows:
            for v, c in row[k]:
                col.append(v)
                val.append(c)
            rp.append(len(col))
        mats.append((np.array(rp, np.uint64), np.array(col, np.uint32), fr_arr(val) if val else np.zeros((0, 4), np.uint64)))
    return mats


def test_bigint_groth16_accepts_and_rejects():
    cs, z = tiny_circuit()
    assert cs.is_satisfied(z)
    g = Fd.SplitMix64(42)
    tox = [g.fr() for _ in range(5)]
    pk = G.setup(cs, *tox)
    proof = G.prove(cs, pk, z, g.fr(), g.fr())
    assert G.verify(pk["vk"], z[1:2], proof)
    assert not G.verify(pk["vk"], [z[1] + 1], proof)
    assert not G.verify(pk["vk"], z[1:2], (proof[0], proof[1], C.add(C.FP, proof[2], C.G1_GEN)))
    bad = list(z)
    bad[2] = 4
    assert not cs.is_satisfied(bad)
    assert not G.verify(pk["vk"], z[1:2], G.prove(cs, pk, bad, 1, 2))
    assert len(G.zkproof_blob(proof)) == 391


def test_c_groth16_equals_bigint(cref):
    cs, z = tiny_circuit()
    g = Fd.SplitMix64(7)
    tox = [g.fr() for _ in range(5)]
    r, s = g.fr(), g.fr()
    pk = G.setup(cs, *tox)
    want = G.proof_to_bytes(G.prove(cs, pk, z, r, s))
    mats = to_csr(cs)
    cpk = GC.setup(cs.num_inputs, cs.num_aux, mats, fr_arr(tox))
    # parameters agree point by point with the big-integer generator
    assert [C.g1_from_bytes(bytes(x)) for x in cpk["h"]] == pk["h"]
    assert [C.g1_from_bytes(bytes(x)) for x in cpk["l"]] == pk["l"]
    assert [C.g1_from_bytes(bytes(x)) for x in cpk["a"]] == pk["a"]
    assert [C.g1_from_bytes(bytes(x)) for x in cpk["b_g1"]] == pk["b_g1"]
    assert [C.g2_from_bytes(bytes(x)) for x in cpk["b_g2"]] == pk["b_g2"]
    assert [C.g1_from_bytes(bytes(x)) for x in cpk["vk"]["ic"]] == pk["vk"]["ic"]
    zz = fr_arr(z)
    a, b, c = GC.prove(cs.num_inputs, cs.num_aux, mats, cpk, zz[:2], zz[2:], fr_arr([r])[0], fr_arr([s])[0])
    assert bytes(GC.proof_bytes(a, b, c)) == want


def test_synthetic_circuit_is_satisfied_and_proves(cref):
    from bazuka_b200 import synth
    ni, na, mats, inputs, aux = synth.build(lanes=3, rounds=4, seed=5, ops=GC.CpuOps)
    ncons = len(mats[0][0]) - 1
    assert ncons == 3 * (12 + 255 + 1) + 1 and na == 3 * (1 + 12 + 255)
    z = np.concatenate([inputs, aux])
    ev = []
    for rp, col, val in mats:
        prod = cref.fr_mul(val, z[col])
        out = np.zeros((ncons, 4), np.uint64)
        for j in range(ncons):
            acc = np.zeros((1, 4), np.uint64)
            for k in range(int(rp[j]), int(rp[j + 1])):
                acc = cref.fr_add(acc, prod[k:k + 1])
            out[j] = acc[0]
        ev.append(out)
    assert (cref.fr_mul(ev[0], ev[1]) == ev[2]).all()
    # ~ one third of the witness is 0/1 valued, like an MPN witness
    zi = fr_ints(aux)
    assert 0.25 < sum(1 for v in zi if v in (0, 1)) / len(zi)
    tox = cref.fr_random(11, 5)
    pk = GC.setup(ni, na, mats, tox)
    r, s = cref.fr_random(12, 2)
    proof = GC.prove(ni, na, mats, pk, inputs, aux, r, s)
    assert GC.verify_py(pk["vk"], inputs[1:], proof)
    tampered = aux.copy()
    tampered[0] = cref.fr_add(tampered[0:1], tampered[1:2])[0]
    assert not GC.verify_py(pk["vk"], inputs[1:], GC.prove(ni, na, mats, pk, inputs, tampered, r, s))


def test_libbzk_host_verifier_accepts_and_rejects(cref):
    """the product's own `groth16_verify` (host pairing in libbzk, no GPU) on proofs made by the C
    oracle: accept; wrong input / tampered proof / swapped key: reject; agrees with the big-integer
    pairing; bincode entry point takes the reference's VK layout (878 + 97*len bytes)."""
    from bazuka_b200 import groth16 as BG
    cs, z = tiny_circuit()
    mats = to_csr(cs)
    pk = GC.setup(cs.num_inputs, cs.num_aux, mats, cref.fr_random(5, 5))
    zz = fr_arr(z)
    r, s = cref.fr_random(6, 2)
    proof = GC.prove(cs.num_inputs, cs.num_aux, mats, pk, zz[:2], zz[2:], r, s)
    assert BG.verify(pk["vk"], zz[1:2], proof)
    assert GC.verify_py(pk["vk"], zz[1:2], proof)
    assert not BG.verify(pk["vk"], fr_arr([z[1] + 1]), proof)
    bad = (proof[0], proof[1], cref.g1_add(proof[2], cref.g1_generator()))
    assert not BG.verify(pk["vk"], zz[1:2], bad)
    other = GC.setup(cs.num_inputs, cs.num_aux, mats, cref.fr_random(77, 5))
    assert not BG.verify(other["vk"], zz[1:2], proof)
    blob = BG.vk_to_bincode(pk["vk"])
    assert blob.size == 878 + 97 * 2
    assert BG.verify_bytes(blob, zz[1:2], GC.proof_bytes(*proof))
    assert not BG.verify_bytes(blob, fr_arr([5]), GC.proof_bytes(*proof))


def test_production_vk_blobs_parse_in_libbzk():
    """the three production keys (tests/golden/mpn_vks.json) go through the bincode entry point: a
    random 'proof' made of valid points must be rejected, not crash."""
    import json, os
    from bazuka_b200 import groth16 as BG
    from oracle import cref
    vks = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mpn_vks.json")))["vks"]
    g1, g2 = cref.g1_generator(), cref.g2_generator()
    fake = np.concatenate([g1[:97], g2[:193], g1[:97]])
    for hx in vks.values():
        blob = np.frombuffer(bytes.fromhex(hx), dtype=np.uint8)
        assert BG.verify_bytes(blob, cref.fr_random(1, 5), fake) is False


def test_generator_constants_match_oracle(cref):
    from bazuka_b200 import groth16 as BG
    assert (BG.G1_GENERATOR == cref.g1_generator()).all() and (BG.G2_GENERATOR == cref.g2_generator()).all()
