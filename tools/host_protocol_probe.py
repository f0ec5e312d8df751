"""CPU-side timings of the native worker protocol's HOST parts at the production shape (A=15, T=3, 256-transaction update batch).
Runs over tests/hostshim/_mpn_shim.so (libbzk's host sources compiled with g++, GPU launches replaced by host stand-ins — see
tests/conftest.py::hostmpn), so the builder / batched-hash numbers here are HOST-Poseidon numbers, not GPU numbers; what the probe
shows is the size of a production work on the wire, the cost of the codec, and what recomputing the entering roots costs when
every hash is a dependent host hash (the reason the prover's `_ctx` variants batch them into 1 + A launches).
    python tools/host_protocol_probe.py > profiles/r02_host_protocol_cpu.txt"""
import ctypes as ct
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import conftest
    from bazuka_b200 import _lib
    from bazuka_b200.mpn import native as N, update as U, wire as Wr, works as Wk
    from bazuka_b200.mpn.ledger import NativeLedger
    from test_wire_cpu import _config
    host = conftest.hostmpn.__wrapped__() if hasattr(conftest.hostmpn, "__wrapped__") else None
    lib = host._l
    A, T, B = 15, 3, 4
    n_acc, n_tx = 512, 256
    t0 = time.time()
    keys = [N.eddsa_keys(b"probe%d" % i) for i in range(n_acc)]
    led = NativeLedger(host, A, T)
    for i, (pk, _) in enumerate(keys):
        led.set_account(i, U.MpnAccount(0, 0, pk, {0: U.Money(U.ZIESHA, 10 ** 12)}))
    print(f"ledger: {n_acc} accounts loaded in {time.time() - t0:.1f} s (host stand-in hashing)")
    t0 = time.time()
    txs = []
    for i in range(n_tx):
        tx = U.MpnTransaction(1, N.jj_compress(keys[i][0]), N.jj_compress(keys[(i + 1) % n_acc][0]), U.Money(U.ZIESHA, 1000), U.Money(U.ZIESHA, 10))
        tx.sign(keys[i][1])
        txs.append(tx)
    print(f"{n_tx} transfers signed in {time.time() - t0:.1f} s (Python EdDSA)")
    cfg = dict(_config(num=(0, 0, 1)), log4_tree_size=A, log4_token_tree_size=T, log4_update_batch_size=B)
    cw = Wr.Writer()
    Wr.enc_config(cw, cfg)
    w = Wr.Writer()
    w.vec([{"nonce": t.nonce, "src_pub_key": tuple(t.src_pub_key), "dst_pub_key": tuple(t.dst_pub_key), "amount": Wk._money_w(t.amount),
            "fee": Wk._money_w(t.fee), "sig": {"r": tuple(t.sig["r"]), "s": t.sig["s"]}} for t in txs], Wr.enc_mpn_tx)
    cb, ub = bytes(cw.b), bytes(w.b)
    canon = lambda v: np.frombuffer((v % N.R).to_bytes(32, "little"), dtype=np.uint64).copy()
    rw, fee, jj_d = np.array([1, 2, 3], np.uint64), canon(U.ZIESHA), canon(N.JJ_D)
    fork, buf, ln, n = ct.c_void_p(), ct.c_void_p(), ct.c_size_t(), ct.c_uint64()
    t0 = time.time()
    st = lib.bzk_mpn_prepare_works(host._h, led._h, cb, len(cb), None, 0, None, 0, ub, len(ub), ct.c_void_p(rw.ctypes.data), 1, ct.c_void_p(fee.ctypes.data),
                                   ct.byref(fork), ct.byref(buf), ct.byref(ln), ct.byref(n))
    assert st == 0 and n.value == 1
    dt = time.time() - t0
    resp = ct.string_at(buf, ln.value)
    lib.bzk_buffer_free(buf)
    print(f"bzk_mpn_prepare_works: 1 update batch of {n_tx} transactions in {dt:.2f} s with HOST hashing (on the GPU the hashing is "
          f"A + T + 4 batched launches); GetMpnWorkResponse image {len(resp)} bytes")
    work = resp[16:]                                   # u64 count, u64 id, then the work
    reps = 20
    t0 = time.time()
    for _ in range(reps):
        h = ct.c_void_p()
        assert lib.bzk_mpn_work_decode(work, len(work), ct.byref(h), None) == 0
        lib.bzk_mpn_work_free(h)
    print(f"bzk_mpn_work_decode: {(time.time() - t0) / reps * 1e3:.2f} ms per production work ({len(work)} bytes)")
    h = ct.c_void_p()
    assert lib.bzk_mpn_work_decode(work, len(work), ct.byref(h), None) == 0
    t0 = time.time()
    for _ in range(reps):
        sz = ct.c_size_t()
        lib.bzk_mpn_work_encode(h, None, 0, ct.byref(sz))
    print(f"bzk_mpn_work_encode: {(time.time() - t0) / reps * 1e3:.2f} ms")
    blob = open(_lib.PARAMS_PATH, "rb").read()
    hasher = ct.c_void_p()
    assert lib.bzk_poseidon_host_create(blob, len(blob), ct.byref(hasher)) == 0
    raws, ext = np.zeros((n_tx, 32 + 9 * T + 6 * A, 4), np.uint64), np.zeros((n_tx, 2, 4), np.uint64)
    t0 = time.time()
    assert lib.bzk_mpn_work_update_rows(h, hasher, ct.c_void_p(jj_d.ctypes.data), ct.c_void_p(fee.ctypes.data), ct.c_void_p(raws.ctypes.data),
                                        ct.c_void_p(ext.ctypes.data)) == 0
    dt = time.time() - t0
    print(f"bzk_mpn_work_update_rows (host Poseidon): {dt * 1e3:.0f} ms for {n_tx} slots = {n_tx} x (1 + {A}) hashes + {n_tx} key decompressions "
          f"-> what the prover's _ctx variant turns into 1 + {A} batched launches")
    one = np.zeros((1, 4), np.uint64)
    one[0, 0] = 1
    t0 = time.time()
    k = 2000
    inp = np.zeros((k, 4, 4), np.uint64)
    out = np.zeros((k, 4), np.uint64)
    lib.bzk_poseidon_host_hash(hasher, 4, ct.c_void_p(inp.ctypes.data), k, ct.c_void_p(out.ctypes.data))
    print(f"host Poseidon-4: {(time.time() - t0) / k * 1e6:.1f} us per hash on one core")
    addr = bytes(range(32))
    t0 = time.time()
    for _ in range(2000):
        c = np.zeros(4, np.uint64)
        lib.bzk_mpn_commitment(addr, 5, ct.c_void_p(c.ctypes.data))
    print(f"bzk_mpn_commitment: {(time.time() - t0) / 2000 * 1e6:.1f} us")
    lib.bzk_mpn_work_free(h)
    lib.bzk_mpn_state_free(fork)
    led.free()


if __name__ == "__main__":
    main()
