"""ORACLE (test infrastructure only): Groth16 setup / prove on the C oracle for CSR constraint
systems — bellman 0.14.0 `generate_parameters` / `create_proof` restated (see oracle/py/groth16.py
for the conventions and the parity status: unpinned by the reference, pinned by the pairing check)."""
import ctypes as ct

import numpy as np

from . import cref
from .py import field as Fd


class CpuOps:
    """field-array backend for bazuka_b200.synth on the C oracle."""
    add = staticmethod(lambda a, b: cref.fr_add(a, b))
    mul = staticmethod(lambda a, b: cref.fr_mul(a, b))


def _p(a):
    return a.ctypes.data_as(ct.c_void_p)


def _csr_args(mats):
    out = []
    for rp, col, val in mats:
        out += [_p(rp), _p(col), _p(val)]
    return out


def density(num_inputs, num_aux, mats):
    nv = num_inputs + num_aux
    pres = []
    for rp, col, val in mats[:2]:
        d = np.zeros(nv, dtype=bool)
        d[col[val.any(axis=1)]] = True
        pres.append(d)
    a_idx = np.concatenate([np.arange(num_inputs), num_inputs + np.nonzero(pres[0][num_inputs:])[0]])
    return a_idx.astype(np.int64), np.nonzero(pres[1])[0].astype(np.int64)


def setup(num_inputs, num_aux, mats, toxic, threads=cref.NCPU):
    """toxic [5,4] Montgomery (tau, alpha, beta, gamma, delta) -> bellman Parameters as wire images."""
    L = cref.lib()
    mats = [(np.ascontiguousarray(rp, np.uint64), np.ascontiguousarray(col, np.uint32), np.ascontiguousarray(val, np.uint64)) for rp, col, val in mats]
    ncons = len(mats[0][0]) - 1
    nv = num_inputs + num_aux
    rows, log_m = ncons + num_inputs, 0
    while (1 << log_m) < rows:
        log_m += 1
    m = 1 << log_m
    toxic = np.ascontiguousarray(toxic, dtype=np.uint64).reshape(5, 4)
    h_k = np.zeros((max(m - 1, 1), 4), np.uint64)
    at, bt = np.zeros((nv, 4), np.uint64), np.zeros((nv, 4), np.uint64)
    eic, el = np.zeros((num_inputs, 4), np.uint64), np.zeros((max(num_aux, 1), 4), np.uint64)
    got = L.bzko_groth16_setup_scalars(ct.c_uint64(num_inputs), ct.c_uint64(num_aux), ct.c_uint64(ncons), *_csr_args(mats),
                                       _p(toxic), _p(h_k), _p(at), _p(bt), _p(eic), _p(el), ct.c_int(threads))
    assert got == log_m
    g1, g2 = cref.g1_generator(), cref.g2_generator()

    def fb1(k):
        k = np.ascontiguousarray(k, np.uint64).reshape(-1, 4)
        out = np.zeros((len(k), 104), np.uint8)
        L.bzko_g1_fixed_base_mul(_p(g1), _p(k), ct.c_size_t(len(k)), _p(out), ct.c_int(threads))
        return out

    def fb2(k):
        k = np.ascontiguousarray(k, np.uint64).reshape(-1, 4)
        out = np.zeros((len(k), 200), np.uint8)
        L.bzko_g2_fixed_base_mul(_p(g2), _p(k), ct.c_size_t(len(k)), _p(out), ct.c_int(threads))
        return out

    a_idx, b_idx = density(num_inputs, num_aux, mats)
    vk1, vk2 = fb1(toxic[[1, 2, 4]]), fb2(toxic[[2, 3, 4]])
    return {
        "log_m": log_m,
        "vk": {"alpha_g1": vk1[0], "beta_g1": vk1[1], "delta_g1": vk1[2], "beta_g2": vk2[0], "gamma_g2": vk2[1],
               "delta_g2": vk2[2], "ic": fb1(eic)},
        "h": fb1(h_k[: m - 1]), "l": fb1(el[:num_aux]),
        "a": fb1(at[a_idx]), "b_g1": fb1(bt[b_idx]), "b_g2": fb2(bt[b_idx]),
        "a_idx": a_idx, "b_idx": b_idx,
    }


def prove(num_inputs, num_aux, mats, params, inputs, aux, r, s, threads=cref.NCPU):
    """-> (a[104], b[200], c[104]) wire images, bellman create_proof order of operations."""
    L = cref.lib()
    mats = [(np.ascontiguousarray(rp, np.uint64), np.ascontiguousarray(col, np.uint32), np.ascontiguousarray(val, np.uint64)) for rp, col, val in mats]
    ncons = len(mats[0][0]) - 1
    z = np.ascontiguousarray(np.concatenate([np.asarray(inputs, np.uint64).reshape(-1, 4), np.asarray(aux, np.uint64).reshape(-1, 4)]))
    m = 1 << params["log_m"]
    h = np.zeros((max(m - 1, 1), 4), np.uint64)
    L.bzko_groth16_h(ct.c_uint64(num_inputs), ct.c_uint64(num_aux), ct.c_uint64(ncons), *_csr_args(mats), _p(z), _p(h), ct.c_int(threads))
    h_ans = cref.msm_g1(params["h"], h[: m - 1], threads)
    l_ans = cref.msm_g1(params["l"], z[num_inputs:], threads)
    a_ans = cref.msm_g1(params["a"], z[params["a_idx"]], threads)
    b1_ans = cref.msm_g1(params["b_g1"], z[params["b_idx"]], threads)
    b2_ans = cref.msm_g2(params["b_g2"], z[params["b_idx"]], threads)
    vk = params["vk"]
    oa, ob, oc = np.zeros(104, np.uint8), np.zeros(200, np.uint8), np.zeros(104, np.uint8)
    r = np.ascontiguousarray(r, np.uint64).reshape(4)
    s = np.ascontiguousarray(s, np.uint64).reshape(4)
    L.bzko_groth16_assemble(_p(vk["alpha_g1"]), _p(vk["beta_g1"]), _p(vk["beta_g2"]), _p(vk["delta_g1"]), _p(vk["delta_g2"]),
                            _p(a_ans), _p(b1_ans), _p(b2_ans), _p(h_ans), _p(l_ans), _p(r), _p(s), _p(oa), _p(ob), _p(oc))
    return oa, ob, oc


def proof_bytes(a, b, c):
    return np.concatenate([a[:97], b[:193], c[:97]])


def verify_py(vk, public_inputs_mont, proof):
    """pairing check on the big-integer oracle (slow, a few seconds)."""
    from .py import curve as C, groth16 as G
    vkp = {k: (C.g1_from_bytes(bytes(v)) if v.shape[-1] == 104 else C.g2_from_bytes(bytes(v))) for k, v in vk.items() if k != "ic"}
    vkp["ic"] = [C.g1_from_bytes(bytes(x)) for x in vk["ic"]]
    pub = [Fd.fr_from_mont_bytes(bytes(np.asarray(x, np.uint64).tobytes())) for x in public_inputs_mont]
    a, b, c = proof
    return G.verify(vkp, pub, (C.g1_from_bytes(bytes(a)), C.g2_from_bytes(bytes(b)), C.g1_from_bytes(bytes(c))))
