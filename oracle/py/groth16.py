"""ORACLE (test infrastructure only — never imported by the product path).

Groth16 over BLS12-381 exactly as bellman 0.14.0 implements it (`groth16::generator`,
`groth16::prover`, `groth16::verifier`; bellman is an un-vendored crates.io dependency,
/root/reference/Cargo.toml:27).  Reference call sites:
  setup   `generate_random_parameters`  /root/reference/src/config/blockchain.rs:372-400, every gadget test
  prove   `create_random_proof`         /root/reference/src/mpn/circuits/test.rs:135,175,215
  verify  `prepare_verifying_key` + `verify_proof`  /root/reference/src/zk/groth16/mod.rs:97-120
PARITY STATUS: the reference pins no proof bytes (OsRng everywhere) — "parity
unpinned"; this restatement is pinned by the verifier equation (a proof made here must verify, a
tampered one must not) and is the arbiter for the C oracle and the GPU prover at fixed
(parameters, r, s, witness).

Conventions restated from bellman:
  * variables: Input(0) = ONE, Input(1..), Aux(0..); z = inputs ++ aux.
  * both generator and prover append one constraint  Input(i) * 0 = 0  per public input.
  * domain size m = next power of two >= #constraints (after the appended ones).
  * h has m-1 bases  g1 * (tau^i * Z(tau) / delta); l[j] = g1 * (beta A_j + alpha B_j + C_j)(tau) / delta
    for aux j; ic[i] likewise / gamma for inputs; a, b_g1, b_g2 = g * A_i(tau), g * B_i(tau) with
    identity entries filtered out (which is what the prover's density trackers skip).
  * prover: a,b,c evaluations -> ifft, coset_fft, a*b-c, /Z on the coset, icoset_fft, first m-1
    coefficients = h scalars; A = alpha + r delta + sum z_i a_i; B = beta + s delta + sum z_i b_i;
    C = (sum z_i a_i) s + (sum z_i b1_i) r + r s delta + s alpha + r beta + h + l.
"""
from . import curve as C
from . import ntt as N
from .field import R_MOD


class R1CS:
    """rows of (A, B, C); each row side is a list of (var_index, coeff) with var_index into z."""

    def __init__(self, num_inputs, num_aux):
        self.num_inputs = num_inputs  # including ONE
        self.num_aux = num_aux
        self.rows = []

    def enforce(self, a, b, c):
        self.rows.append((list(a), list(b), list(c)))

    def with_input_constraints(self):
        """the constraint list bellman's generator/prover actually work on."""
        rows = list(self.rows)
        for i in range(self.num_inputs):
            rows.append(([(i, 1)], [], []))
        return rows

    def is_satisfied(self, z):
        def ev(lc):
            return sum(c * z[v] for v, c in lc) % R_MOD
        return all(ev(a) * ev(b) % R_MOD == ev(c) for a, b, c in self.rows)


def domain_log(n_constraints):
    m, e = 1, 0
    while m < n_constraints:
        m, e = m * 2, e + 1
    return e


def setup(r1cs, tau, alpha, beta, gamma, delta, g1=C.G1_GEN, g2=C.G2_GEN):
    rows = r1cs.with_input_constraints()
    log_m = domain_log(len(rows))
    m = 1 << log_m
    nv = r1cs.num_inputs + r1cs.num_aux
    # Lagrange basis at tau: L_j(tau) = ifft(powers of tau)[j]
    lag = N.ifft([pow(tau, i, R_MOD) for i in range(m)], log_m)
    At, Bt, Ct = [0] * nv, [0] * nv, [0] * nv
    for j, (a, b, c) in enumerate(rows):
        for v, co in a:
            At[v] = (At[v] + co * lag[j]) % R_MOD
        for v, co in b:
            Bt[v] = (Bt[v] + co * lag[j]) % R_MOD
        for v, co in c:
            Ct[v] = (Ct[v] + co * lag[j]) % R_MOD
    zt = (pow(tau, m, R_MOD) - 1) % R_MOD
    dinv, ginv = pow(delta, -1, R_MOD), pow(gamma, -1, R_MOD)
    G1 = lambda k: C.mul(C.FP, g1, k % R_MOD)
    G2 = lambda k: C.mul(C.FP2, g2, k % R_MOD)
    h = [G1(pow(tau, i, R_MOD) * zt % R_MOD * dinv) for i in range(m - 1)]
    ext = [(beta * At[v] + alpha * Bt[v] + Ct[v]) % R_MOD for v in range(nv)]
    ic = [G1(ext[v] * ginv) for v in range(r1cs.num_inputs)]
    l = [G1(ext[v] * dinv) for v in range(r1cs.num_inputs, nv)]
    a_all = [G1(At[v]) for v in range(nv)]
    b1_all = [G1(Bt[v]) for v in range(nv)]
    b2_all = [G2(Bt[v]) for v in range(nv)]
    return {
        "log_m": log_m,
        "vk": {"alpha_g1": G1(alpha), "beta_g1": G1(beta), "beta_g2": G2(beta), "gamma_g2": G2(gamma),
               "delta_g1": G1(delta), "delta_g2": G2(delta), "ic": ic},
        "h": h, "l": l,
        # unfiltered per-variable columns plus the filtered vectors bellman stores
        "a_all": a_all, "b1_all": b1_all, "b2_all": b2_all,
        "a": [p for p in a_all if p is not None],
        "b_g1": [p for p in b1_all if p is not None],
        "b_g2": [p for p in b2_all if p is not None],
    }


def h_scalars(r1cs, z):
    rows = r1cs.with_input_constraints()
    log_m = domain_log(len(rows))
    m = 1 << log_m

    def ev(lc):
        return sum(c * z[v] for v, c in lc) % R_MOD
    a = [ev(r[0]) for r in rows] + [0] * (m - len(rows))
    b = [ev(r[1]) for r in rows] + [0] * (m - len(rows))
    c = [ev(r[2]) for r in rows] + [0] * (m - len(rows))
    a, b, c = (N.coset_fft(N.ifft(v, log_m), log_m) for v in (a, b, c))
    q = N.divide_by_z_on_coset([(x * y - w) % R_MOD for x, y, w in zip(a, b, c)], log_m)
    return N.icoset_fft(q, log_m)[: m - 1]


def prove(r1cs, params, z, r, s):
    """z = inputs ++ aux (z[0] = 1).  Returns (A in G1, B in G2, C in G1) affine or None."""
    F1, F2 = C.FP, C.FP2
    hs = h_scalars(r1cs, z)
    msm1 = lambda bases, sc: C.msm_naive(F1, bases, sc)
    vk = params["vk"]
    a_ans = msm1(params["a_all"], z)          # identity columns contribute nothing: same as density-filtered sums
    b1_ans = msm1(params["b1_all"], z)
    b2_ans = C.msm_naive(F2, params["b2_all"], z)
    h_ans = msm1(params["h"], hs)
    l_ans = msm1(params["l"], z[r1cs.num_inputs:])
    add1 = lambda p, q: C.add(F1, p, q)
    g_a = add1(add1(C.mul(F1, vk["delta_g1"], r), vk["alpha_g1"]), a_ans)
    g_b = C.add(F2, C.add(F2, C.mul(F2, vk["delta_g2"], s), vk["beta_g2"]), b2_ans)
    g_c = C.mul(F1, vk["delta_g1"], r * s % R_MOD)
    g_c = add1(g_c, C.mul(F1, vk["alpha_g1"], s))
    g_c = add1(g_c, C.mul(F1, vk["beta_g1"], r))
    g_c = add1(g_c, C.mul(F1, a_ans, s))
    g_c = add1(g_c, C.mul(F1, b1_ans, r))
    g_c = add1(add1(g_c, h_ans), l_ans)
    return g_a, g_b, g_c


def verify(vk, public_inputs, proof):
    """public_inputs excludes ONE (bellman `verify_proof(pvk, proof, &inputs)`)."""
    if len(public_inputs) + 1 != len(vk["ic"]):
        return False
    acc = vk["ic"][0]
    for x, p in zip(public_inputs, vk["ic"][1:]):
        acc = C.add(C.FP, acc, C.mul(C.FP, p, x % R_MOD))
    A, B, Cc = proof
    neg = lambda p: C.neg(C.FP, p)
    # e(A,B) * e(-acc, gamma) * e(-C, delta) * e(-alpha, beta) == 1
    return C.pairing_product_is_one([(A, B), (neg(acc), vk["gamma_g2"]), (neg(Cc), vk["delta_g2"]),
                                     (neg(vk["alpha_g1"]), vk["beta_g2"])])


def proof_to_bytes(proof):
    """387-byte bincode image of `Groth16Proof {a, b, c}` (/root/reference/src/zk/groth16/mod.rs:33-38):
    (Fp,Fp,bool) | ((Fp,Fp),(Fp,Fp),bool) | (Fp,Fp,bool), Montgomery limbs, 1-byte bools."""
    a, b, c = proof
    return C.g1_to_bytes(a)[:97] + C.g2_to_bytes(b)[:193] + C.g1_to_bytes(c)[:97]


def zkproof_blob(proof):
    """391-byte bincode of `ZkProof::Groth16(Box<Groth16Proof>)` (u32 variant tag 0)."""
    return (0).to_bytes(4, "little") + proof_to_bytes(proof)
