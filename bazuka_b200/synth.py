"""Synthetic MPN-like constraint systems + witnesses for benchmarks and parity tests.

NOT the MPN circuits (restating `UpdateCircuit::synthesize`, /root/reference/src/mpn/circuits/
update_circuit.rs:49-494, is SURVEY.md §8 row C1 and still to come): a generator of R1CS instances
with the same *statistics* the MPN update circuit has (SURVEY.md §8a census) —
  * long chains of x^5 S-box constraints with additive round constants (the Poseidon gadget is
    ~80 % of an update transaction's constraints, /root/reference/src/zk/groth16/gadgets/poseidon/mod.rs:27-56),
  * boolean decompositions (`to_bits_le_strict`-style: 255 booleanity constraints + one 255-term
    packing row), which make ~30 % of the witness 0/1 valued,
  * one public output tying the lanes together.
`lanes` independent chains are evaluated in lock-step so witness generation is a few hundred
vectorised field operations, done by whatever backend `ops` provides (libbzk kernels on the GPU in
bench.py, the C oracle in CPU-side tests) — the generator itself only moves data.
"""
import numpy as np

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_R = (1 << 256) % R_MOD


def mont(x: int) -> np.ndarray:
    v = (x % R_MOD) * _R % R_MOD
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def _splitmix(seed):
    s = seed & 0xFFFFFFFFFFFFFFFF
    while True:
        s = (s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        yield z ^ (z >> 31)


def _fr_stream(seed):
    g = _splitmix(seed)
    while True:
        yield sum(next(g) << (64 * i) for i in range(4)) % R_MOD


class GpuOps:
    """field-array backend on libbzk kernels (numpy [n,4] in / out)."""

    def __init__(self, ctx):
        import torch
        self.ctx, self.t = ctx, torch

    def _bin(self, op, a, b):
        t = self.t
        da = t.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
        db = t.from_numpy(np.ascontiguousarray(b).view(np.int64)).cuda()
        o = t.empty_like(da)
        self.ctx.fr_binop_dev(op, da, db, o, da.shape[0])
        self.ctx.synchronize()
        return o.cpu().numpy().view(np.uint64)

    def add(self, a, b):
        return self._bin(0, a, b)

    def mul(self, a, b):
        return self._bin(2, a, b)


def build(lanes: int, rounds: int, seed: int, ops, bits: int = 255):
    """-> (num_inputs, num_aux, (A, B, C) CSR triples, inputs [2,4], aux [num_aux,4]).

    Variables: inputs = [ONE, out];  aux per lane: s0, then per round (t1, t2, s_next), then
    `bits` booleans decomposing the lane's final state.
    Constraints per lane: rounds x { (s+c)(s+c)=t1, t1*t1=t2, t2*(s+c)=s_next },
    bits x { b(b-1)=0 }, 1 x { (sum 2^i b_i) * 1 = s_final };  plus 1 x { (sum_l s_final_l) * 1 = out }."""
    L, T = int(lanes), int(rounds)
    fs = _fr_stream(seed)
    rc = [next(fs) for _ in range(T)]
    one = mont(1)
    per_lane_aux = 1 + 3 * T + bits
    num_aux = L * per_lane_aux
    num_inputs = 2
    base = num_inputs + np.arange(L, dtype=np.int64) * per_lane_aux  # z-index of lane's s0

    # ---------------- witness (vectorised over lanes)
    s = np.stack([mont(next(fs)) for _ in range(L)])
    aux = np.zeros((L, per_lane_aux, 4), dtype=np.uint64)
    aux[:, 0] = s
    for t in range(T):
        u = ops.add(s, np.repeat(mont(rc[t]).reshape(1, 4), L, axis=0))
        t1 = ops.mul(u, u)
        t2 = ops.mul(t1, t1)
        s = ops.mul(t2, u)
        aux[:, 1 + 3 * t], aux[:, 2 + 3 * t], aux[:, 3 + 3 * t] = t1, t2, s
    # canonical value of the final state -> bits
    canon = ops.mul(s, np.repeat(np.array([[1, 0, 0, 0]], dtype=np.uint64), L, axis=0))  # from_mont
    for i in range(bits):
        bit = (canon[:, i // 64] >> np.uint64(i % 64)) & np.uint64(1)
        aux[:, 1 + 3 * T + i] = np.where(bit[:, None] == 1, one[None, :], np.uint64(0))
    if bits < 255:
        # only the low `bits` bits are decomposed: constrain against the truncated value instead
        raise ValueError("bits must be 255 (full decomposition)")
    total = s[0:1]
    for l in range(1, L):
        total = ops.add(total, s[l:l + 1])
    inputs = np.stack([one, total[0]])

    # ---------------- constraints (COO per matrix, then CSR)
    ncons = L * (3 * T + bits + 1) + 1
    coo = [([], [], []) for _ in range(3)]  # rows, cols, vals(python ints)

    def put(k, rows, cols, val_int):
        coo[k][0].append(np.asarray(rows, dtype=np.int64))
        coo[k][1].append(np.asarray(cols, dtype=np.int64))
        coo[k][2].append(np.broadcast_to(mont(val_int), (len(np.atleast_1d(rows)), 4)))

    lane_row0 = np.arange(L, dtype=np.int64) * (3 * T + bits + 1)
    zero_col = np.zeros(L, dtype=np.int64)
    for t in range(T):
        r0 = lane_row0 + 3 * t
        s_prev = base + (0 if t == 0 else 3 * t)          # s0 or previous s_next
        t1c, t2c, snc = base + 1 + 3 * t, base + 2 + 3 * t, base + 3 + 3 * t
        # (s + c)(s + c) = t1
        for k in (0, 1):
            put(k, r0, s_prev, 1); put(k, r0, zero_col, rc[t])
        put(2, r0, t1c, 1)
        # t1 * t1 = t2
        put(0, r0 + 1, t1c, 1); put(1, r0 + 1, t1c, 1); put(2, r0 + 1, t2c, 1)
        # t2 * (s + c) = s_next
        put(0, r0 + 2, t2c, 1); put(1, r0 + 2, s_prev, 1); put(1, r0 + 2, zero_col, rc[t]); put(2, r0 + 2, snc, 1)
    s_fin = base + 3 * T
    for i in range(bits):
        r0 = lane_row0 + 3 * T + i
        bc = base + 1 + 3 * T + i
        # b * (b - 1) = 0
        put(0, r0, bc, 1); put(1, r0, bc, 1); put(1, r0, zero_col, R_MOD - 1)
        # packing row: sum 2^i b_i
        put(0, lane_row0 + 3 * T + bits, bc, 1 << i)
    put(1, lane_row0 + 3 * T + bits, zero_col, 1)
    put(2, lane_row0 + 3 * T + bits, s_fin, 1)
    last = np.full(L, ncons - 1, dtype=np.int64)
    put(0, last, s_fin, 1)
    put(1, [ncons - 1], [0], 1)
    put(2, [ncons - 1], [1], 1)

    mats = []
    for rows, cols, vals in coo:
        rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
        order = np.argsort(rows, kind="stable")
        rp = np.zeros(ncons + 1, dtype=np.uint64)
        np.cumsum(np.bincount(rows, minlength=ncons), out=rp[1:])
        mats.append((rp, cols[order].astype(np.uint32), np.ascontiguousarray(vals[order])))
    return num_inputs, num_aux, mats, inputs, aux.reshape(-1, 4)
