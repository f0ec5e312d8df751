"""GPU tier (`-m gpu`): every libbzk kernel, called through the C ABI, against the CPU oracle on the
same seeded inputs — bit-exact (all of this path is integer arithmetic) — plus the reference's
Poseidon known answers and size-independent identities at the BASELINE.json sizes."""
import json
import os

import numpy as np
import pytest

from conftest import fr_ints, fr_arr

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _t():
    import torch
    return torch


def dev_u64(arr):
    t = _t()
    return t.from_numpy(np.ascontiguousarray(arr).view(np.int64)).cuda()


def host_u64(tensor):
    return tensor.cpu().numpy().view(np.uint64)


# ------------------------------------------------------------------ field arithmetic
def test_fr_binops(ctx, cref):
    t = _t()
    n = 1 << 16
    a, b = cref.fr_random(31, n), cref.fr_random(32, n)
    edge = fr_arr([0, 1, 2, (1 << 255) % (2**255), 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000000])
    a[: len(edge)] = edge
    b[: len(edge)] = edge[::-1]
    da, db = dev_u64(a), dev_u64(b)
    out = t.empty_like(da)
    for op, f in ((0, cref.fr_add), (1, cref.fr_sub), (2, cref.fr_mul)):
        ctx.fr_binop_dev(op, da, db, out, n)
        ctx.synchronize()
        assert (host_u64(out).reshape(-1, 4) == f(a, b)).all(), op


def test_fp_mul(ctx, cref):
    import random
    t = _t()
    P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
    rnd = random.Random(7)
    n = 1 << 14
    vals = [rnd.randrange(P) for _ in range(n - 4)] + [0, 1, P - 1, P - 2]
    x = np.frombuffer(b"".join(v.to_bytes(48, "little") for v in vals), dtype=np.uint64).reshape(-1, 6).copy()
    y = x[::-1].copy()
    dx, dy = dev_u64(x), dev_u64(y)
    out = t.empty_like(dx)
    ctx.fp_mul_dev(dx, dy, out, n)
    ctx.synchronize()
    assert (host_u64(out).reshape(-1, 6) == cref.fp_mul(x, y)).all()


def test_fr_random_stream(ctx, cref):
    t = _t()
    n = 5000
    d = t.empty((n, 4), dtype=t.int64, device="cuda")
    ctx.fr_random_dev(77, n, d)
    ctx.synchronize()
    assert (host_u64(d) == cref.fr_random(77, n)).all()


# ------------------------------------------------------------------ Poseidon
def test_poseidon_reference_kats(ctx):
    kats = json.load(open(f"{G}/poseidon_kats.json"))["expected_decimal"]
    for n, want in enumerate(kats, 1):
        got = fr_ints(ctx.poseidon(fr_arr(list(range(n))).reshape(1, n, 4)))[0]
        assert got == int(want), n


@pytest.mark.parametrize("arity", list(range(1, 17)))
def test_poseidon_vs_oracle(ctx, cref, arity):
    n = 3000 if arity <= 7 else 300
    inp = cref.fr_random(200 + arity, n * arity).reshape(n, arity, 4)
    inp[0] = 0
    assert (ctx.poseidon(inp) == cref.poseidon(inp)).all()


def test_poseidon_ragged_and_empty(ctx, cref):
    assert ctx.poseidon(np.zeros((0, 4, 4), dtype=np.uint64)).shape == (0, 4)
    for n in (1, 127, 129):
        inp = cref.fr_random(n, n * 4).reshape(n, 4, 4)
        assert (ctx.poseidon(inp) == cref.poseidon(inp)).all()


# ------------------------------------------------------------------ NTT
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 6, 7, 10, 13, 16])
def test_ntt_vs_oracle(ctx, cref, log_n):
    a = cref.fr_random(300 + log_n, 1 << log_n)
    for op in range(4):
        assert (ctx.ntt(a, op) == cref.ntt(a, op)).all(), (log_n, op)


def test_ntt_2_20_vs_oracle(ctx, cref):
    a = cref.fr_random(3, 1 << 20)
    assert (ctx.ntt(a, 0) == cref.ntt(a, 0)).all()
    assert (ctx.ntt(a, 3) == cref.ntt(a, 3)).all()


def test_ntt_2_24_roundtrip_and_linearity(ctx, cref):
    """BASELINE config 3 size: identities that hold for any n."""
    t = _t()
    log_n, n = 24, 1 << 24
    a = t.empty((n, 4), dtype=t.int64, device="cuda")
    ctx.fr_random_dev(3, n, a)
    ref = a.clone()
    for fwd, inv in ((0, 1), (2, 3)):
        ctx.ntt_dev(a, log_n, fwd)
        ctx.ntt_dev(a, log_n, inv)
        ctx.synchronize()
        assert t.equal(a, ref)
    # NTT(x) + NTT(y) == NTT(x + y)
    b = t.empty_like(a)
    ctx.fr_random_dev(4, n, b)
    s = t.empty_like(a)
    ctx.fr_binop_dev(0, a, b, s, n)
    ctx.ntt_dev(a, log_n, 0); ctx.ntt_dev(b, log_n, 0); ctx.ntt_dev(s, log_n, 0)
    ctx.fr_binop_dev(0, a, b, a, n)
    ctx.synchronize()
    assert t.equal(a, s)
    # spot value: output 0 of the forward transform is the plain sum of inputs — check on 2^12 prefix problem
    small = cref.fr_random(3, 1 << 12)
    assert (ctx.ntt(small, 0) == cref.ntt(small, 0)).all()


def test_groth16_h_pipeline(ctx, cref):
    t = _t()
    log_n = 12
    n = 1 << log_n
    a, b, c = (cref.fr_random(s, n) for s in (41, 42, 43))
    da, db, dc = dev_u64(a), dev_u64(b), dev_u64(c)
    ctx.groth16_h_dev(da, db, dc, log_n)
    ctx.synchronize()
    ea, eb, ec = (cref.ntt(cref.ntt(v, 1), 2) for v in (a, b, c))
    h = cref.divide_by_z_on_coset(cref.fr_sub(cref.fr_mul(ea, eb), ec))
    h = cref.ntt(h, 3)
    assert (host_u64(da).reshape(-1, 4) == h).all()


# ------------------------------------------------------------------ MSM
def test_random_bases_match_oracle(ctx, cref):
    t = _t()
    n = 300
    d = t.empty((n, 104), dtype=t.uint8, device="cuda")
    ctx.g1_random_bases_dev(2, n, d)
    ctx.synchronize()
    assert (d.cpu().numpy() == cref.g1_random_bases(2, n)).all()
    d2 = t.empty((40, 200), dtype=t.uint8, device="cuda")
    ctx.g2_random_bases_dev(7, 40, d2)
    ctx.synchronize()
    assert (d2.cpu().numpy() == cref.g2_random_bases(7, 40)).all()


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 257, 1000, 4096, 1 << 14, 1 << 16])
def test_msm_g1_vs_oracle(ctx, cref, n):
    bases = cref.g1_random_bases(2, n)
    scalars = cref.fr_random(1, n)
    assert (ctx.msm_g1(bases, scalars) == cref.msm_g1(bases, scalars)).all()


def test_msm_g1_edge_cases(ctx, cref):
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    inf = np.zeros(104, dtype=np.uint8)
    inf[96] = 1
    # empty
    out = ctx.msm_g1(np.zeros((0, 104), dtype=np.uint8), np.zeros((0, 4), dtype=np.uint64))
    assert out[96] == 1
    n = 2000
    bases = cref.g1_random_bases(5, n)
    # all-zero scalars -> identity image (x=0, y=R(1), flag)
    z = ctx.msm_g1(bases, fr_arr([0] * n))
    assert (z == cref.msm_g1(bases, fr_arr([0] * n))).all() and z[96] == 1
    # witness-like: many 0/1, small values, r-1, powers of two on window edges
    vals = fr_ints(cref.fr_random(6, n))
    for i in range(0, n, 3):
        vals[i] = i % 2
    for i in range(1, 200, 3):
        vals[i] = i
    vals[7], vals[8], vals[9], vals[10] = R - 1, 1 << 15, (1 << 16) - 1, (1 << 254) + (1 << 32)
    vals[11] = (1 << 255) % R
    sc = fr_arr(vals)
    assert (ctx.msm_g1(bases, sc) == cref.msm_g1(bases, sc)).all()
    # repeated bases (forces the P+P doubling branch in buckets), opposite bases (P-P), identities
    rep = np.repeat(bases[:1], n, axis=0)
    assert (ctx.msm_g1(rep, sc) == cref.msm_g1(rep, sc)).all()
    ones = fr_arr([1] * n)
    assert (ctx.msm_g1(rep, ones) == cref.msm_g1(rep, ones)).all()
    mixed = bases.copy()
    mixed[::5] = inf
    mixed[1::5] = bases[0]
    assert (ctx.msm_g1(mixed, sc) == cref.msm_g1(mixed, sc)).all()
    # s*P + (r-s)*P = identity
    two = np.repeat(bases[:1], 2, axis=0)
    s = vals[3]
    assert ctx.msm_g1(two, fr_arr([s, R - s]))[96] == 1


def test_msm_g1_resident_and_offsets(ctx, cref):
    n = 5000
    bases = cref.g1_random_bases(9, n)
    scalars = cref.fr_random(10, n)
    rb = ctx.g1_bases(bases, check_on_curve=True)
    assert len(rb) == n
    assert (ctx.msm_g1_resident(rb, scalars) == cref.msm_g1(bases, scalars)).all()
    assert (ctx.msm_g1_resident(rb, scalars[1000:3000], offset=1000) == cref.msm_g1(bases[1000:3000], scalars[1000:3000])).all()
    rb.free()
    bad = bases.copy()
    bad[17, 0] ^= 1
    import bazuka_b200 as B
    with pytest.raises(B.BzkError) as e:
        ctx.g1_bases(bad, check_on_curve=True)
    assert e.value.status == -4


def test_msm_g1_2_20_sharded_sum_and_linearity(ctx, cref):
    """BASELINE config 2 size (2^20): MSM over the whole vector == sum of MSMs over 4 shards
    (the multi-GPU reduction identity) and MSM(s)+MSM(t) == MSM(s+t); plus the full-size value
    against the 8-thread C oracle."""
    t = _t()
    n = 1 << 20
    d_img = t.empty((n, 104), dtype=t.uint8, device="cuda")
    ctx.g1_random_bases_dev(2, n, d_img)
    rb = ctx.g1_bases_from_dev(d_img, n)
    s = t.empty((n, 4), dtype=t.int64, device="cuda")
    u = t.empty_like(s)
    ctx.fr_random_dev(1, n, s)
    ctx.fr_random_dev(99, n, u)
    full = ctx.msm_g1_resident(rb, s)
    acc = None
    q = n // 4
    for k in range(4):
        part = ctx.msm_g1_resident(rb, s[k * q:(k + 1) * q], offset=k * q)
        acc = part if acc is None else ctx.g1_add(acc, part)
    assert (acc == full).all()
    su = t.empty_like(s)
    ctx.fr_binop_dev(0, s, u, su, n)
    lhs = ctx.g1_add(full, ctx.msm_g1_resident(rb, u))
    assert (lhs == ctx.msm_g1_resident(rb, su)).all()
    # full-size oracle value (a few seconds on 8 host threads)
    bases = d_img.cpu().numpy()
    scal = host_u64(s).reshape(-1, 4)
    assert (full == cref.msm_g1(bases, scal)).all()
    rb.free()


@pytest.mark.parametrize("n", [1, 5, 100, 3000])
def test_msm_g2_vs_oracle(ctx, cref, n):
    bases = cref.g2_random_bases(7, n)
    scalars = cref.fr_random(8, n)
    if n >= 100:
        v = fr_ints(scalars)
        v[0], v[1], v[2] = 0, 1, 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000000
        scalars = fr_arr(v)
    assert (ctx.msm_g2(bases, scalars) == cref.msm_g2(bases, scalars)).all()


def test_group_add_helpers(ctx, cref):
    bs = cref.g1_random_bases(3, 2)
    assert (ctx.g1_add(bs[0], bs[1]) == cref.g1_add(bs[0], bs[1])).all()
    assert (ctx.g1_add(bs[0], bs[0]) == cref.g1_add(bs[0], bs[0])).all()
    b2 = cref.g2_random_bases(3, 2)
    assert (ctx.g2_add(b2[0], b2[1]) == cref.g2_add(b2[0], b2[1])).all()


def test_launch_counter_moves(ctx, cref):
    before = ctx.launch_count
    ctx.poseidon(cref.fr_random(1, 8).reshape(2, 4, 4))
    assert ctx.launch_count > before


# ------------------------------------------------------------------ 4-ary Poseidon Merkle tree
def test_merkle4_build_prove_root(ctx, cref):
    """dense 4^6-leaf tree: every level equals the host SparseTree4 (the state manager's hash
    structure), proofs equal `prove()`, and recomputed roots equal the tree root; a wrong leaf gives a
    different root."""
    t = _t()
    from bazuka_b200.mpn import native as N
    log4 = 6
    n = 1 << (2 * log4)
    total = (4 ** (log4 + 1) - 1) // 3
    leaves = cref.fr_random(55, n)
    nodes = t.zeros((total, 4), dtype=t.int64, device="cuda")
    nodes[:n] = dev_u64(leaves)
    ctx.merkle4_build_dev(nodes, log4)
    ctx.synchronize()
    host = host_u64(nodes).reshape(-1, 4)
    # level 1 against the batched hash oracle, root against the big-int tree
    lvl1 = cref.poseidon(leaves.reshape(n // 4, 4, 4))
    assert (host[n:n + n // 4] == lvl1).all()
    tree = N.SparseTree4(log4, 0)
    ints = fr_ints(leaves)
    for i, v in enumerate(ints):
        tree.set_leaf(i, v)
    assert fr_ints(host[-1:])[0] == tree.root
    idx = np.array([0, 1, 5, 1000, n - 1, 2731], dtype=np.uint64)
    d_idx = t.from_numpy(idx.view(np.int64)).cuda()
    proofs = t.empty((len(idx), log4, 3, 4), dtype=t.int64, device="cuda")
    ctx.merkle4_prove_dev(nodes, log4, d_idx, proofs)
    ctx.synchronize()
    hp = host_u64(proofs).reshape(len(idx), log4, 3, 4)
    for k, i in enumerate(idx):
        want = tree.prove(int(i))
        got = [[fr_ints(hp[k, l, s:s + 1])[0] for s in range(3)] for l in range(log4)]
        assert got == want
    d_leaves = dev_u64(leaves[idx.astype(np.int64)])
    roots = t.empty((len(idx), 4), dtype=t.int64, device="cuda")
    ctx.merkle4_root_dev(log4, d_idx, d_leaves, proofs, roots)
    ctx.synchronize()
    assert all(r == tree.root for r in fr_ints(host_u64(roots).reshape(-1, 4)))
    bad_leaves = dev_u64(leaves[(idx.astype(np.int64) + 1) % n])
    ctx.merkle4_root_dev(log4, d_idx, bad_leaves, proofs, roots)
    ctx.synchronize()
    assert all(r != tree.root for r in fr_ints(host_u64(roots).reshape(-1, 4)))


# ------------------------------------------------------------------ fixed-base tables (resident proving-key columns)
@pytest.mark.parametrize("levels", [2, 3, 7, 16])
def test_msm_g1_fixed_base_table_equals_oracle(ctx, cref, levels):
    """bzk_g1_bases_precompute: level t = [2^(c*G*t)] P.  Same group element as the plain sum for uniform, witness-like and
    adversarial inputs (identity bases, repeated bases -> equal points inside a bucket, r-1, zero), and for sub-ranges."""
    n = 6000
    bases = cref.g1_random_bases(9, n)
    inf = np.zeros(104, dtype=np.uint8)
    inf[96] = 1
    bases[::7] = inf
    bases[1::7] = bases[1]
    vals = fr_ints(cref.fr_random(10, n))
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    for i in range(0, n, 3):
        vals[i] = i % 2
    vals[5], vals[6], vals[8] = R - 1, 0, (1 << 255) % R
    scalars = fr_arr(vals)
    rb = ctx.g1_bases(bases)
    got_levels = rb.precompute(levels)
    assert 2 <= got_levels <= levels and rb.levels == got_levels
    assert (ctx.msm_g1_resident(rb, scalars) == cref.msm_g1(bases, scalars)).all()
    assert (ctx.msm_g1_resident(rb, scalars[1000:3001], offset=1000) == cref.msm_g1(bases[1000:3001], scalars[1000:3001])).all()
    ones = fr_arr([1] * n)
    assert (ctx.msm_g1_resident(rb, ones) == cref.msm_g1(bases, ones)).all()
    rb.free()


@pytest.mark.parametrize("levels", [2, 16])
def test_msm_g2_fixed_base_table_equals_oracle(ctx, cref, levels):
    n = 1500
    bases = cref.g2_random_bases(7, n)
    bases[3] = bases[4]
    v = fr_ints(cref.fr_random(8, n))
    v[0], v[1], v[2] = 0, 1, 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000000
    scalars = fr_arr(v)
    rb = ctx.g2_bases(bases)
    rb.precompute(levels)
    assert (ctx.msm_g2_resident(rb, scalars) == cref.msm_g2(bases, scalars)).all()
    assert (ctx.msm_g2_resident(rb, scalars[100:900], offset=100) == cref.msm_g2(bases[100:900], scalars[100:900])).all()
    rb.free()


def test_msm_g1_2_20_fixed_base_table_vs_oracle(ctx, cref):
    """BASELINE configs[1] size through the table path (what bench.py times): value against the threaded C oracle."""
    t = _t()
    n = 1 << 20
    d_img = t.empty((n, 104), dtype=t.uint8, device="cuda")
    ctx.g1_random_bases_dev(2, n, d_img)
    rb = ctx.g1_bases_from_dev(d_img, n)
    rb.precompute(16)
    s = t.empty((n, 4), dtype=t.int64, device="cuda")
    ctx.fr_random_dev(1, n, s)
    ctx.synchronize()
    got = ctx.msm_g1_resident(rb, s)
    assert (got == cref.msm_g1(d_img.cpu().numpy(), host_u64(s).reshape(-1, 4))).all()
    rb.free()


@pytest.mark.parametrize("rounds", [1, 2, 5])
def test_msm_batched_affine_rounds_equal_oracle(ctx, cref, rounds):
    """csrc/msm_affine.cuh: R rounds of pairwise affine sums (one shared inversion per round) in front of the XYZZ
    accumulation give the same group element — uniform, witness-like, all-equal scalars, repeated bases (P + P inside a
    bucket), identity bases, P - P pairs, for G1 and G2, with and without a table."""
    n = 5000
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    bases = cref.g1_random_bases(9, n)
    inf = np.zeros(104, dtype=np.uint8)
    inf[96] = 1
    bases[::7] = inf
    bases[1::7] = bases[1]
    vals = fr_ints(cref.fr_random(10, n))
    for i in range(0, n, 3):
        vals[i] = i % 2
    vals[5], vals[6], vals[8] = R - 1, 0, (1 << 255) % R
    vals[15], vals[22] = 777, R - 777          # same base (index 1 mod 7), opposite scalars: P - P in one bucket
    scalars = fr_arr(vals)
    ctx.set_msm_affine_rounds(rounds, rounds)
    try:
        rb = ctx.g1_bases(bases)
        assert (ctx.msm_g1_resident(rb, scalars) == cref.msm_g1(bases, scalars)).all()
        rb.precompute(16)
        assert (ctx.msm_g1_resident(rb, scalars) == cref.msm_g1(bases, scalars)).all()
        ones = fr_arr([5] * n)
        assert (ctx.msm_g1_resident(rb, ones) == cref.msm_g1(bases, ones)).all()
        rb.free()
        b2 = cref.g2_random_bases(7, 1500)
        b2[3] = b2[4]
        s2 = cref.fr_random(8, 1500)
        r2 = ctx.g2_bases(b2)
        r2.precompute(16)
        assert (ctx.msm_g2_resident(r2, s2) == cref.msm_g2(b2, s2)).all()
        r2.free()
    finally:
        ctx.set_msm_affine_rounds(-1, -1)
