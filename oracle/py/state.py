"""ORACLE (test infrastructure only — never imported by the product path).

State-model algebra of the 4-ary Poseidon Merkle state.
Follows /root/reference/src/zk/mod.rs:401-423 (`ZkStateModel::compress_default`):
Scalar -> 0, Struct -> H(children defaults), List -> H([d;4]) iterated log4_size times.
Pinned by the empty-MPN-root constant in /root/reference/src/node/api/get_explorer_blocks.rs:29.
"""
from .poseidon import poseidon

SCALAR = ("scalar",)


def struct(*fields):
    return ("struct", tuple(fields))


def list_(log4_size, item):
    return ("list", log4_size, item)


def compress_default(model):
    if model[0] == "scalar":
        return 0
    if model[0] == "struct":
        return poseidon([compress_default(f) for f in model[1]])
    d = compress_default(model[2])
    for _ in range(model[1]):
        d = poseidon([d, d, d, d])
    return d


def mpn_state_model(log4_tree_size, log4_token_tree_size):
    """/root/reference/src/mpn/mod.rs:219-240"""
    return list_(log4_tree_size, struct(SCALAR, SCALAR, SCALAR, SCALAR,
                                        list_(log4_token_tree_size, struct(SCALAR, SCALAR))))


# ---------------------------------------------------------------------------------------------
# Reference semantics of the batched ("versioned") tree update the GPU transition builder uses
# (bazuka_b200/csrc/poseidon.cu k_tree4_versioned_level).  TEST INFRASTRUCTURE: plain Python, sequential.
# Ground truth is the reference's own loop — one `set_data` root path per write with `prove` in between
# (/root/reference/src/zk/state/mod.rs:218-264,310-420) — which is what this function literally does on
# dictionaries; tests check the level-synchronous kernel against it.
# ---------------------------------------------------------------------------------------------
def sequential_tree_updates(depth, tree_ids, indices, leaf_values, init_proofs, hash4):
    """apply the writes one at a time to sparse trees known only through `init_proofs` (the proof of each
    written leaf in the pre-batch tree).  -> (vals [depth+1][n], proofs [n][depth][3]) like the kernel."""
    n = len(indices)
    nodes = {}  # (tree, level, node index) -> value, for nodes some write of the batch has produced
    vals = [[0] * n for _ in range(depth + 1)]
    proofs = []
    for e in range(n):
        t, idx, cur = tree_ids[e], indices[e], leaf_values[e]
        vals[0][e] = cur
        nodes[(t, 0, idx)] = cur
        proof = []
        for lvl in range(depth):
            pos, base = idx & 3, (idx >> 2) << 2
            init = list(init_proofs[e][lvl])
            sib, w = [], 0
            for k in range(4):
                if k == pos:
                    continue
                sib.append(nodes.get((t, lvl, base + k), init[w]))
                w += 1
            proof.append(sib)
            kids = list(sib)
            kids.insert(pos, cur)
            cur = hash4(kids)
            idx >>= 2
            vals[lvl + 1][e] = cur
            nodes[(t, lvl + 1, idx)] = cur
        proofs.append(proof)
    return vals, proofs


class HostTreeHasher:
    """drop-in for bazuka_b200.mpn.batch_update.GpuTreeHasher in CPU tests (oracle Poseidon, sequential trees)."""

    def __init__(self, poseidon):
        self.poseidon = poseidon

    def poseidon_batch(self, rows):
        return [self.poseidon(list(r)) for r in rows]

    def tree_update(self, depth, tree_ids, indices, leaf_values, init_proofs):
        if not indices:
            return [[] for _ in range(depth + 1)], []
        return sequential_tree_updates(depth, tree_ids, indices, leaf_values, init_proofs, self.poseidon)
