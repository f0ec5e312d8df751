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
