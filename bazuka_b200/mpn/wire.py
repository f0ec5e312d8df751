"""The prover worker's wire protocol: bincode images of what a Bazuka node hands out and takes back.

  MpnWork {config, public_inputs, data, new_root, reward}      /root/reference/src/mpn/mod.rs:264-270
  MpnConfig, MpnWorkData, ZkPublicInputs                       /root/reference/src/mpn/mod.rs:203-262
  {Deposit,Withdraw,Update}Transition                          /root/reference/src/mpn/mod.rs:427-537
  GetMpnWork{Request,Response}, PostMpnSolution{Request,Response}, PostMpnWorker{Request,Response}
                                                               /root/reference/src/client/messages.rs:368-397
  MpnWork::verify's commitment                                 /root/reference/src/mpn/mod.rs:281-295

bincode 1.x default options (what `bincode::serialize` gives the reference): little-endian fixed-width integers, `usize`
and every length prefix as u64, enum variant index as u32, bool and Option tag as one byte, `PhantomData` as nothing,
tuples / structs / fixed arrays as their fields back to back.  Field elements travel as their raw MONTGOMERY limbs
(`ZkScalar([u64;4])`, serde derive on the tuple struct, /root/reference/src/zk/mod.rs:202-206); this module keeps them as
canonical Python ints and converts at the edge.  Two encodings come from crates that are not vendored and are restated
from their documented behaviour (ext): an ed25519 `Address` is `serialize_bytes` of the 32 key bytes (u64 length 32 +
bytes = 40 B), an ed25519 signature likewise 8 + 64 B.

Values are plain dicts / lists named after the Rust fields, so a decoded work re-encodes to the same bytes
(`HashMap`s keep the order they arrived in; a Rust peer may emit another order, compare decoded values)."""
import hashlib
import struct

from .native import R as R_MOD

_RINV = pow(1 << 256, -1, R_MOD)


class Writer:
    def __init__(self):
        self.b = bytearray()

    def u8(self, v): self.b.append(v & 0xFF)
    def bool(self, v): self.b.append(1 if v else 0)
    def u32(self, v): self.b += struct.pack("<I", v)
    def u64(self, v): self.b += struct.pack("<Q", v)
    def raw(self, v): self.b += bytes(v)
    def fr(self, v): self.b += (((int(v) % R_MOD) << 256) % R_MOD).to_bytes(32, "little")   # canonical -> Montgomery limbs
    def bytes_(self, v): self.u64(len(v)); self.raw(v)
    def string(self, v): self.bytes_(v.encode())

    def vec(self, items, enc):
        self.u64(len(items))
        for it in items:
            enc(self, it)

    def option(self, v, enc):
        if v is None:
            self.u8(0)
        else:
            self.u8(1)
            enc(self, v)


class Reader:
    def __init__(self, data):
        self.d, self.o = memoryview(bytes(data)), 0

    def _take(self, n):
        if self.o + n > len(self.d):
            raise ValueError("bincode: input ends early")
        v = self.d[self.o:self.o + n]
        self.o += n
        return v

    def u8(self): return self._take(1)[0]

    def bool(self):
        v = self.u8()
        if v > 1:
            raise ValueError("bincode: invalid bool")
        return bool(v)

    def u32(self): return struct.unpack("<I", self._take(4))[0]
    def u64(self): return struct.unpack("<Q", self._take(8))[0]
    def raw(self, n): return bytes(self._take(n))

    def fr(self):
        m = int.from_bytes(self._take(32), "little")
        if m >= R_MOD:
            raise ValueError("bincode: scalar limbs not reduced")
        return m * _RINV % R_MOD

    def bytes_(self, limit=1 << 24):
        n = self.u64()
        if n > limit:
            raise ValueError("bincode: length prefix too large")
        return self.raw(n)

    def string(self): return self.bytes_().decode()

    def vec(self, dec, limit=1 << 24):
        n = self.u64()
        if n > limit:
            raise ValueError("bincode: length prefix too large")
        return [dec(self) for _ in range(n)]

    def option(self, dec):
        tag = self.u8()
        if tag > 1:
            raise ValueError("bincode: invalid Option tag")
        return dec(self) if tag else None

    def done(self):
        if self.o != len(self.d):
            raise ValueError("bincode: trailing bytes")


# ------------------------------------------------------------------ leaves
def enc_address(w, a):            # ed25519 public key, 32 bytes (ext: serialize_bytes)
    assert len(a) == 32
    w.bytes_(a)


def dec_address(r):
    a = r.bytes_(64)
    if len(a) != 32:
        raise ValueError("bincode: ed25519 key is 32 bytes")
    return a


def enc_contract_id(w, c):        # ContractId::{Null, Ziesha, Custom(scalar)} <- None / "ziesha" / int
    if c is None:
        w.u32(0)
    elif c == "ziesha":
        w.u32(1)
    else:
        w.u32(2)
        w.fr(c)


def dec_contract_id(r):
    t = r.u32()
    if t == 0:
        return None
    if t == 1:
        return "ziesha"
    if t == 2:
        return r.fr()
    raise ValueError("bincode: ContractId variant")


def contract_id_scalar(c):
    """`impl From<ContractId> for ZkScalar` (/root/reference/src/zk/mod.rs:280-288)"""
    return 0 if c is None else 1 if c == "ziesha" else c


def scalar_contract_id(s):
    return None if s == 0 else "ziesha" if s == 1 else s


def enc_money(w, m): enc_contract_id(w, m["token_id"]); w.u64(m["amount"])
def dec_money(r): return {"token_id": dec_contract_id(r), "amount": r.u64()}
def enc_point(w, p): w.fr(p[0]); w.fr(p[1])                       # jubjub::PointAffine
def dec_point(r): return (r.fr(), r.fr())
def enc_pubkey(w, k): w.fr(k[0]); w.bool(k[1])                     # jubjub::PublicKey(PointCompressed(x, is_odd))
def dec_pubkey(r): return (r.fr(), r.bool())
def enc_zk_sig(w, s): enc_point(w, s["r"]); w.fr(s["s"])           # jubjub::Signature {r, s}
def dec_zk_sig(r): return {"r": dec_point(r), "s": r.fr()}
def enc_proof3(w, p): w.vec(p, lambda w_, row: [w_.fr(x) for x in row])
def dec_proof3(r): return r.vec(lambda r_: [r_.fr(), r_.fr(), r_.fr()])


def enc_account(w, a):            # zk::MpnAccount
    w.u32(a["tx_nonce"]); w.u32(a["withdraw_nonce"]); enc_point(w, a["address"])
    w.u64(len(a["tokens"]))
    for k, m in a["tokens"].items():
        w.u64(k)
        enc_money(w, m)


def dec_account(r):
    a = {"tx_nonce": r.u32(), "withdraw_nonce": r.u32(), "address": dec_point(r), "tokens": {}}
    for _ in range(r.u64()):
        k = r.u64()
        a["tokens"][k] = dec_money(r)
    return a


def enc_mpn_tx(w, t):             # zk::MpnTransaction
    w.u32(t["nonce"]); enc_pubkey(w, t["src_pub_key"]); enc_pubkey(w, t["dst_pub_key"])
    enc_money(w, t["amount"]); enc_money(w, t["fee"]); enc_zk_sig(w, t["sig"])


def dec_mpn_tx(r):
    return {"nonce": r.u32(), "src_pub_key": dec_pubkey(r), "dst_pub_key": dec_pubkey(r), "amount": dec_money(r), "fee": dec_money(r),
            "sig": dec_zk_sig(r)}


def enc_contract_deposit(w, p):   # core::ContractDeposit
    w.string(p["memo"]); enc_contract_id(w, p["contract_id"]); w.u32(p["deposit_circuit_id"]); w.fr(p["calldata"])
    enc_address(w, p["src"]); enc_money(w, p["amount"]); enc_money(w, p["fee"]); w.u32(p["nonce"])
    w.option(p["sig"], lambda w_, s: w_.bytes_(s))


def dec_contract_deposit(r):
    return {"memo": r.string(), "contract_id": dec_contract_id(r), "deposit_circuit_id": r.u32(), "calldata": r.fr(), "src": dec_address(r),
            "amount": dec_money(r), "fee": dec_money(r), "nonce": r.u32(), "sig": r.option(lambda r_: r_.bytes_(128))}


def enc_contract_withdraw(w, p):  # core::ContractWithdraw
    w.string(p["memo"]); enc_contract_id(w, p["contract_id"]); w.u32(p["withdraw_circuit_id"]); w.fr(p["calldata"])
    enc_address(w, p["dst"]); enc_money(w, p["amount"]); enc_money(w, p["fee"])


def dec_contract_withdraw(r):
    return {"memo": r.string(), "contract_id": dec_contract_id(r), "withdraw_circuit_id": r.u32(), "calldata": r.fr(), "dst": dec_address(r),
            "amount": dec_money(r), "fee": dec_money(r)}


def enc_mpn_deposit(w, d): enc_pubkey(w, d["mpn_address"]); enc_contract_deposit(w, d["payment"])
def dec_mpn_deposit(r): return {"mpn_address": dec_pubkey(r), "payment": dec_contract_deposit(r)}


def enc_mpn_withdraw(w, d):
    enc_pubkey(w, d["mpn_address"]); w.u32(d["mpn_withdraw_nonce"]); enc_zk_sig(w, d["mpn_sig"]); enc_contract_withdraw(w, d["payment"])


def dec_mpn_withdraw(r):
    return {"mpn_address": dec_pubkey(r), "mpn_withdraw_nonce": r.u32(), "mpn_sig": dec_zk_sig(r), "payment": dec_contract_withdraw(r)}


# ------------------------------------------------------------------ transitions (field order = the Rust structs')
_UPDATE_FIELDS = [("enabled", "bool"), ("tx", "mpn_tx"), ("src_before", "account"), ("src_before_balances_hash", "fr"), ("src_before_balance", "money"),
                  ("src_before_fee_balance", "money"), ("src_proof", "proof3"), ("src_index", "u64"), ("src_token_index", "u64"),
                  ("src_balance_proof", "proof3"), ("src_fee_token_index", "u64"), ("src_fee_balance_proof", "proof3"), ("dst_before", "account"),
                  ("dst_before_balances_hash", "fr"), ("dst_before_balance", "money"), ("dst_proof", "proof3"), ("dst_index", "u64"),
                  ("dst_token_index", "u64"), ("dst_balance_proof", "proof3")]
_DEPOSIT_FIELDS = [("enabled", "bool"), ("tx", "mpn_deposit"), ("before", "account"), ("before_balances_hash", "fr"), ("before_balance", "money"),
                   ("proof", "proof3"), ("account_index", "u64"), ("token_index", "u64"), ("balance_proof", "proof3")]
_WITHDRAW_FIELDS = [("enabled", "bool"), ("tx", "mpn_withdraw"), ("before", "account"), ("before_token_balance", "money"), ("before_fee_balance", "money"),
                    ("proof", "proof3"), ("account_index", "u64"), ("token_index", "u64"), ("token_balance_proof", "proof3"),
                    ("before_token_hash", "fr"), ("fee_token_index", "u64"), ("fee_balance_proof", "proof3")]
_ENC = {"bool": Writer.bool, "u64": Writer.u64, "fr": Writer.fr, "money": enc_money, "account": enc_account, "proof3": enc_proof3,
        "mpn_tx": enc_mpn_tx, "mpn_deposit": enc_mpn_deposit, "mpn_withdraw": enc_mpn_withdraw}
_DEC = {"bool": Reader.bool, "u64": Reader.u64, "fr": Reader.fr, "money": dec_money, "account": dec_account, "proof3": dec_proof3,
        "mpn_tx": dec_mpn_tx, "mpn_deposit": dec_mpn_deposit, "mpn_withdraw": dec_mpn_withdraw}


def _enc_struct(fields):
    return lambda w, v: [_ENC[k](w, v[name]) for name, k in fields]


def _dec_struct(fields):
    return lambda r: {name: _DEC[k](r) for name, k in fields}


# ------------------------------------------------------------------ keys, proofs, config, work
def enc_verifier_key(w, vk_blob):   # ZkVerifierKey::Groth16(Box<Groth16VerifyingKey>): u32 tag 0 + the 878+97n byte image
    w.u32(0)
    w.raw(vk_blob)


def dec_verifier_key(r):
    if r.u32() != 0:
        raise ValueError("bincode: ZkVerifierKey variant")
    head = r.raw(870)
    n = r.u64()
    if n > 4096:
        raise ValueError("bincode: verifying key too long")
    return head + struct.pack("<Q", n) + r.raw(97 * n)


def enc_zkproof(w, proof387):       # ZkProof::Groth16(Box<Groth16Proof>): 391 bytes (/root/reference/src/zk/mod.rs:646-651)
    assert len(proof387) == 387
    w.u32(0)
    w.raw(proof387)


def dec_zkproof(r):
    if r.u32() != 0:
        raise ValueError("bincode: ZkProof variant")
    return r.raw(387)


_CONFIG_U8 = ["log4_tree_size", "log4_token_tree_size", "log4_deposit_batch_size", "log4_withdraw_batch_size", "log4_update_batch_size"]
_CONFIG_USIZE = ["mpn_num_update_batches", "mpn_num_deposit_batches", "mpn_num_withdraw_batches"]
_CONFIG_VK = ["deposit_vk", "withdraw_vk", "update_vk"]


def enc_config(w, c):
    for k in _CONFIG_U8:
        w.u8(c[k])
    enc_contract_id(w, c["mpn_contract_id"])
    for k in _CONFIG_USIZE:
        w.u64(c[k])
    for k in _CONFIG_VK:
        enc_verifier_key(w, c[k])


def dec_config(r):
    c = {k: r.u8() for k in _CONFIG_U8}
    c["mpn_contract_id"] = dec_contract_id(r)
    c.update({k: r.u64() for k in _CONFIG_USIZE})
    c.update({k: dec_verifier_key(r) for k in _CONFIG_VK})
    return c


_KINDS = ["deposit", "withdraw", "update"]           # MpnWorkData variant order
_TRANSITION = {"deposit": _DEPOSIT_FIELDS, "withdraw": _WITHDRAW_FIELDS, "update": _UPDATE_FIELDS}


def enc_work(w, work):
    enc_config(w, work["config"])
    p = work["public_inputs"]
    w.u64(p["height"]); w.fr(p["state"]); w.fr(p["aux_data"]); w.fr(p["next_state"])
    kind, items = work["data"]
    w.u32(_KINDS.index(kind))
    w.vec(items, _enc_struct(_TRANSITION[kind]))
    w.fr(work["new_root"]["state_hash"]); w.u64(work["new_root"]["state_size"])
    w.u64(work["reward"])


def dec_work(r):
    config = dec_config(r)
    p = {"height": r.u64(), "state": r.fr(), "aux_data": r.fr(), "next_state": r.fr()}
    tag = r.u32()
    if tag > 2:
        raise ValueError("bincode: MpnWorkData variant")
    kind = _KINDS[tag]
    items = r.vec(_dec_struct(_TRANSITION[kind]), limit=1 << 16)
    return {"config": config, "public_inputs": p, "data": (kind, items), "new_root": {"state_hash": r.fr(), "state_size": r.u64()}, "reward": r.u64()}


def work_to_bytes(work):
    w = Writer()
    enc_work(w, work)
    return bytes(w.b)


def work_from_bytes(b):
    r = Reader(b)
    work = dec_work(r)
    r.done()
    return work


# ------------------------------------------------------------------ messages (client/messages.rs:368-397)
def get_mpn_work_request(address):
    w = Writer()
    enc_address(w, address)
    return bytes(w.b)


def get_mpn_work_response_to_bytes(works):
    w = Writer()
    w.u64(len(works))
    for wid, work in works.items():
        w.u64(wid)
        enc_work(w, work)
    return bytes(w.b)


def get_mpn_work_response_from_bytes(b):
    r = Reader(b)
    works = {}
    n = r.u64()
    if n > 1 << 16:
        raise ValueError("bincode: too many works")
    for _ in range(n):
        wid = r.u64()
        works[wid] = dec_work(r)
    r.done()
    return works


def post_mpn_solution_request(prover, proofs):
    """proofs: {work id: 387-byte Groth16Proof image}"""
    w = Writer()
    enc_address(w, prover)
    w.u64(len(proofs))
    for wid, p in proofs.items():
        w.u64(wid)
        enc_zkproof(w, bytes(p))
    return bytes(w.b)


def post_mpn_solution_request_from_bytes(b):
    r = Reader(b)
    prover = dec_address(r)
    proofs = {}
    for _ in range(r.u64()):
        wid = r.u64()
        proofs[wid] = dec_zkproof(r)
    r.done()
    return prover, proofs


def post_mpn_solution_response_from_bytes(b):
    r = Reader(b)
    accepted = r.u64()
    r.done()
    return accepted


def post_mpn_worker_request(address):
    return get_mpn_work_request(address)


def post_mpn_worker_response_from_bytes(b):
    r = Reader(b)
    ok = r.bool()
    r.done()
    return ok


# ------------------------------------------------------------------ MpnWork::verify's commitment
def commitment(prover_address, reward):
    """`ZkScalar::new(Hasher::hash(&bincode::serialize(&(prover, reward))))` (/root/reference/src/mpn/mod.rs:283-285; the same
    formula chain-side, src/blockchain/ops/apply_tx/update_contract/mod.rs:29-32): sha3-256 of the 48-byte image
    (8 + 32 address bytes, u64 reward), read as a little-endian integer and reduced mod r (zk/mod.rs:262-271)."""
    w = Writer()
    enc_address(w, prover_address)
    w.u64(reward)
    return int.from_bytes(hashlib.sha3_256(bytes(w.b)).digest(), "little") % R_MOD
