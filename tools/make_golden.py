#!/usr/bin/env python3
"""Extract the reference's own fixtures for this path into tests/golden/ (run in the build
container, where /root/reference is mounted; the GPU box only sees the committed JSON).

  poseidon_kats.json   16 known answers of `test_hash_samples`
                       (/root/reference/src/zk/poseidon/mod.rs:116-133): poseidon([0..n)), n=1..16
  mpn_vks.json         the three production verifying keys, 1460-byte bincode blobs of Montgomery
                       limbs (/root/reference/src/config/blockchain.rs:32-37)
  empty_root.json      the empty MPN state root printed in the explorer test vector
                       (/root/reference/src/node/api/get_explorer_blocks.rs:29), with the state
                       model sizes it was produced with
  genesis_mpn_addresses.json  the `jub…` MPN public keys of the genesis allocation
                       (/root/reference/src/config/initials.rs:13027+): PointCompressed(x, is_odd) strings, every one
                       must decompress (Fr square root + parity rule, src/crypto/jubjub/curve.rs:78-88) to a curve point
"""
import json, os, re

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    src = open(f"{REF}/zk/poseidon/mod.rs").read()
    body = src[src.index("fn test_hash_samples"):]
    kats = re.findall(r'"(\d{60,80})"', body)[:16]
    assert len(kats) == 16
    json.dump({"source": "src/zk/poseidon/mod.rs:116-133", "inputs": "poseidon([0,1,..,n-1]) for n=1..16",
               "expected_decimal": kats}, open(f"{OUT}/poseidon_kats.json", "w"), indent=1)

    cfg = open(f"{REF}/config/blockchain.rs").read()
    vks = {}
    for name in ("MPN_UPDATE_VK", "MPN_DEPOSIT_VK", "MPN_WITHDRAW_VK"):
        m = re.search(name + r'[^"]*hex::decode\("([0-9a-f]+)"', cfg, re.S)
        vks[name] = m.group(1)
        assert len(vks[name]) == 2 * 1460
    json.dump({"source": "src/config/blockchain.rs:32-37", "format": "bincode(Groth16VerifyingKey), Montgomery limbs",
               "vks": vks}, open(f"{OUT}/mpn_vks.json", "w"), indent=1)

    ex = open(f"{REF}/node/api/get_explorer_blocks.rs").read()
    m = re.search(r'ExplorerMpnAccount|state_model', ex)
    roots = re.findall(r'0x[0-9a-f]{64}', ex)
    json.dump({"source": "src/node/api/get_explorer_blocks.rs:29", "hex_scalars_in_vector": sorted(set(roots))},
              open(f"{OUT}/empty_root.json", "w"), indent=1)
    ini = open(f"{REF}/config/initials.rs").read()
    addrs = sorted(set(re.findall(r'"(jub[23][0-9a-f]{64})"', ini)))
    assert len(addrs) > 100
    json.dump({"source": "src/config/initials.rs:13027+", "format": "jub{2|3 = y parity}{x as 64 hex digits, big-endian} (src/crypto/jubjub/mod.rs:60-106)",
               "addresses": addrs}, open(f"{OUT}/genesis_mpn_addresses.json", "w"), indent=0)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
