#!/usr/bin/env python3
"""Build bazuka_b200/data/poseidon_params.bin from the reference's hadeshash text dumps.

The 16 files under /root/reference/src/zk/poseidon/params/ are *data* (round constants and
MDS matrices of the Poseidon instances x^5, t=2..17 over BLS12-381 Fr). They are parsed the
way the reference parses them (src/zk/poseidon/params/mod.rs:39-57: line 1 -> t, R_F, R_P;
line 4 -> round constants; line 16 -> MDS, row-major chunks of t) and re-emitted as one
compact binary table so the repo carries no reference text.

Layout (little-endian):
  magic  "BZKPOSv1"                       8 B
  u32 n_widths (=16)
  per width, in order t=2..17:
     u32 t, u32 R_F, u32 R_P, u32 n_rc (= t*(R_F+R_P))
     n_rc  x 32 B  round constants, canonical little-endian integers < r
     t*t   x 32 B  MDS matrix, row-major (row j multiplies the state to give lane j)
Only needs to be re-run if the reference's parameter files change (they never have).
"""
import struct, sys, os, re, hashlib

REF = "/root/reference/src/zk/poseidon/params"
OUT = os.path.join(os.path.dirname(__file__), "..", "bazuka_b200", "data", "poseidon_params.bin")
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def consts(line):
    return [int(h, 16) for h in re.findall(r"0x([0-9a-fA-F]+)", line)]


def main():
    blob = bytearray(b"BZKPOSv1") + struct.pack("<I", 16)
    for t in range(2, 18):
        lines = open(f"{REF}/poseidon_params_n255_t{t}_alpha5_M128.txt").read().splitlines()
        opts = [s.strip() for s in lines[0].split(",")]
        tt = int(opts[1].split("=")[1]); rf = int(opts[4].split("=")[1]); rp = int(opts[5].split("=")[1])
        assert tt == t
        rc = consts(lines[3]); mds = consts(lines[15])
        assert len(rc) == t * (rf + rp), (t, len(rc))
        assert len(mds) == t * t
        assert all(c < R for c in rc + mds)
        blob += struct.pack("<IIII", t, rf, rp, len(rc))
        for c in rc + mds:
            blob += c.to_bytes(32, "little")
    with open(OUT, "wb") as f:
        f.write(blob)
    print(len(blob), hashlib.sha256(blob).hexdigest())


if __name__ == "__main__":
    main()
