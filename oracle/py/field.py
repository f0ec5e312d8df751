"""ORACLE (test infrastructure only — never imported by the product path).

Big-integer model of the two BLS12-381 prime fields and their Montgomery byte images.

Follows:
  * Fr  = `ZkScalar`, /root/reference/src/zk/mod.rs:202-206  (ff-derive: modulus r, generator 7,
          little-endian repr, in-memory = 4 x u64 Montgomery limbs with R = 2^256)
  * Fp  = `groth16::Fp([u64;6])`, /root/reference/src/zk/groth16/mod.rs:19-20 (transmuted image of
          bls12_381 0.8.0 `Fp`: 6 x u64 Montgomery limbs with R = 2^384)
  * `ZkScalar::new` (mod-r reduction of LE bytes), /root/reference/src/zk/mod.rs:262-271

bls12_381 0.8.0 / ff 0.13 are crates.io dependencies (Cargo.toml:19,28) that are NOT vendored in
/root/reference; their published conventions (Montgomery form, R=2^(64*limbs)) are restated here.
"""

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001  # Fr modulus r
P_MOD = int(
    "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab", 16
)  # Fp modulus p

FR_R = (1 << 256) % R_MOD  # Montgomery radix for Fr
FP_R = (1 << 384) % P_MOD  # Montgomery radix for Fp
FR_RINV = pow(FR_R, -1, R_MOD)
FP_RINV = pow(FP_R, -1, P_MOD)

FR_GENERATOR = 7  # PrimeFieldGenerator = "7" (src/zk/mod.rs:204)
FR_S = 32  # two-adicity of r-1
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_S, R_MOD)  # primitive 2^32-th root (ff-derive rule)

assert (R_MOD - 1) % (1 << FR_S) == 0 and ((R_MOD - 1) >> FR_S) % 2 == 1
assert pow(FR_ROOT_OF_UNITY, 1 << 31, R_MOD) == R_MOD - 1


# ---------------------------------------------------------------- Montgomery byte images
def fr_to_mont_bytes(x: int) -> bytes:
    """canonical integer -> 32-byte image of `ZkScalar([u64;4])` (Montgomery limbs, LE)."""
    return ((x % R_MOD) * FR_R % R_MOD).to_bytes(32, "little")


def fr_from_mont_bytes(b: bytes) -> int:
    v = int.from_bytes(b[:32], "little")
    assert v < R_MOD, "non-canonical Montgomery limbs"
    return v * FR_RINV % R_MOD


def fp_to_mont_bytes(x: int) -> bytes:
    return ((x % P_MOD) * FP_R % P_MOD).to_bytes(48, "little")


def fp_from_mont_bytes(b: bytes) -> int:
    v = int.from_bytes(b[:48], "little")
    assert v < P_MOD, "non-canonical Montgomery limbs"
    return v * FP_RINV % P_MOD


def zkscalar_new(le_bytes: bytes) -> int:
    """`ZkScalar::new` — LE integer reduced mod r (src/zk/mod.rs:262-271)."""
    return int.from_bytes(le_bytes, "little") % R_MOD


def fr_vec_to_mont(xs) -> bytes:
    return b"".join(fr_to_mont_bytes(x) for x in xs)


def fr_vec_from_mont(b: bytes):
    return [fr_from_mont_bytes(b[i : i + 32]) for i in range(0, len(b), 32)]


# ---------------------------------------------------------------- deterministic test inputs
class SplitMix64:
    """SplitMix64 stream — the input generator named in SURVEY.md §8(d) configs 2/3."""

    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def fr(self) -> int:
        """4 limbs -> 256-bit integer reduced mod r."""
        v = 0
        for i in range(4):
            v |= self.next() << (64 * i)
        return v % R_MOD
