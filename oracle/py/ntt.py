"""ORACLE (test infrastructure only — never imported by the product path).

Radix-2 NTT over Fr exactly as bellman 0.14.0 `domain::EvaluationDomain` defines it.  bellman is an
un-vendored crates.io dependency (/root/reference/Cargo.toml:27); its published algorithm is restated:

  from_coeffs : m = next power of two >= len, omega = ROOT_OF_UNITY^(2^(S-exp)),
                geninv = 7^-1, minv = m^-1
  fft         : natural order in -> natural order out, A[k] = sum_j a[j] omega^(jk)
                (serial_fft: bit-reversal permutation, then log m DIT butterfly passes)
  ifft        : fft with omega^-1, then * minv
  coset_fft   : a[i] *= 7^i, then fft          icoset_fft : ifft, then a[i] *= 7^-i
  divide_by_z_on_coset : a[i] *= (7^m - 1)^-1
Call sites in the reference: every `create_random_proof` (src/mpn/circuits/test.rs:135,175,215).
"""
from .field import R_MOD, FR_ROOT_OF_UNITY, FR_S, FR_GENERATOR


def omega_for(log_n: int) -> int:
    w = FR_ROOT_OF_UNITY
    for _ in range(log_n, FR_S):
        w = w * w % R_MOD
    return w


def _bitrev(k, bits):
    return int(format(k, f"0{bits}b")[::-1], 2) if bits else 0


def serial_fft(a, omega, log_n):
    n = 1 << log_n
    assert len(a) == n
    a = list(a)
    for k in range(n):
        rk = _bitrev(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    for _ in range(log_n):
        w_m = pow(omega, n // (2 * m), R_MOD)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[k + j + m] * w % R_MOD
                a[k + j + m] = (a[k + j] - t) % R_MOD
                a[k + j] = (a[k + j] + t) % R_MOD
                w = w * w_m % R_MOD
        m *= 2
    return a


def fft(a, log_n):
    return serial_fft(a, omega_for(log_n), log_n)


def ifft(a, log_n):
    minv = pow(1 << log_n, -1, R_MOD)
    return [x * minv % R_MOD for x in serial_fft(a, pow(omega_for(log_n), -1, R_MOD), log_n)]


def distribute_powers(a, g):
    out, u = [], 1
    for x in a:
        out.append(x * u % R_MOD)
        u = u * g % R_MOD
    return out


def coset_fft(a, log_n):
    return fft(distribute_powers(a, FR_GENERATOR), log_n)


def icoset_fft(a, log_n):
    return distribute_powers(ifft(a, log_n), pow(FR_GENERATOR, -1, R_MOD))


def divide_by_z_on_coset(a, log_n):
    zi = pow((pow(FR_GENERATOR, 1 << log_n, R_MOD) - 1) % R_MOD, -1, R_MOD)
    return [x * zi % R_MOD for x in a]


def dft_naive(a, log_n):
    """the definition, O(n^2) — arbiter for serial_fft."""
    n = 1 << log_n
    w = omega_for(log_n)
    return [sum(a[j] * pow(w, j * k, R_MOD) for j in range(n)) % R_MOD for k in range(n)]
