"""ORACLE (test infrastructure only): ctypes face of oracle/_build/libbzk_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Arrays are numpy uint64 ([n,4] Fr / [n,6] Fp, Montgomery images) or uint8
([n,104] G1 / [n,200] G2 affine images: x | y | inf | pad)."""
import ctypes as ct
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbzk_oracle.so")
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, "c", f) for f in ("bzk_oracle.c", "mont_tmpl.h", "ec_tmpl.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ct.CDLL(_SO)
        _lib.bzko_init()
        blob = open(os.path.join(_HERE, "..", "bazuka_b200", "data", "poseidon_params.bin"), "rb").read()
        rc = _lib.bzko_poseidon_load(blob, ct.c_size_t(len(blob)))
        assert rc == 0, rc
    return _lib


def _p(a):
    return a.ctypes.data_as(ct.c_void_p)


def _u64(a, w):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == w
    return a


def usable_cpus():
    """host threads this process may really use: the scheduler affinity mask, capped by the cgroup
    CPU quota (cpu.max of cgroup v2 / cfs_quota of v1).  os.cpu_count() alone reports the machine's
    cores even inside a container limited to a fraction of them — the round-1 CPU arm oversubscribed
    such a box 16x and ran 5x slower than on an unrestricted one."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    try:
        q, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(period)
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / period
        except Exception:
            pass
    if quota is not None:
        n = max(1, min(n, int(quota + 0.5)))
    return max(1, n)


def cpu_info():
    """what the CPU arm ran on (printed in every bench line that times the oracle)."""
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    try:
        aff = len(os.sched_getaffinity(0))
    except AttributeError:
        aff = os.cpu_count() or 1
    return {"model": model, "os_cpu_count": os.cpu_count(), "affinity": aff, "usable": usable_cpus()}


NCPU = usable_cpus()


def _binop(name, w):
    def f(a, b):
        a, b = _u64(a, w), _u64(b, w)
        r = np.empty_like(a)
        getattr(lib(), name)(_p(a), _p(b), _p(r), ct.c_size_t(a.size // w))
        return r
    return f


def _unop(name, w):
    def f(a):
        a = _u64(a, w)
        r = np.empty_like(a)
        getattr(lib(), name)(_p(a), _p(r), ct.c_size_t(a.size // w))
        return r
    return f


fr_mul, fr_add, fr_sub = _binop("bzko_fr_mul", 4), _binop("bzko_fr_add", 4), _binop("bzko_fr_sub", 4)
fp_mul, fp_add, fp_sub = _binop("bzko_fp_mul", 6), _binop("bzko_fp_add", 6), _binop("bzko_fp_sub", 6)
fr_inv, fr_to_mont, fr_from_mont = _unop("bzko_fr_inv", 4), _unop("bzko_fr_to_mont", 4), _unop("bzko_fr_from_mont", 4)
fp_inv, fp_to_mont, fp_from_mont = _unop("bzko_fp_inv", 6), _unop("bzko_fp_to_mont", 6), _unop("bzko_fp_from_mont", 6)


def fr_random(seed, n):
    out = np.empty((n, 4), dtype=np.uint64)
    lib().bzko_fr_random(ct.c_uint64(seed), _p(out), ct.c_size_t(n))
    return out


def g1_generator():
    out = np.zeros(104, dtype=np.uint8)
    lib().bzko_g1_generator(_p(out))
    return out


def g2_generator():
    out = np.zeros(200, dtype=np.uint8)
    lib().bzko_g2_generator(_p(out))
    return out


def g1_on_curve(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    return bool(lib().bzko_g1_on_curve(_p(img)))


def g2_on_curve(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    return bool(lib().bzko_g2_on_curve(_p(img)))


def _ec2(name, nbytes):
    def f(a, b):
        a = np.ascontiguousarray(a, dtype=np.uint8); b = np.ascontiguousarray(b, dtype=np.uint8)
        out = np.zeros(nbytes, dtype=np.uint8)
        getattr(lib(), name)(_p(a), _p(b), _p(out))
        return out
    return f


g1_add, g2_add = _ec2("bzko_g1_add", 104), _ec2("bzko_g2_add", 200)


def g1_mul(a, k_mont):
    a = np.ascontiguousarray(a, dtype=np.uint8); k = _u64(k_mont, 4)
    out = np.zeros(104, dtype=np.uint8)
    lib().bzko_g1_mul(_p(a), _p(k), _p(out))
    return out


def g2_mul(a, k_mont):
    a = np.ascontiguousarray(a, dtype=np.uint8); k = _u64(k_mont, 4)
    out = np.zeros(200, dtype=np.uint8)
    lib().bzko_g2_mul(_p(a), _p(k), _p(out))
    return out


def g1_random_bases(seed, n, threads=NCPU):
    out = np.zeros((n, 104), dtype=np.uint8)
    lib().bzko_g1_random_bases(ct.c_uint64(seed), _p(out), ct.c_size_t(n), ct.c_int(threads))
    return out


def g2_random_bases(seed, n, threads=NCPU):
    out = np.zeros((n, 200), dtype=np.uint8)
    lib().bzko_g2_random_bases(ct.c_uint64(seed), _p(out), ct.c_size_t(n), ct.c_int(threads))
    return out


def _msm(name, nbytes, naive=False):
    def f(bases, scalars, threads=NCPU):
        bases = np.ascontiguousarray(bases, dtype=np.uint8); scalars = _u64(scalars, 4)
        n = scalars.size // 4
        assert bases.size == n * nbytes
        out = np.zeros(nbytes, dtype=np.uint8)
        if naive:
            getattr(lib(), name)(_p(bases), _p(scalars), ct.c_size_t(n), _p(out))
        else:
            getattr(lib(), name)(_p(bases), _p(scalars), ct.c_size_t(n), _p(out), ct.c_int(threads))
        return out
    return f


msm_g1, msm_g2 = _msm("bzko_msm_g1", 104), _msm("bzko_msm_g2", 200)
msm_g1_naive, msm_g2_naive = _msm("bzko_msm_g1_naive", 104, True), _msm("bzko_msm_g2_naive", 200, True)


def poseidon(inputs, threads=NCPU):
    """inputs [n, arity, 4] Montgomery -> [n, 4] digests."""
    inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
    n, arity, _ = inputs.shape
    out = np.empty((n, 4), dtype=np.uint64)
    rc = lib().bzko_poseidon(_p(inputs), ct.c_size_t(n), ct.c_uint32(arity), _p(out), ct.c_int(threads))
    assert rc == 0, rc
    return out


NTT_FFT, NTT_IFFT, NTT_COSET_FFT, NTT_ICOSET_FFT = 0, 1, 2, 3


def ntt(a, op, threads=NCPU):
    """returns the transformed copy; a is [2^k, 4] Montgomery."""
    a = np.array(a, dtype=np.uint64, copy=True)
    n = a.size // 4
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().bzko_ntt(_p(a), ct.c_uint(log_n), ct.c_int(op), ct.c_int(threads))
    assert rc == 0
    return a


def divide_by_z_on_coset(a):
    a = np.array(a, dtype=np.uint64, copy=True)
    n = a.size // 4
    lib().bzko_divide_by_z_on_coset(_p(a), ct.c_uint(n.bit_length() - 1))
    return a
