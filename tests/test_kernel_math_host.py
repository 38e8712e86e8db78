"""CPU: the kernel's arithmetic header (ml_unigen_b200/csrc/ug_math.cuh) compiled for the host
must agree bit for bit with the C oracle: exp (both scaling variants), log, Philox noise
streams, confidence, bf16 CFG mix.  This is what makes GPU-vs-oracle equality a by-construction
property rather than a statistical one."""
import ctypes
import os
import subprocess
import struct

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostmath") / "libug_hostmath.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--shared", "-Xcompiler", "-fPIC,-ffp-contract=off",
                           "-I" + os.path.join(ROOT, "ml_unigen_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_math_check.cu"), "-o", out])
    L = ctypes.CDLL(out)
    for name in ("ugk_host_expf_exact", "ugk_host_expf_fast", "ugk_host_logf"):
        getattr(L, name).restype = ctypes.c_float
        getattr(L, name).argtypes = [ctypes.c_float]
    L.ugk_host_race_q.restype = ctypes.c_float
    L.ugk_host_race_q.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
    L.ugk_host_select_u.restype = ctypes.c_float
    L.ugk_host_select_u.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
    L.ugk_host_confidence.restype = ctypes.c_float
    L.ugk_host_confidence.argtypes = [ctypes.c_float] * 3
    L.ugk_host_cfg_mix.restype = ctypes.c_uint32
    L.ugk_host_cfg_mix.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float]
    L.ugk_host_ordered_key.restype = ctypes.c_uint32
    L.ugk_host_ordered_key.argtypes = [ctypes.c_float]
    return L


def _bits(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


def test_exp_bitwise_equal_to_oracle(hostlib):
    ol = O.lib()
    fn = ctypes.cast(ol.ug_oracle_expf, ctypes.c_void_p)
    me, mf = ctypes.c_uint64(), ctypes.c_uint64()
    # all negative floats from -0.0 down to -110 with stride 61 (~18M points)
    hostlib.ugk_host_expf_scan.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                           ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    hostlib.ugk_host_expf_scan(0x80000000, _bits(-110.0), 61, fn, ctypes.byref(me), ctypes.byref(mf))
    assert me.value == 0 and mf.value == 0


def test_log_bitwise_equal_to_oracle(hostlib):
    ol = O.lib()
    fn = ctypes.cast(ol.ug_oracle_logf, ctypes.c_void_p)
    m = ctypes.c_uint64()
    hostlib.ugk_host_logf_scan.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                           ctypes.POINTER(ctypes.c_uint64)]
    hostlib.ugk_host_logf_scan(_bits(1e-20), _bits(3.4e38), 97, fn, ctypes.byref(m))
    assert m.value == 0


def test_noise_streams_equal_to_oracle(hostlib):
    ol = O.lib()
    rng = np.random.default_rng(0)
    for _ in range(20000):
        seed, off, row = (int(x) for x in rng.integers(0, 2**63, size=3))
        i = int(rng.integers(0, 16384))
        assert _bits(hostlib.ugk_host_race_q(seed, off, row, i)) == _bits(ol.ug_oracle_race_q(seed, off, row, i))
        n = int(rng.integers(0, 4096))
        assert _bits(hostlib.ugk_host_select_u(seed, off, row, n)) == _bits(ol.ug_oracle_select_u(seed, off, row, n))


def test_cfg_mix_equal_to_torch_bf16(hostlib, golden_dir):
    g = np.load(os.path.join(golden_dir, "cfg_mix.npz"))
    c = g["cond"].reshape(-1)[:4096]
    u = g["uncond"].reshape(-1)[:4096]
    for s in (2.0, 6.0, 1.75):
        ref = g[f"mixed_{s}"].reshape(-1)[:4096]
        got = np.array([hostlib.ugk_host_cfg_mix(int(a), int(b), s) for a, b in zip(c, u)], np.uint16)
        assert np.array_equal(got, ref)


def test_ordered_key_is_monotone(hostlib):
    xs = np.sort(np.concatenate([np.random.default_rng(1).standard_normal(2000).astype(np.float32) * 50,
                                 np.array([0.0, -0.0, 88.72284, -46.0], np.float32)]))
    keys = [hostlib.ugk_host_ordered_key(float(x)) for x in xs]
    for a, b, ka, kb in zip(xs[:-1], xs[1:], keys[:-1], keys[1:]):
        assert (ka < kb) == (a < b)
