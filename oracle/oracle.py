"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/decode_step_oracle.c (the CPU restatement of one MaskGIT decode
step, models/unigen.py:407-453 + models/sampling.py:20-46 of apple/ml-unigen) plus a numpy
restatement of the host-side schedule arithmetic (models/sampling.py:49-88,
models/unigen.py:425-444).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  Nothing under ml_unigen_b200/ does.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libug_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc) if needed and return the path of the shared object."""
    src = os.path.join(_HERE, "decode_step_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.ug_oracle_expf.restype = ctypes.c_float
        L.ug_oracle_expf.argtypes = [ctypes.c_float]
        L.ug_oracle_logf.restype = ctypes.c_float
        L.ug_oracle_logf.argtypes = [ctypes.c_float]
        L.ug_oracle_race_q.restype = ctypes.c_float
        L.ug_oracle_race_q.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
        L.ug_oracle_select_u.restype = ctypes.c_float
        L.ug_oracle_select_u.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
        L.ug_oracle_expf_max_ulp.restype = ctypes.c_double
        L.ug_oracle_expf_max_ulp.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_uint32]
        L.ug_oracle_logf_max_ulp.restype = ctypes.c_double
        L.ug_oracle_logf_max_ulp.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_uint32]
        L.ug_oracle_export_noise.restype = None
        L.ug_oracle_export_noise.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_void_p]
        L.ug_oracle_decode_step.restype = ctypes.c_int
        L.ug_oracle_decode_step.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_float]
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().ug_oracle_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def export_noise(seed: int, offset: int, B: int, N: int, V: int, batch_offset: int = 0,
                 want_q: bool = True, want_u: bool = True):
    """The noise the Philox mode of the decode step uses, as (q[B,N,V] | None, u[B,N] | None)."""
    q = np.empty((B, N, V), np.float32) if want_q else None
    u = np.empty((B, N), np.float32) if want_u else None
    lib().ug_oracle_export_noise(seed, offset, batch_offset, B, N, V, _ptr(q), _ptr(u))
    return q, u


def decode_step(logits_bf16: np.ndarray, ids: np.ndarray, *, mask_token_id: int, mask_len_cap: int,
                temperature: float, q_noise: Optional[np.ndarray] = None, u_noise: Optional[np.ndarray] = None,
                seed: int = 0, offset: int = 0, batch_offset: int = 0,
                uncond_bf16: Optional[np.ndarray] = None, guidance_scale: float = 0.0):
    """One decode step.  logits_bf16: uint16 [B,N,V] (bf16 bit patterns); ids: int64 [B,N].

    Returns dict(ids=new ids_inout, sampled, mask (uint8), sel_prob, conf, rowsum)."""
    assert logits_bf16.dtype == np.uint16 and logits_bf16.ndim == 3
    B, N, V = logits_bf16.shape
    logits_bf16 = np.ascontiguousarray(logits_bf16)
    if uncond_bf16 is not None:
        uncond_bf16 = np.ascontiguousarray(uncond_bf16)
        assert uncond_bf16.shape == logits_bf16.shape and uncond_bf16.dtype == np.uint16
    ids = np.ascontiguousarray(ids.astype(np.int64)).copy()
    sampled = np.empty((B, N), np.int64)
    mask = np.empty((B, N), np.uint8)
    sel = np.empty((B, N), np.float32)
    conf = np.empty((B, N), np.float32)
    rowsum = np.empty((B, N), np.float32)
    if q_noise is not None:
        q_noise = np.ascontiguousarray(q_noise.reshape(B, N, V).astype(np.float32))
    if u_noise is not None:
        u_noise = np.ascontiguousarray(u_noise.reshape(B, N).astype(np.float32))
    rc = lib().ug_oracle_decode_step(_ptr(logits_bf16), _ptr(uncond_bf16), float(guidance_scale),
                                     _ptr(ids), _ptr(sampled), _ptr(mask), _ptr(sel), _ptr(conf), _ptr(rowsum),
                                     _ptr(q_noise), _ptr(u_noise), seed, offset, batch_offset,
                                     B, N, V, mask_token_id, int(mask_len_cap), float(np.float32(temperature)))
    if rc != 0:
        raise ValueError(f"ug_oracle_decode_step rc={rc}")
    return dict(ids=ids, sampled=sampled, mask=mask, sel_prob=sel, conf=conf, rowsum=rowsum)


# ----------------------------------------------------------------------------------------------
# Host-side schedule arithmetic (numpy float32 restatement)
# ----------------------------------------------------------------------------------------------
def cosine_schedule_f32(ratio: float) -> np.float32:
    """models/sampling.py:49-50 on a 0-d float32 tensor: cos(float32(ratio) * fl32(pi*0.5))."""
    t = np.float32(ratio)
    return np.cos(np.float32(t * np.float32(math.pi * 0.5)), dtype=np.float32)


def mask_len_cap(N: int, mask_ratio) -> int:
    """floor(N * mask_ratio) in float32 (models/unigen.py:436); may be -1 at ratio 1."""
    return int(np.floor(np.float32(N) * np.float32(mask_ratio)))


def step_schedule(timesteps: int, N: int, temperature: float = 1.0, schedule=cosine_schedule_f32):
    """Per-step (mask_len_cap, annealed temperature) exactly as models/unigen.py:425-426, 444 compute them.

    The temperature compounds in Python double arithmetic (unigen.py:444 overwrites the variable)."""
    out = []
    temp = float(temperature)
    for step in range(timesteps):
        ratio = 1.0 * (step + 1) / timesteps
        cap = mask_len_cap(N, schedule(ratio))
        temp = temp * (1.0 - ratio)
        out.append((cap, temp))
    return out


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bf16 bit patterns (uint16)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


# ----------------------------------------------------------------------------------------------
# Platform-independent synthetic inputs for the full-size goldens (tests/golden/decode_full_*.npz store only seeds, ids
# and masks; the test box regenerates logits and noise).  Everything is integer RNG (numpy PCG64 `integers`) followed by
# exact IEEE operations, so the bits do not depend on libm / MKL / vector ISA of the host.
# ----------------------------------------------------------------------------------------------
def synthetic_logits_bits(seed: int, step: int, B: int, N: int, V: int, scale: float) -> np.ndarray:
    """bf16 bit patterns [B,N,V] of `scale` x an Irwin-Hall(4) variate (approximately N(0,1), |x| <= 3.46)."""
    rng = np.random.default_rng([int(seed), int(step), 0x10617])
    s = np.zeros((B, N, V), np.int32)
    for _ in range(4):
        s += rng.integers(0, 65536, size=(B, N, V), dtype=np.uint16)
    x = (s - np.int32(2 * 65535)).astype(np.float32)                       # exact: |value| < 2^18
    sigma = np.float32(math.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0))
    return f32_to_bf16_bits(x * np.float32(np.float32(scale) / sigma))     # one float32 multiply, one rounding to bf16


def integer_noise(seed: int, step: int, B: int, N: int, V: int):
    """(q [B,N,V], u [B,N]) float32, exactly representable: k * 2^-24 with k uniform on [1, 2^24).  q stands in for the Exp(1)
    tensor of torch.multinomial (the contract is `id = argmax p_i / q_i` for whatever q is injected), u for gumbel_noise's."""
    rng = np.random.default_rng([int(seed), int(step), 0x9015E])
    q = rng.integers(1, 1 << 24, size=(B, N, V), dtype=np.uint32).astype(np.float32) * np.float32(2.0 ** -24)
    u = rng.integers(1, 1 << 24, size=(B, N), dtype=np.uint32).astype(np.float32) * np.float32(2.0 ** -24)
    return q, u


def noise_checksum(a: np.ndarray) -> int:
    """Order-dependent 64-bit checksum of a float32 array's bit patterns (guards generator-replayed noise across hosts)."""
    w = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64).ravel()
    idx = (np.arange(w.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(32)
    return int(np.bitwise_xor.reduce(w * (idx | np.uint64(1))) ^ np.uint64(w.size))
