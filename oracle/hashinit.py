"""ORACLE — test infrastructure only.  Deterministic pseudo-normal weight generator.

numpy restatement + ctypes binding of oracle/hashinit.c (see that file for the definition).
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
import torch

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_IH4_STD = math.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0)    # std of the sum of four u16 uniforms
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def fnv1a64(name: str) -> int:
    h = 0xCBF29CE484222325
    for b in name.encode("utf-8"):
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def tensor_seed(global_seed: int, name: str) -> int:
    return (fnv1a64(name) ^ ((global_seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF


def scale_for_std(std: float) -> np.float32:
    return np.float32(std / _IH4_STD)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def hash_normal_numpy(n: int, seed: int, std: float, mean: float = 0.0) -> np.ndarray:
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        x = _splitmix64(np.uint64(seed) + i * _GOLDEN)
    m = np.uint64(0xFFFF)
    s = ((x & m) + ((x >> np.uint64(16)) & m) + ((x >> np.uint64(32)) & m) + ((x >> np.uint64(48)) & m))
    s = s.astype(np.int64) - 131070
    return (np.float32(mean) + s.astype(np.float32) * scale_for_std(std)).astype(np.float32)


def hash_u8_numpy(n: int, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        x = _splitmix64(np.uint64(seed) + i * _GOLDEN)
    return (x >> np.uint64(56)).astype(np.uint8)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_hashinit.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-s", "-C", _HERE])
        lib = ctypes.CDLL(path)
        lib.teo_oracle_hash_normal_f32.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64,
                                                   ctypes.c_float, ctypes.c_float]
        lib.teo_oracle_hash_u8.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64]
        _LIB = lib
    return _LIB


def hash_normal(shape, seed: int, std: float, mean: float = 0.0) -> torch.Tensor:
    """fp32 tensor of the given shape (C fast path)."""
    out = torch.empty(shape, dtype=torch.float32)
    _lib().teo_oracle_hash_normal_f32(out.data_ptr(), out.numel(), ctypes.c_uint64(seed),
                                      ctypes.c_float(float(scale_for_std(std))), ctypes.c_float(mean))
    return out


def hash_u8(shape, seed: int) -> torch.Tensor:
    out = torch.empty(shape, dtype=torch.uint8)
    _lib().teo_oracle_hash_u8(out.data_ptr(), out.numel(), ctypes.c_uint64(seed))
    return out
