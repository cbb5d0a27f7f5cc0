"""ctypes binding of oracle/c/dm_oracle.c.  TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdmoracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "c", f) for f in ("dm_oracle.c", "dm_oracle_bench.c")]
    if force or not os.path.exists(_SO) or any(
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO) for src in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.dmo_create.restype = C.c_void_p
        L.dmo_create.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_uint32)]
        L.dmo_destroy.argtypes = [C.c_void_p]
        L.dmo_process.restype = C.c_uint64
        L.dmo_process.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
        for fn in ("dmo_known_count", "dmo_unknown_count"):
            getattr(L, fn).restype = C.c_uint64
            getattr(L, fn).argtypes = [C.c_void_p, C.c_int]
        for fn in ("dmo_total_lines", "dmo_total_anomalies"):
            getattr(L, fn).restype = C.c_uint64
            getattr(L, fn).argtypes = [C.c_void_p]
        L.dmo_export_known.restype = C.c_uint64
        L.dmo_export_known.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        L.dmo_fp64.restype = C.c_uint64
        L.dmo_fp64.argtypes = [C.c_char_p, C.c_uint32]
        L.dmo_bench_threads.restype = C.c_int
        L.dmo_bench_threads.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                        C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def fp64(value: bytes) -> int:
    return int(lib().dmo_fp64(value, len(value)))


def bench_threads(keys: Sequence[bytes], train, detect, n_threads: int, min_seconds: float = 2.0, n_samples: int = 5):
    """Records/s of the C restatement on `n_threads` pinned worker threads (oracle/c/dm_oracle_bench.c):
    returns (threads used, [rate per sample], alerts seen)."""
    keys = [bytes(k) for k in keys]
    blob = b"".join(keys)
    lens = (C.c_uint32 * len(keys))(*[len(k) for k in keys])
    tr = np.frombuffer(train, dtype=np.uint8) if not isinstance(train, np.ndarray) else train
    de = np.frombuffer(detect, dtype=np.uint8) if not isinstance(detect, np.ndarray) else detect
    rates = (C.c_double * n_samples)()
    anom = C.c_uint64()
    used = lib().dmo_bench_threads(len(keys), blob, lens, tr.ctypes.data, int(tr.size), de.ctypes.data, int(de.size),
                                   int(n_threads), float(min_seconds), int(n_samples), rates, C.byref(anom))
    if used < 0:
        raise RuntimeError(f"dmo_bench_threads failed: {used}")
    return used, [float(r) for r in rates], int(anom.value)


class NativeOracle:
    """Raw-line NewValueDetector over a fixed list of monitored keys (global scope)."""

    def __init__(self, keys: Sequence[bytes]):
        self.keys = [bytes(k) for k in keys]
        blob = b"".join(self.keys)
        lens = (C.c_uint32 * len(self.keys))(*[len(k) for k in self.keys])
        self._h = lib().dmo_create(len(self.keys), blob, lens)
        if not self._h:
            raise ValueError("dmo_create failed")

    def close(self) -> None:
        if self._h:
            lib().dmo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process(self, buf, n_train_lines: int = 0, want_masks: bool = False
                ) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
        """buf: bytes or a uint8 numpy array.  Returns (flags u8, scores f32, masks u32|None)."""
        arr = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
        n = int(arr.size)
        nl = int(np.count_nonzero(arr == 10))
        cap = nl + 1
        flags = np.zeros(cap, dtype=np.uint8)
        scores = np.zeros(cap, dtype=np.float32)
        masks = np.zeros(cap, dtype=np.uint32) if want_masks else None
        got = lib().dmo_process(self._h, arr.ctypes.data, n, int(n_train_lines),
                                flags.ctypes.data, scores.ctypes.data,
                                masks.ctypes.data if want_masks else None)
        got = int(got)
        return flags[:got], scores[:got], (masks[:got] if want_masks else None)

    def known_count(self, k: int) -> int:
        return int(lib().dmo_known_count(self._h, k))

    def unknown_count(self, k: int) -> int:
        return int(lib().dmo_unknown_count(self._h, k))

    def known_values(self, k: int) -> List[bytes]:
        need = int(lib().dmo_export_known(self._h, k, None, 0))
        buf = (C.c_uint8 * max(need, 1))()
        lib().dmo_export_known(self._h, k, buf, need)
        raw = bytes(buf)[:need]
        out, off = [], 0
        while off < need:
            ln = int.from_bytes(raw[off:off + 4], "little")
            out.append(raw[off + 4:off + 4 + ln])
            off += 4 + ln
        return out
