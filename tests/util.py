"""Shared helpers for the parity tests: fuzz inputs and oracle drivers."""
import numpy as np

from oracle.native import NativeOracle

FUZZ_KEYS = [b"a", b"ab", b"type", b"res", b"k1"]          # 5 distinct last bytes: candidate pre-filter off
FUZZ_KEYS_FEW = [b"a", b"ab", b"type", b"sa", b"ra"]        # 3 distinct last bytes: pre-filter on


def fuzz_lines(seed: int, n_lines: int, max_len: int = 600, alphabet: bytes = b"ab=  '\"typeresk1xq") -> bytes:
    """Random records over an alphabet rich in '=', space, both quote characters and the
    letters of the monitored keys -- stresses quote parity, field starts and duplicates."""
    r = np.random.Generator(np.random.PCG64(seed))
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    parts = []
    for _ in range(n_lines):
        mode = r.integers(0, 10)
        if mode == 0:
            ln = 0
        elif mode < 6:
            ln = int(r.integers(1, 64))
        elif mode < 9:
            ln = int(r.integers(64, 300))
        else:
            ln = int(r.integers(300, max_len + 1))
        parts.append(alpha[r.integers(0, alpha.size, ln)].tobytes())
    return b"\n".join(parts) + b"\n"


def oracle_run(keys, train_msg: bytes, n_train: int, detect_msgs):
    """Train the C oracle on the first n_train records of train_msg (the rest of it is
    detected), then detect every message of detect_msgs.  Returns list of (flags, scores, masks)."""
    o = NativeOracle(keys)
    out = [o.process(train_msg, n_train, want_masks=True)]
    for m in detect_msgs:
        out.append(o.process(m, 0, want_masks=True))
    return o, out
