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


def lookalike_lines(seed: int, n_lines: int, keys=(b"key", b"type", b"res"), max_tokens: int = 260) -> bytes:
    """Records built from TOKENS rather than bytes: true fields, duplicates of them, quoted look-alikes
    (`q="x res=v y"`), apostrophe-led field starts, stray quotes that flip the parity for the rest of the
    record, dense runs of short fields (more than four '=' in 32 bytes) and long fillers -- up to a few KiB
    per record with dozens of candidates each.  Exercises the chained re-check of the stream kernel
    (dmx_verify_chain: carried state across batches, record starts inside a batch, fall-backs)."""
    r = np.random.Generator(np.random.PCG64(seed))
    vals = [b"aaa", b"bbb", b"cccc", b"success", b"quoted", b"xx1", b"xx2", b"new%d"]
    out = []
    for _ in range(n_lines):
        nt = int(r.integers(0, 8)) if r.random() < 0.6 else int(r.integers(8, max_tokens))
        toks = []
        dense_ok = r.random() < 0.04                               # (a dense row switches the chained re-check off around it)
        for _t in range(nt):
            u = r.random()
            if 0.64 <= u < 0.70 and not dense_ok:
                u = 0.9
            key = keys[int(r.integers(0, len(keys)))]
            v = vals[int(r.integers(0, len(vals)))]
            if b"%d" in v:
                v = v % int(r.integers(0, 40))
            if u < 0.30:
                toks.append(key + b"=" + v)
            elif u < 0.50:
                toks.append(b'q%d="v %s=%s x"' % (int(r.integers(0, 9)), key, v))
            elif u < 0.56:
                toks.append(b"'" + key + b"=" + v)
            elif u < 0.60:
                toks.append(b'"')                                  # stray quote: everything behind it is quoted
            elif u < 0.64:
                toks.append(b"x" + key + b"=" + v)                 # no field start
            elif u < 0.70:
                toks.append(b" ".join(b"%c=%d" % (97 + int(r.integers(0, 26)), int(r.integers(0, 10))) for _ in range(8)))
            elif u < 0.75:
                toks.append(key + b'="' + v + b' ' + v + b'"')     # quoted value with a space
            else:
                toks.append(b"f%d=%x" % (int(r.integers(0, 99)), int(r.integers(0, 1 << 24))))
        out.append(b" ".join(toks))
    return b"\n".join(out) + b"\n"


def oracle_run(keys, train_msg: bytes, n_train: int, detect_msgs):
    """Train the C oracle on the first n_train records of train_msg (the rest of it is
    detected), then detect every message of detect_msgs.  Returns list of (flags, scores, masks)."""
    o = NativeOracle(keys)
    out = [o.process(train_msg, n_train, want_masks=True)]
    for m in detect_msgs:
        out.append(o.process(m, 0, want_masks=True))
    return o, out
