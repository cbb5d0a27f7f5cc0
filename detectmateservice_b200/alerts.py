"""Host-side alert materialisation: DetectorSchema bytes for the (rare) anomalous records.

The decision -- which monitored fields of which records hold unknown values -- is made on
the GPU (unknown-field bitmask per anomalous record).  What remains for the host is text
formatting: pull the flagged fields' value strings out of the record so the alert can say
``Unknown value: '<value>'`` like the reference's NewValueDetector does
(/root/reference/docs/getting_started.md:510).  This module re-tokenises ONLY those records
(R-tok rules) and never decides anything.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

SP, DQ, SQ, EQ = 0x20, 0x22, 0x27, 0x3D


def record_fields(line: bytes, wanted: Sequence[bytes]) -> Dict[bytes, bytes]:
    """First-occurrence values of the wanted keys in one record (R-tok L2-L6)."""
    out: Dict[bytes, bytes] = {}
    want = set(wanted)
    n = len(line)
    inq = False
    for p in range(n):
        c = line[p]
        if not inq and (p == 0 or line[p - 1] == SP or line[p - 1] == SQ):
            q = p
            while q < n and line[q] not in (EQ, SP, DQ, SQ):
                q += 1
            if q < n and line[q] == EQ and q > p:
                key = line[p:q]
                if key in want and key not in out:
                    e, vq = q + 1, False
                    while e < n and not (line[e] == SP and not vq):
                        if line[e] == DQ:
                            vq = not vq
                        e += 1
                    out[key] = line[q + 1:e]
                    if len(out) == len(want):
                        break
        if c == DQ:
            inq = not inq
    return out


def record_time(line: bytes) -> Optional[int]:
    """Seconds of the audit stamp ``msg=audit(<sec>.<ms>:<serial>)`` (R-tok L7) or None."""
    v = record_fields(line, [b"msg"]).get(b"msg")
    if not v or not v.startswith(b"audit("):
        return None
    j = len(v)
    for i in range(6, len(v)):
        if v[i] in (0x3A, 0x29):
            j = i
            break
    try:
        return int(float(v[6:j]))
    except ValueError:
        return None


def record_at(msg: bytes, offset: int) -> bytes:
    if isinstance(msg, memoryview):                  # a lent receive buffer: look in 4 KiB windows
        pos, n = offset, len(msg)
        while pos < n:
            win = bytes(msg[pos:pos + 4096])
            k = win.find(b"\n")
            if k >= 0:
                return bytes(msg[offset:pos + k])
            pos += len(win)
        return bytes(msg[offset:])
    end = msg.find(b"\n", offset)
    return bytes(msg[offset:] if end < 0 else msg[offset:end])      # bytes or bytearray


def alert_text(value: bytes) -> str:
    return "Unknown value: '%s'" % value.decode("utf-8", "replace")


def count_records(data) -> int:
    """Records of a raw-line message (R-tok L1): every '\n' ends one, a non-empty tail is one."""
    b = bytes(data)
    return b.count(b"\n") + (1 if b and not b.endswith(b"\n") else 0)
