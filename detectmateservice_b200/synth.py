"""Synthetic audit-log workloads of BASELINE.json's configs.

Distributions mirror /root/reference/tests/library_integration/audit.log (15 record
types with the empirical counts, 312 pids, 306 sessions, 6 executables, 6 terminals,
2 accounts, res success:failed = 99:1).  Deterministic: numpy PCG64, seed 20260921.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

SEED = 20260921
MONITORED_KEYS = ["type", "exe", "terminal", "acct", "res"]     # K = 5
LINE_BYTES = 256
BATCH_LINES = 65536

_TYPES = [("CRED_ACQ", 308), ("USER_START", 306), ("USER_ACCT", 305), ("LOGIN", 304), ("CRED_DISP", 302),
          ("USER_END", 302), ("SERVICE_START", 241), ("SERVICE_STOP", 230), ("AVC", 4), ("SYSCALL", 4),
          ("PROCTITLE", 4), ("USER_LOGIN", 3), ("USER_AUTH", 1), ("USER_CMD", 1), ("CRED_REFR", 1)]
_EXES = ["/usr/sbin/cron", "/lib/systemd/systemd", "/usr/sbin/sshd", "/bin/su", "/usr/bin/sudo", "/sbin/apparmor_parser"]
_TERMS = ["cron", "?", "/dev/pts/0", "/dev/pts/1", "pts/1", "ssh"]
_ACCTS = ["root", "jhall"]
_OPS = ["PAM:accounting", "PAM:setcred", "PAM:session_open", "PAM:session_close", "PAM:authentication", "login"]


class AuditSynth:
    def __init__(self, seed: int = SEED, anomaly_rate: float = 1e-3):
        self.rng = np.random.Generator(np.random.PCG64(seed))
        w = np.array([c for _, c in _TYPES], dtype=np.float64)
        self.type_p = w / w.sum()
        self.pids = self.rng.choice(np.arange(300, 40000), size=312, replace=False)
        self.sess = self.rng.choice(np.arange(1, 5000), size=306, replace=False)
        self.novel = ["NOVEL_%04d" % i for i in range(1024)]
        self.anomaly_rate = anomaly_rate
        self.serial = 375
        self.sec = 1642723741

    # ---------------------------------------------------------------- fixed 256-byte records
    def batch(self, n_lines: int, inject: bool, line_bytes: int = LINE_BYTES) -> Tuple[bytes, np.ndarray]:
        """n_lines records of exactly line_bytes bytes (incl. '\\n').  Returns (message,
        injected uint8[n]) where injected marks records that received a novel value."""
        r = self.rng
        t_idx = r.choice(len(_TYPES), size=n_lines, p=self.type_p)
        pid = self.pids[r.integers(0, 312, n_lines)]
        ses = self.sess[r.integers(0, 306, n_lines)]
        exe = r.integers(0, len(_EXES), n_lines)
        term = r.integers(0, len(_TERMS), n_lines)
        acct = r.integers(0, len(_ACCTS), n_lines)
        op = r.integers(0, len(_OPS), n_lines)
        res_fail = r.random(n_lines) < 0.01
        ms = r.integers(0, 1000, n_lines)
        uid = r.integers(0, 3, n_lines) * 33
        inj = (r.random(n_lines) < self.anomaly_rate) if inject else np.zeros(n_lines, dtype=bool)
        inj_field = r.integers(0, 5, n_lines)
        inj_val = r.integers(0, len(self.novel), n_lines)
        out: List[str] = []
        for i in range(n_lines):
            T = _TYPES[t_idx[i]][0]
            e = '"%s"' % _EXES[exe[i]]
            tm = _TERMS[term[i]]
            ac = '"%s"' % _ACCTS[acct[i]]
            rs = "failed'" if res_fail[i] else "success'"
            if inj[i]:
                nv = self.novel[inj_val[i]]
                f = inj_field[i]
                if f == 0:
                    T = nv
                elif f == 1:
                    e = '"/opt/%s"' % nv
                elif f == 2:
                    tm = nv
                elif f == 3:
                    ac = '"%s"' % nv
                else:
                    rs = nv + "'"
            self.serial += 1
            if (i & 7) == 0:
                self.sec += 1
            body = ("type=%s msg=audit(%d.%03d:%d): pid=%d uid=%d auid=4294967295 ses=%d "
                    "msg='op=%s acct=%s exe=%s hostname=? addr=? terminal=%s res=%s") % (
                        T, self.sec, ms[i], self.serial, pid[i], uid[i], ses[i], _OPS[op[i]], ac, e, tm, rs)
            room = line_bytes - 1 - len(body)
            if room >= 6:
                body += " pad=" + "x" * (room - 5)
            elif room > 0:
                body += " " * room
            elif room < 0:
                body = body[:line_bytes - 1]
            out.append(body)
        msg = ("\n".join(out) + "\n").encode("ascii")
        return msg, inj.astype(np.uint8)

    # ---------------------------------------------------------------- variable length (config 5)
    def batch_varlen(self, n_lines: int, inject: bool, min_len: int = 32, max_len: int = 4096
                     ) -> Tuple[bytes, np.ndarray]:
        """Records of mixed length: 70 % log-uniform[32,256], 25 % [256,1024], 5 % [1024,4096]
        bytes (incl. '\\n'); long records carry extra kN= fields and quoted values with spaces."""
        r = self.rng
        u = r.random(n_lines)
        lo = np.where(u < 0.70, 32.0, np.where(u < 0.95, 256.0, 1024.0))
        hi = np.where(u < 0.70, 256.0, np.where(u < 0.95, 1024.0, 4096.0))
        lens = np.exp(r.uniform(np.log(lo), np.log(hi))).astype(np.int64)
        lens = np.clip(lens, min_len, max_len)
        t_idx = r.choice(len(_TYPES), size=n_lines, p=self.type_p)
        exe = r.integers(0, len(_EXES), n_lines)
        term = r.integers(0, len(_TERMS), n_lines)
        acct = r.integers(0, len(_ACCTS), n_lines)
        res_fail = r.random(n_lines) < 0.01
        inj = (r.random(n_lines) < self.anomaly_rate) if inject else np.zeros(n_lines, dtype=bool)
        inj_field = r.integers(0, 5, n_lines)
        inj_val = r.integers(0, len(self.novel), n_lines)
        filler = r.integers(0, 1 << 30, n_lines)
        out: List[str] = []
        for i in range(n_lines):
            target = int(lens[i]) - 1
            T = _TYPES[t_idx[i]][0]
            fields = ["exe=\"%s\"" % _EXES[exe[i]], "terminal=%s" % _TERMS[term[i]],
                      "acct=\"%s\"" % _ACCTS[acct[i]], "res=%s" % ("failed" if res_fail[i] else "success")]
            if inj[i]:
                nv = self.novel[inj_val[i]]
                f = inj_field[i]
                if f == 0:
                    T = nv
                else:
                    fields[f - 1] = fields[f - 1].split("=")[0] + "=" + nv
            self.serial += 1
            line = "type=%s msg=audit(%d.000:%d):" % (T, self.sec, self.serial)
            if len(line) > target:
                line = "type=%s msg=audit(1.000:1): k=v" % T
            j = 0
            # monitored fields first while they fit, then filler fields, some quoted with spaces
            for f in fields:
                if len(line) + 1 + len(f) <= target:
                    line += " " + f
            while len(line) + 8 <= target:
                if j % 3 == 2 and len(line) + 24 <= target:
                    f = 'k%d="v %x res=quoted x"' % (j, (filler[i] + j) & 0xFFFF)
                else:
                    f = "k%d=%x" % (j, (filler[i] * (j + 1)) & 0xFFFFFF)
                if len(line) + 1 + len(f) > target:
                    break
                line += " " + f
                j += 1
            if len(line) < target:
                line += " " + "p" * (target - len(line) - 1) if target - len(line) >= 1 else ""
            out.append(line[:max(target, 0)] if len(line) > target else line)
        msg = ("\n".join(out) + "\n").encode("ascii")
        return msg, inj.astype(np.uint8)


def config2_stream(n_lines: int = 1_000_000, batch_lines: int = BATCH_LINES, train_lines: int = BATCH_LINES,
                   seed: int = SEED, line_bytes: int = LINE_BYTES):
    """Yield (message, n_train_lines_in_message) for BASELINE config 2: the first
    train_lines records are an anomaly-free training window, the rest carry p=1e-3 injections."""
    g = AuditSynth(seed)
    done = 0
    while done < n_lines:
        n = min(batch_lines, n_lines - done)
        n_train = max(0, min(n, train_lines - done))
        if n_train == n or n_train == 0:
            msg, _ = g.batch(n, inject=(n_train == 0), line_bytes=line_bytes)
        else:
            a, _ = g.batch(n_train, inject=False, line_bytes=line_bytes)
            b, _ = g.batch(n - n_train, inject=True, line_bytes=line_bytes)
            msg = a + b
        yield msg, n_train
        done += n
