"""MatcherParser fused in front of the detector: log_format header
extraction + `<*>` template matching.  Regex oracle (oracle/rfmt.py) vs the host mirror vs
the kernel source on the CPU emulator, and (GPU tier) the real library / component.
Pinned known answer: the nginx example of docs/getting_started.md:395-435,498-510."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from detectmateservice_b200 import wire
from detectmateservice_b200.logformat import LogFormat, load_templates, norm_flags
from oracle.nvd import NewValueDetectorOracle
from oracle.rfmt import FormatParser

NGINX = '<IP> - - [<Time>] "<Method> <URL> <Protocol>" <Status> <Bytes> "<Referer>" "<UserAgent>"'
AUDIT = "type=<type> msg=audit(<Time>): <Content>"
NGINX_CFG = {"detectors": {"NewValueDetector": {                      # container/config/detector_config.yaml
    "method_type": "new_value_detector", "data_use_training": 2, "auto_config": False,
    "global": {"global_instance": {"header_variables": [{"pos": "URL"}]}}}}}


def nginx_line(url, ip="172.18.0.1", status=404):
    return ('%s - - [18/Mar/2026:15:39:43 +0000] "GET %s HTTP/1.1" %d 162 "-" "curl/8.5.0"' % (ip, url, status)).encode()


def audit_templates(golden_dir):
    return load_templates(os.path.join(golden_dir, "audit_templates.txt"))


class FormatOracle:
    """FormatParser (regex) + NewValueDetectorOracle: the CPU restatement of parser + detector."""

    def __init__(self, cfg, log_format, templates=(), norm=(False, False, False)):
        self.parser = FormatParser(log_format, templates, remove_spaces=norm[0], remove_punctuation=norm[1], lowercase=norm[2])
        self.nvd = NewValueDetectorOracle(config=cfg, clock=lambda: 1773848383)

    def process_lines(self, buf):
        lines = buf.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        flags, scores, alerts, bad = [], [], [], 0
        for ln in lines:
            rec = self.parser.parse_line(ln)
            if rec is None:
                bad += 1
                rec = {"EventID": None, "variables": [], "logFormatVariables": {}}
            f, s, a = self.nvd.step(rec)
            flags.append(int(f)); scores.append(float(s)); alerts.append(a)
        return flags, scores, alerts, bad


# ------------------------------------------------------------------------------------------ CPU tier
def test_docs_nginx_golden_oracle_and_host():
    """docs/getting_started.md: train on /hello and /world, /foobar alerts with this exact text."""
    buf = b"\n".join(nginx_line(u) for u in ("/hello", "/world", "/foobar", "/hello")) + b"\n"
    orc = FormatOracle(NGINX_CFG, NGINX)
    flags, scores, alerts, bad = orc.process_lines(buf)
    assert flags == [0, 0, 1, 0] and scores == [0.0, 0.0, 1.0, 0.0] and bad == 0
    assert alerts[2] == {"Global - URL": "Unknown value: '/foobar'"}
    lf = LogFormat(NGINX)
    eid, variables, lfv = lf.parse(nginx_line("/foobar"))
    assert (eid, variables) == (-1, []) and lfv["URL"] == b"/foobar" and lfv["Time"] == b"18/Mar/2026:15:39:43 +0000"
    assert lfv == FormatParser(NGINX).parse_line(nginx_line("/foobar"))["logFormatVariables"]


def test_host_mirror_matches_regex_on_audit_log(golden_dir):
    tm = audit_templates(golden_dir)
    a, b = FormatParser(AUDIT, tm), LogFormat(AUDIT, tm)
    eids = []
    for ln in open(os.path.join(golden_dir, "audit_sample.log"), "rb").read().split(b"\n")[:-1]:
        r, s = a.parse_line(ln), b.parse(ln)
        assert (r["EventID"], r["variables"], r["logFormatVariables"]) == s
        eids.append(s[0])
    assert set(eids) >= {0, 1, 2} and -1 not in eids


def fuzz_formats(r, n):
    """Random chains over a tiny alphabet so that literals recur inside captured text."""
    alpha = b"ab =:"
    out = []
    for _ in range(n):
        k = int(r.integers(1, 5))
        lits = [bytes(alpha[int(i)] for i in r.integers(0, len(alpha), int(r.integers(1, 4)))) for _ in range(k + 1)]
        if r.random() < 0.4:
            lits[0] = b""
        ends = r.random() < 0.5
        out.append((lits, ends))
    return out


def fuzz_text(r, lits, ends):
    alpha = b"ab =:xy"
    if r.random() < 0.25:                                        # arbitrary text, mostly non-matching
        return bytes(alpha[int(i)] for i in r.integers(0, len(alpha), int(r.integers(0, 24))))
    t = b""
    for i, lit in enumerate(lits):
        t += lit
        if i < len(lits) - 1 or ends:
            t += bytes(alpha[int(j)] for j in r.integers(0, len(alpha), int(r.integers(0, 9))))
    if r.random() < 0.15:
        t = t[:int(r.integers(0, len(t) + 1))]
    return t


def test_sequential_matcher_equals_regex_fuzz():
    r = np.random.Generator(np.random.PCG64(5))
    n_match = 0
    for lits, ends in fuzz_formats(r, 300):
        names = ["c%d" % i for i in range(len(lits) - (0 if ends else 1))]
        fmt = b"".join(l + (b"<%s>" % names[i].encode() if i < len(names) else b"") for i, l in enumerate(lits))
        tmpl = b"<*>".join(lits) + (b"<*>" if ends else b"")
        a, b = FormatParser(fmt), LogFormat(fmt)
        ta, tb = FormatParser("<Content>", [tmpl]), LogFormat("<Content>", [tmpl])
        for _ in range(40):
            text = fuzz_text(r, lits, ends)
            ra, rb = a.parse_line(text), b.parse(text)
            assert (ra is None) == (rb is None), (fmt, text)
            if ra is not None:
                n_match += 1
                assert ra["logFormatVariables"] == rb[2], (fmt, text)
            va, vb = ta.parse_line(text), tb.parse(text)
            assert (va["EventID"], va["variables"]) == (vb[0], vb[1]), (tmpl, text)
    assert n_match > 2000


NORMS = [(True, True, True), (True, False, False), (False, True, False), (False, False, True), (True, False, True)]


def norm_fuzz_template(r):
    """A template over an alphabet with spaces, punctuation and capitals, whose literals keep at least one
    letter (so that normalisation never empties the text between two wildcards)."""
    alpha = b"aB =:c-"
    k = int(r.integers(1, 4))
    lits = []
    for _ in range(k + 1):
        raw = bytes(alpha[int(i)] for i in r.integers(0, len(alpha), int(r.integers(1, 4))))
        at = int(r.integers(0, len(raw) + 1))
        lits.append(raw[:at] + (b"a", b"B", b"c")[int(r.integers(0, 3))] + raw[at:])
    if r.random() < 0.3:
        lits[0] = b""
    ends = r.random() < 0.5
    if not ends and r.random() < 0.3:
        lits[-1] = b"'"                                          # punctuation only: vanishes under remove_punctuation
    return lits, ends


def norm_fuzz_text(r, lits, ends):
    alpha = b"aB =:c-xY\t.A"
    t = b""
    for i, lit in enumerate(lits):
        t += lit if r.random() < 0.8 else lit.swapcase().replace(b" ", b"  ")
        if i < len(lits) - 1 or ends:
            t += bytes(alpha[int(j)] for j in r.integers(0, len(alpha), int(r.integers(0, 9))))
    if r.random() < 0.2:
        t = bytes(alpha[int(i)] for i in r.integers(0, len(alpha), int(r.integers(0, 24))))
    return t


def test_normalisation_switches_host_mirror_equals_regex(golden_dir):
    """R-norm (remove_spaces / remove_punctuation / lowercase): the sequential matcher on the normalised
    text against Python's re on the same; then the reference's audit templates with all three on."""
    from detectmateservice_b200.logformat import normalise
    assert normalise(b" A-b\tC\r\n.d_E ", norm_flags(True, True, True)) == b"abcde"
    assert normalise(b" A-b\tC ", norm_flags(True, False, False)) == b"A-bC"
    assert normalise(b" A-b\tC\x80\xc3\x84", norm_flags(False, True, True)) == b" ab\tc\x80\xc3\x84"
    r = np.random.Generator(np.random.PCG64(7))
    n_match = 0
    for it in range(300):
        lits, ends = norm_fuzz_template(r)
        norm = NORMS[it % len(NORMS)]
        tmpl = b"<*>".join(lits) + (b"<*>" if ends else b"")
        ta = FormatParser("<Content>", [tmpl], remove_spaces=norm[0], remove_punctuation=norm[1], lowercase=norm[2])
        tb = LogFormat("<Content>", [tmpl], flags=norm_flags(*norm))
        for _ in range(40):
            text = norm_fuzz_text(r, lits, ends)
            va, vb = ta.parse_line(text), tb.parse(text)
            assert (va["EventID"], va["variables"]) == (vb[0], vb[1]), (tmpl, text, norm)
            n_match += va["EventID"] == 0
    assert n_match > 3000
    tm = audit_templates(golden_dir)
    a, b = FormatParser(AUDIT, tm, remove_spaces=True, remove_punctuation=True, lowercase=True), LogFormat(AUDIT, tm, flags=7)
    plain = LogFormat(AUDIT, tm)
    for ln in open(os.path.join(golden_dir, "audit_sample.log"), "rb").read().split(b"\n")[:-1]:
        x, y = a.parse_line(ln), b.parse(ln)
        assert (x["EventID"], x["variables"], x["logFormatVariables"]) == y
        assert y[0] == plain.parse(ln)[0] and y[2] == plain.parse(ln)[2]     # same template, verbatim header
    first = b.parse(b"type=USER_ACCT msg=audit(1642723741.072:375): pid=10125 uid=0 auid=4294967295 ses=4294967295 "
                    b"msg='op=PAM:accounting acct=\"root\" exe=\"/usr/sbin/cron\" hostname=? addr=? terminal=cron res=success'")
    assert first[1] == [b"10125", b"0", b"4294967295", b"4294967295", b"pamaccounting", b"root", b"usrsbincron", b"", b"",
                        b"cron", b"success"]
    with pytest.raises(ValueError, match="nothing between"):
        LogFormat("<Content>", ["a<*> - <*>b"], flags=norm_flags(True, True, False))
    with pytest.raises(ValueError):
        FormatParser("<Content>", ["a<*> - <*>b"], remove_spaces=True, remove_punctuation=True)


def test_config_errors():
    with pytest.raises(ValueError):
        LogFormat("<A><B> x")
    with pytest.raises(ValueError):
        LogFormat("<A> <A>")
    with pytest.raises(ValueError):
        LogFormat("<A> x", ["a<*><*>b"])
    with pytest.raises(ValueError):
        LogFormat("<A> x", ["a<*>"])                             # templates need a <Content> capture
    assert LogFormat("a<b c>d <e>").header.names == ["e"]          # '<b c>' is literal text


def _monitor_array(mons):
    from detectmateservice_b200 import _lib
    arr = (_lib.Monitor * max(1, len(mons)))()
    for i, m in enumerate(mons):
        arr[i].event_id = m.event_id if m.event_id is not None else 0
        arr[i].has_event = 0 if m.event_id is None else 1
        if m.source == "header":
            kb = m.pos.encode()
            arr[i].source, arr[i].key_len = 0, len(kb)
            for j, c in enumerate(kb):
                arr[i].key[j] = c
        else:
            arr[i].source, arr[i].var_index = 1, m.pos
    return arr


EMU_KERNEL = "lanes"


@pytest.fixture(params=["lanes"])
def emu_kernel(request):
    """The matcher kernel: one thread per record."""
    global EMU_KERNEL
    EMU_KERNEL = request.param
    yield request.param
    EMU_KERNEL = "lanes"


class EmuFormat:
    def __init__(self, cfg, log_format, templates=(), name="NewValueDetector", norm=(False, False, False)):
        import emu_harness
        from detectmateservice_b200.component import parse_monitors, select_component_config
        self.mons = parse_monitors(select_component_config(cfg, name))
        self.lib = C.CDLL(emu_harness.build())
        self.det = emu_harness.EmuDetector([m.key for m in self.mons], table_log2=12)
        L = self.lib
        L.emu_set_format.restype = C.c_char_p
        L.emu_set_format.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_char_p), C.c_uint32]
        L.emu_process_format.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        ts = [t if isinstance(t, bytes) else t.encode() for t in templates]
        fmt = log_format if isinstance(log_format, bytes) else log_format.encode()
        err = L.emu_set_format(_monitor_array(self.mons), len(self.mons), fmt, b"Content", len(ts), (C.c_char_p * max(1, len(ts)))(*ts),
                               norm_flags(*norm))
        if err:
            raise ValueError(err.decode())

    def process(self, buf, n_train):
        cap = buf.count(b"\n") + 2
        f = np.full(cap, 9, np.uint8); s = np.full(cap, -1, np.float32)
        n, na = C.c_uint64(), C.c_uint64()
        assert self.lib.emu_process_format(self.det.h, buf, len(buf), n_train, f.ctypes.data, s.ctypes.data, cap,
                                           C.byref(n), C.byref(na)) == 0
        k = n.value
        assert na.value == int(f[:k].sum())
        masks = {a[0]: a[1] for a in self.det.anomalies()}
        return f[:k].tolist(), s[:k].tolist(), masks


def _mask_of(alerts, mons):
    return sum(1 << i for i, m in enumerate(mons) if m.alert_key in alerts)


def test_emu_kernel_docs_nginx_golden(emu_kernel):
    buf = b"\n".join(nginx_line(u) for u in ("/hello", "/world", "/foobar", "/hello", "/x y")) + b"\n"
    emu = EmuFormat(NGINX_CFG, NGINX)
    f, s, masks = emu.process(buf, 2)
    assert f == [0, 0, 1, 0, 1] and s == [0.0, 0.0, 1.0, 0.0, 1.0] and masks == {2: 1, 4: 1}


AUDIT_CFG = {"detectors": {"NewValueDetector": {
    "method_type": "new_value_detector", "data_use_training": 250, "auto_config": False,
    "global": {"g": {"header_variables": [{"pos": "type"}]}},
    "events": {0: {"pam": {"variables": [{"pos": 4, "name": "op"}, {"pos": 6, "name": "exe"}, {"pos": 9, "name": "terminal"}],
                           "header_variables": [{"pos": "type"}]}},
               1: {"unit": {"variables": [{"pos": 4, "name": "unit"}, {"pos": 30}]}},
               2: {"login": {"variables": [{"pos": 3, "name": "auid"}, {"pos": 7, "name": "res"}]}},
               40: {"never": {"variables": [{"pos": 0}]}}}}}}


def test_emu_kernel_audit_templates(golden_dir, emu_kernel):
    tm = audit_templates(golden_dir)
    buf = open(os.path.join(golden_dir, "audit_sample.log"), "rb").read()
    buf += b"type=WEIRD msg=audit(1.0:1): totally unlike=any template\nnot an audit line at all\n\n"
    buf += b"type=LOGIN msg=audit(1642723741.076:377): pid=1 uid=0 old-auid=4294967295 auid=77777 tty=(none) old-ses=4294967295 ses=65 res=9"
    orc = FormatOracle(AUDIT_CFG, AUDIT, tm)
    wf, ws, wa, bad = orc.process_lines(buf)
    emu = EmuFormat(AUDIT_CFG, AUDIT, tm)
    f, s, masks = emu.process(buf, 250)
    assert f == wf and s == ws
    assert masks == {i: _mask_of(a, emu.mons) for i, a in enumerate(wa) if a}
    assert sum(wf) >= 5 and bad == 2 and wf[-1] == 1 and ws[-1] == 2.0          # the unterminated last record counts
    stats = (C.c_uint64 * 40)()
    emu.lib.emu_get_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    emu.lib.emu_get_stats(emu.det.h, stats)
    assert stats[7] == bad


def test_emu_kernel_fuzz_formats(emu_kernel):
    r = np.random.Generator(np.random.PCG64(17))
    checked = 0
    for lits, ends in fuzz_formats(r, 25):
        names = ["c%d" % i for i in range(len(lits) - (0 if ends else 1))]
        if not names:
            continue
        fmt = b"".join(l + (b"<%s>" % names[i].encode() if i < len(names) else b"") for i, l in enumerate(lits))
        cfg = {"detectors": {"NewValueDetector": {"method_type": "new_value_detector", "data_use_training": 30,
                                                  "global": {"g": {"header_variables": [{"pos": n} for n in names]}}}}}
        lines = [fuzz_text(r, lits, ends).replace(b"\n", b"") for _ in range(90)]
        buf = b"\n".join(lines) + b"\n"
        wf, ws, wa, _ = FormatOracle(cfg, fmt).process_lines(buf)
        emu = EmuFormat(cfg, fmt)
        f, s, masks = emu.process(buf, 30)
        assert f == wf and s == ws, fmt
        assert masks == {i: _mask_of(a, emu.mons) for i, a in enumerate(wa) if a}, fmt
        checked += sum(wf)
    assert checked > 100


SYNTH_TMPL = "pid=<*> uid=<*> auid=<*> ses=<*> msg='op=<*> acct=<*> exe=<*> hostname=<*> addr=<*> terminal=<*> res=<*>' pad=<*>"
SYNTH_CFG = {"detectors": {"NewValueDetector": {"method_type": "new_value_detector", "data_use_training": 2000,
             "global": {"g": {"header_variables": [{"pos": "type"}]}},
             "events": {0: {"pam": {"variables": [{"pos": 5, "name": "acct"}, {"pos": 6, "name": "exe"},
                                                  {"pos": 9, "name": "terminal"}, {"pos": 10, "name": "res"}]}}}}}}


def test_emu_kernel_synthetic_header_and_variable_monitors(emu_kernel):
    """Header-capture and template-variable monitors side by side (different source lanes), on
    the config-2 synthetic records with their injected anomalies."""
    from detectmateservice_b200.synth import AuditSynth
    g = AuditSynth(seed=99)
    train, _ = g.batch(2000, inject=False)
    big, truth = g.batch(24000, inject=True)
    lines = big.split(b"\n")[:-1]
    sel = []
    for i in np.flatnonzero(truth):
        sel += lines[max(0, i - 1):i + 1]
    test = b"\n".join(sel) + b"\n"
    orc = FormatOracle(SYNTH_CFG, AUDIT, [SYNTH_TMPL])
    orc.process_lines(train)
    wf, ws, wa, bad = orc.process_lines(test)
    emu = EmuFormat(SYNTH_CFG, AUDIT, [SYNTH_TMPL])
    emu.process(train, 2000)
    f, s, masks = emu.process(test, 0)
    assert bad == 0 and f == wf and s == ws and sum(wf) >= 15
    assert masks == {i: _mask_of(a, emu.mons) for i, a in enumerate(wa) if a}
    assert len({m for m in masks.values()}) >= 3                  # several different monitors fired


AUDIT_NORM_CFG = {"detectors": {"NewValueDetector": {
    "method_type": "new_value_detector", "data_use_training": 250, "auto_config": False,
    "global": {"g": {"header_variables": [{"pos": "type"}]}},
    "events": {0: {"pam": {"variables": [{"pos": 4, "name": "op"}, {"pos": 6, "name": "exe"}, {"pos": 9, "name": "terminal"}]}},
               1: {"unit": {"variables": [{"pos": 4, "name": "unit"}]}},
               2: {"login": {"variables": [{"pos": 3, "name": "auid"}, {"pos": 7, "name": "res"}]}}}}}}
# records that differ from trained ones only in what R-norm removes (no alert) or in more (alert)
AUDIT_NORM_TAIL = (
    b"type=USER_ACCT msg=audit(1642723741.072:375): pid=10125 uid=0 auid=4294967295 ses=4294967295 "
    b"msg='op=pam:ACCOUNTING acct=\"root\" exe=\"/usr/sbin/CRON\" hostname=? addr=? terminal=cron res=success'\n"
    b"type=USER_ACCT msg=audit(1642723741.072:375): pid=10125   uid=0 AUID=4294967295 ses=4294967295 "
    b"msg='op=PAM:accounting acct=\"root\" exe=\"/usr/sbin/cron2\" hostname=? addr=? terminal=cron res=success'\n"
    b"type=LOGIN msg=audit(1642723741.076:377): pid=1 uid=0 old-auid=4294967295 auid=77777 tty=(none) old-ses=4294967295 ses=65 res=9")


def test_emu_kernel_normalisation_switches(golden_dir, emu_kernel):
    """R-norm on the device path: the reference's audit templates with all three switches on
    (tests/library_integration/test_pipe_filereader_matcher_nvd.py:82-84), then random templates
    under every switch combination, against the regular-expression oracle."""
    tm = audit_templates(golden_dir)
    buf = open(os.path.join(golden_dir, "audit_sample.log"), "rb").read() + AUDIT_NORM_TAIL
    norm = (True, True, True)
    wf, ws, wa, bad = FormatOracle(AUDIT_NORM_CFG, AUDIT, tm, norm).process_lines(buf)
    emu = EmuFormat(AUDIT_NORM_CFG, AUDIT, tm, norm=norm)
    f, s, masks = emu.process(buf, 250)
    assert f == wf and s == ws and bad == 0
    assert masks == {i: _mask_of(a, emu.mons) for i, a in enumerate(wa) if a}
    assert wf[-3:] == [0, 1, 1] and wa[-2] == {"EventID 0 - exe": "Unknown value: 'usrsbincron2'"}
    r = np.random.Generator(np.random.PCG64(29))
    checked = 0
    for it in range(20):
        lits, ends = norm_fuzz_template(r)
        norm = NORMS[it % len(NORMS)]
        n_caps = len(lits) - (0 if ends else 1)
        tmpl = b"<*>".join(lits) + (b"<*>" if ends else b"")
        cfg = {"detectors": {"NewValueDetector": {"method_type": "new_value_detector", "data_use_training": 30,
               "global": {"g": {"header_variables": [{"pos": "h"}]}},
               "events": {0: {"e": {"variables": [{"pos": i} for i in range(n_caps)]}}}}}}
        lines = [b"%d|" % int(r.integers(0, 3)) + norm_fuzz_text(r, lits, ends) for _ in range(120)]
        buf = b"\n".join(lines) + b"\n"
        wf, ws, wa, _ = FormatOracle(cfg, "<h>|<Content>", [tmpl], norm).process_lines(buf)
        emu = EmuFormat(cfg, "<h>|<Content>", [tmpl], norm=norm)
        f, s, masks = emu.process(buf, 30)
        assert f == wf and s == ws, (tmpl, norm)
        assert masks == {i: _mask_of(a, emu.mons) for i, a in enumerate(wa) if a}, (tmpl, norm)
        checked += sum(wf)
    assert checked > 100


def test_emu_set_format_errors():
    with pytest.raises(ValueError, match="nothing between"):
        EmuFormat(NGINX_CFG, "<Content>", ["a<*> - <*>b"], norm=(True, True, False))
    with pytest.raises(ValueError, match="nothing between"):
        EmuFormat(NGINX_CFG, "<A><B>")
    with pytest.raises(ValueError, match="twice"):
        EmuFormat(NGINX_CFG, "<A> <A>")
    with pytest.raises(ValueError, match="Content"):
        EmuFormat(NGINX_CFG, "<A> x", ["y<*>"])
    with pytest.raises(ValueError, match="more than 32"):
        EmuFormat(NGINX_CFG, " ".join("<c%d>" % i for i in range(40)))


# ------------------------------------------------------------------------------------------ GPU tier
def _component(cfg, params, name="NewValueDetector"):
    from detectmateservice_b200.component import B200NewValueDetector
    det = dict(cfg["detectors"]["NewValueDetector"])
    det["params"] = dict(det.get("params") or {}, **params)
    c = B200NewValueDetector(name=name, config={"detectors": {name: det}})
    c.clock = lambda: 1773848383
    return c


@pytest.mark.gpu
def test_gpu_docs_nginx_pipeline_golden(golden_dir):
    """docs/getting_started.md:395-435,498-510 with parser + detector fused: raw access-log lines in,
    the documented DetectorSchema out."""
    g = json.load(open(os.path.join(golden_dir, "docs_golden.json")))
    comp = _component(NGINX_CFG, {"log_format": NGINX})
    outs = [comp.process(nginx_line(u) + b"\n") for u in ("/hello", "/world", "/foobar")]
    assert outs[0] is None and outs[1] is None
    a = wire.decode_detector_schema(outs[2])
    e = g["expected"]
    assert a["alertsObtain"] == e["alertsObtain"] == {"Global - URL": "Unknown value: '/foobar'"}
    assert (a["detectorID"], a["detectorType"], a["alertID"], a["score"], a["description"]) == (
        e["detectorID"], e["detectorType"], e["alertID"], e["score"], e["description"])
    assert a["extractedTimestamps"] == [1773848383]              # the nginx Time is not epoch seconds: detection time
    # a whole message of lines, one of them not in the format
    buf = b"\n".join([nginx_line("/hello"), b"garbage line", nginx_line("/admin", ip="10.0.0.9"), nginx_line("/world")]) + b"\n"
    alerts = [wire.decode_detector_schema(b) for b in wire.split_delimited(comp.process(buf))]
    assert [x["alertsObtain"] for x in alerts] == [{"Global - URL": "Unknown value: '/admin'"}]
    assert alerts[0]["logIDs"] == ["5"] and comp.stats()["bad_records"] == 1
    comp.close()


@pytest.mark.gpu
def test_gpu_audit_templates_component_vs_oracle(golden_dir):
    tm = audit_templates(golden_dir)
    buf = open(os.path.join(golden_dir, "audit_sample.log"), "rb").read()
    buf += b"type=WEIRD msg=audit(1.0:1): totally unlike=any template\nnot an audit line at all\n\n"
    buf += b"type=LOGIN msg=audit(1642723741.076:377): pid=1 uid=0 old-auid=4294967295 auid=77777 tty=(none) old-ses=4294967295 ses=65 res=9\n"
    wf, ws, wa, bad = FormatOracle(AUDIT_CFG, AUDIT, tm).process_lines(buf)
    comp = _component(AUDIT_CFG, {"log_format": AUDIT, "path_templates": os.path.join(golden_dir, "audit_templates.txt")})
    cut = buf.index(b"\n", len(buf) // 3) + 1                    # two messages; training ends inside the first or second
    got = []
    for part in (buf[:cut], buf[cut:]):
        out = comp.process(part)
        got += [wire.decode_detector_schema(b) for b in wire.split_delimited(out)] if out else []
    want = [(i, a) for i, a in enumerate(wa) if a]
    assert [(int(x["logIDs"][0]), x["alertsObtain"]) for x in got] == want
    assert [x["score"] for x in got] == [ws[i] for i, _ in want]
    st = comp.stats()
    assert st["bad_records"] == bad == 2 and st["anomalies"] == sum(wf) and st["lines"] == len(wf)
    comp.close()
    from detectmateservice_b200.component import decode_compact
    comp = _component(AUDIT_CFG, {"log_format": AUDIT, "templates": [t.decode() for t in tm], "output_format": "compact"})
    f, s = decode_compact(comp.process(buf))
    assert f.tolist() == wf and s.tolist() == ws
    comp.close()


@pytest.mark.gpu
def test_gpu_normalisation_switches(golden_dir):
    """The reference's audit parser config as it stands -- remove_spaces, remove_punctuation, lowercase all
    true (tests/library_integration/test_pipe_filereader_matcher_nvd.py:74-88) -- through the component,
    then random templates under every switch combination through the C-ABI, against the oracle."""
    from detectmateservice_b200.detector import DeviceDetector
    from detectmateservice_b200.component import parse_monitors, select_component_config
    tm = audit_templates(golden_dir)
    buf = open(os.path.join(golden_dir, "audit_sample.log"), "rb").read() + AUDIT_NORM_TAIL + b"\n"
    norm = (True, True, True)
    wf, ws, wa, bad = FormatOracle(AUDIT_NORM_CFG, AUDIT, tm, norm).process_lines(buf)
    comp = _component(AUDIT_NORM_CFG, {"log_format": AUDIT, "path_templates": os.path.join(golden_dir, "audit_templates.txt"),
                                       "remove_spaces": True, "remove_punctuation": True, "lowercase": True})
    cut = buf.index(b"\n", len(buf) // 2) + 1
    got = []
    for part in (buf[:cut], buf[cut:]):
        out = comp.process(part)
        got += [wire.decode_detector_schema(b) for b in wire.split_delimited(out)] if out else []
    want = [(i, a) for i, a in enumerate(wa) if a]
    assert [(int(x["logIDs"][0]), x["alertsObtain"]) for x in got] == want and len(want) >= 5
    assert want[-2][1] == {"EventID 0 - exe": "Unknown value: 'usrsbincron2'"}
    assert comp.stats()["bad_records"] == bad == 0
    comp.close()
    r = np.random.Generator(np.random.PCG64(31))
    checked = 0
    for it in range(40):
        lits, ends = norm_fuzz_template(r)
        norm = NORMS[it % len(NORMS)]
        n_caps = len(lits) - (0 if ends else 1)
        tmpl = b"<*>".join(lits) + (b"<*>" if ends else b"")
        cfg = {"detectors": {"NewValueDetector": {"method_type": "new_value_detector", "data_use_training": 100,
               "global": {"g": {"header_variables": [{"pos": "h"}]}},
               "events": {0: {"e": {"variables": [{"pos": i} for i in range(n_caps)]}}}}}}
        mons = parse_monitors(select_component_config(cfg, "NewValueDetector"))
        lines = [b"%d|" % int(r.integers(0, 3)) + norm_fuzz_text(r, lits, ends) for _ in range(500)]
        buf = b"\n".join(lines) + b"\n"
        wf, ws, wa, _ = FormatOracle(cfg, "<h>|<Content>", [tmpl], norm).process_lines(buf)
        det = DeviceDetector([m.key for m in mons], max_batch_bytes=1 << 20)
        det.set_monitors([{"event_id": m.event_id, "source": m.source, "pos": m.pos} for m in mons])
        det.set_format("<h>|<Content>", [tmpl], norm_flags=norm_flags(*norm))
        f, s = det.process_lines(buf, n_train_lines=100)
        assert f.tolist() == wf and s.tolist() == ws, (tmpl, norm)
        assert {a[0]: a[1] for a in det.anomalies()} == {i: _mask_of(a, mons) for i, a in enumerate(wa) if a}
        checked += sum(wf)
        det.close()
    assert checked > 300


@pytest.mark.gpu
def test_gpu_fuzz_formats_and_switching():
    from detectmateservice_b200.detector import DeviceDetector
    from detectmateservice_b200._lib import DmError
    from detectmateservice_b200.component import parse_monitors, select_component_config
    r = np.random.Generator(np.random.PCG64(23))
    checked = 0
    for lits, ends in fuzz_formats(r, 60):
        names = ["c%d" % i for i in range(len(lits) - (0 if ends else 1))]
        if not names:
            continue
        fmt = b"".join(l + (b"<%s>" % names[i].encode() if i < len(names) else b"") for i, l in enumerate(lits))
        cfg = {"detectors": {"NewValueDetector": {"method_type": "new_value_detector", "data_use_training": 100,
                                                  "global": {"g": {"header_variables": [{"pos": n} for n in names]}}}}}
        mons = parse_monitors(select_component_config(cfg, "NewValueDetector"))
        lines = [fuzz_text(r, lits, ends).replace(b"\n", b"") for _ in range(400)]
        buf = b"\n".join(lines) + b"\n"
        wf, ws, wa, _ = FormatOracle(cfg, fmt).process_lines(buf)
        det = DeviceDetector([m.key for m in mons], max_batch_bytes=1 << 20)
        det.set_monitors([{"event_id": m.event_id, "source": m.source, "pos": m.pos} for m in mons])
        det.set_format(fmt.decode())
        f, s = det.process_lines(buf, n_train_lines=100)
        assert f.tolist() == wf and s.tolist() == ws, fmt
        assert {a[0]: a[1] for a in det.anomalies()} == {i: _mask_of(a, mons) for i, a in enumerate(wa) if a}
        checked += sum(wf)
        det.close()
    assert checked > 500
    # errors, and switching between log_format and key=value tokenisation on one handle
    det = DeviceDetector([b"type"], max_batch_bytes=1 << 20)
    with pytest.raises(DmError):
        det.set_format("<A> <B>")                                # before set_monitors
    det.set_monitors([{"event_id": None, "source": "header", "pos": "type"}])
    for bad in ("<A><B>", "<A> <A>"):
        with pytest.raises(DmError):
            det.set_format(bad)
    with pytest.raises(DmError):
        det.set_format("<A> x", ["y<*>"])
    lines = b"type=A msg=audit(1.0:1): k=v\ntype=B msg=audit(1.0:2): k=v\ntype=A msg=audit(1.0:3): k=v\n"
    det.set_format(AUDIT)
    assert det.process_lines(lines, n_train_lines=1)[0].tolist() == [0, 1, 0]
    with pytest.raises(DmError):
        det.submit(lines)                                        # the pipelined path is key=value only
    det.set_format(None)
    det.reset()
    assert det.process_lines(lines, n_train_lines=2)[0].tolist() == [0, 0, 0]
    det.close()


@pytest.mark.gpu
def test_gpu_format_mode_full_size_message():
    """BASELINE config-2 sized message (64k x 256 B) in log_format + template mode against the oracle."""
    from detectmateservice_b200.detector import DeviceDetector
    from detectmateservice_b200.component import parse_monitors, select_component_config
    from detectmateservice_b200.synth import AuditSynth
    tmpl, cfg = SYNTH_TMPL, SYNTH_CFG
    mons = parse_monitors(select_component_config(cfg, "NewValueDetector"))
    g = AuditSynth(seed=99)
    train, _ = g.batch(65536, inject=False)
    test, truth = g.batch(65536, inject=True)
    det = DeviceDetector([m.key for m in mons], max_batch_bytes=32 << 20)
    det.set_monitors([{"event_id": m.event_id, "source": m.source, "pos": m.pos} for m in mons])
    det.set_format(AUDIT, [tmpl])
    det.process_lines(train, n_train_lines=65536)
    f, s = det.process_lines(test)
    orc = FormatOracle(cfg, AUDIT, [tmpl])
    orc.nvd.data_use_training = 65536
    orc.process_lines(train)
    wf, ws, _, bad = orc.process_lines(test)
    assert bad == 0 and f.tolist() == wf and s.tolist() == ws
    assert int(f.sum()) == int(np.asarray(truth).astype(bool).sum()) > 20       # every injected anomaly, nothing else
    det.close()
