"""R-spec -- CPU restatement of NewValueDetector (and DummyDetector).  TEST INFRASTRUCTURE.

PARITY UNPINNED for NewValueDetector in general: the arithmetic lives in
detectmatelibrary 0.1.0 @ ecdda558 (/root/reference/uv.lock:240-251;
``detectmatelibrary.detectors.new_value_detector.NewValueDetector`` selected at
/root/reference/container/config/detector_settings.yaml:2 and
/root/reference/tests/config/service_settings.yaml:1-2), which is not in /root/reference
and not installable here.  What IS pinned: the documented example
(docs/getting_started.md:423-435,498-510) and the DummyDetector pattern
(tests/library_integration/test_detector_integration.py:83-84,89-115) -- see
tests/test_oracle_golden.py.

R-spec.  A record is a ParserSchema-like mapping
{EventID, variables[], logFormatVariables{}, logID}; raw lines are first turned into
one by oracle/rtok.py (every key=value field becomes a header variable, EventID unset).

 1. Records 1..data_use_training (arrival order) are training: for each configured
    monitor present in the record, insert the value into that monitor's known set;
    process() returns None.  (docs/getting_started.md:427,435)
 2. Later records: for each configured monitor present in the record, value not in
    known set => alert.  Detection never inserts.
 3. Monitors (container/config/detector_config.yaml:1-9, tests/config/detector_config.yaml:1-17,
    tests/test_reconfigure_params.py:35-52,185-204):
      global.<instance>.header_variables[pos=KEY]        every record, key "Global - KEY"
      events.<EventID>.<instance>.header_variables[pos=KEY]  records with that EventID,
                                                         key "EventID <id> - KEY"
      events.<EventID>.<instance>.variables[pos=i,name=N]    value variables[i],
                                                         key "EventID <id> - <N or i>"
    Missing field => monitor skipped for that record.
 4. score = float32(number of alerts); anomaly iff score > 0.
 5. Output only when anomalous: DetectorSchema{__version__ "1.0.0", detectorID name,
    detectorType method_type, alertID str(start_id + alerts so far) with start_id 10,
    detectionTimestamp = receivedTimestamp = int(time()), logIDs [logID],
    extractedTimestamps [int(float(Time))] (else the detection timestamp), score,
    description "<name> detects values not encountered in training as anomalies.",
    alertsObtain {key: "Unknown value: '<value>'"}}   (docs/getting_started.md:510)
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

from . import rtok
from .fingerprint import table_key
from .schemas import DetectorSchema, ParserSchema

GLOBAL = None


@dataclass
class Monitor:
    event_id: Optional[int]          # None = global scope
    source: str                      # "header" | "variable"
    pos: Any                         # str key (header) or int index (variable)
    label: str                       # alert-key suffix

    @property
    def alert_key(self) -> str:
        if self.event_id is None:
            return f"Global - {self.label}"
        return f"EventID {self.event_id} - {self.label}"


def select_component_config(config: Optional[dict], name: str) -> dict:
    """The loader hands the component the whole ServiceConfig dump
    {"detectors": {...}, "parsers": None, "readers": None}
    (src/service/core.py:127-133,144-148); the library picks the entry named after the
    component (docs/interfaces.md:162-167).  ``params`` are flattened into the top level,
    an ``all_`` prefix is stripped (docs/interfaces.md:74-84)."""
    cfg = dict(config or {})
    if "detectors" in cfg and isinstance(cfg["detectors"], dict):
        dets = cfg["detectors"]
        cfg = dict(dets.get(name) or (next(iter(dets.values())) if dets else {}))
    params = cfg.pop("params", None) or {}
    for k, v in params.items():
        cfg[k[4:] if k.startswith("all_") else k] = v
    return cfg


def parse_monitors(cfg: dict) -> List[Monitor]:
    mons: List[Monitor] = []

    def instances(scope: dict, event_id: Optional[int]) -> None:
        for _inst_name, inst in (scope or {}).items():
            inst = inst or {}
            for hv in inst.get("header_variables") or []:
                mons.append(Monitor(event_id, "header", str(hv["pos"]), str(hv["pos"])))
            for var in inst.get("variables") or []:
                pos = int(var["pos"])
                mons.append(Monitor(event_id, "variable", pos, str(var.get("name", pos))))

    instances(cfg.get("global") or {}, None)
    for eid, scope in (cfg.get("events") or {}).items():
        instances(scope or {}, int(eid))
    return mons


def _b(x: Any) -> bytes:
    return x if isinstance(x, bytes) else str(x).encode("utf-8")


class NewValueDetectorOracle:
    def __init__(self, name: str = "NewValueDetector", config: Optional[dict] = None,
                 clock=time.time) -> None:
        self.name = name
        cfg = select_component_config(config, name)
        self.method_type = cfg.get("method_type", "new_value_detector")
        self.data_use_training = int(cfg.get("data_use_training") or 0)
        self.start_id = int(cfg.get("start_id", 10))
        self.monitors = parse_monitors(cfg)
        self.known: List[set] = [set() for _ in self.monitors]
        self.n_seen = 0
        self.n_alerts = 0
        self.clock = clock

    # -- record level -----------------------------------------------------------------
    def _values(self, rec: dict) -> List[Tuple[int, bytes]]:
        out = []
        eid = rec.get("EventID")
        lfv = rec.get("logFormatVariables") or {}
        var = rec.get("variables") or []
        for i, m in enumerate(self.monitors):
            if m.event_id is not None and m.event_id != eid:
                continue
            if m.source == "header":
                v = lfv.get(m.pos)
            else:
                v = var[m.pos] if 0 <= m.pos < len(var) else None
            if v is not None:
                out.append((i, _b(v)))
        return out

    def step(self, rec: dict) -> Tuple[bool, float, Dict[str, str]]:
        """One record through train-or-detect.  Returns (flag, score, alerts)."""
        self.n_seen += 1
        vals = self._values(rec)
        if self.n_seen <= self.data_use_training:
            for i, v in vals:
                self.known[i].add(v)
            return False, 0.0, {}
        alerts: Dict[str, str] = {}
        score = 0.0
        for i, v in vals:
            if v not in self.known[i]:
                alerts[self.monitors[i].alert_key] = "Unknown value: '%s'" % v.decode("utf-8", "replace")
                score += 1.0
        return score > 0, score, alerts

    def make_output(self, rec: dict, score: float, alerts: Dict[str, str]):
        now = int(self.clock())
        out = DetectorSchema()
        out.__setattr__("__version__", "1.0.0")
        out.detectorID = self.name
        out.detectorType = self.method_type
        out.alertID = str(self.start_id + self.n_alerts)
        self.n_alerts += 1
        out.detectionTimestamp = now
        out.receivedTimestamp = now
        out.logIDs.append(str(rec.get("logID", "")))
        t = (rec.get("logFormatVariables") or {}).get("Time")
        try:
            out.extractedTimestamps.append(int(float(t)))
        except (TypeError, ValueError):
            out.extractedTimestamps.append(now)
        out.score = score
        out.description = f"{self.name} detects values not encountered in training as anomalies."
        for k, v in alerts.items():
            out.alertsObtain[k] = v
        return out

    # -- the CoreComponent.process contract (docs/interfaces.md:23-35) -----------------
    def process(self, data: bytes) -> Optional[bytes]:
        msg = ParserSchema()
        msg.ParseFromString(data)
        rec = {
            "EventID": msg.EventID if msg.HasField("EventID") else None,
            "variables": list(msg.variables),
            "logFormatVariables": dict(msg.logFormatVariables),
            "logID": msg.logID,
        }
        flag, score, alerts = self.step(rec)
        if not flag:
            return None
        return self.make_output(rec, score, alerts).SerializeToString()

    # -- raw mode ----------------------------------------------------------------------
    def step_line(self, line: bytes) -> Tuple[bool, float, Dict[str, str]]:
        fields = rtok.tokenize_line(line)
        rec = {"EventID": None, "variables": [],
               "logFormatVariables": {k.decode("latin-1"): v for k, v in fields.items()}}
        return self.step(rec)

    def process_lines(self, buf: bytes) -> Tuple[List[int], List[float]]:
        flags, scores = [], []
        for line in rtok.split_records(buf):
            f, s, _ = self.step_line(line)
            flags.append(int(f))
            scores.append(s)
        return flags, scores

    # -- fingerprint audit (what makes 64-bit fingerprints safe for bit-exact flags) ----
    def known_keys(self) -> List[int]:
        keys: Dict[int, Tuple[int, bytes]] = {}
        for i, s in enumerate(self.known):
            for v in s:
                k = table_key(i, v)
                if k in keys and keys[k] != (i, v):
                    raise AssertionError(f"fingerprint collision: {keys[k]} vs {(i, v)}")
                keys[k] = (i, v)
        return sorted(keys)


class DummyDetectorOracle:
    """detectmatelibrary.detectors.dummy_detector.DummyDetector as pinned by
    /root/reference/tests/library_integration/test_detector_integration.py:83-84,89-115,
    143-144: alternates False, True, False, ...; a detection has score 1.0,
    description "Dummy detection process" and an alert whose value contains
    "Anomaly detected by DummyDetector"."""

    def __init__(self, name: str = "DummyDetector", config: Optional[dict] = None, clock=time.time):
        self.name = name
        self.n = 0
        self.n_alerts = 0
        self.clock = clock

    def process(self, data: bytes) -> Optional[bytes]:
        msg = ParserSchema()
        msg.ParseFromString(data)
        self.n += 1
        if self.n % 2 == 1:
            return None
        now = int(self.clock())
        out = DetectorSchema()
        out.__setattr__("__version__", "1.0.0")
        out.detectorID = self.name
        out.detectorType = "dummy_detector"
        out.alertID = str(10 + self.n_alerts)
        self.n_alerts += 1
        out.detectionTimestamp = now
        out.receivedTimestamp = now
        out.logIDs.append(msg.logID)
        out.score = 1.0
        out.description = "Dummy detection process"
        out.alertsObtain["type"] = "Anomaly detected by DummyDetector"
        return out.SerializeToString()
