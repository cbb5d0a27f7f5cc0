"""R-combo -- CPU restatement of NewValueComboDetector.  TEST INFRASTRUCTURE.

PARITY UNPINNED: the class lives in detectmatelibrary 0.1.0 @ ecdda558
(/root/reference/uv.lock:240-251), which is not in /root/reference; the reference only names
it (src/service/features/component_resolver.py:15,39-40, component_loader.py:22) and shows
its config shape -- the same ``events.<EventID>.<instance>.{variables, header_variables}``
tree as NewValueDetector with ``method_type: new_value_combo_detector`` and several
instances per event (tests/test_reconfigure_params.py:149-169).  No test or document pins
an output, so the semantics below are DEFINED BY US ("tuple-of-fields
hashing, reuses the same table"):

 1. Every instance (``global.<instance>`` or ``events.<EventID>.<instance>``) is ONE
    combination: the ordered tuple of its fields, header_variables first, then variables,
    each in config order (the order oracle/nvd.py::parse_monitors lists them).
 2. A record yields the combination iff the scope applies (global: always; event:
    record.EventID == EventID) and ALL member fields are present; otherwise the
    combination is skipped for that record.
 3. Records 1..data_use_training train (insert the value tuple), later records alert when
    the tuple is not in the known set; detection never inserts.
 4. score = float32(number of unknown combinations), anomaly iff score > 0.
 5. Alert key "<scope> - (<label>, <label>, ...)" with <scope> and <label> as in R-spec 3;
    text "Unknown value combination: ('<v1>', '<v2>', ...)".
 6. Output DetectorSchema as R-spec 5 with detectorType = method_type and description
    "<name> detects value combinations not encountered in training as anomalies."
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Tuple

from .nvd import Monitor, NewValueDetectorOracle, _b, select_component_config


def parse_combos(cfg: dict) -> Tuple[List[Monitor], List[List[int]]]:
    """(member monitors in parse_monitors order, combos as lists of member indices)."""
    mons: List[Monitor] = []
    combos: List[List[int]] = []

    def instances(scope: dict, event_id: Optional[int]) -> None:
        for _inst_name, inst in (scope or {}).items():
            inst = inst or {}
            members = []
            for hv in inst.get("header_variables") or []:
                members.append(len(mons))
                mons.append(Monitor(event_id, "header", str(hv["pos"]), str(hv["pos"])))
            for var in inst.get("variables") or []:
                pos = int(var["pos"])
                members.append(len(mons))
                mons.append(Monitor(event_id, "variable", pos, str(var.get("name", pos))))
            if members:
                combos.append(members)

    instances(cfg.get("global") or {}, None)
    for eid, scope in (cfg.get("events") or {}).items():
        instances(scope or {}, int(eid))
    return mons, combos


def combo_alert_key(mons: List[Monitor], members: List[int]) -> str:
    m0 = mons[members[0]]
    scope = "Global" if m0.event_id is None else f"EventID {m0.event_id}"
    return "%s - (%s)" % (scope, ", ".join(mons[i].label for i in members))


def combo_alert_text(values: List[bytes]) -> str:
    return "Unknown value combination: (%s)" % ", ".join("'%s'" % v.decode("utf-8", "replace") for v in values)


class NewValueComboDetectorOracle(NewValueDetectorOracle):
    def __init__(self, name: str = "NewValueComboDetector", config: Optional[dict] = None, clock=time.time) -> None:
        super().__init__(name=name, config=config, clock=clock)
        cfg = select_component_config(config, name)
        self.method_type = cfg.get("method_type", "new_value_combo_detector")
        self.monitors, self.combos = parse_combos(cfg)
        self.known_combos: List[set] = [set() for _ in self.combos]

    def step(self, rec: dict) -> Tuple[bool, float, Dict[str, str]]:
        self.n_seen += 1
        vals = dict(self._values(rec))                      # member index -> value (scope already applied)
        alerts: Dict[str, str] = {}
        score = 0.0
        for c, members in enumerate(self.combos):
            if not all(i in vals for i in members):
                continue
            tup = tuple(vals[i] for i in members)
            if self.n_seen <= self.data_use_training:
                self.known_combos[c].add(tup)
            elif tup not in self.known_combos[c]:
                alerts[combo_alert_key(self.monitors, members)] = combo_alert_text(list(tup))
                score += 1.0
        return score > 0, score, alerts

    def make_output(self, rec: dict, score: float, alerts: Dict[str, str]):
        out = super().make_output(rec, score, alerts)
        out.description = f"{self.name} detects value combinations not encountered in training as anomalies."
        return out
