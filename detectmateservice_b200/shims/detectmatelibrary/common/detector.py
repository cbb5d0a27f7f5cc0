"""`CoreDetectorConfig` as the reference's tests use it
(/root/reference/tests/test_reconfigure_params.py:10,16,75-79,142-146): the schema class a
detector service hands to its ConfigManager.  It carries the library's detector defaults
(`parser`, `start_id`, `comp_type`), which `to_dict()` strips again so that a persisted YAML
file holds user-specified values only."""
from __future__ import annotations

from typing import Any, Dict, Optional

from .core import CoreConfig


class CoreDetectorConfig(CoreConfig):
    comp_type: str = "detectors"
    method_type: str = "core_detector"
    parser: str = "<PLACEHOLDER>"
    data_use_training: Optional[int] = None
    events: Optional[Dict[Any, Any]] = None
    global_: Optional[Dict[str, Any]] = None
