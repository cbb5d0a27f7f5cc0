"""CoreComponent / CoreConfig as pinned by the reference's own tests
(/root/reference/tests/test_component_loader/test_detectmatelibrary_import.py:13-27:
``CoreConfig(start_id=100)``, ``CoreComponent(name=..., config=...)``, ``.name``,
``.config.start_id``) and by docs/interfaces.md:9-57."""
from __future__ import annotations

from typing import Any, Optional

from pydantic import BaseModel, ConfigDict


class CoreConfig(BaseModel):
    model_config = ConfigDict(extra="allow")

    start_id: int = 10
    method_type: str = "core"
    comp_type: str = "core"
    auto_config: bool = False

    def to_dict(self) -> dict:
        """User-specified values only (the reference strips defaults before persisting,
        tests/test_reconfigure_params.py:142-146)."""
        return self.model_dump(exclude_defaults=True)


class CoreComponent:
    def __init__(self, name: str = "CoreComponent", config: Optional[Any] = None, **_: Any) -> None:
        self.name = name
        self.config = config if config is not None else CoreConfig()

    def process(self, data: bytes) -> Optional[bytes]:
        return data

    def __repr__(self) -> str:
        return f"<{type(self).__name__} {self.name}>"
