import os
import sys
from pathlib import Path

import pytest

_SWITCHES = ("SKR_FORCE_INTERP", "SKR_NO_PINNED", "SKR_IN_MODE", "SKR_STAGES", "SKR_CTAS")
ROOT = Path(__file__).resolve().parent.parent
for extra in (ROOT, ROOT / "tests"):
    if str(extra) not in sys.path:
        sys.path.insert(0, str(extra))


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config: pytest.Config, items: list[pytest.Item]) -> None:
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _switch_hygiene():
    """Tests that flip the library's development switches (SKR_FORCE_INTERP, SKR_NO_PINNED) call
    ``native.reset_switches()`` after setting them; this re-reads the restored environment afterwards (monkeypatch,
    requested by the test, is undone before this autouse fixture finishes) and drops plans made under the switches."""
    before = {k: os.environ.get(k) for k in _SWITCHES}
    yield
    import sys

    native = sys.modules.get("skrample_b200.native")
    if native is not None and native._lib is not None:
        now = {k: os.environ.get(k) for k in _SWITCHES}
        if now != before or getattr(native, "_switches_touched", False):
            native._switches_touched = False
            native.reset_switches()
