import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dex-tts_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "run_last: GPU tests committed before they could be run on a B200 as pytest items; they are "
                                       "ordered after every validated test so that -x cannot hide the validated suite behind them")


def pytest_collection_modifyitems(config, items):
    def last_key(it):                        # (0, 0) for validated items; run_last(n) items after them, in increasing n
        m = it.get_closest_marker("run_last")
        return (0, 0) if m is None else (1, m.args[0] if m.args else 0)
    items.sort(key=last_key)                 # stable: everything else keeps its order
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
