import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def state_dict():
    from nomad_b200.weights import random_state_dict
    return random_state_dict(1234)


_ACHIEVED = []


@pytest.fixture(scope="session")
def record():
    """``record(test, **achieved)``: the GPU parity tests log the errors they ACHIEVED (not just that a threshold held);
    the session writes them to ``$NOMAD_B200_PARITY_LOG`` (default ``gpurun_out/parity_achieved.jsonl``)."""
    def rec(test, **kw):
        row = {"test": test}
        row.update({k: (float(v) if hasattr(v, "__float__") else v) for k, v in kw.items()})
        _ACHIEVED.append(row)
        print("ACHIEVED", row)
    return rec


def pytest_sessionfinish(session, exitstatus):
    if not _ACHIEVED:
        return
    import json
    path = os.environ.get("NOMAD_B200_PARITY_LOG", os.path.join(ROOT, "gpurun_out", "parity_achieved.jsonl"))
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "a") as f:
            for row in _ACHIEVED:
                f.write(json.dumps(row) + "\n")
    except OSError:
        pass
