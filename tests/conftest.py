import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """CPU-side artefacts (oracle, host loader, generator).  The CUDA library is built by
    __graft_entry__.build(); GPU tests fail loudly if it is missing."""
    need = [os.path.join(ROOT, "oracle", "liboracle.so"), os.path.join(ROOT, "mmseq_b200", "libmmq_host.so"),
            os.path.join(ROOT, "mmseq_b200", "libmmq_synth.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-C", ROOT, "synth", "hostlib", "oracle"], stdout=subprocess.DEVNULL)


_ensure_built()


@pytest.fixture(scope="session")
def small_synth():
    from mmseq_b200 import synth
    return synth.Synth(20260101 + 1, 300, 20000)


@pytest.fixture(scope="session")
def small_problem(small_synth):
    """Collapsed hit classes of a 300-transcript / 20k-fragment synthetic sample."""
    from mmseq_b200 import hostlib
    s = small_synth
    return hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, layout=hostlib.LAYOUT_COLLAPSED)


@pytest.fixture(scope="session")
def small_problem_pf(small_synth):
    from mmseq_b200 import hostlib
    s = small_synth
    return hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, layout=hostlib.LAYOUT_PER_FRAGMENT)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
