"""The C-ABI library loads and exports every symbol include/mmq.h declares.
No compute calls (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mmq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmq_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    from mmseq_b200 import capi
    assert _declared() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "mmseq_b200", "libmmseq_b200.so")
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(lib, name), name
    lib.mmq_version.restype = ctypes.c_char_p
    assert b"mmseq-b200" in lib.mmq_version()


def test_create_without_gpu_fails_loudly():
    """No CPU fallback: on a box without a CUDA device mmq_create must fail with a message."""
    import numpy as np
    from mmseq_b200 import capi
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(capi.MmqError) as e:
        capi.Handle([0, 1], [0], [3], np.array([1e-3]))
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_create_rejects_bad_problems_before_touching_the_gpu():
    import numpy as np
    from mmseq_b200 import capi
    with pytest.raises(capi.MmqError):
        capi.Handle([1, 2], [0], [3], np.array([1e-3]))          # row_ptr[0] != 0
    with pytest.raises(capi.MmqError):
        capi.Handle([0, 1], [0], [3], np.array([1e-3]), alpha=0.0)
