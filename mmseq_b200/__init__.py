"""mmseq_b200 — B200 (sm_100a) implementation of mmseq's EM + Gibbs hot path.

The product is native: ``libmmseq_b200.so`` (CUDA kernels behind the C ABI of
``include/mmq.h``), ``libmmq_host.so`` (the .hits loader / hit-class builder) and
the ``mmseq`` host program.  This Python package is only the ctypes view of
those libraries used by the tests and ``bench.py``; it contains no compute and
no CPU fallback — importing :mod:`mmseq_b200.capi` fails loudly if the CUDA
library has not been built (``python -c "import __graft_entry__ as g; g.build()"``).
"""
from . import capi, hostlib, synth  # noqa: F401

__all__ = ["capi", "hostlib", "synth"]
