"""Known-answer vectors of the allocation contract (include/mmq_sampler.h: mmq_alloc_row) and of a
few Gamma draws, written to tests/golden/alloc_kat.npz.  They pin the random-stream contract —
which Philox stream, block and word every draw of a class uses, the order of the running sums —
so that a later change to the sampler that keeps CPU replay and kernels consistent with each other
but silently moves the chain is caught (tests/test_oracle_samplers.py::test_alloc_known_answers).

    python tools/make_golden_alloc.py        # regenerate after a DELIBERATE contract change
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

CASES = [  # (seed, p, k): class ids 0..7, sweep 0 (orc_draw_alloc)
    (1234, [3.0, 0.0, 1.0, 1e-30, 6.0], 1),
    (1234, [3.0, 0.0, 1.0, 1e-30, 6.0], 2),
    (1234, [0.5, 2.5, 1.0, 4.0], 7),
    (99, [0.5, 2.5, 1.0, 4.0], 64),
    (99, [0.5, 2.5, 1.0, 4.0], 65),
    (7, [1.0, 1.0], 1000),
    (7, [2.0, 1.0, 1.0, 5.0, 0.25, 0.25, 3.0, 1.5, 0.5], 8192),
    (7, [2.0, 1.0, 1.0, 5.0, 0.25, 0.25, 3.0, 1.5, 0.5], 8193),
    (5, [1e-3, 1.0, 1e3], 123456),
]


def main():
    out = {}
    for i, (seed, p, k) in enumerate(CASES):
        out[f"seed_{i}"] = np.int64(seed)
        out[f"p_{i}"] = np.array(p)
        out[f"k_{i}"] = np.int64(k)
        out[f"x_{i}"] = orc.draw_alloc(seed, 8, np.array(p), k)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "alloc_kat.npz"), n_cases=np.int64(len(CASES)), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
