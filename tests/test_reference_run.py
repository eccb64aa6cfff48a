"""Parity pinned on the REFERENCE's own code: eturro/mmseq's src/mmseq.cpp, hitsio.cpp, uh.cpp and
sokal.cc compiled UNMODIFIED against oracle/shim (minimal Boost/GSL stand-ins) into
oracle/_ref/mmseq_ref, and its outputs for a small hits file committed under
tests/golden/ref_small/ (tools/make_golden_ref_small.py).

What this pins, without a GPU:
  * the oracle's restatement of main() (oracle/tables.py) and the product's loader reproduce the
    reference's class construction, numbering and ordering: .k and .M byte for byte;
  * every column that does not depend on the random stream — feature ids and their order,
    effective_length, true_length, unique_hits (uh()), log_mu_em (the EM), observed, ntranscripts,
    the closed forms of hit-less features, NA / -nan / inf placement — cell for cell;
  * the stochastic columns of the oracle's Philox chain agree with the reference's MT19937/GSL-style
    chain within Monte-Carlo standard error (north_star c).
The GPU-side counterpart (the `mmseq` host program against the same reference outputs) is
tests/test_gpu_cli.py::test_cli_matches_reference_run."""
import os
import subprocess

import numpy as np
import pytest

from oracle import tables
from tests.ref_case import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_small")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "mmseq_ref")

# columns that do not depend on the random stream, by table
DET = {
    "mmseq": ["feature_id", "effective_length", "true_length", "unique_hits", "log_mu_em", "observed", "ntranscripts"],
    "identical": ["feature_id", "effective_length", "true_length", "unique_hits", "observed", "ntranscripts"],
    "gene": ["feature_id", "true_length", "unique_hits", "ntranscripts", "observed"],
}


def compare_with_reference(ref_rows, our_rows, kind, z_max=5.0, frac=0.97):
    """Deterministic columns equal; stochastic log_mu within z_max combined MCSE for >= frac of the
    observed features; hit-less features (closed forms) equal in every non-simulated column."""
    assert ref_rows[0] == our_rows[0] and ref_rows[1] == our_rows[1]      # '# Mapped fragments' + header
    hdr = ref_rows[1]
    assert len(ref_rows) == len(our_rows)
    col = {name: i for i, name in enumerate(hdr)}
    z = []
    for a, b in zip(ref_rows[2:], our_rows[2:]):
        for name in DET[kind]:
            assert tables.cells_match(a[col[name]], b[col[name]], rtol=2e-5), (kind, a[0], name, a[col[name]], b[col[name]])
        observed = a[col["observed"]] == "1"
        if not observed and kind != "gene":        # closed forms: log_mu, sd, mcse, iact are literals of alpha, beta, length
            for name in ("log_mu", "sd", "mcse", "iact"):
                assert tables.cells_match(a[col[name]], b[col[name]], rtol=2e-5), (kind, a[0], name)
        if observed:
            la, lb = float(a[col["log_mu"]]), float(b[col["log_mu"]])
            ma, mb = float(a[col["mcse"]]), float(b[col["mcse"]])
            if np.isfinite(la) and np.isfinite(lb):
                z.append(abs(la - lb) / max(np.hypot(ma, mb), 1e-12))
    z = np.array(z)
    assert len(z) > 0 and np.mean(z <= z_max) >= frac, np.sort(z)[-8:]
    return z


def _golden(kind):
    name = {"mmseq": "ref.mmseq", "identical": "ref.identical.mmseq", "gene": "ref.gene.mmseq"}[kind]
    return tables.read_table(os.path.join(GOLD, name))


@pytest.fixture(scope="module")
def oracle_run(tmp_path_factory):
    path = make_case(tmp_path_factory.mktemp("refcase"), "text")
    return path, tables.run(path, gibbs_iter=4096, seed=99, percentiles=(5.0, 50.0, 95.0))


def test_oracle_matches_reference_golden(oracle_run):
    path, want = oracle_run
    assert open(os.path.join(GOLD, "ref.k")).read().split("\n")[:-1] == want["k"]
    assert open(os.path.join(GOLD, "ref.M")).read().split("\n")[:-1] == want["M"]
    for kind in ("mmseq", "identical", "gene"):
        compare_with_reference(_golden(kind), want[kind], kind)
    # the reference prints -nan for sd_probit_proportion of single-isoform genes; so does the oracle
    ref = _golden("mmseq")
    col = {n: i for i, n in enumerate(ref[1])}
    for a, b in zip(ref[2:], want["mmseq"][2:]):
        if a[col["ntranscripts"]] == "1":
            assert a[col["sd_probit_proportion"]] == b[col["sd_probit_proportion"]] == "-nan"
            assert a[col["mean_probit_proportion"]] == b[col["mean_probit_proportion"]] == "inf"


def test_product_loader_matches_reference_k_and_M(oracle_run):
    """libmmq_host.so builds the classes the reference builds (same numbering and order)."""
    from mmseq_b200 import hostlib
    path, _ = oracle_run
    h = hostlib.load_hits(path)
    k_ref = [int(x) for x in open(os.path.join(GOLD, "ref.k")).read().split()]
    assert list(h.k) == k_ref
    m_lines = open(os.path.join(GOLD, "ref.M")).read().split("\n")[:-1]
    assert m_lines[0] == "#" + "".join("\t" + h.names[i] for i in h.col2hdr)
    pairs = [tuple(map(int, ln.split("\t"))) for ln in m_lines[1:]]
    ours = [(i, int(c)) for i in range(h.m) for c in h.col[h.row_ptr[i]:h.row_ptr[i + 1]]]
    assert pairs == ours


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/mmseq_ref not built (reference tree absent)")
def test_reference_binary_live_text_and_binary_inputs(tmp_path, oracle_run):
    """Run the reference itself on both hits schemas: same classes, same deterministic columns."""
    _, want = oracle_run
    for fmt in ("text", "binary"):
        path = make_case(tmp_path, fmt)
        base = str(tmp_path / f"ref_{fmt}")
        r = subprocess.run([REF_BIN, "-gibbs_iter", "2048", "-seed", "5", "-percentiles", "5,50,95", path, base],
                           capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="2"), timeout=600)
        assert r.returncode == 0, r.stderr
        assert open(base + ".k").read().split("\n")[:-1] == want["k"]
        assert open(base + ".M").read().split("\n")[:-1] == want["M"]
        for kind, ext in (("mmseq", ".mmseq"), ("identical", ".identical.mmseq"), ("gene", ".gene.mmseq")):
            compare_with_reference(tables.read_table(base + ext), want[kind], kind, z_max=6.0, frac=0.95)
