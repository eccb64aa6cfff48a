"""Command-line behaviour of the `mmseq` host program that needs no GPU: usage, version,
validation messages and exit codes of src/mmseq.cpp:207-296."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "mmseq_b200", "bin", "mmseq")


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(BIN):
        env = dict(os.environ); env.pop("CXX", None); env.pop("CC", None)
        subprocess.check_call(["make", "-C", ROOT, "lib", "cli"], env=env, stdout=subprocess.DEVNULL)


def run(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True, timeout=60)


def test_help_and_version_exit_1_on_stderr():
    for flag in ("-h", "-help", "--help"):
        r = run(flag)
        assert r.returncode == 1 and "Usage: mmseq [OPTIONS...] hits_file output_base" in r.stderr
        assert "-gibbs_iter INT    number of Gibbs iterations (default: 16384)" in r.stderr
    for flag in ("-v", "-version", "--version"):
        r = run(flag)
        assert r.returncode == 1 and r.stderr.startswith("mmseq-")


def test_argument_errors():
    r = run()
    assert r.returncode == 1 and "Error: mandatory arguments missing." in r.stderr
    r = run("-bogus", "a", "b")
    assert r.returncode == 1 and "Error: unrecognised option -bogus." in r.stderr
    r = run("-gibbs_iter", "1000", "-gibbs_ss", "7", "a", "b")
    assert r.returncode == 1 and "Error: gibbs_iter must be divisible by gibbs_ss." in r.stderr
    r = run("-percentiles", "5,101", "a", "b")
    assert r.returncode == 1 and "Percentiles must be in (0,100)" in r.stderr
    r = run("-gibbs_iter", "-16", "a", "b")
    assert r.returncode == 1 and "no. of iteratons or trace length <= 0" in r.stderr


def test_missing_hits_file_message(tmp_path):
    r = run(str(tmp_path / "nope.hits"), str(tmp_path / "out"))
    assert r.returncode == 1 and 'Error reading hits file "' in r.stderr


def test_bad_header_message(tmp_path):
    p = tmp_path / "bad.hits"
    p.write_text("@TranscriptMetaData\tA\t10\t20\n@TranscriptMetaData\tB\t10\t20\n@GeneIsoforms\tg\tA\n>r\nA\n")
    r = run(str(p), str(tmp_path / "out"))
    assert r.returncode == 1 and "does not belong to a gene in the @GeneIsoforms header entries." in r.stderr
