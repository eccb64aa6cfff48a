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


def test_fast_g_format_equals_printf():
    """fmt_g6.h (the trace files' number format) gives the characters of printf's %g == operator<< at precision 6."""
    import numpy as np
    from mmseq_b200 import hostlib
    rng = np.random.default_rng(3)
    edge = [0.0, -0.0, 1.0, -1.5, 1e5, 999999.0, 999999.5, 1e6, 123456.5, 1e-4, 1e-5, 0.000099999949, 9.999995e-5, 1e22, 1e23, 1e27,
            float("inf"), float("-inf"), 5e-324, 2.5, 0.5, 1234565.0, 1234575.0, 1e-17, 9.99999e-18, 1.0000005, 0.1, 1 / 3, 2 / 3, 1e-300, 1e300]
    pw = np.array([m * 10.0 ** k for k in range(-40, 40) for m in (1.0, 9.999995, 9.9999949, 1.0000005, 1.234565, 0.9999995)])
    pw = np.concatenate([pw, np.nextafter(pw, 0), np.nextafter(pw, np.inf)])
    vals = np.concatenate([np.array(edge), pw, np.exp(rng.uniform(-45, 45, 400000)), rng.integers(0, 10 ** 8, 200000) / rng.integers(1, 10 ** 6, 200000),
                           rng.gamma(0.1, 1.0, 200000), rng.standard_normal(50000)])
    got = hostlib.fmt_g6(vals)
    want = ["%g" % v for v in vals]
    assert got == want
    assert hostlib.fmt_g6([float("nan")])[0].lstrip("-") == "nan"


def test_huffman_only_gzip_members_decode_with_zlib():
    """huff_gz.h (the trace files' compressor): any inflate must read its members; concatenated members read as one stream."""
    import gzip
    import io
    import zlib
    import numpy as np
    from mmseq_b200 import hostlib
    rng = np.random.default_rng(9)
    texts = [b"", b"a", b"aaaaaaaaaaaaaaaa", bytes(range(256)) * 3,
             " ".join("%g" % v for v in np.exp(rng.normal(0, 3, 200000))).encode() + b"\n",
             rng.integers(0, 256, 100000, dtype=np.uint8).tobytes(),
             # a very skewed histogram: code lengths beyond 15 bits before the limit is enforced
             b"".join(bytes([i]) * (1 << min(i, 22)) for i in range(26))]
    whole = b""
    for t in texts:
        z = hostlib.gz_huffman(t)
        assert gzip.decompress(z) == t
        assert zlib.decompress(z, 15 + 16) == t
        whole += z
    assert gzip.GzipFile(fileobj=io.BytesIO(whole)).read() == b"".join(texts)
    digits = texts[4]
    assert len(hostlib.gz_huffman(digits)) < 1.02 * len(zlib.compress(digits, 1)) / 1.05   # at least zlib level 1's ratio on digit text


def test_trace_file_writer_round_trip(tmp_path, monkeypatch):
    """trace_writer.h: ids line, one line per slot, "%g" of every kept feature (src/mmseq.cpp:1033-1108), through the built-in
    Huffman-only members and through zlib (MMQ_GZIP_LEVEL): the decompressed text is the same."""
    import gzip
    import numpy as np
    from mmseq_b200 import hostlib
    rng = np.random.default_rng(4)
    n, L = 3000, 64
    tr = np.exp(rng.normal(0, 3, (n, L)))
    tr[5, :] = 0.0
    tr[6, 3] = np.inf
    ids = [f"T{i}" for i in range(n)]
    keep = (rng.random(n) < 0.9).astype(np.uint8)
    texts = []
    monkeypatch.setenv("MMQ_TRACE_BLOCK_BYTES", str(2 * 12 * n))       # two lines per thread and round: several rounds, the writer thread overlaps
    for level in (None, "6"):
        if level:
            monkeypatch.setenv("MMQ_GZIP_LEVEL", level)
        for kp in (None, keep):
            path = str(tmp_path / f"t{level}{kp is None}.gz")
            hostlib.write_trace_gz(path, ids, tr, kp)
            lines = gzip.open(path, "rt").read().split("\n")
            sel = np.arange(n) if kp is None else np.flatnonzero(kp)
            assert lines[0].split() == [ids[i] for i in sel] and lines[0].endswith(" ")
            assert len(lines) == L + 2 and lines[-1] == ""
            for i in (0, 1, L // 2, L - 1):
                assert lines[1 + i].split() == ["%g" % v for v in tr[sel, i]]
            texts.append((kp is None, "\n".join(lines)))
    assert texts[0][1] == texts[2][1] and texts[1][1] == texts[3][1]
