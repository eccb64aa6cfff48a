"""The product's .hits loader / class builder (libmmq_host.so) against the
Python restatement of the reference reader and of src/mmseq.cpp:395-441.  CPU only."""
import os

import numpy as np
import pytest

from mmseq_b200 import hostlib, synth
from oracle import oracle as orc


def _same(h, o):
    assert (h.n, h.m, h.N) == (o["n"], o["m"], o["N"])
    assert np.array_equal(h.row_ptr, o["row_ptr"]) and np.array_equal(h.col, o["col"]) and np.array_equal(h.k, o["k"])
    assert [h.names[i] for i in h.col2hdr] == o["names_by_col"]
    assert np.array_equal(h.doublehits, o["doublehits"])
    assert np.allclose(h.len, o["len"], rtol=1e-15)


@pytest.mark.parametrize("fmt", ["text", "binary"])
def test_loader_matches_oracle_reader(tmp_path, small_synth, fmt):
    s = small_synth
    path = str(tmp_path / f"x.{fmt}.hits")
    ident = [[0, 1], [5, 6, 7]]
    (synth.write_hits_text if fmt == "text" else synth.write_hits_binary)(s, path, identical=ident)
    h = hostlib.load_hits(path)
    hf = orc.HitsFile(path)
    assert h.schema == hf.schema == (0 if fmt == "text" else 1)
    assert h.names == hf.names and h.T == s.T
    assert np.allclose(h.efflen, [hf.efflen[n] for n in hf.names]) and list(h.truelen) == [hf.truelen[n] for n in hf.names]
    assert h.gene_names == sorted(hf.genes)                      # std::map order
    for g, name in enumerate(h.gene_names):
        mem = [h.names[i] for i in h.gene_members[h.gene_ptr[g]:h.gene_ptr[g + 1]]]
        assert mem == hf.genes[name]
    assert [[h.names[i] for i in h.ident_members[h.ident_ptr[j]:h.ident_ptr[j + 1]]] for j in range(h.I)] == hf.identical
    _same(h, orc.build_classes(hf))
    assert h.k.sum() == s.N


def test_text_and_binary_and_in_memory_agree(tmp_path, small_synth):
    s = small_synth
    a = str(tmp_path / "a.hits"); b = str(tmp_path / "b.hits")
    synth.write_hits_text(s, a); synth.write_hits_binary(s, b)
    ha, hb = hostlib.load_hits(a), hostlib.load_hits(b)
    hc = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid)
    for x in (hb, hc):
        assert np.array_equal(ha.row_ptr, x.row_ptr) and np.array_equal(ha.col, x.col) and np.array_equal(ha.k, x.k)
        assert np.allclose(ha.len, x.len)
    assert os.path.getsize(b) < os.path.getsize(a) / 3            # "~7x" smaller per src/README.md:26


def test_per_fragment_layouts_expand_the_same_classes(small_synth):
    s = small_synth
    c = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, layout=hostlib.LAYOUT_COLLAPSED)
    for layout in (hostlib.LAYOUT_PER_FRAGMENT, hostlib.LAYOUT_PER_FRAGMENT_SORTED, hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH):
        p = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, layout=layout)
        assert p.m == s.N and p.k is None and p.n == c.n and p.n_classes == c.m
        # every per-fragment row is one of the classes; multiplicities equal k
        key = lambda rp, col, i: tuple(col[rp[i]:rp[i + 1]])
        cls = {key(c.row_ptr, c.col, i): i for i in range(c.m)}
        cnt = np.zeros(c.m, np.int64)
        ids = [cls[key(p.row_ptr, p.col, i)] for i in range(p.m)]
        np.add.at(cnt, ids, 1)
        assert np.array_equal(cnt, c.k)
        if layout == hostlib.LAYOUT_PER_FRAGMENT_SORTED:
            assert ids == sorted(ids)
        if layout == hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH:
            d = np.diff(p.row_ptr)
            assert (np.diff(d) >= 0).all()                         # grouped by class size ...
            rows = [key(p.row_ptr, p.col, i) for i in range(p.m)]
            assert all(rows[i] <= rows[i + 1] for i in range(p.m - 1) if d[i] == d[i + 1])   # ... then by member columns


def test_duplicates_and_first_appearance_order(tmp_path):
    txt = ("@TranscriptMetaData\tA\t1000\t1180\n@TranscriptMetaData\tB\t500.5\t680\n@TranscriptMetaData\tC\t2000\t2180\n"
           "@TranscriptMetaData\tD\t100\t280\n@GeneIsoforms\tg2\tC\tD\n@GeneIsoforms\tg1\tA\tB\n@IdenticalTranscripts\tA\tB\n"
           ">r1\nC\nA\n>r2\nA\nC\nA\n>r3\nB\n>r4\nA\nC\n>r5\n")
    p = tmp_path / "t.hits"
    p.write_text(txt)
    h = hostlib.load_hits(str(p))
    assert h.N == 4                                    # the trailing empty record is dropped with a warning
    assert [h.names[i] for i in h.col2hdr] == ["C", "A", "B"]   # first appearance; D never hit => not a column
    assert list(h.hdr2col) == [1, 2, 0, -1]
    assert list(h.row_ptr) == [0, 2, 3] and list(h.col) == [0, 1, 2] and list(h.k) == [3, 1]
    assert list(h.doublehits) == [0, 1, 0]
    assert h.gene_names == ["g1", "g2"]
    assert np.allclose(h.len, np.array([2000, 1000, 500.5]) * 4 / 1e9)
    _same(h, orc.build_classes(orc.HitsFile(str(p))))


@pytest.mark.parametrize("txt,msg", [
    ("@TranscriptMetaData\tA\t10\t20\n@TranscriptMetaData\tA\t10\t20\n@GeneIsoforms\tg\tA\n>r\nA\n", "duplicate transcripts in @TranscriptMetaData"),
    ("@TranscriptMetaData\tA\t10\t20\n@TranscriptMetaData\tB\t10\t20\n@GeneIsoforms\tg\tA\n>r\nA\n", "does not belong to a gene"),
    ("@TranscriptMetaData\tA\t10\t20\n@GeneIsoforms\tg\tA\n@GeneIsoforms\th\tA\n>r\nA\n", "nested within genes"),
    ("@TranscriptMetaData\tA\t10\t20\n@GeneIsoforms\tg\tA\n>r\nZ\n", "has no length"),
    ("@TranscriptMetaData\tA\t0\t20\n@GeneIsoforms\tg\tA\n>r\nA\n", "length of zero"),
    ("hello\n", "does not seem to be a hits file"),
])
def test_loader_errors_follow_the_reference(tmp_path, txt, msg):
    p = tmp_path / "bad.hits"
    p.write_text(txt)
    with pytest.raises(RuntimeError) as e:
        hostlib.load_hits(str(p))
    assert msg in str(e.value)


def test_missing_file(tmp_path):
    with pytest.raises(RuntimeError) as e:
        hostlib.load_hits(str(tmp_path / "nope.hits"))
    assert "Error reading hits file" in str(e.value)


def test_host_special_functions_match_scipy():
    """psi, psi_1 and the probit used for the closed-form rows of features without hits
    (src/mmseq.cpp:1372-1373, :1286; GSL gsl_sf_psi / gsl_sf_psi_n / gsl_cdf_ugaussian_Pinv)."""
    import ctypes as C
    from scipy import special
    L = hostlib.lib()
    for f in (L.mmq_host_digamma, L.mmq_host_trigamma, L.mmq_host_ndtri):
        f.restype = C.c_double
        f.argtypes = [C.c_double]
    for x in (0.01, 0.1, 0.5, 1.0, 2.5, 9.99, 10.0, 37.0, 1e4):
        assert np.isclose(L.mmq_host_digamma(x), special.digamma(x), rtol=1e-13, atol=1e-14)
        assert np.isclose(L.mmq_host_trigamma(x), special.polygamma(1, x), rtol=1e-13)
    for p in (1e-9, 0.01, 0.3, 0.5, 0.999999999):
        assert np.isclose(L.mmq_host_ndtri(p), special.ndtri(p), rtol=1e-13, atol=1e-15)


def test_header_order_columns_is_a_pure_renumbering(small_synth):
    """LAYOUT_HEADER_ORDER_COLUMNS: same classes and counts, transcripts numbered by header index."""
    s = small_synth
    rng = np.random.default_rng(1)
    w = rng.uniform(0.5, 2.0, len(s.frag_tid)).astype(np.float32)
    for layout, fw in ((hostlib.LAYOUT_COLLAPSED, None), (hostlib.LAYOUT_PER_FRAGMENT, w), (hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH, None)):
        a = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, frag_w=fw, layout=layout)
        b = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, frag_w=fw, layout=layout | hostlib.LAYOUT_HEADER_ORDER_COLUMNS)
        assert (np.diff(b.col2hdr) > 0).all() and sorted(a.col2hdr) == list(b.col2hdr)
        assert a.m == b.m and a.n == b.n and np.array_equal(np.diff(a.row_ptr) if layout != hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH else np.sort(np.diff(a.row_ptr)), np.diff(b.row_ptr) if layout != hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH else np.sort(np.diff(b.row_ptr)))
        for i in range(0, b.m, 97):   # rows ascending in the new numbering
            seg = b.col[b.row_ptr[i]:b.row_ptr[i + 1]]
            assert (np.diff(seg) > 0).all()
        if layout != hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH:
            # same header transcripts row by row (and the weights travel with their transcript)
            for i in range(0, a.m, 53):
                ha = a.col2hdr[a.col[a.row_ptr[i]:a.row_ptr[i + 1]]]; hb = b.col2hdr[b.col[b.row_ptr[i]:b.row_ptr[i + 1]]]
                assert sorted(ha) == sorted(hb)
                if fw is not None:
                    wa = dict(zip(ha, a.w[a.row_ptr[i]:a.row_ptr[i + 1]])); wb = dict(zip(hb, b.w[b.row_ptr[i]:b.row_ptr[i + 1]]))
                    assert wa == wb
        assert np.allclose(np.sort(a.len), np.sort(b.len))


def test_schema2_weighted_hits_round_trip(tmp_path):
    """Schema 2 (this package's extension): binary records with one fp32 weight per hit.  The loader gives the
    per-fragment CSR with the weights aligned to the columns — the same arrays as from the in-memory records."""
    s = synth.Synth(11, 120, 3000, weights=True)
    p = str(tmp_path / "w.hits")
    synth.write_hits_binary(s, p, weights=s.frag_w)
    pf = str(tmp_path / "wf.hits")
    s.write_hits_fast(pf, weights=True)
    assert open(p, "rb").read() == open(pf, "rb").read()          # the C++ writer emits the same bytes
    for lay in (hostlib.LAYOUT_PER_FRAGMENT, hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH):
        a = hostlib.load_hits(p, layout=lay)
        b = hostlib.from_records(s.T, s.efflen, s.frag_ptr, s.frag_tid, frag_w=s.frag_w, layout=lay)
        assert a.schema == 2 and a.w is not None and a.m == s.N
        assert np.array_equal(a.row_ptr, b.row_ptr) and np.array_equal(a.col2hdr[a.col], b.col2hdr[b.col]) and np.array_equal(a.w, b.w)
    with pytest.raises(RuntimeError, match="per-fragment layout"):
        hostlib.load_hits(p)                                        # the collapsed layout cannot carry weights


def test_truncated_and_degenerate_files_are_refused(tmp_path):
    """ADVICE (round 1): a zlib stream cut short, a binary record cut mid-way, a record without transcripts and a hit
    count above the number of transcripts are errors, not silently shorter samples."""
    import struct
    import zlib
    s = synth.Synth(5, 80, 2000)
    p = str(tmp_path / "ok.hits")
    synth.write_hits_binary(s, p)
    raw = open(p, "rb").read()
    assert hostlib.load_hits(p).N == 2000
    cut = str(tmp_path / "cut.hits")
    open(cut, "wb").write(raw[:len(raw) // 2])                      # truncated zlib stream
    with pytest.raises(RuntimeError, match="corrupt"):
        hostlib.load_hits(cut)
    plain = zlib.decompress(raw)
    mid = str(tmp_path / "mid.hits")
    open(mid, "wb").write(zlib.compress(plain[:-6], 1))             # complete stream, last record cut inside its index list
    with pytest.raises(RuntimeError, match="malformed"):
        hostlib.load_hits(mid)
    empty = str(tmp_path / "empty.hits")
    open(empty, "wb").write(zlib.compress(plain + b"rx\n" + struct.pack("<I", 0), 1))   # a record with no transcripts
    with pytest.raises(RuntimeError, match="without any mapping transcripts"):
        hostlib.load_hits(empty)
    big = str(tmp_path / "big.hits")
    open(big, "wb").write(zlib.compress(plain + b"ry\n" + struct.pack("<I", 4000000000), 1))   # an absurd hit count
    with pytest.raises(RuntimeError, match="malformed"):
        hostlib.load_hits(big)
    txt = str(tmp_path / "t.hits")
    synth.write_hits_text(synth.Synth(5, 20, 50), txt)
    body = open(txt).read().replace(">r10\n", ">r10\n>r10b\n", 1)   # text: a read name directly followed by the next one
    open(txt, "w").write(body)
    with pytest.raises(RuntimeError, match="without any mapping transcripts"):
        hostlib.load_hits(txt)
