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


@pytest.mark.parametrize("parallel", [False, True])
def test_truncated_and_degenerate_files_are_refused(tmp_path, monkeypatch, parallel):
    """(parallel: the all-threads loader path, which small files do not take by default.)  ADVICE (round 1): a zlib stream cut short, a binary record cut mid-way, a record without transcripts and a hit
    count above the number of transcripts are errors, not silently shorter samples."""
    import struct
    import zlib
    if parallel:
        monkeypatch.setenv("MMQ_LOADER_PAR_MIN_BYTES", "0")
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


# ---- the loader on all host threads: parallel inflate (inflate_par.h) and parallel class building ------------------

def _streams():
    import zlib
    rng = np.random.default_rng(11)
    for it in range(60):
        n = int(2 ** rng.uniform(8, 23))
        kind = it % 5
        if kind == 0:
            raw = rng.integers(0, 256, n, dtype=np.uint8).tobytes()                      # incompressible: stored blocks
        elif kind == 1:
            raw = bytes(rng.choice(np.frombuffer(b"ACGT\n>r", np.uint8), n))
        elif kind == 2:
            raw = (np.arange(n) // 3 % 251).astype(np.uint8).tobytes()                   # long matches
        elif kind == 3:
            raw = np.where(rng.random(n) < 0.9, 0, rng.integers(0, 256, n)).astype(np.uint8).tobytes()
        else:
            raw = b"".join(b"\n\x02%d\n\x00" % i + int(i % 977).to_bytes(4, "little") * (1 + i % 5) for i in range(n // 16))
        level = [1, 1, 6, 9, 0][int(rng.integers(5))]
        strategy = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE][int(rng.integers(5))]
        c = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        yield raw, c.compress(raw) + c.flush(), int(rng.integers(1, 9))


def test_parallel_inflate_equals_zlib_on_every_block_type():
    refused = 0
    for raw, z, threads in _streams():
        out = hostlib.inflate_parallel(z, threads, len(raw) + 16)
        if out is None:
            refused += 1          # allowed (the loader then falls back to zlib), but must be rare
            continue
        assert out == raw
    assert refused <= 3


def test_parallel_inflate_refuses_damaged_streams():
    import zlib
    rng = np.random.default_rng(5)
    raw = bytes(rng.choice(np.frombuffer(b"ACGT\n>r0123", np.uint8), 3 << 20))
    z = bytearray(zlib.compress(raw, 1))
    for it in range(24):
        zz = bytearray(z)
        if it % 2:
            del zz[len(zz) * (1 + it % 7) // 9:]                  # truncated
        else:
            zz[len(zz) // 3 + it * 1013] ^= 1 << (it % 8)         # one flipped bit
        out = hostlib.inflate_parallel(bytes(zz), 4, len(raw) + 16)
        assert out is None or out == raw                           # never wrong data


class _Recs:
    """A synthetic sample whose records repeat transcripts (doublehits, src/mmseq.cpp:404-409) and include very long ones."""
    def __init__(self, s, seed=3):
        rng = np.random.default_rng(seed)
        self.__dict__.update({k: getattr(s, k) for k in ("T", "G", "efflen", "truelen", "gene_ptr")})
        self._s = s
        ptr, tid = [0], []
        for r in range(s.N):
            ids = list(s.frag_tid[s.frag_ptr[r]:s.frag_ptr[r + 1]])
            u = rng.random()
            if u < 0.05:
                ids = ids + ids[:1]                                # a repeated transcript
            elif u < 0.06:
                ids = list(rng.integers(0, s.T, 50)) + ids * 3     # a long record with repeats
            elif u < 0.10:
                ids = ids[::-1]                                    # same set, other order
            tid += ids
            ptr.append(len(tid))
        self.frag_ptr = np.array(ptr, np.int64)
        self.frag_tid = np.array(tid, np.int32)
        self.N = s.N
    def transcript_name(self, t): return self._s.transcript_name(t)
    def gene_name(self, g): return self._s.gene_name(g)


def _load_env(path, layout, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return hostlib.load_hits(path, layout)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


LAYOUTS = [hostlib.LAYOUT_COLLAPSED, hostlib.LAYOUT_PER_FRAGMENT, hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH,
           hostlib.LAYOUT_COLLAPSED | hostlib.LAYOUT_HEADER_ORDER_COLUMNS]


def _equal(a, b):
    assert (a.n, a.m, a.N, a.nnz) == (b.n, b.m, b.N, b.nnz)
    for f in ("row_ptr", "col", "col2hdr", "hdr2col", "doublehits"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert (a.k is None and b.k is None) or np.array_equal(a.k, b.k)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_parallel_loader_equals_the_sequential_walk(tmp_path, small_synth, layout):
    """Whole file inflated in memory + parallel builders (first-appearance numbering reconstructed afterwards) against the
    record-by-record path: identical classes, numbering, counts, doublehits."""
    recs = _Recs(small_synth)
    path = str(tmp_path / "dups.hits")
    synth.write_hits_binary(recs, path)
    serial = _load_env(path, layout, MMQ_LOADER_SERIAL_INFLATE=1)
    for builders, workers in ((1, 1), (3, 2), (5, 3)):
        par = _load_env(path, layout, MMQ_LOADER_PAR_MIN_BYTES=0, MMQ_LOADER_BUILDERS=builders, MMQ_LOADER_THREADS=workers)
        _equal(par, serial)
    if layout == hostlib.LAYOUT_COLLAPSED:
        _same(serial, orc.build_classes(orc.HitsFile(path)))      # and both equal the restated reference reader


def test_in_memory_records_parallel_equals_sequential(small_synth):
    recs = _Recs(small_synth, seed=8)
    for layout in LAYOUTS + [hostlib.LAYOUT_COLLAPSED | hostlib.LAYOUT_IDENTITY_COLUMNS]:
        os.environ["MMQ_LOADER_SERIAL_RECORDS"] = "1"
        try:
            a = hostlib.from_records(recs.T, recs.efflen, recs.frag_ptr, recs.frag_tid, layout=layout)
        finally:
            os.environ.pop("MMQ_LOADER_SERIAL_RECORDS")
        b = hostlib.from_records(recs.T, recs.efflen, recs.frag_ptr, recs.frag_tid, layout=layout)
        _equal(a, b)


WLAYOUTS = [hostlib.LAYOUT_PER_FRAGMENT, hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH, hostlib.LAYOUT_PER_FRAGMENT_SORTED,
            hostlib.LAYOUT_PER_FRAGMENT_BY_LENGTH | hostlib.LAYOUT_HEADER_ORDER_COLUMNS]


def _equal_w(a, b):
    _equal(a, b)
    assert a.w is not None and b.w is not None and np.array_equal(a.w, b.w)     # bit for bit: same sums in the same order


@pytest.mark.parametrize("layout", WLAYOUTS)
def test_weighted_parallel_loader_equals_the_sequential_walk(tmp_path, small_synth, layout):
    """Schema 2 (one fp32 weight per hit): weights of repeated transcripts are summed in record order, every record's
    weights follow its class's column order — the parallel path against the record-by-record one."""
    recs = _Recs(small_synth, seed=5)
    rng = np.random.default_rng(12)
    wts = rng.lognormal(0.0, 0.5, len(recs.frag_tid)).astype(np.float32)
    path = str(tmp_path / "w.hits")
    synth.write_hits_binary(recs, path, weights=wts)
    serial = _load_env(path, layout, MMQ_LOADER_SERIAL_INFLATE=1)
    assert serial.schema == 2
    for builders, workers in ((1, 1), (4, 3)):
        par = _load_env(path, layout, MMQ_LOADER_PAR_MIN_BYTES=0, MMQ_LOADER_BUILDERS=builders, MMQ_LOADER_THREADS=workers)
        _equal_w(par, serial)
    os.environ["MMQ_LOADER_SERIAL_RECORDS"] = "1"
    try:
        a = hostlib.from_records(recs.T, recs.efflen, recs.frag_ptr, recs.frag_tid, frag_w=wts, layout=layout)
    finally:
        os.environ.pop("MMQ_LOADER_SERIAL_RECORDS")
    b = hostlib.from_records(recs.T, recs.efflen, recs.frag_ptr, recs.frag_tid, frag_w=wts, layout=layout)
    _equal_w(a, b)
    _equal_w(a, serial)


def test_weighted_parallel_loader_refuses_bad_weights(tmp_path, small_synth):
    s = small_synth
    wts = np.ones(len(s.frag_tid), np.float32)
    wts[len(wts) // 2] = -1.0
    path = str(tmp_path / "bad.hits")
    synth.write_hits_binary(s, path, weights=wts)
    for env in ({"MMQ_LOADER_SERIAL_INFLATE": 1}, {"MMQ_LOADER_PAR_MIN_BYTES": 0}):
        with pytest.raises(RuntimeError) as e:
            _load_env(path, hostlib.LAYOUT_PER_FRAGMENT, **env)
        assert "weights" in str(e.value)
