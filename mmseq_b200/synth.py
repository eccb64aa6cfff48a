"""Seeded synthetic workloads (harness): ctypes view of libmmq_synth.so plus
writers for real ``.hits`` files in both of the reference's schemas
(src/hitsio.cpp:162-240; README.md:388-403).  The reference ships no sample
data, so every test and bench input is generated here."""
import ctypes as C
import os
import struct
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmq_synth.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built: run `make synth`")
        L = C.CDLL(LIB_PATH)
        L.mmq_synth_create.restype = C.c_void_p
        L.mmq_synth_create.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int]
        L.mmq_synth_destroy.argtypes = [C.c_void_p]
        L.mmq_synth_write_hits.restype = C.c_int
        L.mmq_synth_write_hits.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        for nm in ["T", "G", "N", "nhits"]:
            f = getattr(L, "mmq_synth_" + nm)
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p]
        for nm in ["gene_of", "gene_ptr", "efflen", "truelen", "mu", "frag_ptr", "frag_tid", "frag_w"]:
            f = getattr(L, "mmq_synth_" + nm)
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _arr(ptr, n, ct, dtype):
    if not ptr:
        return None
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).astype(dtype, copy=True)


class Synth:
    """Transcriptome + per-fragment hit lists (header transcript indices)."""

    def __init__(self, seed, T, N, haplo=False, weights=False, threads=0, frag_seed=0):
        L = lib()
        h = L.mmq_synth_create(seed, frag_seed, T // (2 if haplo else 1), N, int(haplo), int(weights), threads)
        self.T = int(L.mmq_synth_T(h)); self.G = int(L.mmq_synth_G(h)); self.N = int(L.mmq_synth_N(h))
        nh = int(L.mmq_synth_nhits(h))
        self.gene_of = _arr(L.mmq_synth_gene_of(h), self.T, C.c_int32, np.int32)
        self.gene_ptr = _arr(L.mmq_synth_gene_ptr(h), self.G + 1, C.c_int64, np.int64)
        self.efflen = _arr(L.mmq_synth_efflen(h), self.T, C.c_double, np.float64)
        self.truelen = _arr(L.mmq_synth_truelen(h), self.T, C.c_int32, np.int32)
        self.mu = _arr(L.mmq_synth_mu(h), self.T, C.c_double, np.float64)
        self.frag_ptr = _arr(L.mmq_synth_frag_ptr(h), self.N + 1, C.c_int64, np.int64)
        self.frag_tid = _arr(L.mmq_synth_frag_tid(h), nh, C.c_int32, np.int32)
        self.frag_w = _arr(L.mmq_synth_frag_w(h), nh, C.c_float, np.float32)
        self._keep = None
        L.mmq_synth_destroy(h)
        self.haplo = haplo
        self._args = (seed, frag_seed, T // (2 if haplo else 1), N, int(haplo), int(weights), threads)

    def write_hits_fast(self, path, binary=True, weights=False):
        """C++ writer for large files (no identical-transcript records); same bytes as the Python writers.
        weights=True: schema 2 (binary + one fp32 weight per hit; the generator must have been made with weights)."""
        L = lib()
        h = L.mmq_synth_create(*self._args)
        rc = L.mmq_synth_write_hits(h, os.fsencode(path), 2 if weights else (1 if binary else 0), int(self.haplo))
        L.mmq_synth_destroy(h)
        if rc:
            raise RuntimeError(f"cannot write {path}")

    # names: gene ids are zero-padded so that std::map (byte-wise) order == numeric order
    def transcript_name(self, t):
        if self.haplo:
            return f"T{t // 2:07d}_{'AB'[t % 2]}"
        return f"T{t:07d}"

    def gene_name(self, g):
        return f"G{g:07d}"


def _fmt_g6(x):
    """operator<<(ostream, double) with default precision 6 (src/hitsio.cpp:168, :7-12)."""
    return "%g" % x


def write_hits_text(s, path, identical=None):
    """Schema 0 (src/hitsio.cpp:162-187)."""
    with open(path, "w") as f:
        for t in range(s.T):
            f.write(f"@TranscriptMetaData\t{s.transcript_name(t)}\t{_fmt_g6(s.efflen[t])}\t{int(s.truelen[t])}\n")
        for g in range(s.G):
            mem = "\t".join(s.transcript_name(t) for t in range(int(s.gene_ptr[g]), int(s.gene_ptr[g + 1])))
            f.write(f"@GeneIsoforms\t{s.gene_name(g)}\t{mem}\n")
        for grp in identical or []:
            f.write("@IdenticalTranscripts\t" + "\t".join(s.transcript_name(t) for t in grp) + "\n")
        fp, ft = s.frag_ptr, s.frag_tid
        for r in range(s.N):
            f.write(f">r{r}\n")
            for q in range(int(fp[r]), int(fp[r + 1])):
                f.write(s.transcript_name(int(ft[q])) + "\n")


def _u32(v):
    return struct.pack("<I", v)


def _small(v):
    return bytes([v]) if v < 255 else b"\xff" + _u32(v)


def write_hits_binary(s, path, identical=None, weights=None):
    """Schema 1, whole stream zlib-compressed (src/hitsio.cpp:189-213, :232-240, :77-99).  With `weights` (fp32, one
    per hit, aligned with s.frag_tid): schema 2, this package's extension — the weights follow a record's indices."""
    out = bytearray()
    out += b"MMSEQ_HITSFILE\n" + _u32(1 if weights is None else 2)
    out += _u32(s.T)
    for t in range(s.T):
        out += s.transcript_name(t).encode() + b"\n" + _fmt_g6(s.efflen[t]).encode() + b"\n" + _u32(int(s.truelen[t]))
    out += _u32(s.G)
    for g in range(s.G):
        b, e = int(s.gene_ptr[g]), int(s.gene_ptr[g + 1])
        out += s.gene_name(g).encode() + b"\n" + _u32(e - b)
        for t in range(b, e):
            out += s.transcript_name(t).encode() + b"\n"
    ident = identical or []
    out += _u32(len(ident))
    for grp in ident:
        out += _u32(len(grp))
        for t in grp:
            out += s.transcript_name(t).encode() + b"\n"
    prev = ""
    fp, ft = s.frag_ptr, s.frag_tid
    for r in range(s.N):
        name = f"r{r}"
        nb = 0
        while nb < min(len(prev), len(name)) and prev[nb] == name[nb]:
            nb += 1
        ne = 0
        while nb + ne < min(len(prev), len(name)) and prev[len(prev) - 1 - ne] == name[len(name) - 1 - ne]:
            ne += 1
        if nb == 0 and ne == 0:
            out += name.encode() + b"\n"
        else:
            out += b"\n" + _small(nb) + name[nb:len(name) - ne].encode() + b"\n" + _small(ne)
        prev = name
        b, e = int(fp[r]), int(fp[r + 1])
        out += _u32(e - b)
        out += np.asarray(ft[b:e], dtype="<u4").tobytes()
        if weights is not None:
            out += np.asarray(weights[b:e], dtype="<f4").tobytes()
    with open(path, "wb") as f:
        f.write(zlib.compress(bytes(out), 1))
