"""oracle.py — Python side of the CPU ORACLE (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product never does.

Contents
  * ctypes view of oracle/liboracle.so (mmseq_oracle.cpp: EM, Gibbs replay,
    GSL-style chain, Sokal, uh) and of oracle/_ref/libsokal_ref.so (the
    reference's own src/sokal.cc compiled from /root/reference);
  * pure-Python restatements, for small inputs, of the hits reader
    (src/hitsio.cpp:250-447), of class construction (src/mmseq.cpp:395-441)
    and of the posterior summaries (src/mmseq.cpp:938-1395).

PARITY STATUS: see the header of mmseq_oracle.cpp — pinned against the reference's own sources
(sokal.cc directly; mmseq.cpp/hitsio.cpp/uh.cpp through oracle/shim) except for GSL's arithmetic.
"""
import ctypes as C
import os
import struct
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_SOKAL_PATH = os.path.join(_HERE, "_ref", "libsokal_ref.so")

_lib = None
_ref = None

vp, i32, i64, u32, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_double


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle`")
        L = C.CDLL(LIB_PATH)
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]
        L.orc_draw_uniform.argtypes = [u32, i64, vp]
        L.orc_draw_gamma.argtypes = [u32, i64, dbl, dbl, vp]
        L.orc_draw_normal.argtypes = [u32, i64, vp]
        L.orc_draw_binomial.argtypes = [u32, i64, i64, dbl, vp]
        L.orc_draw_alloc.argtypes = [u32, i64, i32, vp, i64, vp]
        L.orc_math.argtypes = [i32, i64, vp, vp]
        L.orc_init_mu.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, vp]
        L.orc_loglik.restype = dbl
        L.orc_loglik.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp]
        L.orc_em.restype = i32
        L.orc_em.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, i32, dbl, vp, vp]
        L.orc_sweep_replay.argtypes = [i64, i64, vp, vp, vp, vp, vp, dbl, dbl, u32, u32, i64, vp, vp, vp, i32]
        L.orc_sweep_replay_ids.argtypes = [i64, i64, vp, vp, vp, vp, vp, dbl, dbl, u32, u32, vp, vp, vp, vp, i32]
        L.orc_gamma_replay.argtypes = [i64, vp, vp, dbl, dbl, u32, u32, vp]
        L.orc_sweep_counts.argtypes = [i64, i64, vp, vp, vp, vp, u32, u32, i64, vp, vp, vp, i32]
        L.orc_em_partial.restype = dbl
        L.orc_em_partial.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, i32]
        L.orc_cls_plan_replay.restype = C.c_int
        L.orc_cls_plan_replay.argtypes = [i64, i64, vp, vp, vp, vp, i64, vp, u32, u32, vp, vp]
        L.orc_gibbs_replay.argtypes = [i64, i64, vp, vp, vp, vp, vp, dbl, dbl, u32, i64, i64, i32, i32, vp, vp]
        L.orc_prior_replay.argtypes = [i64, vp, vp, dbl, dbl, u32, i32, vp]
        L.orc_gsl_binomial.argtypes = [u32, i64, i64, dbl, vp]
        L.orc_gsl_gamma.argtypes = [u32, i64, dbl, dbl, vp]
        L.orc_gsl_multinomial.argtypes = [u32, i64, i32, vp, i64, vp]
        L.orc_max_threads.restype = i32
        L.orc_gibbs_gsl.restype = dbl
        L.orc_gibbs_gsl.argtypes = [i64, i64, vp, vp, vp, vp, dbl, dbl, i32, i32, i64, i32, i32, vp, vp]
        L.orc_sokal.restype = i32
        L.orc_sokal.argtypes = [i32, vp, vp, vp, vp]
        L.orc_uh_literal.argtypes = [i64, vp, vp, vp, i64, vp, vp, vp]
        L.orc_cov.argtypes = [vp, i32, i64, vp]
        L.orc_mean_corrs.argtypes = [vp, vp, i64, i32, vp, i64, dbl, vp, vp]
        _lib = L
    return _lib


def ref_sokal_lib():
    """The reference's own sokal() (src/sokal.cc), or None if oracle/_ref was not built."""
    global _ref
    if _ref is None and os.path.exists(REF_SOKAL_PATH):
        R = C.CDLL(REF_SOKAL_PATH)
        R.sokal.restype = i32
        R.sokal.argtypes = [vp, vp, vp, vp, vp]
        _ref = R
    return _ref


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


# ------------------------------------------------------------------ samplers

def philox(ctr, key):
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def draw_uniform(seed, cnt):
    o = np.zeros(cnt); lib().orc_draw_uniform(seed, cnt, _p(o)); return o


def draw_normal(seed, cnt):
    o = np.zeros(cnt); lib().orc_draw_normal(seed, cnt, _p(o)); return o


def draw_gamma(seed, cnt, a, rate):
    o = np.zeros(cnt); lib().orc_draw_gamma(seed, cnt, a, rate, _p(o)); return o


def draw_binomial(seed, cnt, n, p):
    o = np.zeros(cnt, np.int64); lib().orc_draw_binomial(seed, cnt, n, p, _p(o)); return o


def draw_alloc(seed, cnt, p, k):
    p = _c(p, np.float64); o = np.zeros((cnt, len(p)), np.int32)
    lib().orc_draw_alloc(seed, cnt, len(p), _p(p), k, _p(o)); return o


def math_fn(which, x):
    x = _c(x, np.float64); o = np.zeros_like(x)
    lib().orc_math({"log": 0, "exp": 1, "ndtri": 2, "log1p": 3, "cos2pi_u32": 4}[which], x.size, _p(x), _p(o)); return o


def gsl_binomial(seed, cnt, n, p):
    o = np.zeros(cnt, np.int64); lib().orc_gsl_binomial(seed, cnt, n, p, _p(o)); return o


def gsl_gamma(seed, cnt, a, scale):
    o = np.zeros(cnt); lib().orc_gsl_gamma(seed, cnt, a, scale, _p(o)); return o


def gsl_multinomial(seed, cnt, p, k):
    p = _c(p, np.float64); o = np.zeros((cnt, len(p)), np.int32)
    lib().orc_gsl_multinomial(seed, cnt, len(p), _p(p), k, _p(o)); return o


# ------------------------------------------------------------- EM and Gibbs

class Problem:
    """CSR hit-class matrix + k + l, as mmq_problem."""

    def __init__(self, row_ptr, col, k, length, weight=None, alpha=0.1, beta=0.1):
        self.row_ptr = _c(row_ptr, np.int64); self.col = _c(col, np.int32)
        self.k = _c(k, np.int32); self.w = _c(weight, np.float32); self.len = _c(length, np.float64)
        self.m = len(self.row_ptr) - 1; self.n = len(self.len); self.nnz = int(self.row_ptr[-1])
        self.alpha = alpha; self.beta = beta

    def init_mu(self):
        mu = np.zeros(self.n); uh = np.zeros(self.n, np.int32); cs = np.zeros((self.n, 100), np.int32)
        lib().orc_init_mu(self.m, self.n, _p(self.row_ptr), _p(self.col), _p(self.k), _p(self.len), _p(mu), _p(uh), _p(cs))
        return mu, uh, cs

    def loglik(self, mu):
        mu = _c(mu, np.float64)
        return lib().orc_loglik(self.m, self.n, _p(self.row_ptr), _p(self.col), _p(self.k), _p(self.w), _p(self.len), _p(mu))

    def em(self, mu, max_iter=1000, eps=0.1):
        mu = np.array(mu, np.float64); ll = dbl(); llr = dbl()
        it = lib().orc_em(self.m, self.n, _p(self.row_ptr), _p(self.col), _p(self.k), _p(self.w), _p(self.len), _p(mu),
                          max_iter, eps, C.byref(ll), C.byref(llr))
        return mu, it, ll.value, llr.value

    def sweep_replay(self, mu, seed, sweep, class_id_base=0, do_gamma=True, rows=None):
        """One sweep on the shared Philox stream; returns (x, counts, mu_after)."""
        mu = np.array(mu, np.float64)
        rp, col, k, w = self.row_ptr, self.col, self.k, self.w
        if rows is not None:  # a shard: rows [a, b)
            a, b = rows
            lo, hi = int(rp[a]), int(rp[b])
            col = np.ascontiguousarray(col[lo:hi]); w = None if w is None else np.ascontiguousarray(w[lo:hi])
            k = None if k is None else np.ascontiguousarray(k[a:b]); rp = np.ascontiguousarray(rp[a:b + 1] - lo)
        m = len(rp) - 1
        x = np.zeros(int(rp[-1]), np.int32); counts = np.zeros(self.n, np.int32)
        lib().orc_sweep_replay(m, self.n, _p(rp), _p(col), _p(k), _p(w), _p(self.len), self.alpha, self.beta,
                               seed, sweep, class_id_base, _p(mu), _p(x), _p(counts), int(do_gamma))
        return x, counts, mu

    def cls_plan_replay(self, mu, seed, sweep, class_id=None, class_id_base=0):
        """Counts of one sweep through the product's class plan built on the CPU (collapsed shards):
        returns (counts, stats dict) or (None, None) when the plan does not apply."""
        mu = np.array(mu, np.float64)
        cid = None if class_id is None else _c(class_id, np.int64)
        counts = np.zeros(self.n, np.int32); stats = np.zeros(8, np.int64)
        rc = lib().orc_cls_plan_replay(self.m, self.n, _p(self.row_ptr), _p(self.col), _p(self.k), _p(cid), class_id_base,
                                       _p(mu), seed, sweep, _p(counts), _p(stats))
        if rc:
            return None, None
        return counts, dict(zip(["in_use", "small_classes", "packed_slots", "class_slots", "rest_classes", "rest_nnz", "chain_classes", "chain_slots"], [int(v) for v in stats]))

    def sweep_replay_ids(self, mu, seed, sweep, class_id, do_gamma=True):
        """One sweep with explicit Philox counters per class (any class order)."""
        mu = np.array(mu, np.float64); cid = _c(class_id, np.int64)
        x = np.zeros(self.nnz, np.int32); counts = np.zeros(self.n, np.int32)
        lib().orc_sweep_replay_ids(self.m, self.n, _p(self.row_ptr), _p(self.col), _p(self.k), _p(self.w), _p(self.len), self.alpha,
                                   self.beta, seed, sweep, _p(cid), _p(mu), _p(x), _p(counts), int(do_gamma))
        return x, counts, mu

    def sweep_counts(self, mu, seed, sweep, class_id_base=0, class_id=None, threads=0):
        """Allocation step of one sweep on the shared Philox stream, on `threads` host threads: counts only."""
        mu = _c(mu, np.float64); cid = None if class_id is None else _c(class_id, np.int64)
        counts = np.zeros(self.n, np.int32)
        lib().orc_sweep_counts(self.m, self.n, _p(self.row_ptr), _p(self.col), _p(self.k), _p(self.w), seed, sweep, class_id_base,
                               _p(cid), _p(mu), _p(counts), threads)
        return counts

    def em_partial(self, mu, threads=0):
        """This shard's part of one EM iteration: (acc[n], sum_i k_i log D_i)."""
        mu = _c(mu, np.float64); acc = np.zeros(self.n)
        ll = lib().orc_em_partial(self.m, self.n, _p(self.row_ptr), _p(self.col), _p(self.k), _p(self.w), _p(mu), _p(acc), threads)
        return acc, ll

    def gamma_replay(self, counts, seed, sweep):
        counts = _c(counts, np.int32); mu = np.zeros(self.n)
        lib().orc_gamma_replay(self.n, _p(counts), _p(self.len), self.alpha, self.beta, seed, sweep, _p(mu))
        return mu

    def gibbs_replay(self, mu, seed, first_sweep, n_sweeps, stride, trace_len):
        mu = np.array(mu, np.float64); trace = np.zeros((self.n, trace_len)) if trace_len else None
        lib().orc_gibbs_replay(self.m, self.n, _p(self.row_ptr), _p(self.col), _p(self.k), _p(self.w), _p(self.len),
                               self.alpha, self.beta, seed, first_sweep, n_sweeps, stride, trace_len, _p(mu), _p(trace))
        return mu, trace

    def gibbs_gsl(self, mu, seed, n_sweeps, stride=16, trace_len=0, threads=0):
        """The reference's own loop (MT19937 per thread, GSL-style samplers). Returns (mu, trace, seconds)."""
        assert self.w is None, "the reference has no per-hit weights"
        mu = np.array(mu, np.float64); trace = np.zeros((self.n, trace_len)) if trace_len else None
        sec = lib().orc_gibbs_gsl(self.m, self.n, _p(self.row_ptr), _p(self.col), _p(self.k), _p(self.len), self.alpha,
                                  self.beta, seed, threads, n_sweeps, stride, trace_len, _p(mu), _p(trace))
        return mu, trace, sec


def max_threads():
    return int(lib().orc_max_threads())


def prior_replay(ids, lscaled, alpha, beta, seed, trace_len):
    ids = _c(ids, np.int64); ls = _c(lscaled, np.float64); out = np.zeros((len(ids), trace_len))
    lib().orc_prior_replay(len(ids), _p(ids), _p(ls), alpha, beta, seed, trace_len, _p(out))
    return out


def sokal(x):
    """Restated sokal(); returns (rc, var, tau, m)."""
    x = np.array(x, np.float64); var = dbl(); tau = dbl(); m = i32()
    rc = lib().orc_sokal(len(x), _p(x), C.byref(var), C.byref(tau), C.byref(m))
    return rc, var.value, tau.value, m.value


def sokal_reference(x):
    """The reference's own compiled sokal() (None if oracle/_ref is absent)."""
    R = ref_sokal_lib()
    if R is None:
        return None
    x = np.array(x, np.float64); n = i32(len(x)); var = dbl(); tau = dbl(); m = i32()
    rc = R.sokal(C.byref(n), _p(x), C.byref(var), C.byref(tau), C.byref(m))
    return rc, var.value, tau.value, m.value


def uh_literal(row_ptr, col, k, set_ptr, set_members):
    rp = _c(row_ptr, np.int64); col = _c(col, np.int32); k = _c(k, np.int32)
    sp = _c(set_ptr, np.int64); sm = _c(set_members, np.int32); out = np.zeros(len(sp) - 1, np.int32)
    lib().orc_uh_literal(len(rp) - 1, _p(rp), _p(col), _p(k), len(sp) - 1, _p(sp), _p(sm), _p(out))
    return out


def trace_cov(M):
    """get_corrs()'s per-sample step, src/mmcollapse.cpp:553-558: cov() of the L x C trace matrix M[i, c] (non-finite -> 0).
    Returns the C x C matrix (symmetric)."""
    M = np.asarray(M, np.float64)
    L, Cn = M.shape
    Mf = np.asfortranarray(M)                       # column c contiguous, as arma::mat
    R = np.zeros((Cn, Cn), np.float64, order="F")
    lib().orc_cov(Mf.ctypes.data_as(vp), L, Cn, R.ctypes.data_as(vp))
    return R


def mean_corrs(R, S, ts, sdpenalty=0.0, V=None, W=None):
    """mean_corrs(), src/mmcollapse.cpp:483-511.  R: (ns, C, C) covariance slices, S: (C, ns) 0/1, ts: rows to refresh.
    Returns (V, W), C x C (entries outside the rows / columns of ts keep what was passed in, zeros by default)."""
    R = np.ascontiguousarray(R, np.float64)          # symmetric slices: row- and column-major coincide
    ns, Cn, _ = R.shape
    Sf = np.asfortranarray(np.asarray(S, np.uint8))
    ts = _c(ts, np.int32)
    V = np.zeros((Cn, Cn), np.float64, order="F") if V is None else np.asfortranarray(V, np.float64)
    W = np.zeros((Cn, Cn), np.float64, order="F") if W is None else np.asfortranarray(W, np.float64)
    lib().orc_mean_corrs(_p(R), Sf.ctypes.data_as(vp), Cn, ns, _p(ts), len(ts), float(sdpenalty), V.ctypes.data_as(vp), W.ctypes.data_as(vp))
    return V, W


# ------------------------------------------------- hits reader (pure Python)

class HitsFile:
    """src/hitsio.cpp:250-447 restated: header tables + list of records (lists of transcript NAMES)."""

    def __init__(self, path):
        raw = open(path, "rb").read()
        if raw[:1] == b"\x78":  # :258
            raw = zlib.decompress(raw)
        first = raw.split(b"\n", 1)[0]
        self.names, self.efflen, self.truelen = [], {}, {}
        self.genes, self.identical, self.records = {}, [], []
        if first.split()[:1] == [b"@TranscriptMetaData"]:
            self.schema = 0
            self._text(raw.decode())
        else:
            self.schema = 1
            self._binary(raw)

    def _text(self, s):  # :286-347
        lines = s.split("\n")
        i = 0
        while i < len(lines) and not lines[i].startswith(">"):
            tok = lines[i].split()
            if not tok:
                i += 1
                continue
            if tok[0] == "@TranscriptMetaData":
                self.names.append(tok[1]); self.efflen.setdefault(tok[1], float(tok[2])); self.truelen.setdefault(tok[1], int(tok[3]))
            elif tok[0] == "@GeneIsoforms":
                self.genes.setdefault(tok[1], tok[2:])
            elif tok[0] == "@IdenticalTranscripts":
                self.identical.append(tok[1:])
            else:
                raise ValueError("Hits file looks malformed.")
            i += 1
        cur = None
        for ln in lines[i:]:
            if ln.startswith(">"):
                cur = []
                self.records.append(cur)
            elif ln != "" or cur is None:
                cur.append(ln)
        if self.records and not self.records[-1]:
            self.records.pop()  # trailing record without transcripts: warning + stop (:336-340)

    def _binary(self, b):  # :349-439
        pos = b.index(b"\n") + 1
        (schema,) = struct.unpack_from("<I", b, pos); pos += 4
        assert schema == 1

        def line():
            nonlocal pos
            e = b.index(b"\n", pos); s = b[pos:e].decode(); pos = e + 1
            return s

        def u32v():
            nonlocal pos
            (v,) = struct.unpack_from("<I", b, pos); pos += 4
            return v

        def small():
            nonlocal pos
            v = b[pos]; pos += 1
            return u32v() if v == 255 else v

        for _ in range(u32v()):
            nm = line(); el = line(); tl = u32v()
            self.names.append(nm); self.efflen.setdefault(nm, float(el)); self.truelen.setdefault(nm, tl)
        for _ in range(u32v()):
            g = line(); c = u32v()
            self.genes.setdefault(g, [line() for _ in range(c)])
        for _ in range(u32v()):
            c = u32v()
            self.identical.append([line() for _ in range(c)])
        while pos < len(b):
            nm = line()
            if nm == "":
                small(); line(); small()
            c = u32v()
            self.records.append([self.names[u32v()] for _ in range(c)])


def build_classes(hf):
    """src/mmseq.cpp:395-441 restated.  Returns dict(n, m, N, row_ptr, col, k, names_by_col, doublehits)."""
    sid_index, index_sid, doublehits = {}, [], []
    index_comb, k, rows = {}, [], []
    N = 0
    for rec in hf.records:
        N += 1
        comb = []
        for name in rec:
            if name not in sid_index:
                sid_index[name] = len(index_sid); index_sid.append(name); doublehits.append(0)
            c = sid_index[name]
            if c not in comb:
                comb.append(c)
            else:
                doublehits[c] += 1
        comb = tuple(sorted(comb))
        if comb not in index_comb:
            index_comb[comb] = len(rows); rows.append(comb); k.append(0)
        k[index_comb[comb]] += 1
    row_ptr = np.zeros(len(rows) + 1, np.int64)
    row_ptr[1:] = np.cumsum([len(r) for r in rows])
    col = np.array([c for r in rows for c in r], np.int32)
    length = np.array([hf.efflen[nm] * N / 1000000000.0 for nm in index_sid])  # :603
    return dict(n=len(index_sid), m=len(rows), N=N, row_ptr=row_ptr, col=col, k=np.array(k, np.int32),
                names_by_col=index_sid, doublehits=np.array(doublehits, np.int32), len=length, sid_index=sid_index)


# --------------------------------------------------- summaries (numpy, small)

def summaries_transcripts(trace, pct=(5, 25, 50, 75, 95)):
    """log_mu, sd, mcse, iact, percentiles for a (rows, L) raw-scale trace
    (src/mmseq.cpp:1111-1146, :1203-1227, :1308-1324)."""
    trace = np.asarray(trace, np.float64)
    rows, L = trace.shape
    idx = [int(np.floor(p / 100.0 * (L - 1) + 0.5)) for p in pct]  # :1113 (C round(): half away from zero)
    srt = np.sort(trace, axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        lg = np.log(trace)
    out = dict(log_mu=lg.mean(axis=1) if L else None, pct=srt[:, idx])
    # mean as the reference accumulates it: sequential sum then divide
    out["log_mu"] = np.array([np.add.reduce(lg[r]) / L for r in range(rows)])
    sd = np.zeros(rows); mcse = np.zeros(rows); iact = np.zeros(rows); var = np.zeros(rows); win = np.zeros(rows, np.int32)
    for r in range(rows):
        rc, v, tau, m = sokal(lg[r])
        var[r] = v; win[r] = m
        if rc != 0:
            mcse[r] = L; iact[r] = np.nan
        else:
            with np.errstate(invalid="ignore"):
                mcse[r] = np.sqrt(tau * v / L)
            iact[r] = tau
        with np.errstate(invalid="ignore"):
            sd[r] = np.sqrt(v)
    out.update(sd=sd, mcse=mcse, iact=iact, var=var, win=win)
    return out
