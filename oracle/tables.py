"""tables.py — CPU ORACLE (test infrastructure) for the whole `mmseq` run: what the
reference's main() computes from a hits file, src/mmseq.cpp:312-1669, restated with
numpy on top of oracle.py.  The Gibbs chain is the shared-Philox replay, so the expected
traces — and therefore every output column — are reproducible on the CPU and can be
compared with the CUDA program cell by cell.

Returned tables are lists of rows of strings, formatted as the reference's
operator<< would (default precision 6 == "%g"; "nan"/"-nan"/"inf"/"-inf"; literal
"NA", "0", "1" where the reference writes literals).  PARITY STATUS: pinned against the outputs of
the reference's own main() (oracle/_ref/mmseq_ref, tests/golden/ref_small/) for every column that does
not depend on the random stream; see mmseq_oracle.cpp."""
import gzip

import numpy as np
from scipy import special

from . import oracle as orc


def g6(x):
    """operator<<(ostream&, double) at default precision."""
    if isinstance(x, str):
        return x
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    x = float(x)
    if np.isnan(x):
        return "nan"
    return "%g" % x


def run(path, alpha=0.1, beta=0.1, max_em_iter=1000, epsilon=0.1, gibbs_iter=16384, seed=1234,
        percentiles=(5.0, 25.0, 50.0, 75.0, 95.0)):
    L = 1024                                   # trace_length, src/mmseq.cpp:191
    stride = gibbs_iter // L                   # :284
    hf = orc.HitsFile(path)
    cl = orc.build_classes(hf)
    n, m, N = cl["n"], cl["m"], cl["N"]
    names = hf.names
    sid_index = cl["sid_index"]
    genes = dict(sorted(hf.genes.items()))     # std::map order
    t2g = {t: g for g, ts in genes.items() for t in ts}
    P = orc.Problem(cl["row_ptr"], cl["col"], cl["k"], cl["len"], alpha=alpha, beta=beta)
    mu0, uh, _ = P.init_mu()
    mu_em, em_iters, loglik, llr = P.em(mu0, max_em_iter, epsilon)
    _, trace = P.gibbs_replay(mu_em, seed, 0, gibbs_iter, stride, L)

    # unique hits of sets (src/uh.cpp) — literal restatement
    def set_csr(sets):
        ptr = [0]; mem = []
        for s in sets:
            mem += [sid_index[t] for t in s if t in sid_index]
            ptr.append(len(mem))
        return np.array(ptr, np.int64), np.array(mem, np.int32)

    iptr, imem = set_csr(hf.identical)
    gptr, gmem = set_csr(genes.values())
    uh_ident = orc.uh_literal(cl["row_ptr"], cl["col"], cl["k"], iptr, imem) if hf.identical else np.zeros(0, np.int32)
    uh_gene = orc.uh_literal(cl["row_ptr"], cl["col"], cl["k"], gptr, gmem)

    # prior-simulated traces for transcripts without hits (:971-978), PRIOR stream keyed by header index
    unobs = [i for i, t in enumerate(names) if t not in sid_index]
    lsc = np.array([hf.efflen[names[i]] * N / 1000000000.0 for i in unobs])
    simu = orc.prior_replay(unobs, lsc, alpha, beta, seed, L) if unobs else np.zeros((0, L))
    simu_of = {names[i]: simu[j] for j, i in enumerate(unobs)}

    ident_trace = np.zeros((len(hf.identical), L))
    for s, grp in enumerate(hf.identical):
        for t in grp:
            if t in sid_index:
                ident_trace[s] += trace[sid_index[t]]
    gene_trace = np.zeros((len(genes), L))
    gene_index = {}
    for g, (gname, ts) in enumerate(genes.items()):
        gene_index[gname] = g
        extra = np.zeros(L)
        for t in ts:
            if t in sid_index:
                gene_trace[g] += trace[sid_index[t]]
            else:
                extra += simu_of[t]
        gene_trace[g] += extra               # the device adds the simulated sum last
    pidx = [int(np.floor(p / 100.0 * (L - 1) + 0.5)) for p in percentiles]   # C round()

    St = orc.summaries_transcripts(trace, percentiles)
    Si = orc.summaries_transcripts(ident_trace, percentiles) if len(hf.identical) else None
    Sg = orc.summaries_transcripts(gene_trace, percentiles)

    def prop_stats(prop, multi):
        with np.errstate(invalid="ignore", over="ignore"):
            mp = np.add.reduce(prop) / L
            if multi:
                z = special.ndtri(np.minimum(np.maximum(prop, 0.000000001), 0.999999999))
                s1 = z.sum(); s2 = (z * z).sum()
                sdv = np.sqrt((s2 - s1 * s1 / L) / (L - 1.0))
                return mp, s1 / L, g6(sdv)
            return mp, np.inf, "-nan"         # inf - inf on x86-64: the default NaN prints as -nan

    digalpha = special.digamma(alpha)
    sqrtpolyg = np.sqrt(special.polygamma(1, alpha))
    hdr_pct = ",".join(g6(p) for p in percentiles)

    # ---- .mmseq (:1469-1554)
    mm = [["# Mapped fragments: %d" % N],
          ["feature_id", "log_mu", "sd", "mcse", "iact", "effective_length", "true_length", "unique_hits", "mean_proportion",
           "mean_probit_proportion", "sd_probit_proportion", "log_mu_em", "observed", "ntranscripts", "percentiles" + hdr_pct,
           "percentiles_proportion" + hdr_pct]]
    for t in names:
        gname = t2g[t]
        ntr = len(genes[gname])
        gtr = gene_trace[gene_index[gname]]
        if t in sid_index:
            c = sid_index[t]
            with np.errstate(invalid="ignore", divide="ignore"):
                prop = trace[c] / gtr
            mp, mpp, sdp = prop_stats(prop, ntr > 1)
            with np.errstate(divide="ignore"):
                lme = np.log(mu_em[c])
            mm.append([t, g6(St["log_mu"][c]), g6(St["sd"][c]), g6(St["mcse"][c]), g6(St["iact"][c]), g6(hf.efflen[t]), g6(hf.truelen[t]),
                       g6(int(uh[c])), g6(mp), g6(mpp), sdp, g6(lme), "1", str(ntr), ",".join(g6(v) for v in St["pct"][c]),
                       ",".join(g6(v) for v in np.sort(prop)[pidx])])
        else:
            sv = simu_of[t]
            prop = sv / gtr
            mp, mpp, sdp = prop_stats(prop, ntr > 1)
            mm.append([t, g6(digalpha - np.log(beta + hf.efflen[t] * N / 1000000000.0)), g6(sqrtpolyg), "0", "1", g6(hf.efflen[t]),
                       g6(hf.truelen[t]), "0", g6(mp), g6(mpp), sdp, "NA", "0", str(ntr), ",".join(g6(v) for v in np.sort(sv)[pidx]),
                       ",".join(g6(v) for v in np.sort(prop)[pidx])])

    # ---- .identical.mmseq (:1556-1613)
    im = [["# Mapped fragments: %d" % N],
          ["feature_id", "log_mu", "sd", "mcse", "iact", "effective_length", "true_length", "unique_hits", "observed", "ntranscripts",
           "percentiles" + hdr_pct]]
    for s, grp in enumerate(hf.identical):
        fid = "".join(t + ("+" if t != grp[-1] else "") for t in grp)
        first = grp[0]
        if np.isfinite(Si["log_mu"][s]):
            im.append([fid, g6(Si["log_mu"][s]), g6(Si["sd"][s]), g6(Si["mcse"][s]), g6(Si["iact"][s]), g6(hf.efflen[first]), g6(hf.truelen[first]),
                       g6(int(uh_ident[s])), "1", str(len(grp)), ",".join(g6(v) for v in Si["pct"][s])])
        else:
            last = grp[-1]
            im.append([fid, g6(np.log(len(grp)) + digalpha - np.log(beta + hf.efflen[last] * N / 1000000000.0)), g6(sqrtpolyg), "0", "NA",
                       g6(hf.efflen[first]), g6(hf.truelen[first]), "0", "0", str(len(grp)), ",".join("NA" for _ in percentiles)])

    # ---- .gene.mmseq (:1615-1669)
    gm = [["# Mapped fragments: %d" % N],
          ["feature_id", "log_mu", "sd", "mcse", "iact", "effective_length", "true_length", "unique_hits", "ntranscripts", "observed",
           "percentiles" + hdr_pct]]
    for g, (gname, ts) in enumerate(genes.items()):
        glen = 0.0
        if np.isfinite(Sg["log_mu"][g]):       # :1380-1392
            num = 0.0; den = 0.0
            for t in ts:
                if t in sid_index:
                    e = np.exp(St["log_mu"][sid_index[t]])
                else:
                    e = np.exp(digalpha - np.log(beta + hf.efflen[t] * N / 1000000000.0))
                num += hf.efflen[t] * e; den += e
            glen = num / den
        obs = any(t in sid_index for t in ts)
        if obs:
            gm.append([gname, g6(Sg["log_mu"][g]), g6(Sg["sd"][g]), g6(Sg["mcse"][g]), g6(Sg["iact"][g]), g6(glen), "NA", g6(int(uh_gene[g])),
                       str(len(ts)), "1", ",".join(g6(v) for v in Sg["pct"][g])])
        else:
            gm.append([gname, g6(Sg["log_mu"][g]), g6(Sg["sd"][g]), g6(Sg["sd"][g] / np.sqrt(L)), "1", g6(glen), "NA", "0", str(len(ts)), "0",
                       ",".join(g6(v) for v in Sg["pct"][g])])

    k_lines = [str(int(v)) for v in cl["k"]]
    M_lines = ["#" + "".join("\t" + nm for nm in cl["names_by_col"])]
    rp, col = cl["row_ptr"], cl["col"]
    for i in range(m):
        for q in range(int(rp[i]), int(rp[i + 1])):
            M_lines.append("%d\t%d" % (i, col[q]))
    with np.errstate(divide="ignore", invalid="ignore"):
        prop_trace = np.array([trace[sid_index[t]] / gene_trace[gene_index[t2g[t]]] for t in cl["names_by_col"]])
    return dict(mmseq=mm, identical=im, gene=gm, k=k_lines, M=M_lines, trace=trace, ident_trace=ident_trace, gene_trace=gene_trace,
                prop_trace=prop_trace, names_by_col=cl["names_by_col"], gene_names=list(genes), em_iters=em_iters, N=N, n=n, m=m,
                identical_ids=["".join(t + ("+" if t != grp[-1] else "") for t in grp) for grp in hf.identical])


def read_table(path):
    rows = []
    for ln in open(path).read().split("\n"):
        if ln == "":
            continue
        rows.append(ln.split("\t") if not ln.startswith("#") else [ln])
    return rows


def read_trace_gz(path):
    lines = gzip.open(path, "rt").read().split("\n")
    ids = lines[0].split(" ")[:-1] if lines[0] else []
    vals = [[float(v) for v in ln.split(" ")[:-1]] for ln in lines[1:] if ln != ""]
    return ids, np.array(vals).T if vals and ids else np.zeros((len(ids), 0))


def cells_match(a, b, rtol=2e-5):
    """Two table cells: literal equality, or (comma lists of) numbers equal to the printed precision."""
    if a == b:
        return True
    pa, pb = a.split(","), b.split(",")
    if len(pa) != len(pb):
        return False
    for x, y in zip(pa, pb):
        if x == y:
            continue
        try:
            fx, fy = float(x), float(y)
        except ValueError:
            return False
        if np.isnan(fx) or np.isnan(fy) or np.isinf(fx) or np.isinf(fy):
            return False
        if not np.isclose(fx, fy, rtol=rtol, atol=1e-300):
            return False
    return True
