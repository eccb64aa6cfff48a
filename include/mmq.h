/* mmq.h — C ABI of libmmseq_b200.so: the B200 (sm_100a) implementation of
 * mmseq's expression-estimation hot path (EM initialisation + Gibbs sampler
 * over hit classes), eturro/mmseq 1.0.11 src/mmseq.cpp:593-918 and the
 * summaries of :938-1395.
 *
 * The reference has no plugin or FFI interface for this path: it is one
 * main() (src/mmseq.cpp:179-1728).  The seams this ABI cuts are therefore the
 * loop nests of that function; every entry point cites the lines it replaces.
 * The host program (`mmseq`, mmseq_b200/csrc/mmseq_main.cpp) keeps the
 * reference's command line and file formats and calls only these functions.
 *
 * Conventions
 *   - plain C: pointers and sizes, no C++ or torch types;
 *   - all array arguments are HOST pointers unless the name ends in _dev;
 *     the library owns every device allocation behind the opaque handle;
 *   - every function returns 0 on success, non-zero on failure;
 *     mmq_last_error() gives the message (the reference prints to cerr and
 *     exit(1)s, e.g. src/mmseq.cpp:278-296, :595-607 — the CLI wrapper does that);
 *   - a handle is used from one host thread at a time; handles are independent;
 *   - results do not depend on the GPU count or launch geometry: every random
 *     draw is a pure function of (seed, stream, class or transcript id, sweep)
 *     through Philox4x32-10 (include/mmq_sampler.h).
 * There is no CPU fallback: without a CUDA device mmq_create fails.
 */
#ifndef MMQ_H
#define MMQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMQ_OK 0
#define MMQ_ERR_ARG 1
#define MMQ_ERR_CUDA 2
#define MMQ_ERR_STATE 3
#define MMQ_ERR_NCCL 4

typedef struct mmq_handle mmq_handle;

/* The hit-class matrix M of src/mmseq.cpp:117 (boolean, m x n, row-major
 * compressed) in CSR form, with the per-class fragment counts k (:385, :440)
 * and the scaled lengths l (:593-608).  One mmq_problem is one GPU's shard:
 * a contiguous block of classes (rows) and ALL n transcripts (columns). */
typedef struct mmq_problem {
  int64_t n;              /* transcripts with at least one hit (columns), :388 */
  int64_t m;              /* hit classes in this shard (rows), :387            */
  int64_t nnz;            /* row_ptr[m]                                        */
  const int64_t* row_ptr; /* [m+1]                                             */
  const int32_t* col;     /* [nnz] ascending within a row (:412 sort(comb))    */
  const int32_t* k;       /* [m] fragments per class, or NULL meaning all 1
                             (the per-fragment layout)                         */
  const float* weight;    /* [nnz] per-hit weights or NULL.  NULL = reference
                             semantics (M is boolean, :72, :437-438).          */
  const double* len;      /* [n] l[t] = effective_length[t]*N/1e9, :603        */
  double alpha, beta;     /* Gamma prior on mu, :184-185                       */
  int64_t class_id_base;  /* global index of this shard's row 0: the Philox
                             counter of row i is class_id_base + i             */
  const int64_t* class_id; /* [m] explicit Philox counters, or NULL.  Lets the caller
                             hand the classes over in any order (e.g. sorted by cost so
                             that a warp's classes do similar work) and still get the
                             chain of the canonical order.  k != NULL only.       */
} mmq_problem;

/* Upload a shard (H2D) and validate it on the device.  The transcript-major transpose
 * (replaces Mt = trans(M), src/mmseq.cpp:578, and Mrowsum/Mcolsum :582-589) is built on the
 * device by the first call that needs it (mmq_init_mu, mmq_em, a TRANSPOSED sweep). */
int mmq_create(const mmq_problem* problem, int device, mmq_handle** out);
void mmq_destroy(mmq_handle* h);
/* Message of the last failure on h (h may be NULL: failure of mmq_create). */
const char* mmq_last_error(const mmq_handle* h);
/* Run every later call on this cudaStream_t (default: a stream owned by h). */
int mmq_set_stream(mmq_handle* h, void* cuda_stream);
void* mmq_get_stream(mmq_handle* h);
int mmq_synchronize(mmq_handle* h);
/* Device memory held by the handle, bytes. */
int64_t mmq_device_bytes(const mmq_handle* h);

/* Multi-GPU: one process (or thread) per GPU, each with its own shard.  Rank 0
 * calls mmq_comm_id and ships the 128 bytes to the others (torch.distributed,
 * MPI, a file — the caller's plumbing); everyone calls mmq_comm_init.  After
 * that mmq_init_mu / mmq_em / mmq_gibbs end each pass with one NCCL all-reduce
 * over NVLink of the per-transcript vector (int32 counts for Gibbs — the
 * reference's thread-partials sum, src/mmseq.cpp:895-899 — fp64 partial sums
 * for EM).  libnccl.so.2 is dlopen'ed on first use. */
int mmq_comm_id(char id[128]);
int mmq_comm_init(mmq_handle* h, const char id[128], int rank, int nranks);
/* Hand the communicator of `from` (which keeps none) to `to`: lets a long-lived process load
 * the next sample's shard without paying NCCL initialisation again.  A peer-memory attachment
 * (mmq_p2p_attach*) moves along when both handles have the same n; every rank must then make the
 * same move between the same two sweeps. */
int mmq_comm_move(mmq_handle* from, mmq_handle* to);

/* Fused count exchange over NVLink peer memory (optional, on top of the communicator).
 * After mmq_p2p_attach the Gibbs sweep makes no NCCL call: the Gamma kernel of every rank
 * signals "my allocation is done" into its peers' flag words, waits for theirs, then reads the
 * count vectors of ALL ranks straight from peer memory (P2P loads), sums them and draws —
 * the all-reduce (src/mmseq.cpp:895-899) and the Gamma update (:904-908) are ONE kernel.
 * Counts are double-buffered by sweep parity so no second barrier is needed.
 *   multi-process: every rank calls mmq_p2p_export (64-byte cudaIpcMemHandle), the caller
 *     gathers the handles of all ranks (rank order) and passes them to mmq_p2p_attach;
 *   single process, one handle per GPU: mmq_p2p_attach_local with the handles in rank order.
 * All ranks must then issue the same sequence of sweeps (they already must, for NCCL). */
int mmq_p2p_export(mmq_handle* h, char ipc_handle[64]);
int mmq_p2p_attach(mmq_handle* h, const char* ipc_handles, int rank, int nranks);
int mmq_p2p_attach_local(mmq_handle** handles, int nranks);
/* Ranks attached through mmq_p2p_attach*(), 0 when the handle exchanges counts through NCCL (or is alone). */
int mmq_p2p_attached(const mmq_handle* h);

/* mu0[t] = (sum_{i containing t} k[i]/|i|)/l[t] and unique_hits[t] =
 * counts_shared[t][0]; src/mmseq.cpp:617-638.  Leaves mu0 as the current mu.
 * unique_hits_out (int32[n]) may be NULL. */
int mmq_init_mu(mmq_handle* h, int32_t* unique_hits_out);
int mmq_set_mu(mmq_handle* h, const double* mu);
int mmq_get_mu(mmq_handle* h, double* mu_out);

/* sum_i k[i] log(sum_{s in i} mu[s]) - sum_t mu[t] l[t]; src/mmseq.cpp:745-754. */
int mmq_loglik(mmq_handle* h, double* loglik_out);

/* EM from the current mu; src/mmseq.cpp:756-811: llr starts at eps+1, loop
 * while iter < max_iter && llr > eps.  Outputs may be NULL. */
int mmq_em(mmq_handle* h, int max_iter, double eps, int* iters_out, double* loglik_out,
           double* llr_out);

/* Gibbs sweeps first_sweep .. first_sweep+n_sweeps-1 from the current mu;
 * src/mmseq.cpp:851-918.  Sweep s with s % stride == 0 stores mu into trace
 * slot s/stride (slots >= trace_len are dropped), :911-917.  Asynchronous on
 * the handle's stream; mmq_synchronize / mmq_get_* wait for it.
 * flags: MMQ_GIBBS_* below. */
#define MMQ_GIBBS_DEFAULT 0
#define MMQ_GIBBS_TRANSPOSED 1 /* materialise X, reduce over the transposed CSR
                                  (atomic-free); default is the fused path   */
#define MMQ_GIBBS_NO_GRAPH 2   /* plain launches instead of a CUDA graph      */
#define MMQ_GIBBS_TIME_KERNELS 4 /* bracket every k_alloc / k_gamma launch with
                                  CUDA events on the handle's stream; read the
                                  totals with mmq_kernel_times               */
#define MMQ_GIBBS_GENERIC_KERNEL 8 /* k == 1 shards: use the general multinomial
                                  kernel instead of the categorical fast path
                                  (same results; for tests and comparison)   */
#define MMQ_GIBBS_RAGGED_KERNEL 16 /* k == 1 shards: use the row-pointer driven (TMA-staged)
                                  kernel even when the by-length segment plan exists */
#define MMQ_GIBBS_ROWS_KERNEL 64 /* by-length k == 1 shards: the row plan (columns once per run of identical rows, member-major
                                  weights: 1.6x fewer bytes than the default segment kernel, measured no faster); same results */
int mmq_gibbs(mmq_handle* h, uint32_t seed, int64_t first_sweep, int64_t n_sweeps, int stride,
              int trace_len, int flags);

/* What the class plan of a collapsed shard (classes with counts k; built by mmq_create) holds:
 * out[0] 1 if the plan is in use, out[1] classes of the small set (k <= mmq_cat_limit(d) fragments, at most
 * 64 members: categorical draws, stands in for the multinomial of src/mmseq.cpp:880 for those classes), out[2] packed column
 * slots and out[3] class slots streamed per sweep for them, out[4] classes and out[5] CSR entries
 * left to the general kernel (more than 64 members), out[6] classes of the chain set (more than
 * mmq_cat_limit(d) fragments: gsl_ran_multinomial's conditional-binomial chain, one class per lane) and out[7]
 * their packed column slots.  Used by bench.py for the algorithmic-bytes figure. */
int mmq_cls_stats(const mmq_handle* h, int64_t out[8]);

/* What the row plan of a by-length k == 1 shard holds (mmq_rows.cu; built by this call if no sweep has asked for it yet):
 * out[0] 1 if in use, out[1] rows with >= 2 members, out[2] runs of identical rows ("sets"), out[3] set
 * column entries, out[4] weight slots (0 without weights), out[5] chunks of 128 rows, out[6] bytes streamed
 * per sweep (4 B per weight slot and set column, 32 B per chunk), out[7] single-member rows (not visited). */
int mmq_rows_stats(mmq_handle* h, int64_t out[8]);

/* Launch-geometry knobs of the class-plan sweep, for measurements (tools/gpu_tune_cls.py); value 0 restores the
 * default.  knob 0: bit mask of pieces NOT launched (1 k >= 2 small classes, 2 rest, 4 k >= 2 large classes, 8 chain,
 * 16 single-fragment classes: timing experiments only, the chain is then wrong); knobs 1..4: resident CTAs per SM the
 * grid of the chain / large / small / single-fragment kernel is capped at; knob 5: launch order variant; knob 6: 1 =
 * the multi-GPU count exchange as reduce-scatter + all-gather (k_gamma_rs) instead of the all-reduce inside the Gamma
 * kernel (all ranks must set it alike).  Results do not depend on knobs 1..6. */
int mmq_tune(mmq_handle* h, int knob, int value);

/* Device time of the launches made under MMQ_GIBBS_TIME_KERNELS since the last
 * call (waits for the stream): total milliseconds and launch count of the
 * allocation kernel and of the Gamma kernel.  Any output may be NULL. */
int mmq_kernel_times(mmq_handle* h, double* alloc_ms, int64_t* alloc_launches, double* gamma_ms,
                     int64_t* gamma_launches);

/* One sweep with its integer state exposed, for bit-exact parity against the
 * CPU replay: x_out int32[nnz] (CSR order; the X matrix of :842-847, :884),
 * counts_out int32[n] (Xcolsum, :895-899; the all-reduced vector when a
 * communicator is attached), mu_out fp64[n] after the Gamma step.  Any may be
 * NULL.  flags as mmq_gibbs (TRANSPOSED or fused); both give the same integers. */
int mmq_sweep_debug(mmq_handle* h, uint32_t seed, int64_t sweep, int flags, int32_t* x_out,
                    int32_t* counts_out, double* mu_out);

/* trace_out[t*trace_len + s] = mu[t] at slot s — the layout of mu_trace,
 * src/mmseq.cpp:827, :914. */
int mmq_get_trace(mmq_handle* h, double* trace_out);
int mmq_trace_len(const mmq_handle* h);

/* ---- posterior summaries on the device (src/mmseq.cpp:938-1363) ---------- */

/* Groups of transcripts whose traces are summed: identical-transcript sets
 * (:938-954) and genes (:959-982).  member ids index the n observed
 * transcripts; `extra` (fp64[ngroups*trace_len], may be NULL) is added to the
 * group trace — the prior-simulated Gamma draws of hit-less isoforms, :971-978. */
#define MMQ_GROUP_IDENTICAL 0
#define MMQ_GROUP_GENE 1
int mmq_set_groups(mmq_handle* h, int kind, int64_t ngroups, const int64_t* group_ptr,
                   const int32_t* members, const double* extra);

/* Per-feature summaries of a trace matrix.  which: 0 transcripts (n rows),
 * 1 identical sets, 2 genes.  All outputs fp64[rows] unless noted; NULL skips.
 *   log_mean   mean of log(trace)                                  :1203-1227
 *   var,tau,win  sokal() on the log trace (win int32)        :1308-1363, sokal.cc:33-87
 *   pct        fp64[rows*npct], sorted raw trace at index pct_idx[j]   :1111-1172 */
int mmq_summarize(mmq_handle* h, int which, double* log_mean, double* var, double* tau,
                  int32_t* win, int32_t* sokal_status, int npct, const int32_t* pct_idx, double* pct);
/* Group traces back to the host, [g*trace_len + s]; which = 1 or 2 as above. */
int mmq_get_group_trace(mmq_handle* h, int which, double* out);

/* Proportion summaries of the observed transcripts (:985-1008, :1236-1265):
 * prop = mu_t / gene trace; mean_prop; sum and sum of squares of
 * Phi^-1(clamp(prop, 1e-9, 1-1e-9)) (callers finish :1261-1265 on the host);
 * pct of the sorted proportion trace.  gene_of[t] = gene index of transcript t
 * (int32[n]); multi_iso[t] != 0 when the gene has more than one isoform. */
int mmq_prop_summaries(mmq_handle* h, const int32_t* gene_of, const uint8_t* multi_iso,
                       double* mean_prop, double* sum_probit, double* sumsq_probit, int npct,
                       const int32_t* pct_idx, double* pct, double* prop_trace_out);

/* uh(): unique hits of transcript sets, src/uh.cpp:3-26 — sum of k[i] over the
 * classes all of whose members lie in one set.  set_of int32[n]: the set of
 * each transcript or -1.  out int32[nsets] (all-reduced over shards). */
int mmq_unique_hits_sets(mmq_handle* h, const int32_t* set_of, int64_t nsets, int32_t* out);

/* Prior draws for transcripts without hits, src/mmseq.cpp:971-978:
 * out[u*trace_len + s] ~ Gamma(alpha, rate[u]) with rate[u] = beta + len*N/1e9, from the
 * PRIOR Philox stream keyed by (seed, ids[u], s) — the reference draws them from rg[0]. */
int mmq_prior_draws(int device, int64_t count, const int64_t* ids, const double* rate, double alpha,
                    uint32_t seed, int trace_len, double* out);

/* Stand-alone batched Sokal on host data: rows x len fp64 (len a power of two,
 * 4..2048); src/sokal.cc:33-87. */
int mmq_sokal_batch(int device, int64_t rows, int len, const double* x, double* var, double* tau,
                    int32_t* win, int32_t* status);

/* ---- the consumer of the traces: mmcollapse's covariance step (src/mmcollapse.cpp:483-561) ----------
 * The one dense contraction of the package: tensor cores (tcgen05, bf16 split products, fp32 accumulation in
 * TMEM), everything around it fp64.  nsplit = bf16 terms per value: 1 (|dr| <~ 4e-3 in correlation units),
 * 2 (<~ 2e-5, the default of the callers here), 3 (<~ 2e-6). */

/* get_corrs()'s per-sample step, src/mmcollapse.cpp:553-558: R = cov(M) of the L x C trace matrix (column c =
 * the L posterior draws of feature c, contiguous: Armadillo's layout of myM and the layout of mmq_get_trace),
 * non-finite -> 0.  R: C x C fp64, both triangles, exactly symmetric, diagonal = the columns' variances.
 * L a multiple of 64 (mmcollapse: TRACELEN 1024).  Host pointers; H2D, kernels and D2H inside. */
int mmq_trace_cov(int device, const double* M, int L, int64_t C, int nsplit, double* R);
/* Same on device-resident data, asynchronous on cuda_stream, no allocation inside: workspace_dev must hold
 * mmq_trace_cov_workspace_bytes(L, C, nsplit) bytes (256-byte aligned). */
int64_t mmq_trace_cov_workspace_bytes(int L, int64_t C, int nsplit);
int mmq_trace_cov_dev(const double* M_dev, int L, int64_t C, int nsplit, double* R_dev, void* workspace_dev, void* cuda_stream);
/* Straight from the trace a handle recorded (mmq_gibbs with trace_len = L): features[c] = index of an observed
 * transcript; no trace file round trip (the reference re-reads *.trace_gibbs.gz, src/mmcollapse.cpp:530-551).
 * R_out (host) and / or R_dev_out (device, C x C fp64) receive the matrix; either may be NULL. */
int mmq_handle_trace_cov(mmq_handle* h, const int32_t* features, int64_t C, int nsplit, double* R_out, double* R_dev_out);

/* mean_corrs(), src/mmcollapse.cpp:483-511.  R: C x C x ns cube (slice s at R + s C C, symmetric slices as
 * mmq_trace_cov writes them), S: C x ns column-major, 1 where feature c was observed in sample s (:686-695);
 * for every t in ts[0..nts) and v in 0..C-1: V(t,v) = V(v,t) = mean over the samples with S(t,s) S(v,s) = 1 of
 * R(t,v,s)/sqrt(R(t,t,s))/sqrt(R(v,v,s)) + sdpenalty * W(t,v), W = their sd (0 for one sample or if not finite).
 * V, W: C x C fp64, read and written (entries outside the rows / columns of ts stay). */
int mmq_mean_corrs(int device, const double* R, const uint8_t* S, int64_t C, int ns, const int32_t* ts, int64_t nts, double sdpenalty,
                   double* V, double* W);
int mmq_mean_corrs_dev(const double* R_dev, const uint8_t* S_dev, int64_t C, int ns, const int32_t* ts_dev, int64_t nts, double sdpenalty,
                       double* V_dev, double* W_dev, void* cuda_stream);

/* Device memory given back by mmq_destroy and by the set-up steps' temporaries is kept in a per-device cache and handed
 * out again (cudaMalloc / cudaFree of these sizes cost 1-100 ms each and made mmq_create vary between 8 ms and 2 s; a
 * process that runs one sample after another — config 5 — now pays them once).  At most MMQ_DEVICE_CACHE_MB (environment,
 * default 16384; 0 disables) are held per device; this call returns them to the driver (device < 0: every device). */
int mmq_release_cache(int device);

/* Kernel launches issued by this process through the library so far. */
int64_t mmq_launch_count(void);
/* Bring up the CUDA context of `device` (the first CUDA call of a process costs 1-2 s): a host
 * program calls this from a side thread while it parses its input.  No reference counterpart. */
int mmq_warmup(int device);

const char* mmq_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MMQ_H */
