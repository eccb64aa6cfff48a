/* mmq_post.cu — posterior summaries on the device (sm_100a): group (gene /
 * identical-set) traces, log-scale means, batched Sokal IACT, percentiles,
 * isoform-proportion and probit moments, unique hits of transcript sets.
 *
 * Reference code replaced (eturro/mmseq 1.0.11, /root/reference):
 *   k_group_trace      src/mmseq.cpp:938-982   identical-set and gene trace sums
 *   k_row_sokal        src/mmseq.cpp:1203-1227 (log + mean), :1308-1363 and
 *                      src/sokal.cc:33-87      (var, tau, window)
 *   k_row_percentiles  src/mmseq.cpp:1111-1172 (sort each trace, pick indices)
 *   k_prop             src/mmseq.cpp:985-1008, :1236-1257
 *   k_uh_sets          src/uh.cpp:3-26 restated O(nnz)
 * All traces are [row * trace_len + slot] (the reference's mu_trace layout).
 */
#include <algorithm>
#include <cstdio>

#include "../../include/mmq_sampler.h"
#include "mmq_internal.h"

#define MMQ_POST_THREADS 256
#define MMQ_MAX_ROW_LEN 2048

__device__ __forceinline__ double post_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

/* deterministic block sum, result broadcast to every thread */
__device__ __forceinline__ double post_block_sum(double v) {
  __shared__ double s_part[33];
  v = post_warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) r += s_part[i];
    s_part[32] = r;
  }
  __syncthreads();
  return s_part[32];
}

/* group_trace[g*L+s] = sum_{t in group g} trace[t*L+s] (+ extra[g*L+s]) */
__global__ void k_group_trace(const double* __restrict__ trace, const int64_t* __restrict__ gptr,
                              const int32_t* __restrict__ members, const double* __restrict__ extra,
                              double* __restrict__ out, int64_t ngroups, int L) {
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int64_t b = gptr[g], e = gptr[g + 1];
    for (int s = threadIdx.x; s < L; s += blockDim.x) {
      double v = 0.0;
      for (int64_t q = b; q < e; ++q) v += trace[(int64_t)members[q] * L + s];
      if (extra) v += extra[g * L + s];
      out[g * L + s] = v;
    }
  }
}

/* In-place forward FFT of (re, im) in shared memory, length len = 2^lg.
 * dif = true : natural order in, bit-reversed order out (Gentleman-Sande)
 * dif = false: bit-reversed order in, natural order out (Cooley-Tukey) */
__device__ void smem_fft(double* re, double* im, int len, bool dif) {
  const int nb = len >> 1;
  if (dif) {
    for (int half = nb; half >= 1; half >>= 1) {
      for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int j = b & (half - 1);
        const int a = ((b - j) << 1) + j;
        const int c = a + half;
        double sn, cs;
        sincospi(-(double)j / (double)half, &sn, &cs);
        const double xr = re[a] - re[c], xi = im[a] - im[c];
        re[a] += re[c]; im[a] += im[c];
        re[c] = xr * cs - xi * sn;
        im[c] = xr * sn + xi * cs;
      }
      __syncthreads();
    }
  } else {
    for (int half = 1; half <= nb; half <<= 1) {
      for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int j = b & (half - 1);
        const int a = ((b - j) << 1) + j;
        const int c = a + half;
        double sn, cs;
        sincospi(-(double)j / (double)half, &sn, &cs);
        const double xr = re[c] * cs - im[c] * sn, xi = re[c] * sn + im[c] * cs;
        re[c] = re[a] - xr; im[c] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
      __syncthreads();
    }
  }
}

/* One CTA per row: optional log, mean, then Sokal's estimator (src/sokal.cc:33-87):
 * FFT -> |.|^2 -> zero DC -> FFT -> var = acov0/(n(n-1)) -> rho -> adaptive window. */
__global__ void __launch_bounds__(MMQ_POST_THREADS)
k_row_sokal(const double* __restrict__ src, int64_t rows, int len, int do_log, double* __restrict__ mean_out,
            double* __restrict__ var_out, double* __restrict__ tau_out, int32_t* __restrict__ win_out,
            int32_t* __restrict__ status_out) {
  extern __shared__ double sh[];
  double* re = sh;
  double* im = sh + len;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    double part = 0.0;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      double v = src[row * len + i];
      if (do_log) v = log(v);
      re[i] = v;
      im[i] = 0.0;
      part += v;
    }
    const double total = post_block_sum(part); /* also orders the smem writes */
    if (threadIdx.x == 0 && mean_out) mean_out[row] = total / (double)len;
    smem_fft(re, im, len, true);
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      const double p = re[i] * re[i] + im[i] * im[i];
      re[i] = (i == 0) ? 0.0 : p; /* bit-reversal maps bin 0 to slot 0 */
      im[i] = 0.0;
    }
    __syncthreads();
    smem_fft(re, im, len, false);
    if (threadIdx.x == 0) {
      const double x0 = re[0];
      const double var = x0 / ((double)len * (double)(len - 1));
      const double c = 1.0 / x0;
      double sum = -0.333333333333333333333;
      int m = len + 1;
      for (int i = 0; i < len; ++i) {
        sum += re[i] * c - 0.166666666666666666666;
        if (sum < 0) { m = i + 1; break; }
      }
      if (var_out) var_out[row] = var;
      if (tau_out) tau_out[row] = 2 * (sum + (m - 1.0) / 6.0);
      if (win_out) win_out[row] = m;
      if (status_out) status_out[row] = 0;
    }
    __syncthreads();
  }
}

/* One CTA per row: bitonic sort of the row in shared memory (len a power of two),
 * then pct[row*npct + j] = sorted[idx[j]].  src/mmseq.cpp:1139-1146. */
__global__ void __launch_bounds__(MMQ_POST_THREADS)
k_row_percentiles(const double* __restrict__ src, int64_t rows, int len, int npct,
                  const int32_t* __restrict__ idx, double* __restrict__ pct) {
  extern __shared__ double sh[];
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    for (int i = threadIdx.x; i < len; i += blockDim.x) sh[i] = src[row * len + i];
    __syncthreads();
    for (int k = 2; k <= len; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
          const int p = i ^ j;
          if (p > i) {
            const double a = sh[i], b = sh[p];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { sh[i] = b; sh[p] = a; }
          }
        }
        __syncthreads();
      }
    for (int j = threadIdx.x; j < npct; j += blockDim.x) pct[row * npct + j] = sh[idx[j]];
    __syncthreads();
  }
}

/* prop[s] = trace[t][s] / gene_trace[gene_of[t]][s]  (src/mmseq.cpp:995-996);
 * mean (:1244, :1258), probit sum / sum of squares (:1249-1255). */
__global__ void __launch_bounds__(MMQ_POST_THREADS)
k_prop(const double* __restrict__ trace, const double* __restrict__ gene_trace, const int32_t* __restrict__ gene_of,
       const uint8_t* __restrict__ multi_iso, int64_t n, int L, double* __restrict__ prop_trace,
       double* __restrict__ mean_prop, double* __restrict__ sum_probit, double* __restrict__ sumsq_probit) {
  for (int64_t t = blockIdx.x; t < n; t += gridDim.x) {
    const int64_t g = gene_of[t];
    const bool multi = multi_iso[t] != 0;
    double sp = 0.0, sz = 0.0, szz = 0.0;
    for (int s = threadIdx.x; s < L; s += blockDim.x) {
      const double p = trace[t * L + s] / gene_trace[g * L + s];
      prop_trace[t * L + s] = p;
      sp += p;
      double z;
      if (multi) {
        double pc = p; /* min(max(p,1e-9),1-1e-9); a NaN passes through both, as with std::max/min */
        if (pc < 0.000000001) pc = 0.000000001;
        if (pc > 0.999999999) pc = 0.999999999;
        z = mmq_ndtri(pc);
      } else {
        z = mmq_u2d(0x7ff0000000000000ull);
      }
      sz += z;
      szz += z * z;
    }
    sp = post_block_sum(sp);
    sz = post_block_sum(sz);
    szz = post_block_sum(szz);
    if (threadIdx.x == 0) {
      mean_prop[t] = sp / (double)L;
      sum_probit[t] = sz;
      sumsq_probit[t] = szz;
    }
  }
}

/* uh(): a class counts for set s when all its members lie in s. src/uh.cpp:12-23. */
template <bool HAS_K>
__global__ void k_uh_sets(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                          const int32_t* __restrict__ kk, const int32_t* __restrict__ set_of, int64_t m,
                          int32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = row_ptr[i], e = row_ptr[i + 1];
    const int32_t s0 = set_of[col[b]];
    bool uniq = s0 >= 0;
    for (int64_t q = b + 1; q < e && uniq; ++q) uniq = set_of[col[q]] == s0;
    if (uniq) atomicAdd(out + s0, HAS_K ? kk[i] : 1);
  }
}

/* Gamma(alpha, rate_u) prior draws, one thread per (transcript, slot). src/mmseq.cpp:971-978. */
__global__ void k_prior(const int64_t* __restrict__ ids, const double* __restrict__ rate, int64_t count, int L,
                        double alpha, uint32_t seed, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count * L; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = i / L;
    const int s = (int)(i - u * L);
    mmq_rng g;
    mmq_rng_init(&g, seed, MMQ_STREAM_PRIOR, (uint64_t)ids[u], (uint32_t)s);
    out[i] = mmq_gamma(&g, alpha, rate[u]);
  }
}

/* ------------------------------------------------------------------ host */

static bool pow2_ok(int len) { return len >= 4 && len <= MMQ_MAX_ROW_LEN && (len & (len - 1)) == 0; }

int mmq_ensure_trace_groups(mmq_handle* h) {
  for (int kind = 0; kind < 2; ++kind) {
    mmq_group_set& g = h->groups[kind];
    if (g.ngroups == 0 || g.trace_valid) continue;
    if (!h->trace) return mmq_fail(h, MMQ_ERR_STATE, "no trace recorded yet");
    if (!g.trace_dev) {
      int rc = mmq_dev_alloc(h, (void**)&g.trace_dev, sizeof(double) * (size_t)g.ngroups * (size_t)h->trace_len);
      if (rc) return rc;
    }
    const int grid = (int)std::min<int64_t>(g.ngroups, (int64_t)h->num_sms * 16);
    k_group_trace<<<grid, 128, 0, h->stream>>>(h->trace, g.ptr_dev, g.members_dev, g.extra_dev, g.trace_dev, g.ngroups, h->trace_len);
    MMQ_LAUNCHED(h);
    g.trace_valid = true;
  }
  return MMQ_OK;
}

static int run_row_summaries(mmq_handle* h, cudaStream_t st, const double* src_dev, int64_t rows, int len, int do_log,
                             double* log_mean, double* var, double* tau, int32_t* win, int32_t* status, int npct,
                             const int32_t* pct_idx, double* pct) {
  if (rows <= 0) return MMQ_OK;
  double *d_mean = nullptr, *d_var = nullptr, *d_tau = nullptr, *d_pct = nullptr;
  int32_t *d_win = nullptr, *d_status = nullptr, *d_idx = nullptr;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t b) { if (e == cudaSuccess) e = cudaMalloc(p, b ? b : 16); };
  const bool want_sokal = log_mean || var || tau || win || status;
  int ret = MMQ_OK;
  if (want_sokal) {
    A((void**)&d_mean, sizeof(double) * rows); A((void**)&d_var, sizeof(double) * rows); A((void**)&d_tau, sizeof(double) * rows);
    A((void**)&d_win, sizeof(int32_t) * rows); A((void**)&d_status, sizeof(int32_t) * rows);
  }
  if (npct > 0 && pct) { A((void**)&d_pct, sizeof(double) * rows * npct); A((void**)&d_idx, sizeof(int32_t) * npct); }
  if (e == cudaSuccess && want_sokal) {
    const int grid = (int)std::min<int64_t>(rows, 148 * 8);
    k_row_sokal<<<grid, MMQ_POST_THREADS, sizeof(double) * 2 * len, st>>>(src_dev, rows, len, do_log, d_mean, d_var, d_tau, d_win, d_status);
    g_mmq_launches.fetch_add(1);
    e = cudaGetLastError();
    auto D = [&](void* dst, const void* src, size_t b) { if (e == cudaSuccess && dst) e = cudaMemcpyAsync(dst, src, b, cudaMemcpyDeviceToHost, st); };
    D(log_mean, d_mean, sizeof(double) * rows); D(var, d_var, sizeof(double) * rows); D(tau, d_tau, sizeof(double) * rows);
    D(win, d_win, sizeof(int32_t) * rows); D(status, d_status, sizeof(int32_t) * rows);
  }
  if (e == cudaSuccess && npct > 0 && pct) {
    e = cudaMemcpyAsync(d_idx, pct_idx, sizeof(int32_t) * npct, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
      const int grid = (int)std::min<int64_t>(rows, 148 * 8);
      k_row_percentiles<<<grid, MMQ_POST_THREADS, sizeof(double) * len, st>>>(src_dev, rows, len, npct, d_idx, d_pct);
      g_mmq_launches.fetch_add(1);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(pct, d_pct, sizeof(double) * rows * npct, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) ret = mmq_cuda_fail(h, e, "row summaries", __FILE__, __LINE__);
  cudaFree(d_mean); cudaFree(d_var); cudaFree(d_tau); cudaFree(d_win); cudaFree(d_status); cudaFree(d_pct); cudaFree(d_idx);
  return ret;
}

extern "C" {

int mmq_set_groups(mmq_handle* h, int kind, int64_t ngroups, const int64_t* group_ptr, const int32_t* members, const double* extra) {
  if (!h || kind < 0 || kind > 1 || ngroups < 0 || (ngroups > 0 && (!group_ptr || group_ptr[0] != 0)))
    return mmq_fail(h, MMQ_ERR_ARG, "mmq_set_groups: bad arguments");
  MMQ_CUDA(h, cudaSetDevice(h->device));
  mmq_group_set& g = h->groups[kind];
  if (g.ptr_dev) { mmq_dev_free(h, g.ptr_dev); g.ptr_dev = nullptr; }
  if (g.members_dev) { mmq_dev_free(h, g.members_dev); g.members_dev = nullptr; }
  if (g.extra_dev) { mmq_dev_free(h, g.extra_dev); g.extra_dev = nullptr; }
  if (g.trace_dev) { mmq_dev_free(h, g.trace_dev); g.trace_dev = nullptr; }
  g.ngroups = ngroups;
  g.trace_valid = false;
  if (ngroups == 0) return MMQ_OK;
  const int64_t nm = group_ptr[ngroups];
  for (int64_t q = 0; q < nm; ++q)
    if (members[q] < 0 || members[q] >= h->n) return mmq_fail(h, MMQ_ERR_ARG, "mmq_set_groups: member index out of range");
  int rc;
  if ((rc = mmq_dev_alloc(h, (void**)&g.ptr_dev, sizeof(int64_t) * (size_t)(ngroups + 1)))) return rc;
  if ((rc = mmq_dev_alloc(h, (void**)&g.members_dev, sizeof(int32_t) * (size_t)std::max<int64_t>(nm, 1)))) return rc;
  MMQ_CUDA(h, cudaMemcpyAsync(g.ptr_dev, group_ptr, sizeof(int64_t) * (size_t)(ngroups + 1), cudaMemcpyHostToDevice, h->stream));
  if (nm) MMQ_CUDA(h, cudaMemcpyAsync(g.members_dev, members, sizeof(int32_t) * (size_t)nm, cudaMemcpyHostToDevice, h->stream));
  if (extra) {
    if (h->trace_len <= 0) return mmq_fail(h, MMQ_ERR_STATE, "mmq_set_groups: extra given before any trace exists");
    const size_t b = sizeof(double) * (size_t)ngroups * (size_t)h->trace_len;
    if ((rc = mmq_dev_alloc(h, (void**)&g.extra_dev, b))) return rc;
    MMQ_CUDA(h, cudaMemcpyAsync(g.extra_dev, extra, b, cudaMemcpyHostToDevice, h->stream));
  }
  MMQ_CUDA(h, cudaStreamSynchronize(h->stream));
  return MMQ_OK;
}

int mmq_summarize(mmq_handle* h, int which, double* log_mean, double* var, double* tau, int32_t* win,
                  int32_t* sokal_status, int npct, const int32_t* pct_idx, double* pct) {
  if (!h || which < 0 || which > 2) return mmq_fail(h, MMQ_ERR_ARG, "mmq_summarize: bad arguments");
  if (!h->trace) return mmq_fail(h, MMQ_ERR_STATE, "mmq_summarize: no trace recorded");
  MMQ_CUDA(h, cudaSetDevice(h->device));
  const int L = h->trace_len;
  for (int j = 0; j < npct; ++j)
    if (pct_idx[j] < 0 || pct_idx[j] >= L) return mmq_fail(h, MMQ_ERR_ARG, "mmq_summarize: percentile index out of range");
  int rc = mmq_ensure_trace_groups(h);
  if (rc) return rc;
  const double* src = which == 0 ? h->trace : h->groups[which - 1].trace_dev;
  const int64_t rows = which == 0 ? h->n : h->groups[which - 1].ngroups;
  if (rows == 0) return MMQ_OK;
  if (!pow2_ok(L)) {
    /* sokal() fails for such lengths (src/sokal.cc:36-40, :108-126); callers map
     * failure to mcse = trace_length, iact = NaN (src/mmseq.cpp:1316-1318).
     * Percentiles and log means are not produced either. */
    if (sokal_status) for (int64_t r = 0; r < rows; ++r) sokal_status[r] = L < 4 ? 200 : (L > MMQ_MAX_ROW_LEN ? 100 : 201);
    return mmq_fail(h, MMQ_ERR_ARG, "mmq_summarize: trace_len must be a power of two in [4, 2048]");
  }
  return run_row_summaries(h, h->stream, src, rows, L, 1, log_mean, var, tau, win, sokal_status, npct, pct_idx, pct);
}

int mmq_get_group_trace(mmq_handle* h, int which, double* out) {
  if (!h || !out || which < 1 || which > 2) return mmq_fail(h, MMQ_ERR_ARG, "mmq_get_group_trace: bad arguments");
  MMQ_CUDA(h, cudaSetDevice(h->device));
  int rc = mmq_ensure_trace_groups(h);
  if (rc) return rc;
  mmq_group_set& g = h->groups[which - 1];
  if (g.ngroups == 0) return MMQ_OK;
  MMQ_CUDA(h, cudaMemcpyAsync(out, g.trace_dev, sizeof(double) * (size_t)g.ngroups * (size_t)h->trace_len, cudaMemcpyDeviceToHost, h->stream));
  MMQ_CUDA(h, cudaStreamSynchronize(h->stream));
  return MMQ_OK;
}

int mmq_prop_summaries(mmq_handle* h, const int32_t* gene_of, const uint8_t* multi_iso, double* mean_prop,
                       double* sum_probit, double* sumsq_probit, int npct, const int32_t* pct_idx, double* pct,
                       double* prop_trace_out) {
  if (!h || !gene_of || !multi_iso) return mmq_fail(h, MMQ_ERR_ARG, "mmq_prop_summaries: NULL argument");
  if (!h->trace) return mmq_fail(h, MMQ_ERR_STATE, "mmq_prop_summaries: no trace recorded");
  mmq_group_set& G = h->groups[MMQ_GROUP_GENE];
  if (G.ngroups == 0) return mmq_fail(h, MMQ_ERR_STATE, "mmq_prop_summaries: gene groups not set");
  MMQ_CUDA(h, cudaSetDevice(h->device));
  const int L = h->trace_len;
  const int64_t n = h->n;
  for (int64_t t = 0; t < n; ++t)
    if (gene_of[t] < 0 || gene_of[t] >= G.ngroups) return mmq_fail(h, MMQ_ERR_ARG, "mmq_prop_summaries: gene index out of range");
  int rc = mmq_ensure_trace_groups(h);
  if (rc) return rc;
  int32_t* d_gene = nullptr; uint8_t* d_multi = nullptr;
  double *d_prop = nullptr, *d_mean = nullptr, *d_s = nullptr, *d_ss = nullptr;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t b) { if (e == cudaSuccess) e = cudaMalloc(p, b ? b : 16); };
  A((void**)&d_gene, sizeof(int32_t) * n); A((void**)&d_multi, n); A((void**)&d_prop, sizeof(double) * n * L);
  A((void**)&d_mean, sizeof(double) * n); A((void**)&d_s, sizeof(double) * n); A((void**)&d_ss, sizeof(double) * n);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_gene, gene_of, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_multi, multi_iso, n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    const int grid = (int)std::min<int64_t>(n, (int64_t)h->num_sms * 8);
    k_prop<<<grid, MMQ_POST_THREADS, 0, h->stream>>>(h->trace, G.trace_dev, d_gene, d_multi, n, L, d_prop, d_mean, d_s, d_ss);
    g_mmq_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  auto D = [&](void* dst, const void* src, size_t b) { if (e == cudaSuccess && dst) e = cudaMemcpyAsync(dst, src, b, cudaMemcpyDeviceToHost, h->stream); };
  D(mean_prop, d_mean, sizeof(double) * n); D(sum_probit, d_s, sizeof(double) * n); D(sumsq_probit, d_ss, sizeof(double) * n);
  D(prop_trace_out, d_prop, sizeof(double) * n * L);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  int ret = MMQ_OK;
  if (e != cudaSuccess) ret = mmq_cuda_fail(h, e, "mmq_prop_summaries", __FILE__, __LINE__);
  if (ret == MMQ_OK && npct > 0 && pct) {
    if (!pow2_ok(L)) ret = mmq_fail(h, MMQ_ERR_ARG, "mmq_prop_summaries: trace_len must be a power of two in [4, 2048]");
    else ret = run_row_summaries(h, h->stream, d_prop, n, L, 0, nullptr, nullptr, nullptr, nullptr, nullptr, npct, pct_idx, pct);
  }
  cudaFree(d_gene); cudaFree(d_multi); cudaFree(d_prop); cudaFree(d_mean); cudaFree(d_s); cudaFree(d_ss);
  return ret;
}

int mmq_unique_hits_sets(mmq_handle* h, const int32_t* set_of, int64_t nsets, int32_t* out) {
  if (!h || !set_of || !out || nsets < 0) return mmq_fail(h, MMQ_ERR_ARG, "mmq_unique_hits_sets: bad arguments");
  if (nsets == 0) return MMQ_OK;
  for (int64_t t = 0; t < h->n; ++t)
    if (set_of[t] >= nsets) return mmq_fail(h, MMQ_ERR_ARG, "mmq_unique_hits_sets: set index out of range");
  MMQ_CUDA(h, cudaSetDevice(h->device));
  int32_t *d_set = nullptr, *d_out = nullptr;
  MMQ_CUDA(h, cudaMalloc(&d_set, sizeof(int32_t) * (size_t)h->n));
  cudaError_t e = cudaMalloc(&d_out, sizeof(int32_t) * (size_t)nsets);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_set, set_of, sizeof(int32_t) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_out, 0, sizeof(int32_t) * (size_t)nsets, h->stream);
  int ret = MMQ_OK;
  if (e == cudaSuccess && h->m > 0) {
    const int grid = mmq_grid_for(h->m, 256, h->num_sms * 8);
    if (h->has_k) k_uh_sets<true><<<grid, 256, 0, h->stream>>>(h->row_ptr, h->col, h->k, d_set, h->m, d_out);
    else k_uh_sets<false><<<grid, 256, 0, h->stream>>>(h->row_ptr, h->col, h->k, d_set, h->m, d_out);
    g_mmq_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) ret = mmq_allreduce(h, d_out, (size_t)nsets, 0);
  if (e == cudaSuccess && ret == MMQ_OK) e = cudaMemcpyAsync(out, d_out, sizeof(int32_t) * (size_t)nsets, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) ret = mmq_cuda_fail(h, e, "mmq_unique_hits_sets", __FILE__, __LINE__);
  cudaFree(d_set); cudaFree(d_out);
  return ret;
}

int mmq_prior_draws(int device, int64_t count, const int64_t* ids, const double* rate, double alpha, uint32_t seed,
                    int trace_len, double* out) {
  if (count < 0 || trace_len <= 0 || (count > 0 && (!ids || !rate || !out)) || !(alpha > 0.0))
    return mmq_fail(nullptr, MMQ_ERR_ARG, "mmq_prior_draws: bad arguments");
  if (count == 0) return MMQ_OK;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return mmq_cuda_fail(nullptr, e, "cudaSetDevice", __FILE__, __LINE__);
  int64_t* d_ids = nullptr; double *d_rate = nullptr, *d_out = nullptr;
  auto A = [&](void** p, size_t b) { if (e == cudaSuccess) e = cudaMalloc(p, b); };
  A((void**)&d_ids, sizeof(int64_t) * count); A((void**)&d_rate, sizeof(double) * count);
  A((void**)&d_out, sizeof(double) * count * trace_len);
  if (e == cudaSuccess) e = cudaMemcpy(d_ids, ids, sizeof(int64_t) * count, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_rate, rate, sizeof(double) * count, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_prior<<<mmq_grid_for(count * trace_len, 128, 148 * 16), 128>>>(d_ids, d_rate, count, trace_len, alpha, seed, d_out);
    g_mmq_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(double) * count * trace_len, cudaMemcpyDeviceToHost);
  cudaFree(d_ids); cudaFree(d_rate); cudaFree(d_out);
  if (e != cudaSuccess) return mmq_cuda_fail(nullptr, e, "mmq_prior_draws", __FILE__, __LINE__);
  return MMQ_OK;
}

int mmq_sokal_batch(int device, int64_t rows, int len, const double* x, double* var, double* tau, int32_t* win, int32_t* status) {
  if (rows < 0 || !x) return mmq_fail(nullptr, MMQ_ERR_ARG, "mmq_sokal_batch: bad arguments");
  if (!pow2_ok(len)) {
    /* the reference's return codes: src/sokal.cc:36-40 (100), :108-111 (200), :119-126 (201) */
    const int code = len < 4 ? 200 : (len > MMQ_MAX_ROW_LEN ? 100 : 201);
    if (status) for (int64_t r = 0; r < rows; ++r) status[r] = code;
    return MMQ_OK;
  }
  if (rows == 0) return MMQ_OK;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return mmq_cuda_fail(nullptr, e, "cudaSetDevice", __FILE__, __LINE__);
  double* d_x = nullptr;
  e = cudaMalloc(&d_x, sizeof(double) * (size_t)rows * (size_t)len);
  if (e != cudaSuccess) return mmq_cuda_fail(nullptr, e, "cudaMalloc", __FILE__, __LINE__);
  e = cudaMemcpy(d_x, x, sizeof(double) * (size_t)rows * (size_t)len, cudaMemcpyHostToDevice);
  int ret = MMQ_OK;
  if (e != cudaSuccess) ret = mmq_cuda_fail(nullptr, e, "cudaMemcpy", __FILE__, __LINE__);
  else ret = run_row_summaries(nullptr, 0, d_x, rows, len, 0, nullptr, var, tau, win, status, 0, nullptr, nullptr);
  cudaFree(d_x);
  return ret;
}

} /* extern "C" */
