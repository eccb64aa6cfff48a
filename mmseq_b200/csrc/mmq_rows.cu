/* mmq_rows.cu — K2 for k == 1 shards laid out BY LENGTH (one CSR row per fragment, rows grouped by
 * class size and, inside a size, by hit class: the loader's LAYOUT_PER_FRAGMENT_BY_LENGTH) — the only
 * layout that can carry per-hit weights (north_star (1): likelihood, insert-size and bias weights), i.e.
 * BASELINE config 4's stream.  Replaces, for such shards, the allocation loop of src/mmseq.cpp:862-891.
 *
 * What is compulsory per sweep in such a shard is ONE fp32 weight per hit; the transcript indices are not:
 * the 30 M fragments of the config-2 sample fall into 3.6 M distinct hit classes, so consecutive rows
 * mostly repeat the previous row's columns.  The row plan (built on the device the first time a sweep asks for it:
 * MMQ_GIBBS_ROWS_KERNEL) therefore stores
 *   - the columns ONCE per run of identical rows ("set"), set-major;
 *   - the weights member-major in chunks of 128 rows (entry (j, r) of a chunk at 128 j + r), so that a lane's
 *     four rows are one 16-byte load per member and a warp's load is 512 contiguous bytes;
 *   - per chunk 32 bytes: a 128-bit mask "row r opens a new set", the set of row 0, the run, the weight offset;
 *   - nothing for single-member rows (x = 1 is deterministic: summed once into seg_base[], mmq_seg.cu).
 * A warp takes a chunk: lane l owns rows 4l..4l+3 = the four classes of ONE Philox block of the CAT stream
 * (include/mmq_sampler.h), finds their sets from the mask (popcounts), gathers mu for a set once and reuses
 * it while the set repeats (unweighted: the running sums too), multiplies by the lane's weights, draws.
 * Arithmetic and its order are those of the k == 1 branch of mmq_alloc_row, so the counts equal the CPU
 * replay's bit for bit (tests/test_gpu_parity.py).
 *
 * Algorithmic HBM bytes per sweep: 4 B per weight slot + 4 B per set column + 32 B per chunk
 * (mmq_rows_stats; the segment kernel of mmq_seg.cu streams 8 B per hit).
 *
 * STATUS (round 2, measured on the config-2 sample, profiles/README.md): 1.6x fewer bytes than the segment kernel but
 * no faster — 247 against 245 us per sweep with weights, 177 against 145 us without: the kernel is bound by instruction
 * issue (about 420 warp instructions per 128 rows), not by HBM.  The segment kernel therefore stays the default for
 * these shards; this one runs when MMQ_GIBBS_ROWS_KERNEL is passed.
 */
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mmq_device.cuh"
#include "mmq_internal.h"

#define MMQ_ROWS_CHUNK 128
#define MMQ_ROWS_WARPS 4
#define MMQ_ROWS_MAXRUN 96
#define MMQ_ROWS_DREG 8 /* class sizes up to this are register-resident template instances */

struct mmq_rows_run {
  int64_t cid_virtual; /* class id of the run's virtual row 0: a multiple of 4 */
  int64_t col_base;    /* offset of the run's first set in set_col */
  int64_t q0;          /* CSR offset of the run's first real row (plan build only) */
  int32_t row_lo;      /* virtual rows [0, row_lo) are dummies in front of the first real row */
  int32_t vrows;       /* virtual rows (dummies included) */
  int32_t d;
  int32_t chunk0;
  int32_t set_base;    /* global index of the run's first set */
  int32_t pad;
};
struct mmq_rows_meta { /* one per chunk, 32 bytes */
  uint32_t mask[4];  /* bit r: row r of the chunk opens a new set (bit 0 is not used) */
  int32_t set_first; /* set of row 0, relative to the run's first set (-1 for the dummies in front of a run) */
  int32_t run;
  int64_t woff;      /* offset of the chunk's weights in w_mm */
};

/* ------------------------------------------------------------------ plan build (device) */

/* flag[v] = 1 when virtual row v of the run is a real row whose columns differ from the previous row's */
__global__ void k_rows_flags(mmq_rows_run R, const int32_t* __restrict__ col, int32_t* __restrict__ flag, int64_t vbase, int64_t vtotal) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < vtotal; v += (int64_t)gridDim.x * blockDim.x) {
    int f = 0;
    if (v >= R.row_lo && v < R.vrows) {
      const int32_t* c = col + R.q0 + (v - R.row_lo) * R.d;
      if (v == R.row_lo) f = 1;
      else
        for (int j = 0; j < R.d; ++j)
          if (c[j] != c[j - R.d]) { f = 1; break; }
    }
    flag[vbase + v] = f;
  }
}

/* set columns, member-major weights and chunk metadata of one run; one warp per 32 virtual rows */
template <bool HAS_W>
__global__ void k_rows_fill(mmq_rows_run R, int run_index, const int32_t* __restrict__ col, const float* __restrict__ w,
                            const int32_t* __restrict__ flag, const int32_t* __restrict__ incl, int64_t vbase, int64_t vtotal,
                            int64_t wbase, int32_t* __restrict__ set_col, float* __restrict__ w_mm, mmq_rows_meta* __restrict__ meta) {
  const int lane = threadIdx.x & 31;
  for (int64_t v0 = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * 32; v0 < vtotal; v0 += ((int64_t)gridDim.x * blockDim.x >> 5) * 32) {
    const int64_t v = v0 + lane;
    const int f = flag[vbase + v];
    const unsigned bits = __ballot_sync(0xffffffffu, f != 0);
    const int64_t chunk = R.chunk0 + v / MMQ_ROWS_CHUNK;
    const int r = (int)(v % MMQ_ROWS_CHUNK);
    const int64_t woff = wbase + (v / MMQ_ROWS_CHUNK) * (int64_t)MMQ_ROWS_CHUNK * R.d;
    if (lane == 0) {
      meta[chunk].mask[r >> 5] = bits;
      if (r == 0) {
        meta[chunk].set_first = incl[vbase + v] - 1 - R.set_base; /* -1: the chunk starts with the run's leading dummies */
        meta[chunk].run = run_index;
        meta[chunk].woff = woff;
      }
    }
    const bool real = v >= R.row_lo && v < R.vrows;
    const int64_t q = R.q0 + (v - R.row_lo) * R.d;
    if (f) {
      int32_t* dst = set_col + R.col_base + (int64_t)(incl[vbase + v] - 1 - R.set_base) * R.d;
      for (int j = 0; j < R.d; ++j) dst[j] = col[q + j];
    }
    if (HAS_W)
      for (int j = 0; j < R.d; ++j) w_mm[woff + (int64_t)j * MMQ_ROWS_CHUNK + r] = real ? w[q + j] : 0.f;
  }
}

/* ------------------------------------------------------------------ the sweep kernel */

__device__ __forceinline__ double rows_w2d(float f) { return (double)f; }

template <int D, bool HAS_W>
__device__ __forceinline__ void rows_chunk(int row_lo, int vrows, int s0, unsigned nb, int vrow0, uint4 wq4,
                                           const int32_t* __restrict__ setp, const float* __restrict__ wchunk, const double* __restrict__ mu,
                                           int32_t* __restrict__ counts, int lane) {
  float4 wv[HAS_W ? D : 1];
  if (HAS_W) {
    const float4* wp = reinterpret_cast<const float4*>(wchunk) + lane;
#pragma unroll
    for (int j = 0; j < D; ++j) wv[j] = __ldg(wp + j * (MMQ_ROWS_CHUNK / 4));
  }
  /* the set of each of the lane's four rows; straight-line code: a row always re-reads its set's columns and mu (L1 hits
   * when the set repeats) — cheaper than branching on "same set as the row before" (measured: the divergent reload
   * branches and their reconvergence were 12 % of the instructions of the first version) */
  int sr[4];
  sr[0] = s0; /* -1: the dummy rows in front of a run (clamped where it is used) */
#pragma unroll
  for (int i = 1; i < 4; ++i) sr[i] = sr[i - 1] + (int)((nb >> (i - 1)) & 1u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int32_t* cp = setp + (int64_t)(sr[i] < 0 ? 0 : sr[i]) * D;
    double p[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double gj = mu[__ldg(cp + j)];
      if (HAS_W) {
        const float wf = i == 0 ? wv[j].x : i == 1 ? wv[j].y : i == 2 ? wv[j].z : wv[j].w;
        p[j] = gj * rows_w2d(wf);
      } else {
        p[j] = gj;
      }
    }
#pragma unroll
    for (int j = 1; j < D; ++j) p[j] = p[j - 1] + p[j]; /* running sums, left to right */
    const double target = mmq_uniform32(i == 0 ? wq4.x : i == 1 ? wq4.y : i == 2 ? wq4.z : wq4.w) * p[D - 1];
    int chosen = D - 1;
#pragma unroll
    for (int j = 0; j < D - 1; ++j) chosen -= (target < p[j]) ? 1 : 0; /* non-decreasing: first j with target < S_j */
    const int vr = vrow0 + i;
    cat_red(counts, (vr >= row_lo && vr < vrows) ? __ldg(cp + chosen) : -1, lane);
  }
}

/* class sizes 9..16: mu of the set in registers (reused while the set repeats), the lane's weights read row by row
 * (the chunk's weight lines stay in L1 between its four rows), two passes over the members instead of stored sums */
template <int D, bool HAS_W>
__device__ __forceinline__ void rows_chunk_mid(int D_, int row_lo, int vrows, int s0, unsigned nb, int vrow0, uint4 wq4,
                                            const int32_t* __restrict__ setp, const float* __restrict__ wchunk, const double* __restrict__ mu,
                                            int32_t* __restrict__ counts, int lane) {
  double g[D];
  const int32_t* cp = setp;
  int cur = -2;
  int s = s0;
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    if (i > 0) s += (int)((nb >> (i - 1)) & 1u);
    if (s != cur) {
      cur = s;
      cp = setp + (int64_t)(s < 0 ? 0 : s) * D;
#pragma unroll
      for (int j = 0; j < D; ++j) g[j] = mu[__ldg(cp + j)];
    }
    const float* wp = wchunk + 4 * lane + i;
    double p[D];
#pragma unroll
    for (int j = 0; j < D; ++j) p[j] = HAS_W ? g[j] * (double)__ldg(wp + j * MMQ_ROWS_CHUNK) : g[j];
    double norm = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) norm += p[j];
    const double target = mmq_uniform32(i == 0 ? wq4.x : i == 1 ? wq4.y : i == 2 ? wq4.z : wq4.w) * norm;
    double acc = 0.0;
    int chosen = D - 1;
#pragma unroll
    for (int j = 0; j < D - 1; ++j) { acc += p[j]; chosen -= (target < acc) ? 1 : 0; }
    const int vr = vrow0 + i;
    cat_red(counts, (vr >= row_lo && vr < vrows) ? __ldg(cp + chosen) : -1, lane);
  }
}

/* any class size: members re-read per row (L1 / L2 hits) */
template <bool HAS_W>
__device__ __noinline__ void rows_chunk_any(int D_, int row_lo, int vrows, int s0, unsigned nb, int vrow0, uint4 wq4,
                                            const int32_t* __restrict__ setp, const float* __restrict__ wchunk, const double* __restrict__ mu,
                                            int32_t* __restrict__ counts, int lane) {
  const int D = D_;
  int s = s0;
  for (int i = 0; i < 4; ++i) {
    if (i > 0) s += (int)((nb >> (i - 1)) & 1u);
    const int32_t* cp = setp + (int64_t)(s < 0 ? 0 : s) * D;
    const float* wp = wchunk + 4 * lane + i;
    double norm = 0.0;
    for (int j = 0; j < D; ++j) norm += HAS_W ? mu[cp[j]] * rows_w2d(wp[(int64_t)j * MMQ_ROWS_CHUNK]) : mu[cp[j]];
    const double target = mmq_uniform32(i == 0 ? wq4.x : i == 1 ? wq4.y : i == 2 ? wq4.z : wq4.w) * norm;
    double acc = 0.0;
    int chosen = D - 1;
    for (int j = 0; j < D - 1; ++j) {
      acc += HAS_W ? mu[cp[j]] * rows_w2d(wp[(int64_t)j * MMQ_ROWS_CHUNK]) : mu[cp[j]];
      if (target < acc) { chosen = j; break; }
    }
    const int vr = vrow0 + i;
    cat_red(counts, (vr >= row_lo && vr < vrows) ? cp[chosen] : -1, lane);
  }
}

/* MID = false: the chunks of class sizes 2..MMQ_ROWS_DREG ([chunk_begin, chunk_end) = the front of the plan, runs are in
 * ascending class size); MID = true: the larger sizes (their own launch: the register budgets differ) */
template <bool HAS_W, bool MID, int MINB>
__global__ void __launch_bounds__(MMQ_ROWS_WARPS * 32, MINB)
k_alloc_rows(const mmq_rows_run* __restrict__ runs, int nruns, int chunk_begin, int chunks, const mmq_rows_meta* __restrict__ meta,
             const int32_t* __restrict__ set_col, const float* __restrict__ w_mm, const double* __restrict__ mu,
             int32_t* __restrict__ counts, uint32_t seed, uint32_t sweep, const uint32_t* __restrict__ sweep_base) {
  if (sweep_base) sweep += *sweep_base;
  __shared__ mmq_rows_run s_run[MMQ_ROWS_MAXRUN];
  for (int i = threadIdx.x; i < nruns; i += blockDim.x) s_run[i] = runs[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * MMQ_ROWS_WARPS;
  /* the chunk's 32 bytes of metadata are fetched one chunk ahead: their latency is off the critical path */
  int chunk = chunk_begin + blockIdx.x * MMQ_ROWS_WARPS + (threadIdx.x >> 5);
  uint4 ma = make_uint4(0u, 0u, 0u, 0u), mb = ma;
  if (chunk < chunks) {
    const uint4* mp = reinterpret_cast<const uint4*>(meta + chunk); /* the same 32 bytes for all lanes: one transaction */
    ma = __ldg(mp); mb = __ldg(mp + 1);
  }
  for (; chunk < chunks; chunk += nwarps) {
    mmq_rows_meta M;
    M.mask[0] = ma.x; M.mask[1] = ma.y; M.mask[2] = ma.z; M.mask[3] = ma.w;
    M.set_first = (int32_t)mb.x; M.run = (int32_t)mb.y;
    M.woff = (int64_t)(((unsigned long long)mb.w << 32) | mb.z);
    if (chunk + nwarps < chunks) {
      const uint4* mp = reinterpret_cast<const uint4*>(meta + chunk + nwarps);
      ma = __ldg(mp); mb = __ldg(mp + 1);
    }
    const mmq_rows_run& R = s_run[M.run];
    const int vrow0 = (chunk - R.chunk0) * MMQ_ROWS_CHUNK + 4 * lane;
    const int row_lo = R.row_lo, vrows = R.vrows, Dr = R.d;
    const int32_t* setp = set_col + R.col_base;
    const float* wchunk = w_mm + M.woff;
    /* set of row 4 lane: the chunk's first set + the "new set" bits of rows 1 .. 4 lane */
    const int wq = lane >> 3, bq = (4 * lane) & 31;
    int s0 = M.set_first;
    s0 += wq > 0 ? __popc(M.mask[0]) : 0;
    s0 += wq > 1 ? __popc(M.mask[1]) : 0;
    s0 += wq > 2 ? __popc(M.mask[2]) : 0;
    const uint32_t mw = wq == 0 ? M.mask[0] : wq == 1 ? M.mask[1] : wq == 2 ? M.mask[2] : M.mask[3];
    s0 += __popc(mw & ((2u << bq) - 1u));
    s0 -= (int)(M.mask[0] & 1u); /* bit 0 of the chunk is not a boundary inside the chunk */
    const unsigned nb = (mw >> (bq + 1)) & 7u; /* rows 4 lane + 1 .. + 3 */
    const uint64_t cid = (uint64_t)(R.cid_virtual + vrow0);
    uint32_t wd[4] = {(uint32_t)(cid >> 2), (uint32_t)(cid >> 34), sweep, 0u};
    mmq_philox4x32_10(wd, seed, MMQ_STREAM_CAT);
#define MMQ_ROWS_CASE(DD) case DD: rows_chunk<DD, HAS_W>(row_lo, vrows, s0, nb, vrow0, make_uint4(wd[0], wd[1], wd[2], wd[3]), setp, wchunk, mu, counts, lane); break;
    if (!MID) {
      switch (Dr) {
        MMQ_ROWS_CASE(2) MMQ_ROWS_CASE(3) MMQ_ROWS_CASE(4) MMQ_ROWS_CASE(5) MMQ_ROWS_CASE(6) MMQ_ROWS_CASE(7) MMQ_ROWS_CASE(8)
        default: break;
      }
    } else {
#define MMQ_ROWS_MID(DD) case DD: rows_chunk_mid<DD, HAS_W>(DD, row_lo, vrows, s0, nb, vrow0, make_uint4(wd[0], wd[1], wd[2], wd[3]), setp, wchunk, mu, counts, lane); break;
      switch (Dr) {
        MMQ_ROWS_MID(9) MMQ_ROWS_MID(10) MMQ_ROWS_MID(11) MMQ_ROWS_MID(12) MMQ_ROWS_MID(13) MMQ_ROWS_MID(14) MMQ_ROWS_MID(15) MMQ_ROWS_MID(16)
        default: rows_chunk_any<HAS_W>(Dr, row_lo, vrows, s0, nb, vrow0, make_uint4(wd[0], wd[1], wd[2], wd[3]), setp, wchunk, mu, counts, lane); break;
      }
#undef MMQ_ROWS_MID
    }
#undef MMQ_ROWS_CASE
  }
}

/* ------------------------------------------------------------------ host */

int mmq_rows_plan(mmq_handle* h) {
  h->rows_ready = false;
  h->rows_tried = true;
  static const bool off = [] { const char* e = getenv("MMQ_ROWS_OFF"); return e && atoi(e) != 0; }();
  if (off || !h->seg_scan_ok || h->has_k || h->m == 0) return MMQ_OK;
  std::vector<mmq_rows_run> runs;
  std::vector<int64_t> vbase, wbase;
  int64_t chunks = 0, chunks_small = 0, vtot = 0, wtot = 0, rows = 0;
  int last_d = 0;
  for (const auto& r : h->seg_runs) {
    if (r.d == 1) continue;
    if (r.d > 0xffff || (r.r1 - r.r0 + 3) > 0x7ffffff0ll || r.d < last_d) return MMQ_OK; /* runs must come in ascending class size */
    last_d = r.d;
    mmq_rows_run R;
    const int lead = (int)((h->class_id_base + r.r0) & 3); /* dummy rows: the virtual first class id is a multiple of 4 */
    R.cid_virtual = h->class_id_base + r.r0 - lead;
    R.col_base = 0; R.q0 = r.q0;
    R.row_lo = lead;
    R.vrows = (int32_t)(r.r1 - r.r0 + lead);
    R.d = r.d;
    R.chunk0 = (int32_t)chunks;
    R.set_base = 0; R.pad = 0;
    const int64_t nch = ((int64_t)R.vrows + MMQ_ROWS_CHUNK - 1) / MMQ_ROWS_CHUNK;
    if (r.d <= MMQ_ROWS_DREG) chunks_small = chunks + nch;
    vbase.push_back(vtot); wbase.push_back(wtot);
    chunks += nch; vtot += nch * MMQ_ROWS_CHUNK; wtot += nch * MMQ_ROWS_CHUNK * r.d;
    rows += r.r1 - r.r0;
    if (chunks > 0x7fff0000ll || vtot > 0x7fff0000ll) return MMQ_OK;
    runs.push_back(R);
  }
  if (runs.empty() || (int)runs.size() > MMQ_ROWS_MAXRUN) return MMQ_OK;
  int rc;
  /* 1. new-set flags, 2. inclusive scan = set numbering, 3. per-run set counts back to the host, 4. fill */
  int32_t *flag = nullptr, *incl = nullptr;
  void* temp = nullptr;
  size_t temp_bytes = 0;
  MMQ_CUDA(h, cudaMalloc(&flag, sizeof(int32_t) * (size_t)vtot));
  MMQ_CUDA(h, cudaMalloc(&incl, sizeof(int32_t) * (size_t)vtot));
  auto cleanup = [&] { cudaFree(flag); cudaFree(incl); if (temp) cudaFree(temp); };
  for (size_t i = 0; i < runs.size(); ++i) {
    const int64_t vt = ((int64_t)runs[i].vrows + MMQ_ROWS_CHUNK - 1) / MMQ_ROWS_CHUNK * MMQ_ROWS_CHUNK;
    k_rows_flags<<<mmq_grid_for(vt, 256, h->num_sms * 8), 256, 0, h->stream>>>(runs[i], h->col, flag, vbase[i], vt);
    g_mmq_launches.fetch_add(1);
  }
  cudaError_t e = cub::DeviceScan::InclusiveSum(nullptr, temp_bytes, flag, incl, (int)vtot, h->stream);
  if (e == cudaSuccess) e = cudaMalloc(&temp, temp_bytes);
  if (e == cudaSuccess) e = cub::DeviceScan::InclusiveSum(temp, temp_bytes, flag, incl, (int)vtot, h->stream);
  if (e != cudaSuccess) { cleanup(); return mmq_cuda_fail(h, e, "row plan scan", __FILE__, __LINE__); }
  std::vector<int32_t> last(runs.size());
  for (size_t i = 0; i < runs.size(); ++i) {
    const int64_t vend = (i + 1 < runs.size() ? vbase[i + 1] : vtot) - 1;
    e = cudaMemcpyAsync(&last[i], incl + vend, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    if (e != cudaSuccess) { cleanup(); return mmq_cuda_fail(h, e, "row plan copy", __FILE__, __LINE__); }
  }
  e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { cleanup(); return mmq_cuda_fail(h, e, "row plan sync", __FILE__, __LINE__); }
  int64_t sets = 0, set_cols = 0;
  for (size_t i = 0; i < runs.size(); ++i) {
    const int64_t prev = i ? last[i - 1] : 0;
    runs[i].set_base = (int32_t)prev;
    runs[i].col_base = set_cols;
    sets += last[i] - prev;
    set_cols += (int64_t)(last[i] - prev) * runs[i].d;
  }
  if ((rc = mmq_dev_alloc(h, (void**)&h->rows_set_col, sizeof(int32_t) * (size_t)std::max<int64_t>(set_cols, 1)))) { cleanup(); return rc; }
  if (h->has_w && (rc = mmq_dev_alloc(h, (void**)&h->rows_w, sizeof(float) * (size_t)wtot))) { cleanup(); return rc; }
  if ((rc = mmq_dev_alloc(h, &h->rows_meta, sizeof(mmq_rows_meta) * (size_t)chunks))) { cleanup(); return rc; }
  if ((rc = mmq_dev_alloc(h, &h->rows_runs, sizeof(mmq_rows_run) * runs.size()))) { cleanup(); return rc; }
  for (size_t i = 0; i < runs.size(); ++i) {
    const int64_t vt = ((int64_t)runs[i].vrows + MMQ_ROWS_CHUNK - 1) / MMQ_ROWS_CHUNK * MMQ_ROWS_CHUNK;
    const int grid = mmq_grid_for(vt, 256, h->num_sms * 8);
    if (h->has_w) k_rows_fill<true><<<grid, 256, 0, h->stream>>>(runs[i], (int)i, h->col, h->w, flag, incl, vbase[i], vt, wbase[i], h->rows_set_col, h->rows_w, (mmq_rows_meta*)h->rows_meta);
    else k_rows_fill<false><<<grid, 256, 0, h->stream>>>(runs[i], (int)i, h->col, h->w, flag, incl, vbase[i], vt, wbase[i], h->rows_set_col, h->rows_w, (mmq_rows_meta*)h->rows_meta);
    g_mmq_launches.fetch_add(1);
  }
  e = cudaMemcpyAsync(h->rows_runs, runs.data(), sizeof(mmq_rows_run) * runs.size(), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cleanup();
  if (e != cudaSuccess) return mmq_cuda_fail(h, e, "row plan fill", __FILE__, __LINE__);
  h->rows_nruns = (int)runs.size();
  h->rows_chunks = chunks;
  h->rows_chunks_small = chunks_small;
  h->rows_rows = rows;
  h->rows_sets = sets;
  h->rows_set_cols = set_cols;
  h->rows_wslots = h->has_w ? wtot : 0;
  h->rows_ready = true;
  return MMQ_OK;
}

extern "C" int mmq_rows_stats(mmq_handle* h, int64_t out[8]) {
  if (!h || !out) return MMQ_ERR_ARG;
  if (!h->rows_ready && !h->rows_tried && !h->has_k) { /* the plan is built on first use */
    MMQ_CUDA(h, cudaSetDevice(h->device));
    int rc = mmq_rows_plan(h);
    if (rc) return rc;
  }
  out[0] = h->rows_ready ? 1 : 0;
  out[1] = h->rows_rows;
  out[2] = h->rows_sets;
  out[3] = h->rows_set_cols;
  out[4] = h->rows_wslots;
  out[5] = h->rows_chunks;
  out[6] = 4 * h->rows_wslots + 4 * h->rows_set_cols + (int64_t)sizeof(mmq_rows_meta) * h->rows_chunks; /* bytes streamed per sweep */
  out[7] = h->seg_singletons;
  return MMQ_OK;
}

int mmq_rows_launch(mmq_handle* h, uint32_t seed, uint32_t sweep, const uint32_t* sweep_base) {
  int rc = mmq_seg_add_base(h, true);
  if (rc) return rc;
  if (h->rows_chunks == 0) return MMQ_OK;
  const int chunks = (int)h->rows_chunks, csmall = (int)h->rows_chunks_small;
#define MMQ_ROWS_GO(W, MID, MINB, c0, c1)                                                                                        \
  do {                                                                                                                           \
    const int grid = std::min(((c1) - (c0) + MMQ_ROWS_WARPS - 1) / MMQ_ROWS_WARPS, h->num_sms * MINB);                             \
    k_alloc_rows<W, MID, MINB><<<grid, MMQ_ROWS_WARPS * 32, 0, h->stream>>>((const mmq_rows_run*)h->rows_runs, h->rows_nruns, (c0), (c1), \
                                                                           (const mmq_rows_meta*)h->rows_meta, h->rows_set_col, \
                                                                           h->rows_w, h->mu, h->counts, seed, sweep, sweep_base); \
    MMQ_LAUNCHED(h);                                                                                                             \
  } while (0)
  if (chunks > csmall) { /* the long rows first: few chunks, long dependent work */
    if (h->has_w) MMQ_ROWS_GO(true, true, 4, csmall, chunks);
    else MMQ_ROWS_GO(false, true, 4, csmall, chunks);
  }
  if (csmall > 0) {
    if (h->has_w) MMQ_ROWS_GO(true, false, 5, 0, csmall);
    else MMQ_ROWS_GO(false, false, 8, 0, csmall);
  }
#undef MMQ_ROWS_GO
  g_mmq_launches.fetch_sub(1, std::memory_order_relaxed); /* the caller counts one */
  return MMQ_OK;
}
