/* mmq_seg.cu — K2 for k == 1 shards laid out BY LENGTH (the loader's
 * LAYOUT_PER_FRAGMENT_BY_LENGTH): the rows of the shard come in a few long runs of equal
 * class size d.  Replaces, for such shards, the allocation loop of src/mmseq.cpp:862-891.
 *
 * What the plan (built once in mmq_create) buys every sweep:
 *   - no row pointers are read at all: inside a run, row r starts at e0 + r*d;
 *   - classes with a single member (d == 1) are not visited: their allocation is
 *     deterministic (x = k, no random number — include/mmq_sampler.h), so their counts
 *     are summed once into seg_base[] and the Gamma kernel restarts counts[] from it;
 *   - every warp works on 64 consecutive classes of ONE length, two per lane, with
 *     straight-line code specialised on d (2..8): 8- or 16-byte vector loads of the lane's
 *     2d contiguous columns (runs are re-packed 16-byte aligned), 2d independent mu
 *     gathers in flight, left-to-right running sums in registers, no divergence, no shared
 *     memory; one Philox block per lane serves both classes.
 * Arithmetic and its order are those of the k == 1 branch of mmq_alloc_row, so the counts
 * equal the CPU replay's bit for bit (tests/test_gpu_parity.py).
 *
 * Algorithmic HBM bytes per sweep: 4 B (+4 B weight) per CSR entry of the classes with
 * d >= 2 — nothing else is streamed.
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mmq_device.cuh"
#include "mmq_internal.h"

#define MMQ_SEG_MAX 96    /* more runs than this: not a by-length layout, use the ragged kernel */
#define MMQ_SEG_WARPS 8
#define MMQ_SEG_ROWS 64

struct mmq_seg {
  int64_t e_virtual;   /* packed-array offset of the (possibly dummy) virtual first row; multiple of 4 */
  int64_t cid_virtual; /* class id of the virtual first row; EVEN, so lane pairs are Philox pairs */
  int32_t row_lo;      /* 0, or 1 when the virtual first row is a dummy */
  int32_t rows;        /* virtual row count (dummy included) */
  int32_t d;           /* class size of the run */
  int32_t pad_;
  int64_t chunk0;      /* first 64-row chunk of this run in the global chunk numbering */
};

__global__ void k_fill_i32(int32_t* __restrict__ p, int64_t count, int32_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_axpy_i32(int32_t* __restrict__ y, const int32_t* __restrict__ x, int64_t count, int sign) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) y[i] += sign * x[i];
}
/* base[col[q]] += 1 over the CSR entries [q0, q1) of a run of singleton classes */
__global__ void k_count_singletons(const int32_t* __restrict__ col, int64_t q0, int64_t q1, int32_t* __restrict__ base) {
  for (int64_t q = q0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < q1; q += (int64_t)gridDim.x * blockDim.x) atomicAdd(base + col[q], 1);
}

/* Both classes of a lane, class size D known at compile time.  c[0..D) / c[D..2D) are the
 * columns of class a / b (sentinel for an absent class), p the matching probabilities.
 * Same arithmetic and order as mmq_alloc_row's k == 1 branch. */
template <int D>
__device__ __forceinline__ int32_t seg_pick(const int32_t* c, const double* p, double u) {
  double S[D];
  S[0] = p[0];
#pragma unroll
  for (int j = 1; j < D; ++j) S[j] = S[j - 1] + p[j];
  const double target = u * S[D - 1];
  int chosen = -1;
#pragma unroll
  for (int j = D - 1; j >= 0; --j)
    if (target < S[j]) chosen = j; /* descending: the smallest hit index wins */
  if (chosen < 0) { /* rounding at the top end or an all-zero row: last member with p > 0, else the last */
    chosen = D - 1;
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (p[j] > 0.0) chosen = j;
    bool any = false;
#pragma unroll
    for (int j = 0; j < D; ++j) any |= p[j] > 0.0;
    if (!any) chosen = D - 1;
  }
  int32_t out = c[0];
#pragma unroll
  for (int j = 1; j < D; ++j)
    if (chosen == j) out = c[j];
  return out;
}

template <int D, bool HAS_W>
__device__ __forceinline__ void seg_chunk_fixed(const int32_t* __restrict__ colp, const float* __restrict__ wp, int64_t e,
                                                bool va, bool vb, double ua, double ub, const double* __restrict__ mu,
                                                int32_t sentinel, int32_t& out_a, int32_t& out_b) {
  int32_t c[2 * D];
  float wv[2 * D];
  if (va || vb) { /* the lane's 2D columns are contiguous and 8-byte aligned (16-byte when D is even) */
    if (D % 2 == 0) {
#pragma unroll
      for (int j = 0; j < 2 * D; j += 4) {
        const int4 v = *reinterpret_cast<const int4*>(colp + e + j);
        c[j] = v.x; c[j + 1] = v.y; c[j + 2] = v.z; c[j + 3] = v.w;
        if (HAS_W) {
          const float4 f = *reinterpret_cast<const float4*>(wp + e + j);
          wv[j] = f.x; wv[j + 1] = f.y; wv[j + 2] = f.z; wv[j + 3] = f.w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 2 * D; j += 2) {
        const int2 v = *reinterpret_cast<const int2*>(colp + e + j);
        c[j] = v.x; c[j + 1] = v.y;
        if (HAS_W) {
          const float2 f = *reinterpret_cast<const float2*>(wp + e + j);
          wv[j] = f.x; wv[j + 1] = f.y;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < D; ++j) { /* absent classes (dummy first row, tail of the run) gather the sentinel: p = 0 */
    if (!va) { c[j] = sentinel; if (HAS_W) wv[j] = 0.f; }
    if (!vb) { c[D + j] = sentinel; if (HAS_W) wv[D + j] = 0.f; }
  }
  /* Class a first, then class b in the same registers.  Rows are grouped by class, so the lane's two
   * classes usually have the same members: b then reuses a's mu (weights stay per row) instead of
   * gathering again.  Running sums are recomputed in the scan instead of being kept: D doubles of mu
   * are all the state a class needs, which lets sizes up to 12 stay in registers. */
  bool same = true;
#pragma unroll
  for (int j = 0; j < D; ++j) same &= c[j] == c[D + j];
  double g[D];
#pragma unroll
  for (int j = 0; j < D; ++j) g[j] = mu[c[j]];
  auto pick = [&](const int32_t* cc, const float* ww, double u) -> int32_t {
    double norm = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) norm += HAS_W ? g[j] * (double)ww[j] : g[j];
    const double target = u * norm;
    double acc = 0.0;
    int chosen = -1, lastpos = -1;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double pj = HAS_W ? g[j] * (double)ww[j] : g[j];
      acc += pj;
      if (chosen < 0 && target < acc) chosen = j;
      if (pj > 0.0) lastpos = j;
    }
    if (chosen < 0) chosen = lastpos >= 0 ? lastpos : D - 1; /* rounding at the top end / all-zero row */
    int32_t out = cc[0];
#pragma unroll
    for (int j = 1; j < D; ++j)
      if (chosen == j) out = cc[j];
    return out;
  };
  out_a = va ? pick(c, wv, ua) : -1;
#pragma unroll
  for (int j = 0; j < D; ++j)
    if (!same) g[j] = mu[c[D + j]];
  out_b = vb ? pick(c + D, wv + D, ub) : -1;
}

/* One class of compile-time size D (7..12), everything in registers: D columns by vector loads,
 * D independent mu gathers, sum, scan.  The chosen column is re-read (an L1 hit) instead of being
 * selected from registers. */
template <int D, bool HAS_W>
__device__ __noinline__ int32_t seg_row_fixed(const int32_t* __restrict__ cp, const float* __restrict__ wq,
                                                 const double* __restrict__ mu, double u) {
  int32_t c[D];
  float wv[HAS_W ? D : 1];
  if (D % 2 == 0) { /* the row start is 8-byte aligned for even D (two rows of a lane start 16-byte aligned) */
#pragma unroll
    for (int j = 0; j < D; j += 2) {
      const int2 v = *reinterpret_cast<const int2*>(cp + j);
      c[j] = v.x; c[j + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < D; ++j) c[j] = cp[j];
  }
  if (HAS_W) {
#pragma unroll
    for (int j = 0; j < D; ++j) wv[j] = wq[j];
  }
  double g[D];
#pragma unroll
  for (int j = 0; j < D; ++j) g[j] = HAS_W ? mu[c[j]] * (double)wv[j] : mu[c[j]];
  double norm = 0.0;
#pragma unroll
  for (int j = 0; j < D; ++j) norm += g[j];
  const double target = u * norm;
  double acc = 0.0;
  int chosen = -1, lastpos = -1;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    acc += g[j];
    if (chosen < 0 && target < acc) chosen = j;
    if (g[j] > 0.0) lastpos = j;
  }
  if (chosen < 0) chosen = lastpos >= 0 ? lastpos : D - 1;
  return cp[chosen];
}

/* Any class size: two passes with batched gathers straight from global memory. */
template <bool HAS_W>
__device__ __noinline__ int32_t seg_row_generic(const int32_t* __restrict__ c, const float* __restrict__ wv, int d,
                                                   const double* __restrict__ mu, double u) {
#define MMQ_PG(j) (HAS_W ? mu[c[j]] * (double)wv[j] : mu[c[j]])
  double norm = 0.0;
  int j = 0;
  for (; j + 4 <= d; j += 4) {
    const double a0 = MMQ_PG(j), a1 = MMQ_PG(j + 1), a2 = MMQ_PG(j + 2), a3 = MMQ_PG(j + 3);
    norm += a0; norm += a1; norm += a2; norm += a3;
  }
  for (; j < d; ++j) norm += MMQ_PG(j);
  const double target = u * norm;
  double acc = 0.0;
  int chosen = -1;
  for (j = 0; j + 4 <= d && chosen < 0; j += 4) {
    const double a0 = MMQ_PG(j), a1 = MMQ_PG(j + 1), a2 = MMQ_PG(j + 2), a3 = MMQ_PG(j + 3);
    acc += a0; if (chosen < 0 && target < acc) chosen = j;
    acc += a1; if (chosen < 0 && target < acc) chosen = j + 1;
    acc += a2; if (chosen < 0 && target < acc) chosen = j + 2;
    acc += a3; if (chosen < 0 && target < acc) chosen = j + 3;
  }
  for (; j < d && chosen < 0; ++j) { acc += MMQ_PG(j); if (target < acc) chosen = j; }
  if (chosen < 0) {
    chosen = d - 1;
    for (j = d - 1; j >= 0; --j)
      if (MMQ_PG(j) > 0.0) { chosen = j; break; }
  }
#undef MMQ_PG
  return c[chosen];
}

template <bool HAS_W, int MAXD, int OCC>
__global__ void __launch_bounds__(MMQ_SEG_WARPS * 32, OCC)
k_alloc_seg(const mmq_seg* __restrict__ segs, int nsegs, int64_t total_chunks, const int32_t* __restrict__ colp,
            const float* __restrict__ wp, const double* __restrict__ mu, int32_t* __restrict__ counts, uint32_t seed,
            uint32_t sweep, int32_t sentinel, int red_mode, int dbg_dmin, int dbg_dmax, const uint32_t* __restrict__ sweep_base) {
  if (sweep_base) sweep += *sweep_base; /* CUDA-graph replays: the sweep counter lives on the device */
  __shared__ mmq_seg s_seg[MMQ_SEG_MAX];
  for (int i = threadIdx.x; i < nsegs; i += blockDim.x) s_seg[i] = segs[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * MMQ_SEG_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * MMQ_SEG_WARPS;
  int si = 0;
  for (int64_t chunk = warp0; chunk < total_chunks; chunk += nwarps) {
    while (si + 1 < nsegs && chunk >= s_seg[si + 1].chunk0) ++si; /* warp-uniform */
    const mmq_seg sg = s_seg[si];
    const int D = sg.d;
    if (D < dbg_dmin || D > dbg_dmax) continue; /* timing experiments only (MMQ_DEBUG_DMIN / _DMAX) */
    const int rv = (int)(chunk - sg.chunk0) * MMQ_SEG_ROWS + 2 * lane; /* virtual row of class a */
    const bool va = rv >= sg.row_lo && rv < sg.rows;
    const bool vb = rv + 1 < sg.rows; /* rv + 1 >= 1 >= row_lo always */
    const int64_t e = sg.e_virtual + (int64_t)rv * D;
    /* one Philox block per lane: classes cid (even) and cid + 1 */
    const uint64_t cid = (uint64_t)(sg.cid_virtual + rv);
    uint32_t wd[4] = {(uint32_t)(cid >> 1), (uint32_t)(cid >> 33), sweep, 0u};
    mmq_philox4x32_10(wd, seed, MMQ_STREAM_CAT);
    const double ua = cat_u52(wd[0], wd[1]), ub = cat_u52(wd[2], wd[3]);
    int32_t ca = -1, cb = -1;
    /* warp-uniform dispatch on the class size */
#define MMQ_SEG_CASE(DD) else if (MAXD >= DD && D == DD) seg_chunk_fixed<(MAXD >= DD ? DD : 2), HAS_W>(colp, wp, e, va, vb, ua, ub, mu, sentinel, ca, cb);
#define MMQ_SEG_ROWCASE(DD)                                                                            \
  else if (MAXD >= DD && D == DD) {                                                                    \
    if (va) ca = seg_row_fixed<(MAXD >= DD ? DD : 7), HAS_W>(colp + e, wp + e, mu, ua);               \
    if (vb) cb = seg_row_fixed<(MAXD >= DD ? DD : 7), HAS_W>(colp + e + DD, wp + e + DD, mu, ub);     \
  }
    if (D == 2) seg_chunk_fixed<2, HAS_W>(colp, wp, e, va, vb, ua, ub, mu, sentinel, ca, cb);
    MMQ_SEG_CASE(3) MMQ_SEG_CASE(4) MMQ_SEG_CASE(5) MMQ_SEG_CASE(6)
    MMQ_SEG_ROWCASE(7) MMQ_SEG_ROWCASE(8) MMQ_SEG_ROWCASE(9) MMQ_SEG_ROWCASE(10) MMQ_SEG_ROWCASE(11) MMQ_SEG_ROWCASE(12)
#undef MMQ_SEG_CASE
#undef MMQ_SEG_ROWCASE
    else {
      if (va) ca = seg_row_generic<HAS_W>(colp + e, wp + e, D, mu, ua);
      if (vb) cb = seg_row_generic<HAS_W>(colp + e + D, wp + e + D, D, mu, ub);
    }
    cat_red(counts, ca, lane);
    cat_red(counts, cb, lane);
  }
}


/* ---- entry-parallel variant -------------------------------------------------------------
 * Same plan, other mapping of work to lanes.  The chunk's 64*D contiguous columns are fetched
 * by ONE TMA bulk copy into the warp's shared-memory double buffer (issued one chunk ahead).
 * Phase E: lane l takes entries l, l+32, ... of the chunk: neighbouring lanes read neighbouring
 *   entries, i.e. the members of the same few classes, whose mu lie in the same one or two
 *   cache lines (header-order numbering keeps a gene's isoforms adjacent) — a gather request
 *   costs 1-3 L1 wavefronts instead of one per lane — and stores p = mu[col]*w to shared memory.
 * Phase R: lane l owns classes 2l, 2l+1: running sums left to right from shared memory, one
 *   Philox block for both, chosen column read back from the staged columns.
 * Row pitch in the p buffer is odd (D | 1) so the owners' strided reads stay 2-way conflicted
 * at worst. */
#define MMQ_SEG2_DMAX 6
#define MMQ_SEG2_WARPS 8

template <int D, bool HAS_W>
__device__ __forceinline__ void seg2_chunk(const int32_t* __restrict__ sc, const float* __restrict__ sw, double* __restrict__ sp,
                                           int row_lo, int row_hi, double ua, double ub, const double* __restrict__ mu,
                                           int32_t sentinel, int lane, int32_t& out_a, int32_t& out_b) {
  constexpr int PITCH = D | 1;
  /* phase E */
#pragma unroll
  for (int q = 0; q < 2 * D; ++q) {
    const int e = q * 32 + lane;
    const int r = e / D;
    const int j = e - r * D;
    int32_t c = sc[e];
    if (r < row_lo || r >= row_hi) c = sentinel; /* dummy first row / rows past the end of the run */
    double p = mu[c];
    if (HAS_W) p *= (double)sw[e];
    sp[r * PITCH + j] = p;
  }
  __syncwarp();
  /* phase R */
  const int ra = 2 * lane, rb = ra + 1;
  double S[2 * D];
#pragma unroll
  for (int j = 0; j < D; ++j) { S[j] = sp[ra * PITCH + j]; S[D + j] = sp[rb * PITCH + j]; }
  auto pick = [&](const double* p, double u) -> int {
    double R[D];
    R[0] = p[0];
#pragma unroll
    for (int j = 1; j < D; ++j) R[j] = R[j - 1] + p[j];
    const double target = u * R[D - 1];
    int chosen = -1;
#pragma unroll
    for (int j = D - 1; j >= 0; --j)
      if (target < R[j]) chosen = j;
    if (chosen < 0) { /* rounding at the top end or all-zero row: last member with p > 0, else the last */
      chosen = D - 1;
#pragma unroll
      for (int j = 0; j < D; ++j)
        if (p[j] > 0.0) chosen = j;
      bool any = false;
#pragma unroll
      for (int j = 0; j < D; ++j) any |= p[j] > 0.0;
      if (!any) chosen = D - 1;
    }
    return chosen;
  };
  const bool va = ra >= row_lo && ra < row_hi, vb = rb >= row_lo && rb < row_hi;
  out_a = va ? sc[ra * D + pick(S, ua)] : -1;
  out_b = vb ? sc[rb * D + pick(S + D, ub)] : -1;
  __syncwarp(); /* sp is reused by the next chunk */
}

template <bool HAS_W>
__global__ void __launch_bounds__(MMQ_SEG2_WARPS * 32, 3)
k_alloc_seg2(const mmq_seg* __restrict__ segs, int nsegs, int64_t total_chunks, const int32_t* __restrict__ colp,
             const float* __restrict__ wp, const double* __restrict__ mu, int32_t* __restrict__ counts, uint32_t seed,
             uint32_t sweep, int32_t sentinel, const uint32_t* __restrict__ sweep_base) {
  if (sweep_base) sweep += *sweep_base;
  constexpr int CAP = MMQ_SEG_ROWS * MMQ_SEG2_DMAX;                  /* staged entries per buffer */
  constexpr int PCAP = MMQ_SEG_ROWS * (MMQ_SEG2_DMAX | 1);           /* p buffer, doubles */
  constexpr int PER_WARP = 16 + PCAP * 8 + 2 * CAP * 4 * (HAS_W ? 2 : 1);
  extern __shared__ __align__(16) unsigned char seg2_smem[];
  __shared__ mmq_seg s_seg[MMQ_SEG_MAX];
  for (int i = threadIdx.x; i < nsegs; i += blockDim.x) s_seg[i] = segs[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char* wbase = seg2_smem + wib * PER_WARP;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase);
  double* sp = reinterpret_cast<double*>(wbase + 16);
  int32_t* sc = reinterpret_cast<int32_t*>(wbase + 16 + PCAP * 8);            /* [2][CAP] */
  float* sw = reinterpret_cast<float*>(wbase + 16 + PCAP * 8 + 2 * CAP * 4);  /* [2][CAP] when HAS_W */
  const int64_t warp0 = (int64_t)blockIdx.x * MMQ_SEG2_WARPS + wib;
  const int64_t nwarps = (int64_t)gridDim.x * MMQ_SEG2_WARPS;
  if (warp0 >= total_chunks) return;
  if (lane == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  struct Desc { int64_t e0, cid0; int d, row_lo, row_hi; bool staged; };
  int si = 0;
  auto describe = [&](int64_t chunk, int buf) -> Desc { /* also launches the bulk copy of a staged chunk */
    while (si + 1 < nsegs && chunk >= s_seg[si + 1].chunk0) ++si;
    const mmq_seg& sg = s_seg[si];
    Desc ds;
    const int r0 = (int)(chunk - sg.chunk0) * MMQ_SEG_ROWS;
    ds.d = sg.d;
    ds.e0 = sg.e_virtual + (int64_t)r0 * sg.d;
    ds.cid0 = sg.cid_virtual + r0;
    ds.row_lo = r0 == 0 ? sg.row_lo : 0;
    ds.row_hi = sg.rows - r0 < MMQ_SEG_ROWS ? sg.rows - r0 : MMQ_SEG_ROWS;
    ds.staged = sg.d <= MMQ_SEG2_DMAX;
    if (ds.staged && lane == 0) {
      const uint32_t bytes = (uint32_t)(MMQ_SEG_ROWS * sg.d * 4); /* 256*d: a multiple of 16; the packed array has slack */
      mbar_expect_tx(&mbar[buf], HAS_W ? 2 * bytes : bytes);
      bulk_g2s(sc + buf * CAP, colp + ds.e0, bytes, &mbar[buf]);
      if (HAS_W) bulk_g2s(sw + buf * CAP, wp + ds.e0, bytes, &mbar[buf]);
    }
    return ds;
  };
  uint32_t uses0 = 0, uses1 = 0;
  Desc cur = describe(warp0, 0);
  int it = 0;
  for (int64_t chunk = warp0; chunk < total_chunks; chunk += nwarps, ++it) {
    const int buf = it & 1;
    Desc next = cur;
    if (chunk + nwarps < total_chunks) next = describe(chunk + nwarps, buf ^ 1);
    const int D = cur.d;
    const uint64_t cid = (uint64_t)(cur.cid0 + 2 * lane);
    uint32_t wd[4] = {(uint32_t)(cid >> 1), (uint32_t)(cid >> 33), sweep, 0u};
    mmq_philox4x32_10(wd, seed, MMQ_STREAM_CAT);
    const double ua = cat_u52(wd[0], wd[1]), ub = cat_u52(wd[2], wd[3]);
    int32_t ca = -1, cb = -1;
    if (cur.staged) {
      mbar_wait(&mbar[buf], (buf ? uses1 : uses0) & 1u);
      if (buf) ++uses1; else ++uses0;
      const int32_t* c0 = sc + buf * CAP;
      const float* w0 = sw + buf * CAP;
      if (D == 2) seg2_chunk<2, HAS_W>(c0, w0, sp, cur.row_lo, cur.row_hi, ua, ub, mu, sentinel, lane, ca, cb);
      else if (D == 3) seg2_chunk<3, HAS_W>(c0, w0, sp, cur.row_lo, cur.row_hi, ua, ub, mu, sentinel, lane, ca, cb);
      else if (D == 4) seg2_chunk<4, HAS_W>(c0, w0, sp, cur.row_lo, cur.row_hi, ua, ub, mu, sentinel, lane, ca, cb);
      else if (D == 5) seg2_chunk<5, HAS_W>(c0, w0, sp, cur.row_lo, cur.row_hi, ua, ub, mu, sentinel, lane, ca, cb);
      else seg2_chunk<6, HAS_W>(c0, w0, sp, cur.row_lo, cur.row_hi, ua, ub, mu, sentinel, lane, ca, cb);
    } else {
      const int ra = 2 * lane;
      const int64_t e = cur.e0 + (int64_t)ra * D;
      if (ra >= cur.row_lo && ra < cur.row_hi) ca = seg_row_generic<HAS_W>(colp + e, wp + e, D, mu, ua);
      if (ra + 1 < cur.row_hi) cb = seg_row_generic<HAS_W>(colp + e + D, wp + e + D, D, mu, ub);
    }
    cat_red(counts, ca, lane);
    cat_red(counts, cb, lane);
    __syncwarp();
    cur = next;
  }
}

/* ------------------------------------------------------------------ host */

int mmq_seg_add_base(mmq_handle* h, bool want_in_counts) {
  if (!h->seg_base || h->seg_base_in_counts == want_in_counts) return MMQ_OK;
  k_axpy_i32<<<mmq_grid_for(h->n, 256, h->num_sms * 8), 256, 0, h->stream>>>(h->counts, h->seg_base, h->n, want_in_counts ? 1 : -1);
  MMQ_LAUNCHED(h);
  h->seg_base_in_counts = want_in_counts;
  return MMQ_OK;
}

/* Host-only part of the plan: runs of equal class size.  Called right after the H2D copies are
 * queued, so the scan overlaps the DMA. */
int mmq_seg_scan(mmq_handle* h, const int64_t* rp) {
  h->seg_runs.clear();
  h->seg_scan_ok = false;
  const int64_t m = h->m;
  if (m == 0 || h->has_k) return MMQ_OK;
  for (int64_t i = 0; i < m;) {
    const int64_t d = rp[i + 1] - rp[i];
    if (d > 0x7fffffff || d <= 0) return MMQ_OK;
    int64_t j = i + 1;
    while (j < m && rp[j + 1] - rp[j] == d) ++j;
    h->seg_runs.push_back({i, j, rp[i], (int)d});
    if ((int)h->seg_runs.size() > MMQ_SEG_MAX) { h->seg_runs.clear(); return MMQ_OK; } /* ragged shard: the row-pointer kernel handles it */
    i = j;
  }
  h->seg_scan_ok = true;
  return MMQ_OK;
}

int mmq_seg_plan(mmq_handle* h) {
  h->seg_ready = false;
  if (!h->seg_scan_ok) return MMQ_OK;
  struct Run { int64_t r0, r1, q0; int d; };
  std::vector<Run> runs;
  for (const auto& r : h->seg_runs) runs.push_back({r.r0, r.r1, r.q0, r.d});
  std::vector<mmq_seg> segs;
  int64_t packed = 0, chunks = 0, entries = 0, rows = 0, singles = 0;
  for (const Run& r : runs) {
    if (r.d == 1) { singles += r.r1 - r.r0; continue; }
    if ((r.r1 - r.r0 + 1) > 0x7ffffff0ll) return MMQ_OK;
    mmq_seg sg;
    const int parity = (int)((h->class_id_base + r.r0) & 1);
    packed = (packed + 3) & ~(int64_t)3;
    sg.e_virtual = packed;
    sg.cid_virtual = h->class_id_base + r.r0 - parity;
    sg.row_lo = parity;
    sg.rows = (int32_t)(r.r1 - r.r0 + parity);
    sg.d = r.d;
    sg.pad_ = 0;
    sg.chunk0 = chunks;
    packed += (int64_t)sg.rows * r.d;
    chunks += (sg.rows + MMQ_SEG_ROWS - 1) / MMQ_SEG_ROWS;
    entries += (r.r1 - r.r0) * r.d;
    rows += r.r1 - r.r0;
    segs.push_back(sg);
  }
  if (segs.empty() && singles == 0) return MMQ_OK;
  packed = ((packed + 3) & ~(int64_t)3) + 64 * 8 + 64; /* slack: the bulk copy of a run's last chunk always moves 64 rows */
  int rc;
  if ((rc = mmq_dev_alloc(h, (void**)&h->seg_col, sizeof(int32_t) * (size_t)packed))) return rc;
  k_fill_i32<<<mmq_grid_for(packed, 256, h->num_sms * 8), 256, 0, h->stream>>>(h->seg_col, packed, (int32_t)h->n);
  MMQ_LAUNCHED(h);
  if (h->has_w) {
    if ((rc = mmq_dev_alloc(h, (void**)&h->seg_w, sizeof(float) * (size_t)packed))) return rc;
    MMQ_CUDA(h, cudaMemsetAsync(h->seg_w, 0, sizeof(float) * (size_t)packed, h->stream));
  }
  size_t si = 0;
  for (const Run& r : runs) {
    if (r.d == 1) continue;
    const mmq_seg& sg = segs[si++];
    const int64_t dst = sg.e_virtual + (int64_t)sg.row_lo * r.d;
    const size_t cnt = (size_t)(r.r1 - r.r0) * (size_t)r.d;
    MMQ_CUDA(h, cudaMemcpyAsync(h->seg_col + dst, h->col + r.q0, sizeof(int32_t) * cnt, cudaMemcpyDeviceToDevice, h->stream));
    if (h->has_w) MMQ_CUDA(h, cudaMemcpyAsync(h->seg_w + dst, h->w + r.q0, sizeof(float) * cnt, cudaMemcpyDeviceToDevice, h->stream));
  }
  if (singles > 0) {
    if ((rc = mmq_dev_alloc(h, (void**)&h->seg_base, sizeof(int32_t) * (size_t)h->n))) return rc;
    MMQ_CUDA(h, cudaMemsetAsync(h->seg_base, 0, sizeof(int32_t) * (size_t)h->n, h->stream));
    for (const Run& r : runs)
      if (r.d == 1) {
        k_count_singletons<<<mmq_grid_for(r.r1 - r.r0, 256, h->num_sms * 8), 256, 0, h->stream>>>(h->col, r.q0, r.q0 + (r.r1 - r.r0), h->seg_base);
        MMQ_LAUNCHED(h);
      }
    MMQ_CUDA(h, cudaMemcpyAsync(h->counts, h->seg_base, sizeof(int32_t) * (size_t)h->n, cudaMemcpyDeviceToDevice, h->stream));
    h->seg_base_in_counts = true;
  }
  if (!segs.empty()) {
    if ((rc = mmq_dev_alloc(h, &h->seg_table, sizeof(mmq_seg) * segs.size()))) return rc;
    MMQ_CUDA(h, cudaMemcpyAsync(h->seg_table, segs.data(), sizeof(mmq_seg) * segs.size(), cudaMemcpyHostToDevice, h->stream));
  }
  MMQ_CUDA(h, cudaStreamSynchronize(h->stream)); /* segs is a host temporary */
  h->seg_count = (int)segs.size();
  h->seg_chunks = chunks;
  h->seg_entries = entries;
  h->seg_rows = rows;
  h->seg_singletons = singles;
  h->seg_ready = true;
  return MMQ_OK;
}

int mmq_seg_launch(mmq_handle* h, uint32_t seed, uint32_t sweep, const uint32_t* sweep_base) {
  int rc = mmq_seg_add_base(h, true);
  if (rc) return rc;
  if (h->seg_count == 0) return MMQ_OK; /* only singletons: nothing random to do */
  static const int variant = [] { const char* e = getenv("MMQ_SEG_KERNEL"); return e ? atoi(e) : 1; }(); /* 1 row-parallel, 2 entry-parallel */
#define MMQ_SEG_ARGS (const mmq_seg*)h->seg_table, h->seg_count, h->seg_chunks, h->seg_col, h->seg_w, h->mu, h->counts, seed, sweep, (int32_t)h->n
#define MMQ_SEG_ARGS1 MMQ_SEG_ARGS, red_mode, dbg_dmin, dbg_dmax, sweep_base
  if (variant == 2) {
    const int64_t want2 = (h->seg_chunks + MMQ_SEG2_WARPS - 1) / MMQ_SEG2_WARPS;
    const int grid2 = (int)std::min<int64_t>(want2, (int64_t)h->num_sms * 3);
    constexpr int CAP = MMQ_SEG_ROWS * MMQ_SEG2_DMAX, PCAP = MMQ_SEG_ROWS * (MMQ_SEG2_DMAX | 1);
    if (h->has_w) {
      constexpr int SM = MMQ_SEG2_WARPS * (16 + PCAP * 8 + 2 * CAP * 4 * 2);
      MMQ_CUDA(h, cudaFuncSetAttribute(k_alloc_seg2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
      k_alloc_seg2<true><<<grid2, MMQ_SEG2_WARPS * 32, SM, h->stream>>>(MMQ_SEG_ARGS, sweep_base);
    } else {
      constexpr int SM = MMQ_SEG2_WARPS * (16 + PCAP * 8 + 2 * CAP * 4);
      MMQ_CUDA(h, cudaFuncSetAttribute(k_alloc_seg2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
      k_alloc_seg2<false><<<grid2, MMQ_SEG2_WARPS * 32, SM, h->stream>>>(MMQ_SEG_ARGS, sweep_base);
    }
    return MMQ_OK;
  }
  const int64_t want = (h->seg_chunks + MMQ_SEG_WARPS - 1) / MMQ_SEG_WARPS;
  static const int red_mode = [] { const char* e = getenv("MMQ_DEBUG_RED"); return e ? atoi(e) : 2; }(); /* 2 = aggregated (product) */
  static const int dbg_dmin = [] { const char* e = getenv("MMQ_DEBUG_DMIN"); return e ? atoi(e) : 0; }();
  static const int dbg_dmax = [] { const char* e = getenv("MMQ_DEBUG_DMAX"); return e ? atoi(e) : 0x7fffffff; }();
  static const int maxd_env = [] { const char* e = getenv("MMQ_SEG_MAXD"); return e ? atoi(e) : 0; }(); /* tuning knob */
  const int maxd = maxd_env ? maxd_env : (h->has_w ? 8 : 12); /* largest class size with a register-resident specialisation */
  static const int occ_env = [] { const char* e = getenv("MMQ_SEG_OCC"); return e ? atoi(e) : 0; }();
  const int occ = occ_env ? occ_env : (h->has_w ? 3 : 4);
  const int grid = (int)std::min<int64_t>(want, (int64_t)h->num_sms * occ);
#define MMQ_SEG_GO(W, MD, OC) k_alloc_seg<W, MD, OC><<<grid, MMQ_SEG_WARPS * 32, 0, h->stream>>>(MMQ_SEG_ARGS1)
  if (h->has_w) {
    if (maxd > 8) MMQ_SEG_GO(true, 12, 2);
    else if (maxd > 6) MMQ_SEG_GO(true, 8, 3);
    else if (maxd > 4) MMQ_SEG_GO(true, 6, 3);
    else if (occ == 4) MMQ_SEG_GO(true, 4, 4);
    else MMQ_SEG_GO(true, 4, 3);
  } else {
    if (maxd > 8) { if (occ == 2) MMQ_SEG_GO(false, 12, 2); else if (occ == 4) MMQ_SEG_GO(false, 12, 4); else if (occ == 5) MMQ_SEG_GO(false, 12, 5); else MMQ_SEG_GO(false, 12, 3); }
    else if (maxd > 6) MMQ_SEG_GO(false, 8, 3);
    else if (maxd > 4) MMQ_SEG_GO(false, 6, 3);
    else if (occ == 4) MMQ_SEG_GO(false, 4, 4);
    else MMQ_SEG_GO(false, 4, 3);
  }
#undef MMQ_SEG_GO
#undef MMQ_SEG_ARGS
#undef MMQ_SEG_ARGS1
  return MMQ_OK;
}
