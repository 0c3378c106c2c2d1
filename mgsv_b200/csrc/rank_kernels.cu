// Similarity -> top-k selection and ground-truth rank, replacing the full argsort + python walk
// of utils/util_test.py:32-97 (Recall_metrics, dedup=True) and the fp64 sum of
// test-MaDe.py:401-403.
//
// One CTA per query row.  Scores are s = double(single) + double(dual) exactly as the reference
// forms them (numpy float32 + float64 -> float64).  Work per row:
//   1. s* = max score over the ground-truth id's columns (chain walk over prev_same),
//   2. rank = number of DISTINCT music ids whose best column beats s* (strictly) — a column counts
//      iff s > s* and no earlier column of the same id also has s > s* (prev_same chain),
//   3. exact top-k, score descending, ties broken by the lower column index.
// Tie rule (documented difference): the reference orders by np.argsort(sim)[:, ::-1] — an unstable sort, reversed —
// so among EXACTLY equal float64 scores its order is implementation-defined (in practice the higher column first,
// and an id tied with the ground truth may be counted ahead of it).  Here ties go to the lower column and only
// strictly greater scores count for the rank.  Exact float64 ties need identical columns (a repeated track — which
// dedup merges into one id anyway) or saturated scores; tests compare indices modulo exactly tied scores.
//
// Staged kernel (rows up to 49 152 columns): the row is read ONCE from HBM and kept in shared memory
// as order-preserving 32-bit keys of fl32(single + dual).  Rounding is monotone, so a column that
// beats another in fp64 never has the smaller fp32 key: the fp32 keys select a SUPERSET of the true
// top-k and decide every rank comparison that is more than two fp32 steps away from s*; only the
// selected columns and the near-ties of s* are re-read (L2) and compared in fp64.  The candidate
// threshold comes from GROUP MAXIMA: while staging, every thread tracks the maximum key of its own
// columns; the k-th largest of those 256 (or 1024) maxima bounds the k-th largest key from below and
// leaves ~1.3 k columns at or above it, so a 1024-bin histogram of the maxima alone (1 - 4 shared-
// memory atomics per thread instead of one per column: ATOMS costs 2 cycles per lane, and the
// per-column histogram was 2/3 of the kernel) replaces the histogram of the row.  One sweep of the
// staged keys then counts the rank and gathers the candidates, which are ordered by counting.
// Rows where more than 512 columns pass that threshold (heavy ties) fall back to the per-column
// value histogram (refined 10 bits at a time while more than k + 32 keys sit at or above the chosen
// bin); rows where more than 512 columns share one fp32 key around rank k, and rows too long to
// stage, take the exact 8-pass MSB radix select on the 64-bit keys (re-reading global memory) + a
// bitonic sort instead.
#include <type_traits>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace made {

constexpr int kRankThreads = 256;
constexpr int kMaxK = 256;
constexpr int kMaxSmemCols = 49152;              // 49152 * 4 B = 192 KB of 32-bit keys
constexpr int kValueBinBits = 10;
constexpr int kValueBins = 1 << kValueBinBits;   // value-histogram bins of the staged path
constexpr int kMaxCand = 512;                    // candidates the staged path orders (>= kMaxK + kCandSlack)
constexpr int kCandSlack = 32;                   // refine the threshold while more than k + 32 keys pass it

// NaN scores (an all-zero pooled row divides 0 by 0, exactly as the reference does) order BELOW every number:
// key 0, so they never enter a top-k ahead of a real score and never count as beating the ground truth.
__device__ __forceinline__ unsigned long long f64_key(double x) {
  if (x != x) return 0ull;
  unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(x));
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double(static_cast<long long>(b));
}
__device__ __forceinline__ unsigned int f32_key(float x) {
  const unsigned int b = __float_as_uint(x);
  // negative: ~b, positive: b | 0x80000000 (= b ^ 0x80000000, the sign bit being clear): one shift, one three-input logic op
  const unsigned int key = b ^ (static_cast<unsigned int>(static_cast<int>(b) >> 31) | 0x80000000u);
  return x != x ? 0u : key;
}

// exact 64-bit keys of one row, read from global memory
struct RowView {
  const float* a;
  const float* b;  // may be null
  __device__ __forceinline__ unsigned long long key(int64_t j) const {
    double s = static_cast<double>(a[j]);
    if (b) s += static_cast<double>(b[j]);
    return f64_key(s);
  }
  __device__ __forceinline__ unsigned int key32(int64_t j) const {
    return f32_key(b ? __fadd_rn(a[j], b[j]) : a[j]);
  }
};

struct RankSmem {
  int hist[kValueBins];                     // value histogram / the 256 radix buckets
  int red[kRankThreads / 32];
  unsigned int red32[3][kRankThreads / 32];
  unsigned long long sel_keys[kMaxCand];
  int sel_idx[kMaxCand];
  unsigned long long prefix, gt_key;
  unsigned int gt_key32;
  int krem, count, eq_taken, bin, ncand;
};

__device__ __forceinline__ int block_sum_int(int v, int* red) {
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (kRankThreads / 32) ? red[threadIdx.x] : 0;
    t = __reduce_add_sync(0xffffffffu, t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

// bitonic sort of n (power of two) (key, idx) pairs: key descending, idx ascending.  Called by the whole CTA
// (blockDim.x a multiple of 32).  A compare-exchange step of stride <= 32 stays inside the 64-element segment
// [64 w, 64 w + 64) of the warp w = t / 32 that owns it, so consecutive narrow steps only need a warp barrier; the CTA
// barrier is kept where a step reads what other warps wrote (27 of the 28 steps of a 128-element sort are narrow).
__device__ void bitonic_sort_desc(unsigned long long* keys, int* idx, int n) {
  bool prev_wide = true;   // the input was written by arbitrary threads
  const bool idle = static_cast<int>(threadIdx.x & ~31u) >= n / 2;   // a warp without a compare-exchange in any step
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const bool wide = stride > 32;
      if (wide || prev_wide) __syncthreads(); else if (!idle) __syncwarp();
      prev_wide = wide;
      if (idle) continue;
      for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
        int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        int hi = lo + stride;
        bool desc_block = ((lo & size) == 0);
        unsigned long long kl = keys[lo], kh = keys[hi];
        int il = idx[lo], ih = idx[hi];
        // "lo should come first" when (key larger) or (equal and idx smaller)
        bool lo_first = (kl > kh) || (kl == kh && il < ih);
        if (lo_first != desc_block) {
          keys[lo] = kh; keys[hi] = kl;
          idx[lo] = ih; idx[hi] = il;
        }
      }
    }
  }
  __syncthreads();
}

// ground-truth score of the row: thread 0 walks the id's columns (exact fp64 keys)
__device__ __forceinline__ void gt_walk(const RowView& rv, int64_t row, int64_t n_cols, const int32_t* gt_col,
                                        const double* gt_score_in, const int32_t* prev_same, RankSmem& sm) {
  unsigned long long best = 0;  // smaller than any real key
  unsigned int best32 = 0;
  if (gt_score_in) {
    best = f64_key(gt_score_in[row]);
    best32 = f32_key(__double2float_rn(gt_score_in[row]));
  } else {
    int32_t g = gt_col ? gt_col[row] : -1;
    int32_t guard = 0;
    while (g >= 0 && g < n_cols && guard++ < (1 << 20)) {
      const unsigned long long kk = rv.key(g);
      if (kk > best) { best = kk; best32 = rv.key32(g); }
      g = prev_same ? prev_same[g] : -1;
    }
  }
  sm.gt_key = best;
  sm.gt_key32 = best32;
}

// Exact top-k of one row by an 8-pass MSB radix select on the 64-bit keys (global memory), ordered
// append of the ties at the k-th key, bitonic sort of the winners.  Called by the whole CTA.
__device__ void radix_topk_row(const RowView& rv, int64_t n_cols, int kk, int k, int64_t row, int32_t col_offset,
                               int32_t* __restrict__ topk_idx, double* __restrict__ topk_score, RankSmem& sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    sm.prefix = 0ull;
    sm.krem = kk;
    sm.count = 0;
    sm.eq_taken = 0;
  }
  __syncthreads();
  for (int pass = 7; pass >= 0; --pass) {
    sm.hist[threadIdx.x] = 0;  // kRankThreads == 256 buckets
    __syncthreads();
    const unsigned long long prefix = sm.prefix;
    const int shift = pass * 8;
    const unsigned long long himask = pass == 7 ? 0ull : (~0ull << (shift + 8));
    for (int64_t j = threadIdx.x; j < n_cols; j += kRankThreads) {
      unsigned long long key = rv.key(j);
      if ((key & himask) == prefix) atomicAdd(&sm.hist[(key >> shift) & 0xFF], 1);
    }
    __syncthreads();
    // find the bucket that holds the k-th largest key: warp 0 scans the 256 buckets from the top,
    // 8 buckets per lane + a shuffle scan (a single thread walking 256 buckets cost ~1.3 us per pass)
    if (threadIdx.x < 32) {
      const int rem = sm.krem;
      int loc[8], tot = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {          // lane 0 owns buckets 255..248, lane 1 247..240, ...
        loc[i] = sm.hist[255 - (lane * 8 + i)];
        tot += loc[i];
      }
      int incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      int acc = incl - tot;                   // keys in buckets above this lane's range
      const bool here = acc < rem && incl >= rem;      // the k-th largest falls inside this lane's 8 buckets
      const bool none = __ballot_sync(0xffffffffu, here) == 0u;   // fewer than `rem` keys in total: bucket 0
      if (here) {
        int b = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (acc + loc[i] >= rem) { b = 255 - (lane * 8 + i); break; }
          acc += loc[i];
        }
        sm.krem = rem - acc;  // still needed inside bucket b
        sm.prefix = prefix | (static_cast<unsigned long long>(b) << shift);
      } else if (none && lane == 31) {
        sm.krem = rem - (incl - loc[7]);
        sm.prefix = prefix;                   // bucket 0
      }
    }
    __syncthreads();
  }
  const unsigned long long thr = sm.prefix;  // exact k-th largest key
  const int need_eq = sm.krem;               // how many == thr to take (lowest indices first)
  // strictly-greater elements: unordered append
  for (int64_t j = threadIdx.x; j < n_cols; j += kRankThreads) {
    unsigned long long key = rv.key(j);
    if (key > thr) {
      int slot = atomicAdd(&sm.count, 1);
      if (slot < kMaxK) { sm.sel_keys[slot] = key; sm.sel_idx[slot] = static_cast<int>(j); }
    }
  }
  __syncthreads();
  const int n_gt = sm.count;
  // ties: ordered append, chunk by chunk
  for (int64_t base = 0; base < n_cols; base += kRankThreads) {
    if (sm.eq_taken >= need_eq) break;
    int64_t j = base + threadIdx.x;
    bool eq = j < n_cols && rv.key(j) == thr;
    unsigned ballot = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) sm.red[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kRankThreads / 32; ++w) {
      if (w < warp) before += sm.red[w];
      total += sm.red[w];
    }
    before += __popc(ballot & ((1u << lane) - 1u));
    int taken = sm.eq_taken;
    if (eq && taken + before < need_eq) {
      int slot = n_gt + taken + before;
      if (slot < kMaxK) { sm.sel_keys[slot] = thr; sm.sel_idx[slot] = static_cast<int>(j); }
    }
    __syncthreads();
    if (threadIdx.x == 0) sm.eq_taken = taken + total;
    __syncthreads();
  }
  // pad to a power of two and sort
  int n_pow = 1;
  while (n_pow < kk) n_pow <<= 1;
  for (int t = kk + threadIdx.x; t < n_pow; t += kRankThreads) {
    sm.sel_keys[t] = 0ull;
    sm.sel_idx[t] = 0x7FFFFFFF;
  }
  bitonic_sort_desc(sm.sel_keys, sm.sel_idx, n_pow);
  for (int t = threadIdx.x; t < k; t += kRankThreads) {
    if (t < kk) {
      topk_idx[row * k + t] = sm.sel_idx[t] + col_offset;
      if (topk_score) topk_score[row * k + t] = key_f64(sm.sel_keys[t]);
    } else {
      topk_idx[row * k + t] = -1;
      if (topk_score) topk_score[row * k + t] = -INFINITY;
    }
  }
}

// Rows too long to stage: every pass re-reads global memory (L2 for rows up to a few MB).
__global__ void __launch_bounds__(kRankThreads)
rank_topk_kernel(const float* __restrict__ single, const float* __restrict__ dual, int64_t ld,
                 int64_t n_cols, const int32_t* __restrict__ gt_col,
                 const double* __restrict__ gt_score_in, const int32_t* __restrict__ prev_same,
                 int32_t col_offset, int k, int32_t* __restrict__ topk_idx,
                 double* __restrict__ topk_score, int32_t* __restrict__ rank_out,
                 double* __restrict__ gt_score_out) {
  __shared__ RankSmem sm;
  const int64_t row = blockIdx.x;
  RowView rv;
  rv.a = single + row * ld;
  rv.b = dual ? dual + row * ld : nullptr;
  if (rank_out != nullptr) {
    if (threadIdx.x == 0) gt_walk(rv, row, n_cols, gt_col, gt_score_in, prev_same, sm);
    __syncthreads();
    const unsigned long long gk = sm.gt_key;
    int cnt = 0;
    for (int64_t j = threadIdx.x; j < n_cols; j += kRankThreads) {
      if (rv.key(j) > gk) {
        bool first = true;
        if (prev_same) {
          int32_t p = prev_same[j];
          int32_t guard = 0;
          while (p >= 0 && guard++ < (1 << 20)) {
            if (rv.key(p) > gk) { first = false; break; }
            p = prev_same[p];
          }
        }
        cnt += first ? 1 : 0;
      }
    }
    cnt = block_sum_int(cnt, sm.red);
    if (threadIdx.x == 0) {
      rank_out[row] = cnt;
      if (gt_score_out) gt_score_out[row] = gk ? key_f64(gk) : -INFINITY;
    }
  }
  if (k <= 0 || topk_idx == nullptr) return;
  const int kk = n_cols < k ? static_cast<int>(n_cols) : k;
  radix_topk_row(rv, n_cols, kk, k, row, col_offset, topk_idx, topk_score, sm);
}

// Warp 0 scans the value histogram from the top bin down (32 bins per lane + a shuffle scan) for the bin that holds
// the need-th largest item: sm.bin = that bin, sm.krem = items in the bins above it, sm.ncand = items in it.  sm.bin is
// left untouched when the histogram holds fewer than `need` items.
__device__ __forceinline__ void scan_hist_from_top(RankSmem& sm, int need, int lane) {
  constexpr int kPer = kValueBins / 32;
  int tot = 0;
  const int top = kValueBins - 1 - lane * kPer;   // this lane owns bins top, top-1, ..., top-kPer+1
#pragma unroll 8
  for (int i = 0; i < kPer; ++i) tot += sm.hist[top - ((i + lane) & (kPer - 1))];   // rotated: no bank conflicts
  int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  int acc = incl - tot;
  if (acc < need && incl >= need) {
    for (int i = 0; i < kPer; ++i) {
      const int c = sm.hist[top - i];
      if (acc + c >= need) { sm.bin = top - i; sm.krem = acc; sm.ncand = c; break; }
      acc += c;
    }
  }
}

// Staged rows: see the header of this file.  WIDE (rows of more than 8192 columns, whose 32+ KB of keys leave room for
// at most 3 - 6 CTAs per SM anyway): 3 CTAs per SM, 85 registers, EIGHT 16-byte word pairs in flight per thread while
// staging (64 KB per CTA; with four, a 128 KB row took four HBM round trips and staging was 40 % of the kernel).
// Otherwise 7 CTAs per SM: the job's own 2000 rows fit in two waves of 148 x 7 instead of three of 148 x 6.
template <bool WIDE>
__global__ void __launch_bounds__(kRankThreads, WIDE ? 3 : 7)
rank_topk_staged_kernel(const float* __restrict__ single, const float* __restrict__ dual, int64_t ld,
                        int64_t n_cols, const int32_t* __restrict__ gt_col,
                        const double* __restrict__ gt_score_in, const int32_t* __restrict__ prev_same,
                        int32_t col_offset, int k, int32_t* __restrict__ topk_idx,
                        double* __restrict__ topk_score, int32_t* __restrict__ rank_out,
                        double* __restrict__ gt_score_out, int use_group_maxima) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ RankSmem sm;
  unsigned int* cache = reinterpret_cast<unsigned int*>(dyn_smem);
  const int64_t row = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RowView rv;
  rv.a = single + row * ld;
  rv.b = dual ? dual + row * ld : nullptr;
  const bool want_rank = rank_out != nullptr;
  const bool want_topk = k > 0 && topk_idx != nullptr;
  const int kk = n_cols < k ? static_cast<int>(n_cols) : k;

  // ---- 1. ground-truth score.  Distinct ids with the score read from the row (the common call): every thread knows
  //         the ground-truth column and one of them re-reads it AFTER its staging loads (an L2 hit) - a walk by
  //         thread 0 (two dependent loads before its own staging starts) kept the whole CTA at the barrier below for a
  //         memory round trip.  Otherwise (repeated ids / score given): thread 0 walks. -----------------
  const bool gt_inline = want_rank && gt_score_in == nullptr && prev_same == nullptr;
  int g_col = -1;
  if (gt_inline) {
    g_col = gt_col ? gt_col[row] : -1;
    if (g_col < 0 || g_col >= n_cols) {
      g_col = -1;
      if (threadIdx.x == 0) { sm.gt_key = 0ull; sm.gt_key32 = 0u; }   // no ground truth in this row: every key beats it
    }
  } else if (want_rank && threadIdx.x == 0) {
    gt_walk(rv, row, n_cols, gt_col, gt_score_in, prev_same, sm);
  }
  auto gt_here = [&](float av, float bv) {    // called by one thread
    sm.gt_key = f64_key(rv.b ? static_cast<double>(av) + static_cast<double>(bv) : static_cast<double>(av));
    sm.gt_key32 = f32_key(rv.b ? __fadd_rn(av, bv) : av);
  };

  // ---- 2. stage the row as 32-bit keys, tracking the maximum of each of this thread's four column groups
  //         (group c of thread t = component c of the 16-byte words t, t + 256, ...: 1024 groups per row) --------
  uint4 gmax = make_uint4(0u, 0u, 0u, 0u);
  unsigned int kmax = 0u, imin = 0xFFFFFFFFu;    // largest key of the row, smallest non-empty group maximum
  {
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(single) & 15) == 0) &&
                     (!dual || (reinterpret_cast<uintptr_t>(dual) & 15) == 0);
    const int n4 = vec ? static_cast<int>(n_cols / 4) : 0;      // staged rows have <= 49152 columns: 32-bit indices
    const float4* a4 = reinterpret_cast<const float4*>(rv.a);
    const float4* b4 = reinterpret_cast<const float4*>(rv.b);
#pragma unroll (WIDE ? 8 : 4)
    for (int q = threadIdx.x; q < n4; q += kRankThreads) {
      float4 x = __ldg(a4 + q);
      if (rv.b) {
        const float4 y = __ldg(b4 + q);
        x.x = __fadd_rn(x.x, y.x); x.y = __fadd_rn(x.y, y.y); x.z = __fadd_rn(x.z, y.z); x.w = __fadd_rn(x.w, y.w);
      }
      const uint4 kq = make_uint4(f32_key(x.x), f32_key(x.y), f32_key(x.z), f32_key(x.w));
      *reinterpret_cast<uint4*>(cache + q * 4) = kq;
      gmax.x = max(gmax.x, kq.x); gmax.y = max(gmax.y, kq.y); gmax.z = max(gmax.z, kq.z); gmax.w = max(gmax.w, kq.w);
    }
    // the ground-truth column once more, by the thread that staged it (a test inside the loop above keeps the
    // compiler from batching its loads: measured 20 % slower on long rows): an L2 hit after the staging loads, under
    // the reductions below
    asm volatile("" ::: "memory");
    if (g_col >= 0 && static_cast<int>(threadIdx.x) == ((g_col >> 2) & (kRankThreads - 1)))
      gt_here(__ldg(rv.a + g_col), rv.b ? __ldg(rv.b + g_col) : 0.f);
    for (int j = n4 * 4 + threadIdx.x; j < static_cast<int>(n_cols); j += kRankThreads) {
      const unsigned int key = rv.key32(j);
      cache[j] = key;
      gmax.x = max(gmax.x, key);
    }
    kmax = max(max(gmax.x, gmax.y), max(gmax.z, gmax.w));
    // key 0 = an empty group (or one of NaN scores only): not a threshold item
    if (gmax.x != 0u) imin = gmax.x;
    if (gmax.y != 0u) imin = min(imin, gmax.y);
    if (gmax.z != 0u) imin = min(imin, gmax.z);
    if (gmax.w != 0u) imin = min(imin, gmax.w);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    imin = __reduce_min_sync(0xffffffffu, imin);
    if (lane == 0) { sm.red32[0][warp] = kmax; sm.red32[1][warp] = imin; }
  }
  for (int t = threadIdx.x; t < kValueBins; t += kRankThreads) sm.hist[t] = 0;
  if (threadIdx.x == 0) { sm.count = 0; sm.ncand = 0; sm.bin = -1; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kRankThreads / 32; ++w) {
    kmax = max(kmax, sm.red32[0][w]);
    imin = min(imin, sm.red32[1][w]);
  }
  const bool select = want_topk && kk > 0;

  // A column beats s* when its fp32 key is more than two steps above that of s* (then the fp64
  // scores differ too); within two steps the exact fp64 keys decide.
  const unsigned long long gk = want_rank ? sm.gt_key : 0ull;
  const unsigned int gk32 = want_rank ? sm.gt_key32 : 0u;
  // k32 > g_hi: more than two fp32 steps above s* (beats it); k32 < g_lo: more than two below (does not); saturated
  // bounds make the impossible side unreachable
  const unsigned int g_hi = gk32 > 0xFFFFFFFDu ? 0xFFFFFFFFu : gk32 + 2u;
  const unsigned int g_lo = gk32 < 2u ? 0u : gk32 - 2u;
  const unsigned int g_win = g_hi - g_lo;      // k32 - g_lo <= g_win (unsigned): within two steps of s*
  auto beats = [&](int j, unsigned int k32) -> bool {
    if (k32 > g_hi) return true;
    if (k32 < g_lo) return false;
    return rv.key(j) > gk;
  };
  const int nc = static_cast<int>(n_cols);
  const int nv = nc >> 2;                      // the row as 16-byte words of four keys (the staging buffer is 16-byte aligned)
  const uint4* cache4 = reinterpret_cast<const uint4*>(cache);
  // a top-k candidate: only its column is recorded here.  The exact fp64 keys are fetched (L2) by order_and_write, one
  // candidate per thread in ONE round trip: fetched here, a thread with three candidates spent three dependent round
  // trips on them while the other 255 waited at the barrier behind the sweep.
  auto take = [&](int j) {
    const int slot = atomicAdd(&sm.count, 1);
    if (slot < kMaxCand) sm.sel_idx[slot] = j;
  };
  constexpr std::true_type kYes{};
  constexpr std::false_type kNo{};
  // One sweep over the staged keys.  do_rank: count the distinct ids ahead of the ground truth; gather: append the
  // columns at or above the candidate threshold.  The common case of a key is two compares: the exact fp64 compare
  // (keys within two fp32 steps of s*) and the append (~k of the n columns) sit behind ONE rarely taken branch per
  // four keys.
  auto sweep = [&](auto do_rank_t, auto gather_t, unsigned int thr) {
    constexpr bool do_rank = decltype(do_rank_t)::value, gather = decltype(gather_t)::value;   // compile-time: no flag tests in the loops
    int cnt = 0;
    if (prev_same == nullptr) {
      auto rare = [&](int j, unsigned int k32) {
        if (do_rank && k32 - g_lo <= g_win && rv.key(j) > gk) ++cnt;
        if (gather && k32 >= thr) take(j);
      };
      // A 16-byte word with a rare key only sets a bit (its iteration number: rows of <= 49152 columns give a thread
      // <= 48 words); the words are revisited after the loop.  Rare per thread is not rare per warp - ~k of the n / 4
      // words hold a candidate, more than half of the warp iterations have one in some lane - and the revisit costs a
      // warp as many passes as its busiest lane has words (2 - 3) instead of one pass per such iteration.
      unsigned long long revisit = 0ull;
      int it = 0;
#pragma unroll 2
      for (int q = threadIdx.x; q < nv; q += kRankThreads, ++it) {
        const uint4 kq = cache4[q];
        if (do_rank) cnt += (kq.x > g_hi ? 1 : 0) + (kq.y > g_hi ? 1 : 0) + (kq.z > g_hi ? 1 : 0) + (kq.w > g_hi ? 1 : 0);
        const unsigned int near = min(min(kq.x - g_lo, kq.y - g_lo), min(kq.z - g_lo, kq.w - g_lo));
        const unsigned int top = max(max(kq.x, kq.y), max(kq.z, kq.w));
        const bool hit = (do_rank && near <= g_win) || (gather && top >= thr);
        revisit |= static_cast<unsigned long long>(hit ? 1u : 0u) << it;
      }
      while (revisit) {
        const int b = __ffsll(static_cast<long long>(revisit)) - 1;
        revisit &= revisit - 1ull;
        const int q = threadIdx.x + b * kRankThreads;
        const uint4 kq = cache4[q];
        rare(4 * q, kq.x); rare(4 * q + 1, kq.y); rare(4 * q + 2, kq.z); rare(4 * q + 3, kq.w);
      }
      for (int j = nv * 4 + threadIdx.x; j < nc; j += kRankThreads) {
        const unsigned int k32 = cache[j];
        if (do_rank && k32 > g_hi) ++cnt;
        rare(j, k32);
      }
    } else {
      for (int j = threadIdx.x; j < nc; j += kRankThreads) {
        const unsigned int k32 = cache[j];
        if (gather && k32 >= thr) take(j);
        if (do_rank && beats(j, k32)) {
          bool first = true;
          int32_t p = prev_same[j];
          int32_t guard = 0;
          while (p >= 0 && guard++ < (1 << 20)) {
            if (beats(p, cache[p])) { first = false; break; }
            p = prev_same[p];
          }
          cnt += first ? 1 : 0;
        }
      }
    }
    if (do_rank) {
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0) sm.red[warp] = cnt;
    }
    __syncthreads();                        // ONE barrier publishes the candidates and the warps' counts
    if (do_rank && threadIdx.x == 0) {
      int tot = 0;
#pragma unroll
      for (int w = 0; w < kRankThreads / 32; ++w) tot += sm.red[w];
      rank_out[row] = tot;
      if (gt_score_out) gt_score_out[row] = gk ? key_f64(gk) : -INFINITY;
    }
  };
  // Order the n_cand <= kMaxCand gathered candidates (score descending, lower column first) and write the best kk: every
  // candidate's position is the number of candidates that come before it.  The count runs on the HIGH words of the fp64
  // keys (sign, exponent, 20 mantissa bits: one broadcast 4-byte shared-memory load and two compares per pair), split
  // over 256 / n threads per candidate; only a candidate that shares its high word with another one (a relative
  // score difference below 1e-6) recounts with the full keys and the column tie rule.  Two barriers instead of the
  // 28 dependent compare-exchange steps of a 128-element bitonic sort (a quarter of the kernel on long rows).
  auto order_and_write = [&](int n_cand) {
    for (int t = threadIdx.x; t < n_cand; t += kRankThreads) sm.sel_keys[t] = rv.key(sm.sel_idx[t]);
    __syncthreads();
    int n_pow = 32;
    while (n_pow < n_cand) n_pow <<= 1;
    const int parts = n_pow <= kRankThreads ? kRankThreads / n_pow : 1;   // n_pow >= 32: the lanes of a warp share one part
    const int part = parts > 1 ? static_cast<int>(threadIdx.x) / n_pow : 0;
    const int per = (n_cand + parts - 1) / parts;
    const int u0 = part * per, u1 = min(n_cand, u0 + per);
    const uint2* k2 = reinterpret_cast<const uint2*>(sm.sel_keys);   // .y = high word (little endian)
    int* partial = sm.hist;                                            // [parts][n_pow] (<= 512 entries), free after the sweep
    for (int c = static_cast<int>(threadIdx.x) & (n_pow - 1); c < n_pow; c += kRankThreads) {
      int gt = 0, eq = 0;
      if (c < n_cand) {
        const unsigned int khi = k2[c].y;
#pragma unroll 8
        for (int u = u0; u < u1; ++u) {
          const unsigned int h = k2[u].y;
          gt += h > khi ? 1 : 0;
          eq += h == khi ? 1 : 0;
        }
      }
      partial[part * n_pow + c] = gt | (eq << 16);
    }
    __syncthreads();
    if (part == 0) {
      for (int c = static_cast<int>(threadIdx.x) & (n_pow - 1); c < n_cand; c += kRankThreads) {
        int acc = 0;
        for (int p = 0; p < parts; ++p) acc += partial[p * n_pow + c];
        int pos = acc & 0xFFFF;
        const unsigned long long key = sm.sel_keys[c];
        if ((acc >> 16) != 1) {      // another candidate shares the high word: exact recount (larger key, or equal key and lower column)
          const int idx = sm.sel_idx[c];
          pos = 0;
          for (int u = 0; u < n_cand; ++u) {
            const unsigned long long ku = sm.sel_keys[u];
            pos += (ku > key || (ku == key && sm.sel_idx[u] < idx)) ? 1 : 0;
          }
        }
        if (pos < kk) {
          topk_idx[row * k + pos] = sm.sel_idx[c] + col_offset;
          if (topk_score) topk_score[row * k + pos] = key_f64(key);
        }
      }
    }
    for (int t = kk + threadIdx.x; t < k; t += kRankThreads) {
      topk_idx[row * k + t] = -1;
      if (topk_score) topk_score[row * k + t] = -INFINITY;
    }
  };

  // ---- 3. candidate threshold from the group maxima ------------------------------------------------------------
  // The kk-th largest of the G = 1024 group maxima is a lower bound of the kk-th largest key (the kk groups at or
  // above it hold kk distinct columns), and about -G ln(1 - kk / G) keys of a row lie at or above it (105 for
  // kk = 100): locating it takes a histogram of G items instead of n_cols, and ONE sweep then serves the rank and
  // gathers the candidates.  The threshold is the low edge of the value bin that holds that maximum; the 1024 bins
  // span the maxima only (a fraction of the row's key range: float keys are logarithmic in the value), refined 10 bits
  // at a time while more than 8 maxima share the bin.
  bool rank_done = false;
  if (select && use_group_maxima && imin <= kmax) {
    const unsigned int gm[4] = {gmax.x, gmax.y, gmax.z, gmax.w};
    const unsigned int mrange = kmax - imin;
    const int mbits = mrange ? 32 - __clz(static_cast<int>(mrange)) : 0;
    unsigned int base = imin, thr_g = 0u;
    int shift = mbits > kValueBinBits ? mbits - kValueBinBits : 0, need = kk;
    unsigned int span = kValueBins;
    bool have = false;
    for (int level = 0; level < 4; ++level) {
      if (level > 0) {
        for (int t = threadIdx.x; t < kValueBins; t += kRankThreads) sm.hist[t] = 0;
        __syncthreads();
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // leaving a group out (key 0, see above) only lowers the threshold
        if (gm[i] != 0u && gm[i] >= base) {
          const unsigned int d = (gm[i] - base) >> shift;
          if (d < span) atomicAdd(&sm.hist[d], 1);
        }
      }
      __syncthreads();
      if (warp == 0) scan_hist_from_top(sm, need, lane);
      __syncthreads();
      const int bin = sm.bin, acc = sm.krem, c = sm.ncand;
      if (bin < 0) break;                       // level 0 only: fewer than kk non-empty groups (a short row)
      thr_g = base + (static_cast<unsigned int>(bin) << shift);
      have = true;
      if (c <= 8 || shift == 0) break;
      need -= acc;
      base = thr_g;
      span = 1u << (shift > kValueBinBits ? kValueBinBits : shift);
      shift = shift > kValueBinBits ? shift - kValueBinBits : 0;
    }
    if (have) {
      if (want_rank) sweep(kYes, kYes, thr_g); else sweep(kNo, kYes, thr_g);
      const int n_cand = sm.count;
      if (n_cand <= kMaxCand) {
        order_and_write(n_cand);
        return;
      }
      rank_done = true;                         // too many columns at the threshold (heavy ties): the histogram path decides
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kValueBins; t += kRankThreads) sm.hist[t] = 0;
    if (threadIdx.x == 0) sm.count = 0;
    __syncthreads();
  }

  // ---- 4. rows the group maxima do not settle: the rank sweep, then the per-column value histogram ---------------
  if (want_rank && !rank_done) sweep(kYes, kNo, 0u);
  if (!want_topk) return;

  // ---- 5. threshold by value, 10 bits of the key range per level -------------------------------
  bool ok = false;
  unsigned int thr = 0u;
  int n_cand = 0;
  if (select) {
    unsigned int kmin = 0xFFFFFFFFu;
    for (int q = threadIdx.x; q < nv; q += kRankThreads) {
      const uint4 kq = cache4[q];
      kmin = min(kmin, min(min(kq.x, kq.y), min(kq.z, kq.w)));
    }
    for (int j = nv * 4 + threadIdx.x; j < nc; j += kRankThreads) kmin = min(kmin, cache[j]);
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    __syncthreads();                                         // red32 / bin of the sections above are no longer read
    if (lane == 0) sm.red32[2][warp] = kmin;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kRankThreads / 32; ++w) kmin = min(kmin, sm.red32[2][w]);
    const unsigned int range = kmax - kmin;
    const int bits = range ? 32 - __clz(static_cast<int>(range)) : 0;
    const int vshift = bits > kValueBinBits ? bits - kValueBinBits : 0;   // (range >> vshift) < kValueBins
    auto fill = [&](unsigned int k32) { atomicAdd(&sm.hist[(k32 - kmin) >> vshift], 1); };
    for (int q = threadIdx.x; q < nv; q += kRankThreads) {
      const uint4 kq = cache4[q];
      fill(kq.x); fill(kq.y); fill(kq.z); fill(kq.w);
    }
    for (int j = nv * 4 + threadIdx.x; j < nc; j += kRankThreads) fill(cache[j]);
    __syncthreads();
    unsigned int base = kmin;
    int shift = vshift, need = kk, above = 0;
    unsigned int span = kValueBins;   // sub-bins the bin chosen at the previous level splits into
    for (int level = 0; level < 4; ++level) {
      if (level > 0) {
        for (int t = threadIdx.x; t < kValueBins; t += kRankThreads) sm.hist[t] = 0;
        __syncthreads();
        auto refine = [&](unsigned int k32) {
          const unsigned int d = (k32 - base) >> shift;
          if (k32 >= base && d < span) atomicAdd(&sm.hist[d], 1);   // keys of the chosen bin only
        };
        for (int q = threadIdx.x; q < nv; q += kRankThreads) {
          const uint4 kq = cache4[q];
          refine(kq.x); refine(kq.y); refine(kq.z); refine(kq.w);
        }
        for (int j = nv * 4 + threadIdx.x; j < nc; j += kRankThreads) refine(cache[j]);
        __syncthreads();
      }
      if (warp == 0) scan_hist_from_top(sm, need, lane);
      __syncthreads();
      const int bin = sm.bin, acc = sm.krem, c = sm.ncand;
      n_cand = above + acc + c;
      // accept once the candidate set is close to k; at shift 0 the bin is a single key value and cannot be split further
      if (n_cand <= kk + kCandSlack || (shift == 0 && n_cand <= kMaxCand)) {
        thr = base + (static_cast<unsigned int>(bin) << shift);
        ok = true;
        break;
      }
      if (shift == 0) break;   // more than kMaxCand columns share one fp32 key: the radix path orders them
      above += acc;
      need -= acc;
      base += static_cast<unsigned int>(bin) << shift;
      span = 1u << (shift > kValueBinBits ? kValueBinBits : shift);
      shift = shift > kValueBinBits ? shift - kValueBinBits : 0;
    }
  }
  if (ok) {
    // ---- 6. gather the candidates with their exact fp64 keys, order them ----------
    sweep(kNo, kYes, thr);
    order_and_write(n_cand);
    return;
  }
  __syncthreads();
  radix_topk_row(rv, n_cols, kk, k, row, col_offset, topk_idx, topk_score, sm);
}

// Merge G per-shard candidate lists per row ([n_rows, n_cand] scores + global indices, -1 = empty)
// into the global top-k: score descending, index ascending.
__global__ void __launch_bounds__(kRankThreads)
topk_merge_kernel(const double* __restrict__ cand_score, const int32_t* __restrict__ cand_idx,
                  int n_cand, int k, int32_t* __restrict__ out_idx, double* __restrict__ out_score) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  int n_pow = 1;
  while (n_pow < n_cand) n_pow <<= 1;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(dyn_smem);
  int* idx = reinterpret_cast<int*>(keys + n_pow);
  const int64_t row = blockIdx.x;
  for (int t = threadIdx.x; t < n_pow; t += kRankThreads) {
    if (t < n_cand && cand_idx[row * n_cand + t] >= 0) {
      keys[t] = f64_key(cand_score[row * n_cand + t]);
      idx[t] = cand_idx[row * n_cand + t];
    } else {
      keys[t] = 0ull;
      idx[t] = 0x7FFFFFFF;
    }
  }
  bitonic_sort_desc(keys, idx, n_pow);
  for (int t = threadIdx.x; t < k; t += kRankThreads) {
    bool ok = t < n_pow && idx[t] != 0x7FFFFFFF;
    out_idx[row * k + t] = ok ? idx[t] : -1;
    out_score[row * k + t] = ok ? key_f64(keys[t]) : -INFINITY;
  }
}

// dual-tower cosine (modules/loss.py:52-56): out[i,j] = <a_i/|a_i|, b_j/|b_j|>, all fp32.
// 64x64 output tile per CTA, 256 threads, 4x4 outputs per thread, D = 256 staged in two halves.
constexpr int kCosTile = 64;
__global__ void __launch_bounds__(256)
cosine_sim_kernel(const float* __restrict__ a, int64_t n, const float* __restrict__ b, int64_t m,
                  int d, float* __restrict__ out, int64_t ld) {
  __shared__ float sa[kCosTile][33], sb[kCosTile][33];
  __shared__ float na[kCosTile], nb[kCosTile];
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * kCosTile;
  const int64_t j0 = static_cast<int64_t>(blockIdx.x) * kCosTile;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  // row norms (one warp handles 8 rows of each side)
  {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < 2 * kCosTile; r += 8) {
      bool is_a = r < kCosTile;
      int rr = is_a ? r : r - kCosTile;
      int64_t g = (is_a ? i0 : j0) + rr;
      int64_t lim = is_a ? n : m;
      const float* p = (is_a ? a : b) + g * d;
      float s = 0.f;
      if (g < lim)
        for (int c = lane; c < d; c += 32) { float v = p[c]; s = fmaf(v, v, s); }
      s = warp_sum(s);
      if (lane == 0) (is_a ? na : nb)[rr] = sqrtf(s);
    }
  }
  float acc[4][4] = {};
  for (int c0 = 0; c0 < d; c0 += 32) {
    __syncthreads();
    for (int t = threadIdx.x; t < kCosTile * 32; t += 256) {
      int r = t / 32, c = t % 32;
      int64_t gi = i0 + r, gj = j0 + r;
      // normalise on load: x / |x| (one rounding, like the reference's elementwise divide)
      sa[r][c] = (gi < n && c0 + c < d) ? __fdiv_rn(a[gi * d + c0 + c], na[r]) : 0.f;
      sb[r][c] = (gj < m && c0 + c < d) ? __fdiv_rn(b[gj * d + c0 + c], nb[r]) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      float av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { av[u] = sa[ty * 4 + u][c]; bv[u] = sb[tx * 4 + u][c]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      int64_t gi = i0 + ty * 4 + u, gj = j0 + tx * 4 + v;
      if (gi < n && gj < m) out[gi * ld + gj] = acc[u][v];
    }
}

// Tensor-core route of the dual-tower cosine for 256-d embeddings: each L2-normalised fp32 row is
// split into an fp16 (hi, lo) pair, x = hi + lo to ~2^-22, and  <x, y> ~= hi.hi' + hi.lo' + lo.hi'
// is ONE K = 768 GEMM of [hi | hi | lo] against [hi' | lo' | hi'] with fp32 accumulation
// (error ~1e-7, the dropped lo.lo' term is ~2^-24).  Warp per row.
__global__ void cos_split_kernel(const float* __restrict__ x, int64_t rows, int is_track, op_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[8];
  const float4* p = reinterpret_cast<const float4*>(x + row * 256 + lane * 8);
  const float4 a = p[0], b = p[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s = fmaf(v[j], v[j], s);
  const float nrm = sqrtf(warp_sum(s));
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float x0 = __fdiv_rn(v[2 * j], nrm), x1 = __fdiv_rn(v[2 * j + 1], nrm);
    const op2_t h = floats2op2(x0, x1);
    const float2 hf = op2_to_f2(h);
    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
    lo[j] = pack_op2(x0 - hf.x, x1 - hf.y);
  }
  const uint4 H = make_uint4(hi[0], hi[1], hi[2], hi[3]), L = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  uint4* o = reinterpret_cast<uint4*>(out + row * 768 + lane * 8);
  o[0] = H;                          // columns   0..255
  o[32] = is_track ? L : H;          // columns 256..511
  o[64] = is_track ? H : L;          // columns 512..767
}

}  // namespace made

using namespace made;

extern "C" {

int made_rank_topk(const float* single, const float* dual, int64_t ld, int64_t n_rows,
                   int64_t n_cols, const int32_t* gt_col, const double* gt_score_in,
                   const int32_t* prev_same, int32_t col_offset, int k, int32_t* topk_idx,
                   double* topk_score, int32_t* rank_out, double* gt_score_out, void* stream) {
  if (n_rows == 0) return MADE_OK;
  MADE_REQUIRE(single, "rank_topk: null similarity matrix");
  MADE_REQUIRE(k >= 0 && k <= kMaxK, "rank_topk: k=%d outside [0,%d]", k, kMaxK);
  MADE_REQUIRE(n_cols > 0 && n_cols < (1LL << 31), "rank_topk: n_cols=%lld unsupported",
               (long long)n_cols);
  MADE_REQUIRE(k == 0 || topk_idx, "rank_topk: k>0 needs topk_idx");
  MADE_REQUIRE(!rank_out || gt_col || gt_score_in, "rank_topk: rank needs gt_col or gt_score_in");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof_scope(kProfRank, st);
  if (n_cols <= kMaxSmemCols) {
    const bool wide = n_cols > 8192;
    MADE_TRY(ensure_dynamic_smem(wide ? reinterpret_cast<const void*>(&rank_topk_staged_kernel<true>)
                                      : reinterpret_cast<const void*>(&rank_topk_staged_kernel<false>),
                                 static_cast<int>(kMaxSmemCols * 4)));
    const size_t smem = (static_cast<size_t>(n_cols) * 4 + 15) & ~static_cast<size_t>(15);
    // MADE_RANK_GROUP_MAXIMA=0: the full value histogram decides every row (A/B switch; results are identical)
    const char* gmx = getenv("MADE_RANK_GROUP_MAXIMA");
    const int use_group_maxima = (gmx && gmx[0] == '0') ? 0 : 1;
    if (wide)
      rank_topk_staged_kernel<true><<<static_cast<unsigned>(n_rows), kRankThreads, smem, st>>>(
          single, dual, ld, n_cols, gt_col, gt_score_in, prev_same, col_offset, k, topk_idx, topk_score, rank_out,
          gt_score_out, use_group_maxima);
    else
      rank_topk_staged_kernel<false><<<static_cast<unsigned>(n_rows), kRankThreads, smem, st>>>(
          single, dual, ld, n_cols, gt_col, gt_score_in, prev_same, col_offset, k, topk_idx, topk_score, rank_out,
          gt_score_out, use_group_maxima);
  } else {
    rank_topk_kernel<<<static_cast<unsigned>(n_rows), kRankThreads, 0, st>>>(
        single, dual, ld, n_cols, gt_col, gt_score_in, prev_same, col_offset, k, topk_idx, topk_score, rank_out,
        gt_score_out);
  }
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int made_topk_merge(const double* cand_score, const int32_t* cand_idx, int64_t n_rows, int n_cand,
                    int k, int32_t* out_idx, double* out_score, void* stream) {
  if (n_rows == 0) return MADE_OK;
  MADE_REQUIRE(cand_score && cand_idx && out_idx && out_score, "topk_merge: null pointer");
  MADE_REQUIRE(n_cand > 0 && n_cand <= 4096 && k > 0 && k <= n_cand,
               "topk_merge: n_cand=%d k=%d unsupported", n_cand, k);
  int n_pow = 1;
  while (n_pow < n_cand) n_pow <<= 1;
  size_t smem = static_cast<size_t>(n_pow) * 12;
  MADE_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&topk_merge_kernel), static_cast<int>(4096 * 12)));
  topk_merge_kernel<<<static_cast<unsigned>(n_rows), kRankThreads, smem,
                      static_cast<cudaStream_t>(stream)>>>(cand_score, cand_idx, n_cand, k, out_idx,
                                                           out_score);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int made_cosine_sim(const float* a, int64_t n, const float* b, int64_t m, int d, float* out,
                    int64_t ld, void* stream) {
  if (n == 0 || m == 0) return MADE_OK;
  MADE_REQUIRE(a && b && out && d > 0 && ld >= m, "cosine_sim: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t m_tc = 0;
  // One arithmetic for every gallery width AND every output placement: narrow calls (a last score chunk, the
  // paired-track scores of the gallery index) and calls into a matrix whose row pitch or column offset is not a
  // multiple of 4 floats (a gallery shard cut between two music ids) must give the bits of the wide aligned call —
  // the sharded path is compared with the single-GPU path bit for bit.  The TMA store needs 16-byte alignment, so an
  // unaligned destination is served through an aligned scratch matrix and one strided device copy.
  if (d == 256 && gemm_tma_store_enabled()) {
    const bool aligned = (ld % 4) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const int64_t ld_tc = aligned ? ld : ((m + 3) / 4) * 4;
    // tcgen05 route: every column when outputs leave through TMA stores (the last 256-wide tile reads
    // zero rows past the gallery and its store is clipped at column m), else the whole tiles only
    m_tc = gemm_tma_store_enabled() ? m : (m / 256) * 256;
    // stream-ordered scratch: keep freed blocks in the pool across synchronisations (the default
    // release threshold of 0 hands them back to the OS at every sync, which turned the next
    // cudaMallocAsync into a multi-millisecond allocation whenever a step ended with a sync)
    static thread_local int pool_ready_dev = -1;
    int dev = 0;
    MADE_CUDA(cudaGetDevice(&dev));
    if (pool_ready_dev != dev) {
      cudaMemPool_t pool;
      MADE_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
      uint64_t keep = UINT64_MAX;
      MADE_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      pool_ready_dev = dev;
    }
    op_t *a16 = nullptr, *b16 = nullptr;
    float* out_tc = out;
    if (!aligned)
      MADE_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&out_tc), static_cast<size_t>(n) * ld_tc * sizeof(float), st));
    MADE_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&a16), static_cast<size_t>(n) * 768 * 2, st));
    MADE_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&b16), static_cast<size_t>(m_tc) * 768 * 2, st));
    cos_split_kernel<<<static_cast<unsigned>(ceil_div64(n, 8)), 256, 0, st>>>(a, n, 0, a16);
    MADE_CHECK_LAUNCH();
    cos_split_kernel<<<static_cast<unsigned>(ceil_div64(m_tc, 8)), 256, 0, st>>>(b, m_tc, 1, b16);
    MADE_CHECK_LAUNCH();
    GemmParams p;
    p.M = n;
    p.N = static_cast<int>(ceil_div64(m_tc, 256) * 256);
    p.n_store = static_cast<int>(m_tc);
    p.K = 768;
    p.epi.out_f32 = out_tc;
    p.epi.ld_f32 = ld_tc;
    int rc = gemm_f16_tc(a16, 768, b16, 768, m_tc, p, 256, st);
    MADE_CUDA(cudaFreeAsync(a16, st));
    MADE_CUDA(cudaFreeAsync(b16, st));
    if (!aligned) {
      if (rc == MADE_OK)
        MADE_CUDA(cudaMemcpy2DAsync(out, static_cast<size_t>(ld) * sizeof(float), out_tc, static_cast<size_t>(ld_tc) * sizeof(float),
                                    static_cast<size_t>(m) * sizeof(float), static_cast<size_t>(n), cudaMemcpyDeviceToDevice, st));
      MADE_CUDA(cudaFreeAsync(out_tc, st));
    }
    MADE_TRY(rc);
  }
  if (m_tc < m) {   // remaining columns (and every non-256-d call): fp32 SIMT tiles
    dim3 grid(static_cast<unsigned>(ceil_div64(m - m_tc, kCosTile)), static_cast<unsigned>(ceil_div64(n, kCosTile)));
    cosine_sim_kernel<<<grid, 256, 0, st>>>(a, n, b + m_tc * d, m - m_tc, d, out + m_tc, ld);
    MADE_CHECK_LAUNCH();
  }
  return MADE_OK;
}

}  // extern "C"
