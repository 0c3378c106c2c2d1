// Fused X-Pool scoring: for every (query video v, gallery track m) pair compute
//   sim_single[v,m] = cos( v_hat,  LN3( a + linear_proj(a) ) ),  a = LN2( out_proj( softmax_t(q.K_t/16) V ) )
// i.e. Transformer_XA.forward (modules/transformer.py:156-180, 87-123) followed by
// sim_matrix_music_pooling (modules/metrics.py:10-24), WITHOUT materialising the reference's
// [N_m, N_v, 256] pooled tensor (8.2 GB fp32 at 2k x 4k).
//
// Algebra (all exact in real arithmetic; DESIGN.md "X-Pool folding").  With e = exp(S - max) the
// un-normalised attention weights of a pair, l = sum(e):
//   o - mean(o)          = (1/l) sum_t e_t V''_t      V'' = centred (Wo Wv) LN1(x) + centred biases
//   |o - mean(o)|^2      = e^T G e / l^2              G   = V'' V''^T   (per track, 96x96)
//   LN2 -> (I+Wl) + bl   = alpha Y + b',  Y = e Z'',  alpha = 1 / (l sigma2),  Z'' = V'' W'^T
// and LN3 + the cosine with v_hat only need SUMS over the 256 features of alpha*Y + b':
//   linear in Y with constant weights  (sum Y, sum b'Y, sum g^2 Y, sum g^2 b'Y, sum g*beta Y)
//        = e . w_c with w_c = Z'' c per track  ->  extra columns "W5" next to G (each w_c as an fp16
//          hi/lo pair), so the tensor core produces them in the same MMA as T = e G;
//   quadratic / per-query terms        (sum Y^2, sum g^2 Y^2, sum u Y with u = v_hat*g)
//        = one sweep over the Y accumulator, 4 flops per element.
// Per pair the MMAs are S = q K^T (96), [T | L] = e [G | W5] (112) and Y = e Z'' (256):
// 119 808 executed flops instead of the reference's 360 960.
//
// One CTA owns a 128-query tile (Q resident in shared memory, u resident in TMEM as fp16) and
// streams tracks: TMA brings K / Z'' / [G|W5] of a track into 128B-swizzled shared memory, one thread
// issues tcgen05.mma (S: K-major B; T, Y: MN-major B straight from the row-major [token, feature]
// matrices), and 8 epilogue warps (two per TMEM lane quarter, each owning half of the columns of a
// query row) do softmax -> P (swizzled smem A operand), the quadratic form, the Y sweep and the
// closed-form LN2/LN3/cosine.  GEMM1 of track i+1 overlaps the Y sweep of track i.
//
// Work follows the track's VALID length: nblk = number of 16-segment blocks up to the last valid segment (from the
// mask bits; 13..96 valid segments, 56 on average -> 4.0 of 6 blocks).  S is issued with N = 16 nblk, [T|L] and Y
// with nblk K-steps, and the two column halves of the epilogue split the nblk blocks between them (ceil / floor),
// so masked tails cost neither tensor time nor softmax work.  Masked segments inside the covered blocks (holes,
// the partial last block) are excluded exactly: e = 0.  The softmax denominator l = sum of the fp16-ROUNDED weights
// comes out of the [T|L] MMA through a ones column (106) of the [G | W5 | 1] operand.  The Y sweep runs on packed
// fp32 pairs (FFMA2 / FMUL2: one issue slot per two features).
#include <cstdlib>

#include "common.cuh"
#include "prep.cuh"

namespace made {

constexpr int kXQ = 128;     // queries per tile
constexpr int kXL = 96;      // segments per track
constexpr int kXD = 256;
constexpr int kXG = 112;     // columns of the per-track [G | W5 | 0] operand
constexpr int kXThreads = 384;
constexpr uint32_t kQBytes = kXQ * kXD * 2;          // 64 KB: 4 k-slabs of [128 x 128B]
constexpr uint32_t kKBytes = kXL * kXD * 2;          // 48 KB: 4 k-slabs of [96 x 128B]
constexpr uint32_t kZBytes = kXL * kXD * 2;          // 48 KB: 4 n-slabs of [96 x 128B]
constexpr uint32_t kGBytes = 2 * kXL * 128;          // 24 KB: 2 n-slabs of [96 x 128B]
constexpr uint32_t kPBytes = 2 * kXQ * 128;          // 32 KB: 2 k-slabs of [128 x 128B]
constexpr uint32_t kSlabQ = kXQ * 128;               // 16 KB
constexpr uint32_t kSlabT = kXL * 128;               // 12 KB
constexpr uint32_t kXchgBytes = 2 * kXQ * 4 + 2 * kXQ * 4 * 4;   // max per (half, row) + 4 partial sums per (parity, row)
constexpr uint32_t kXSmem = kQBytes + kKBytes + kZBytes + kGBytes + kPBytes + 1024 + 256 + kXchgBytes;
static_assert(kXSmem <= 232448, "shared memory budget of an sm_100 CTA");

// TMEM columns
constexpr uint32_t kColY = 0;      // 256 fp32 columns
constexpr uint32_t kColT = 256;    // 112 fp32 columns [T | L]; also holds S before the softmax
constexpr uint32_t kColV = 384;    // 128 columns: u = v_hat * gamma3 as packed fp16 pairs

// XpoolConsts (prep.cuh) travels as a __grid_constant__ kernel parameter: parameters live in constant bank 0,
// so gamma3^2 stays an immediate constant-bank FFMA operand in the Y sweep, and every made_ctx has its own copy.

// Tensor maps: the track operands have one map per valid-block count (box of 16 nblk rows), so a track's masked tail
// is never fetched from L2 (the kernel moves ~2 GB from L2 to shared memory per 2000 x 1000 launch otherwise).
struct XpoolMaps {
  CUtensorMap q;
  CUtensorMap k[6], z[6], g[6];
};

struct XpoolParams {
  int64_t n_queries, n_tracks;
  int q_tiles, slices;
  const float* vhat;           // [n_queries, 256] fp32: v / |v| (modules/metrics.py:19)
  const uint32_t* maskbits;    // [n_tracks, 4]
  float* sim;                  // [n_queries, ld]
  int64_t ld;
  int64_t col_offset;
  float ln2_eps, ln3_eps;
  long long* trace;            // diagnostics: clock64 stamps of CTA 0 (scripts/diag_xpool_trace.py), null in production
  int debug;                   // bench-only ablations (MADE_XPOOL_DEBUG): 1 no Y sweep, 2 no exp, 4 no quadratic form,
                               // 8 no Y MMA, 16 no S MMA, 32 no T MMA (results are wrong with any bit set)
};

// packed fp32 pairs (sm_100: FFMA2 / FMUL2 process two lanes per issue slot)
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float f2_hsum(uint64_t v) {
  float lo, hi;
  asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// One sweep over half H of the Y accumulator of this thread's row: sum Y^2, sum g^2 Y^2, sum u Y, two features
// per instruction.  H is a template parameter so that the gamma^2 constants are constant-bank operands; two
// accumulators per sum halve the FFMA2 dependency chains.  (The sweep is bound by the FP32 pipe - FFMA2 occupies it for
// two cycles - not by TMEM latency: keeping the next chunk's loads in flight measured no gain.)
template <int H>
__device__ __forceinline__ void y_sweep(const XpoolConsts& c_xp, uint32_t lane_addr, float& s2, float& sg2, float& su) {
  uint64_t a_s2 = 0ull, a_sg2 = 0ull, a_su = 0ull;     // (+0.0f, +0.0f)
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t y[32], uh[16];
    tmem_ld_x32(lane_addr + kColY + H * 128 + c * 32, y);
    tmem_ld_x16(lane_addr + kColV + H * 64 + c * 16, uh);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 uu = __half22float2(*reinterpret_cast<const __half2*>(&uh[i]));
      const uint64_t yy = f2_pack(__uint_as_float(y[2 * i]), __uint_as_float(y[2 * i + 1]));
      const uint64_t g2 = f2_pack(c_xp.gamma2[H * 128 + c * 32 + 2 * i], c_xp.gamma2[H * 128 + c * 32 + 2 * i + 1]);
      a_s2 = f2_fma(yy, yy, a_s2);
      a_sg2 = f2_fma(f2_mul(yy, yy), g2, a_sg2);
      a_su = f2_fma(f2_pack(uu.x, uu.y), yy, a_su);
    }
  }
  s2 = f2_hsum(a_s2);
  sg2 = f2_hsum(a_sg2);
  su = f2_hsum(a_su);
}

// Blocks of 16 segments a track needs: up to its last valid segment (mask bits 0..95); at least one.
__device__ __forceinline__ int xpool_nblk(const uint4& mb) {
  const int hi = mb.z ? 95 - __clz(mb.z) : (mb.y ? 63 - __clz(mb.y) : (mb.x ? 31 - __clz(mb.x) : 0));
  return (hi >> 4) + 1;
}

// trace slot: [track u][event e] of CTA 0, 24 events per track, 64 tracks.  Compiled in only with -DMADE_XPOOL_TRACE:
// the single-lane stamps of the epilogue diverge a warp right before its warp-synchronous tcgen05.ld / bar.sync, which
// is tolerable for a timing diagnostic and not for the product kernel.
#ifdef MADE_XPOOL_DIAG
#define XP_DBG(bit) ((p.debug & (bit)) != 0)      // MADE_XPOOL_DEBUG ablations (scripts/diag_xpool.py)
#else
#define XP_DBG(bit) false
#endif
#ifdef MADE_XPOOL_TRACE
#define XP_TRACE(e) do { if (p.trace && blockIdx.x == 0 && u < 64) p.trace[u * 24 + (e)] = clock64(); __syncwarp(); } while (0)
#else
#define XP_TRACE(e) do { } while (0)
#endif

__global__ void __launch_bounds__(kXThreads, 1)
xpool_score_kernel(const __grid_constant__ XpoolMaps tm, const __grid_constant__ XpoolConsts c_xp, const XpoolParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;
  uint8_t* sZ = sK + kKBytes;
  uint8_t* sG = sZ + kZBytes;
  uint8_t* sP = sG + kGBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* zg_full = bars + 3;
  uint64_t* zg_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* y_full = bars + 7;
  uint64_t* t_free = bars + 8;
  uint64_t* y_free = bars + 9;
  uint64_t* t_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* xmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][128]
  float* xpart = xmax + 2 * kXQ;                                                     // [2 parities][128][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int slice = blockIdx.x / p.q_tiles;
  const int64_t q0 = static_cast<int64_t>(qt) * kXQ;

  if (warp == 0 && lane < 19) tma_prefetch_desc(&tm.q + lane);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(zg_full, 1);
    mbar_init(zg_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(y_full, 1);
    mbar_init(t_full, 1);
    mbar_init(t_free, 256);
    mbar_init(y_free, 256);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      for (int j = 0; j < 4; ++j) tma_load_2d(sQ + j * kSlabQ, &tm.q, q_full, j * 64, static_cast<int32_t>(q0));
      uint32_t u = 0;
      uint4 mb_next = __ldg(reinterpret_cast<const uint4*>(p.maskbits + static_cast<int64_t>(slice) * 4));
      for (int64_t m = slice; m < p.n_tracks; m += p.slices, ++u) {
        const int nb = XP_DBG(64) ? 5 : xpool_nblk(mb_next) - 1;        // rows [0, 16 (nb + 1)) of the track are fetched
        if (m + p.slices < p.n_tracks) mb_next = __ldg(reinterpret_cast<const uint4*>(p.maskbits + (m + p.slices) * 4));
        const uint32_t blk_bytes = static_cast<uint32_t>(nb + 1) * 16 * 128;      // one slab's share
        const int32_t row = static_cast<int32_t>(m * kXL);
        mbar_wait(k_empty, (u & 1) ^ 1);
        XP_TRACE(0);
        mbar_arrive_expect_tx(k_full, 4 * blk_bytes);
        for (int j = 0; j < 4; ++j) tma_load_2d(sK + j * kSlabT, &tm.k[nb], k_full, j * 64, row);
        mbar_wait(zg_empty, (u & 1) ^ 1);
        XP_TRACE(1);
        mbar_arrive_expect_tx(zg_full, 6 * blk_bytes);
        for (int j = 0; j < 2; ++j) tma_load_2d(sG + j * kSlabT, &tm.g[nb], zg_full, j * 64, row);   // cols >= 112: zero fill
        for (int j = 0; j < 4; ++j) tma_load_2d(sZ + j * kSlabT, &tm.z[nb], zg_full, j * 64, row);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc_y = umma_idesc_f16(128, 256, 0, 1);    // Y = P Z''        (B MN-major)
      constexpr uint32_t idesc_t = umma_idesc_f16(128, kXG, 0, 1);    // [T|L] = P [G|W5] (B MN-major)
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aZ = smem_u32(sZ), aG = smem_u32(sG),
                     aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      uint32_t u = 0;
      uint4 mb_next = __ldg(reinterpret_cast<const uint4*>(p.maskbits + static_cast<int64_t>(slice) * 4));
      for (int64_t m = slice; m < p.n_tracks; m += p.slices, ++u) {
        const int nblk = XP_DBG(128) ? 6 : xpool_nblk(mb_next);          // 16-segment blocks up to the last valid segment
        if (m + p.slices < p.n_tracks) mb_next = __ldg(reinterpret_cast<const uint4*>(p.maskbits + (m + p.slices) * 4));
        const uint32_t idesc_s = umma_idesc_f16(128, static_cast<uint32_t>(nblk) * 16, 0, 0);   // S = Q K^T (B K-major)
        mbar_wait(k_full, u & 1);
        mbar_wait(t_free, (u & 1) ^ 1);
        tc_fence_after_sync();
        if (!XP_DBG(16)) {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const uint32_t off = (ks >> 2) * kSlabQ + (ks & 3) * 32;
            const uint32_t offk = (ks >> 2) * kSlabT + (ks & 3) * 32;
            umma_ss(tmem_base + kColT, umma_smem_desc(aQ + off, 0, 1024), umma_smem_desc(aK + offk, 0, 1024),
                    idesc_s, ks != 0);
          }
        }
        tc_commit(k_empty);
        tc_commit(s_full);

        mbar_wait(p_full, u & 1);
        mbar_wait(zg_full, u & 1);
        tc_fence_after_sync();
        for (int ks = 0; ks < (XP_DBG(32) ? 0 : nblk); ++ks) {
          const uint32_t offp = (ks >> 2) * kSlabQ + (ks & 3) * 32;      // P: K-major A, 16 k = 32 B
          const uint32_t offb = ks * 16 * 128;                            // 16 t-rows of 128 B
          umma_ss(tmem_base + kColT, umma_smem_desc(aP + offp, 0, 1024), umma_smem_desc(aG + offb, kSlabT, 1024),
                  idesc_t, ks != 0);
        }
        tc_commit(t_full);                                                // the quadratic form can start
        mbar_wait(y_free, (u & 1) ^ 1);
        tc_fence_after_sync();
        for (int ks = 0; ks < (XP_DBG(8) ? 0 : nblk); ++ks) {
          const uint32_t offp = (ks >> 2) * kSlabQ + (ks & 3) * 32;
          const uint32_t offb = ks * 16 * 128;
          umma_ss(tmem_base + kColY, umma_smem_desc(aP + offp, 0, 1024), umma_smem_desc(aZ + offb, kSlabT, 1024),
                  idesc_y, ks != 0);
        }
        tc_commit(zg_empty);
        tc_commit(y_full);
      }
    }
  } else if (warp >= 4) {
    // ============================ epilogue ============================
    // thread = (query row r, column half h): h = 0 owns segments 0..47 / features 0..127,
    // h = 1 owns segments 48..95 / features 128..255.  Partner threads (same row, other half)
    // exchange their partial max / sums through shared memory with a 64-thread named barrier.
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const int64_t grow = q0 + r;
    const bool row_ok = grow < p.n_queries;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t bar_id = 1 + q;
    // ---- per-query constants and u = v_hat * gamma3 (fp16) -> TMEM, once ----
    float Cu1 = 0.f, Cu0 = 0.f, Cvb = 0.f;
    {
      // v_hat arrives in fp32 and u = v_hat * gamma3 is rounded to fp16 ONCE (a fp16 v_hat would add a second
      // rounding to the operand of the final dot product: tests/tools/precision_pipeline.py)
      const float4* src = reinterpret_cast<const float4*>(p.vhat + (row_ok ? grow : 0) * kXD);
#pragma unroll
      for (int c = 0; c < 8; ++c) {          // 32 features per step
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = row_ok ? __ldg(src + c * 8 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int f = c * 32 + i * 4 + j * 2;
            const float vx = tt[2 * j], vy = tt[2 * j + 1];
            const __half2 uh = __floats2half2_rn(vx * c_xp.gamma[f], vy * c_xp.gamma[f + 1]);
            const float2 uf = __half22float2(uh);
            Cu1 = fmaf(uf.x, c_xp.bias[f], Cu1);
            Cu1 = fmaf(uf.y, c_xp.bias[f + 1], Cu1);
            Cu0 += uf.x + uf.y;
            Cvb = fmaf(vx, c_xp.beta[f], Cvb);
            Cvb = fmaf(vy, c_xp.beta[f + 1], Cvb);
            w[i * 2 + j] = *reinterpret_cast<const uint32_t*>(&uh);
          }
        }
        if ((c >> 2) == h) tmem_st_x16(lane_addr + kColV + c * 16, w);   // own half of the columns
      }
      tmem_wait_st();
    }
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;
    float* my_max = xmax + h * kXQ + r;
    const float* peer_max = xmax + (h ^ 1) * kXQ + r;
    float* my_part = xpart + (h * kXQ + r) * 4;
    const float* peer_part = xpart + ((h ^ 1) * kXQ + r) * 4;
    uint32_t u = 0;
    uint4 mb_next = __ldg(reinterpret_cast<const uint4*>(p.maskbits + static_cast<int64_t>(slice) * 4));
    for (int64_t m = slice; m < p.n_tracks; m += p.slices, ++u) {
      // ---------------- A. softmax over this half's share of the track's valid blocks ----------------
#ifdef XP_NO_PREFETCH
      const uint4 mb = __ldg(reinterpret_cast<const uint4*>(p.maskbits + m * 4));
#else
      const uint4 mb = mb_next;                   // the mask words of the next track are fetched a track ahead
      if (m + p.slices < p.n_tracks) mb_next = __ldg(reinterpret_cast<const uint4*>(p.maskbits + (m + p.slices) * 4));
#endif
      const int nblk = XP_DBG(128) ? 6 : xpool_nblk(mb);
      const int n0 = (nblk + 1) >> 1;
      const int b0 = h ? n0 : 0;                  // first 16-segment block of this half
      const int nmine = h ? nblk - n0 : n0;       // 0..3 blocks
      mbar_wait(s_full, u & 1);
      if (threadIdx.x == 128) XP_TRACE(8);
      tc_fence_after_sync();
      uint32_t pk[24];
      {
        uint32_t sv[3][16];
        // loads are unconditional (block index clamped into the 96 columns) so that the arrays stay in registers;
        // blocks past this half's share are loaded and ignored
#pragma unroll
        for (int j = 0; j < 3; ++j) tmem_ld_x16(lane_addr + kColT + min(b0 + j, 5) * 16, sv[j]);
        tmem_wait_ld();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j < nmine) {
            const int b = b0 + j;
            const uint32_t w = b < 2 ? mb.x : (b < 4 ? mb.y : mb.z);
            const uint32_t bits = w >> ((b & 1) * 16);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = ((bits >> i) & 1u) ? __uint_as_float(sv[j][i]) : -INFINITY;
              sv[j][i] = __float_as_uint(a);
              mx = fmaxf(mx, a);
            }
          }
        }
        *my_max = mx;
        if (threadIdx.x == 128) XP_TRACE(9);
        named_bar_sync(bar_id, 64);
        if (threadIdx.x == 128) XP_TRACE(10);
        mx = fmaxf(mx, *peer_max);
        // P -> shared memory, 128B-swizzled K-major: logical 16-byte chunk c of row r sits at c ^ (r & 7)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j < nmine && !XP_DBG(2)) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __half2 e = __floats2half2_rn(fast_exp(__uint_as_float(sv[j][2 * i]) - mx),
                                                  fast_exp(__uint_as_float(sv[j][2 * i + 1]) - mx));
              pk[j * 8 + i] = *reinterpret_cast<const uint32_t*>(&e);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int c = (b0 + j) * 2 + k;
              *reinterpret_cast<uint4*>(prow + (c >> 3) * kSlabQ + (((c & 7) ^ sw) << 4)) =
                  make_uint4(pk[j * 8 + 4 * k], pk[j * 8 + 4 * k + 1], pk[j * 8 + 4 * k + 2], pk[j * 8 + 4 * k + 3]);
            }
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      if (threadIdx.x == 128) XP_TRACE(11);
      mbar_arrive(p_full);

      // ---------------- B. quadratic form e^T G e, the five linear sums and l = sum(e) ----------------
      mbar_wait(t_full, u & 1);
      if (threadIdx.x == 128) XP_TRACE(12);
      tc_fence_after_sync();
      float qf = 0.f, l;
      float lin[5];
      {
        uint32_t tv[3][16], t2[16];
#pragma unroll
        for (int j = 0; j < 3; ++j) tmem_ld_x16(lane_addr + kColT + min(b0 + j, 5) * 16, tv[j]);
        tmem_ld_x16(lane_addr + kColT + 96, t2);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j < nmine && !XP_DBG(4)) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&pk[j * 8 + i]));
              qf = fmaf(e.x, __uint_as_float(tv[j][2 * i]), qf);
              qf = fmaf(e.y, __uint_as_float(tv[j][2 * i + 1]), qf);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) lin[i] = __uint_as_float(t2[i]) + __uint_as_float(t2[5 + i]);   // hi + lo
        l = __uint_as_float(t2[10]);          // ones column: sum of the fp16-rounded weights, fp32 accumulation
      }
      tc_fence_before_sync();
      if (threadIdx.x == 128) XP_TRACE(13);
      mbar_arrive(t_free);

      // ---------------- C. one sweep over this half of Y ----------------
      mbar_wait(y_full, u & 1);
      if (threadIdx.x == 128) XP_TRACE(14);
      tc_fence_after_sync();
      float s2 = 0.f, sg2 = 0.f, su = 0.f;
      if (!XP_DBG(1)) {
        if (h == 0) y_sweep<0>(c_xp, lane_addr, s2, sg2, su);
        else y_sweep<1>(c_xp, lane_addr, s2, sg2, su);
      }
      tc_fence_before_sync();
      if (threadIdx.x == 128) XP_TRACE(15);
      mbar_arrive(y_free);

      // ---------------- D. combine the halves, closed-form LN2 / LN3 / cosine ----------------
      if (h == 1) { my_part[0] = qf; my_part[1] = s2; my_part[2] = sg2; my_part[3] = su; }
      named_bar_sync(bar_id, 64);
      if (h == 0) {
        qf += peer_part[0];
        s2 += peer_part[1];
        sg2 += peer_part[2];
        su += peer_part[3];
        const float inv_l = 1.0f / l;
        const float var2 = qf * inv_l * inv_l * (1.0f / kXD);
        const float alpha = rsqrtf(fmaxf(var2, 0.f) + p.ln2_eps) * inv_l;   // 1 / (l * sigma2)
        const float a2 = alpha * alpha;
        // lin = {sum Y, sum b'Y, sum g^2 Y, sum g^2 b' Y, sum g beta Y}
        const float So = fmaf(alpha, lin[0], c_xp.B1);
        const float So2 = fmaf(a2, s2, fmaf(2.f * alpha, lin[1], c_xp.B2));
        const float mean = So * (1.0f / kXD);
        const float var3 = fmaxf(So2 * (1.0f / kXD) - mean * mean, 0.f);
        const float rs = rsqrtf(var3 + p.ln3_eps);
        const float dot = fmaf(rs, fmaf(alpha, su, Cu1) - mean * Cu0, Cvb);
        const float Sg2o = fmaf(a2, sg2, fmaf(2.f * alpha, lin[3], c_xp.G2b2));
        const float Sg1o = fmaf(alpha, lin[2], c_xp.G2b);
        const float A2 = Sg2o - 2.f * mean * Sg1o + mean * mean * c_xp.G2;
        const float A1 = fmaf(alpha, lin[4], c_xp.Gbb) - mean * c_xp.Gb;
        const float n2 = fmaf(rs * rs, A2, fmaf(2.f * rs, A1, c_xp.Bb));
        if (row_ok) p.sim[grow * p.ld + p.col_offset + m] = dot / sqrtf(n2);
      }
      if (threadIdx.x == 128) XP_TRACE(16);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// W5[t, :] = Z''[t, :] . {1, b', g^2, g^2 b', g beta}, stored as an fp16 (hi, lo) pair per sum so that
// the tensor core reproduces the five linear sums to fp32 accuracy: columns 96..100 = hi,
// 101..105 = lo of the [G | W5 | 1] operand, 106 = 1.0 (the softmax denominator rides along), 107..111 zero.  Warp per row; the weight vectors live
// in global memory (lane-indexed reads of __constant__ memory would serialise).
__global__ void __launch_bounds__(256)
xpool_w5_kernel(const op_t* __restrict__ z, int64_t ldz, int64_t rows, const float* __restrict__ c5,
                op_t* __restrict__ gw) {
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  // this lane's 8-column slice of the five weight vectors stays in registers for all its rows
  float c[5][8];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(c5 + k * kXD + lane * 8));
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(c5 + k * kXD + lane * 8 + 4));
    c[k][0] = c0.x; c[k][1] = c0.y; c[k][2] = c0.z; c[k][3] = c0.w;
    c[k][4] = c1.x; c[k][5] = c1.y; c[k][6] = c1.z; c[k][7] = c1.w;
  }
  // rows are strided over the warps; the next two rows are in flight while this one is reduced
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  auto load = [&](int64_t r) { return r < rows ? __ldg(reinterpret_cast<const uint4*>(z + r * ldz + lane * 8)) : zero4; };
  uint4 r0 = load(row), r1 = load(row + n_warps);
  for (; row < rows; row += n_warps) {
    const uint4 raw = r0;
    r0 = r1;
    r1 = load(row + 2 * n_warps);
    const op2_t* hh = reinterpret_cast<const op2_t*>(&raw);
    float zz[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = op2_to_f2(hh[j]);
      zz[2 * j] = f.x;
      zz[2 * j + 1] = f.y;
    }
    float a[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float t = zz[0] * c[k][0];
#pragma unroll
      for (int j = 1; j < 8; ++j) t = fmaf(zz[j], c[k][j], t);
      a[k] = warp_sum(t);
    }
    if (lane == 0) {
      float lo[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) lo[k] = a[k] - op2f(f2op(a[k]));
      uint4* o = reinterpret_cast<uint4*>(gw + row * kXG + kXL);
      o[0] = make_uint4(pack_op2(a[0], a[1]), pack_op2(a[2], a[3]), pack_op2(a[4], lo[0]), pack_op2(lo[1], lo[2]));
      o[1] = make_uint4(pack_op2(lo[3], lo[4]), pack_op2(1.0f, 0.f), 0u, 0u);     // column 106 = 1: l = sum(e) from the MMA
    }
  }
}

// Folded constants of one checkpoint (api.cu load_xpool): `h` is passed by value to every xpool_score launch,
// `c5` ([5][256] fp32) is uploaded into the context's own device buffer for xpool_w5.
void xpool_fill_constants(const float* bias_prime, const float* gamma3, const float* beta3, XpoolConsts* hp, float* c5) {
  XpoolConsts& h = *hp;
  double B1 = 0, B2 = 0, G2 = 0, G2b2 = 0, G2b = 0, Gbb = 0, Gb = 0, Bb = 0;
  for (int i = 0; i < kXD; ++i) {
    const double b = bias_prime[i], g = gamma3[i], be = beta3[i];
    h.bias[i] = bias_prime[i];
    h.gamma[i] = gamma3[i];
    h.gamma2[i] = static_cast<float>(g * g);
    h.beta[i] = beta3[i];
    B1 += b; B2 += b * b; G2 += g * g; G2b2 += g * g * b * b; G2b += g * g * b;
    Gbb += g * be * b; Gb += g * be; Bb += be * be;
    c5[0 * kXD + i] = 1.0f;
    c5[1 * kXD + i] = bias_prime[i];
    c5[2 * kXD + i] = static_cast<float>(g * g);
    c5[3 * kXD + i] = static_cast<float>(g * g * b);
    c5[4 * kXD + i] = static_cast<float>(g * be);
  }
  h.B1 = static_cast<float>(B1); h.B2 = static_cast<float>(B2); h.G2 = static_cast<float>(G2);
  h.G2b2 = static_cast<float>(G2b2); h.G2b = static_cast<float>(G2b); h.Gbb = static_cast<float>(Gbb);
  h.Gb = static_cast<float>(Gb); h.Bb = static_cast<float>(Bb);
}

int xpool_w5(const op_t* z, int64_t ldz, int64_t rows, const float* c5_dev, op_t* gw, cudaStream_t st) {
  if (rows == 0) return MADE_OK;
  const int64_t want = ceil_div64(rows, 8), cap = static_cast<int64_t>(sm_count()) * 3;   // persistent: 3 resident CTAs per SM (78 registers)
  xpool_w5_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, st>>>(z, ldz, rows, c5_dev, gw);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

static long long* g_xpool_trace = nullptr;     // diagnostics only (made_debug_xpool_trace)

// q [n_queries,256] fp16 (pre-scaled by 1/16), vhat fp16, kz [n_tracks*96, ldkz] fp16 with the K block
// at column 0 and the Z'' block at column z_col, gw [n_tracks*96, 112] fp16 = [G | W5 | 0].
int xpool_score(const XpoolConsts& consts, const op_t* q, const float* vhat, int64_t n_queries, const op_t* kz,
                int64_t ldkz, int z_col, const op_t* gw, const uint32_t* maskbits,
                int64_t n_tracks, float* sim, int64_t ld, int64_t col_offset, cudaStream_t st) {
  if (n_queries == 0 || n_tracks == 0) return MADE_OK;
  MADE_REQUIRE(q && vhat && kz && gw && maskbits && sim, "xpool_score: null pointer");
  MADE_REQUIRE(n_tracks * kXL < (1LL << 31), "xpool_score: too many tracks for one launch");
  MADE_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&xpool_score_kernel), static_cast<int>(kXSmem)));
  XpoolMaps tm;
  const uint64_t T = static_cast<uint64_t>(n_tracks) * kXL;
  MADE_TRY(encode_tmap_2d_16b(&tm.q, q, kXD, static_cast<uint64_t>(n_queries), kXD * 2, 64, kXQ));
  for (int nb = 0; nb < 6; ++nb) {
    const uint32_t rows = 16u * (nb + 1);
    MADE_TRY(encode_tmap_2d_16b(&tm.k[nb], kz, kXD, T, static_cast<uint64_t>(ldkz) * 2, 64, rows));
    MADE_TRY(encode_tmap_2d_16b(&tm.z[nb], kz + z_col, kXD, T, static_cast<uint64_t>(ldkz) * 2, 64, rows));
    MADE_TRY(encode_tmap_2d_16b(&tm.g[nb], gw, kXG, T, kXG * 2, 64, rows));
  }
  XpoolParams p;
  p.n_queries = n_queries;
  p.n_tracks = n_tracks;
  p.q_tiles = static_cast<int>((n_queries + kXQ - 1) / kXQ);
  int sms = sm_count();
  int slices = sms / p.q_tiles;
  if (slices < 1) slices = 1;
  if (slices > n_tracks) slices = static_cast<int>(n_tracks);
  p.slices = slices;
  p.vhat = vhat;
  p.maskbits = maskbits;
  p.sim = sim;
  p.ld = ld;
  p.col_offset = col_offset;
  p.ln2_eps = 1e-5f;
  p.ln3_eps = 1e-5f;
  {
    const char* dbg = getenv("MADE_XPOOL_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
    p.trace = g_xpool_trace;
  }
  ProfScope prof_scope(kProfXpool, st);
  xpool_score_kernel<<<p.q_tiles * slices, kXThreads, kXSmem, st>>>(tm, consts, p);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

}  // namespace made

// Diagnostics (not part of include/made_b200.h): device buffer of 64 x 24 int64 clock stamps written by CTA 0 of the
// following xpool_score launches; null turns tracing off.
extern "C" int made_debug_xpool_trace(void* dev_buf) {
  made::g_xpool_trace = static_cast<long long*>(dev_buf);
  return 0;
}
