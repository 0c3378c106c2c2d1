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
constexpr uint32_t kXchgBytes = 2 * kXQ * 4 + 2 * kXQ * 5 * 4;   // max + 5 partial sums per (half, row)
constexpr uint32_t kXSmem = kQBytes + kKBytes + kZBytes + kGBytes + kPBytes + 1024 + 256 + kXchgBytes;
static_assert(kXSmem <= 232448, "shared memory budget of an sm_100 CTA");

// TMEM columns
constexpr uint32_t kColY = 0;      // 256 fp32 columns
constexpr uint32_t kColT = 256;    // 112 fp32 columns [T | L]; also holds S before the softmax
constexpr uint32_t kColV = 384;    // 128 columns: u = v_hat * gamma3 as packed fp16 pairs

// XpoolConsts (prep.cuh) travels as a __grid_constant__ kernel parameter: parameters live in constant bank 0,
// so gamma3^2 stays an immediate constant-bank FFMA operand in the Y sweep, and every made_ctx has its own copy.

struct XpoolParams {
  int64_t n_queries, n_tracks;
  int q_tiles, slices;
  const float* vhat;           // [n_queries, 256] fp32: v / |v| (modules/metrics.py:19)
  const uint32_t* maskbits;    // [n_tracks, 4]
  float* sim;                  // [n_queries, ld]
  int64_t ld;
  int64_t col_offset;
  float ln2_eps, ln3_eps;
};

// One sweep over half H of the Y accumulator of this thread's row: sum Y^2, sum g^2 Y^2, sum u Y.
// H is a template parameter so that the gamma^2 constants are immediate constant-bank operands.
template <int H>
__device__ __forceinline__ void y_sweep(const XpoolConsts& c_xp, uint32_t lane_addr, float& s2, float& sg2, float& su) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t y[32], uh[16];
    tmem_ld_x32(lane_addr + kColY + H * 128 + c * 32, y);
    tmem_ld_x16(lane_addr + kColV + H * 64 + c * 16, uh);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 uu = __half22float2(*reinterpret_cast<const __half2*>(&uh[i]));
      const float y0 = __uint_as_float(y[2 * i]), y1 = __uint_as_float(y[2 * i + 1]);
      const float a0 = y0 * y0, a1 = y1 * y1;
      s2 += a0 + a1;
      sg2 = fmaf(a0, c_xp.gamma2[H * 128 + c * 32 + 2 * i], sg2);
      sg2 = fmaf(a1, c_xp.gamma2[H * 128 + c * 32 + 2 * i + 1], sg2);
      su = fmaf(uu.x, y0, su);
      su = fmaf(uu.y, y1, su);
    }
  }
}

__global__ void __launch_bounds__(kXThreads, 1)
xpool_score_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_g,
                   const __grid_constant__ XpoolConsts c_xp, const XpoolParams p) {
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* xmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][128]
  float* xpart = xmax + 2 * kXQ;                                                     // [2][128][5]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int slice = blockIdx.x / p.q_tiles;
  const int64_t q0 = static_cast<int64_t>(qt) * kXQ;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_z);
    tma_prefetch_desc(&tm_g);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(zg_full, 1);
    mbar_init(zg_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(y_full, 1);
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
      for (int j = 0; j < 4; ++j) tma_load_2d(sQ + j * kSlabQ, &tm_q, q_full, j * 64, static_cast<int32_t>(q0));
      uint32_t u = 0;
      for (int64_t m = slice; m < p.n_tracks; m += p.slices, ++u) {
        const int32_t row = static_cast<int32_t>(m * kXL);
        mbar_wait(k_empty, (u & 1) ^ 1);
        mbar_arrive_expect_tx(k_full, kKBytes);
        for (int j = 0; j < 4; ++j) tma_load_2d(sK + j * kSlabT, &tm_k, k_full, j * 64, row);
        mbar_wait(zg_empty, (u & 1) ^ 1);
        mbar_arrive_expect_tx(zg_full, kZBytes + kGBytes);
        for (int j = 0; j < 4; ++j) tma_load_2d(sZ + j * kSlabT, &tm_z, zg_full, j * 64, row);
        for (int j = 0; j < 2; ++j) tma_load_2d(sG + j * kSlabT, &tm_g, zg_full, j * 64, row);   // cols >= 112: zero fill
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 96, 0, 0);     // S = Q K^T        (B K-major)
      constexpr uint32_t idesc_y = umma_idesc_f16(128, 256, 0, 1);    // Y = P Z''        (B MN-major)
      constexpr uint32_t idesc_t = umma_idesc_f16(128, kXG, 0, 1);    // [T|L] = P [G|W5] (B MN-major)
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aZ = smem_u32(sZ), aG = smem_u32(sG),
                     aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      uint32_t u = 0;
      for (int64_t m = slice; m < p.n_tracks; m += p.slices, ++u) {
        mbar_wait(k_full, u & 1);
        mbar_wait(t_free, (u & 1) ^ 1);
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t off = (ks >> 2) * kSlabQ + (ks & 3) * 32;
          const uint32_t offk = (ks >> 2) * kSlabT + (ks & 3) * 32;
          umma_ss(tmem_base + kColT, umma_smem_desc(aQ + off, 0, 1024), umma_smem_desc(aK + offk, 0, 1024),
                  idesc_s, ks != 0);
        }
        tc_commit(k_empty);
        tc_commit(s_full);

        mbar_wait(p_full, u & 1);
        mbar_wait(zg_full, u & 1);
        mbar_wait(y_free, (u & 1) ^ 1);
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) {
          const uint32_t offp = (ks >> 2) * kSlabQ + (ks & 3) * 32;      // P: K-major A, 16 k = 32 B
          const uint64_t adesc = umma_smem_desc(aP + offp, 0, 1024);
          const uint32_t offb = ks * 16 * 128;                            // 16 t-rows of 128 B
          umma_ss(tmem_base + kColT, adesc, umma_smem_desc(aG + offb, kSlabT, 1024), idesc_t, ks != 0);
          umma_ss(tmem_base + kColY, adesc, umma_smem_desc(aZ + offb, kSlabT, 1024), idesc_y, ks != 0);
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
    float* my_part = xpart + (h * kXQ + r) * 5;
    const float* peer_part = xpart + ((h ^ 1) * kXQ + r) * 5;
    uint32_t u = 0;
    for (int64_t m = slice; m < p.n_tracks; m += p.slices, ++u) {
      // ---------------- A. softmax over this half's 48 segments ----------------
      const uint4 mb = __ldg(reinterpret_cast<const uint4*>(p.maskbits + m * 4));
      const uint64_t lo64 = (static_cast<uint64_t>(mb.y) << 32) | mb.x;
      const uint64_t bits = h == 0 ? lo64 : ((lo64 >> 48) | (static_cast<uint64_t>(mb.z) << 16));   // 48 valid bits
      mbar_wait(s_full, u & 1);
      tc_fence_after_sync();
      uint32_t pk[24];
      float lsum = 0.f;
      {
        uint32_t s0[32], s1[16];
        tmem_ld_x32(lane_addr + kColT + h * 48, s0);
        tmem_ld_x16(lane_addr + kColT + h * 48 + 32, s1);
        tmem_wait_ld();
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float a = ((bits >> i) & 1ull) ? __uint_as_float(s0[i]) : -INFINITY;
          s0[i] = __float_as_uint(a);
          mx = fmaxf(mx, a);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = ((bits >> (32 + i)) & 1ull) ? __uint_as_float(s1[i]) : -INFINITY;
          s1[i] = __float_as_uint(a);
          mx = fmaxf(mx, a);
        }
        *my_max = mx;
        named_bar_sync(bar_id, 64);
        mx = fmaxf(mx, *peer_max);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const __half2 e = __floats2half2_rn(fast_exp(__uint_as_float(s0[2 * i]) - mx),
                                              fast_exp(__uint_as_float(s0[2 * i + 1]) - mx));
          const float2 f = __half22float2(e);
          lsum += f.x + f.y;
          pk[i] = *reinterpret_cast<const uint32_t*>(&e);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __half2 e = __floats2half2_rn(fast_exp(__uint_as_float(s1[2 * i]) - mx),
                                              fast_exp(__uint_as_float(s1[2 * i + 1]) - mx));
          const float2 f = __half22float2(e);
          lsum += f.x + f.y;
          pk[16 + i] = *reinterpret_cast<const uint32_t*>(&e);
        }
      }
      // P -> shared memory, 128B-swizzled K-major: logical 16-byte chunk c of row r sits at c ^ (r & 7)
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int c = h * 6 + k;
        const int slab = c >> 3, cc = c & 7;
        *reinterpret_cast<uint4*>(prow + slab * kSlabQ + ((cc ^ sw) << 4)) =
            make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(p_full);

      // ---------------- B. quadratic form e^T G e and the five linear sums ----------------
      mbar_wait(y_full, u & 1);
      tc_fence_after_sync();
      float qf = 0.f;
      float lin[5];
      {
        uint32_t t0[32], t1[16], t2[16];
        tmem_ld_x32(lane_addr + kColT + h * 48, t0);
        tmem_ld_x16(lane_addr + kColT + h * 48 + 32, t1);
        tmem_ld_x16(lane_addr + kColT + 96, t2);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&pk[i]));
          qf = fmaf(e.x, __uint_as_float(t0[2 * i]), qf);
          qf = fmaf(e.y, __uint_as_float(t0[2 * i + 1]), qf);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&pk[16 + i]));
          qf = fmaf(e.x, __uint_as_float(t1[2 * i]), qf);
          qf = fmaf(e.y, __uint_as_float(t1[2 * i + 1]), qf);
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) lin[i] = __uint_as_float(t2[i]) + __uint_as_float(t2[5 + i]);   // hi + lo
      }
      tc_fence_before_sync();
      mbar_arrive(t_free);

      // ---------------- C. one sweep over this half of Y ----------------
      float s2 = 0.f, sg2 = 0.f, su = 0.f;
      if (h == 0) y_sweep<0>(c_xp, lane_addr, s2, sg2, su);
      else y_sweep<1>(c_xp, lane_addr, s2, sg2, su);
      tc_fence_before_sync();
      mbar_arrive(y_free);

      // ---------------- D. combine the halves, closed-form LN2 / LN3 / cosine ----------------
      my_part[0] = lsum; my_part[1] = qf; my_part[2] = s2; my_part[3] = sg2; my_part[4] = su;
      named_bar_sync(bar_id, 64);
      if (h == 0) {
        const float l = lsum + peer_part[0];
        qf += peer_part[1];
        s2 += peer_part[2];
        sg2 += peer_part[3];
        su += peer_part[4];
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
// 101..105 = lo of the [G | W5 | 0] operand (106..111 zero).  Warp per row; the weight vectors live
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
      o[1] = make_uint4(pack_op2(lo[3], lo[4]), 0u, 0u, 0u);
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

// q [n_queries,256] fp16 (pre-scaled by 1/16), vhat fp16, kz [n_tracks*96, ldkz] fp16 with the K block
// at column 0 and the Z'' block at column z_col, gw [n_tracks*96, 112] fp16 = [G | W5 | 0].
int xpool_score(const XpoolConsts& consts, const op_t* q, const float* vhat, int64_t n_queries, const op_t* kz,
                int64_t ldkz, int z_col, const op_t* gw, const uint32_t* maskbits,
                int64_t n_tracks, float* sim, int64_t ld, int64_t col_offset, cudaStream_t st) {
  if (n_queries == 0 || n_tracks == 0) return MADE_OK;
  MADE_REQUIRE(q && vhat && kz && gw && maskbits && sim, "xpool_score: null pointer");
  MADE_REQUIRE(n_tracks * kXL < (1LL << 31), "xpool_score: too many tracks for one launch");
  MADE_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&xpool_score_kernel), static_cast<int>(kXSmem)));
  CUtensorMap tq, tk, tz, tg;
  const uint64_t T = static_cast<uint64_t>(n_tracks) * kXL;
  MADE_TRY(encode_tmap_2d_16b(&tq, q, kXD, static_cast<uint64_t>(n_queries), kXD * 2, 64, kXQ));
  MADE_TRY(encode_tmap_2d_16b(&tk, kz, kXD, T, static_cast<uint64_t>(ldkz) * 2, 64, kXL));
  MADE_TRY(encode_tmap_2d_16b(&tz, kz + z_col, kXD, T, static_cast<uint64_t>(ldkz) * 2, 64, kXL));
  MADE_TRY(encode_tmap_2d_16b(&tg, gw, kXG, T, kXG * 2, 64, kXL));
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
  ProfScope prof_scope(kProfXpool, st);
  xpool_score_kernel<<<p.q_tiles * slices, kXThreads, kXSmem, st>>>(tq, tk, tz, tg, consts, p);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

}  // namespace made
