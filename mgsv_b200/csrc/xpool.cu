// Fused X-Pool scoring: for every (query video v, gallery track m) pair compute
//   sim_single[v,m] = cos( v_hat,  LN3( a + linear_proj(a) ) ),  a = LN2( out_proj( softmax_t(q.K_t/16) V ) )
// i.e. Transformer_XA.forward (modules/transformer.py:156-180, 87-123) followed by
// sim_matrix_music_pooling (modules/metrics.py:10-24), WITHOUT materialising the reference's
// [N_m, N_v, 256] pooled tensor (8.2 GB fp32 at 2k x 4k).
//
// Algebra (all exact in real arithmetic; DESIGN.md "X-Pool folding"):
//   o - mean(o)          = sum_t a_t V''_t            V'' = centred (Wo Wv) LN1(x) + centred biases
//   |o - mean(o)|^2      = a^T G a                    G   = V'' V''^T   (per track, 96x96)
//   LN2 -> (I+Wl) + bl   = (1/sigma) sum_t a_t Z''_t + b'     Z'' = V'' W'^T,  W' = (I+Wl) diag(g2)
// so per pair only three products remain:  S = q K^T (96),  T = e G (96),  Y = e Z'' (256), with
// e = exp(S - max) kept unnormalised in fp16 and divided by the sum of the rounded weights.
//
// One CTA owns a 128-query tile (Q resident in shared memory, v_hat resident in TMEM as fp16) and
// streams tracks: TMA brings K/Z''/G of a track into 128B-swizzled shared memory, one thread issues
// tcgen05.mma (S: K-major B; T,Y: MN-major B straight from the row-major [token, feature]
// matrices), and 4 epilogue warps (thread == query row) do softmax -> P (swizzled smem A operand)
// and the LN2/LN3/cosine sweep out of TMEM.  GEMM1 of track i+1 overlaps the Y sweep of track i.
#include "common.cuh"

namespace made {

constexpr int kXQ = 128;     // queries per tile
constexpr int kXL = 96;      // segments per track
constexpr int kXD = 256;
constexpr int kXThreads = 256;
constexpr uint32_t kQBytes = kXQ * kXD * 2;          // 64 KB: 4 k-slabs of [128 x 128B]
constexpr uint32_t kKBytes = kXL * kXD * 2;          // 48 KB: 4 k-slabs of [96 x 128B]
constexpr uint32_t kZBytes = kXL * kXD * 2;          // 48 KB: 4 n-slabs of [96 x 128B]
constexpr uint32_t kGBytes = 2 * kXL * 128;          // 24 KB: 2 n-slabs of [96 x 128B]
constexpr uint32_t kPBytes = 2 * kXQ * 128;          // 32 KB: 2 k-slabs of [128 x 128B]
constexpr uint32_t kSlabQ = kXQ * 128;               // 16 KB
constexpr uint32_t kSlabT = kXL * 128;               // 12 KB
constexpr uint32_t kXSmem = kQBytes + kKBytes + kZBytes + kGBytes + kPBytes + 1024 + 256;

// TMEM columns
constexpr uint32_t kColY = 0;      // 256 fp32 columns
constexpr uint32_t kColT = 256;    // 96 fp32 columns; also holds S before the softmax
constexpr uint32_t kColV = 384;    // 128 columns: v_hat as packed fp16 pairs

__constant__ float c_xp_bias[kXD];    // b' = (I + Wl) beta2 + bl
__constant__ float c_xp_gamma3[kXD];
__constant__ float c_xp_beta3[kXD];

struct XpoolParams {
  int64_t n_queries, n_tracks;
  int q_tiles, slices;
  const __half* vhat;          // [n_queries, 256] fp16
  const uint32_t* maskbits;    // [n_tracks, 4]
  float* sim;                  // [n_queries, ld]
  int64_t ld;
  int64_t col_offset;
  float ln2_eps, ln3_eps;
};

__global__ void __launch_bounds__(kXThreads, 1)
xpool_score_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_g,
                   const XpoolParams p) {
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
    mbar_init(p_full, 128);
    mbar_init(y_full, 1);
    mbar_init(t_free, 128);
    mbar_init(y_free, 128);
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
        for (int j = 0; j < 2; ++j) tma_load_2d(sG + j * kSlabT, &tm_g, zg_full, j * 64, row);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 96, 0, 0);    // S = Q K^T   (B K-major)
      constexpr uint32_t idesc_y = umma_idesc_f16(128, 256, 0, 1);   // Y = P Z''   (B MN-major)
      constexpr uint32_t idesc_t = umma_idesc_f16(128, 96, 0, 1);    // T = P G     (B MN-major)
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
          umma_ss(tmem_base + kColY, adesc, umma_smem_desc(aZ + offb, kSlabT, 1024), idesc_y, ks != 0);
          umma_ss(tmem_base + kColT, adesc, umma_smem_desc(aG + offb, kSlabT, 1024), idesc_t, ks != 0);
        }
        tc_commit(zg_empty);
        tc_commit(y_full);
      }
    }
  } else if (warp >= 4) {
    // ============================ epilogue: thread == query row ============================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int64_t grow = q0 + r;
    const bool row_ok = grow < p.n_queries;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // v_hat row -> TMEM (fp16 pairs), once
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.vhat + (row_ok ? grow : 0) * kXD);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 t = row_ok ? __ldg(src + c * 8 + i) : make_uint4(0, 0, 0, 0);
          w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
        }
        tmem_st_x32(lane_addr + kColV + c * 32, w);
      }
      tmem_wait_st();
    }
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;
    uint32_t u = 0;
    for (int64_t m = slice; m < p.n_tracks; m += p.slices, ++u) {
      // ---------------- softmax over the 96 segments ----------------
      const uint4 mb = __ldg(reinterpret_cast<const uint4*>(p.maskbits + m * 4));
      const uint32_t mw[3] = {mb.x, mb.y, mb.z};
      mbar_wait(s_full, u & 1);
      tc_fence_after_sync();
      uint32_t pk[48];
      float lsum = 0.f;
      {
        uint32_t s0[32], s1[32], s2[32];
        tmem_ld_x32(lane_addr + kColT + 0, s0);
        tmem_ld_x32(lane_addr + kColT + 32, s1);
        tmem_ld_x32(lane_addr + kColT + 64, s2);
        tmem_wait_ld();
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float a = ((mw[0] >> i) & 1u) ? __uint_as_float(s0[i]) : -INFINITY;
          float b = ((mw[1] >> i) & 1u) ? __uint_as_float(s1[i]) : -INFINITY;
          float c = ((mw[2] >> i) & 1u) ? __uint_as_float(s2[i]) : -INFINITY;
          s0[i] = __float_as_uint(a); s1[i] = __float_as_uint(b); s2[i] = __float_as_uint(c);
          mx = fmaxf(mx, fmaxf(a, fmaxf(b, c)));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          op2_t e0 = floats2op2(__expf(__uint_as_float(s0[2 * i]) - mx),
                                                    __expf(__uint_as_float(s0[2 * i + 1]) - mx));
          op2_t e1 = floats2op2(__expf(__uint_as_float(s1[2 * i]) - mx),
                                                    __expf(__uint_as_float(s1[2 * i + 1]) - mx));
          op2_t e2 = floats2op2(__expf(__uint_as_float(s2[2 * i]) - mx),
                                                    __expf(__uint_as_float(s2[2 * i + 1]) - mx));
          float2 f0 = op2_to_f2(e0), f1 = op2_to_f2(e1), f2 = op2_to_f2(e2);
          lsum += (f0.x + f0.y) + (f1.x + f1.y) + (f2.x + f2.y);
          pk[i] = *reinterpret_cast<uint32_t*>(&e0);
          pk[16 + i] = *reinterpret_cast<uint32_t*>(&e1);
          pk[32 + i] = *reinterpret_cast<uint32_t*>(&e2);
        }
      }
      // P -> shared memory, 128B-swizzled K-major: logical 16-byte chunk c of row r sits at c ^ (r & 7)
#pragma unroll
      for (int c = 0; c < 12; ++c) {
        const int slab = c >> 3, cc = c & 7;
        *reinterpret_cast<uint4*>(prow + slab * kSlabQ + ((cc ^ sw) << 4)) =
            make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(p_full);

      // ---------------- sigma of LN2 from the quadratic form e^T G e ----------------
      mbar_wait(y_full, u & 1);
      tc_fence_after_sync();
      float qf = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t tt[32];
        tmem_ld_x32(lane_addr + kColT + c * 32, tt);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 e = op2_to_f2(*reinterpret_cast<const op2_t*>(&pk[c * 16 + i]));
          qf = fmaf(e.x, __uint_as_float(tt[2 * i]), qf);
          qf = fmaf(e.y, __uint_as_float(tt[2 * i + 1]), qf);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(t_free);
      const float inv_l = 1.0f / lsum;
      const float var2 = qf * inv_l * inv_l * (1.0f / kXD);
      const float alpha = rsqrtf(fmaxf(var2, 0.f) + p.ln2_eps) * inv_l;   // 1 / (l * sigma)

      // ---------------- LN3 statistics (pass 1) ----------------
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t y[32];
        tmem_ld_x32(lane_addr + kColY + c * 32, y);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float o = fmaf(alpha, __uint_as_float(y[i]), c_xp_bias[c * 32 + i]);
          sum += o;
          sq = fmaf(o, o, sq);
        }
      }
      const float mean = sum * (1.0f / kXD);
      const float var3 = fmaxf(sq * (1.0f / kXD) - mean * mean, 0.f);
      const float rs3 = rsqrtf(var3 + p.ln3_eps);
      // ---------------- LN3 output norm and cosine with v_hat (pass 2) ----------------
      float n2 = 0.f, dot = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t y[32], vh[16];
        tmem_ld_x32(lane_addr + kColY + c * 32, y);
        tmem_ld_x16(lane_addr + kColV + c * 16, vh);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 v = __half22float2(*reinterpret_cast<const __half2*>(&vh[i]));
          float o0 = fmaf(alpha, __uint_as_float(y[2 * i]), c_xp_bias[c * 32 + 2 * i]);
          float o1 = fmaf(alpha, __uint_as_float(y[2 * i + 1]), c_xp_bias[c * 32 + 2 * i + 1]);
          float t0 = fmaf((o0 - mean) * rs3, c_xp_gamma3[c * 32 + 2 * i], c_xp_beta3[c * 32 + 2 * i]);
          float t1 = fmaf((o1 - mean) * rs3, c_xp_gamma3[c * 32 + 2 * i + 1], c_xp_beta3[c * 32 + 2 * i + 1]);
          n2 = fmaf(t0, t0, n2);
          n2 = fmaf(t1, t1, n2);
          dot = fmaf(v.x, t0, dot);
          dot = fmaf(v.y, t1, dot);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(y_free);
      if (row_ok) p.sim[grow * p.ld + p.col_offset + m] = dot / sqrtf(n2);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

int xpool_set_constants(const float* bias_prime, const float* gamma3, const float* beta3, cudaStream_t st) {
  MADE_CUDA(cudaMemcpyToSymbolAsync(c_xp_bias, bias_prime, kXD * 4, 0, cudaMemcpyHostToDevice, st));
  MADE_CUDA(cudaMemcpyToSymbolAsync(c_xp_gamma3, gamma3, kXD * 4, 0, cudaMemcpyHostToDevice, st));
  MADE_CUDA(cudaMemcpyToSymbolAsync(c_xp_beta3, beta3, kXD * 4, 0, cudaMemcpyHostToDevice, st));
  return MADE_OK;
}

// q [n_queries,256] fp16 (pre-scaled by 1/16), vhat fp16, kz [n_tracks*96, ldkz] fp16 with the K block
// at column 0 and the Z'' block at column z_col, gram [n_tracks*96, 96] fp16.
int xpool_score(const op_t* q, const __half* vhat, int64_t n_queries, const op_t* kz,
                int64_t ldkz, int z_col, const op_t* gram, const uint32_t* maskbits,
                int64_t n_tracks, float* sim, int64_t ld, int64_t col_offset, cudaStream_t st) {
  if (n_queries == 0 || n_tracks == 0) return MADE_OK;
  MADE_REQUIRE(q && vhat && kz && gram && maskbits && sim, "xpool_score: null pointer");
  MADE_REQUIRE(n_tracks * kXL < (1LL << 31), "xpool_score: too many tracks for one launch");
  static bool attr_set = false;
  if (!attr_set) {
    MADE_CUDA(cudaFuncSetAttribute(xpool_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kXSmem));
    attr_set = true;
  }
  CUtensorMap tq, tk, tz, tg;
  const uint64_t T = static_cast<uint64_t>(n_tracks) * kXL;
  MADE_TRY(encode_tmap_2d_16b(&tq, q, kXD, static_cast<uint64_t>(n_queries), kXD * 2, 64, kXQ));
  MADE_TRY(encode_tmap_2d_16b(&tk, kz, kXD, T, static_cast<uint64_t>(ldkz) * 2, 64, kXL));
  MADE_TRY(encode_tmap_2d_16b(&tz, kz + z_col, kXD, T, static_cast<uint64_t>(ldkz) * 2, 64, kXL));
  MADE_TRY(encode_tmap_2d_16b(&tg, gram, kXL, T, kXL * 2, 64, kXL));
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
  xpool_score_kernel<<<p.q_tiles * slices, kXThreads, kXSmem, st>>>(tq, tk, tz, tg, p);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

}  // namespace made
