// tcgen05 / TMA / TMEM GEMM with fused epilogues (see gemm_tc.cuh).
#include "gemm_tc.cuh"

#include <cstdlib>
#include <cstring>

namespace made {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                 // 64 fp16 = one 128-byte swizzle row
constexpr int kWarpBoxBytes = 32 * 128;              // one TMA store box: a warp's 32 rows x 128 bytes (swizzled)
constexpr int kAccStages = 2;
constexpr int kGemmThreads = 384;           // 4 control warps + 8 epilogue warps
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KB

// WS = weight-stationary: the whole [BN x K<=256] slice of W stays in shared memory while the CTA
// walks a contiguous range of M tiles (n-major tile order), so only A tiles stream through the ring.
constexpr int kWsKBlocks = 4;                       // K <= 256
// PAIR = CTA pair (cta_group::2): the two CTAs of a 2-CTA cluster compute ONE [256 x BN] tile with 256-row MMAs issued by
// the leader.  Each CTA stages its own 128 rows of A and only its own HALF of the W rows of a k block (32 instead of 48
// KB per k block and CTA.  A single-CTA k block moves 48 KB into shared memory and its four MMAs read 48 KB back: 96 KB
// through a 128 B / cycle shared memory = 750 cycles against 512 cycles of MMA (measured: 767); a pair moves 64 KB
// per CTA and its mainloop runs at 7.6 k instead of 9.2 k cycles per split-2 tile), holds the
// accumulator of its 128 rows in its own TMEM and runs the unchanged epilogue on them.
template <int BN, bool WS = false, bool PAIR = false>
struct GemmCfg {
  static_assert(!PAIR || (BN == 256 && !WS), "CTA pairs: 256-column streaming tiles only");
  static constexpr int kStages = PAIR ? 4 : 3;
  static constexpr int kBRows = PAIR ? BN / 2 : BN;               // W rows of a k block this CTA stages
  static constexpr int kBTileBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = WS ? kATileBytes : kATileBytes + kBTileBytes;
  // WS, BN = 256: the [256 x K] slice of a plain fp16 W (4 k-tiles).  WS, BN = 128: the [128 x K] slices of BOTH halves of
  // a (hi | lo) weight pair (8 k-tiles, 128 KB too) — the weight-stationary form of the split-precision GEMMs, whose
  // streaming form re-reads 1.1-1.7 MB of W tiles per 128 rows
  static constexpr int kResidentTiles = WS ? (BN == 128 ? 2 * kWsKBlocks : kWsKBlocks) : 0;
  static constexpr int kResidentBytes = kResidentTiles * kBTileBytes;
  static constexpr int kAccStride = BN <= 128 ? 128 : 256;      // TMEM columns per acc stage
  static constexpr uint32_t kTmemCols = kAccStride * kAccStages;  // 256 or 512
  static constexpr int kChunks = BN / 32;
  // dynamic smem: tiles + barriers + tmem slot + LN partials
  static constexpr int kVecFloats = 2048 + 2 * 256;         // bias [N <= 2048], LayerNorm gamma / beta [256] (epilogue copies)
  static constexpr int kMiscBytes = 256 + 4 * 128 * 4 + kVecFloats * 4;      // barriers + TMEM slot, LN partial sums, vectors
  static constexpr int kStageOutOffset = ((kResidentBytes + kStages * kStageBytes + kMiscBytes + 1023) / 1024) * 1024;
  // the weight-stationary variant has room for the fp16 boxes only (fp32 outputs go to the streaming variant)
  static constexpr int kStageOutBytes = (WS ? 1 : 2) * kEpiWarps * kWarpBoxBytes;
  static constexpr int kSmemBytes = kStageOutOffset + kStageOutBytes + 1024 /*align slack*/;
};

// GELU(erf) (nn.GELU default, model_Base.py:77).  erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7,
// far below the fp16 rounding of the stored activation): one MUFU.RCP, one MUFU.EX2, 8 FMA-class ops,
// no branches — about half the instructions of erff().
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);        // erf(|x| / sqrt(2))
  return 0.5f * x * (1.0f + copysignf(e, x));
}

// low halves of eight values whose fp16-rounded high halves are `hi`: fp16(v - float(hi))
__device__ __forceinline__ uint4 lo_of(const uint4& hi, const float* v) {
  const op2_t* h = reinterpret_cast<const op2_t*>(&hi);
  const float2 f0 = op2_to_f2(h[0]), f1 = op2_to_f2(h[1]), f2 = op2_to_f2(h[2]), f3 = op2_to_f2(h[3]);
  return make_uint4(pack_op2(v[0] - f0.x, v[1] - f0.y), pack_op2(v[2] - f1.x, v[3] - f1.y),
                    pack_op2(v[4] - f2.x, v[5] - f2.y), pack_op2(v[6] - f3.x, v[7] - f3.y));
}

// Diagnostics (scripts/diag_gemm_trace.py, MADE_GEMM_DEBUG ablations) are compiled in only with -DMADE_GEMM_DIAG
// (`MADE_DIAG=1 python -m mgsv_b200.build`): the product kernel carries none of their branches.
#ifdef MADE_GEMM_DIAG
#define GEMM_TRACE(slot) do { if (p.trace && blockIdx.x == 0) p.trace[slot] = clock64(); } while (0)
#define GEMM_TRACING (p.trace != nullptr)
#define GEMM_DBG(bit) ((p.debug & (bit)) != 0)
#else
#define GEMM_TRACE(slot) do { } while (0)
#define GEMM_TRACING false
#define GEMM_DBG(bit) false
#endif

template <int BN, bool WS, bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_oh, const __grid_constant__ CUtensorMap tmap_of,
               const __grid_constant__ CUtensorMap tmap_ol, const GemmParams p) {
  using Cfg = GemmCfg<BN, WS, PAIR>;
  constexpr int kStages = Cfg::kStages;
  if (threadIdx.x == 0) GEMM_TRACE(0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* resident = smem;                           // WS: W slice, kWsKBlocks tiles of [BN x 64]
  uint8_t* tiles = smem + Cfg::kResidentBytes;
  uint8_t* after = tiles + kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + kAccStages;
  uint64_t* w_full = tmem_empty + kAccStages;         // WS: resident W slice landed
  uint64_t* w_empty = w_full + 1;                     // WS: every MMA that reads the slice has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + 1);
  float* ln_part = reinterpret_cast<float*>(after + 256);   // [2 halves][2][128]
  float* vec_s = ln_part + 4 * 128;                         // bias [<= 2048] | gamma [256] | beta [256]
  uint8_t* stage_out = smem + Cfg::kStageOutOffset;         // per epilogue warp: one fp16 box, then one fp32 box

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_blocks = p.N / BN;
  const int64_t M = p.m_dev ? static_cast<int64_t>(__ldg(p.m_dev)) : p.M;   // ragged batches: rows actually present
  const int64_t m_tiles = (M + p.m_stride - 1) / p.m_stride;
  const int64_t n_tiles = m_tiles * n_blocks;
  const int kb_seg = (p.K + kBlockK - 1) / kBlockK;        // k blocks of one operand half
  // pass s of a split GEMM reads the A / W halves at these column offsets (hi halves at 0, lo halves at K).  The
  // producer and the MMA issuer are ONE thread each: their per-k-block code is a serial chain of dependent
  // instructions, so it is kept free of divisions (a profile showed the producer 70 % busy with index arithmetic and
  // the k blocks arriving ~800 cycles apart because of it)
  auto a_off = [&](int pass) { return (p.split == 2 && pass == 1) ? p.K : 0; };
  auto w_off = [&](int pass) { return (p.split != 0 && pass == p.split) ? p.K : 0; };
  // tile schedule: streaming = round-robin, m-major; WS = one contiguous n-major range per CTA
  const int64_t ws_per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t ws_t0 = static_cast<int64_t>(blockIdx.x) * ws_per;
  // PAIR: the cluster (blockIdx.x / 2) walks [256 x BN] pair tiles round-robin; rank r of the pair owns rows 128 r ..
  // of each (an odd tile count leaves the last pair's second half empty: its loads read zeros, its stores are clipped)
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int64_t n_ptiles = ((m_tiles + 1) / 2) * n_blocks;
  const int64_t n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const int64_t my_tiles = PAIR ? (n_ptiles - cluster_id + n_clusters - 1) / n_clusters
                           : WS ? (ws_t0 >= n_tiles ? 0 : (n_tiles - ws_t0 < ws_per ? n_tiles - ws_t0 : ws_per))
                                : (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  auto decode = [&](int64_t i, int64_t& m_blk, int& n_blk) {
    if constexpr (PAIR) {
      const int64_t t = cluster_id + i * n_clusters;
      m_blk = (t / n_blocks) * 2 + cta_rank;
      n_blk = static_cast<int>(t % n_blocks);
    } else if constexpr (WS) {
      const int64_t t = ws_t0 + i;
      n_blk = static_cast<int>(t / m_tiles);
      m_blk = t % m_tiles;
    } else {
      const int64_t t = blockIdx.x + i * gridDim.x;
      m_blk = t / n_blocks;
      n_blk = static_cast<int>(t % n_blocks);
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.tma_store) {
      if (p.epi.out_h) tma_prefetch_desc(&tmap_oh);
      if (p.epi.out_f32) tma_prefetch_desc(&tmap_of);
      if (p.epi.out_lo) tma_prefetch_desc(&tmap_ol);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], PAIR ? 2 * kEpiWarps : kEpiWarps);    // PAIR: the epilogue warps of both CTAs (leader's copy)
    }
    mbar_init(w_full, 1);
    mbar_init(w_empty, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();     // the peer's barriers exist before anything signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) GEMM_TRACE(1);

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int cur_n = -1;
      uint32_t groups = 0;
      const uint32_t full_bar_leader = PAIR ? mapa_u32(smem_u32(full_bar), 0) : 0;
      for (int64_t i = 0; i < my_tiles; ++i) {
        int64_t m_blk;
        int n_blk;
        decode(i, m_blk, n_blk);
        const int32_t row_a = static_cast<int32_t>(m_blk * p.m_stride);
        const int32_t row_b = p.b_batched ? row_a : n_blk * BN + static_cast<int32_t>(cta_rank) * Cfg::kBRows;
        if constexpr (WS) {
          if (n_blk != cur_n) {     // (re)load the resident W slice once the MMAs of the old one are done
            if (groups > 0) mbar_wait(w_empty, (groups - 1) & 1);
            // every distinct [BN x 64] block of W once: the hi halves, then (split) the lo halves K columns to the right
            const int n_w = kb_seg * (p.split ? 2 : 1);
            mbar_arrive_expect_tx(w_full, static_cast<uint32_t>(n_w) * Cfg::kBTileBytes);
            for (int t = 0; t < n_w; ++t)
              tma_load_2d(resident + t * Cfg::kBTileBytes, &tmap_b, w_full, (t / kb_seg) * p.K + (t % kb_seg) * kBlockK, row_b);
            cur_n = n_blk;
            ++groups;
          }
        }
        // Optional (MADE_GEMM_L2_PREFETCH=1, off by default): pull the A tiles of this CTA's NEXT output tile into L2.
        // Measured: no gain on L2-resident operands (23.6 us either way on the music-chunk GEMM) and 3 - 7 % SLOWER on
        // operands streamed from HBM (303 k rows: 147 -> 157 us), +2 % on the whole job: the HBM latency of A is not
        // what bounds the mainloop.  ONE linear `cp.async.bulk.prefetch.L2` of the next tile's
        // contiguous rows instead of the tensor-map prefetches measured the same (145 -> 153 us).
        if (p.l2_prefetch && i + 1 < my_tiles) {
          int64_t m_nx;
          int n_nx;
          decode(i + 1, m_nx, n_nx);
          if (m_nx != m_blk) {
            const int32_t row_nx = static_cast<int32_t>(m_nx * p.m_stride);
            const int n_a = kb_seg * (p.split == 2 ? 2 : 1);
            for (int kb = 0; kb < n_a; ++kb)
              tma_prefetch_l2_2d(&tmap_a, (kb / kb_seg) * p.K + (kb % kb_seg) * kBlockK, row_nx);
          }
        }
        for (int pass = 0; pass <= p.split; ++pass) {
          int ca = a_off(pass), cw = w_off(pass);
          for (int ks = 0; ks < kb_seg; ++ks, ca += kBlockK, cw += kBlockK) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = tiles + stage * Cfg::kStageBytes;
            if constexpr (PAIR) {
              // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the two stages
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
              const uint32_t fb = full_bar_leader + stage * 8;
              tma_load_2d_pair(sa, &tmap_a, fb, ca, row_a);
              tma_load_2d_pair(sa + kATileBytes, &tmap_b, fb, cw, row_b);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              if (i == 0 && pass == 0 && ks == 0) GEMM_TRACE(2);
              tma_load_2d(sa, &tmap_a, &full_bar[stage], ca, row_a);
              if constexpr (!WS) tma_load_2d(sa + kATileBytes, &tmap_b, &full_bar[stage], cw, row_b);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_f16(PAIR ? 2 * kBlockM : kBlockM, BN, 0, 0);
      const uint32_t tiles_u32 = smem_u32(tiles), resident_u32 = smem_u32(resident);
      int stage = 0;
      uint32_t phase = 0;
      int cur_n = -1;
      uint32_t groups = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        int64_t m_blk;
        int n_blk, n_next = -1;
        decode(it, m_blk, n_blk);
        if (it + 1 < my_tiles) {
          int64_t m2;
          decode(it + 1, m2, n_next);
        }
        const int as = static_cast<int>(it & 1);
        const uint32_t aphase = static_cast<uint32_t>((it >> 1) & 1);
        if constexpr (PAIR) mbar_wait_cluster(&tmem_empty[as], aphase ^ 1);
        else mbar_wait(&tmem_empty[as], aphase ^ 1);
        if constexpr (WS) {
          if (n_blk != cur_n) {
            mbar_wait(w_full, groups & 1);
            cur_n = n_blk;
            ++groups;
          }
        }
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * Cfg::kAccStride;
        for (int pass = 0; pass <= p.split; ++pass) {
          // resident block of a k block: its position inside the operand half (+ kb_seg for the lo halves)
          const int wbase = (p.split != 0 && pass == p.split) ? kb_seg : 0;
          for (int ks = 0; ks < kb_seg; ++ks) {
            const bool first = (pass | ks) == 0, last = pass == p.split && ks == kb_seg - 1;
            mbar_wait(&full_bar[stage], phase);
            if (it == 0 && first) GEMM_TRACE(3);
            if (it == 0 && last) GEMM_TRACE(4);
            tc_fence_after_sync();
            const uint32_t sa = tiles_u32 + stage * Cfg::kStageBytes;
            const uint32_t sb = WS ? resident_u32 + (wbase + ks) * Cfg::kBTileBytes : sa + kATileBytes;
            const uint64_t adesc = umma_smem_desc(sa, 0, 1024);
            const uint64_t bdesc = umma_smem_desc(sb, 0, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle row: +2 in (addr>>4)
              if constexpr (PAIR) umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, !first || k != 0);
              else umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, !first || k != 0);
            }
            if constexpr (PAIR) tc_commit_pair(&empty_bar[stage]);     // the stage is free in BOTH CTAs
            else tc_commit(&empty_bar[stage]);
            if (last) {
              if constexpr (PAIR) tc_commit_pair(&tmem_full[as]);
              else tc_commit(&tmem_full[as]);
              if constexpr (WS) {
                if (n_next != n_blk) tc_commit(w_empty);   // last tile that reads this W slice
              }
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // Thread = accumulator row (TMEM lane); warps 4-7 own the left half of the tile's columns, warps
    // 8-11 the right half, 32 columns per chunk in registers.  Outputs leave through shared-memory
    // staging boxes and TMA bulk stores (p.tma_store): a thread=row epilogue that stores straight to
    // global memory touches 32 different 128-byte lines per instruction and saturates the LSU.
    const GemmEpilogue& e = p.epi;
    const int ew = warp - 4;
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int half = ew >> 2;        // column half this warp owns
    const int r_in_tile = q * 32 + lane;
    const bool do_ln = e.ln_gamma != nullptr;
    const bool two_pass = do_ln || e.l2norm;
    constexpr bool kContig = (Cfg::kChunks % 2) == 0;                 // BN = 256: chunks [4*half, 4*half + 4)
    const int n_my = kContig ? Cfg::kChunks / 2 : (Cfg::kChunks - half + 1) / 2;
    auto chunk_of = [&](int j) { return kContig ? half * (Cfg::kChunks / 2) + j : half + 2 * j; };
    const bool tma_out = p.tma_store != 0;
    // every warp owns its 32 rows of the staging boxes and issues its own bulk stores: no barrier wider
    // than the warp on the store path
    const bool issuer = lane == 0;
    uint8_t* stg_h = stage_out + ew * kWarpBoxBytes;                        // [32 rows][128 B] fp16 box (64 columns)
    uint8_t* stg_f = stage_out + (kEpiWarps + ew) * kWarpBoxBytes;          // [32 rows][128 B] fp32 box (32 columns)
    const int swz = lane & 7;

    // The CTA's shared memory leaves no L1: every __ldg in the epilogue is an L2 round trip (~1000 cycles under the
    // operand traffic).  The vectors every row needs — bias, LayerNorm gamma / beta — are copied to shared memory once.
    const bool bias_smem = e.bias != nullptr && p.N <= 2048;
    float* gam_s = vec_s + 2048;
    float* bet_s = gam_s + 256;
    {
      const int et = threadIdx.x - 4 * 32;
      if (bias_smem)
        for (int i = et; i < p.N; i += kEpiThreads) vec_s[i] = __ldg(e.bias + i);
      if (do_ln && et < BN) {
        gam_s[et] = __ldg(e.ln_gamma + et);
        bet_s[et] = __ldg(e.ln_beta + et);
      }
      named_bar_sync(1, kEpiThreads);
    }

    for (int64_t it = 0; it < my_tiles; ++it) {
      const int as = static_cast<int>(it & 1);
      const uint32_t aphase = static_cast<uint32_t>((it >> 1) & 1);
      int64_t m_blk;
      int n_blk;
      decode(it, m_blk, n_blk);
      const int64_t row0 = m_blk * p.m_stride;
      const int64_t grow = row0 + r_in_tile;
      const bool row_ok = r_in_tile < p.m_valid && grow < M;
      const int64_t srow = row_ok ? grow : 0;   // safe row for loads
      mbar_wait(&tmem_full[as], aphase);
      if (GEMM_TRACING) {      // warp-uniform branch; reconverge before the warp-synchronous TMEM loads
        if (threadIdx.x == 128 && it < 4) GEMM_TRACE(5 + 2 * static_cast<int>(it));
        __syncwarp();
      }
      tc_fence_after_sync();
      const uint32_t t_acc = tmem_base + as * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);
      float keep = 1.f;
      if (e.row_mask) keep = (row_ok && e.row_mask[srow] != 0.f) ? 1.f : 0.f;
      const int64_t hrow = e.h_row_idx ? static_cast<int64_t>(__ldg(e.h_row_idx + srow)) : grow;

      // final values of chunk j (columns col0..col0+31 of this thread's row) -> every requested output
      auto emit = [&](int j, int col0, float (&v)[32]) {
        if (e.row_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= keep;
        }
        if (GEMM_DBG(2)) {
          if (v[0] == 123456.f) stg_h[0] = 1;      // keep the values alive
          return;
        }
        if (tma_out) {
          const bool h_first = (j & 1) == 0;     // an fp16 box holds two chunks
          // the previous bulk stores of this half have finished READING the staging boxes
          if (GEMM_TRACING) { if (threadIdx.x == 128 && it == 1) GEMM_TRACE(18 + 5 * j); __syncwarp(); }
          if (e.out_f32 || h_first) {      // a box is about to be overwritten
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
          if (GEMM_TRACING) { if (threadIdx.x == 128 && it == 1) GEMM_TRACE(19 + 5 * j); __syncwarp(); }
          if (e.out_f32) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(stg_f + lane * 128 + ((i ^ swz) << 4)) =
                  make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (e.out_h) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 hi = make_uint4(pack_op2(v[8 * i], v[8 * i + 1]), pack_op2(v[8 * i + 2], v[8 * i + 3]),
                                          pack_op2(v[8 * i + 4], v[8 * i + 5]), pack_op2(v[8 * i + 6], v[8 * i + 7]));
              *reinterpret_cast<uint4*>(stg_h + lane * 128 + ((((j & 1) * 4 + i) ^ swz) << 4)) = hi;
              if (e.out_lo)     // low halves: what the fp16 rounding of the high halves dropped (fp32 box reused)
                *reinterpret_cast<uint4*>(stg_f + lane * 128 + ((((j & 1) * 4 + i) ^ swz) << 4)) = lo_of(hi, &v[8 * i]);
            }
          }
          if (e.out_f32 || !h_first) {      // a store follows: the generic-proxy writes must be visible to the TMA engine
            fence_proxy_async_smem();
            __syncwarp();
          }
          if (issuer && !GEMM_DBG(1) && (e.out_f32 || !h_first)) {
            const int32_t wrow = static_cast<int32_t>(row0) + q * 32;
            if (e.out_f32) tma_store_2d(&tmap_of, stg_f, col0, wrow);
            if (e.out_h && !h_first) tma_store_2d(&tmap_oh, stg_h, col0 - 32, wrow);
            if (e.out_lo && !h_first) tma_store_2d(&tmap_ol, stg_f, col0 - 32, wrow);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (GEMM_TRACING) { if (threadIdx.x == 128 && it == 1) GEMM_TRACE(20 + 5 * j); __syncwarp(); }
        } else if (row_ok) {
          if (e.out_h) {
            uint4* o = reinterpret_cast<uint4*>(e.out_h + hrow * e.ld_h + col0);
            uint4* ol = e.out_lo ? reinterpret_cast<uint4*>(e.out_lo + hrow * e.ld_h + col0) : nullptr;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 hi = make_uint4(pack_op2(v[8 * i], v[8 * i + 1]), pack_op2(v[8 * i + 2], v[8 * i + 3]),
                                          pack_op2(v[8 * i + 4], v[8 * i + 5]), pack_op2(v[8 * i + 6], v[8 * i + 7]));
              o[i] = hi;
              if (ol) ol[i] = lo_of(hi, &v[8 * i]);
            }
          }
          if (e.out_f32) {
            float4* o = reinterpret_cast<float4*>(e.out_f32 + grow * e.ld_f32 + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
        if (e.out2_h && row_ok) {     // second output out + add2 (DETR src + pos): direct stores
          const uint4* a4 = reinterpret_cast<const uint4*>(e.add2 + grow * e.add2_ld + col0);
          uint4* o = reinterpret_cast<uint4*>(e.out2_h + grow * e.ld_out2 + col0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 t = __ldg(a4 + i);
            const op2_t* h = reinterpret_cast<const op2_t*>(&t);
            float2 f0 = op2_to_f2(h[0]), f1 = op2_to_f2(h[1]);
            float2 f2 = op2_to_f2(h[2]), f3 = op2_to_f2(h[3]);
            o[i] = make_uint4(pack_op2(v[8 * i] + f0.x, v[8 * i + 1] + f0.y),
                              pack_op2(v[8 * i + 2] + f1.x, v[8 * i + 3] + f1.y),
                              pack_op2(v[8 * i + 4] + f2.x, v[8 * i + 5] + f2.y),
                              pack_op2(v[8 * i + 6] + f3.x, v[8 * i + 7] + f3.y));
          }
        }
      };

      float psum = 0.f, psq = 0.f;
      // Row operands of the epilogue — the residual (fp32 row, or an fp16 (hi, lo) pair) or else the position table row —
      // are fetched one 32-column chunk AHEAD of the chunk being processed: their L2 latency overlaps the TMEM read /
      // math / store of the previous chunk instead of stalling every chunk
      const bool res32 = e.residual != nullptr && e.residual_f32;
      const bool res16 = e.residual != nullptr && !e.residual_f32;
      const bool tab_pre = e.row_table != nullptr && e.residual == nullptr;
      const float* res_row = res32 ? static_cast<const float*>(e.residual) + srow * e.res_ld + n_blk * BN : nullptr;
      const op_t* res_hi = res16 ? static_cast<const op_t*>(e.residual) + srow * e.res_ld + n_blk * BN : nullptr;
      const op_t* res_lo = res16 && e.residual_lo ? e.residual_lo + srow * e.res_ld + n_blk * BN : nullptr;
      const float* tab_row = nullptr;
      if (e.row_table) {
        const int64_t trow = (e.row_src ? static_cast<int64_t>(__ldg(e.row_src + srow)) : srow) % e.row_mod;
        tab_row = e.row_table + trow * p.N + n_blk * BN;
      }
      uint4 pre[8];
      auto prefetch = [&](int c) {      // the row operand words of chunk c
        if (res32) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pre[i] = __ldg(reinterpret_cast<const uint4*>(res_row + c * 32) + i);
        } else if (res16) {
#pragma unroll
          for (int i = 0; i < 4; ++i) pre[i] = __ldg(reinterpret_cast<const uint4*>(res_hi + c * 32) + i);
          if (res_lo) {
#pragma unroll
            for (int i = 0; i < 4; ++i) pre[4 + i] = __ldg(reinterpret_cast<const uint4*>(res_lo + c * 32) + i);
          }
        } else if (tab_pre) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pre[i] = __ldg(reinterpret_cast<const uint4*>(tab_row + c * 32) + i);
        }
      };
      prefetch(chunk_of(0));
      // ---------- pass 1: x = act(acc + bias + table + residual); store or stash ----------
      for (int j = 0; j < n_my; ++j) {
        const int c = chunk_of(j);
        uint4 cur[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = pre[i];
        if (j + 1 < n_my) prefetch(chunk_of(j + 1));
        uint32_t acc[32];
        if (GEMM_TRACING) { if (threadIdx.x == 128 && it == 1) GEMM_TRACE(16 + 5 * j); __syncwarp(); }
        if (GEMM_DBG(4)) {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = static_cast<uint32_t>(i + c);
        } else {
          tmem_ld_x32(t_acc + c * 32, acc);
          tmem_wait_ld();
        }
        if (GEMM_TRACING) { if (threadIdx.x == 128 && it == 1) GEMM_TRACE(17 + 5 * j); __syncwarp(); }
        const int col0 = n_blk * BN + c * 32;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
        if (e.bias) {
          if (bias_smem) {
            const float4* b4 = reinterpret_cast<const float4*>(vec_s + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 t = b4[i];
              v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
            }
          } else {
            const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 t = __ldg(b4 + i);
              v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
            }
          }
        }
        if (e.row_table) {
          if (tab_pre) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i] += __uint_as_float(cur[i].x); v[4 * i + 1] += __uint_as_float(cur[i].y);
              v[4 * i + 2] += __uint_as_float(cur[i].z); v[4 * i + 3] += __uint_as_float(cur[i].w);
            }
          } else {
            const float4* t4 = reinterpret_cast<const float4*>(tab_row + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 t = __ldg(t4 + i);
              v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
            }
          }
        }
        if (res32) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[4 * i] += __uint_as_float(cur[i].x); v[4 * i + 1] += __uint_as_float(cur[i].y);
            v[4 * i + 2] += __uint_as_float(cur[i].z); v[4 * i + 3] += __uint_as_float(cur[i].w);
          }
        } else if (res16) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const op2_t* h = reinterpret_cast<const op2_t*>(&cur[i]);
            const op2_t* l = reinterpret_cast<const op2_t*>(&cur[4 + i]);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float2 f = op2_to_f2(h[jj]);
              v[8 * i + 2 * jj] += f.x;
              v[8 * i + 2 * jj + 1] += f.y;
            }
            if (res_lo) {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const float2 f = op2_to_f2(l[jj]);
                v[8 * i + 2 * jj] += f.x;
                v[8 * i + 2 * jj + 1] += f.y;
              }
            }
          }
        }
        if (e.act == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        } else if (e.act == 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (two_pass) {
          uint32_t st[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            psum += v[i];
            psq = fmaf(v[i], v[i], psq);
            st[i] = __float_as_uint(v[i]);
          }
          tmem_st_x32(t_acc + c * 32, st);
        } else {
          emit(j, col0, v);
        }
      }
      if (two_pass) {
        tmem_wait_st();
        // ---------- row statistics across the two column halves (sum and sum of squares) ----------
        float mean = 0.f, scale;
        ln_part[(0 * 2 + half) * 128 + r_in_tile] = psum;
        ln_part[(1 * 2 + half) * 128 + r_in_tile] = psq;
        named_bar_sync(1, kEpiThreads);
        const float tot = ln_part[(0 * 2 + 0) * 128 + r_in_tile] + ln_part[(0 * 2 + 1) * 128 + r_in_tile];
        const float totsq = ln_part[(1 * 2 + 0) * 128 + r_in_tile] + ln_part[(1 * 2 + 1) * 128 + r_in_tile];
        if (do_ln) {
          mean = tot * (1.0f / BN);
          const float var = fmaxf(totsq * (1.0f / BN) - mean * mean, 0.f);
          scale = rsqrtf(var + e.ln_eps);
        } else {
          scale = 1.0f / fmaxf(sqrtf(totsq), 1e-12f);   // F.normalize
        }
        // ---------- pass 2: normalise + store ----------
        for (int j = 0; j < n_my; ++j) {
          const int c = chunk_of(j);
          uint32_t acc[32];
          tmem_ld_x32(t_acc + c * 32, acc);
          tmem_wait_ld();
          const int col0 = n_blk * BN + c * 32;
          float v[32];
          if (do_ln) {
            const float4* g4 = reinterpret_cast<const float4*>(gam_s + c * 32);      // N == BN: column = c * 32
            const float4* b4 = reinterpret_cast<const float4*>(bet_s + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 g = g4[i], bb = b4[i];
              v[4 * i] = (__uint_as_float(acc[4 * i]) - mean) * scale * g.x + bb.x;
              v[4 * i + 1] = (__uint_as_float(acc[4 * i + 1]) - mean) * scale * g.y + bb.y;
              v[4 * i + 2] = (__uint_as_float(acc[4 * i + 2]) - mean) * scale * g.z + bb.z;
              v[4 * i + 3] = (__uint_as_float(acc[4 * i + 3]) - mean) * scale * g.w + bb.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * scale;
          }
          emit(j, col0, v);
        }
        // the partial-sum slots are reused by the next tile: all readers must be done
        named_bar_sync(1, kEpiThreads);
      }
      // release the accumulator stage
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));   // the leader's copy
        else mbar_arrive(&tmem_empty[as]);
      }
      if (GEMM_TRACING) {
        if (threadIdx.x == 128 && it < 4) GEMM_TRACE(6 + 2 * static_cast<int>(it));
        __syncwarp();
      }
    }
    // the staging boxes must outlive the bulk stores that read them
    if (GEMM_TRACING) {
      if (threadIdx.x == 128) GEMM_TRACE(13);
      __syncwarp();
    }
    if (tma_out && issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (GEMM_TRACING) {
      if (threadIdx.x == 128) GEMM_TRACE(14);
      __syncwarp();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();     // neither CTA leaves while the other may still touch its memory
  if (warp == 2) {
    tc_fence_after_sync();
    if constexpr (PAIR) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    if (threadIdx.x == 64) GEMM_TRACE(15);
  }
}

template <int BN, bool WS, bool PAIR = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& toh, const CUtensorMap& tof,
                       const CUtensorMap& tol, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, WS, PAIR>;
  static_assert(Cfg::kSmemBytes <= 232448, "shared memory budget of an sm_100 CTA");
  MADE_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<BN, WS, PAIR>), Cfg::kSmemBytes));
  const int64_t m_tiles = (p.M + p.m_stride - 1) / p.m_stride;
  ProfScope prof_scope(kProfGemm, stream);
  if constexpr (PAIR) {
    const int64_t n_ptiles = ((m_tiles + 1) / 2) * (p.N / BN);
    const int64_t max_pairs = sm_count() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * (n_ptiles < max_pairs ? n_ptiles : max_pairs)));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    MADE_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, WS, PAIR>, ta, tb, toh, tof, tol, p));
  } else {
    const int64_t n_tiles = m_tiles * (p.N / BN);
    int grid = static_cast<int>(n_tiles < sm_count() ? n_tiles : sm_count());
#ifdef MADE_GEMM_DIAG
    if (const char* g = getenv("MADE_GEMM_GRID")) grid = atoi(g) > 0 && atoi(g) < grid ? atoi(g) : grid;   // diagnostics: fewer CTAs
#endif
    gemm_tc_kernel<BN, WS, PAIR><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, toh, tof, tol, p);
  }
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

// MADE_GEMM_PAIR=1 turns the CTA-pair form on (read per call).  OFF by default: it is bit-identical and its mainloop is
// faster (7.6 k against 9.2 k cycles per split-2 tile on L2-resident operands), but a tile's EPILOGUE takes 9.5 - 12 k
// cycles and is what a tile waits for there, and on operands streamed from HBM both forms sit at the same 7 us per tile
// with the epilogue emptied — so it measures within +-3 % of the single-CTA form (scripts/diag_gemm_pair.py,
// diag_gemm_ablate.py, diag_gemm_trace2.py).
bool gemm_pair_enabled() {
  const char* v = getenv("MADE_GEMM_PAIR");
  return v && v[0] == '1';
}

// MADE_GEMM_WS128=1 turns the weight-stationary 128-column form of the split GEMMs on (read per call).  OFF by default:
// it moves 1.5x fewer bytes from L2 but is 2x SLOWER (N = 768, K = 256, split 2: 140 us against 72 us) — the mainloop
// is bound by the round trip of a ring stage (TMA latency under load + MMA + commit, ~2 us) times the number of k
// blocks over a 3-deep ring, not by bytes, and 128-column tiles double the k blocks per output.
bool gemm_ws128_enabled() {
  const char* v = getenv("MADE_GEMM_WS128");
  return v && v[0] == '1';
}

bool gemm_tma_store_enabled() {
  static const bool enabled = [] {
    const char* v = getenv("MADE_GEMM_TMA_STORE");
    return !(v && v[0] == '0');
  }();
  return enabled;
}

static long long* g_gemm_trace = nullptr;     // diagnostics only (made_debug_gemm_trace)

int gemm_f16_tc(const op_t* A, int64_t lda, const op_t* W, int64_t ldb,
                 int64_t w_rows, const GemmParams& p, int block_n, cudaStream_t stream) {
  if (p.M == 0) return MADE_OK;
  MADE_REQUIRE(A && W, "gemm: null operand");
  MADE_REQUIRE(block_n == 256 || block_n == 96, "gemm: block_n=%d unsupported", block_n);
  MADE_REQUIRE(p.N % block_n == 0, "gemm: N=%d not a multiple of %d", p.N, block_n);
  MADE_REQUIRE(p.K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "gemm: K/lda/ldb must be multiples of 8");
  MADE_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  const GemmEpilogue& e = p.epi;
  MADE_REQUIRE(!(e.ln_gamma || e.l2norm) || (p.N == block_n && block_n == 256),
               "gemm: LayerNorm/L2 epilogue needs N == 256");
  MADE_REQUIRE(!(e.ln_gamma && e.l2norm), "gemm: LayerNorm and L2 epilogues are exclusive");
  MADE_REQUIRE(!e.ln_gamma || e.ln_beta, "gemm: LayerNorm needs beta");
  MADE_REQUIRE(!e.out2_h || e.add2, "gemm: out2 needs add2");
  MADE_REQUIRE(e.out_h || e.out_f32 || e.out2_h, "gemm: no output");
  MADE_REQUIRE(p.m_valid >= 1 && p.m_valid <= 128 && p.m_stride >= 1 && p.m_stride <= 128,
               "gemm: bad tile geometry");
  MADE_REQUIRE(p.split >= 0 && p.split <= 2, "gemm: split=%d", p.split);
  MADE_REQUIRE(p.split == 0 || (p.K % kBlockK == 0 && ldb >= 2 * p.K && !p.b_batched && block_n == 256),
               "gemm: split operands need K %% 64 == 0 and [hi | lo] rows of 2K columns");
  MADE_REQUIRE(p.split != 2 || lda >= 2 * p.K, "gemm: split A needs lda >= 2K");
  MADE_REQUIRE(!e.out_lo || e.out_h, "gemm: out_lo needs out_h");
  CUtensorMap ta, tb, toh, tof, tol;
  memset(&toh, 0, sizeof(toh));
  memset(&tof, 0, sizeof(tof));
  memset(&tol, 0, sizeof(tol));
  MADE_TRY(encode_tmap_2d_16b(&ta, A, static_cast<uint64_t>(p.split == 2 ? 2 * p.K : p.K), static_cast<uint64_t>(p.M),
                               static_cast<uint64_t>(lda) * 2, kBlockK, kBlockM));
  MADE_TRY(encode_tmap_2d_16b(&tb, W, static_cast<uint64_t>(p.split ? 2 * p.K : p.K), static_cast<uint64_t>(w_rows),
                               static_cast<uint64_t>(ldb) * 2, kBlockK, static_cast<uint32_t>(block_n)));
  GemmParams pp = p;
  pp.trace = g_gemm_trace;
  {
    const char* pf = getenv("MADE_GEMM_L2_PREFETCH");
    pp.l2_prefetch = (pf && atoi(pf) != 0) ? 1 : 0;
    const char* dbg = getenv("MADE_GEMM_DEBUG");
    pp.debug = dbg ? atoi(dbg) : 0;
  }
  // outputs through TMA bulk stores whenever the tile is a plain [128 x 256] block of the output matrices
  const bool tma_store_enabled = gemm_tma_store_enabled();
  auto aligned16 = [](const void* ptr, int64_t ld, int esz) {
    return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * esz) % 16 == 0;
  };
  // split-precision GEMMs without a row-wide epilogue (LayerNorm / L2) and with a plain fp16 output run weight-
  // stationary on 128-column tiles: both halves of the weight pair of a tile column stay in shared memory
  const int64_t m_tiles_all = (p.M + p.m_stride - 1) / p.m_stride;
  const bool ws128 = block_n == 256 && p.split >= 1 && p.K <= kWsKBlocks * kBlockK && !p.b_batched && !e.ln_gamma && !e.l2norm &&
                     !e.out_f32 && !e.out_lo && e.out_h && p.N % 128 == 0 && p.n_store == 0 && p.m_valid == 128 && p.m_stride == 128 &&
                     m_tiles_all * (p.N / 128) >= 2 * sm_count() && gemm_ws128_enabled();
  if (ws128) MADE_TRY(encode_tmap_2d_16b(&tb, W, static_cast<uint64_t>(2 * p.K), static_cast<uint64_t>(w_rows),
                                          static_cast<uint64_t>(ldb) * 2, kBlockK, 128));
  pp.tma_store = tma_store_enabled && block_n == 256 && p.m_valid == 128 && p.m_stride == 128 && !e.h_row_idx &&
                 (e.out_h || e.out_f32) && (!e.out_h || aligned16(e.out_h, e.ld_h, 2)) &&
                 (!e.out_f32 || aligned16(e.out_f32, e.ld_f32, 4)) &&
                 (!e.out_lo || (aligned16(e.out_lo, e.ld_h, 2) && !e.out_f32));
  if (p.n_store > 0 && p.n_store < p.N && !(pp.tma_store && !e.out2_h && !e.ln_gamma && !e.l2norm)) {
    set_error("gemm: a clipped output width needs the plain TMA-store epilogue");
    return MADE_EUNSUPPORTED;
  }
  if (pp.tma_store) {
    const uint64_t n_ext = static_cast<uint64_t>(p.n_store > 0 && p.n_store < p.N ? p.n_store : p.N);
    if (e.out_h)
      MADE_TRY(encode_tmap_2d(&toh, e.out_h, 2, n_ext, static_cast<uint64_t>(p.M),
                              static_cast<uint64_t>(e.ld_h) * 2, 64, 32));
    if (e.out_f32)
      MADE_TRY(encode_tmap_2d(&tof, e.out_f32, 4, n_ext, static_cast<uint64_t>(p.M),
                              static_cast<uint64_t>(e.ld_f32) * 4, 32, 32));
    if (e.out_lo)
      MADE_TRY(encode_tmap_2d(&tol, e.out_lo, 2, n_ext, static_cast<uint64_t>(p.M),
                              static_cast<uint64_t>(e.ld_h) * 2, 64, 32));
  }
  if (ws128) return launch_gemm<128, true>(ta, tb, toh, tof, tol, pp, stream);
  if (block_n == 256) {
    // weight-stationary when the [256 x K] slice fits next to the A ring and every CTA gets >= 2 tiles;
    // its shared memory has no room for fp32 staging boxes, so fp32 outputs take the streaming variant
    const int64_t m_tiles = (p.M + p.m_stride - 1) / p.m_stride;
    const bool ws = p.K <= kWsKBlocks * kBlockK && !p.b_batched && m_tiles * (p.N / 256) >= 2 * sm_count() &&
                    !(pp.tma_store && (e.out_f32 || e.out_lo)) && p.split == 0;
    if (ws) return launch_gemm<256, true>(ta, tb, toh, tof, tol, pp, stream);
    // CTA pairs when there are at least two waves of [256 x 256] pair tiles: every CTA stages half of the W rows
    const bool pair = !p.b_batched && p.m_stride == 128 && p.m_valid == 128 && w_rows >= p.N &&
                      ((m_tiles + 1) / 2) * (p.N / 256) >= sm_count() && gemm_pair_enabled();
    if (pair) {
      MADE_TRY(encode_tmap_2d_16b(&tb, W, static_cast<uint64_t>(p.split ? 2 * p.K : p.K), static_cast<uint64_t>(w_rows),
                                   static_cast<uint64_t>(ldb) * 2, kBlockK, 128));
      return launch_gemm<256, false, true>(ta, tb, toh, tof, tol, pp, stream);
    }
    return launch_gemm<256, false>(ta, tb, toh, tof, tol, pp, stream);
  }
  return launch_gemm<96, false>(ta, tb, toh, tof, tol, pp, stream);
}

}  // namespace made

// Diagnostics (not part of include/made_b200.h): device buffer of 16 int64 clock stamps written by CTA 0 of the
// following GEMM launches; null turns tracing off.
extern "C" int made_debug_gemm_trace(void* dev_buf) {
  made::g_gemm_trace = static_cast<long long*>(dev_buf);
  return 0;
}
