// Fused feed-forward block:  out = epilogue( act(X W1^T + b1) W2^T + b2 + residual ),  256 -> 1024 -> 256,
// for the temporal encoders (GELU-erf, model/model_Base.py:74-80,89) and the DETR encoder layers (ReLU + post-norm,
// music_detr/transformer.py:204-209).  The [rows, 1024] hidden activation never leaves the SM: the two GEMMs of a
// 128-row tile are chained through shared memory, one 128-column slice of the hidden layer at a time.
//
// Per 128-row tile (persistent CTAs, static round-robin over the tiles):
//   X tile [128 x 256] fp16             -> shared memory, 4 k-slabs (TMA, 128B swizzle), A operand of GEMM 1
//   for j = 0..7 (hidden slices of 128):
//     H_j  = X W1_j^T                   tcgen05.mma 128 x 128 x 256  -> TMEM (double-buffered, 2 x 128 columns)
//     h_j  = act(H_j + b1_j) as fp16    epilogue warps: TMEM -> registers -> swizzled shared memory (A operand)
//     Y   += h_j W2_j^T                 tcgen05.mma 128 x 256 x 128  -> TMEM (256 columns)
//   out   = Y + b2 + residual [LayerNorm]   epilogue warps -> staging boxes -> TMA bulk stores
// The MMA issuer runs GEMM 1 one slice ahead (G1(j+1) is issued before G2(j)), so the activation epilogue of slice
// j overlaps tensor work on both sides of it.  W1 / W2 stream through a 4-stage ring of 32 KB (W1: two [128 x 64]
// k-tiles per stage, W2: one [256 x 64] k-tile per stage): 1 MB of weights per tile from L2, and the ring holds the
// operands of one G2 and one G1 step at once so that their L2 latency is hidden behind the previous step (a 3-stage
// ring with two hidden buffers measured 36 us per tile, bound by exactly that latency; r02 profiles).
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..11 =
// epilogue (thread = row, warps 4-7 the left half of the columns, warps 8-11 the right half).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "prep.cuh"

namespace made {

// MADE_FFN_DEBUG ablations (scripts/diag_ffn.py) exist only in a diagnostics build (MADE_DIAG=1)
#ifdef MADE_FFN_DIAG
#define FFN_DBG(bit) ((p.debug & (bit)) != 0)
#else
#define FFN_DBG(bit) false
#endif

namespace {

constexpr int kM = 128;            // rows per tile
constexpr int kD = 256;            // model width
constexpr int kHid = 1024;         // hidden width
constexpr int kSl = 128;           // hidden columns per slice
constexpr int kSlices = kHid / kSl;
constexpr int kStages = 4;
constexpr int kStageBytes = 32768;
constexpr int kThreads = 384;
constexpr int kEpiThreads = 256;
constexpr uint32_t kXBytes = kM * kD * 2;        // 64 KB: 4 k-slabs of [128 x 128 B]
constexpr uint32_t kHBytes = kM * kSl * 2;       // 32 KB: 2 k-slabs of [128 x 128 B]
constexpr uint32_t kSlab = kM * 128;             // 16 KB
constexpr uint32_t kSmem = kXBytes + kHBytes + kStages * kStageBytes + 1024 /*barriers*/ + 1024 /*alignment slack*/;
static_assert(kSmem <= 232448, "shared memory budget of an sm_100 CTA");

// TMEM columns
constexpr uint32_t kColH = 0;      // 2 x 128 fp32 columns (double-buffered H accumulator)
constexpr uint32_t kColY = 256;    // 256 fp32 columns

// GELU(erf) = relu(x) - |x|/2 * erfc(|x|/sqrt 2), erfc by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7) with the
// argument pre-scaled so that exp(-z^2) is a bare ex2: 13 FP32-pipe instructions + 2 MUFU per element.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  constexpr float kS = 0.8493218002880191f;            // sqrt(log2 e) / sqrt 2:  z' = |x| kS,  z'^2 = z^2 log2 e
  constexpr float kP = 0.3275911f / 1.2011224087864498f;   // p / sqrt(log2 e)
  const float ax = fabsf(x);
  const float zs = ax * kS;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-zs * zs));
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(kP, zs, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float y = p * t * e;                            // erfc(|x| / sqrt 2)
  return fmaf(-0.5f * ax, y, fmaxf(x, 0.f));
}

struct FfnParams {
  int64_t M;                    // upper bound of the row count (grid / TMA maps)
  const int32_t* m_dev;         // device scalar: rows actually present (nullable)
  const float* b1;              // [1024]
  const float* b2;              // [256]
  int act;                      // 1 GELU(erf), 2 ReLU
  const op_t* res_hi;           // residual (hi, lo) pair, row stride res_ld (lo nullable)
  const op_t* res_lo;
  int64_t res_ld;
  const float* ln_gamma;        // optional LayerNorm over the 256 outputs
  const float* ln_beta;
  float ln_eps;
  int has_lo;                   // write the low halves (out_lo tensor map)
  const op_t* add2;             // optional second output fp16(out + add2), direct stores
  int64_t add2_ld;
  op_t* out2;
  int64_t ld_out2;
  int debug;                    // bench-only ablations (MADE_FFN_DEBUG): 1 = no weight TMA traffic after the first ring
                                // fill, 2 = output epilogue without global loads / stores, 4 = identity activation
};

__global__ void __launch_bounds__(kThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                 const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_oh,
                 const __grid_constant__ CUtensorMap tm_ol, const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sH = sX + kXBytes;                   // fp16 hidden slice; reused as the output staging boxes at the end of a tile
  uint8_t* ring = sH + kHBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kStages * kStageBytes);
  uint64_t* full = bars;                        // [4] weight stage landed
  uint64_t* empty = bars + 4;                   // [4] MMAs that read the stage completed
  uint64_t* x_full = bars + 8;
  uint64_t* x_empty = bars + 9;                 // all GEMM-1 MMAs of the tile completed
  uint64_t* hacc_full = bars + 10;              // [2] H accumulator of a slice complete
  uint64_t* hacc_empty = bars + 12;             // [2] epilogue has read it (256 arrivals)
  uint64_t* h_full = bars + 14;                 // fp16 slice in shared memory (256 arrivals)
  uint64_t* h_empty = bars + 15;                // GEMM-2 MMAs that read it completed
  uint64_t* y_full = bars + 16;
  uint64_t* y_empty = bars + 17;                // Y drained by the epilogue (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  // LayerNorm partial sums [2 stats][2 halves][128] alias the first 2 KB of the hidden buffer, which is idle
  // between the last GEMM-2 MMA of a tile and the first staging-box write of its output epilogue
  float* ln_part = reinterpret_cast<float*>(sH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t M = p.m_dev ? static_cast<int64_t>(__ldg(p.m_dev)) : p.M;
  const int64_t m_tiles = (M + kM - 1) / kM;
  const int64_t my_tiles = m_tiles > blockIdx.x ? (m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2);
    tma_prefetch_desc(&tm_oh);
    if (p.has_lo) tma_prefetch_desc(&tm_ol);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hacc_full[i], 1);
      mbar_init(&hacc_empty[i], kEpiThreads);
    }
    mbar_init(h_full, kEpiThreads);
    mbar_init(h_empty, 1);
    mbar_init(y_full, 1);
    mbar_init(y_empty, kEpiThreads);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t fills = 0;
      bool skip = false;
      auto next_stage = [&]() -> uint8_t* {
        mbar_wait(&empty[stage], phase ^ 1);
        skip = FFN_DBG(1) && fills >= kStages;
        ++fills;
        if (skip) mbar_arrive(&full[stage]);      // ablation: stale weights, no L2 traffic
        else mbar_arrive_expect_tx(&full[stage], kStageBytes);
        return ring + stage * kStageBytes;
      };
      auto advance = [&]() { if (++stage == kStages) { stage = 0; phase ^= 1; } };
      auto load_w1 = [&](int j) {      // hidden rows [128 j, +128), 4 k-tiles of 64 -> 2 stages
        for (int h2 = 0; h2 < 2; ++h2) {
          uint8_t* s = next_stage();
          if (!skip) {
            tma_load_2d(s, &tm_w1, &full[stage], (2 * h2) * 64, j * kSl);
            tma_load_2d(s + 16384, &tm_w1, &full[stage], (2 * h2 + 1) * 64, j * kSl);
          }
          advance();
        }
      };
      auto load_w2 = [&](int j) {      // all 256 output rows, hidden columns [128 j, +128) -> 2 stages of one k-tile
        for (int h2 = 0; h2 < 2; ++h2) {
          uint8_t* s = next_stage();
          if (!skip) tma_load_2d(s, &tm_w2, &full[stage], j * kSl + h2 * 64, 0);
          advance();
        }
      };
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int32_t row0 = static_cast<int32_t>((blockIdx.x + it * gridDim.x) * kM);
        mbar_wait(x_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(x_full, kXBytes);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sX + kb * kSlab, &tm_x, x_full, kb * 64, row0);
        // same order as the MMA issuer consumes: G1(0) G1(1) [G2(j) G1(j+2)] ... G2(6) G2(7)
        load_w1(0);
        load_w1(1);
        for (int j = 0; j < kSlices; ++j) {
          load_w2(j);
          if (j + 2 < kSlices) load_w1(j + 2);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_f16(kM, kSl, 0, 0);     // H_j = X W1_j^T
      constexpr uint32_t idesc2 = umma_idesc_f16(kM, kD, 0, 0);      // Y  += h_j W2_j^T
      const uint32_t aX = smem_u32(sX), aH = smem_u32(sH);
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() { if (++stage == kStages) { stage = 0; phase ^= 1; } };
      for (int64_t it = 0; it < my_tiles; ++it) {
        auto g1 = [&](int j) {
          const int b = j & 1;
          const uint32_t use = static_cast<uint32_t>(it * (kSlices / 2) + (j >> 1));
          mbar_wait(&hacc_empty[b], (use & 1) ^ 1);
          tc_fence_after_sync();
          for (int h2 = 0; h2 < 2; ++h2) {
            mbar_wait(&full[stage], phase);
            tc_fence_after_sync();
            const uint32_t sb = smem_u32(ring + stage * kStageBytes);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int kb = 2 * h2 + t;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_ss(tmem_base + kColH + b * kSl, umma_smem_desc(aX + kb * kSlab + k * 32, 0, 1024),
                        umma_smem_desc(sb + t * 16384 + k * 32, 0, 1024), idesc1, (kb | k) != 0);
            }
            tc_commit(&empty[stage]);
            advance();
          }
          tc_commit(&hacc_full[b]);
          if (j == kSlices - 1) tc_commit(x_empty);
        };
        auto g2 = [&](int j) {
          const uint32_t n = static_cast<uint32_t>(it * kSlices + j);
          mbar_wait(h_full, n & 1);
          if (j == 0) mbar_wait(y_empty, (it & 1) ^ 1);
          tc_fence_after_sync();
          for (int h2 = 0; h2 < 2; ++h2) {
            mbar_wait(&full[stage], phase);
            tc_fence_after_sync();
            const uint32_t sb = smem_u32(ring + stage * kStageBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss(tmem_base + kColY, umma_smem_desc(aH + h2 * kSlab + k * 32, 0, 1024),
                      umma_smem_desc(sb + k * 32, 0, 1024), idesc2, (j | h2 | k) != 0);
            tc_commit(&empty[stage]);
            advance();
          }
          tc_commit(h_empty);
          if (j == kSlices - 1) tc_commit(y_full);
        };
        mbar_wait(x_full, it & 1);
        g1(0);
        g1(1);
        for (int j = 0; j < kSlices; ++j) {
          g2(j);
          if (j + 2 < kSlices) g1(j + 2);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    const int ew = warp - 4;
    const int q = warp & 3;            // TMEM lane quarter
    const int half = ew >> 2;          // column half
    const int r = q * 32 + lane;       // row of the tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int sw = r & 7;
    const bool issuer = lane == 0;
    // this warp's staging boxes, one 32-column chunk each: dense [32 rows][64 B] hi, then lo (un-swizzled
    // tensor maps: a 2 KB box is written once per tile chunk, the 4-way bank conflict of its 16-byte stores is noise)
    uint8_t* stg_h = sH + ew * 4096;
    uint8_t* stg_l = stg_h + 2048;
    uint32_t box_off[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) box_off[i] = static_cast<uint32_t>(lane) * 64u + static_cast<uint32_t>(i) * 16u;
    const bool do_ln = p.ln_gamma != nullptr;

    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t row0 = (blockIdx.x + it * gridDim.x) * kM;
      const int64_t grow = row0 + r;
      const bool row_ok = grow < M;
      const int64_t srow = row_ok ? grow : 0;
      // ---------- activation epilogue of the 8 hidden slices ----------
      for (int j = 0; j < kSlices; ++j) {
        const int b = j & 1;
        const uint32_t use = static_cast<uint32_t>(it * (kSlices / 2) + (j >> 1));
        mbar_wait(&hacc_full[b], use & 1);
        tc_fence_after_sync();
        uint32_t a0[32], a1[32];
        tmem_ld_x32(lane_addr + kColH + b * kSl + half * 64, a0);
        tmem_ld_x32(lane_addr + kColH + b * kSl + half * 64 + 32, a1);
        tmem_wait_ld();
        tc_fence_before_sync();
        mbar_arrive(&hacc_empty[b]);
        const float4* bias4 = reinterpret_cast<const float4*>(p.b1 + j * kSl + half * 64);
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 bb = __ldg(bias4 + i);
          const uint32_t* src = i < 8 ? &a0[4 * i] : &a1[4 * (i - 8)];
          float v0 = __uint_as_float(src[0]) + bb.x, v1 = __uint_as_float(src[1]) + bb.y;
          float v2 = __uint_as_float(src[2]) + bb.z, v3 = __uint_as_float(src[3]) + bb.w;
          if (FFN_DBG(4)) {
          } else if (p.act == 1) {
            v0 = gelu_erf_fast(v0); v1 = gelu_erf_fast(v1); v2 = gelu_erf_fast(v2); v3 = gelu_erf_fast(v3);
          } else {
            v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
          }
          pk[2 * i] = pack_op2(v0, v1);
          pk[2 * i + 1] = pack_op2(v2, v3);
        }
        // the GEMM-2 MMAs of the previous slice have finished reading the hidden buffer
        const uint32_t n = static_cast<uint32_t>(it * kSlices + j);
        mbar_wait(h_empty, (n & 1) ^ 1);
        // fp16 slice -> swizzled K-major A operand: this thread's 64 hidden columns = k-slab `half`, chunks 0..7
        uint8_t* hrow = sH + half * kSlab + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(hrow + ((c ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async_smem();
        mbar_arrive(h_full);
      }
      // ---------- output epilogue ----------
      // residual rows (thread = row: 16-byte pieces of 32 different rows per load instruction) are fetched one
      // 32-column chunk ahead; the first chunk is requested before the wait for the last GEMM-2 step
      uint4 rh[4], rl[4];
      auto res_issue = [&](int c) {
        if (p.res_hi) {
          const uint4* r4 = reinterpret_cast<const uint4*>(p.res_hi + srow * p.res_ld + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) rh[i] = __ldg(r4 + i);
          if (p.res_lo) {
            const uint4* l4 = reinterpret_cast<const uint4*>(p.res_lo + srow * p.res_ld + c * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) rl[i] = __ldg(l4 + i);
          }
        }
      };
#pragma unroll
      for (int i = 0; i < 4; ++i) { rh[i] = make_uint4(0u, 0u, 0u, 0u); rl[i] = make_uint4(0u, 0u, 0u, 0u); }
      if (!FFN_DBG(2)) res_issue(half * 4);
      mbar_wait(y_full, it & 1);
      tc_fence_after_sync();
      // every GEMM-2 MMA of the tile has completed: the hidden buffer is free and becomes the staging boxes
      float psum = 0.f, psq = 0.f;
      auto value_chunk = [&](int c, int c_next, float (&v)[32]) {      // acc + b2 + residual for columns [32 c, +32)
        uint4 ch[4], cl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ch[i] = rh[i]; cl[i] = rl[i]; }
        if (c_next >= 0) res_issue(c_next);      // the next chunk's residual rows travel while this chunk is processed
        uint32_t acc[32];
        tmem_ld_x32(lane_addr + kColY + c * 32, acc);
        tmem_wait_ld();
        const float4* b4 = reinterpret_cast<const float4*>(p.b2 + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = __ldg(b4 + i);
          v[4 * i] = __uint_as_float(acc[4 * i]) + t.x;
          v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + t.y;
          v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + t.z;
          v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + t.w;
        }
        if (p.res_hi) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const op2_t* h = reinterpret_cast<const op2_t*>(&ch[i]);
            const op2_t* l = reinterpret_cast<const op2_t*>(&cl[i]);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float2 f = op2_to_f2(h[jj]);
              const float2 g = op2_to_f2(l[jj]);      // zeros when there is no low half
              v[8 * i + 2 * jj] += f.x + g.x;
              v[8 * i + 2 * jj + 1] += f.y + g.y;
            }
          }
        }
      };
      // emit chunk c: fill the warp's 32-column boxes and hand them to the TMA engine
      auto emit = [&](int jl, int c, float (&v)[32]) {
        (void)jl;
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // previous stores have read the boxes
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 hi = make_uint4(pack_op2(v[8 * i], v[8 * i + 1]), pack_op2(v[8 * i + 2], v[8 * i + 3]),
                                      pack_op2(v[8 * i + 4], v[8 * i + 5]), pack_op2(v[8 * i + 6], v[8 * i + 7]));
          *reinterpret_cast<uint4*>(stg_h + box_off[i]) = hi;
          if (p.has_lo) {
            const op2_t* hh = reinterpret_cast<const op2_t*>(&hi);
            uint32_t lo[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float2 f = op2_to_f2(hh[jj]);
              lo[jj] = pack_op2(v[8 * i + 2 * jj] - f.x, v[8 * i + 2 * jj + 1] - f.y);
            }
            *reinterpret_cast<uint4*>(stg_l + box_off[i]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          if (p.out2 && row_ok) {
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(p.add2 + grow * p.add2_ld + c * 32) + i);
            const op2_t* h = reinterpret_cast<const op2_t*>(&t);
            const float2 f0 = op2_to_f2(h[0]), f1 = op2_to_f2(h[1]), f2 = op2_to_f2(h[2]), f3 = op2_to_f2(h[3]);
            reinterpret_cast<uint4*>(p.out2 + grow * p.ld_out2 + c * 32)[i] =
                make_uint4(pack_op2(v[8 * i] + f0.x, v[8 * i + 1] + f0.y), pack_op2(v[8 * i + 2] + f1.x, v[8 * i + 3] + f1.y),
                           pack_op2(v[8 * i + 4] + f2.x, v[8 * i + 5] + f2.y), pack_op2(v[8 * i + 6] + f3.x, v[8 * i + 7] + f3.y));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (issuer) {
          const int32_t wrow = static_cast<int32_t>(row0) + q * 32;
          tma_store_2d(&tm_oh, stg_h, c * 32, wrow);
          if (p.has_lo) tma_store_2d(&tm_ol, stg_l, c * 32, wrow);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      };
      if (FFN_DBG(2)) {
        // ablation: no output epilogue
      } else if (!do_ln) {
#pragma unroll 1
        for (int jl = 0; jl < 4; ++jl) {
          const int c = half * 4 + jl;
          float v[32];
          value_chunk(c, jl < 3 ? c + 1 : -1, v);
          emit(jl, c, v);
        }
      } else {
#pragma unroll 1
        for (int jl = 0; jl < 4; ++jl) {       // pass 1: values back into TMEM, row statistics
          const int c = half * 4 + jl;
          float v[32];
          value_chunk(c, jl < 3 ? c + 1 : -1, v);
          uint32_t st[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            psum += v[i];
            psq = fmaf(v[i], v[i], psq);
            st[i] = __float_as_uint(v[i]);
          }
          tmem_st_x32(lane_addr + kColY + c * 32, st);
        }
        tmem_wait_st();
        ln_part[(0 * 2 + half) * 128 + r] = psum;
        ln_part[(1 * 2 + half) * 128 + r] = psq;
        named_bar_sync(1, kEpiThreads);
        const float tot = ln_part[(0 * 2 + 0) * 128 + r] + ln_part[(0 * 2 + 1) * 128 + r];
        const float totsq = ln_part[(1 * 2 + 0) * 128 + r] + ln_part[(1 * 2 + 1) * 128 + r];
        const float mean = tot * (1.0f / kD);
        const float scale = rsqrtf(fmaxf(totsq * (1.0f / kD) - mean * mean, 0.f) + p.ln_eps);
        named_bar_sync(1, kEpiThreads);        // every partial sum has been read: the staging boxes may be written
#pragma unroll 1
        for (int jl = 0; jl < 4; ++jl) {       // pass 2: normalise + store
          const int c = half * 4 + jl;
          uint32_t acc[32];
          tmem_ld_x32(lane_addr + kColY + c * 32, acc);
          tmem_wait_ld();
          float v[32];
          const float4* g4 = reinterpret_cast<const float4*>(p.ln_gamma + c * 32);
          const float4* be4 = reinterpret_cast<const float4*>(p.ln_beta + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 g = __ldg(g4 + i), bb = __ldg(be4 + i);
            v[4 * i] = (__uint_as_float(acc[4 * i]) - mean) * scale * g.x + bb.x;
            v[4 * i + 1] = (__uint_as_float(acc[4 * i + 1]) - mean) * scale * g.y + bb.y;
            v[4 * i + 2] = (__uint_as_float(acc[4 * i + 2]) - mean) * scale * g.z + bb.z;
            v[4 * i + 3] = (__uint_as_float(acc[4 * i + 3]) - mean) * scale * g.w + bb.w;
          }
          emit(jl, c, v);
        }
      }
      // Y is drained: the next tile's GEMM 2 may overwrite it
      tc_fence_before_sync();
      mbar_arrive(y_empty);
      // the staging boxes alias the hidden buffer of the next tile: every warp's bulk stores must have
      // finished reading them before anyone writes the next h_0
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      named_bar_sync(1, kEpiThreads);
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// x [M, 256] fp16 (row stride ldx), w1 [1024, 256], w2 [256, 1024] fp16 row-major; out_hi / out_lo [M, 256] with row
// stride ld_out (out_lo nullable); residual (hi, lo) pair with stride res_ld (both nullable).
int ffn_fused(const op_t* x, int64_t ldx, const op_t* w1, const float* b1, const op_t* w2, const float* b2, int act,
              const op_t* res_hi, const op_t* res_lo, int64_t res_ld, const float* ln_gamma, const float* ln_beta,
              op_t* out_hi, op_t* out_lo, int64_t ld_out, const op_t* add2, int64_t add2_ld, op_t* out2,
              int64_t ld_out2, int64_t M, const int32_t* m_dev, cudaStream_t st) {
  if (M == 0) return MADE_OK;
  MADE_REQUIRE(x && w1 && b1 && w2 && b2 && out_hi, "ffn_fused: null pointer");
  MADE_REQUIRE(act == 1 || act == 2, "ffn_fused: act=%d", act);
  MADE_REQUIRE(ldx % 8 == 0 && ld_out % 8 == 0 && (!res_hi || res_ld % 8 == 0), "ffn_fused: strides must be multiples of 8");
  MADE_REQUIRE(!ln_gamma || ln_beta, "ffn_fused: LayerNorm needs beta");
  MADE_REQUIRE(!out2 || add2, "ffn_fused: out2 needs add2");
  MADE_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&ffn_fused_kernel), static_cast<int>(kSmem)));
  CUtensorMap tx, tw1, tw2, toh, tol;
  memset(&tol, 0, sizeof(tol));
  MADE_TRY(encode_tmap_2d_16b(&tx, x, kD, static_cast<uint64_t>(M), static_cast<uint64_t>(ldx) * 2, 64, kM));
  MADE_TRY(encode_tmap_2d_16b(&tw1, w1, kD, kHid, kD * 2, 64, kSl));
  MADE_TRY(encode_tmap_2d_16b(&tw2, w2, kHid, kD, kHid * 2, 64, kD));
  // output boxes of 32 columns x 32 rows, dense 64-byte rows
  MADE_TRY(encode_tmap_2d(&toh, out_hi, 2, kD, static_cast<uint64_t>(M), static_cast<uint64_t>(ld_out) * 2, 32, 32, 0));
  if (out_lo)
    MADE_TRY(encode_tmap_2d(&tol, out_lo, 2, kD, static_cast<uint64_t>(M), static_cast<uint64_t>(ld_out) * 2, 32, 32, 0));
  FfnParams p;
  p.M = M;
  p.m_dev = m_dev;
  p.b1 = b1;
  p.b2 = b2;
  p.act = act;
  p.res_hi = res_hi;
  p.res_lo = res_lo;
  p.res_ld = res_ld;
  p.ln_gamma = ln_gamma;
  p.ln_beta = ln_beta;
  p.ln_eps = 1e-5f;
  p.has_lo = out_lo != nullptr;
  p.add2 = add2;
  p.add2_ld = add2_ld;
  p.out2 = out2;
  p.ld_out2 = ld_out2;
  {
    const char* dbg = getenv("MADE_FFN_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  const int64_t m_tiles = (M + kM - 1) / kM;
  const int grid = static_cast<int>(m_tiles < sm_count() ? m_tiles : sm_count());
  ProfScope prof_scope(kProfFfn, st);
  ffn_fused_kernel<<<grid, kThreads, kSmem, st>>>(tx, tw1, tw2, toh, tol, p);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

}  // namespace made
