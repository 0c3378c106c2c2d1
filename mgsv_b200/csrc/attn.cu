// Multi-head self-attention core for the short sequences of MaDe (L = 50 / 96 / 146, 8 heads of
// 32 dims): softmax(Q K^T / sqrt(32) + key_padding_mask) V, as used by nn.MultiheadAttention at
// model_Base.py:87 and music_detr/transformer.py:196.
//
// These are ~5-9 % of the path's FLOPs in 32-wide heads, too small for a 128-row tcgen05 tile, so
// one warp owns one (sequence, head): the K and V head slices live row-major in padded shared memory
// (80-byte rows: conflict-free 16-byte stores, fragment loads and ldmatrix.trans), scores stay in
// registers (whole row, no online softmax), and the two products run on mma.sync m16n8k16 fp16 with
// fp32 accumulation.
// The decoder's single-query cross-attention (1 x 146 per head) is a separate SIMT kernel that
// works on the memory itself (K/V projections folded into the query side).
#include "common.cuh"

namespace made {

__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4],
                                               const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// 16-byte cp.async that writes zeros when !valid (src-size 0)
__device__ __forceinline__ void cp_async_16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
               "r"(valid ? 16 : 0)
               : "memory");
}

constexpr int kHeadDim = 32;
constexpr int kHeadsPerCta = 4;
constexpr int kWarpsPerHead = 2;   // the row tiles of a head alternate between two warps sharing its K/V slices
constexpr int kMhaThreads = kHeadsPerCta * kWarpsPerHead * 32;
constexpr int kKStride = 40;  // fp16 elements per K row in smem (80 B: conflict-free 4-byte frags)

template <int LP>  // padded length, multiple of 16
struct AttnSmem {
  static constexpr int kPerWarp = 2 * LP * kKStride;  // K and V head slices, both [key][32 (+8 pad)] fp16
  static constexpr int kBytes = kHeadsPerCta * kPerWarp * 2 + LP * 4;
};

// B fragments of P.V straight from the row-major V slice: two transposed 8x8 tiles (keys 0-7 / 8-15
// of the k-step, 8 head dims) -> {b0, b1} of mma.m16n8k16.  Lanes 0-15 supply the row addresses.
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&b)[2], const op_t* row_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(b[0]), "=r"(b[1])
               : "r"(smem_u32(row_ptr)));
}

template <int LP>
__global__ void __launch_bounds__(kMhaThreads, 2)
mha_core_kernel(const op_t* __restrict__ Q, int64_t ldq, const op_t* __restrict__ K,
                int64_t ldk, const op_t* __restrict__ V, int64_t ldv,
                const float* __restrict__ key_mask, int L_in, const int32_t* __restrict__ seq_off,
                const int32_t* __restrict__ seq_len, float scale, op_t* __restrict__ O, int64_t ldo) {
  // Dense batches: sequence b occupies rows [b*L, (b+1)*L) and key_mask marks the valid keys.
  // Ragged batches (seq_off != null): rows [seq_off[b], +seq_len[b]) and every key is valid.
  using S = AttnSmem<LP>;
  constexpr int NT = LP / 8, KK = LP / 16;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t b = blockIdx.x;
  const int hl = warp % kHeadsPerCta, part = warp / kHeadsPerCta;   // head within the CTA, which of its two warps
  const int h = blockIdx.y * kHeadsPerCta + hl;
  op_t* Ks = reinterpret_cast<op_t*>(attn_smem) + hl * S::kPerWarp;
  op_t* Vs = Ks + LP * kKStride;
  float* smask = reinterpret_cast<float*>(attn_smem + kHeadsPerCta * S::kPerWarp * 2);
  const int L = seq_off ? min(seq_len[b], LP) : L_in;
  const int64_t row0 = seq_off ? static_cast<int64_t>(seq_off[b]) : b * L_in;
  if (L <= 0) return;
  const float scale2 = scale * 1.4426950408889634f;

  for (int i = threadIdx.x; i < LP; i += blockDim.x)
    smask[i] = (i < L && (seq_off != nullptr || key_mask[row0 + i] != 0.f)) ? 0.f : -INFINITY;

  // ---- stage the K and V head slices (row-major); rows >= L are zero ----
  const op_t* Kg = K + row0 * ldk + h * kHeadDim;
  const op_t* Vg = V + row0 * ldv + h * kHeadDim;
  const int n_nt = (L + 7) >> 3, n_kk = (L + 15) >> 4;   // key tiles that hold at least one real key
  for (int idx = lane + 32 * part; idx < n_kk * 64; idx += 32 * kWarpsPerHead) {  // rows the MMAs touch; all copies in flight, one wait
    const int key = idx >> 2, ch = idx & 3;
    const bool ok = key < L;
    const int kr = ok ? key : 0;
    cp_async_16_zfill(Ks + key * kKStride + ch * 8, Kg + kr * ldk + ch * 8, ok);
    cp_async_16_zfill(Vs + key * kKStride + ch * 8, Vg + kr * ldv + ch * 8, ok);
  }
  // column 32 of the V slice (first pad column) = 1: the P.V MMA of a fifth n-tile then returns the row
  // sums of the ROUNDED weights, i.e. exactly the divisor that makes the weights used sum to one
  for (int key = lane + 32 * part; key < n_kk * 16; key += 32 * kWarpsPerHead) Vs[key * kKStride + kHeadDim] = f2op(1.0f);
  cp_async_wait_all();
  __syncthreads();

  const op_t* Qg = Q + row0 * ldq + h * kHeadDim;
  // Q fragments straight from global (each element is read exactly once), one row tile ahead
  auto load_q = [&](int rt, uint32_t (&qa)[2][4]) {
    const int r0 = rt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int c = ks * 16 + 2 * t;
      qa[ks][0] = r0 < L ? __ldg(reinterpret_cast<const uint32_t*>(Qg + r0 * ldq + c)) : 0u;
      qa[ks][1] = r1 < L ? __ldg(reinterpret_cast<const uint32_t*>(Qg + r1 * ldq + c)) : 0u;
      qa[ks][2] = r0 < L ? __ldg(reinterpret_cast<const uint32_t*>(Qg + r0 * ldq + c + 8)) : 0u;
      qa[ks][3] = r1 < L ? __ldg(reinterpret_cast<const uint32_t*>(Qg + r1 * ldq + c + 8)) : 0u;
    }
  };
  uint32_t qn[2][4];
  if (part * 16 < L) load_q(part, qn);
  for (int rt = part; rt < KK; rt += kWarpsPerHead) {
    const int r0 = rt * 16 + g, r1 = r0 + 8;
    if (rt * 16 >= L) break;
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) qa[ks][i] = qn[ks][i];
    if ((rt + kWarpsPerHead) * 16 < L) load_q(rt + kWarpsPerHead, qn);
    // ---- S = Q K^T ----
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt >= n_nt) continue;       // keys beyond the sequence: masked to -inf below
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t kb[2];
        const op_t* kp = Ks + (nt * 8 + g) * kKStride + ks * 16 + 2 * t;
        kb[0] = *reinterpret_cast<const uint32_t*>(kp);
        kb[1] = *reinterpret_cast<const uint32_t*>(kp + 8);
        mma_f16_16816(s[nt], qa[ks], kb);
      }
    }
    // ---- masked softmax over the whole row (rows g and g+8 of this tile) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float m0 = smask[nt * 8 + 2 * t], m1 = smask[nt * 8 + 2 * t + 1];
      s[nt][0] = fmaf(s[nt][0], scale2, m0);      // logits in log2 units: exp(x) = 2^(x * log2 e)
      s[nt][1] = fmaf(s[nt][1], scale2, m1);
      s[nt][2] = fmaf(s[nt][2], scale2, m0);
      s[nt][3] = fmaf(s[nt][3], scale2, m1);
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    uint32_t pa[KK][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      // unnormalised weights, rounded to fp16 for the MMA
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_op2(exp2_approx(s[nt][0] - mx0), exp2_approx(s[nt][1] - mx0));
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_op2(exp2_approx(s[nt][2] - mx1), exp2_approx(s[nt][3] - mx1));
    }
    // ---- [O | row sums] = P [V | 1] ----
    float o[5][4];
#pragma unroll
    for (int nd = 0; nd < 5; ++nd) {
      o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        if (kk >= n_kk) continue;     // all-zero probabilities
        uint32_t vb[2];
        ldmatrix_x2_trans(vb, Vs + (kk * 16 + (lane & 15)) * kKStride + nd * 8);
        mma_f16_16816(o[nd], pa[kk], vb);
      }
    }
    // column 0 of the fifth n-tile lives in the lanes with t == 0 (rows g and g + 8)
    const float sum0 = __shfl_sync(0xffffffffu, o[4][0], lane & ~3);
    const float sum1 = __shfl_sync(0xffffffffu, o[4][2], lane & ~3);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    op_t* Og = O + row0 * ldo + h * kHeadDim;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      const int c = nd * 8 + 2 * t;
      if (r0 < L) *reinterpret_cast<uint32_t*>(Og + r0 * ldo + c) = pack_op2(o[nd][0] * inv0, o[nd][1] * inv0);
      if (r1 < L) *reinterpret_cast<uint32_t*>(Og + r1 * ldo + c) = pack_op2(o[nd][2] * inv1, o[nd][3] * inv1);
    }
  }
}

// Decoder cross-attention with ONE query per sequence (music_detr/transformer.py:289-292) with the
// per-layer K/V projections of the memory folded into the query side (api.cu load_detr):
//   scores_h[t] = q~_h . (memory + pos)_t          q~ [B, 8*256] fp32 (includes 1/sqrt(32))
//   mbar_h      = sum_t softmax_t(scores_h)[t] * memory_t   -> out [B, 8*256] fp16
// so every layer reads the SAME two [len,256] fp16 matrices of a sequence and the [B*146, 6*512] K/V
// tensors of the reference are never formed.
// One CTA per sequence.  All 256 threads stage the valid rows of (memory + pos) in shared memory with
// cp.async; the 8 heads form the 8 real rows of an m16n8k16 A tile (rows 8-15 are zero), so
//   S[8 x len]  = Q~[8 x 256] . MP^T      (q~ as an fp16 hi + lo pair: two MMAs per k-step)
//   O[8 x 256]  = P[8 x len]  . MEM       (P = un-normalised exps in fp16, B via ldmatrix.trans)
// run on mma.sync; the same row buffer is refilled with the memory rows while the softmax runs.
constexpr int kDecMaxRows = 160;
constexpr int kDecPitch = 264;                          // fp16 elements per staged row (528 B: conflict-free)
constexpr int kDecRowsBytes = kDecMaxRows * kDecPitch * 2;
constexpr int kDecQBytes = 2 * 8 * kDecPitch * 2;       // q~ hi / lo, 8 heads
constexpr int kDecPPitch = kDecMaxRows + 8;             // fp16 elements per P row
constexpr int kDecSmemBytes = kDecRowsBytes + kDecQBytes + 8 * kDecPPitch * 2;

__global__ void __launch_bounds__(256)
dec_attn_folded_kernel(const float* __restrict__ qt, const op_t* __restrict__ mp,
                       const op_t* __restrict__ mem, const float* __restrict__ key_mask, int L,
                       const int32_t* __restrict__ seq_off, const int32_t* __restrict__ seq_len,
                       op_t* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t dec_smem[];
  op_t* rows = reinterpret_cast<op_t*>(dec_smem);                               // [nvp][264]
  op_t* qhi = reinterpret_cast<op_t*>(dec_smem + kDecRowsBytes);                // [8][264]
  op_t* qlo = qhi + 8 * kDecPitch;                                              // [8][264]
  op_t* ps = reinterpret_cast<op_t*>(dec_smem + kDecRowsBytes + kDecQBytes);    // [8][168] exps (fp16)
  __shared__ float sc[8][kDecMaxRows];
  __shared__ float ssum[8];
  __shared__ short vidx[kDecMaxRows];
  __shared__ int s_nvalid;
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t row0 = seq_off ? static_cast<int64_t>(seq_off[b]) : b * L;
  if (seq_off) {     // ragged batch: the sequence's rows are exactly its valid keys
    const int n = min(seq_len[b], kDecMaxRows);
    for (int i = tid; i < n; i += 256) vidx[i] = static_cast<short>(i);
    if (tid == 0) s_nvalid = n;
  } else if (warp == 0) {   // ordered compaction of the valid key positions
    int base = 0;
    for (int t0 = 0; t0 < L; t0 += 32) {
      const int tt = t0 + lane;
      const bool ok = tt < L && key_mask[b * L + tt] != 0.f;
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (ok) vidx[base + __popc(bal & ((1u << lane) - 1u))] = static_cast<short>(tt);
      base += __popc(bal);
    }
    if (lane == 0) s_nvalid = base;
  }
  // q~ -> fp16 hi + lo in shared memory (warp = head, lane = 8 features)
  {
    const float4* qp = reinterpret_cast<const float4*>(qt + b * 2048 + warp * 256 + lane * 8);
    const float4 a = __ldg(qp), c = __ldg(qp + 1);
    const float q[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const op2_t h = floats2op2(q[2 * j], q[2 * j + 1]);
      const float2 hf = op2_to_f2(h);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = pack_op2(q[2 * j] - hf.x, q[2 * j + 1] - hf.y);
    }
    *reinterpret_cast<uint4*>(qhi + warp * kDecPitch + lane * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(qlo + warp * kDecPitch + lane * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  __syncthreads();
  const int nv = s_nvalid;
  const int nvp = (nv + 15) & ~15;                       // keys padded to the MMA k-step
  // ---- stage (memory + pos) rows; rows [nv, nvp) are zero-filled ----
  for (int i = warp; i < nvp; i += 8)
    cp_async_16_zfill(rows + i * kDecPitch + lane * 8, mp + (row0 + (i < nv ? vidx[i] : 0)) * 256 + lane * 8, i < nv);
  cp_async_wait_all();
  __syncthreads();
  // ---- S = Q~ MP^T: warp w takes key tiles w, w+8, ... (8 keys each) ----
  for (int nt = warp; nt * 8 < nv; nt += 8) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    const op_t* krow = rows + (nt * 8 + g) * kDecPitch + 2 * t;
    const op_t* qh = qhi + g * kDecPitch + 2 * t;
    const op_t* ql = qlo + g * kDecPitch + 2 * t;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      uint32_t kb[2], ah[4], al[4];
      kb[0] = *reinterpret_cast<const uint32_t*>(krow + ks * 16);
      kb[1] = *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 8);
      ah[0] = *reinterpret_cast<const uint32_t*>(qh + ks * 16);
      ah[2] = *reinterpret_cast<const uint32_t*>(qh + ks * 16 + 8);
      al[0] = *reinterpret_cast<const uint32_t*>(ql + ks * 16);
      al[2] = *reinterpret_cast<const uint32_t*>(ql + ks * 16 + 8);
      ah[1] = ah[3] = al[1] = al[3] = 0u;                 // rows 8..15 of the A tile are zero
      mma_f16_16816(c, ah, kb);
      mma_f16_16816(c, al, kb);
    }
    const int key = nt * 8 + 2 * t;                       // c[0], c[1]: head g, keys key, key + 1
    if (key < nv) sc[g][key] = c[0];
    if (key + 1 < nv) sc[g][key + 1] = c[1];
  }
  __syncthreads();                 // scores complete; every warp is done with the (memory + pos) rows
  for (int i = warp; i < nvp; i += 8)
    cp_async_16_zfill(rows + i * kDecPitch + lane * 8, mem + (row0 + (i < nv ? vidx[i] : 0)) * 256 + lane * 8, i < nv);
  // ---- softmax of head `warp` (overlaps the refill): un-normalised exps -> fp16 ----
  {
    float mx = -INFINITY;
    for (int i = lane; i < nv; i += 32) mx = fmaxf(mx, sc[warp][i]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < nvp; i += 32) {
      const op_t e = i < nv ? f2op(fast_exp(sc[warp][i] - mx)) : f2op(0.f);
      ps[warp * kDecPPitch + i] = e;
      sum += op2f(e);                                      // divisor = sum of the ROUNDED weights
    }
    sum = warp_sum(sum);
    if (lane == 0) ssum[warp] = sum;
  }
  cp_async_wait_all();
  __syncthreads();
  // ---- O = P MEM: warp w takes feature tiles 4w .. 4w+3 (8 features each) ----
  const op_t* prow = ps + g * kDecPPitch + 2 * t;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int nd = warp * 4 + j;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kk = 0; kk * 16 < nvp; ++kk) {
      uint32_t a[4], vb[2];
      a[0] = *reinterpret_cast<const uint32_t*>(prow + kk * 16);
      a[2] = *reinterpret_cast<const uint32_t*>(prow + kk * 16 + 8);
      a[1] = a[3] = 0u;
      ldmatrix_x2_trans(vb, rows + (kk * 16 + (lane & 15)) * kDecPitch + nd * 8);
      mma_f16_16816(c, a, vb);
    }
    const float inv = 1.f / ssum[g];                       // c[0], c[1]: head g, features nd*8 + 2t, +1
    *reinterpret_cast<uint32_t*>(out + b * 2048 + g * 256 + nd * 8 + 2 * t) = pack_op2(c[0] * inv, c[1] * inv);
  }
}

}  // namespace made

using namespace made;

template <int LP>
static int launch_mha(const op_t* Q, int64_t ldq, const op_t* K, int64_t ldk,
                      const op_t* V, int64_t ldv, const float* mask, int64_t B, int L, const int32_t* seq_off,
                      const int32_t* seq_len, op_t* O, int64_t ldo, cudaStream_t st) {
  MADE_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&mha_core_kernel<LP>), static_cast<int>(AttnSmem<LP>::kBytes)));
  dim3 grid(static_cast<unsigned>(B), 8 / kHeadsPerCta);
  ProfScope prof_scope(kProfAttn, st);
  mha_core_kernel<LP><<<grid, kMhaThreads, AttnSmem<LP>::kBytes, st>>>(
      Q, ldq, K, ldk, V, ldv, mask, L, seq_off, seq_len, 0.17677669529663687f /* 1/sqrt(32) */, O, ldo);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

namespace made {
int mha_core(const op_t* Q, int64_t ldq, const op_t* K, int64_t ldk,
             const op_t* V, int64_t ldv, const float* key_mask, int64_t B, int L,
             op_t* O, int64_t ldo, cudaStream_t st, const int32_t* seq_off, const int32_t* seq_len) {
  if (B == 0) return MADE_OK;
  MADE_REQUIRE(Q && K && V && O && (key_mask || (seq_off && seq_len)), "mha_core: null pointer");
  MADE_REQUIRE(L > 0 && L <= 160, "mha_core: L=%d unsupported (max 160)", L);
  MADE_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "mha_core: bad strides");
  if (L <= 64) return launch_mha<64>(Q, ldq, K, ldk, V, ldv, key_mask, B, L, seq_off, seq_len, O, ldo, st);
  if (L <= 96) return launch_mha<96>(Q, ldq, K, ldk, V, ldv, key_mask, B, L, seq_off, seq_len, O, ldo, st);
  return launch_mha<160>(Q, ldq, K, ldk, V, ldv, key_mask, B, L, seq_off, seq_len, O, ldo, st);
}

int dec_attn_folded(const float* qt, const op_t* mp, const op_t* mem, const float* key_mask, int64_t B,
                    int L, op_t* out, cudaStream_t st, const int32_t* seq_off, const int32_t* seq_len) {
  if (B == 0) return MADE_OK;
  MADE_REQUIRE(qt && mp && mem && (key_mask || (seq_off && seq_len)) && out && L > 0 && L <= 160,
               "dec_attn_folded: bad arguments");
  MADE_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&dec_attn_folded_kernel), static_cast<int>(kDecSmemBytes)));
  ProfScope prof_scope(kProfAttn, st);
  dec_attn_folded_kernel<<<static_cast<unsigned>(B), 256, kDecSmemBytes, st>>>(qt, mp, mem, key_mask, L, seq_off,
                                                                              seq_len, out);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}
}  // namespace made
