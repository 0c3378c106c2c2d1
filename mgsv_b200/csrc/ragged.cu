// Ragged (token-packed) batches.  The reference computes every padded position densely and zeroes
// it afterwards (model_Base.py:533-541); padded keys are masked out of every softmax
// and padded query rows never reach an output, so only the VALID tokens of a batch need to exist.
// Mean valid lengths of MGSV-EC are 23/50 frames and 56/96 segments: packing the valid tokens into
// a dense [total, features] matrix removes ~45 % of all GEMM rows and attention work.
//
//   seq_len[b]  = number of valid tokens of sequence b          (mask != 0)
//   seq_off[b]  = exclusive prefix sum of seq_len               (first packed row of sequence b)
//   total[0]    = sum of seq_len                                (device scalar: no host sync)
//   tok_src[i]  = b * L + t of packed row i                     (position t = tok_src % L)
//
// Host code never learns `total`: GEMM grids are sized for the padded upper bound and the kernels
// read the device scalar.
#include "common.cuh"
#include "prep.cuh"

namespace made {

// warp per sequence
__global__ void ragged_len_kernel(const float* __restrict__ mask, int64_t B, int L, int32_t* __restrict__ seq_len) {
  const int lane = threadIdx.x & 31;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int n = 0;
  for (int t0 = 0; t0 < L; t0 += 32) {
    const int t = t0 + lane;
    n += __popc(__ballot_sync(0xffffffffu, t < L && mask[b * L + t] != 0.f));
  }
  if (lane == 0) seq_len[b] = n;
}

// single block: exclusive scan of seq_len -> seq_off, total
__global__ void __launch_bounds__(1024)
ragged_scan_kernel(const int32_t* __restrict__ seq_len, int64_t B, int32_t* __restrict__ seq_off,
                   int32_t* __restrict__ total) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < B; base += 1024) {
    const int64_t b = base + tid;
    const int v = b < B ? seq_len[b] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      warp_tot[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const int before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + incl - v;
    if (b < B) seq_off[b] = before;
    __syncthreads();
    if (tid == 1023) carry = before + v;
    __syncthreads();
  }
  if (tid == 0) *total = carry;
}

// warp per sequence: ordered compaction of the valid positions
__global__ void ragged_fill_kernel(const float* __restrict__ mask, int64_t B, int L,
                                   const int32_t* __restrict__ seq_off, int32_t* __restrict__ tok_src) {
  const int lane = threadIdx.x & 31;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int base = seq_off[b];
  for (int t0 = 0; t0 < L; t0 += 32) {
    const int t = t0 + lane;
    const bool ok = t < L && mask[b * L + t] != 0.f;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (ok) tok_src[base + __popc(bal & ((1u << lane) - 1u))] = static_cast<int32_t>(b * L + t);
    base += __popc(bal);
  }
}

// packed row i <- cast(feats[tok_src[i], :])   (model_Base.py:556/595 masked_fill + cast, valid rows only)
template <int kIn>
__global__ void ingest_gather_kernel(const void* __restrict__ in_, const int32_t* __restrict__ tok_src,
                                     const int32_t* __restrict__ total, int dim, op_t* __restrict__ out, int split) {
  const int vec = dim / 8;
  const int64_t n = static_cast<int64_t>(*total) * vec;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / vec;
    const int c = static_cast<int>(i % vec) * 8;
    const int64_t src = tok_src[row];
    float v[8];
    if constexpr (kIn == MADE_DTYPE_F32) {
      const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(in_) + src * dim + c);
      const float4 a = __ldcs(p), b = __ldcs(p + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if constexpr (kIn == MADE_DTYPE_BF16) {
      const uint4 a = __ldcs(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(in_) + src * dim + c));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
    } else {
      const uint4 a = __ldcs(reinterpret_cast<const uint4*>(static_cast<const __half*>(in_) + src * dim + c));
      const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
    }
    const uint4 hi = make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]), pack_op2(v[4], v[5]), pack_op2(v[6], v[7]));
    if (!split) {
      *reinterpret_cast<uint4*>(out + row * dim + c) = hi;
    } else {     // rows of [hi | lo]
      const op2_t* hh = reinterpret_cast<const op2_t*>(&hi);
      uint32_t lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = op2_to_f2(hh[j]);
        lo[j] = pack_op2(v[2 * j] - f.x, v[2 * j + 1] - f.y);
      }
      *reinterpret_cast<uint4*>(out + row * 2 * dim + c) = hi;
      *reinterpret_cast<uint4*>(out + row * 2 * dim + dim + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// pooled[b] = normalize( sum of the sequence's packed rows / seq_len[b] )  (model_Base.py:579-580 / :615-616)
__global__ void __launch_bounds__(256)
pool_norm_ragged_kernel(const float* __restrict__ seq, const int32_t* __restrict__ seq_off,
                        const int32_t* __restrict__ seq_len, float* __restrict__ pooled) {
  __shared__ float red[8];
  const int64_t b = blockIdx.x;
  const int d = threadIdx.x;
  const int n = seq_len[b];
  const float* p = seq + static_cast<int64_t>(seq_off[b]) * 256 + d;
  float acc = 0.f;
  for (int t = 0; t < n; ++t) acc += p[static_cast<int64_t>(t) * 256];
  const float v = acc / static_cast<float>(n);
  const float s = warp_sum(v * v);
  if ((d & 31) == 0) red[d >> 5] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  pooled[b * 256 + d] = v / fmaxf(sqrtf(tot), 1e-12f);
}

// padded[tok_src[i], :] = packed[i, :]  (256 fp32 features; the padded buffer was zero-filled)
__global__ void scatter_rows_f32_kernel(const float* __restrict__ packed, const int32_t* __restrict__ tok_src,
                                        const int32_t* __restrict__ total, float* __restrict__ padded) {
  const int lane = threadIdx.x & 31;
  const int n = *total;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n;
       row += static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5)) {
    const float4* s = reinterpret_cast<const float4*>(packed + row * 256 + lane * 8);
    float4* d = reinterpret_cast<float4*>(padded + static_cast<int64_t>(tok_src[row]) * 256 + lane * 8);
    d[0] = s[0];
    d[1] = s[1];
  }
}

// padded[tok_src[i], :] = hi[i, :] + lo[i, :]  (fp16 pair rows of 256 -> fp32)
__global__ void scatter_rows_pair_kernel(const op_t* __restrict__ hi, const op_t* __restrict__ lo,
                                         const int32_t* __restrict__ tok_src, const int32_t* __restrict__ total,
                                         float* __restrict__ padded) {
  const int lane = threadIdx.x & 31;
  const int n = *total;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n;
       row += static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5)) {
    const uint4 h = *reinterpret_cast<const uint4*>(hi + row * 256 + lane * 8);
    const op2_t* hh = reinterpret_cast<const op2_t*>(&h);
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = op2_to_f2(hh[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
    if (lo) {
      const uint4 l = *reinterpret_cast<const uint4*>(lo + row * 256 + lane * 8);
      const op2_t* ll = reinterpret_cast<const op2_t*>(&l);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = op2_to_f2(ll[j]); v[2 * j] += f.x; v[2 * j + 1] += f.y; }
    }
    float4* d = reinterpret_cast<float4*>(padded + static_cast<int64_t>(tok_src[row]) * 256 + lane * 8);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ---- DETR input assembly on packed tokens (model_Uni.py:207-216 + position_encoding.py:51-71) ----
// mask[b] = cat(frame_mask[b], seg_mask[track(b)])   [B,146]
__global__ void detr_mask_kernel(const float* __restrict__ frame_mask, const float* __restrict__ seg_mask,
                                 const int32_t* __restrict__ track_idx, int64_t seq_offset, int64_t B,
                                 float* __restrict__ mask_out) {
  constexpr int LV = 50, LM = 96, L = 146;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  const int64_t b = i / L;
  const int t = static_cast<int>(i % L);
  const int64_t tr = track_idx ? track_idx[b] : seq_offset + b;
  mask_out[i] = t < LV ? frame_mask[b * LV + t] : seg_mask[tr * LM + (t - LV)];
}

// Warp per packed token i = (b, t): src row from the frame / segment features; the sine position
// embedding depends only on the rank j of the token among its sequence's valid tokens:
//   x = cumsum(mask)[t] = j + 1;  x / (n + 1e-6) * 2pi;  pos[2k] = sin(x / dim_t[2k]), pos[2k+1] = cos(...)
__global__ void __launch_bounds__(256)
detr_prep_ragged_kernel(const op_t* __restrict__ frame_out, const op_t* __restrict__ seg_out,
                        const int32_t* __restrict__ track_idx, int64_t seq_offset,
                        const int32_t* __restrict__ tok_src, const int32_t* __restrict__ seq_off,
                        const int32_t* __restrict__ seq_len, const int32_t* __restrict__ total,
                        const float* __restrict__ inv_dim_t, op_t* __restrict__ src, op_t* __restrict__ pos,
                        op_t* __restrict__ srcpos) {
  constexpr int LV = 50, LM = 96, L = 146;
  const int lane = threadIdx.x & 31;
  const int n_tok = *total;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_tok;
       i += static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5)) {
    const int s = tok_src[i];
    const int64_t b = s / L;
    const int t = s % L;
    const int64_t tr = track_idx ? track_idx[b] : seq_offset + b;
    const op_t* sp = t < LV ? frame_out + (b * LV + t) * 256 : seg_out + (tr * LM + (t - LV)) * 256;
    const uint4 raw = *reinterpret_cast<const uint4*>(sp + lane * 8);
    const op2_t* h = reinterpret_cast<const op2_t*>(&raw);
    const float cum = static_cast<float>(static_cast<int>(i) - seq_off[b] + 1);
    const float x = cum / (static_cast<float>(seq_len[b]) + 1e-6f) * 6.283185307179586f;
    float pv[8], sv[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = x * inv_dim_t[lane * 8 + 2 * j];   // same dim_t for the (sin, cos) pair
      pv[2 * j] = sinf(a);
      pv[2 * j + 1] = cosf(a);
      const float2 f = op2_to_f2(h[j]);
      sv[2 * j] = f.x;
      sv[2 * j + 1] = f.y;
    }
    const int64_t o = i * 256 + lane * 8;
    *reinterpret_cast<uint4*>(src + o) = raw;
    *reinterpret_cast<uint4*>(pos + o) = make_uint4(pack_op2(pv[0], pv[1]), pack_op2(pv[2], pv[3]),
                                                    pack_op2(pv[4], pv[5]), pack_op2(pv[6], pv[7]));
    *reinterpret_cast<uint4*>(srcpos + o) =
        make_uint4(pack_op2(sv[0] + pv[0], sv[1] + pv[1]), pack_op2(sv[2] + pv[2], sv[3] + pv[3]),
                   pack_op2(sv[4] + pv[4], sv[5] + pv[5]), pack_op2(sv[6] + pv[6], sv[7] + pv[7]));
  }
}

// row_off[b] = seq_off[b] + base, row_len[b] = seq_len[b]  (sequence rows inside a multi-chunk buffer)
__global__ void offset_rows_kernel(const int32_t* __restrict__ seq_off, const int32_t* __restrict__ seq_len,
                                   int64_t B, int32_t base, int32_t* __restrict__ row_off,
                                   int32_t* __restrict__ row_len) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= B) return;
  row_off[b] = seq_off[b] + base;
  row_len[b] = seq_len[b];
}

// ---- launchers ---------------------------------------------------------------------------
int offset_rows(const Ragged& rb, int32_t base, int32_t* row_off, int32_t* row_len, cudaStream_t st) {
  offset_rows_kernel<<<static_cast<unsigned>(ceil_div64(rb.B, 256)), 256, 0, st>>>(rb.seq_off, rb.seq_len, rb.B, base,
                                                                                  row_off, row_len);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int scatter_rows_f32_nozero(const float* packed, const Ragged& rb, float* padded, cudaStream_t st) {
  scatter_rows_f32_kernel<<<static_cast<unsigned>(sm_count() * 4), 256, 0, st>>>(packed, rb.tok_src, rb.total, padded);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int scatter_rows_pair_nozero(const op_t* hi, const op_t* lo, const Ragged& rb, float* padded, cudaStream_t st) {
  scatter_rows_pair_kernel<<<static_cast<unsigned>(sm_count() * 4), 256, 0, st>>>(hi, lo, rb.tok_src, rb.total, padded);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

size_t ragged_index_words(int64_t B, int L) {
  // [seq_len B][seq_off B][total 1, padded to 4][tok_src B*L], each block 16-byte aligned
  const size_t b4 = (static_cast<size_t>(B) + 3) & ~size_t(3);
  return 2 * b4 + 4 + static_cast<size_t>(B) * L;
}

int ragged_build(const float* mask, int64_t B, int L, int32_t* idx, Ragged* out, cudaStream_t st) {
  MADE_REQUIRE(mask && idx && out, "ragged_build: null pointer");
  MADE_REQUIRE(B > 0 && L > 0 && B * L < (1LL << 31), "ragged_build: bad shape");
  const size_t b4 = (static_cast<size_t>(B) + 3) & ~size_t(3);
  int32_t* seq_len = idx;
  int32_t* seq_off = idx + b4;
  int32_t* total = idx + 2 * b4;
  int32_t* tok_src = idx + 2 * b4 + 4;
  const unsigned wb = static_cast<unsigned>(ceil_div64(B, 8));
  ragged_len_kernel<<<wb, 256, 0, st>>>(mask, B, L, seq_len);
  MADE_CHECK_LAUNCH();
  ragged_scan_kernel<<<1, 1024, 0, st>>>(seq_len, B, seq_off, total);
  MADE_CHECK_LAUNCH();
  ragged_fill_kernel<<<wb, 256, 0, st>>>(mask, B, L, seq_off, tok_src);
  MADE_CHECK_LAUNCH();
  out->seq_len = seq_len;
  out->seq_off = seq_off;
  out->total = total;
  out->tok_src = tok_src;
  out->B = B;
  out->L = L;
  return MADE_OK;
}

int ingest_gather(const void* in, int in_dtype, const Ragged& rb, int dim, op_t* out, bool split, cudaStream_t st) {
  MADE_REQUIRE(dim % 8 == 0, "ingest: dim must be a multiple of 8");
  const int64_t max_items = rb.B * rb.L * (dim / 8);
  int64_t blocks = ceil_div64(max_items, 256);
  int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  // Pinned host memory (UVA): the kernel pulls the valid rows over PCIe itself; two CTAs per SM keep
  // far more reads in flight than the link needs without occupying the SMs.
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, in) == cudaSuccess && attr.type == cudaMemoryTypeHost)
    cap = static_cast<int64_t>(sm_count()) * 2;
  else
    (void)cudaGetLastError();
  if (blocks > cap) blocks = cap;
  const unsigned g = static_cast<unsigned>(blocks);
  if (in_dtype == MADE_DTYPE_F32)
    ingest_gather_kernel<MADE_DTYPE_F32><<<g, 256, 0, st>>>(in, rb.tok_src, rb.total, dim, out, split ? 1 : 0);
  else if (in_dtype == MADE_DTYPE_BF16)
    ingest_gather_kernel<MADE_DTYPE_BF16><<<g, 256, 0, st>>>(in, rb.tok_src, rb.total, dim, out, split ? 1 : 0);
  else
    ingest_gather_kernel<MADE_DTYPE_F16><<<g, 256, 0, st>>>(in, rb.tok_src, rb.total, dim, out, split ? 1 : 0);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int pool_norm_ragged(const float* seq_packed, const Ragged& rb, float* pooled, cudaStream_t st) {
  pool_norm_ragged_kernel<<<static_cast<unsigned>(rb.B), 256, 0, st>>>(seq_packed, rb.seq_off, rb.seq_len, pooled);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int scatter_rows_f32(const float* packed, const Ragged& rb, float* padded, cudaStream_t st) {
  MADE_CUDA(cudaMemsetAsync(padded, 0, static_cast<size_t>(rb.B) * rb.L * 256 * 4, st));
  scatter_rows_f32_kernel<<<static_cast<unsigned>(sm_count() * 4), 256, 0, st>>>(packed, rb.tok_src, rb.total, padded);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int detr_mask(const float* frame_mask, const float* seg_mask, const int32_t* track_idx, int64_t seq_offset,
              int64_t B, float* mask_out, cudaStream_t st) {
  detr_mask_kernel<<<static_cast<unsigned>(ceil_div64(B * 146, 256)), 256, 0, st>>>(frame_mask, seg_mask, track_idx,
                                                                                   seq_offset, B, mask_out);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int detr_prep_ragged(const op_t* frame_out, const op_t* seg_out, const int32_t* track_idx, int64_t seq_offset,
                     const Ragged& rb, const float* inv_dim_t, op_t* src, op_t* pos, op_t* srcpos,
                     cudaStream_t st) {
  int64_t blocks = ceil_div64(rb.B * rb.L, 8);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  detr_prep_ragged_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(frame_out, seg_out, track_idx, seq_offset,
                                                                         rb.tok_src, rb.seq_off, rb.seq_len, rb.total,
                                                                         inv_dim_t, src, pos, srcpos);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

}  // namespace made
