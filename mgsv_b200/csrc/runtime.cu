// Host runtime glue: error strings, driver entry points (TMA descriptor encode), device queries.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace made {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", static_cast<int>(e), cudaGetErrorString(e), file,
            line, what);
  return MADE_ECUDA;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int encode_tmap_2d_16b(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return MADE_ECUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu stride=%llu box=%ux%u",
              static_cast<int>(r), base, (unsigned long long)inner, (unsigned long long)outer,
              (unsigned long long)row_stride_bytes, box_inner, box_outer);
    return MADE_ECUDA;
  }
  return MADE_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace made

extern "C" {

const char* made_last_error_string(void) { return made::g_err; }

int made_abi_version(void) { return MADE_ABI_VERSION; }

int made_h2d_valid_rows(const void* host_feats, int feats_dtype, const float* host_masks, int64_t B, int L,
                        int dim, void* dev_staging, int64_t* bytes_copied, void* stream) {
  using namespace made;
  if (bytes_copied) *bytes_copied = 0;
  if (B == 0) return MADE_OK;
  MADE_REQUIRE(host_feats && host_masks && dev_staging, "h2d_valid_rows: null pointer");
  MADE_REQUIRE(feats_dtype >= MADE_DTYPE_F32 && feats_dtype <= MADE_DTYPE_F16, "h2d_valid_rows: bad dtype %d",
               feats_dtype);
  MADE_REQUIRE(L > 0 && dim > 0, "h2d_valid_rows: bad shape");
  const size_t esz = feats_dtype == MADE_DTYPE_F32 ? 4 : 2;
  const size_t row = static_cast<size_t>(dim) * esz, seq = row * static_cast<size_t>(L);
  // one copy per sequence: rows [0, last valid row]; runs of fully valid sequences are merged
  static thread_local std::vector<void*> dsts, srcs;
  static thread_local std::vector<size_t> sizes;
  dsts.clear(); srcs.clear(); sizes.clear();
  const char* hsrc = static_cast<const char*>(host_feats);
  char* ddst = static_cast<char*>(dev_staging);
  size_t total = 0;
  bool open = false;   // the previous copy ends exactly at the start of this sequence
  for (int64_t b = 0; b < B; ++b) {
    const float* m = host_masks + b * L;
    int n = L;
    while (n > 0 && m[n - 1] == 0.f) --n;
    if (n == 0) { open = false; continue; }
    const size_t bytes = row * static_cast<size_t>(n);
    if (open) {
      sizes.back() += bytes;
    } else {
      dsts.push_back(ddst + b * seq);
      srcs.push_back(const_cast<char*>(hsrc + b * seq));
      sizes.push_back(bytes);
    }
    total += bytes;
    open = (n == L);
  }
  if (bytes_copied) *bytes_copied = static_cast<int64_t>(total);
  if (dsts.empty()) return MADE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaMemcpyAttributes attr;
  memset(&attr, 0, sizeof(attr));
  attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
  attr.flags = cudaMemcpyFlagPreferOverlapWithCompute;
  size_t attr_idx = 0, fail = 0;
  cudaError_t e = st == nullptr ? cudaErrorNotSupported   // the batch API rejects the legacy default stream
                                : cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &attr,
                                                       &attr_idx, 1, &fail, st);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    for (size_t i = 0; i < dsts.size(); ++i)
      MADE_CUDA(cudaMemcpyAsync(dsts[i], srcs[i], sizes[i], cudaMemcpyHostToDevice, st));
  }
  return MADE_OK;
}

int made_device_check(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return made::cuda_fail(e, "cudaGetDeviceProperties", __FILE__, __LINE__);
  if (prop.major != 10) {
    made::set_error("made_b200 needs an sm_100 (B200) device, found sm_%d%d (%s)", prop.major,
                    prop.minor, prop.name);
    return MADE_EUNSUPPORTED;
  }
  return MADE_OK;
}

}  // extern "C"
