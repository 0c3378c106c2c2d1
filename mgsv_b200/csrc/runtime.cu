// Host runtime glue: error strings, driver entry points (TMA descriptor encode), device queries.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <atomic>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>

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

int encode_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                   uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return MADE_ECUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p esz=%d inner=%llu outer=%llu stride=%llu box=%ux%u",
              static_cast<int>(r), base, elem_bytes, (unsigned long long)inner, (unsigned long long)outer,
              (unsigned long long)row_stride_bytes, box_inner, box_outer);
    return MADE_ECUDA;
  }
  return MADE_OK;
}

int encode_tmap_2d_16b(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  return encode_tmap_2d(map, base, 2, inner, outer, row_stride_bytes, box_inner, box_outer, 128);
}

int sm_count() {
  constexpr int kMaxDev = 64;
  static std::atomic<int> cache[kMaxDev];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDev) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

int ensure_dynamic_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;     // (kernel, device) -> bytes granted
  int dev = 0;
  MADE_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(mu);
  auto key = std::make_pair(kernel, dev);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return MADE_OK;
  MADE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done[key] = bytes;
  return MADE_OK;
}

// ---- kernel-family profiler (bench.py roofline) ------------------------------------------------------------
namespace {
struct ProfRec {
  cudaEvent_t a = nullptr, b = nullptr;
  int kind = 0;
};
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;       // records of the current window
std::vector<ProfRec> g_prof_pool;  // events are recycled across windows
}  // namespace

ProfScope::ProfScope(int kind, cudaStream_t st_) : slot(-1), st(st_) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    (void)cudaGetLastError();
    return;
  }
  std::lock_guard<std::mutex> g(g_prof_mu);
  ProfRec r;
  if (!g_prof_pool.empty()) {
    r = g_prof_pool.back();
    g_prof_pool.pop_back();
  } else if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
    (void)cudaGetLastError();
    return;
  }
  r.kind = kind;
  cudaEventRecord(r.a, st);
  slot = static_cast<int>(g_prof.size());
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (slot < static_cast<int>(g_prof.size())) cudaEventRecord(g_prof[slot].b, st);
}

// ---- small persistent host thread pool (fp32 -> fp16 conversion of host features) -----------------
class HostPool {
 public:
  explicit HostPool(int n) : n_(n < 1 ? 1 : n) {
    for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return n_; }
  // runs fn(worker_index) on every worker (the caller is worker 0) and waits for all of them
  void run(const std::function<void(int)>& fn) {
    {
      std::lock_guard<std::mutex> g(m_);
      fn_ = &fn;
      pending_ = n_ - 1;
      ++gen_;
    }
    cv_.notify_all();
    fn(0);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int idx) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* fn;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        fn = fn_;
      }
      (*fn)(idx);
      {
        std::lock_guard<std::mutex> g(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

static HostPool* host_pool(int n_threads) {
  static std::mutex m;
  static HostPool* pool = nullptr;
  std::lock_guard<std::mutex> g(m);
  if (!pool || pool->size() != n_threads) {
    delete pool;
    pool = new HostPool(n_threads);
  }
  return pool;
}

// fp32 -> IEEE fp16, round to nearest even, saturating at +-65504 (same as the device's cvt.rn.satfinite)
__attribute__((target("avx2,f16c"))) static void cvt_f32_f16(const float* src, uint16_t* dst, size_t n) {
  const __m256 lim = _mm256_set1_ps(65504.0f), nlim = _mm256_set1_ps(-65504.0f);
  size_t i = 0;
  for (; i + 16 <= n; i += 16) {
    __m256 a = _mm256_loadu_ps(src + i), b = _mm256_loadu_ps(src + i + 8);
    a = _mm256_max_ps(_mm256_min_ps(a, lim), nlim);
    b = _mm256_max_ps(_mm256_min_ps(b, lim), nlim);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm256_cvtps_ph(a, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 8), _mm256_cvtps_ph(b, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
  }
  for (; i < n; ++i) {
    float f = src[i];
    f = f > 65504.0f ? 65504.0f : (f < -65504.0f ? -65504.0f : f);
    dst[i] = _cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
  }
}

}  // namespace made

extern "C" {

const char* made_last_error_string(void) { return made::g_err; }

int made_abi_version(void) { return MADE_ABI_VERSION; }

int made_prof_enable(int on) {
  made::g_prof_on.store(on ? 1 : 0);
  return MADE_OK;
}

int made_prof_collect(double* ms_by_kind, int64_t* launches_by_kind, int n_kinds) {
  using namespace made;
  MADE_REQUIRE(ms_by_kind && launches_by_kind && n_kinds >= kProfKinds, "prof_collect: need %d slots", (int)kProfKinds);
  MADE_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> g(g_prof_mu);
  for (int i = 0; i < n_kinds; ++i) {
    ms_by_kind[i] = 0.0;
    launches_by_kind[i] = 0;
  }
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ms_by_kind[r.kind] += ms;
      launches_by_kind[r.kind] += 1;
    } else {
      (void)cudaGetLastError();
    }
    g_prof_pool.push_back(r);
  }
  g_prof.clear();
  return MADE_OK;
}

int made_h2d_valid_rows(const void* host_feats, int feats_dtype, const float* host_masks, int64_t B, int L,
                        int dim, void* host_stage16, int n_threads, void* dev_staging, int64_t* bytes_copied,
                        void* stream) {
  using namespace made;
  if (bytes_copied) *bytes_copied = 0;
  if (B == 0) return MADE_OK;
  MADE_REQUIRE(host_feats && host_masks && dev_staging, "h2d_valid_rows: null pointer");
  MADE_REQUIRE(feats_dtype >= MADE_DTYPE_F32 && feats_dtype <= MADE_DTYPE_F16, "h2d_valid_rows: bad dtype %d",
               feats_dtype);
  MADE_REQUIRE(L > 0 && dim > 0, "h2d_valid_rows: bad shape");
  const bool convert = host_stage16 != nullptr;
  MADE_REQUIRE(!convert || feats_dtype == MADE_DTYPE_F32, "h2d_valid_rows: host fp16 conversion needs fp32 features");
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
  if (dsts.empty()) return MADE_OK;
  if (convert) {
    // The host threads round the valid rows to fp16 into the pinned staging tensor (same [B, L, dim]
    // layout), which halves the bytes that cross PCIe; the copy list is rewritten to move the fp16 rows.
    const size_t n_el = total / 4;
    HostPool* pool = host_pool(n_threads < 1 ? 1 : n_threads);
    const int W = pool->size();
    const size_t per = ((n_el + W - 1) / W + 15) & ~size_t(15);
    char* stage = static_cast<char*>(host_stage16);
    const std::vector<void*>& srcs_c = srcs;
    const std::vector<size_t>& sizes_c = sizes;
    pool->run([&](int w) {
      size_t lo = static_cast<size_t>(w) * per, hi = lo + per;
      if (hi > n_el) hi = n_el;
      size_t pos = 0;   // element offset of copy i inside the concatenation of all copies
      for (size_t i = 0; i < srcs_c.size() && pos < hi; ++i) {
        const size_t cnt = sizes_c[i] / 4;
        if (pos + cnt > lo) {
          const size_t a = lo > pos ? lo - pos : 0, b = (hi - pos < cnt) ? hi - pos : cnt;
          const float* s0 = static_cast<const float*>(srcs_c[i]);
          const size_t off_el = static_cast<size_t>(static_cast<const char*>(srcs_c[i]) - hsrc) / 4;
          cvt_f32_f16(s0 + a, reinterpret_cast<uint16_t*>(stage) + off_el + a, b - a);
        }
        pos += cnt;
      }
    });
    for (size_t i = 0; i < srcs.size(); ++i) {
      const size_t off = static_cast<size_t>(static_cast<char*>(srcs[i]) - hsrc);
      srcs[i] = stage + off / 2;
      dsts[i] = ddst + off / 2;
      sizes[i] /= 2;
    }
    total /= 2;
  }
  if (bytes_copied) *bytes_copied = static_cast<int64_t>(total);
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
