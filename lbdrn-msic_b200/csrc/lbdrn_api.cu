// lbdrn_api.cu -- extern "C" entry points of liblbdrn_b200.so (see include/lbdrn.h for the contract).
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "lbdrn_infer_fp32.cuh"
#include "lbdrn_internal.h"
#include "lbdrn_train_fp32.cuh"

using namespace lbdrn;

namespace lbdrn {
std::atomic<long long> g_launches{0};
static thread_local std::string g_err;
const char* last_error() { return g_err.c_str(); }
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
}  // namespace lbdrn

namespace {

// LBDRN_PATH_AUTO: largest K for which the decode kernels evaluate the sine with MUFU.SIN alone (no explicit range
// reduction).  Measured on 8.4 M sub-pixels per K with reference-trained weights (tools/sine_parity.py,
// profiles/r2_sine_parity.txt): its mismatch rate against the reference's fp32 path equals that of the variant with the
// exact reduction (6.2 vs 6.4 ppm at K = 5, 31 vs 29 ppm at K = 8; bar 100 ppm) -- what separates both from the reference
// is MUFU.SIN's own 4e-7, not the one extra rounding of the argument.
constexpr int kAutoMufuMaxK = 8;

// ---- descriptor -> Net ----------------------------------------------------------------------------------
int resolve(const LbdrnDesc* d, Net& n, bool need_rows = true) {
  if (!d) return fail(LBDRN_E_INVALID, "null descriptor");
  memset(&n, 0, sizeof n);
  if (d->C < 1 || d->C > kMaxC) return fail(LBDRN_E_UNSUPPORTED, "C=%d bands (supported 1..%d)", d->C, kMaxC);
  if (d->H < 1 || d->W < 1) return fail(LBDRN_E_INVALID, "bad image size %dx%d", d->H, d->W);
  if (d->K < 1 || d->K > 15) return fail(LBDRN_E_INVALID, "K=%d outside 1..15 (4-bit header field)", d->K);
  if (d->D < 0 || d->D > 15) return fail(LBDRN_E_INVALID, "D=%d outside 0..15", d->D);
  if (d->nl < 1 || d->nl > kMaxLayers - 1) return fail(LBDRN_E_INVALID, "nl=%d outside 1..15", d->nl);
  if (d->bc != 32 && d->bc != 64 && d->bc != 128 && d->bc != 256)
    return fail(LBDRN_E_UNSUPPORTED, "bc=%d (kernels are built for 32/64/128/256)", d->bc);
  const bool coords = d->flags & LBDRN_USE_COORDINATES, emb = d->flags & LBDRN_EMBEDDING;
  const bool colors = d->flags & LBDRN_USE_COLORS;
  n.C = d->C; n.H = d->H; n.W = d->W; n.K = d->K; n.D = d->D; n.n = 2 * d->D + 1;
  n.bc = d->bc; n.nl = d->nl;
  n.tabw = 2 * d->n_freq * (emb ? 1 : 0) + 1;                    // LBDRNdataset.py:105
  n.nco = coords ? 2 * n.tabw : 0;
  n.ncol = colors ? d->C * n.n * n.n : 0;                        // LBDRNdataset.py:104
  n.dim_in = n.nco + n.ncol;
  if (n.dim_in <= 0) return fail(LBDRN_E_INVALID, "feature set is empty (USE_COORDINATES and USE_COLORS both off)");
  if (coords && (d->H < 2 || d->W < 2)) return fail(LBDRN_E_INVALID, "coordinates need H,W >= 2 (division by H-1)");
  if (colors && (d->D >= d->H || d->D >= d->W)) return fail(LBDRN_E_INVALID, "reflect padding needs D < H,W");
  n.relative = ((d->flags & LBDRN_RELATIVE) && d->D > 0) ? 1 : 0;  // LBDRNdataset.py:126
  n.relu = (d->flags & LBDRN_ACT_RELU) ? 1 : 0;
  n.w0 = d->w0;
  if (colors && d->msb_max == 0)
    return fail(LBDRN_E_INVALID, "MSB.max()==0: the reference's features are 0/0=NaN for this K (degenerate)");
  n.maxv = (float)d->msb_max;
  n.maxv_dev = d->msb_max_dev;
  n.qmax = (float)((1 << d->K) - 1);
  if (d->msb_dtype != LBDRN_U8 && d->msb_dtype != LBDRN_U16) return fail(LBDRN_E_INVALID, "bad msb_dtype");
  n.msb_u16 = d->msb_dtype == LBDRN_U16;
  n.lsb_u16 = d->K > 8;
  n.row0 = d->row0; n.row1 = d->row1; n.buf_row0 = d->buf_row0; n.buf_rows = d->buf_rows;
  if (need_rows) {
    if (!(0 <= n.row0 && n.row0 < n.row1 && n.row1 <= n.H)) return fail(LBDRN_E_INVALID, "bad row range [%d,%d)", n.row0, n.row1);
    const int lo = n.row0 - n.D > 0 ? n.row0 - n.D : 0, hi = n.row1 + n.D < n.H ? n.row1 + n.D : n.H;
    if (n.buf_row0 > lo || n.buf_row0 + n.buf_rows < hi || n.buf_row0 < 0 || n.buf_row0 + n.buf_rows > n.H)
      return fail(LBDRN_E_INVALID, "buffer rows [%d,%d) do not cover stripe+halo [%d,%d)", n.buf_row0,
                  n.buf_row0 + n.buf_rows, lo, hi);
  }
  int off = 0, din = n.dim_in;
  for (int l = 0; l <= n.nl; ++l) {                               // state_dict order, LBDRNmodel.py:62-77
    const int out = l < n.nl ? n.bc : n.C;
    n.woff[l] = off; off += out * din;
    n.boff[l] = off; off += out;
    din = n.bc;
  }
  n.P = off;
  return LBDRN_OK;
}

}  // namespace

// ---- per-device scratch (struct Scratch in lbdrn_internal.h) ----------------------------------------------
namespace lbdrn {
static std::mutex g_mu;
static Scratch g_scratch[64];

int get_scratch(int P, Scratch*& out) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(LBDRN_E_UNSUPPORTED, "device ordinal %d", dev);
  std::lock_guard<std::mutex> lk(g_mu);
  Scratch& s = g_scratch[dev];
  if (!s.sms) {
    CUDA_TRY(cudaDeviceGetAttribute(&s.sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&s.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CUDA_TRY(cudaMalloc(&s.partials, 4096 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&s.counter, sizeof(unsigned int)));
    CUDA_TRY(cudaMemset(s.counter, 0, sizeof(unsigned int)));
  }
  if ((size_t)P > s.wpack_n) {
    if (s.wpack) CUDA_TRY(cudaFree(s.wpack));
    s.wpack = nullptr; s.wpack_n = 0;
    CUDA_TRY(cudaMalloc(&s.wpack, ((size_t)P + 64) * sizeof(float)));
    s.wpack_n = (size_t)P;
  }
  out = &s;
  return LBDRN_OK;
}
}  // namespace lbdrn

namespace {

template <int MODE>
int run_infer(const LbdrnDesc* d, const void* msb, const void* lsb, const float* params, const float* tab, void* out,
              double* sse_out, void* stream, const int* skip_flag = nullptr) {
  Net n;
  int rc = resolve(d, n);
  if (rc) return rc;
  if (!msb || !params || (MODE != MODE_SSE && !out)) return fail(LBDRN_E_INVALID, "null device pointer");
  if (n.nco && !tab) return fail(LBDRN_E_INVALID, "USE_COORDINATES set but coord_tab_dev is NULL");
  if (MODE == MODE_SSE && (!lsb || !sse_out)) return fail(LBDRN_E_INVALID, "null lsb/sse pointer");
  Scratch* sc = nullptr;
  rc = get_scratch(n.P, sc);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  launch_pack_params(n, params, sc->wpack, st);
  CUDA_TRY(cudaGetLastError());
  InferArgs a;
  memset(&a, 0, sizeof a);
  a.net = n; a.msb = msb; a.lsb = lsb; a.wpack = sc->wpack; a.tab = tab; a.out = out;
  a.partials = sc->partials; a.counter = sc->counter; a.sse_out = sse_out;
  a.skip_flag = skip_flag;
  if (MODE == MODE_DECODE) return infer_fp32_decode(a, *sc, st);
  if (MODE == MODE_PREDICT) return infer_fp32_predict(a, *sc, st);
  return infer_fp32_sse(a, *sc, st);
}

// ---- small elementwise kernels ----------------------------------------------------------------------------
__global__ void split_kernel(const uint16_t* __restrict__ img, long long n, int K, int msb_u16, void* msb, void* lsb) {
  const unsigned mask = (1u << K) - 1u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned v = img[i], m = v >> K, l = v & mask;                // MSB = img>>K; LSB = img-(MSB<<K)  (LBDRNdataset.py:95-96)
    if (msb_u16) ((uint16_t*)msb)[i] = (uint16_t)m; else ((uint8_t*)msb)[i] = (uint8_t)m;
    if (K > 8) ((uint16_t*)lsb)[i] = (uint16_t)l; else ((uint8_t*)lsb)[i] = (uint8_t)l;
  }
}

__global__ void max_shifted_kernel(const uint16_t* __restrict__ img, long long n, int K, unsigned int* out) {
  unsigned m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = max(m, (unsigned)img[i] >> K);
  for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// a10: pseudo-random permutation of 0..n-1 without a sort: an unbalanced Feistel network over the index bits (each half is
// XORed with a keyed hash of the other: invertible whatever the hash), cycle-walked into [0, n).  2^bits < 2n, so an index
// needs fewer than two walks on average.
constexpr int kRandpermRounds = 6;
struct RandpermKeys {
  uint32_t key[kRandpermRounds];
  int bits_lo, bits_hi;
};
__device__ __forceinline__ uint32_t mix32(uint32_t h) {    // murmur3 finalizer
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
static __global__ void randperm_kernel(long long n, RandpermKeys k, int64_t* __restrict__ out) {
  const uint32_t mlo = (uint32_t)((1ull << k.bits_lo) - 1ull), mhi = (uint32_t)((1ull << k.bits_hi) - 1ull);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long x = (unsigned long long)i;
    do {
      uint32_t lo = (uint32_t)x & mlo, hi = (uint32_t)(x >> k.bits_lo) & mhi;
#pragma unroll
      for (int r = 0; r < kRandpermRounds; r += 2) {
        hi ^= mix32(lo ^ k.key[r]) & mhi;
        lo ^= mix32(hi ^ k.key[r + 1]) & mlo;
      }
      x = ((unsigned long long)hi << k.bits_lo) | lo;
    } while (x >= (unsigned long long)n);
    out[i] = (int64_t)x;
  }
}

__global__ void sse_u16_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b, long long n,
                               unsigned long long* out) {
  unsigned long long acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long d = (long long)a[i] - (long long)b[i];
    acc += (unsigned long long)(d * d);
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);      // integer sum: exact, order-independent
}

}  // namespace

// ---- training handle ------------------------------------------------------------------------------------
struct LbdrnTrain {
  Net net;
  LbdrnTrainCfg cfg;
  TrainPlan plan;
  int dev = 0, sms = 0;
  float *params = nullptr, *wpack = nullptr, *m = nullptr, *v = nullptr, *partial = nullptr;
  uint32_t *imsb = nullptr, *ilsb = nullptr;   // band-interleaved copies of the uint8 planes for the neighbourhood gather
  size_t i_npix = 0;
  bool i_failed = false;           // allocation failed once: gather from the CHW planes
  unsigned* gbar = nullptr;        // arrival counters of the training kernel's split grid barriers
  uint16_t* wimg = nullptr;        // fp16-split training kernel: image of its shared-memory weight operands (padding stays zero)
  float2* adam_tab = nullptr;      // per-step Adam scalars of the launch in flight (ring of two: launches may be queued)
  size_t adam_tab_n = 0;
  unsigned adam_tab_slot = 0;
};

namespace {

int launch_train(LbdrnTrain* t, TrainArgs& a, cudaStream_t st) {
  a.net = t->net; a.params = t->params; a.wpack = t->wpack; a.m = t->m; a.v = t->v;
  a.partial = t->partial; a.pstride = t->plan.pstride; a.dimpad = t->plan.dimpad; a.wimg = t->wimg; a.gbar = t->gbar;
  {
    // both planes of the scene against the device's L2: hints only pay when the window rows are DRAM misses
    int l2 = 0;
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, t->dev);
    const size_t plane_bytes = (size_t)t->net.C * t->net.buf_rows * t->net.W * ((t->net.msb_u16 ? 2 : 1) + (t->net.lsb_u16 ? 2 : 1));
    a.l2_hints = plane_bytes > (size_t)l2 || getenv("LBDRN_TRAIN_L2HINTS") != nullptr;   // the variable forces them (tests)
    a.fast_sine = getenv("LBDRN_TRAIN_FASTSIN") != nullptr;
  }
  a.beta1 = t->cfg.beta1; a.beta2 = t->cfg.beta2;
  a.omb1 = (float)(1.0 - t->cfg.beta1); a.omb2 = (float)(1.0 - t->cfg.beta2);
  a.beta2f = (float)t->cfg.beta2; a.eps = (float)t->cfg.eps;
  if (t->plan.tcx) {        // the streamed variant reads the hidden weights from the global image from its first step on
    launch_tcx_wimg(t->net, t->params, t->wimg, st);
    CUDA_TRY(cudaGetLastError());
  }
  return train_fp32_launch(t->plan, a, st);
}

}  // namespace

extern "C" {

int32_t lbdrn_version(void) { return LBDRN_ABI_VERSION; }
const char* lbdrn_last_error(void) { return lbdrn::last_error(); }
int64_t lbdrn_launch_count(void) { return g_launches.load(); }

int32_t lbdrn_dim_in(const LbdrnDesc* d) {
  Net n;
  int rc = resolve(d, n, false);
  return rc ? rc : n.dim_in;
}

int64_t lbdrn_param_count(const LbdrnDesc* d) {
  Net n;
  int rc = resolve(d, n, false);
  return rc ? rc : n.P;
}

int32_t lbdrn_has_tensor_path(const LbdrnDesc* d) {
  Net n;
  if (resolve(d, n, false)) return 0;
  return (tc_supported(n) || tcw_supported(n)) ? 1 : 0;
}

int32_t lbdrn_split(const uint16_t* img_dev, int64_t n, int32_t K, int32_t msb_dtype, void* msb_dev, void* lsb_dev,
                    void* stream) {
  if (!img_dev || !msb_dev || !lsb_dev || n <= 0 || K < 1 || K > 15) return fail(LBDRN_E_INVALID, "lbdrn_split: bad argument");
  long long blocks = (n + 1023) / 1024;
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img_dev, n, K, msb_dtype == LBDRN_U16, msb_dev, lsb_dev);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

int32_t lbdrn_max_shifted(const uint16_t* img_dev, int64_t n, int32_t K, uint32_t* max_dev, void* stream) {
  if (!img_dev || !max_dev || n <= 0 || K < 0 || K > 15) return fail(LBDRN_E_INVALID, "lbdrn_max_shifted: bad argument");
  long long blocks = (n + 2047) / 2048;
  if (blocks > 148 * 8) blocks = 148 * 8;
  max_shifted_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img_dev, n, K, max_dev);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

int32_t lbdrn_randperm(int64_t n, uint64_t seed, int64_t* out_dev, void* stream) {
  if (!out_dev || n <= 0 || n > ((int64_t)1 << 62)) return fail(LBDRN_E_INVALID, "lbdrn_randperm: bad argument");
  RandpermKeys k;
  int bits = 1;
  while (((uint64_t)1 << bits) < (uint64_t)n) ++bits;
  if (bits < 2) bits = 2;                                   // two non-empty halves
  k.bits_lo = bits / 2;
  k.bits_hi = bits - k.bits_lo;
  uint64_t x = seed;
  for (int r = 0; r < kRandpermRounds; ++r) {               // splitmix64
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    k.key[r] = (uint32_t)((z ^ (z >> 31)) >> 16);
  }
  long long blocks = (n + 1023) / 1024;
  if (blocks > 148 * 16) blocks = 148 * 16;
  randperm_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n, k, out_dev);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

int32_t lbdrn_sse_u16(const uint16_t* a_dev, const uint16_t* b_dev, int64_t n, uint64_t* sse_dev, void* stream) {
  if (!a_dev || !b_dev || !sse_dev || n <= 0) return fail(LBDRN_E_INVALID, "lbdrn_sse_u16: bad argument");
  long long blocks = (n + 4095) / 4096;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sse_u16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a_dev, b_dev, n, (unsigned long long*)sse_dev);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

int32_t lbdrn_decode(const LbdrnDesc* d, const void* msb_dev, const float* params_dev, const float* coord_tab_dev,
                     uint16_t* out_dev, void* stream) {
  Net n;
  int rc = resolve(d, n);
  if (rc) return rc;
  const bool tcw_ok = tcw_supported(n);
  const bool tc_ok = tc_supported(n) || tcw_ok;
  if (d->path == LBDRN_PATH_TENSOR && !tc_ok)
    return fail(LBDRN_E_UNSUPPORTED, "tensor-core decode is not built for this configuration");
  if ((d->path == LBDRN_PATH_TENSOR_FASTSIN || d->path == LBDRN_PATH_TENSOR_FASTSIN2) && !tc_ok)
    return fail(LBDRN_E_UNSUPPORTED, "tensor-core decode is not built for this configuration");
  if (d->path < LBDRN_PATH_AUTO || d->path > LBDRN_PATH_TENSOR_FASTSIN2) return fail(LBDRN_E_INVALID, "path=%d", d->path);
  if (tc_ok && d->path != LBDRN_PATH_PRECISE) {
    if (!msb_dev || !params_dev || !out_dev) return fail(LBDRN_E_INVALID, "null device pointer");
    // AUTO: MUFU sine while the K-bit quantiser leaves >= 10x margin on the 99.99 % identity bar (K <= 8: 99.9996 %
    // identical at K=8 on the parity suite), the 1.3e-7 polynomial above that
    // sine variant of the tensor kernels: 0 polynomial, 1 exact reduction + MUFU.SIN, 2 MUFU.SIN alone (lbdrn_tc.cu)
    int fast = d->path == LBDRN_PATH_TENSOR_FASTSIN2 ? 2 : (d->path == LBDRN_PATH_TENSOR_FASTSIN ? 1 : 0);
    if (d->path == LBDRN_PATH_AUTO) fast = n.K <= kAutoMufuMaxK ? 2 : (n.K <= 8 ? 1 : 0);
    if (tcw_ok) {
      // bc 128/256: two sibling launches of the wide kernel -- fp16-exact weights, or hi + lo weight operands -- and the
      // exactness flag decides ON THE DEVICE which of them decodes the scene (the other exits at once)
      const int* exact_flag = nullptr;
      return tcw_decode(n, msb_dev, params_dev, out_dev, fast != 0, &exact_flag, (cudaStream_t)stream);
    }
    return tc_decode(n, msb_dev, params_dev, coord_tab_dev, out_dev, fast, (cudaStream_t)stream);
  }
  return run_infer<MODE_DECODE>(d, msb_dev, nullptr, params_dev, coord_tab_dev, out_dev, nullptr, stream);
}

int32_t lbdrn_selftest_tc_gemm(const void* a_dev, const void* b_dev, float* d_dev, int32_t K, void* stream) {
  if (!a_dev || !b_dev || !d_dev) return fail(LBDRN_E_INVALID, "null device pointer");
  return tc_selftest(a_dev, b_dev, d_dev, K, (cudaStream_t)stream);
}

int32_t lbdrn_selftest_tc_gemm2(const void* a_dev, const void* b_dev, float* d_dev, int32_t N, int32_t K, int32_t a_mn,
                                int32_t b_mn, void* stream) {
  if (!a_dev || !b_dev || !d_dev) return fail(LBDRN_E_INVALID, "null device pointer");
  return tc_selftest2(a_dev, b_dev, d_dev, N, K, a_mn, b_mn, (cudaStream_t)stream);
}

int32_t lbdrn_selftest_tc_gemm3(const void* a_img_dev, int32_t a_bytes, const void* b_img_dev, int32_t b_bytes, float* d_dev,
                                float* raw_dev, int32_t N, int32_t ksteps, int32_t a_mn, int32_t a_rows, int32_t b_mn,
                                int32_t b_rows, void* stream) {
  if (!a_img_dev || !b_img_dev || !d_dev) return fail(LBDRN_E_INVALID, "null device pointer");
  return tc_selftest3(a_img_dev, a_bytes, b_img_dev, b_bytes, d_dev, raw_dev, N, ksteps, a_mn, a_rows, b_mn, b_rows,
                      (cudaStream_t)stream);
}

int32_t lbdrn_predict(const LbdrnDesc* d, const void* msb_dev, const float* params_dev, const float* coord_tab_dev,
                      float* y_dev, void* stream) {
  return run_infer<MODE_PREDICT>(d, msb_dev, nullptr, params_dev, coord_tab_dev, y_dev, nullptr, stream);
}

int32_t lbdrn_eval_sse(const LbdrnDesc* d, const void* msb_dev, const void* lsb_dev, const float* params_dev,
                       const float* coord_tab_dev, double* sse_dev, void* stream) {
  Net n;
  int rc = resolve(d, n);
  if (rc) return rc;
  if (tc_supported(n) && d->path != LBDRN_PATH_PRECISE) {
    if (!msb_dev || !lsb_dev || !params_dev || !sse_dev) return fail(LBDRN_E_INVALID, "null device pointer");
    return tc_eval_sse(n, msb_dev, lsb_dev, params_dev, coord_tab_dev, sse_dev, (cudaStream_t)stream);
  }
  // bc 128 / 256 with colour features: the wide tensor-core kernel with the low-order weight operands (LBDRN_EVAL_FP32=1
  // keeps the fp32 kernel for A/B)
  if (tcw_supported(n) && d->path != LBDRN_PATH_PRECISE && getenv("LBDRN_EVAL_FP32") == nullptr) {
    if (!msb_dev || !lsb_dev || !params_dev || !sse_dev) return fail(LBDRN_E_INVALID, "null device pointer");
    return tcw_eval_sse(n, msb_dev, lsb_dev, params_dev, sse_dev, (cudaStream_t)stream);
  }
  return run_infer<MODE_SSE>(d, msb_dev, lsb_dev, params_dev, coord_tab_dev, nullptr, sse_dev, stream);
}

int32_t lbdrn_train_create(const LbdrnDesc* d, const LbdrnTrainCfg* cfg, LbdrnTrain** out) {
  if (!cfg || !out) return fail(LBDRN_E_INVALID, "null argument");
  *out = nullptr;
  if (cfg->batch_size < 1) return fail(LBDRN_E_INVALID, "batch_size=%d", cfg->batch_size);
  LbdrnTrain* t = new LbdrnTrain();
  int rc = resolve(d, t->net);
  if (rc) { delete t; return rc; }
  t->cfg = *cfg;
  int supports_coop = 0, max_smem = 0;
  cudaError_t e = cudaGetDevice(&t->dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&t->sms, cudaDevAttrMultiProcessorCount, t->dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&supports_coop, cudaDevAttrCooperativeLaunch, t->dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, t->dev);
  if (e != cudaSuccess) { delete t; return fail(LBDRN_E_CUDA, "device query: %s", cudaGetErrorString(e)); }
  if (!supports_coop) { delete t; return fail(LBDRN_E_UNSUPPORTED, "device lacks cooperative launch"); }
  const Net& n = t->net;
  rc = train_fp32_plan(n, cfg->batch_size, t->sms, max_smem, t->plan);
  if (rc) { delete t; return rc; }
  const size_t pb = ((size_t)n.P + 64) * sizeof(float);
  e = cudaMalloc(&t->params, pb);
  if (e == cudaSuccess) e = cudaMalloc(&t->wpack, pb);
  if (e == cudaSuccess) e = cudaMalloc(&t->m, pb);
  if (e == cudaSuccess) e = cudaMalloc(&t->v, pb);
  if (e == cudaSuccess) e = cudaMalloc(&t->partial, (size_t)t->plan.grid * t->plan.pstride * sizeof(float));
  if (e == cudaSuccess && t->plan.wimg_bytes) e = cudaMalloc(&t->wimg, t->plan.wimg_bytes);
  if (e == cudaSuccess && t->plan.wimg_bytes) e = cudaMemset(t->wimg, 0, t->plan.wimg_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&t->gbar, 64);
  if (e == cudaSuccess) e = cudaMemset(t->params, 0, pb);
  if (e == cudaSuccess) e = cudaMemset(t->wpack, 0, pb);
  if (e == cudaSuccess) e = cudaMemset(t->m, 0, pb);
  if (e == cudaSuccess) e = cudaMemset(t->v, 0, pb);
  if (e == cudaSuccess) e = cudaMemset(t->partial, 0, (size_t)t->plan.grid * t->plan.pstride * sizeof(float));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    lbdrn_train_destroy(t);
    return fail(e == cudaErrorMemoryAllocation ? LBDRN_E_NOMEM : LBDRN_E_CUDA, "train_create: %s", cudaGetErrorString(e));
  }
  *out = t;
  return LBDRN_OK;
}

int32_t lbdrn_train_destroy(LbdrnTrain* t) {
  if (!t) return LBDRN_OK;
  cudaDeviceSynchronize();
  cudaFree(t->params); cudaFree(t->wpack); cudaFree(t->m); cudaFree(t->v); cudaFree(t->partial); cudaFree(t->adam_tab);
  cudaFree(t->wimg); cudaFree(t->gbar); cudaFree(t->imsb); cudaFree(t->ilsb);
  delete t;
  return LBDRN_OK;
}

int32_t lbdrn_train_set_params(LbdrnTrain* t, const float* params_dev, void* stream) {
  if (!t || !params_dev) return fail(LBDRN_E_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(t->params, params_dev, (size_t)t->net.P * sizeof(float), cudaMemcpyDeviceToDevice, st));
  launch_pack_params(t->net, t->params, t->wpack, st);
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

int32_t lbdrn_train_get_params(LbdrnTrain* t, float* params_dev, void* stream) {
  if (!t || !params_dev) return fail(LBDRN_E_INVALID, "null argument");
  CUDA_TRY(cudaMemcpyAsync(params_dev, t->params, (size_t)t->net.P * sizeof(float), cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return LBDRN_OK;
}

int32_t lbdrn_train_steps(LbdrnTrain* t, const void* msb_dev, const void* lsb_dev, const float* coord_tab_dev,
                          const int64_t* perm_dev, int64_t n_perm, int32_t n_steps, int64_t adam_t0, double lr,
                          float* losses_dev, void* stream) {
  if (!t || !msb_dev || !lsb_dev || !perm_dev || !losses_dev) return fail(LBDRN_E_INVALID, "null argument");
  if (t->net.nco && !coord_tab_dev) return fail(LBDRN_E_INVALID, "USE_COORDINATES set but coord_tab_dev is NULL");
  if (n_steps < 1 || (int64_t)(n_steps - 1) * t->cfg.batch_size >= n_perm)
    return fail(LBDRN_E_INVALID, "n_steps=%d does not match n_perm=%lld at batch_size=%d", n_steps, (long long)n_perm,
                t->cfg.batch_size);
  TrainArgs a;
  memset(&a, 0, sizeof a);
  a.msb = msb_dev; a.lsb = lsb_dev; a.tab = coord_tab_dev; a.perm = perm_dev; a.n_perm = n_perm;
  a.bs = t->cfg.batch_size; a.n_steps = n_steps; a.mode = TRAIN_FUSED; a.adam_t0 = adam_t0; a.lr = lr;
  a.losses = losses_dev;
  {
    // Adam's bias corrections per step, exactly as torch/optim/adam.py forms them (python floats = C doubles):
    // step_size = lr / (1 - beta1^t), bias_correction2_sqrt = sqrt(1 - beta2^t)
    if (t->adam_tab_n < (size_t)n_steps) {
      if (t->adam_tab) CUDA_TRY(cudaFree(t->adam_tab));
      t->adam_tab = nullptr; t->adam_tab_n = 0;
      CUDA_TRY(cudaMalloc(&t->adam_tab, 2 * (size_t)n_steps * sizeof(float2)));
      t->adam_tab_n = (size_t)n_steps;
    }
    std::vector<float2> tab((size_t)n_steps);
    for (int s = 0; s < n_steps; ++s) {
      const double tt = (double)(adam_t0 + s + 1);
      tab[s].x = (float)(lr / (1.0 - std::pow(t->cfg.beta1, tt)));
      tab[s].y = (float)std::sqrt(1.0 - std::pow(t->cfg.beta2, tt));
    }
    float2* dst = t->adam_tab + (size_t)(t->adam_tab_slot++ & 1u) * t->adam_tab_n;
    CUDA_TRY(cudaMemcpyAsync(dst, tab.data(), (size_t)n_steps * sizeof(float2), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    a.adam_tab = dst;
  }
  {
    // Band-interleaved copies of the planes (uint8, <= 4 bands, colour windows of 3, 5 or 7: the neighbourhood prefetch
    // of the 5x5 kernels and the chunk gather of the others read one word per pixel instead of one byte per band),
    // rebuilt before every launch -- the caller owns the planes and may have changed them; 2 x (C + 4) bytes per pixel of HBM traffic, ~0.3 ms at 8192^2 against ~200 ms per epoch.
    const Net& n = t->net;
    const bool eligible = !n.msb_u16 && !n.lsb_u16 && n.C <= 4 && (n.n == 3 || n.n == 5 || n.n == 7) && n.nco == 0 && n.ncol != 0 &&
                          getenv("LBDRN_TRAIN_CHW") == nullptr && !t->plan.pf_stride;
    const size_t npix = (size_t)n.buf_rows * n.W;
    if (eligible && !t->i_failed && t->i_npix < npix) {
      cudaFree(t->imsb); cudaFree(t->ilsb);
      t->imsb = t->ilsb = nullptr; t->i_npix = 0;
      cudaError_t e = cudaMalloc(&t->imsb, npix * 4 + 64);          // + slack: rows are read as aligned 32-byte pairs
      if (e == cudaSuccess) e = cudaMalloc(&t->ilsb, npix * 4 + 64);
      if (e == cudaSuccess) e = cudaMemsetAsync(t->imsb + npix, 0, 64, (cudaStream_t)stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(t->ilsb + npix, 0, 64, (cudaStream_t)stream);
      if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(t->imsb); cudaFree(t->ilsb);
        t->imsb = t->ilsb = nullptr; t->i_failed = true;            // not an error: the CHW gather needs no extra memory
      } else {
        t->i_npix = npix;
      }
    }
    if (eligible && t->imsb) {
      launch_interleave_u8(msb_dev, n.C, npix, t->imsb, t->sms, (cudaStream_t)stream);
      launch_interleave_u8(lsb_dev, n.C, npix, t->ilsb, t->sms, (cudaStream_t)stream);
      CUDA_TRY(cudaGetLastError());
      a.imsb = t->imsb; a.ilsb = t->ilsb;
    }
  }
  if (t->plan.pf_stride) {
    // TMA maps of the planes for the neighbourhood gather (nullptr when the buffers are not TMA-addressable)
    const Net& n = t->net;
    const CUtensorMap *tm = nullptr, *tl = nullptr;
    int rc = make_tensor_map_3d(msb_dev, 1, n.W, n.buf_rows, n.C, 32, 5, n.C, t->dev, (cudaStream_t)stream, &tm);
    if (!rc && tm) rc = make_tensor_map_3d(lsb_dev, 1, n.W, n.buf_rows, n.C, 16, 1, n.C, t->dev, (cudaStream_t)stream, &tl);
    if (rc) return rc;
    if (tm && tl) { a.tmap_msb = tm; a.tmap_lsb = tl; a.pf_off = t->plan.pf_off; a.pf_stride = t->plan.pf_stride; }
  }
  return launch_train(t, a, (cudaStream_t)stream);
}

int32_t lbdrn_train_grad(LbdrnTrain* t, const void* msb_dev, const void* lsb_dev, const float* coord_tab_dev,
                         const int64_t* batch_dev, int32_t n_local, int32_t n_global, float* grad_dev, void* stream) {
  if (!t || !msb_dev || !lsb_dev || !batch_dev || !grad_dev) return fail(LBDRN_E_INVALID, "null argument");
  if (t->net.nco && !coord_tab_dev) return fail(LBDRN_E_INVALID, "USE_COORDINATES set but coord_tab_dev is NULL");
  if (n_local < 1 || n_global < n_local || n_local > t->cfg.batch_size)
    return fail(LBDRN_E_INVALID, "bad local/global batch %d/%d", n_local, n_global);
  TrainArgs a;
  memset(&a, 0, sizeof a);
  a.msb = msb_dev; a.lsb = lsb_dev; a.tab = coord_tab_dev; a.perm = batch_dev; a.n_perm = n_local;
  a.bs = n_local; a.n_steps = 1; a.mode = TRAIN_GRAD_ONLY; a.n_global = n_global; a.grad_out = grad_dev;
  return launch_train(t, a, (cudaStream_t)stream);
}

int32_t lbdrn_train_apply(LbdrnTrain* t, const float* grad_dev, int64_t adam_t, double lr, void* stream) {
  if (!t || !grad_dev || adam_t < 1) return fail(LBDRN_E_INVALID, "bad argument");
  const double bc1 = 1.0 - std::pow(t->cfg.beta1, (double)adam_t), bc2 = 1.0 - std::pow(t->cfg.beta2, (double)adam_t);
  launch_adam_apply(t->net, grad_dev, t->params, t->wpack, t->m, t->v, (float)(1.0 - t->cfg.beta1),
                    (float)(1.0 - t->cfg.beta2), (float)t->cfg.beta2, (float)t->cfg.eps, (float)(lr / bc1),
                    (float)std::sqrt(bc2), (cudaStream_t)stream);
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

}  // extern "C"
