// lbdrn_tc.cu -- tcgen05 tensor-core decode kernel (sm_100a): the "TENSOR" path of lbdrn_decode.
//
// Per 128-pixel tile (one CTA = one warpgroup = the 128 TMEM lanes; 3 CTAs per SM overlap each other's phases):
//   patch   : (tile + D halo) of every band, MSB integers as fp16, staged in smem (reflect at true borders)
//   A1      : per pixel the (2D+1)^2*C INTEGER differences m_nbr - m_ctr (exact in fp16 for MSB <= 2048), written by
//             the pixel's thread straight into the UMMA canonical K-major / no-swizzle layout
//   MMA 1   : tcgen05.mma kind::f16, M=128 N=bc K=16 per instruction, D in TMEM (fp32), B1 = W1 * 2^s1 resident in
//             smem as fp16 (fpzip prec<=16 weights are bf16 values => exact after a power-of-two scale)
//   epi 1   : tcgen05.ld -> z = acc*(2^-s1/max) + b1 -> h = sin(w0 z) -> split h = hi + lo (two fp16 terms, |err| <=
//             2^-22 |h|) -> A2 = [hi | lo] written back in the same canonical layout
//   MMA l   : K = 2*bc against the SAME resident B_l for the hi and the lo half (no duplication of weights)
//   epi last: tcgen05.ld -> z -> sin -> output layer (bc x C) in fp32 FFMA -> sigmoid -> round_half_even -> (m<<K)+r
// Why split precision: rounding activations once to fp16/bf16/tf32 leaves only 99.95 % of pixels identical to the
// reference's fp32 path; hi+lo keeps 99.998 % (measured in SURVEY.md 7.2-1 and re-checked by tests/ on the GPU).
//
// Weights that are NOT exactly representable (e.g. -prec 32 streams) are detected by the prep kernel on the device;
// the tensor kernel then exits immediately and the fp32 kernel, launched behind it with the same flag, does the work.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "lbdrn_infer_fp32.cuh"
#include "lbdrn_internal.h"
#include "lbdrn_umma.cuh"

namespace lbdrn {

namespace {

// K layout of the first layer's operands (the order of the contraction index is ours to choose: A1 and B1 only have to
// agree).  0: the reference's column order c*n^2 + dy*n + dx.  1 (`tc_raw8`: uint8 planes, C = 4, D = 2): every window row
// (c, dy) padded from 5 to 6 entries, k' = (c*5 + dy)*6 + dx, so that a row is three fp16 pairs built from two aligned
// 32-bit words of the staged bytes; the sixth entry carries a zero weight.
bool tc_raw8(const Net& n) { return !n.msb_u16 && n.C == 4 && n.D == 2 && n.ncol != 0 && getenv("LBDRN_TC_NO_RAW8") == nullptr; }

void plan_block(const Net& n, TcHeader& h) {
  memset(&h, 0, sizeof h);
  h.k1 = n.ncol;                                       // the MMA contracts over the colour features only; coordinate /
  h.k1pad = align_up(n.ncol > 0 ? n.ncol : 1, 16);     // positional features enter through fp32 row / column tables
  h.klayout = tc_raw8(n) ? 1 : 0;
  if (h.klayout == 1) h.k1pad = align_up(n.C * 5 * 6, 16);
  h.nl = n.nl;
  int off = align_up((int)sizeof(TcHeader), 16);
  h.off_bias = off; off += n.nl * TC_BC * 4;
  h.off_w3 = off;   off += (TC_BC * 8 + 8) * 4;          // W3^T [bc][8] fp32 then b3[8]
  off = align_up(off, 128);
  for (int l = 0; l < n.nl; ++l) {
    h.off_b[l] = off;
    off += (l == 0 ? h.k1pad : TC_BC) * TC_BC * 2;
  }
  h.hi_bytes = align_up(off, 16);
  off = h.hi_bytes;
  for (int l = 0; l < n.nl; ++l) {
    h.off_blo[l] = off;
    off += (l == 0 ? h.k1pad : TC_BC) * TC_BC * 2;
  }
  h.total = align_up(off, 16);
}

// One block: per-layer max -> power-of-two scale -> fp16 operands in UMMA layout, exactness flag, biases, W3^T.
__global__ void tc_prep_kernel(Net net, TcHeader hdr, const float* __restrict__ params, uint8_t* __restrict__ blk) {
  __shared__ float s_max[32];
  __shared__ int s_exact, s_guard;
  const int tid = threadIdx.x;
  if (tid == 0) { s_exact = 1; s_guard = 0; }
  if (tid == 0) check_maxv_bound(net);
  for (int i = tid; i < hdr.total / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(blk)[i] = 0u;
  __syncthreads();
  TcHeader* H = reinterpret_cast<TcHeader*>(blk);
  for (int l = 0; l < net.nl; ++l) {
    const int K = l == 0 ? net.dim_in : net.bc;
    const int k0 = l == 0 ? net.nco : 0;                 // first colour column of W1 (coordinate columns come first)
    const float* W = params + net.woff[l];
    float m = 0.f;
    for (int i = tid; i < K * net.bc; i += blockDim.x)
      if (i % K >= k0) m = fmaxf(m, fabsf(W[i]));
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((tid & 31) == 0) s_max[tid >> 5] = m;
    __syncthreads();
    m = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, s_max[i]);
    __syncthreads();
    const int s = (m > 0.f && isfinite(m)) ? 13 - ilogbf(m) : 0;       // max|W| * 2^s in [2^13, 2^14)
    const float up = ldexpf(1.0f, s);
    __half* B = reinterpret_cast<__half*>(blk + hdr.off_b[l]);
    __half* Blo = reinterpret_cast<__half*>(blk + hdr.off_blo[l]);
    int bad = 0;
    for (int i = tid; i < K * net.bc; i += blockDim.x) {
      const int nrow = i / K;
      int k = i - nrow * K - k0;
      if (k < 0) continue;
      if (l == 0 && hdr.klayout == 1) k = (k / 5) * 6 + k % 5;            // window rows padded to 6 entries (plan_block)
      const float v = W[i] * up;                                          // exact (power of two)
      const __half hv = __float2half_rn(v);
      const float rem = v - __half2float(hv);                             // exact in fp32
      if (rem != 0.f) bad = 1;
      B[umma_off(TC_BC, nrow, k) / 2] = hv;
      Blo[umma_off(TC_BC, nrow, k) / 2] = __float2half_rn(rem);           // second term: |err| <= 2^-22 |v|
    }
    if (bad) atomicAnd(&s_exact, 0);
    // sine path: a = w0*(acc*scale + b) is evaluated as one FFMA, acc*(w0*scale) + w0*b (w0 folded here)
    const float fold = net.relu ? 1.0f : net.w0;
    if (tid == 0) H->scale[l] = fold * (l == 0 ? __fdiv_rn(ldexpf(1.0f, -s), net_maxv(net)) : ldexpf(1.0f, -s));
    float* bias = reinterpret_cast<float*>(blk + hdr.off_bias) + l * TC_BC;
    for (int i = tid; i < net.bc; i += blockDim.x) bias[i] = fold * params[net.boff[l] + i];
    // Can the sine's argument leave the range its fast reduction is exact for?  |w0 z_u| <= w0 (sum_k |W[u][k]| + |b_u|):
    // every input is at most 1 in magnitude (features are differences of values divided by the global maximum, hidden
    // inputs are sines, represented as hi + lo <= 1 + 2^-11; fp16 rounding of the weights adds 2^-11: the 1.001).  The
    // kernel skips the per-value guard of a layer whose bound is safe; coordinate features and non-finite weights keep it.
    float bound = 0.f;
    for (int u = tid; u < net.bc; u += blockDim.x) {
      float sabs = 0.f;
      for (int k = k0; k < K; ++k) sabs += fabsf(W[(size_t)u * K + k]);
      bound = fmaxf(bound, fabsf(fold) * (1.001f * sabs + fabsf(params[net.boff[l] + u])));
      if (!isfinite(sabs) || !isfinite(params[net.boff[l] + u])) bound = INFINITY;
    }
    if (!(bound <= 19000.0f) || (l == 0 && net.nco != 0)) atomicOr(&s_guard, 1 << l);
  }
  float* w3t = reinterpret_cast<float*>(blk + hdr.off_w3);
  for (int i = tid; i < net.bc * net.C; i += blockDim.x) {
    const int c = i / net.bc, u = i - c * net.bc;
    w3t[u * 8 + c] = params[net.woff[net.nl] + i];
  }
  for (int i = tid; i < net.C; i += blockDim.x) w3t[TC_BC * 8 + i] = params[net.boff[net.nl] + i];
  __syncthreads();
  if (tid == 0) {
    H->exact = s_exact;
    H->guard_mask = s_guard;
    H->k1 = hdr.k1; H->k1pad = hdr.k1pad; H->nl = hdr.nl; H->klayout = hdr.klayout;
    H->off_bias = hdr.off_bias; H->off_w3 = hdr.off_w3; H->total = hdr.total; H->hi_bytes = hdr.hi_bytes;
    for (int l = 0; l < net.nl; ++l) { H->off_b[l] = hdr.off_b[l]; H->off_blo[l] = hdr.off_blo[l]; }
  }
}

// USE_COORDINATES (LBDRNdataset.py:108-118): the coordinate / positional-encoding columns of a pixel are [row block of y |
// column block of x], so their contribution to the first layer separates: W1[:, :tabw] . rowtab[y] + W1[:, tabw:2 tabw] .
// coltab[x].  One thread per (row or column, unit) forms that dot product in fp32; the decode epilogue adds
// rtab[y][u] + ctab[x][u] to the scaled tensor-core accumulator.  rtab also carries the bias; both carry the w0 fold.
__global__ void tc_coord_tables_kernel(Net net, const float* __restrict__ params, const float* __restrict__ tab,
                                       float* __restrict__ rtab, float* __restrict__ ctab) {
  const int total = (net.H + net.W) * TC_BC;
  const float fold = net.relu ? 1.0f : net.w0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int pos = i / TC_BC, u = i - pos * TC_BC;
    const bool is_col = pos >= net.H;
    const float* t = tab + (size_t)pos * net.tabw;                                   // row table then column table
    const float* w = params + net.woff[0] + (size_t)u * net.dim_in + (is_col ? net.tabw : 0);
    float acc = is_col ? 0.f : params[net.boff[0] + u];
    for (int k = 0; k < net.tabw; ++k) acc = fmaf(w[k], t[k], acc);
    (is_col ? ctab + (size_t)(pos - net.H) * TC_BC : rtab + (size_t)pos * TC_BC)[u] = fold * acc;
  }
}

struct TcArgs {
  const float* rtab;      // USE_COORDINATES: [H][bc] row-table contribution to layer 0 (+ bias), w0-folded; else nullptr
  const float* ctab;      // [W][bc] column-table contribution
  const CUtensorMap* tmap_dev;  // 3-D tiled map of the MSB buffer (W, buf_rows, C) in global memory; valid when use_tma
  int no_trap;            // debug: do not trap on a barrier timeout
  int use_tma, box_w, box_lead;  // TMA patch staging for interior tiles: the box starts `box_lead` elements left of the
                                 // tile (a multiple of 16 B: the innermost TMA coordinate must be 16 B-aligned --
                                 // measured: an unaligned start raises "illegal instruction") and is box_w wide
  Net net;
  const void* msb;
  const uint8_t* blk;     // packed weight block (TcHeader + operands) in global memory
  uint16_t* out;
  int tiles_x, n_tiles;
  int a_bytes, w_bytes;   // smem: A region (per warpgroup), weight block copy
  int wg_bytes;           // smem: per-warpgroup staging region (fp16 patch + TMA landing box)
  int run_if_exact;       // 1: run only when the weights are fp16-exact; 0: only when they are not; -1: always
  const void* lsb;        // SSE mode: LSB codes
  double* partials;       // SSE mode: [gridDim.x]
  unsigned int* counter;  // SSE mode: zero on entry / exit
  double* sse_out;        // SSE mode
};

enum { TC_DECODE = 0, TC_SSE = 1 };

// sin(w0 z).  SINE_POLY: 2-term Cody-Waite(pi) + degree-9 polynomial (1.3e-7 abs).  SINE_CW_MUFU: exact 2-term reduction
// to [-pi, pi] then MUFU.SIN (abs err ~4e-7).  SINE_MUFU: sin.approx alone -- SASS is FMUL.RZ by 1/(2 pi) + MUFU.SIN, the
// unit takes the fractional turn itself, so the only error on top of MUFU's own is the one rounding (toward zero) of the
// argument expressed in turns, <= |a| * 1.2e-7 rad: 3 instructions per hidden unit instead of 7 (with the scale + bias FFMA).
enum { SINE_POLY = 0, SINE_CW_MUFU = 1, SINE_MUFU = 2 };
template <int SINE>
__device__ __forceinline__ float tc_sine(float a) {
  if (SINE == SINE_POLY) return sin_pi9_core(a);
  if (SINE == SINE_MUFU) return __sinf(a);
  const float t = fmaf(a, 0.15915494309189535f, 12582912.0f);
  const float k = t - 12582912.0f;
  float r = fmaf(k, -6.28318548202514648f, a);
  r = fmaf(k, 1.74845553146e-7f, r);       // 2*pi = 6.28318548202514648 - 1.74845553146e-7 (fp32 hi + lo)
  return __sinf(r);
}

// two sines at once: the range reduction of SINE_CW_MUFU runs on packed fp32 pairs (same roundings as tc_sine)
template <int SINE>
__device__ __forceinline__ void tc_sine2(float a0, float a1, float& s0, float& s1) {
  if (SINE != SINE_CW_MUFU) {
    s0 = tc_sine<SINE>(a0);
    s1 = tc_sine<SINE>(a1);
    return;
  }
  float t0, t1, k0, k1, r0, r1;
  ffma2(t0, t1, a0, a1, 0.15915494309189535f, 0.15915494309189535f, 12582912.0f, 12582912.0f);
  fadd2(k0, k1, t0, t1, -12582912.0f, -12582912.0f);
  ffma2(r0, r1, k0, k1, -6.28318548202514648f, -6.28318548202514648f, a0, a1);
  ffma2(r0, r1, k0, k1, 1.74845553146e-7f, 1.74845553146e-7f, r0, r1);
  s0 = __sinf(r0);
  s1 = __sinf(r1);
}

// nn.Sigmoid with the approximate exponential / reciprocal units (SINE_MUFU kernels only): 2^-22 relative on each, i.e.
// |dy| < 1e-7, an order of magnitude below what the MUFU sine already contributes
__device__ __forceinline__ float sigmoidf_fast(float z) { return __fdividef(1.0f, 1.0f + __expf(-z)); }

// A1 rows of one pixel straight from the staged bytes (RAW8; K layout 1 of plan_block).  rp: the aligned 32-bit word that
// holds the first byte of the pixel's top-left window element in band 0; bw4: words per staged row; sel_a: byte selector
// of window elements 0..3 inside the two words of a window row, sel_e: selector that expands element 4 (byte o of the
// second word) into an fp16 pair with itself (the sixth entry of a row has a zero weight); both fixed per thread.
// 0x64mm is the fp16 number 1024 + mm, so (pair) - (1024 + centre) is the exact integer difference.  Returns the four
// centre bytes packed (band c in byte c) for the final (m << K) + residual.
template <int CC>
__device__ __forceinline__ uint32_t build_a1_raw8(const uint32_t* __restrict__ rp, int bw4, uint32_t sel_a, uint32_t sel_e,
                                                  bool rel, uint8_t* sA, int tid) {
  constexpr int TRW_ = TC_TH + 4;                              // 12 staged rows per band (D = 2)
  constexpr uint32_t C64 = 0x64646464u;
  uint32_t cpk = 0;
  uint32_t dq[4];
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    uint32_t w0[5], w1[5];
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
      const uint32_t lo = rp[(c * TRW_ + dy) * bw4], hi = rp[(c * TRW_ + dy) * bw4 + 1];
      w0[dy] = __byte_perm(lo, hi, sel_a);
      w1[dy] = hi;                                           // element 4 is byte o of the second word: expanded from there
    }
    const uint32_t cpu = rel ? __byte_perm(w0[2], C64, 0x4242) : 0x64006400u;   // (1024 + centre) twice
    const __half2 cp = *reinterpret_cast<const __half2*>(&cpu);
    cpk = __byte_perm(cpk, w0[2], c == 0 ? 0x3216 : (c == 1 ? 0x3260 : (c == 2 ? 0x3610 : 0x6210)));
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint32_t pu = j < 2 ? __byte_perm(w0[dy], C64, j == 1 ? 0x4342 : 0x4140) : __byte_perm(w1[dy], C64, sel_e);
        const __half2 dv = __hsub2(*reinterpret_cast<const __half2*>(&pu), cp);
        const int q = (c * 5 + dy) * 3 + j;                  // pair index in K order; four pairs per 16-byte chunk
        dq[q & 3] = *reinterpret_cast<const uint32_t*>(&dv);
        if ((q & 3) == 3)
          *reinterpret_cast<uint4*>(sA + (size_t)((q >> 2) * 128 + tid) * 16) = make_uint4(dq[0], dq[1], dq[2], dq[3]);
      }
    }
  }
  // K is padded from C*30 to a multiple of 16: the pad multiplies zero weights but must be finite
#pragma unroll
  for (int kc = CC * 15 / 4; kc < (CC * 30 + 15) / 16 * 2; ++kc)
    *reinterpret_cast<uint4*>(sA + (size_t)(kc * 128 + tid) * 16) = make_uint4(0u, 0u, 0u, 0u);
  return cpk;
}

constexpr int TC_PF = 16;  // patch elements prefetched per thread (covers C*(8+2D)*(16+2D) <= 2048)

// CC/DD > 0: bands / radius known at compile time (feature offsets fold into immediates); CC == 0: generic tables.
// WLO: the weights carry a low-order fp16 term (fp32 weights during training / -prec 32 streams): extra MMAs against
// the lo operands.  MODE: TC_DECODE writes the reconstruction, TC_SSE accumulates sum((y - label)^2) (encode.py:105-108).
// NWG: warpgroups per CTA.  NWG=1: one tile in flight per CTA, 3 CTAs/SM, register-prefetched patches (any input).
// NWG=2 (TMA-addressable inputs, exact weights): two independent warpgroups share ONE copy of the weights, 2 CTAs/SM
// = 16 resident warps per SM instead of 12; each warpgroup has its own A region, staging buffers, mbarriers and TMEM
// columns and synchronises on its own named barrier.
// NWG=4 (TMA-addressable inputs, low-order weight operands resident: the per-epoch evaluation of fp32 weights): the weight
// block is 47 KB with the lo halves, so one CTA per SM with FOUR warpgroups sharing it keeps 16 warps resident where two
// single-warpgroup CTAs kept 8 (evaluation of an 8192^2 scene 14.9 -> see DESIGN 4.1).
// COORDS: USE_COORDINATES feature sets (table-driven colour offsets; fp32 row / column tables added in the first epilogue).
// RAW8 (uint8 planes, C = 4, D = 2; K layout 1 of plan_block): A1 is built straight from the staged BYTES -- per window row two
// aligned 32-bit shared-memory loads, byte permutes into fp16 pairs (0x6400 | m is the fp16 number 1024 + m, so the pair
// minus (1024 + centre) is the exact integer difference) -- no fp16 copy of the patch, 40 loads per pixel instead of 108.
// PF: the next 16 accumulator columns are requested from tensor memory before the current 16 are consumed.
template <int SINE, int CC, int DD, bool WLO, int MODE, int NWG, bool COORDS = false, bool RAW8 = false, bool PF = false>
__global__ void __launch_bounds__(TC_THREADS * NWG, NWG >= 4 ? 1 : (NWG == 2 ? 2 : (WLO ? 2 : 3))) tc_decode_kernel(const TcArgs a) {
  constexpr int THREADS = TC_THREADS * NWG;
  const Net& net = a.net;
  const int gtid = threadIdx.x, wg = gtid >> 7, tid = gtid & 127, warp = tid >> 5;
  const int C = CC ? CC : net.C, D = CC ? DD : net.D, n = 2 * D + 1;
  const int trows = TC_TH + 2 * D, twp = TC_TW + 2 * D;
  const int n_patch = C * trows * twp;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem + (size_t)wg * a.a_bytes;
  uint8_t* sW = smem + (size_t)NWG * a.a_bytes;
  uint8_t* region = sW + a.w_bytes + (size_t)wg * a.wg_bytes;
  __half* patch = reinterpret_cast<__half*>(region);
  uint8_t* raw = region + align_up(n_patch * 2, 128);                                    // TMA landing box (128 B aligned)
  uint16_t* koff = reinterpret_cast<uint16_t*>(sW + a.w_bytes + (size_t)NWG * a.wg_bytes);  // [k1pad] patch offset of feature k
  uint16_t* kctr = koff + TC_MAX_K1 + 16;                                                // [k1pad] patch offset of its centre
  __shared__ __align__(8) uint64_t s_mbar_tma[NWG];
  __shared__ __align__(8) uint64_t s_mbar[NWG];
  __shared__ uint32_t s_tmem;
  auto wg_sync = [&]() {
    if (NWG == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + wg), "r"(TC_THREADS) : "memory");
  };

  // ---- one-time setup: weight block -> smem, feature offset tables, TMEM, mbarrier ------------------------------------
  {
    const int4* src = reinterpret_cast<const int4*>(a.blk);
    int4* dst = reinterpret_cast<int4*>(sW);
    for (int i = gtid; i < a.w_bytes / 16; i += THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const TcHeader* H = reinterpret_cast<const TcHeader*>(sW);
  if (a.run_if_exact >= 0 && (H->exact != 0) != (a.run_if_exact != 0)) return;   // the sibling launch handles this scene
  __shared__ double s_red[THREADS / 32];
  double sse_local = 0.0;
  const int k1 = H->k1, k1pad = H->k1pad, NL = H->nl;
  if (CC == 0) {
    for (int k = gtid; k < k1pad; k += THREADS) {
      int off = 0, ctr = 0;
      if (k < k1) {
        const int c = k / (n * n), rem = k - c * n * n, dy = rem / n, dx = rem - dy * n;
        off = (c * trows + dy) * twp + dx;
        ctr = (c * trows + D) * twp + D;
      }
      koff[k] = (uint16_t)off;
      kctr[k] = (uint16_t)ctr;
    }
  }
  if (gtid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"(TC_TMEM_COLS * NWG)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (gtid == 0) {
    for (int g = 0; g < NWG; ++g) {
      mbar_init(smem_u32(&s_mbar[g]), 1);
      mbar_init(smem_u32(&s_mbar_tma[g]), 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem + (uint32_t)(wg * TC_TMEM_COLS);       // this warpgroup's accumulator columns
  const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);     // this warp's 32 lanes
  const uint32_t mbar = smem_u32(&s_mbar[wg]);
  const uint32_t idesc = umma_idesc_f16(128, TC_BC);
  const uint32_t sA_u = smem_u32(sA);
  const float* bias = reinterpret_cast<const float*>(sW + H->off_bias);
  const float* w3t = reinterpret_cast<const float*>(sW + H->off_w3);
  const bool rel = net.relative != 0;
  uint32_t phase = 0;
  const int tile0 = blockIdx.x * NWG + wg, tile_step = gridDim.x * NWG;

  // ---- patch element -> (band, row, col), fixed for the whole kernel; tile loads are prefetched one tile ahead -------
  constexpr int PFN = (CC == 4 && DD <= 2) ? 8 : TC_PF;   // patch elements per thread
  constexpr bool REGPF = NWG == 1 && !RAW8;               // per-thread register prefetch of non-TMA tiles
  const bool pf_ok = n_patch <= PFN * TC_THREADS;
  const int es = net.msb_u16 ? 2 : 1, box_w = a.box_w;
  int pe[REGPF ? PFN : 1];                         // band << 16 | row << 8 | col, or -1
  uint32_t pf[REGPF ? PFN : 1];
  long long pe_off[REGPF ? PFN : 1];               // element offset relative to the patch origin (interior tiles)
  int pe_src[RAW8 ? 1 : PFN];                      // offset of patch element i inside the TMA box (fixed per kernel), or -1
  if (!RAW8) {
#pragma unroll
    for (int i = 0; i < PFN; ++i) {
      const int e = tid + i * TC_THREADS;
      const int c = e / (trows * twp), rem = e - c * trows * twp, r = rem / twp, x = rem - r * twp;
      pe_src[RAW8 ? 0 : i] = e < n_patch ? (c * trows + r) * box_w + (a.box_lead - D) + x : -1;
      if (REGPF) {
        pe[REGPF ? i : 0] = e < n_patch ? ((c << 16) | (r << 8) | x) : -1;
        pe_off[REGPF ? i : 0] = e < n_patch ? ((long long)c * net.buf_rows + r) * net.W + x : 0;
      }
    }
  }
  // (y0, x0): top-left corner of the halo'd tile
  auto issue_patch_loads = [&](int y0, int x0) {
    if (!REGPF) return;
    if (y0 >= 0 && x0 >= 0 && y0 + trows <= net.H && x0 + twp <= net.W) {      // no reflection needed
      const long long origin = (long long)(y0 - net.buf_row0) * net.W + x0;
#pragma unroll
      for (int i = 0; i < (REGPF ? PFN : 1); ++i)
        if (pe[i] >= 0) pf[i] = load_msb_int(a.msb, net.msb_u16, (size_t)(origin + pe_off[i]));
      return;
    }
#pragma unroll
    for (int i = 0; i < (REGPF ? PFN : 1); ++i) {
      if (pe[i] >= 0) {
        const int gy = reflect_clamp(y0 + ((pe[i] >> 8) & 255), net.H), gx = reflect_clamp(x0 + (pe[i] & 255), net.W);
        pf[i] = load_msb_int(a.msb, net.msb_u16, ((size_t)(pe[i] >> 16) * net.buf_rows + (gy - net.buf_row0)) * net.W + gx);
      }
    }
  };
  // TMA staging of interior tiles: one elected thread issues the box load for the NEXT tile; it lands in `raw` while
  // this tile computes.  Border tiles (reflection needed) and buffers TMA cannot address (row pitch not a multiple of
  // 16 B) use the per-thread prefetch above (NWG=1, fp16 patch) or plain loads at the top of the tile.
  const uint32_t mbar_tma = smem_u32(&s_mbar_tma[wg]), raw_u = smem_u32(raw);
  const uint32_t box_bytes = (uint32_t)(box_w * trows * C * es);
  // tile index -> image coordinates of its first pixel: one division per tile, for the NEXT tile only
  auto tile_xy = [&](int tile, int& y0, int& x0) {
    const int ty = tile / a.tiles_x;
    y0 = net.row0 + ty * TC_TH;
    x0 = (tile - ty * a.tiles_x) * TC_TW;
  };
  auto interior = [&](int y0, int x0) {             // the halo'd tile lies inside the image and TMA can address the buffer
    return a.use_tma && y0 - D >= 0 && x0 - D >= 0 && y0 - D + trows <= net.H && x0 - D + twp <= net.W;
  };
  auto stage_tile = [&](bool valid, int y0, int x0, bool by_tma) {   // called by all threads of the warpgroup
    if (!valid) return;
    if (by_tma) {
      if (warp == 0) {             // warp-uniform; one elected lane issues (operands stay in uniform registers)
        if (elect_one()) {
          mbar_expect_tx(mbar_tma, box_bytes);
          tma_load_3d(raw_u, a.tmap_dev, x0 - a.box_lead, y0 - D - net.buf_row0, 0, mbar_tma);
        }
        __syncwarp();
      }
    } else if (REGPF && pf_ok) {
      issue_patch_loads(y0 - D, x0 - D);
    }
  };
  uint32_t phase_tma = 0;
  int cy0 = 0, cx0 = 0;
  bool cur_tma = false;                             // staging mechanism of the tile about to be consumed
  if (tile0 < a.n_tiles) {
    tile_xy(tile0, cy0, cx0);
    cur_tma = interior(cy0, cx0);
  }
  stage_tile(tile0 < a.n_tiles, cy0, cx0, cur_tma);
  const int pr = tid >> 4, px = tid & 15;

  for (int t = tile0; t < a.n_tiles; t += tile_step) {
    const int ty0 = cy0, tx0 = cx0;
    const bool has_next = t + tile_step < a.n_tiles;
    int ny0 = 0, nx0 = 0;
    bool nxt_tma = false;
    if (has_next) {
      tile_xy(t + tile_step, ny0, nx0);
      nxt_tma = interior(ny0, nx0);
    }
    uint32_t mctr[RAW8 ? 1 : kMaxC];                // centre MSB integers for the final (m << K) + residual (RAW8: packed bytes)

    if constexpr (RAW8) {
      // ---- staged bytes of (tile + halo), TMA box layout [band][row][box_w]; A1 straight from them ------------------------
      if (cur_tma) {
        mbar_wait(mbar_tma, phase_tma, 2, t, a.no_trap);
        phase_tma ^= 1;
      } else {
        for (int e = tid; e < n_patch; e += TC_THREADS) {
          const int c = e / (trows * twp), rem = e - c * trows * twp, r = rem / twp, x = rem - r * twp;
          const int gy = reflect_clamp(ty0 - D + r, net.H), gx = reflect_clamp(tx0 - D + x, net.W);
          raw[(c * trows + r) * box_w + (a.box_lead - D) + x] =
              (uint8_t)load_msb_int(a.msb, 0, ((size_t)c * net.buf_rows + (gy - net.buf_row0)) * net.W + gx);
        }
        wg_sync();
      }
      const int bw4 = box_w >> 2;                                // box_w is a multiple of 16 bytes
      const int bcol = a.box_lead - DD + px;                     // byte column of the window's first element
      const uint32_t o = (uint32_t)bcol & 3u;
      const uint32_t cpk = build_a1_raw8<CC ? CC : 1>(reinterpret_cast<const uint32_t*>(raw) + pr * bw4 + (bcol >> 2), bw4,
                                                      0x3210u + o * 0x1111u, 0x4040u + o * 0x0101u, rel, sA, tid);
      mctr[0] = cpk;
    } else {
    // ---- patch: (tile + halo) MSB integers as fp16 ---------------------------------------------------------------------
    if (cur_tma) {
      mbar_wait(mbar_tma, phase_tma, 2, t, a.no_trap);
      phase_tma ^= 1;
      if (pf_ok) {                                                   // box offsets precomputed in pe_src[]
#pragma unroll
        for (int i = 0; i < PFN; ++i) {
          if (pe_src[RAW8 ? 0 : i] >= 0) {
            const uint32_t v = net.msb_u16 ? (uint32_t)reinterpret_cast<const uint16_t*>(raw)[pe_src[RAW8 ? 0 : i]] : (uint32_t)raw[pe_src[RAW8 ? 0 : i]];
            patch[tid + i * TC_THREADS] = __uint2half_rn(v);
          }
        }
      } else {
        for (int e = tid; e < n_patch; e += TC_THREADS) {
          const int row = e / twp, x = e - row * twp;               // row = band * trows + patch row
          const int src = row * box_w + (a.box_lead - D) + x;
          const uint32_t v = net.msb_u16 ? (uint32_t)reinterpret_cast<const uint16_t*>(raw)[src] : (uint32_t)raw[src];
          patch[e] = __uint2half_rn(v);
        }
      }
      fence_async_smem();      // our generic-proxy reads of `raw` are ordered before the next TMA write into it
    } else if (REGPF && pf_ok) {
#pragma unroll
      for (int i = 0; i < (REGPF ? PFN : 1); ++i)
        if (pe[i] >= 0) patch[tid + i * TC_THREADS] = __uint2half_rn(pf[i]);
    } else {
      for (int e = tid; e < n_patch; e += TC_THREADS) {
        const int c = e / (trows * twp), rem = e - c * trows * twp, r = rem / twp, x = rem - r * twp;
        const int gy = reflect_clamp(ty0 - D + r, net.H), gx = reflect_clamp(tx0 - D + x, net.W);
        patch[e] = __uint2half_rn(load_msb_int(a.msb, net.msb_u16, ((size_t)c * net.buf_rows + (gy - net.buf_row0)) * net.W + gx));
      }
    }
    wg_sync();
    stage_tile(has_next, ny0, nx0, nxt_tma);                                         // lands while this tile computes

    // ---- A1 row of this thread's pixel: integer differences (exact in fp16), 16 B per K chunk ---------------------------
    const __half* pme = patch + pr * twp + px;
    if (CC) {
      constexpr int N_ = 2 * DD + 1, NN_ = N_ * N_, K1_ = (CC ? CC : 1) * NN_, TWP_ = TC_TW + 2 * DD, TRW_ = TC_TH + 2 * DD;
      __half2 ctr2[CC ? CC : 1];
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        const __half cv = rel ? pme[(c * TRW_ + DD) * TWP_ + DD] : __half(0);
        ctr2[c] = __halves2half2(cv, cv);
      }
#pragma unroll
      for (int kc = 0; kc < (K1_ + 15) / 16 * 2; ++kc) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k0 = kc * 8 + 2 * e, k1_ = k0 + 1;
          const int c0 = k0 / NN_, c1 = k1_ / NN_;
          const int o0 = (c0 * TRW_ + (k0 % NN_) / N_) * TWP_ + (k0 % NN_) % N_;
          const int o1 = (c1 * TRW_ + (k1_ % NN_) / N_) * TWP_ + (k1_ % NN_) % N_;
          __half2 v = __halves2half2(k0 < K1_ ? pme[o0] : __half(0), k1_ < K1_ ? pme[o1] : __half(0));
          if (k0 < K1_) {
            // both lanes of the pair belong to the same band except across a band boundary
            const __half2 cpair = (c0 == c1 || k1_ >= K1_) ? ctr2[c0 < CC ? c0 : 0]
                                                           : __halves2half2(__low2half(ctr2[c0 < CC ? c0 : 0]),
                                                                            __low2half(ctr2[c1 < CC ? c1 : 0]));
            v = __hsub2(v, (k1_ < K1_) ? cpair : __halves2half2(__low2half(cpair), __half(0)));
          }
          w[e] = *reinterpret_cast<const uint32_t*>(&v);
        }
        *reinterpret_cast<uint4*>(sA + (size_t)(kc * 128 + tid) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      for (int kc = 0; kc < k1pad / 8; ++kc) {
        __half2 v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = kc * 8 + 2 * e;
          __half x0 = pme[koff[k]], x1 = pme[koff[k + 1]];
          if (rel) {
            x0 = __hsub(x0, pme[kctr[k]]);
            x1 = __hsub(x1, pme[kctr[k + 1]]);
          }
          v[e] = __halves2half2(k < k1 ? x0 : __half(0), k + 1 < k1 ? x1 : __half(0));
        }
        *reinterpret_cast<uint4*>(sA + (size_t)(kc * 128 + tid) * 16) =
            make_uint4(*reinterpret_cast<uint32_t*>(&v[0]), *reinterpret_cast<uint32_t*>(&v[1]),
                       *reinterpret_cast<uint32_t*>(&v[2]), *reinterpret_cast<uint32_t*>(&v[3]));
      }
    }
    // centre MSB integers for the final (m << K) + residual
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      mctr[RAW8 ? 0 : c] = c < C ? (uint32_t)__half2int_rn(patch[(c * trows + pr + D) * twp + px + D]) : 0u;
    }

    float yacc[kMaxC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) yacc[c] = 0.f;

    for (int l = 0; l < NL; ++l) {
      // ---- MMA for hidden layer l: one elected thread issues, completion arrives on the mbarrier -----------------------
      fence_async_smem();          // generic-proxy smem writes (and, RAW8, our reads of `raw`) -> ordered before the async proxy
      tc_fence_before();
      wg_sync();
      if (warp == 0) {
       // warp 0 of the warpgroup, warp-uniform control flow, ONE elected lane issues: issuing from inside `if (tid == 0)`
       // makes the compiler wrap every tcgen05 instruction in an R2UR.BROADCAST / ELECT loop
       tc_fence_after();
       if (elect_one()) {
        const uint32_t sB_u = smem_u32(sW + H->off_b[l]);
        const uint32_t sBlo_u = smem_u32(sW + H->off_blo[l]);
        if (l == 0) {
          for (int i = 0; i < k1pad / 16; ++i)
            umma_f16(tmem, umma_desc(sA_u + i * 2 * 2048, 2048, 128), umma_desc(sB_u + i * 2 * 1024, 1024, 128), idesc, i > 0);
          if (WLO)
            for (int i = 0; i < k1pad / 16; ++i)
              umma_f16(tmem, umma_desc(sA_u + i * 2 * 2048, 2048, 128), umma_desc(sBlo_u + i * 2 * 1024, 1024, 128), idesc, 1);
        } else {
          for (int i = 0; i < 2 * TC_BC / 16; ++i)       // hi half then lo half of A2, both against the same B_l
            umma_f16(tmem, umma_desc(sA_u + i * 2 * 2048, 2048, 128),
                     umma_desc(sB_u + (i % (TC_BC / 16)) * 2 * 1024, 1024, 128), idesc, i > 0);
          if (WLO)                                       // hi half of A2 against the low-order weight term
            for (int i = 0; i < TC_BC / 16; ++i)
              umma_f16(tmem, umma_desc(sA_u + i * 2 * 2048, 2048, 128), umma_desc(sBlo_u + i * 2 * 1024, 1024, 128), idesc, 1);
        }
        umma_commit(mbar);
       }
       __syncwarp();
      }
      if (RAW8 && l == 0) stage_tile(has_next, ny0, nx0, nxt_tma);    // every thread's reads of `raw` precede the barrier above
      mbar_wait(mbar, phase, 1, t, a.no_trap);
      phase ^= 1;
      tc_fence_after();

      // ---- epilogue of layer l: thread = pixel = TMEM lane ------------------------------------------------------------
      const float scale = H->scale[l];
      const bool guard = ((H->guard_mask >> l) & 1) != 0;
      const float* bl = bias + l * TC_BC;
      const bool last = l + 1 == NL;
      const bool coords = COORDS && l == 0;
      const float4* rrow = reinterpret_cast<const float4*>(a.rtab + (size_t)min(ty0 + pr, net.H - 1) * TC_BC);
      const float4* crow = reinterpret_cast<const float4*>(a.ctab + (size_t)min(tx0 + px, net.W - 1) * TC_BC);
      // one block of 16 hidden units: scale + bias, activation, then either the hi/lo operand of the next layer or the
      // output layer's partial sums
      auto block16 = [&](int cb, float (&acc)[16]) {
        float bterm[16];
        if (coords) {                 // fp32 coordinate contribution (includes the bias) instead of the bias alone
#pragma unroll
          for (int q = 0; q < (COORDS ? 4 : 0); ++q) {
            const float4 r = __ldg(rrow + (cb >> 2) + q), c = __ldg(crow + (cb >> 2) + q);
            bterm[4 * q] = r.x + c.x; bterm[4 * q + 1] = r.y + c.y; bterm[4 * q + 2] = r.z + c.z; bterm[4 * q + 3] = r.w + c.w;
          }
        } else {                      // 4 x LDS.128 (broadcast) instead of 16 scalar loads
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b4 = *reinterpret_cast<const float4*>(bl + cb + 4 * q);
            bterm[4 * q] = b4.x; bterm[4 * q + 1] = b4.y; bterm[4 * q + 2] = b4.z; bterm[4 * q + 3] = b4.w;
          }
        }
        float h[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2)                        // = w0 * z (w0 folded into scale and bias), two per FFMA2
          ffma2(acc[j], acc[j + 1], acc[j], acc[j + 1], scale, scale, bterm[j], bterm[j + 1]);
        if (net.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) h[j] = fmaxf(acc[j], 0.f);
        } else {
#pragma unroll
          for (int j = 0; j < 16; j += 2) tc_sine2<SINE>(acc[j], acc[j + 1], h[j], h[j + 1]);
          if (guard) {                                          // uniform: tc_prep_kernel could not bound |w0 z| for this layer
            float amax = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(acc[j]));
            if (__builtin_expect(!(amax <= 20000.0f), 0)) {     // huge argument: library slow path, out of line
#pragma unroll
              for (int j = 0; j < 16; ++j) h[j] = sin_slow(acc[j]);
            }
          }
        }
        if (!last) {
          // h = hi + lo, both fp16; A2 chunk index: hi -> (cb+j)/8, lo -> bc/8 + (cb+j)/8
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const __half2 hh = __floats2half2_rn(h[j], h[j + 1]);
            const float2 back = __half22float2(hh);
            float l0, l1;
            ffma2(l0, l1, back.x, back.y, -1.0f, -1.0f, h[j], h[j + 1]);      // h - hi, exact
            const __half2 ll = __floats2half2_rn(l0, l1);
            hi[j >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
            lo[j >> 1] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          const int kc = cb >> 3;
          *reinterpret_cast<uint4*>(sA + (size_t)(kc * 128 + tid) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(sA + (size_t)((kc + 1) * 128 + tid) * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          *reinterpret_cast<uint4*>(sA + (size_t)((TC_BC / 8 + kc) * 128 + tid) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<uint4*>(sA + (size_t)((TC_BC / 8 + kc + 1) * 128 + tid) * 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        } else {
          // output layer in fp32: y_c += W3[c][u] * h[u]   (W3^T rows of 8 floats: one or two LDS.128, broadcast)
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 wa = *reinterpret_cast<const float4*>(w3t + (cb + j) * 8);
            ffma2(yacc[0], yacc[1], wa.x, wa.y, h[j], h[j], yacc[0], yacc[1]);
            ffma2(yacc[2], yacc[3], wa.z, wa.w, h[j], h[j], yacc[2], yacc[3]);
            if (C > 4) {
              const float4 wb = *reinterpret_cast<const float4*>(w3t + (cb + j) * 8 + 4);
              ffma2(yacc[4], yacc[5], wb.x, wb.y, h[j], h[j], yacc[4], yacc[5]);
              ffma2(yacc[6], yacc[7], wb.z, wb.w, h[j], h[j], yacc[6], yacc[7]);
            }
          }
        }
      };
      if constexpr (PF) {
        // ping-pong: the load of block i+1 is in flight while block i is consumed
        uint32_t ra[16], rb[16];
        tmem_ld16_issue(tmem_row, ra);
#pragma unroll 1
        for (int cb = 0; cb < TC_BC; cb += 32) {
          float acc[16];
          tmem_ld16_wait(ra);
          tmem_ld16_issue(tmem_row + cb + 16, rb);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(ra[j]);
          block16(cb, acc);
          tmem_ld16_wait(rb);
          if (cb + 32 < TC_BC) tmem_ld16_issue(tmem_row + cb + 32, ra);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(rb[j]);
          block16(cb + 16, acc);
        }
      } else {
#pragma unroll 1
        for (int cb = 0; cb < TC_BC; cb += 16) {
          float acc[16];
          tmem_ld16(tmem_row + cb, acc);
          block16(cb, acc);
        }
      }
    }

    // ---- sigmoid, inverse quantisation, integer write (decode.py:131-134) ----------------------------------------------
    const int gy = ty0 + pr, gx = tx0 + px;
    if (gy < net.row1 && gx < net.W) {
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) {
        if (c < C) {
          const float yz = yacc[c] + w3t[TC_BC * 8 + c];
          const float y = SINE == SINE_MUFU ? sigmoidf_fast(yz) : sigmoidf_rn(yz);
          const size_t off = ((size_t)c * net.buf_rows + (gy - net.buf_row0)) * net.W + gx;
          if (MODE == TC_DECODE) {
            const int res = (int)rintf(y * net.qmax);
            const uint32_t m = RAW8 ? ((mctr[0] >> (8 * (c & 3))) & 0xFFu) : mctr[RAW8 ? 0 : c];
            a.out[off] = (uint16_t)((m << net.K) + (uint32_t)res);
          } else {
            const uint32_t code = net.lsb_u16 ? (uint32_t)((const uint16_t*)a.lsb)[off] : (uint32_t)((const uint8_t*)a.lsb)[off];
            const float d = y - __fdiv_rn((float)code, net.qmax);
            sse_local += (double)(d * d);
          }
        }
      }
    }
    cy0 = ny0; cx0 = nx0; cur_tma = nxt_tma;
    tc_fence_before();      // our tcgen05.ld's are ordered before the next tile's MMA (issued after the next barrier)
  }

  if (MODE == TC_SSE) {
    // deterministic: lanes -> warp -> CTA partial -> the last CTA sums the partials in index order
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sse_local += __shfl_xor_sync(0xffffffffu, sse_local, off);
    if ((gtid & 31) == 0) s_red[gtid >> 5] = sse_local;
  }
  __syncthreads();
  if (MODE == TC_SSE && gtid == 0) {
    double sum = 0.0;
    for (int i = 0; i < THREADS / 32; ++i) sum += s_red[i];
    a.partials[blockIdx.x] = sum;
    __threadfence();
    if (atomicAdd(a.counter, 1u) == gridDim.x - 1) {
      __threadfence();
      double tot = 0.0;
      for (unsigned int i = 0; i < gridDim.x; ++i) tot += ((volatile double*)a.partials)[i];
      *a.sse_out = tot;
      *a.counter = 0u;
    }
  }
  if (gtid < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(TC_TMEM_COLS * NWG) : "memory");
}

// ---- pipelined decode kernel of the paper's configuration ----------------------------------------------------------------
// uint8 planes, C = 4, D = 2, colour features, nl = 2, fp16-exact weights, TMA-addressable planes (the headline workload).
// Same arithmetic as tc_decode_kernel<SINE, 4, 2, false, TC_DECODE, 2, false, true>; what changes is the ORDER of a
// warpgroup's work, so that it never sleeps on a tensor-core completion.  Each warpgroup owns TWO accumulators (128 TMEM
// columns; 2 warpgroups x 2 CTAs fill the SM's 512) and keeps one A region.  Per tile i (p = i & 1):
//   a  wait MMA1(i)               (issued during tile i-1's step e, finished long ago)
//   b  epilogue 1: acc[p] -> sine -> hi/lo -> A2 in the A region; issue MMA2(i) -> acc[p]
//   c  while MMA2(i) runs: sigmoid / quantise / store of tile i-1 (its four output sums were kept in registers) and the
//      coordinates of tile i+2
//   d  wait MMA2(i)               (the A region is free again)
//   e  bytes of tile i+1 (TMA issued during tile i) -> A1 in the A region; issue MMA1(i+1) -> acc[p^1]; TMA for tile i+2
//   f  while MMA1(i+1) runs: epilogue 2 of tile i: acc[p] -> sine -> output layer sums
template <int SINE>
__global__ void __launch_bounds__(TC_THREADS * 2, 2) tc_pipe_kernel(const TcArgs a) {
  constexpr int NWG = 2, THREADS = TC_THREADS * NWG, CC = 4, DD = 2;
  constexpr int TROWS = TC_TH + 2 * DD, TWP = TC_TW + 2 * DD, NPATCH = CC * TROWS * TWP;
  const Net& net = a.net;
  const int gtid = threadIdx.x, wg = gtid >> 7, tid = gtid & 127, warp = tid >> 5;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem + (size_t)wg * a.a_bytes;
  uint8_t* sW = smem + (size_t)NWG * a.a_bytes;
  uint8_t* raw = sW + a.w_bytes + (size_t)wg * a.wg_bytes + align_up(NPATCH * 2, 128);   // same carve-up as tc_decode_kernel
  __shared__ __align__(8) uint64_t s_mbar_tma[NWG];
  __shared__ __align__(8) uint64_t s_mbar1[NWG];
  __shared__ __align__(8) uint64_t s_mbar2[NWG];
  __shared__ __align__(8) uint64_t s_mbar_rdy[NWG];      // "operands written / accumulator read": 128 arrivals per phase
  __shared__ uint32_t s_tmem;
  auto wg_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + wg), "r"(TC_THREADS) : "memory"); };

  {
    const int4* src = reinterpret_cast<const int4*>(a.blk);
    int4* dst = reinterpret_cast<int4*>(sW);
    for (int i = gtid; i < a.w_bytes / 16; i += THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const TcHeader* H = reinterpret_cast<const TcHeader*>(sW);
  if (a.run_if_exact >= 0 && (H->exact != 0) != (a.run_if_exact != 0)) return;   // the sibling launch handles this scene
  if (gtid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"(2 * TC_TMEM_COLS * NWG)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (gtid == 0) {
    for (int g = 0; g < NWG; ++g) {
      mbar_init(smem_u32(&s_mbar1[g]), 1);
      mbar_init(smem_u32(&s_mbar2[g]), 1);
      mbar_init(smem_u32(&s_mbar_tma[g]), 1);
      mbar_init(smem_u32(&s_mbar_rdy[g]), TC_THREADS);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem0 = s_tmem + (uint32_t)(wg * 2 * TC_TMEM_COLS);       // this warpgroup's two accumulators
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;                   // this warp's 32 lanes
  const uint32_t mbar1 = smem_u32(&s_mbar1[wg]), mbar2 = smem_u32(&s_mbar2[wg]), mbar_tma = smem_u32(&s_mbar_tma[wg]);
  const uint32_t mbar_rdy = smem_u32(&s_mbar_rdy[wg]);
  const uint32_t idesc = umma_idesc_f16(128, TC_BC);
  const uint32_t sA_u = smem_u32(sA), raw_u = smem_u32(raw);
  const float* bias = reinterpret_cast<const float*>(sW + H->off_bias);
  const float* w3t = reinterpret_cast<const float*>(sW + H->off_w3);
  const bool rel = net.relative != 0, relu = net.relu != 0;
  const int k1pad = H->k1pad;
  const int box_w = a.box_w, bw4 = box_w >> 2;
  const uint32_t box_bytes = (uint32_t)(box_w * TROWS * CC);
  const int pr = tid >> 4, px = tid & 15;
  const int bcol = a.box_lead - DD + px;
  const uint32_t o = (uint32_t)bcol & 3u;
  const uint32_t sel_a = 0x3210u + o * 0x1111u, sel_e = 0x4040u + o * 0x0101u;
  const uint32_t* rp = reinterpret_cast<const uint32_t*>(raw) + pr * bw4 + (bcol >> 2);
  const int tile0 = blockIdx.x * NWG + wg, tile_step = gridDim.x * NWG;

  auto tile_xy = [&](int tile, int& y0, int& x0) {
    const int ty = tile / a.tiles_x;
    y0 = net.row0 + ty * TC_TH;
    x0 = (tile - ty * a.tiles_x) * TC_TW;
  };
  auto interior = [&](int y0, int x0) {
    return a.use_tma && y0 - DD >= 0 && x0 - DD >= 0 && y0 - DD + TROWS <= net.H && x0 - DD + TWP <= net.W;
  };
  auto issue_tma = [&](int y0, int x0) {                 // all threads of the warpgroup; `raw` must be free
    if (warp == 0) {
      if (elect_one()) {
        mbar_expect_tx(mbar_tma, box_bytes);
        tma_load_3d(raw_u, a.tmap_dev, x0 - a.box_lead, y0 - DD - net.buf_row0, 0, mbar_tma);
      }
      __syncwarp();
    }
  };
  uint32_t ph_tma = 0, ph1 = 0, ph2 = 0, ph_rdy = 0;
  // staged bytes of (tile + halo) -> A1 in the A region; returns the packed centre bytes
  auto stage_and_build = [&](int y0, int x0, bool by_tma, int tile) {
    if (by_tma) {
      mbar_wait(mbar_tma, ph_tma, 2, tile, a.no_trap);
      ph_tma ^= 1;
    } else {                                              // border tile: plain loads with reflection, TMA box layout
      for (int e = tid; e < NPATCH; e += TC_THREADS) {
        const int c = e / (TROWS * TWP), rem = e - c * TROWS * TWP, r = rem / TWP, x = rem - r * TWP;
        const int gy = reflect_clamp(y0 - DD + r, net.H), gx = reflect_clamp(x0 - DD + x, net.W);
        raw[(c * TROWS + r) * box_w + (a.box_lead - DD) + x] =
            (uint8_t)load_msb_int(a.msb, 0, ((size_t)c * net.buf_rows + (gy - net.buf_row0)) * net.W + gx);
      }
      wg_sync();
    }
    return build_a1_raw8<CC>(rp, bw4, sel_a, sel_e, rel, sA, tid);
  };
  // Hand-off to the tensor core WITHOUT a warpgroup barrier.  Every thread announces that its operand rows are written
  // (and its reads of the accumulator about to be overwritten, and of `raw`, are done) by arriving on `mbar_rdy`; only
  // warp 0 waits for the 128 arrivals, then one elected lane issues.  The other three warps go straight on to work that
  // does not depend on this MMA.  A thread can never run two phases ahead: between two of its arrivals it waits for an
  // MMA completion, which the issuer only produces after the previous phase completed.
  auto issue_mma = [&](int layer, uint32_t acc, uint32_t mbar) {
    fence_async_smem();          // generic-proxy writes of the operand (and reads of `raw`) -> ordered before the async proxy
    tc_fence_before();
    mbar_arrive(mbar_rdy);
    if (warp == 0) {
      mbar_wait(mbar_rdy, ph_rdy, 4, layer, a.no_trap);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sB_u = smem_u32(sW + H->off_b[layer]);
        if (layer == 0) {
          for (int i = 0; i < k1pad / 16; ++i)
            umma_f16(acc, umma_desc(sA_u + i * 2 * 2048, 2048, 128), umma_desc(sB_u + i * 2 * 1024, 1024, 128), idesc, i > 0);
        } else {
          for (int i = 0; i < 2 * TC_BC / 16; ++i)       // hi half then lo half of A2, both against the same B
            umma_f16(acc, umma_desc(sA_u + i * 2 * 2048, 2048, 128),
                     umma_desc(sB_u + (i % (TC_BC / 16)) * 2 * 1024, 1024, 128), idesc, i > 0);
        }
        umma_commit(mbar);
      }
      __syncwarp();
    }
    ph_rdy ^= 1;
  };
  // 16 hidden units of one layer: scale + bias, activation (result left in h)
  auto activate16 = [&](int layer, int cb, uint32_t acc_addr, float (&h)[16]) {
    const float scale = H->scale[layer];
    const float* bl = bias + layer * TC_BC;
    float bterm[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b4 = *reinterpret_cast<const float4*>(bl + cb + 4 * q);
      bterm[4 * q] = b4.x; bterm[4 * q + 1] = b4.y; bterm[4 * q + 2] = b4.z; bterm[4 * q + 3] = b4.w;
    }
    float acc[16];
    tmem_ld16(acc_addr + cb, acc);
#pragma unroll
    for (int j = 0; j < 16; j += 2)                          // = w0 * z (w0 folded into scale and bias), two per FFMA2
      ffma2(acc[j], acc[j + 1], acc[j], acc[j + 1], scale, scale, bterm[j], bterm[j + 1]);
    if (relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) h[j] = fmaxf(acc[j], 0.f);
    } else {
#pragma unroll
      for (int j = 0; j < 16; j += 2) tc_sine2<SINE>(acc[j], acc[j + 1], h[j], h[j + 1]);
      if ((H->guard_mask >> layer) & 1) {                    // uniform: tc_prep_kernel could not bound |w0 z| for this layer
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(acc[j]));
        if (__builtin_expect(!(amax <= 20000.0f), 0)) {
#pragma unroll
          for (int j = 0; j < 16; ++j) h[j] = sin_slow(acc[j]);
        }
      }
    }
  };
  auto epilogue_hidden = [&](uint32_t acc_addr) {          // layer 0: h = hi + lo (fp16 pairs) -> A2 chunks
#pragma unroll 1
    for (int cb = 0; cb < TC_BC; cb += 16) {
      float h[16];
      activate16(0, cb, acc_addr, h);
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const __half2 hh = __floats2half2_rn(h[j], h[j + 1]);
        const float2 back = __half22float2(hh);
        float l0, l1;
        ffma2(l0, l1, back.x, back.y, -1.0f, -1.0f, h[j], h[j + 1]);          // h - hi, exact
        const __half2 ll = __floats2half2_rn(l0, l1);
        hi[j >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
        lo[j >> 1] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      const int kc = cb >> 3;
      *reinterpret_cast<uint4*>(sA + (size_t)(kc * 128 + tid) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sA + (size_t)((kc + 1) * 128 + tid) * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      *reinterpret_cast<uint4*>(sA + (size_t)((TC_BC / 8 + kc) * 128 + tid) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(sA + (size_t)((TC_BC / 8 + kc + 1) * 128 + tid) * 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    }
  };
  auto epilogue_last = [&](uint32_t acc_addr, float (&y)[4]) {   // layer 1 + output layer sums (fp32 FFMA)
    y[0] = y[1] = y[2] = y[3] = 0.f;
#pragma unroll 1
    for (int cb = 0; cb < TC_BC; cb += 16) {
      float h[16];
      activate16(1, cb, acc_addr, h);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 wa = *reinterpret_cast<const float4*>(w3t + (cb + j) * 8);
        ffma2(y[0], y[1], wa.x, wa.y, h[j], h[j], y[0], y[1]);
        ffma2(y[2], y[3], wa.z, wa.w, h[j], h[j], y[2], y[3]);
      }
    }
  };
  // sigmoid, inverse quantisation, integer write (decode.py:131-134)
  auto finish = [&](const float (&y)[4], uint32_t cpk, int ty0, int tx0) {
    const int gy = ty0 + pr, gx = tx0 + px;
    if (gy < net.row1 && gx < net.W) {
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        const float yz = y[c] + w3t[TC_BC * 8 + c];
        const float yy = SINE == SINE_MUFU ? sigmoidf_fast(yz) : sigmoidf_rn(yz);
        const int res = (int)rintf(yy * net.qmax);
        const uint32_t m = (cpk >> (8 * c)) & 0xFFu;
        a.out[((size_t)c * net.buf_rows + (gy - net.buf_row0)) * net.W + gx] = (uint16_t)((m << net.K) + (uint32_t)res);
      }
    }
  };

  if (tile0 < a.n_tiles) {
    int cy = 0, cx = 0, ny = 0, nx = 0, py = 0, pxx = 0;
    tile_xy(tile0, cy, cx);
    bool c_tma = interior(cy, cx);
    if (c_tma) issue_tma(cy, cx);
    bool n_valid = tile0 + tile_step < a.n_tiles, n_tma = false;
    if (n_valid) {
      tile_xy(tile0 + tile_step, ny, nx);
      n_tma = interior(ny, nx);
    }
    uint32_t cpk = stage_and_build(cy, cx, c_tma, tile0), cpk_prev = 0;
    issue_mma(0, tmem0, mbar1);
    if (n_valid && n_tma) issue_tma(ny, nx);
    float yprev[4] = {0.f, 0.f, 0.f, 0.f};
    bool have_prev = false;
    uint32_t p = 0;
    for (int t = tile0; t < a.n_tiles; t += tile_step) {
      const uint32_t acc_p = tmem0 + p * TC_TMEM_COLS + lane_off;
      // a: first hidden layer's accumulator
      mbar_wait(mbar1, ph1, 1, t, a.no_trap);
      ph1 ^= 1;
      tc_fence_after();
      // b: epilogue 1 -> A2, second layer's MMAs
      epilogue_hidden(acc_p);
      issue_mma(1, tmem0 + p * TC_TMEM_COLS, mbar2);
      // c: the previous tile's pixels; where tile i+2 lies
      if (have_prev) finish(yprev, cpk_prev, py, pxx);
      int n2y = 0, n2x = 0;
      const bool n2_valid = t + 2 * tile_step < a.n_tiles;
      bool n2_tma = false;
      if (n2_valid) {
        tile_xy(t + 2 * tile_step, n2y, n2x);
        n2_tma = interior(n2y, n2x);
      }
      // d: second hidden layer's accumulator complete, A region free
      mbar_wait(mbar2, ph2, 3, t, a.no_trap);
      ph2 ^= 1;
      tc_fence_after();
      // e: next tile's operand and first-layer MMAs
      uint32_t cpk_next = 0;
      if (n_valid) {
        cpk_next = stage_and_build(ny, nx, n_tma, t + tile_step);
        issue_mma(0, tmem0 + (p ^ 1u) * TC_TMEM_COLS, mbar1);
        if (n2_valid && n2_tma) issue_tma(n2y, n2x);
      }
      // f: epilogue 2 of this tile
      epilogue_last(acc_p, yprev);
      tc_fence_before();           // our tcgen05.ld's are ordered before the MMAs that next write this accumulator
      cpk_prev = cpk; py = cy; pxx = cx; have_prev = true;
      cy = ny; cx = nx; cpk = cpk_next;
      ny = n2y; nx = n2x; n_tma = n2_tma; n_valid = n2_valid;
      p ^= 1u;
    }
    if (have_prev) finish(yprev, cpk_prev, py, pxx);
  }
  tc_fence_before();
  __syncthreads();
  if (gtid < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(2 * TC_TMEM_COLS * NWG) : "memory");
}

// ---- self-test: D[128][64] = A[128][K] * B[64][K]^T through the same descriptors / layouts / TMEM path ----------------
__global__ void __launch_bounds__(TC_THREADS) tc_selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ B,
                                                                float* __restrict__ Dout, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  __shared__ __align__(8) uint64_t s_mbar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K; i += TC_THREADS) {
    const int r = i / K, k = i - r * K;
    *reinterpret_cast<__half*>(sA + umma_off(128, r, k)) = A[i];
  }
  for (int i = tid; i < TC_BC * K; i += TC_THREADS) {
    const int r = i / K, k = i - r * K;
    *reinterpret_cast<__half*>(sB + umma_off(TC_BC, r, k)) = B[i];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) mbar_init(smem_u32(&s_mbar), 1);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(128, TC_BC);
    for (int i = 0; i < K / 16; ++i)
      umma_f16(tmem, umma_desc(smem_u32(sA) + i * 2 * 2048, 2048, 128), umma_desc(smem_u32(sB) + i * 2 * 1024, 1024, 128),
               idesc, i > 0);
    umma_commit(smem_u32(&s_mbar));
  }
  mbar_wait(smem_u32(&s_mbar), 0);
  tc_fence_after();
  for (int cb = 0; cb < TC_BC; cb += 16) {
    float acc[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + cb, acc);
    for (int j = 0; j < 16; ++j) Dout[tid * TC_BC + cb + j] = acc[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
}

// ---- self-test 2: D[128][N] = A[128][K] * B[N][K]^T with either operand in the MN-major "[group of 8][k][8]" layout
// (the layout the kernels write activations in, read here with the roles of the two dimensions swapped) -----------------
__global__ void __launch_bounds__(TC_THREADS) tc_selftest2_kernel(const __half* __restrict__ A, const __half* __restrict__ B,
                                                                 float* __restrict__ Dout, int N, int K, int a_mn, int b_mn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  __shared__ __align__(8) uint64_t s_mbar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K; i += TC_THREADS) {
    const int r = i / K, k = i - r * K;
    *reinterpret_cast<__half*>(sA + (a_mn ? umma_off_mn(K, r, k) : umma_off(128, r, k))) = A[i];
  }
  for (int i = tid; i < N * K; i += TC_THREADS) {
    const int r = i / K, k = i - r * K;
    *reinterpret_cast<__half*>(sB + (b_mn ? umma_off_mn(K, r, k) : umma_off(N, r, k))) = B[i];
  }
  const int cols = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) mbar_init(smem_u32(&s_mbar), 1);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16_major(128, N, a_mn, b_mn);
    for (int i = 0; i < K / 16; ++i) {
      // K-major: k-chunk stride rows*16 (LBO), 8-row group stride 128 (SBO).  MN-major: k-group stride 128 (LBO),
      // MN-group stride K*16 (SBO); one instruction consumes 16 k = 2 k-groups -> advance 256 B per step.
      const uint64_t da = a_mn ? umma_desc(smem_u32(sA) + i * 256, 128, K * 16) : umma_desc(smem_u32(sA) + i * 2 * 2048, 2048, 128);
      const uint64_t db = b_mn ? umma_desc(smem_u32(sB) + i * 256, 128, K * 16) : umma_desc(smem_u32(sB) + i * 2 * N * 16, N * 16, 128);
      umma_f16(tmem, da, db, idesc, i > 0);
    }
    umma_commit(smem_u32(&s_mbar));
  }
  mbar_wait(smem_u32(&s_mbar), 0);
  tc_fence_after();
  for (int cb = 0; cb < N; cb += 16) {
    float acc[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + cb, acc);
    for (int j = 0; j < 16; ++j) Dout[tid * N + cb + j] = acc[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
}


// ---- self-test 3: D[64][N] = A . B^T with M = 64 (the shape of the training step's chunk GEMMs), operands given as raw
// "images" (lbdrn_umma.cuh: img_off) and consumed through the K-major or the MN-major view; read back with the 16x256b
// shape (Dout) and, for diagnosis, with 32x32b over all 128 lanes (Raw[128][N]) -------------------------------------------
__global__ void __launch_bounds__(TC_THREADS) tc_selftest3_kernel(const uint8_t* __restrict__ A, int a_bytes,
                                                                 const uint8_t* __restrict__ B, int b_bytes,
                                                                 float* __restrict__ Dout, float* __restrict__ Raw, int N,
                                                                 int ksteps, int a_mn, int a_rows, int b_mn, int b_rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((a_bytes + 127) & ~127);
  __shared__ __align__(8) uint64_t s_mbar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < a_bytes; i += TC_THREADS) sA[i] = A[i];
  for (int i = tid; i < b_bytes; i += TC_THREADS) sB[i] = B[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) mbar_init(smem_u32(&s_mbar), 1);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  // poison the accumulator region so that untouched lanes are visible in Raw
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc_f16_major(64, N, a_mn, b_mn);
    for (int i = 0; i < ksteps; ++i)
      umma_f16(tmem, img_desc(smem_u32(sA), a_rows, a_mn, i), img_desc(smem_u32(sB), b_rows, b_mn, i), idesc, i > 0);
    umma_commit(smem_u32(&s_mbar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&s_mbar), 0);
  tc_fence_after();
  const int g = lane >> 2, t = lane & 3;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[4];
    tmem_ld16x256_x1(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait4(r);
    Dout[(16 * warp + g) * N + c0 + 2 * t] = __uint_as_float(r[0]);
    Dout[(16 * warp + g) * N + c0 + 2 * t + 1] = __uint_as_float(r[1]);
    Dout[(16 * warp + g + 8) * N + c0 + 2 * t] = __uint_as_float(r[2]);
    Dout[(16 * warp + g + 8) * N + c0 + 2 * t + 1] = __uint_as_float(r[3]);
  }
  if (Raw != nullptr)
    for (int c0 = 0; c0 < N; c0 += 16) {
      float acc[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
      for (int j = 0; j < 16 && c0 + j < N; ++j) Raw[tid * N + c0 + j] = acc[j];
    }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

std::mutex& tc_mu() {
  static std::mutex m;
  return m;
}
uint8_t* g_blk[64] = {nullptr};
constexpr int kBlkBytes = 1 << 18;

}  // namespace

bool tc_supported(const Net& n) {
  // bc = 64; colour features as integer differences (exact in fp16 up to 2048), coordinate / positional features as
  // fp32 row / column tables added in the first epilogue; up to 8 bands, D <= 3
  return n.bc == TC_BC && n.ncol <= TC_MAX_K1 && n.maxv <= 2048.0f && n.C <= kMaxC && n.nl >= 1 && n.nl <= 4 &&
         (n.ncol == 0 || n.D <= 3);
}

// 3-D tiled TMA descriptor of CHW planes: dims (W, rows, C) of `es`-byte elements, box (box_w, box_h, box_c).  *out stays
// nullptr when TMA cannot address the buffer (base / pitches not multiples of 16 B, driver entry point missing).  The
// descriptor is copied into a per-device ring in global memory (a slot per call, so queued launches do not race).
int make_tensor_map_3d(const void* base, int es, int W, int rows, int C, int box_w, int box_h, int box_c, int dev,
                       cudaStream_t st, const CUtensorMap** out) {
  *out = nullptr;
  const size_t pitch = (size_t)W * es, plane = pitch * rows;
  if (getenv("LBDRN_NO_TMA") || ((uintptr_t)base % 16) != 0 || pitch % 16 != 0 || plane % 16 != 0 || box_w > 256 ||
      box_h > 256 || (box_w * es) % 16 != 0)
    return LBDRN_OK;
  // resolved through the runtime so the library has no link-time dependency on libcuda (it must load on CPU-only boxes)
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeFn>(fn);
  }
  if (!encode) return LBDRN_OK;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)rows, (cuuint64_t)C};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)plane};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_c};
  cuuint32_t estr[3] = {1, 1, 1};
  alignas(64) CUtensorMap tm;
  CUresult r = encode(&tm, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                      const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return LBDRN_OK;
  static std::mutex mu;
  static CUtensorMap* ring[64] = {nullptr};
  static unsigned slot[64] = {0};
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 0 || dev >= 64) return fail(LBDRN_E_UNSUPPORTED, "device ordinal %d", dev);
  if (!ring[dev]) CUDA_TRY(cudaMalloc(&ring[dev], 64 * sizeof(CUtensorMap)));
  CUtensorMap* dst = ring[dev] + (slot[dev]++ % 64);
  CUDA_TRY(cudaMemcpyAsync(dst, &tm, sizeof tm, cudaMemcpyHostToDevice, st));
  *out = dst;
  return LBDRN_OK;
}

namespace {

using KernT = void (*)(const TcArgs);

// sine: SINE_POLY / SINE_CW_MUFU / SINE_MUFU (the last one is instantiated for the paper's configuration -- uint8 planes,
// C = 4, D = 2, colour features -- and falls back to SINE_CW_MUFU elsewhere).  pf: accumulator loads one block ahead
// (two-warpgroup decode kernel of that configuration only).
template <int SINE, bool WLO, int MODE, int NWG>
KernT pick_kernel(const Net& n, bool pf = false) {
  constexpr int S = SINE == SINE_MUFU ? SINE_CW_MUFU : SINE;      // variant used outside the RAW8 family
  const bool raw8 = tc_raw8(n);
  if (n.nco) {                                       // coordinate features (USE_COLORS off: k1 = 0, table-driven kernel)
    if (n.ncol && n.C == 4 && n.D == 2)
      return raw8 ? tc_decode_kernel<S, 4, 2, WLO, MODE, NWG, true, true> : tc_decode_kernel<S, 4, 2, WLO, MODE, NWG, true>;
    return tc_decode_kernel<S, 0, 0, WLO, MODE, NWG, true>;
  }
  if (n.C == 4 && n.D == 2) {
    if (raw8) {
      if constexpr (!WLO && MODE == TC_DECODE && NWG == 2 && SINE != SINE_POLY)
        if (pf) return tc_decode_kernel<SINE, 4, 2, WLO, MODE, NWG, false, true, true>;
      return tc_decode_kernel<SINE, 4, 2, WLO, MODE, NWG, false, true>;
    }
    return tc_decode_kernel<S, 4, 2, WLO, MODE, NWG>;
  }
  if (n.C == 8 && n.D == 2) return tc_decode_kernel<S, 8, 2, WLO, MODE, NWG>;
  if (n.C == 4 && n.D == 1) return tc_decode_kernel<S, 4, 1, WLO, MODE, NWG>;
  if (n.C == 4 && n.D == 3) return tc_decode_kernel<S, 4, 3, WLO, MODE, NWG>;
  return tc_decode_kernel<S, 0, 0, WLO, MODE, NWG>;
}

template <bool WLO, int NWG>
KernT pick_decode(const Net& n, int sine, bool pf) {
  if (sine == SINE_MUFU) return pick_kernel<SINE_MUFU, WLO, TC_DECODE, NWG>(n, pf);
  if (sine == SINE_CW_MUFU) return pick_kernel<SINE_CW_MUFU, WLO, TC_DECODE, NWG>(n, pf);
  return pick_kernel<SINE_POLY, WLO, TC_DECODE, NWG>(n, pf);
}

// TMA descriptor of the MSB planes: dims (W, buf_rows, C); needs a 16 B-aligned base and row / plane pitches.
int tc_setup_tma(TcArgs& a, int dev, cudaStream_t st) {
  const Net& n = a.net;
  const int es = n.msb_u16 ? 2 : 1, trows = TC_TH + 2 * n.D;
  const int al = 16 / es, lead = (n.D + al - 1) / al * al, box_w = lead + TC_TW + lead;
  const size_t pitch = (size_t)n.W * es, plane = pitch * n.buf_rows;
  a.use_tma = 0;
  a.box_w = box_w;
  a.box_lead = lead;
  if (n.W < box_w) return LBDRN_OK;
  const CUtensorMap* dst = nullptr;
  int rc = make_tensor_map_3d(a.msb, es, n.W, n.buf_rows, n.C, box_w, trows, n.C, dev, st, &dst);
  if (rc) return rc;
  if (!dst) return LBDRN_OK;
  a.tmap_dev = dst;
  a.use_tma = 1;
  return LBDRN_OK;
}

// dynamic shared memory of one CTA (fills the region sizes the kernel carves it up with)
size_t tc_smem_bytes(TcArgs& a, const TcHeader& h, bool wlo, int nwg) {
  const Net& n = a.net;
  const int kmax = h.k1pad > 2 * TC_BC ? h.k1pad : 2 * TC_BC;
  a.a_bytes = align_up(128 * kmax * 2, 1024);
  a.w_bytes = wlo ? h.total : h.hi_bytes;
  const int n_patch = n.C * (TC_TH + 2 * n.D) * (TC_TW + 2 * n.D);
  const int raw_bytes = a.box_w * (n.msb_u16 ? 2 : 1) * (TC_TH + 2 * n.D) * n.C;
  a.wg_bytes = align_up(n_patch * 2, 128) + align_up(raw_bytes, 128);
  return (size_t)nwg * a.a_bytes + a.w_bytes + (size_t)nwg * a.wg_bytes + 2 * (TC_MAX_K1 + 16) * 2 + 64;
}

// one launch of the tensor kernel family (wlo: smem holds the low-order weight operands too; nwg: warpgroups per CTA)
int tc_launch(KernT kern, TcArgs& a, const TcHeader& h, bool wlo, int nwg, int dev, cudaStream_t st, int acc_per_wg = 1) {
  const Net& n = a.net;
  const int threads = TC_THREADS * nwg;
  const size_t smem = tc_smem_bytes(a, h, wlo, nwg);
  int sms = 0, max_smem = 0, smem_sm = 0, regs_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
  if (smem > (size_t)max_smem) return fail(LBDRN_E_UNSUPPORTED, "tensor-core decode needs %zu B of shared memory", smem);
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // Resident CTAs per SM.  cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for kernels that allocate tensor
  // memory (measured on B200 / CUDA 12.9) although the hardware co-schedules as many CTAs as shared memory, registers
  // and the 512 TMEM columns allow (3 here: 4.3 vs 1.7 Gpix/s), so the limit is computed from the kernel's attributes.
  cudaFuncAttributes fa;
  CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
  int occ = (int)(smem_sm / (smem + fa.sharedSizeBytes + 1024));   // +1 KB: per-CTA reservation of the driver
  const int regs_cta = ((fa.numRegs + 7) / 8 * 8) * threads;
  if (regs_cta > 0 && regs_sm / regs_cta < occ) occ = regs_sm / regs_cta;
  const int tmem_cta = TC_TMEM_COLS * nwg * acc_per_wg;
  if (occ * tmem_cta > 512) occ = 512 / tmem_cta;                 // TMEM: 512 columns per SM
  const int occ_max = nwg >= 4 ? 1 : (nwg == 2 ? 2 : 3);          // measured optimum for NWG=1 (4 CTAs: 3.0 vs 4.3 Gpix/s)
  if (occ > occ_max) occ = occ_max;
  if (const char* e = getenv("LBDRN_TC_OCC")) occ = atoi(e);
  if (occ < 1) return fail(LBDRN_E_UNSUPPORTED, "tensor-core decode kernel cannot be made resident");
  a.tiles_x = (n.W + TC_TW - 1) / TC_TW;
  a.n_tiles = a.tiles_x * ((n.row1 - n.row0 + TC_TH - 1) / TC_TH);
  int grid = sms * occ;                                          // persistent: whole CTAs per SM
  if (grid * nwg > a.n_tiles) grid = (a.n_tiles + nwg - 1) / nwg;
  if (getenv("LBDRN_DEBUG"))
    fprintf(stderr, "[lbdrn] tc kernel: smem dyn %zu static %zu regs %d occ %d grid %d x %d thr tiles %d wlo %d tma %d box_w %d\n",
            smem, fa.sharedSizeBytes, fa.numRegs, occ, grid, threads, a.n_tiles, (int)wlo, a.use_tma, a.box_w);
  a.no_trap = getenv("LBDRN_DEBUG") != nullptr;
  kern<<<grid, threads, smem, st>>>(a);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  if (a.no_trap) {
    int h4[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaMemcpyFromSymbol(h4, g_tc_timeout, sizeof h4));
    if (h4[3]) fprintf(stderr, "[lbdrn] tc kernel: %d barrier timeouts; last: barrier %d (1=mma 2=tma) tile %d block %d\n", h4[3], h4[0], h4[1], h4[2]);
  }
  return LBDRN_OK;
}

int tc_prepare(const Net& n, const float* params, TcHeader& h, uint8_t*& blk, int& dev, cudaStream_t st) {
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(LBDRN_E_UNSUPPORTED, "device ordinal %d", dev);
  {
    std::lock_guard<std::mutex> lk(tc_mu());
    if (!g_blk[dev]) CUDA_TRY(cudaMalloc(&g_blk[dev], kBlkBytes));
  }
  blk = g_blk[dev];
  plan_block(n, h);
  if (h.total > kBlkBytes) return fail(LBDRN_E_UNSUPPORTED, "tensor-core weight block too large (%d B)", h.total);
  tc_prep_kernel<<<1, 1024, 0, st>>>(n, h, params, blk);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

float* g_ctab[64] = {nullptr};
size_t g_ctab_n[64] = {0};

// row / column contribution tables of the coordinate features (no-op without USE_COORDINATES)
int tc_coord_tables(const Net& n, const float* params, const float* tab, int dev, TcArgs& a, cudaStream_t st) {
  a.rtab = a.ctab = nullptr;
  if (!n.nco) return LBDRN_OK;
  if (!tab) return fail(LBDRN_E_INVALID, "USE_COORDINATES set but coord_tab_dev is NULL");
  const size_t need = (size_t)(n.H + n.W) * TC_BC;
  {
    std::lock_guard<std::mutex> lk(tc_mu());
    if (g_ctab_n[dev] < need) {
      if (g_ctab[dev]) CUDA_TRY(cudaFree(g_ctab[dev]));     // synchronises: no launch still reads the old tables
      g_ctab[dev] = nullptr; g_ctab_n[dev] = 0;
      CUDA_TRY(cudaMalloc(&g_ctab[dev], need * sizeof(float)));
      g_ctab_n[dev] = need;
    }
  }
  a.rtab = g_ctab[dev];
  a.ctab = g_ctab[dev] + (size_t)n.H * TC_BC;
  const int blocks = (int)((need + 255) / 256);
  tc_coord_tables_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, st>>>(n, params, tab, g_ctab[dev], g_ctab[dev] + (size_t)n.H * TC_BC);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

}  // namespace

int tc_decode(const Net& n, const void* msb, const float* params, const float* tab, uint16_t* out, int fast_sine,
              cudaStream_t st) {
  TcHeader h;
  uint8_t* blk = nullptr;
  int dev = 0, rc = tc_prepare(n, params, h, blk, dev, st);
  if (rc) return rc;
  TcArgs a;
  memset(&a, 0, sizeof a);
  a.net = n; a.msb = msb; a.blk = blk; a.out = out;
  rc = tc_coord_tables(n, params, tab, dev, a, st);
  if (rc) return rc;
  rc = tc_setup_tma(a, dev, st);
  if (rc) return rc;
  // two sibling launches; the exactness flag computed by tc_prep_kernel decides ON THE DEVICE which one does the work
  a.run_if_exact = 1;
  const bool two = a.use_tma && !getenv("LBDRN_TC_NWG1");     // TMA-addressable input: two warpgroups per CTA
  const char* pfe = getenv("LBDRN_TC_PF");
  const bool pf = pfe ? atoi(pfe) != 0 : true;
  // the paper's configuration on TMA-addressable planes: the software-pipelined kernel (two accumulators per warpgroup)
  const bool pipe = two && tc_raw8(n) && n.nl == 2 && n.nco == 0 && getenv("LBDRN_TC_NOPIPE") == nullptr;
  if (pipe) {
    KernT k = fast_sine == SINE_MUFU ? (KernT)tc_pipe_kernel<SINE_MUFU>
                                     : (fast_sine == SINE_CW_MUFU ? (KernT)tc_pipe_kernel<SINE_CW_MUFU> : (KernT)tc_pipe_kernel<SINE_POLY>);
    rc = tc_launch(k, a, h, false, 2, dev, st, 2);
  } else {
    rc = two ? tc_launch(pick_decode<false, 2>(n, fast_sine, pf), a, h, false, 2, dev, st)
             : tc_launch(pick_decode<false, 1>(n, fast_sine, pf), a, h, false, 1, dev, st);
  }
  if (rc) return rc;
  a.run_if_exact = 0;
  return tc_launch(pick_decode<true, 1>(n, fast_sine, pf), a, h, true, 1, dev, st);
}

int tc_eval_sse(const Net& n, const void* msb, const void* lsb, const float* params, const float* tab, double* sse_out,
                cudaStream_t st) {
  TcHeader h;
  uint8_t* blk = nullptr;
  int dev = 0, rc = tc_prepare(n, params, h, blk, dev, st);
  if (rc) return rc;
  Scratch* sc = nullptr;
  rc = get_scratch(n.P, sc);
  if (rc) return rc;
  TcArgs a;
  memset(&a, 0, sizeof a);
  a.net = n; a.msb = msb; a.blk = blk; a.lsb = lsb;
  a.partials = sc->partials; a.counter = sc->counter; a.sse_out = sse_out;
  a.run_if_exact = -1;
  rc = tc_coord_tables(n, params, tab, dev, a, st);
  if (rc) return rc;
  rc = tc_setup_tma(a, dev, st);
  if (rc) return rc;
  int max_smem = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // four warpgroups sharing one weight block where that fits (C = 4, D = 2: 193 KB); wider inputs keep one warpgroup per CTA
  if (a.use_tma && !getenv("LBDRN_TC_NWG1") && tc_smem_bytes(a, h, true, 4) + 2048 <= (size_t)max_smem)
    return tc_launch(pick_kernel<SINE_CW_MUFU, true, TC_SSE, 4>(n), a, h, true, 4, dev, st);
  return tc_launch(pick_kernel<SINE_CW_MUFU, true, TC_SSE, 1>(n), a, h, true, 1, dev, st);
}

int tc_selftest2(const void* a_dev, const void* b_dev, float* d_dev, int N, int K, int a_mn, int b_mn, cudaStream_t st) {
  if (K % 16 || K < 16 || K > 256 || N % 16 || N < 16 || N > 256) return fail(LBDRN_E_INVALID, "selftest2 N=%d K=%d", N, K);
  const size_t smem = (size_t)(128 + N) * K * 2;
  CUDA_TRY(cudaFuncSetAttribute(tc_selftest2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest2_kernel<<<1, TC_THREADS, smem, st>>>((const __half*)a_dev, (const __half*)b_dev, d_dev, N, K, a_mn, b_mn);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

int tc_selftest3(const void* a_img, int a_bytes, const void* b_img, int b_bytes, float* d_dev, float* raw_dev, int N,
                 int ksteps, int a_mn, int a_rows, int b_mn, int b_rows, cudaStream_t st) {
  if (N % 8 || N < 8 || N > 256 || ksteps < 1 || a_bytes <= 0 || b_bytes <= 0 || a_bytes + b_bytes > 200 * 1024)
    return fail(LBDRN_E_INVALID, "selftest3 N=%d ksteps=%d", N, ksteps);
  const size_t smem = (size_t)((a_bytes + 127) & ~127) + b_bytes + 128;
  CUDA_TRY(cudaFuncSetAttribute(tc_selftest3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest3_kernel<<<1, TC_THREADS, smem, st>>>((const uint8_t*)a_img, a_bytes, (const uint8_t*)b_img, b_bytes, d_dev,
                                                  raw_dev, N, ksteps, a_mn, a_rows, b_mn, b_rows);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

int tc_selftest(const void* a_dev, const void* b_dev, float* d_dev, int K, cudaStream_t st) {
  if (K % 16 || K < 16 || K > 256) return fail(LBDRN_E_INVALID, "selftest K=%d", K);
  const size_t smem = (size_t)(128 + TC_BC) * K * 2;
  CUDA_TRY(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, TC_THREADS, smem, st>>>((const __half*)a_dev, (const __half*)b_dev, d_dev, K);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

}  // namespace lbdrn
