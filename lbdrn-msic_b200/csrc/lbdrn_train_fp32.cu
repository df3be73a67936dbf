// lbdrn_train_fp32.cu -- instantiations, planning and launch of the fused fp32 training kernel.
#include <cstdio>
#include <cstdlib>

#include "lbdrn_train_fp32.cuh"
#include "lbdrn_internal.h"

namespace lbdrn {
namespace {

#ifndef LBDRN_TRAIN_THREADS
#define LBDRN_TRAIN_THREADS 512
#endif
constexpr int kTT = LBDRN_TRAIN_THREADS;     // threads per CTA = 4*US warps working on one 64-pixel chunk

template <int BC, int CP, int TM>
int plan_t(const Net& n, int batch_size, int sms, int max_smem, TrainPlan& t) {
  constexpr int kTrainLDP = train_ldp(TM), kTrainNPIX = train_npix(TM);
  const size_t acts = ((size_t)t.dimpad + 2 * (size_t)n.nl * BC + (2 + kTT / 128) * CP) * kTrainLDP;
  const size_t wts = (size_t)round4(n.P) + (size_t)(n.nl - 1) * BC * BC;
  const size_t with_w = (acts + wts) * sizeof(float), without = acts * sizeof(float);
  // the opt-in limit covers dynamic + static shared memory of the kernel (pixel coordinates, reduction scratch); the static
  // part differs between instantiations, so every candidate is checked against its own (fits() below)
  const int dev_max_smem = max_smem;
  auto fits = [&](void* kern, size_t dyn) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) { cudaGetLastError(); return false; }
    return dyn + fa.sharedSizeBytes <= (size_t)dev_max_smem;
  };
  {
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, (const void*)train_fp32_kernel<BC, CP, false, kTT, TM>));
    max_smem -= (int)fa.sharedSizeBytes;        // placement of the optional TMA boxes (place_boxes) uses this estimate
  }
  // 64-pixel chunks with bc a multiple of 32: the chunk's GEMMs run as warp-level 3xTF32 tensor-core MMAs
  // (32-pixel chunks: bc 256, where a 64-pixel chunk's activations do not fit -- 3xTF32 only)
  constexpr bool kMma = kTT == 512 && BC % 32 == 0 && ((TM == 4 && BC <= 128) || (TM == 2 && BC % 64 == 0));
  const bool mma = kMma && getenv("LBDRN_TRAIN_FFMA") == nullptr;
  // landing boxes of the TMA neighbourhood gather (uint8 planes, 5x5 window, <= 4 bands, colours only, TMA-addressable
  // rows): 32 x 5 x C bytes of MSB (rounded up to 128) + a 128 B slot for the 16 x 1 x C label box, per pixel.
  // OPT-IN (LBDRN_TRAIN_TMA=1): measured on B200 it does not beat the register prefetch -- 37.5 vs 36.9 us/step at 8192^2,
  // 35.7 vs 33.9 at 2048^2: issuing gets cheap (3.2k vs 7.3k cycles at 8192^2) but the copies then contend with the
  // L2-latency-bound reduction phase (4.7k -> 10.6k cycles), and 128 B-aligned landing slots put the same byte of every
  // pixel in the same bank (the normalisation pass reads them with ~8-way conflicts).  Kept for the record and for
  // scenes that do not fit L2 on parts with a slower LSU path.
  size_t pf_bytes = 0;
  if (TM == 4 && n.nco == 0 && n.ncol != 0 && n.n == 5 && n.C <= 4 && !n.msb_u16 && !n.lsb_u16 && n.W % 16 == 0 &&
      ((size_t)n.W * n.buf_rows) % 16 == 0 && n.buf_row0 == 0 && n.buf_rows == n.H && getenv("LBDRN_TRAIN_TMA") != nullptr) {
    t.pf_stride = ((32 * 5 * n.C + 127) & ~127) + 128;
    pf_bytes = (size_t)kTrainNPIX * t.pf_stride + 128;      // + slack: the boxes are aligned to 128 B at run time
  }
  auto place_boxes = [&](size_t base) {        // boxes go behind the activations / weights if they still fit
    const size_t off = (base + 127) & ~(size_t)127;
    if (pf_bytes && off + pf_bytes <= (size_t)max_smem) { t.pf_off = (int)off; return off + pf_bytes; }
    t.pf_off = 0; t.pf_stride = 0;
    return base;
  };
  // fp16 hi+lo split operands + ldmatrix (MMA = 2): bc a multiple of 64, everything resident in shared memory
  constexpr bool kH2 = kMma && TM == 4 && BC % 64 == 0;
  if constexpr (kH2) {
    const int KP0 = round16(n.dim_in), L = n.nl;
    const size_t h2 = ((size_t)t.dimpad * kTrainLDP + (size_t)BC * kTrainLDP + (size_t)L * BC * kLDH +
                       (size_t)(2 + kTT / 128) * CP * kTrainLDP + (size_t)round4(L * BC + n.C * BC + n.C) +
                       (size_t)KP0 * kLDH + (size_t)(L - 1) * BC * kLDH + (size_t)BC * (KP0 + 8) +
                       (size_t)(L - 1) * BC * (BC + 8)) * sizeof(float);
    if (mma && getenv("LBDRN_TRAIN_TF32") == nullptr && h2 <= (size_t)max_smem &&
        fits((void*)train_fp32_kernel<BC, CP, true, kTT, TM, 2>, h2)) {
      t.wsmem = true; t.smem = place_boxes(h2);
      t.kernel = (void*)train_fp32_kernel<BC, CP, true, kTT, TM, 2>;
      t.h2 = true;
      t.wimg_bytes = ((size_t)BC * (KP0 + 8) + (size_t)(L - 1) * BC * (BC + 8)) * sizeof(float);
    }
  }
  // tcgen05 variant (MMA = 3): M = 64 chunk GEMMs with the accumulators in tensor memory; bc = 64, operand images of the
  // whole step resident in shared memory, gradient accumulators within the 512 TMEM columns.  Preferred over MMA = 2 when it
  // fits: it overrides the MMA = 2 selection above (LBDRN_TRAIN_H2=1 keeps the warp-level kernel for A/B).
  if constexpr (kMma && TM == 4 && BC == 64) {
    const int KP0 = round16(n.dim_in), L = n.nl;
    // (the feature staging buffer shares the bytes of the hidden-output / dz images: it must fit there)
    const size_t fl = (size_t)L * BC * kTrainLDP + (size_t)CP * kTrainLDP + (size_t)round4(L * BC + n.C * BC + n.C);
    const bool x_fits = (size_t)t.dimpad * kTrainLDP * 4 <= (size_t)L * 2 * (BC + 8) * kTrainNPIX * 2 + (size_t)L * 2 * BC * kTrainNPIX * 2;
    const size_t wimg = ((size_t)2 * KP0 * BC + (size_t)(L - 1) * 2 * BC * BC) * 2;
    const size_t imgs = wimg + (size_t)2 * (KP0 + 8) * kTrainNPIX * 2 + (size_t)L * 2 * (BC + 8) * kTrainNPIX * 2 +
                        (size_t)L * 2 * BC * kTrainNPIX * 2 + (size_t)2 * 16 * kTrainNPIX * 2;
    const size_t t5 = fl * sizeof(float) + 128 + imgs;
    const int tmem_cols = 2 * BC + 32 + (KP0 + 8) + (L - 1) * (BC + 8);
    if (mma && getenv("LBDRN_TRAIN_TF32") == nullptr && getenv("LBDRN_TRAIN_H2") == nullptr && tmem_cols <= 512 && x_fits &&
        n.C <= 8 && KP0 + 8 <= 256 && t5 <= (size_t)max_smem && fits((void*)train_fp32_kernel<BC, CP, true, kTT, TM, 3>, t5)) {
      t.wsmem = true; t.smem = t5;
      t.pf_off = 0; t.pf_stride = 0;     // (the optional TMA landing boxes belong to the warp-level kernels)
      t.kernel = (void*)train_fp32_kernel<BC, CP, true, kTT, TM, 3>;
      t.h2 = true;                       // same host-side handling: split weight image kept current by the Adam phase
      t.wimg_bytes = wimg;
    }
  }
  // streamed tcgen05 variant (MMA = 4) for bc 256: transposed GEMMs with M = 128 units, 32-pixel chunks, hidden weights
  // streamed from the global operand image (LBDRN_TRAIN_TF32=1 keeps the 3xTF32 warp-level kernel for A/B)
  if constexpr (kMma && TM == 2 && BC == 256) {
    const int KP0 = round16(n.dim_in), L = n.nl;
    const size_t xfl = (size_t)t.dimpad * kTrainLDP * 4 > 32768 ? (size_t)t.dimpad * kTrainLDP : 8192;
    const size_t fl = xfl + 2 * (size_t)CP * kTrainLDP + (size_t)round4(L * BC + n.C * BC + n.C);
    const size_t wimg = ((size_t)2 * KP0 * BC + (size_t)(L - 1) * 2 * BC * BC + (size_t)2 * 16 * BC) * 2;
    const size_t imgs = (size_t)2 * 16 * BC * 2 + (size_t)2 * (KP0 + 8) * kTrainNPIX * 2 + (size_t)L * 2 * (BC + 8) * kTrainNPIX * 2 +
                        (size_t)L * 2 * BC * kTrainNPIX * 2 + (size_t)2 * 16 * kTrainNPIX * 2;
    const size_t tx = fl * sizeof(float) + 128 + imgs;
    if (mma && getenv("LBDRN_TRAIN_TF32") == nullptr && KP0 + 8 <= 256 && n.C <= 8 && tx <= (size_t)max_smem &&
        fits((void*)train_fp32_kernel<BC, CP, false, kTT, TM, 4>, tx)) {
      t.wsmem = false; t.smem = tx;
      t.pf_off = 0; t.pf_stride = 0;
      t.kernel = (void*)train_fp32_kernel<BC, CP, false, kTT, TM, 4>;
      t.h2 = true; t.tcx = true;
      t.wimg_bytes = wimg;
    }
  }
  void* k_with = mma ? (void*)train_fp32_kernel<BC, CP, true, kTT, TM, kMma> : (void*)train_fp32_kernel<BC, CP, true, kTT, TM>;
  void* k_without = mma ? (void*)train_fp32_kernel<BC, CP, false, kTT, TM, kMma> : (void*)train_fp32_kernel<BC, CP, false, kTT, TM>;
  if (t.h2) {
  } else if (with_w <= (size_t)max_smem && fits(k_with, with_w)) {
    t.wsmem = true; t.smem = place_boxes(with_w);
    t.kernel = k_with;
  } else if (without <= (size_t)max_smem && fits(k_without, without)) {
    t.wsmem = false; t.smem = place_boxes(without);
    t.kernel = k_without;
  } else {
    return fail(LBDRN_E_UNSUPPORTED, "training working set (%zu B) exceeds shared memory for bc=%d nl=%d dim_in=%d",
                without, BC, n.nl, n.dim_in);
  }
  if (t.pf_off && !fits(t.kernel, t.smem)) {     // the optional TMA boxes pushed it over: drop them
    t.smem = (size_t)t.pf_off;
    t.pf_off = 0; t.pf_stride = 0;
  }
  CUDA_TRY(cudaFuncSetAttribute(t.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, t.kernel, kTT, t.smem));
  if (occ < 1) return fail(LBDRN_E_UNSUPPORTED, "training kernel cannot be made resident");
  const int cap = occ * sms;                  // cooperative launch: every CTA must be co-resident
  int want = (batch_size + kTrainNPIX - 1) / kTrainNPIX;
  if (want < 16) want = 16;
  t.grid = want < cap ? want : cap;
  return LBDRN_OK;
}

}  // namespace

int train_fp32_plan(const Net& n, int batch_size, int sms, int max_smem, TrainPlan& t) {
  t.dimpad = (n.dim_in + 7) & ~7;
  t.pstride = (round4(n.P + 1) + 28 + 31) & ~31;   // keep CTAs' partials on distinct 128 B lines
  const bool c4 = n.C <= 4;
  switch (n.bc) {
    case 32: return c4 ? plan_t<32, 4, 4>(n, batch_size, sms, max_smem, t) : plan_t<32, 8, 4>(n, batch_size, sms, max_smem, t);
    case 64: return c4 ? plan_t<64, 4, 4>(n, batch_size, sms, max_smem, t) : plan_t<64, 8, 4>(n, batch_size, sms, max_smem, t);
    case 128: return c4 ? plan_t<128, 4, 4>(n, batch_size, sms, max_smem, t) : plan_t<128, 8, 4>(n, batch_size, sms, max_smem, t);
    // bc = 256 (BASELINE config 3): activations of a 64-pixel chunk (339 KB at D=3, nl=2) exceed shared memory, so the
    // chunk is 32 pixels (TM = 2) and the weights stay in L2
    case 256: return c4 ? plan_t<256, 4, 2>(n, batch_size, sms, max_smem, t) : plan_t<256, 8, 2>(n, batch_size, sms, max_smem, t);
    default: return fail(LBDRN_E_UNSUPPORTED, "fused training is built for bc=32/64/128/256 (got %d)", n.bc);
  }
}

int train_fp32_launch(const TrainPlan& plan, TrainArgs& a, cudaStream_t st) {
  static long long* prof_dev = nullptr;
  const bool prof = getenv("LBDRN_TRAIN_PROF") != nullptr;
  if (prof) {
    if (!prof_dev) CUDA_TRY(cudaMalloc(&prof_dev, 24 * sizeof(long long)));
    CUDA_TRY(cudaMemsetAsync(prof_dev, 0, 24 * sizeof(long long), st));
    a.prof = prof_dev;
  }
  CUDA_TRY(cudaMemsetAsync(a.gbar, 0, 2 * sizeof(unsigned), st));
  void* kargs[] = {(void*)&a};
  CUDA_TRY(cudaLaunchCooperativeKernel(plan.kernel, dim3(plan.grid), dim3(kTT), kargs, plan.smem, st));
  ++g_launches;
  if (prof) {
    long long h[24];
    CUDA_TRY(cudaMemcpyAsync(h, prof_dev, sizeof h, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    static const char* names[24] = {"reload", "gather", "fwd_barriers", "out+loss", "bwd_rest", "wait1", "reduce+adam", "wait2",
                                    "bwd_out", "bwd_dh0", "bwd_dW0", "bwd_dh1", "bwd_dW1", "pf_commit", "pf_issue", "fwd_gemm",
                                    "fwd_act", "red_batches", "red_tail", "red_sync", "arrive2", "head", "g_coords", "g_rows"};
    fprintf(stderr, "[lbdrn] train phases (cycles/step on CTA 0, %d steps, grid %d x %d thr):", a.n_steps, plan.grid, kTT);
    for (int i = 0; i < 24; ++i) fprintf(stderr, " %s=%lld", names[i], h[i] / (a.n_steps > 0 ? a.n_steps : 1));
    fprintf(stderr, "\n");
  }
  return LBDRN_OK;
}

// CHW uint8 planes -> one 32-bit word per pixel (band c in byte c, unused bytes zero).  HBM-bound: C + 4 bytes per pixel.
// VEC: 4 pixels per thread (one 32-bit load per plane, one 16-byte store) when the plane length is a multiple of 4.
template <bool VEC>
static __global__ void interleave_u8_kernel(const uint8_t* __restrict__ planes, int C, size_t npix, uint32_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (VEC) {
    const size_t n4 = npix >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      for (int c = 0; c < C; ++c) w[c] = reinterpret_cast<const uint32_t*>(planes + (size_t)c * npix)[i];
      const uint32_t t01 = __byte_perm(w[0], w[1], 0x5140), t23 = __byte_perm(w[2], w[3], 0x5140);   // pixels 0, 1: bands (0,1) / (2,3)
      const uint32_t u01 = __byte_perm(w[0], w[1], 0x7362), u23 = __byte_perm(w[2], w[3], 0x7362);   // pixels 2, 3
      uint4 o;
      o.x = __byte_perm(t01, t23, 0x5410);
      o.y = __byte_perm(t01, t23, 0x7632);
      o.z = __byte_perm(u01, u23, 0x5410);
      o.w = __byte_perm(u01, u23, 0x7632);
      reinterpret_cast<uint4*>(out)[i] = o;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
      uint32_t w = 0u;
      for (int c = 0; c < C; ++c) w |= (uint32_t)planes[(size_t)c * npix + i] << (8 * c);
      out[i] = w;
    }
  }
}

// Global operand image of the streamed tcgen05 variant from the master parameters (before every launch; the Adam phase keeps
// it current inside a launch): per hidden layer [hi Kp*BC | lo Kp*BC] halves with rows = BC, then W_o as a 16-row image.
static __global__ void tcx_wimg_kernel(Net net, const float* __restrict__ params, uint16_t* __restrict__ wimg) {
  const int BC = net.bc, KP0 = round16(net.dim_in), L = net.nl;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < net.P; i += gridDim.x * blockDim.x) {
    uint32_t img = 0u;
    bool done = false;
    for (int l = 0; l < L && !done; ++l) {
      const int K = l == 0 ? net.dim_in : BC, Kp = l == 0 ? KP0 : BC, o = i - net.woff[l];
      if (o >= 0 && o < K * BC) {
        const int u = o / K, k = o - u * K;
        uint16_t hi, lo;
        split_h1(params[i] * kWScale, hi, lo);
        const uint32_t e = img + (uint32_t)(img_off(BC, u, k) >> 1);
        wimg[e] = hi;
        wimg[e + (uint32_t)Kp * BC] = lo;
        done = true;
      }
      img += 2u * (uint32_t)Kp * BC;
    }
    const int oo = i - net.woff[L];
    if (!done && oo >= 0 && oo < net.C * BC) {
      const int c = oo / BC, u = oo - c * BC;
      uint16_t hi, lo;
      split_h1(params[i] * kWScale, hi, lo);
      const uint32_t e = img + (uint32_t)(img_off(16, c, u) >> 1);
      wimg[e] = hi;
      wimg[e + 16u * BC] = lo;
    }
  }
}

void launch_tcx_wimg(const Net& n, const float* params, uint16_t* wimg, cudaStream_t st) {
  tcx_wimg_kernel<<<(n.P + 255) / 256, 256, 0, st>>>(n, params, wimg);
  ++g_launches;
}

void launch_interleave_u8(const void* planes, int C, size_t npix, uint32_t* out, int sms, cudaStream_t st) {
  const bool vec = npix % 4 == 0 && reinterpret_cast<uintptr_t>(planes) % 4 == 0;
  if (vec) interleave_u8_kernel<true><<<sms * 8, 256, 0, st>>>((const uint8_t*)planes, C, npix, out);
  else interleave_u8_kernel<false><<<sms * 8, 256, 0, st>>>((const uint8_t*)planes, C, npix, out);
  ++g_launches;
}

void launch_adam_apply(const Net& n, const float* grad, float* params, float* wpack, float* m, float* v, float omb1,
                       float omb2, float beta2, float eps, float step_size, float bc2_sqrt, cudaStream_t st) {
  adam_apply_kernel<<<(n.P + 255) / 256, 256, 0, st>>>(n, grad, params, wpack, m, v, omb1, omb2, beta2, eps, step_size,
                                                      bc2_sqrt);
  ++g_launches;
}

}  // namespace lbdrn
