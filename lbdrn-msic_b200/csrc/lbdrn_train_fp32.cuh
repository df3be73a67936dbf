// lbdrn_train_fp32.cuh -- fused, persistent training kernel (fp32 master weights and accumulation): neighbourhood gather
// by pixel index, forward, LBDRNLoss (MSE), backward, deterministic cross-CTA gradient reduction and Adam, for many
// consecutive optimiser steps in ONE cooperative launch.  Replaces the per-step ~35-40 library launches + H2D + host
// sync of the reference (modified_ignite_engine.py:18-27, encode.py:69-70,84-85,95).
//
// Per step:   [each CTA]  64-pixel chunks of the batch: (prefetched) gather -> fwd -> bwd -> gradient partial in L2
//             arrive 1 ; issue the next step's neighbourhood loads ; wait 1            (split grid barrier)
//             [all threads] fixed-order sum of the partials, Adam update of (param, m, v), weight image refresh
//             arrive 2 ; next coordinates, commit of the prefetched neighbourhoods ; wait 2 ; weights -> smem (cp.async)
// The chunk GEMMs run on the tensor cores through warp-level MMAs (template parameter MMA: 2 = fp16 hi+lo split operands
// read with ldmatrix, 1 = 3xTF32, 0 = FFMA).  Latency-bound by construction (81 920 dependent steps for an 8192^2
// scene): the budget is the two grid barriers, the L2 round trips of the reduction and of the weight reload, not flops.
#pragma once
#include <cooperative_groups.h>
#include <type_traits>

#include "lbdrn_common.cuh"
#include "lbdrn_umma.cuh"

namespace lbdrn {
namespace cg = cooperative_groups;

// Chunk geometry: TM pixels per thread group, 16 groups -> NPIX = 16*TM pixels per chunk; rows of the activation
// buffers are padded to LDP = NPIX + 4 floats (conflict-free LDS.128 down a column of rows).  TM = 4 (64-pixel chunks)
// wherever the working set fits in shared memory; TM = 2 (32-pixel chunks) for bc = 256.
constexpr int train_npix(int tm) { return 16 * tm; }
constexpr int train_ldp(int tm) { return 16 * tm + 4; }

enum { TRAIN_FUSED = 0, TRAIN_GRAD_ONLY = 1 };

struct TrainArgs {
  Net net;
  const void* msb;
  const void* lsb;
  const float* tab;
  const int64_t* perm;     // pixel indices; step s uses [s*bs, min((s+1)*bs, n_perm))
  long long n_perm;
  int bs, n_steps, mode;
  long long adam_t0;
  double lr, beta1, beta2;
  float omb1, omb2, beta2f, eps;   // (float)(1-beta1), (float)(1-beta2), (float)beta2, (float)eps
  int n_global;            // GRAD_ONLY: global batch size for the 2/(B*C) factor
  float* params;           // master parameters, reference layout
  float* wpack;            // packed copy (hidden W transposed) kept in sync by the Adam phase (MMA = 0 / 1 kernels)
  float* m;
  float* v;
  float* partial;          // [gridDim.x][pstride]; slot P holds the chunk squared-error sum
  int pstride;
  float* losses;           // FUSED: [n_steps]
  float* grad_out;         // GRAD_ONLY: [P+1]
  int dimpad;              // dim_in rounded up to 8 (rows of the X buffer, padding rows are zero)
  long long* prof;         // optional (LBDRN_TRAIN_PROF=1): clock64 cycles per phase accumulated by CTA 0 / thread 0
  const CUtensorMap* tmap_msb;   // TMA neighbourhood gather: 3-D maps of the MSB / LSB planes (boxes 32x5xC / 16x1xC bytes);
  const CUtensorMap* tmap_lsb;   // nullptr: register prefetch only
  int pf_off, pf_stride;   // byte offset of the landing boxes inside dynamic shared memory, bytes per pixel
  uint16_t* wimg;          // fp16-split kernel: global image of the shared-memory weight operands (hi | lo halves per layer,
                           // padding zero), kept current by the Adam phase so that a reload is one asynchronous copy
  const uint32_t* imsb;    // band-interleaved copies of the uint8 planes (one 32-bit word per pixel, band c in byte c), built
  const uint32_t* ilsb;    // by the library before each launch; nullptr: gather from the CHW planes
  int l2_hints;            // planes larger than L2: prefetch.global.L2 hints for the next step's neighbourhoods
  unsigned* gbar;          // two monotonic arrival counters of the split grid barriers (zeroed before every launch)
  int fast_sine;           // tcgen05 kernel: hidden sine / cosine through MUFU (A/B switch, off by default)
  const float2* adam_tab;  // FUSED: per step {lr / (1 - beta1^t), sqrt(1 - beta2^t)}, formed on the host in double like
                           // torch's python floats (two pow() per step in fp64 cost ~3k cycles of every step on the device)
};

template <int NPIX>
__device__ __forceinline__ float row_sum(const float* __restrict__ row) {
  float s = 0.f;
#pragma unroll
  for (int p = 0; p < NPIX; p += 4) {
    float4 v = *reinterpret_cast<const float4*>(row + p);
    s += (v.x + v.y) + (v.z + v.w);
  }
  return s;
}

template <int NPIX>
__device__ __forceinline__ float row_dot(const float* __restrict__ a, const float* __restrict__ b) {
  float s0 = 0.f, s1 = 0.f;          // two chains: the 64-term dot product was one dependent FFMA chain
#pragma unroll
  for (int p = 0; p < NPIX; p += 8) {
    float4 x = *reinterpret_cast<const float4*>(a + p), y = *reinterpret_cast<const float4*>(b + p);
    float4 x2 = *reinterpret_cast<const float4*>(a + p + 4), y2 = *reinterpret_cast<const float4*>(b + p + 4);
    s0 = fmaf(x.x, y.x, s0); s0 = fmaf(x.y, y.y, s0); s0 = fmaf(x.z, y.z, s0); s0 = fmaf(x.w, y.w, s0);
    s1 = fmaf(x2.x, y2.x, s1); s1 = fmaf(x2.y, y2.y, s1); s1 = fmaf(x2.z, y2.z, s1); s1 = fmaf(x2.w, y2.w, s1);
  }
  return s0 + s1;
}

// index of parameter i inside the packed copy (hidden weights transposed)
__device__ __forceinline__ int packed_index(const Net& net, int i) {
  for (int l = 0; l < net.nl; ++l) {
    int K = l == 0 ? net.dim_in : net.bc;
    int o = i - net.woff[l];
    if (o >= 0 && o < K * net.bc) {
      int n = o / K, k = o - n * K;
      return net.woff[l] + k * net.bc + n;
    }
  }
  return i;
}

// torch.optim.Adam single-tensor update (lr, betas, eps; no weight decay / amsgrad), torch/optim/adam.py
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, float omb1, float omb2,
                                            float beta2, float eps, float step_size, float bc2_sqrt) {
  m = fmaf(omb1, g - m, m);                              // exp_avg.lerp_(grad, 1-beta1)
  v = fmaf(omb2, g * g, v * beta2);                      // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
  float denom = __fdiv_rn(__fsqrt_rn(v), bc2_sqrt) + eps;
  p = p - step_size * __fdiv_rn(m, denom);               // param.addcdiv_(exp_avg, denom, value=-step_size)
}

// load VEC consecutive floats
template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float* out) {
  if (VEC == 4) { float4 v = *reinterpret_cast<const float4*>(p); out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
  else if (VEC == 2) { float2 v = *reinterpret_cast<const float2*>(p); out[0] = v.x; out[1] = v.y; }
  else out[0] = *p;
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float* v) {
  if (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else if (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  else *p = v[0];
}

// acc[i][j] += sum_k act[k*LDP + pg*4 + i] * wt[k*BC + ubase + (j/VEC)*8*VEC + j%VEC]   (both operands k-major)
template <int TM, int TN, int VEC, int BC>
__device__ __forceinline__ void gemm_kmajor_v(float (&acc)[TM][TN], const float* __restrict__ act, const float* wt,
                                              int K, int pg, int ubase) {
  constexpr int kTrainLDP = train_ldp(TM);
  const float* a = act + pg * TM;
  const float* b = wt + ubase;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float av[TM], bv[TN];
    load_vec<TM>(a + (size_t)k * kTrainLDP, av);
#pragma unroll
    for (int j = 0; j < TN; j += VEC) load_vec<VEC>(b + (size_t)k * BC + (j / VEC) * 8 * VEC, bv + j);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// ---- warp-level tensor-core pieces for the 64-pixel chunk: mma.sync m16n8k8 TF32 with the 3xTF32 split ------------------
// A training step is the latency of ONE 64-pixel chunk on one SM, with operands (weights, activations, gradients) that
// change every step and contraction sizes of 64-104: smaller than one tcgen05 tile (M = 128, operands staged in shared
// memory through descriptors, accumulators in TMEM), so the register-fragment MMA is the right-sized instruction here.
// fp32 accuracy is kept by splitting both operands, x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and accumulating
// lo*hi + hi*lo + hi*hi in fp32 (relative error ~2^-21 per product): one MMA instruction replaces 32 FFMA instructions.
// 16-byte asynchronous global -> shared copy through L2 only (.cg: the weights were just rewritten by other SMs).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Split grid barrier on a monotonic counter: arrive (ONE thread, after a __syncthreads() that orders the CTA's writes before
// its fence -- the same cumulativity cooperative_groups' grid.sync() relies on) ... independent work ... wait (one thread
// spins, then a __syncthreads() releases the CTA).  Unlike grid.sync() the two halves can bracket work that does not depend
// on the other CTAs.  Needs all CTAs co-resident: the kernel is still launched cooperatively.
__device__ __forceinline__ void gbar_arrive(unsigned* bar) {
  __threadfence();
  atomicAdd(bar, 1u);
}
__device__ __forceinline__ void gbar_wait(const unsigned* bar, unsigned target) {
  unsigned v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
  } while (v < target);
}

// cvt.rna.tf32.f32 is not a hardware instruction on sm_100a: ptxas expands it to an Inf/NaN test, an integer add, a select
// and a mask (~9 instructions per split element, which made the split -- not the MMAs -- the cost of the chunk GEMMs).
// Same rounding by hand for finite inputs: add half an ulp of the 10-bit mantissa to the magnitude bits (round to nearest,
// ties away from zero) and clear the 13 low bits; lo = x - hi is exact and is truncated to tf32 (|x - hi - lo| <= 2^-21 |x|).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(d, al, bh);      // small terms first
  mma_tf32(d, ah, bl);
  mma_tf32(d, ah, bh);
}

// Fragment coordinates of mma.m16n8k8 (g = lane / 4, t = lane % 4):
//   A (16x8, row): a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
//   B (8x8,  col): b0 (k = t, n = g)  b1 (k = t+4, n = g)
//   C (16x8)     : c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
//
// acc[j][.] += sum_k act[k][p] * wt[k*BC + u] for the warp's 16-pixel tile (m) and its NT unit tiles (n): with PT pixel
// tiles per chunk (4: 64 pixels, 2: 32 pixels) and NG = 16 / PT unit groups, warp w owns pixels 16*(w % PT) .. +15 and unit
// tiles u0 = 8*((w / PT) + NG*j).  act rows are pixel-contiguous with stride LDP (A loads conflict-free: bank = g + 4t for
// LDP = 4 mod 32); act has K rounded up to 8 rows (rows >= K are zero); wt rows >= K are not read.
template <int BC, int LDP, int NT, int PT>
__device__ __forceinline__ void gemm_px_unit_mma(float (&acc)[NT][4], const float* __restrict__ act, const float* wt, int K,
                                                 int warp, int lane) {
  constexpr int NG = 16 / PT;
  static_assert(NT * NG * 8 == BC, "unit tiles must cover the layer");
  const int K8 = (K + 7) & ~7;
  const int g = lane >> 2, t = lane & 3, p0 = 16 * (warp % PT), ng = warp / PT;
  const float* a_ptr = act + (size_t)t * LDP + p0 + g;
  const float* b_ptr = wt + (size_t)t * BC + 8 * ng + g;
#pragma unroll 2
  for (int k0 = 0; k0 < K8; k0 += 8) {
    uint32_t ah[4], al[4];
    split_tf32(a_ptr[(size_t)k0 * LDP], ah[0], al[0]);
    split_tf32(a_ptr[(size_t)k0 * LDP + 8], ah[1], al[1]);
    split_tf32(a_ptr[(size_t)(k0 + 4) * LDP], ah[2], al[2]);
    split_tf32(a_ptr[(size_t)(k0 + 4) * LDP + 8], ah[3], al[3]);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      uint32_t bh[2], bl[2];
      split_tf32(k0 + t < K ? b_ptr[(size_t)k0 * BC + 8 * NG * j] : 0.f, bh[0], bl[0]);
      split_tf32(k0 + t + 4 < K ? b_ptr[(size_t)(k0 + 4) * BC + 8 * NG * j] : 0.f, bh[1], bl[1]);
      mma_3xtf32(acc[j], ah, al, bh, bl);
    }
  }
}

// dst[r*Kin + q] (+)= sum_p G[r][p] * A[q][p] (weight gradient block, natural [BC][Kin] layout) over the chunk's 64 pixels.
// m = rows r of G (BC/16 tiles), n = rows q of A (ceil(Kin/8) tiles), k = pixels.  Warp w owns m-tile w % MT and the n-tiles
// (w / MT) + (16 / MT) * j.  Both operands are pixel-contiguous with stride LDP: conflict-free fragment loads.
template <int BC, int LDP, int NPIX>
__device__ __forceinline__ void grad_weight_mma(const float* __restrict__ G, const float* __restrict__ A, int Kin, int kpad8,
                                                float* __restrict__ dst, bool first, int warp, int lane) {
  constexpr int MT = BC / 16, NW = 16 / MT, MAXN = 4;          // at most 4 n-tiles per warp per pass
  constexpr int KST = NPIX / 8;                                // k steps (8 pixels each)
  const int g = lane >> 2, t = lane & 3, r0 = 16 * (warp % MT), nq = kpad8 >> 3;
  const float* a_ptr = G + (size_t)(r0 + g) * LDP + t;
  // 32-pixel chunks (bc 256: every warp walks ~57 column groups with the same 16 rows of dz): the split A fragments of all
  // k steps stay in registers; 64-pixel chunks re-load them per pass (64 registers would not fit)
  constexpr bool HOIST = KST <= 4;
  uint32_t ahh[HOIST ? KST : 1][4], all_[HOIST ? KST : 1][4];
  if (HOIST) {
#pragma unroll
    for (int ks = 0; ks < KST; ++ks) {
      const int p0 = 8 * ks;
      split_tf32(a_ptr[p0], ahh[ks][0], all_[ks][0]);
      split_tf32(a_ptr[(size_t)8 * LDP + p0], ahh[ks][1], all_[ks][1]);
      split_tf32(a_ptr[p0 + 4], ahh[ks][2], all_[ks][2]);
      split_tf32(a_ptr[(size_t)8 * LDP + p0 + 4], ahh[ks][3], all_[ks][3]);
    }
  }
  // c0 | c1 and c2 | c3 are neighbours in a row of the gradient block: 8-byte stores when the row pitch allows
  const bool pair_ok = (Kin & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
  for (int nbase = warp / MT; nbase < nq; nbase += NW * MAXN) {
    float acc[MAXN][4];
#pragma unroll
    for (int j = 0; j < MAXN; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KST; ++ks) {
      const int p0 = 8 * ks;
      uint32_t ah[4], al[4];
      if (HOIST) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { ah[i] = ahh[ks][i]; al[i] = all_[ks][i]; }
      } else {
        split_tf32(a_ptr[p0], ah[0], al[0]);
        split_tf32(a_ptr[(size_t)8 * LDP + p0], ah[1], al[1]);
        split_tf32(a_ptr[p0 + 4], ah[2], al[2]);
        split_tf32(a_ptr[(size_t)8 * LDP + p0 + 4], ah[3], al[3]);
      }
#pragma unroll
      for (int j = 0; j < MAXN; ++j) {
        const int nt = nbase + NW * j;
        if (nt < nq) {                                           // warp-uniform
          const float* b_ptr = A + (size_t)(8 * nt + g) * LDP + p0 + t;
          uint32_t bh[2], bl[2];
          split_tf32(b_ptr[0], bh[0], bl[0]);
          split_tf32(b_ptr[4], bh[1], bl[1]);
          mma_3xtf32(acc[j], ah, al, bh, bl);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < MAXN; ++j) {
      const int nt = nbase + NW * j;
      if (nt < nq) {
        if (pair_ok) {
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int r = r0 + g + e2 * 8, q = 8 * nt + 2 * t;
            if (q < Kin) {
              float2* d = reinterpret_cast<float2*>(dst + (size_t)r * Kin + q);
              float2 v = make_float2(acc[j][2 * e2], acc[j][2 * e2 + 1]);
              if (!first) { const float2 o = *d; v.x += o.x; v.y += o.y; }
              *d = v;
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = r0 + g + (e >> 1) * 8, q = 8 * nt + 2 * t + (e & 1);
            if (q < Kin) {
              float* d = dst + (size_t)r * Kin + q;
              *d = first ? acc[j][e] : *d + acc[j][e];
            }
          }
        }
      }
    }
  }
}

// ---- fp16 hi+lo split operands fed by ldmatrix (template parameter MMA == 2) -----------------------------------------------
// Measured on B200 (tools/probe/mma_rate.cu): every warp-level MMA shape issues once per 8 cycles per sub-partition, so
// m16n8k16 on fp16 does twice the work of m16n8k8 on tf32 per issue slot -- and the 3xTF32 loop above spends ~8 instructions
// per MMA loading fp32 fragments word by word and splitting them again in each of the 4 warps that share them.  Here every
// operand is split ONCE by its producer, x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (|x - hi - lo| <= 2^-22 |x| while
// lo is a normal fp16 number, <= 2^-25 absolute below that), and stored as two 16-bit arrays [feature or unit][pixel] with
// 144-byte rows; one ldmatrix.x4 (.trans where the contraction runs down the rows) delivers a whole 16x16 fragment, and the
// product is lo*hi + hi*lo + hi*hi accumulated in fp32: 10 instructions per 6 MMAs of k = 16 in the forward loop.
// Ranges: features and sine outputs are in [-1, 1] and go in unscaled; weights are scaled by 2^8 (fp16 overflows at
// |W| >= 256 -- a SIREN weight of that size means sin(7680 x), the loss would be NaN and visible); back-propagated dz is
// scaled per chunk and layer by the power of two that puts its largest magnitude just under 2^15.
constexpr int kLDH = 72;                      // halves per row (64 pixels + 8): the 8 row addresses of an 8x8 tile hit 8 different
                                              // 16-byte bank groups (144 = 9 * 16)
constexpr float kWScale = 256.f, kWInv = 1.f / 256.f;
__host__ __device__ constexpr int round16(int x) { return (x + 15) & ~15; }

__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split_h1(float x, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn(x - __half2float(h));
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_h2x3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                         uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_f16(d, al, bh0, bh1);      // small terms first
  mma_f16(d, ah, bl0, bl1);
  mma_f16(d, ah, bh0, bh1);
}

// Fragments of mma.m16n8k16 (g = lane / 4, t = lane % 4), each register a pair of halves:
//   A (16x16, row): a0 (g, k 2t..2t+1)  a1 (g+8, same k)  a2 (g, k+8)  a3 (g+8, k+8)
//   B (16x8,  col): b0 (k 2t..2t+1, n g)  b1 (k+8, n g)        C as m16n8k8.
// acc[j] += sum_k act[k][p] * B[k][u]: warp w owns pixels 16 (w & 3) .. +15 and the unit tiles u0 = 8 ((w >> 2) + 4 j).
// act_*: shared addresses of the [k][kLDH] arrays (pixel-contiguous, so the A tiles are read with .trans).
// WT_ROWS_ARE_K false: w_* is W[u][k] (natural weight of the forward pass, rows = output units): B tiles read as stored.
// WT_ROWS_ARE_K true : w_* is W[k][u] (natural weight of the NEXT layer in dh = W^T dz): B tiles read with .trans.
template <int NT, bool WT_ROWS_ARE_K>
__device__ __forceinline__ void gemm_px_unit_h2(float (&acc)[NT][4], uint32_t act_hi, uint32_t act_lo, uint32_t w_hi,
                                                uint32_t w_lo, int ldw, int ksteps, int warp, int lane) {
  static_assert(NT % 2 == 0, "unit tiles are loaded in pairs");
  const int p0 = 16 * (warp & 3), ng = warp >> 2;
  // A tiles: [k0, p0] [k0, p0+8] [k0+8, p0] [k0+8, p0+8]
  const uint32_t a_off = (uint32_t)(((lane & 7) + 8 * ((lane >> 4) & 1)) * kLDH + p0 + 8 * ((lane >> 3) & 1)) * 2u;
  uint32_t b_off[NT / 2];
#pragma unroll
  for (int jj = 0; jj < NT / 2; ++jj) {
    const int u0 = 8 * (ng + 4 * (2 * jj + (lane >> 4)));                 // unit tile of lanes 0-15 / 16-31
    b_off[jj] = WT_ROWS_ARE_K ? (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * ldw + u0) * 2u      // [k0,u][k0+8,u] x 2 tiles
                              : (uint32_t)((u0 + (lane & 7)) * ldw + 8 * ((lane >> 3) & 1)) * 2u;    // [u,k0][u,k0+8] x 2 tiles
  }
  const uint32_t b_step = WT_ROWS_ARE_K ? (uint32_t)(16 * ldw) * 2u : 32u;
#pragma unroll 2
  for (int ks = 0; ks < ksteps; ++ks) {
    uint32_t ah[4], al[4];
    ldsm_x4_t(ah, act_hi + a_off + (uint32_t)ks * (16u * kLDH * 2u));
    ldsm_x4_t(al, act_lo + a_off + (uint32_t)ks * (16u * kLDH * 2u));
#pragma unroll
    for (int jj = 0; jj < NT / 2; ++jj) {
      uint32_t bh[4], bl[4];
      if (WT_ROWS_ARE_K) {
        ldsm_x4_t(bh, w_hi + b_off[jj] + (uint32_t)ks * b_step);
        ldsm_x4_t(bl, w_lo + b_off[jj] + (uint32_t)ks * b_step);
      } else {
        ldsm_x4(bh, w_hi + b_off[jj] + (uint32_t)ks * b_step);
        ldsm_x4(bl, w_lo + b_off[jj] + (uint32_t)ks * b_step);
      }
      mma_h2x3(acc[2 * jj], ah, al, bh[0], bh[1], bl[0], bl[1]);
      mma_h2x3(acc[2 * jj + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
    }
  }
}

// dst[r*Kin + q] (+)= inv * sum_p G[r][p] * IN[q][p] over the chunk's 64 pixels (weight-gradient block, natural [BC][Kin]).
// g_*: [BC][kLDH] arrays of the scaled dz, in_*: [8 nq][kLDH] arrays of the layer input; both pixel-contiguous = K-contiguous,
// so every tile is read as stored.  Warp w owns row tile w % MT and the column tiles (w / MT) + (16 / MT) j.
template <int BC>
__device__ __forceinline__ void grad_weight_h2(uint32_t g_hi, uint32_t g_lo, uint32_t in_hi, uint32_t in_lo, int Kin, int nq,
                                               float inv, float* __restrict__ dst, bool first, int warp, int lane) {
  constexpr int MT = BC / 16, NW = 16 / MT, MAXN = 4;
  const int g = lane >> 2, t = lane & 3, r0 = 16 * (warp % MT);
  // A tiles: [r0, p0] [r0+8, p0] [r0, p0+8] [r0+8, p0+8]
  const uint32_t a_off = (uint32_t)((r0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kLDH + 8 * (lane >> 4)) * 2u;
  const bool pair_ok = (Kin & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
  for (int nbase = warp / MT; nbase < nq; nbase += NW * MAXN) {
    float acc[MAXN][4];
#pragma unroll
    for (int j = 0; j < MAXN; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
    uint32_t b_off[MAXN / 2];
#pragma unroll
    for (int jj = 0; jj < MAXN / 2; ++jj) {
      int nt = nbase + NW * (2 * jj + (lane >> 4));
      nt = nt < nq ? nt : nq - 1;                                   // keep the address inside the array; the MMA is skipped
      b_off[jj] = (uint32_t)((8 * nt + (lane & 7)) * kLDH + 8 * ((lane >> 3) & 1)) * 2u;    // [q, p0] [q, p0+8] x 2 tiles
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {                                // 64 pixels
      uint32_t ah[4], al[4];
      ldsm_x4(ah, g_hi + a_off + ks * 32u);
      ldsm_x4(al, g_lo + a_off + ks * 32u);
#pragma unroll
      for (int jj = 0; jj < MAXN / 2; ++jj) {
        if (nbase + NW * 2 * jj < nq) {                             // warp-uniform
          uint32_t bh[4], bl[4];
          ldsm_x4(bh, in_hi + b_off[jj] + ks * 32u);
          ldsm_x4(bl, in_lo + b_off[jj] + ks * 32u);
          mma_h2x3(acc[2 * jj], ah, al, bh[0], bh[1], bl[0], bl[1]);
          if (nbase + NW * (2 * jj + 1) < nq) mma_h2x3(acc[2 * jj + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < MAXN; ++j) {
      const int nt = nbase + NW * j;
      if (nt < nq) {
        if (pair_ok) {
          // c0 | c1 and c2 | c3 are neighbours in a row of the gradient block: 8-byte stores, each warp store fills 8 sectors
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int r = r0 + g + e2 * 8, q = 8 * nt + 2 * t;
            if (q < Kin) {
              float2* d = reinterpret_cast<float2*>(dst + (size_t)r * Kin + q);
              float2 v = make_float2(acc[j][2 * e2] * inv, acc[j][2 * e2 + 1] * inv);
              if (!first) { const float2 o = *d; v.x += o.x; v.y += o.y; }
              *d = v;
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = r0 + g + (e >> 1) * 8, q = 8 * nt + 2 * t + (e & 1);
            if (q < Kin) {
              float* d = dst + (size_t)r * Kin + q;
              const float v = acc[j][e] * inv;
              *d = first ? v : *d + v;
            }
          }
        }
      }
    }
  }
}

// power of two S with max * S in [2^14, 2^15) (S = 1 for max = 0), and its exact inverse
__device__ __forceinline__ void dz_scale(uint32_t max_bits, float& S, float& inv) {
  int f = 268 - (int)(max_bits >> 23);                             // biased exponent field of S
  f = max_bits == 0u ? 127 : (f > 250 ? 250 : (f < 4 ? 4 : f));
  S = __uint_as_float((uint32_t)f << 23);
  inv = __uint_as_float((uint32_t)(254 - f) << 23);
}

// One row (fixed dy) of one band's neighbourhood of one pixel; loads issued before first use.
template <int N_, int kTrainLDP>
__device__ __forceinline__ void gather_row(const void* msb, int u16, size_t rowoff, int gx, const Net& net, float maxv,
                                           float ctr, bool ok, float* d) {
  constexpr int D_ = N_ / 2;
  uint32_t raw[N_];
#pragma unroll
  for (int dx = 0; dx < N_; ++dx) raw[dx] = load_msb_int(msb, u16, rowoff + reflect_clamp(gx + dx - D_, net.W));
#pragma unroll
  for (int dx = 0; dx < N_; ++dx) d[(size_t)dx * kTrainLDP] = ok ? __fdiv_rn((float)raw[dx], maxv) - ctr : 0.f;
}

// The same row read as aligned 8-byte words (two for 8-bit planes, three for 16-bit ones) and shifted into place.
// Rows that touch the left/right border (reflection) or the ends of the buffer take the element-wise path.
template <int N_, int kTrainLDP>
__device__ __forceinline__ void gather_row_wide(const void* msb, size_t total_bytes, int u16, size_t rowoff, int gx,
                                                const Net& net, float maxv, const float* quot, float ctr, bool ok,
                                                float* d) {
  constexpr int D_ = N_ / 2;
  static_assert(N_ <= 8, "a row must fit the shifted words");
  const uintptr_t base = (uintptr_t)msb;
  const uintptr_t p = base + ((rowoff + (size_t)(gx - D_)) << u16);
  const uintptr_t a0 = p & ~(uintptr_t)7;
  const int sh = (int)(p & 7) * 8;
  const bool fast = gx >= D_ && gx + D_ < net.W && a0 >= base && a0 + (u16 ? 24 : 16) <= base + total_bytes;
  uint32_t raw[N_];
  if (fast) {
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(a0);
    const unsigned long long q0 = __ldg(q), q1 = __ldg(q + 1);
    const unsigned long long r0 = sh ? (q0 >> sh) | (q1 << (64 - sh)) : q0;
    if (!u16) {
#pragma unroll
      for (int dx = 0; dx < N_; ++dx) raw[dx] = (uint32_t)(r0 >> (8 * dx)) & 0xffu;
    } else {
      const unsigned long long q2 = __ldg(q + 2);
      const unsigned long long r1 = sh ? (q1 >> sh) | (q2 << (64 - sh)) : q1;
#pragma unroll
      for (int dx = 0; dx < N_; ++dx)
        raw[dx] = (uint32_t)((dx < 4 ? r0 : r1) >> (16 * (dx & 3))) & 0xffffu;
    }
  } else {
#pragma unroll
    for (int dx = 0; dx < N_; ++dx) raw[dx] = load_msb_int(msb, u16, rowoff + reflect_clamp(gx + dx - D_, net.W));
  }
  // 8-bit planes: the quotient comes from the 256-entry table (49 divisions per pixel and band were the gather's cost)
  if (!u16) {
#pragma unroll
    for (int dx = 0; dx < N_; ++dx) d[(size_t)dx * kTrainLDP] = ok ? quot[raw[dx]] - ctr : 0.f;
  } else {
#pragma unroll
    for (int dx = 0; dx < N_; ++dx) d[(size_t)dx * kTrainLDP] = ok ? __fdiv_rn((float)raw[dx], maxv) - ctr : 0.f;
  }
}

// out[r][q] (+)= sum_p G[r][p] * A[q][p] for r < BC, q < Kin (dst = natural [BC][Kin] gradient block).
// Thread (trow = tid%16, tcol = tid/16) owns rows trow + 16x (x<4) and columns tcol + CT*y (y<4, CT = THREADS/16):
// a 4x4 register tile with 16 independent FFMA chains; per 4-pixel step 4 + (#valid y) LDS.128 feed 16*(#valid y)*4/4 FFMAs.
// Row loads of a quarter-warp hit 8 consecutive rows (stride 68 floats = 4 banks apart: conflict-free), column loads are
// broadcasts.  64 rows x 4*CT columns per pass.
template <int BC, int THREADS, int TM>
__device__ __forceinline__ void grad_weight_nt_t(const float* __restrict__ G, const float* __restrict__ A, int Kin,
                                                 int kpad8, float* __restrict__ dst, bool first) {
  constexpr int LDP = train_ldp(TM), kTrainNPIX = train_npix(TM), CT = THREADS / 16;
  const int tid = threadIdx.x, trow = tid & 15, tcol = tid >> 4;
  for (int r0 = 0; r0 < BC; r0 += 64) {
    for (int q0 = 0; q0 < Kin; q0 += 4 * CT) {
      float acc[4][4];
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
      bool qok[4];
#pragma unroll
      for (int y = 0; y < 4; ++y) qok[y] = q0 + tcol + CT * y < kpad8;
      if (qok[0]) {
#pragma unroll 2
        for (int p = 0; p < kTrainNPIX; p += 4) {
          float4 gv[4], av[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int r = r0 + trow + 16 * x;
            gv[x] = r < BC ? *reinterpret_cast<const float4*>(G + (size_t)r * LDP + p) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int y = 0; y < 4; ++y)
            av[y] = qok[y] ? *reinterpret_cast<const float4*>(A + (size_t)(q0 + tcol + CT * y) * LDP + p)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              acc[x][y] = fmaf(gv[x].x, av[y].x, acc[x][y]);
              acc[x][y] = fmaf(gv[x].y, av[y].y, acc[x][y]);
              acc[x][y] = fmaf(gv[x].z, av[y].z, acc[x][y]);
              acc[x][y] = fmaf(gv[x].w, av[y].w, acc[x][y]);
            }
        }
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const int q = q0 + tcol + CT * y, r = r0 + trow + 16 * x;
          if (q < Kin && r < BC) {
            float* d = dst + (size_t)r * Kin + q;
            *d = first ? acc[x][y] : *d + acc[x][y];
          }
        }
    }
  }
}

// THREADS = 128 * US: the chunk's 64 pixels x BC units are tiled as 16 pixel groups (4 px) x 8 lanes x US unit splits,
// so one 64-pixel chunk is worked on by 4*US warps.  The step time is the latency of ONE chunk on ONE SM (every CTA
// has at most one chunk per step at bs <= 64*grid), so more warps per chunk is what shortens the step.
// MMA = 1: the chunk's GEMMs on warp-level 3xTF32 tensor-core MMAs (64-pixel chunks, bc a multiple of 32); MMA = 2: on fp16
// hi+lo split operands prepared once by their producers and read with ldmatrix (bc a multiple of 64, weights in shared
// memory); MMA = 0: FFMA.
template <int BC, int CP, bool WSMEM, int THREADS, int TM, int MMA = 0>
__global__ void __launch_bounds__(THREADS, 1) train_fp32_kernel(const TrainArgs a) {
  static_assert(!MMA || (THREADS == 512 && BC % 32 == 0 && (TM == 4 || ((MMA == 1 || MMA == 4) && TM == 2 && BC % 64 == 0))),
                "MMA paths: 16 warps, 64-pixel chunks (3xTF32 and streamed tcgen05 also 32-pixel chunks: bc 256)");
  static_assert((MMA != 2 && MMA != 3) || TM == 4, "fp16-split and resident tcgen05 paths: 64-pixel chunks");
  static_assert(MMA != 4 || (!WSMEM && TM == 2 && BC % 128 == 0), "streamed tcgen05 path: 32-pixel chunks, weights from the global image");
  constexpr int PT = train_npix(TM) / 16, NGW = 16 / PT;      // MMA = 1: pixel tiles per chunk, unit groups (warp = (tile, group))
  static_assert(MMA != 2 || (WSMEM && BC % 64 == 0), "fp16-split path: weights in shared memory, unit tiles in pairs");
  static_assert(MMA != 3 || (WSMEM && BC == 64), "tcgen05 path: M = 64 chunk GEMMs, everything resident in shared memory");
  constexpr bool H2 = MMA == 2;
  constexpr bool TC5 = MMA == 3;          // chunk GEMMs on tcgen05 (M = 64, accumulators in TMEM), see the TC5 block below
  constexpr bool TCX = MMA == 4;          // wide layers on tcgen05: transposed GEMMs (M = 128 units, N = 32 pixels), weights
                                          // streamed K step by K step from the global operand image, see the TCX block below
  constexpr int NPIX = train_npix(TM), LDP = train_ldp(TM), US = THREADS / 128, TN = BC / 8 / US;
  constexpr int VEC = TN >= 4 ? 4 : TN;
  static_assert(TN >= 1 && BC % (8 * US) == 0, "unit split does not divide bc");
  const Net& net = a.net;
  const int tid = threadIdx.x, us = tid >> 7, t128 = tid & 127, tn = t128 & 7, pg = t128 >> 3;
  const int ubase = us * (BC / US) + tn * VEC;
  auto unit = [&](int j) { return ubase + (j / VEC) * 8 * VEC + (j % VEC); };
  const int C = net.C, D = net.D, n = net.n, L = net.nl, P = net.P;
  const float maxv = net_maxv(net);

  // ---- shared memory carve-up ----------------------------------------------------------------------------
  extern __shared__ float4 smem4[];
  float* X = reinterpret_cast<float*>(smem4);                   // [dimpad][LDP]   features
  // H2: Hbuf holds the fp32 output of the LAST hidden layer only (output layer and its gradients read it); a layer's slot
  // of Gbuf is BC*kLDH floats: act' in fp32 [BC][LDP], overwritten in place by the scaled dz as [hi [BC][kLDH] | lo [BC][kLDH]]
  constexpr int GSZ = H2 ? BC * kLDH : BC * LDP;
  // TCX: the feature staging X shares its 32 KB with the two-stage weight ring (the gather is over before a GEMM starts)
  constexpr int TX_RING = 2 * 16384;
  // TC5: X lives over the images of the hidden outputs and the dz's (all dead between the end of a chunk and its first
  // epilogue): no bytes of its own (41 KB at dim_in 150 -- what lets coordinates + colours fit the tcgen05 variant)
  const size_t x_floats = TCX ? (size_t)((a.dimpad * LDP * 4 > TX_RING ? a.dimpad * LDP * 4 : TX_RING) / 4)
                              : (TC5 ? (size_t)0 : (size_t)a.dimpad * LDP);
  float* Hbuf = X + x_floats;                                   // [L][BC][LDP]    hidden outputs
  float* Gbuf = Hbuf + (size_t)((TC5 || TCX) ? 0 : (H2 ? 1 : L)) * BC * LDP;   // [L][GSZ]  act' then dz
  float* dZo = Gbuf + (size_t)(TCX ? 0 : L) * GSZ;              // [CP][LDP]       output-layer dz
  float* Tl = dZo + ((TC5 || TCX) ? 0 : CP * LDP);              // [CP][LDP]       labels
  float* Pp = Tl + CP * LDP;                                    // [US][CP][LDP]   output-layer partial sums per unit split
  float* wsm = Pp + (TC5 ? 0 : (TCX ? CP * LDP : US * CP * LDP));   // packed weights [P] (+pad) then natural hidden l>=1 (TCX: the Pp slot holds the fp32 dz_o)
  float* wnat_sm = wsm + round4(P);
  // H2: wsm = [biases of the hidden layers L*BC | W_o C*BC | b_o C] in fp32; behind it the split 16-bit operand arrays
  const int KP0 = round16(net.dim_in);                          // layer-0 contraction length, zero-padded to k = 16 steps
  float* h2base = wsm + round4(L * BC + C * BC + C);
  const uint32_t xh_hi = smem_u32(h2base), xh_lo = xh_hi + (uint32_t)KP0 * kLDH * 2u;      // X split: [KP0][kLDH] halves x 2
  const uint32_t hh_base = xh_hi + (uint32_t)KP0 * kLDH * 4u;   // outputs of hidden layers l < L-1: [BC][kLDH] halves x 2 each
  const uint32_t wh_base = hh_base + (uint32_t)(L - 1) * BC * kLDH * 4u;   // W_l natural [BC][ldw_l] halves x 2 (hi | lo)
  auto ldw_of = [&](int l) { return (l == 0 ? KP0 : BC) + 8; };  // row pitch in halves: an odd multiple of 16 bytes
  auto wh_of = [&](int l) { return wh_base + (l == 0 ? 0u : (uint32_t)BC * (KP0 + 8) * 4u + (uint32_t)(l - 1) * BC * (BC + 8) * 4u); };
  // ---- TC5: operand images (lbdrn_umma.cuh: img_off) behind the fp32 part, 128 B-aligned ---------------------------------
  //   weights (byte for byte the global image a.wimg): per hidden layer [hi Kp*BC | lo Kp*BC] halves (the output layer stays in
  //   fp32: it is applied from registers); X split [hi (KP0+8)*NPIX | lo]; hidden outputs h_l [hi (BC+8)*NPIX | lo] (the extra group of 8
  //   is the constant block 1,0,..,0 per pixel in hi, zero in lo: read as one more input row it makes the bias gradient a
  //   column of the weight-gradient GEMM); scaled dz_l [hi BC*NPIX | lo]; scaled output dz [hi 16*NPIX | lo].
  const uint32_t t5_base = (smem_u32(h2base) + 127u) & ~127u;
  auto t5_wbytes = [&](int l) { return (uint32_t)(l == 0 ? KP0 : BC) * BC * 2u; };      // one half (hi or lo) of layer l
  auto t5_w = [&](int l) { return t5_base + (l == 0 ? 0u : 2u * t5_wbytes(0) + (uint32_t)(l - 1) * 2u * t5_wbytes(1)); };
  const uint32_t t5_wend = t5_w(L);
  const uint32_t t5_xbytes = (uint32_t)(KP0 + 8) * NPIX * 2u, t5_hbytes = (uint32_t)(BC + 8) * NPIX * 2u;
  const uint32_t t5_x = t5_wend, t5_h0 = t5_x + 2u * t5_xbytes;
  auto t5_h = [&](int l) { return t5_h0 + (uint32_t)l * 2u * t5_hbytes; };
  const uint32_t t5_dzbytes = (uint32_t)BC * NPIX * 2u;
  auto t5_dz = [&](int l) { return t5_h0 + (uint32_t)L * 2u * t5_hbytes + (uint32_t)l * 2u * t5_dzbytes; };
  const uint32_t t5_dzo = t5_dz(L), t5_dzobytes = 16u * NPIX * 2u;
  if (TC5) X = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(smem4) + (t5_h0 - smem_u32(smem4)));
  // fp32 scratch of the output layer, over the (not yet written) dz_0 image: partial sums per column group, then dz_o
  float* const Pp5 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(smem4) + (t5_dz(0) - smem_u32(smem4)));   // [4][CP][LDP]
  float* const dZo5 = Pp5 + 4 * CP * LDP;                                                                          // [CP][LDP]
  static_assert(!TC5 || 5 * CP * LDP * 4 <= 2 * BC * NPIX * 2, "output-layer scratch must fit in one dz image");
  // TMEM columns: two forward / dh accumulators, then the weight-gradient accumulators
  constexpr uint32_t T5_DWO = 2 * BC + 16, T5_DW0 = 2 * BC + 32;
  auto t5_dwcol = [&](int l) { return T5_DW0 + (l == 0 ? 0u : (uint32_t)(KP0 + 8) + (uint32_t)(l - 1) * (BC + 8)); };
  // ---- TCX: operand images with NPIX = 32 rows behind the fp32 part (same element rule, rows = 32); W_o as a 16-row image
  //   [hi | lo]; X split [hi (KP0+8)*32 | lo]; h_l [hi (BC+8)*32 | lo] for every hidden layer; per layer one region that holds
  //   act' as fp32 [32][BC] until the backward pass overwrites it with the scaled dz_l image [hi BC*32 | lo]; output dz image.
  //   The hidden weights are NOT resident: a.wimg holds their images (rows = BC) and the GEMMs stream K steps of them.
  const uint32_t tx_base = (smem_u32(h2base) + 127u) & ~127u;
  const uint32_t tx_wobytes = 16u * BC * 2u;                                    // one half of the W_o image
  const uint32_t tx_wo = tx_base, tx_x = tx_wo + 2u * tx_wobytes;
  const uint32_t tx_xbytes = (uint32_t)(KP0 + 8) * NPIX * 2u, tx_hbytes = (uint32_t)(BC + 8) * NPIX * 2u;
  const uint32_t tx_h0 = tx_x + 2u * tx_xbytes;
  auto tx_h = [&](int l) { return tx_h0 + (uint32_t)l * 2u * tx_hbytes; };
  const uint32_t tx_dzbytes = (uint32_t)BC * NPIX * 2u;
  auto tx_dz = [&](int l) { return tx_h0 + (uint32_t)L * 2u * tx_hbytes + (uint32_t)l * 2u * tx_dzbytes; };
  const uint32_t tx_dzo = tx_dz(L), tx_dzobytes = 16u * NPIX * 2u, tx_end = tx_dzo + 2u * tx_dzobytes;
  auto tx_g = [&](int l) {                                                       // act' of layer l, fp32 [NPIX][BC], in dz_l's region
    return reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(smem4) + (tx_dz(l) - smem_u32(smem4)));
  };
  // global image: per hidden layer [hi Kp*BC | lo Kp*BC] halves (rows = BC), then W_o [hi 16*BC | lo]
  auto tx_wimg = [&](int l) -> const uint8_t* {                                  // hi half of layer l (l == L: W_o)
    size_t h = 0;
    for (int i = 0; i < l; ++i) h += 2u * (size_t)(i == 0 ? KP0 : BC) * BC;
    return reinterpret_cast<const uint8_t*>(a.wimg) + 2u * h;
  };
  // TMEM columns: two slots of BC/128 x 32 (transposed z / dh: units in lanes, pixels in columns), output accumulator,
  // W_o gradient (BC/128 x 16), then one weight-gradient pass at a time (128 units x (K + 8) columns)
  constexpr uint32_t TX_SLOT = (BC / 128) * 32, TX_ZO = 2 * TX_SLOT, TX_DWO = TX_ZO + 16, TX_DWP = TX_DWO + (BC / 128) * 16;
  __shared__ __align__(8) uint64_t s_tx_free[4];                // TCX: "the MMAs that read ring stage s have retired"
  __shared__ float s_txdb[TCX ? 2 : 1][TCX ? BC : 1];           // TCX: bias-gradient sums of the two pixel halves
  uint32_t tx_fpar = 0u;                                        // bit s: parity of the next wait on s_tx_free[s]
  __shared__ __align__(8) uint64_t s_t5_mbar;
  __shared__ uint32_t s_t5_tmem;
  __shared__ float s_t5_inv[kMaxLayers];                        // TC5: 1 / scale of dz_l (read-out of the gradient accumulators)
  uint32_t t5_phase = 0u;
  __shared__ uint32_t s_dzmax[kMaxLayers];                      // H2: bits of max |dz_l| over the chunk
  __shared__ float s_db[H2 ? 2 : 1][4][H2 ? BC : 1];            // H2: bias-gradient partials per pixel tile (by layer parity)
  __shared__ int s_py[NPIX], s_px[NPIX], s_valid[NPIX];
  __shared__ float s_red[THREADS / 32];
  __shared__ float s_sse;
  __shared__ float s_adam[2];

  const float* w = WSMEM ? wsm : a.wpack;
  const float* wo_p = (H2 || TC5 || TCX) ? wsm + L * BC : w + net.woff[L];   // output layer W_o [C][BC] and b_o [C]
  const float* bo_p = (H2 || TC5 || TCX) ? wsm + L * BC + C * BC : w + net.boff[L];
  // natural (untransposed) W_l for hidden layers l>=1, used as the k-major B operand of dh = W^T dz
  auto wnat = [&](int l) -> const float* {
    return WSMEM ? (wnat_sm + (size_t)(l - 1) * BC * BC) : (a.params + net.woff[l]);
  };
  auto layer_sync = [&]() {
    if (US == 1) __syncwarp();       // a pixel group's columns are private to its warp
    else __syncthreads();            // columns are shared by the US warps holding different units
  };

  for (int i = net.dim_in * LDP + tid; i < a.dimpad * LDP; i += THREADS) X[i] = 0.f;   // padding rows stay zero
  uint32_t t5_tmem = 0u;
  if constexpr (TC5) {
    // tensor memory: one CTA per SM (shared memory), all 512 columns; constant operand blocks; completion barrier
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_t5_tmem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) mbar_init(smem_u32(&s_t5_mbar), 1);
    // everything behind the weights starts as zero (padding features, bands >= C of the output dz, lo halves of the
    // constant blocks), then the hi halves of the constant blocks: element 0 of each pixel's 16-byte row = 1.0
    for (uint32_t o = t5_wend + 16u * tid; o < t5_dzo + 2u * t5_dzobytes; o += 16u * THREADS)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(o), "r"(0u) : "memory");
    __syncthreads();
    for (int i = tid; i < (L + 1) * NPIX; i += THREADS) {
      const int which = i / NPIX, pp = i - which * NPIX;
      const uint32_t blk = which == 0 ? t5_x + (uint32_t)KP0 * NPIX * 2u : t5_h(which - 1) + (uint32_t)BC * NPIX * 2u;
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(blk + 16u * pp), "h"((uint16_t)0x3c00) : "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    t5_tmem = s_t5_tmem;
  }
  if constexpr (TCX) {
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_t5_tmem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
      mbar_init(smem_u32(&s_t5_mbar), 1);
      for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&s_tx_free[i]), 1);
    }
    for (uint32_t o = tx_base + 16u * tid; o < tx_end; o += 16u * THREADS)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(o), "r"(0u) : "memory");
    __syncthreads();
    for (int i = tid; i < (L + 1) * NPIX; i += THREADS) {       // constant blocks: element 0 of each pixel's 16-byte row = 1.0
      const int which = i / NPIX, pp = i - which * NPIX;
      const uint32_t blk = which == 0 ? tx_x + (uint32_t)KP0 * NPIX * 2u : tx_h(which - 1) + (uint32_t)BC * NPIX * 2u;
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(blk + 16u * pp), "h"((uint16_t)0x3c00) : "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    t5_tmem = s_t5_tmem;
  }
  long long t_prev = 0;
  const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
  if (prof) t_prev = clock64();
#define LBDRN_PHASE(idx)                                   \
  if (prof) {                                              \
    const long long t_now = clock64();                     \
    a.prof[idx] += t_now - t_prev;                         \
    t_prev = t_now;                                        \
  }

  // ---- neighbourhood prefetch (uint8 planes, colours only, window <= 5, <= 4 bands, 8 threads per pixel) ----------------
  // The gather is bound by LSU sector throughput, not latency: 64 pixels x 100 scattered byte loads, 32 different sectors per
  // warp instruction = ~10k cycles per step.  Two things fix it:
  //  * a window row (n <= 5 contiguous bytes, interior pixels) is fetched as the TWO aligned 32-bit words that cover it
  //    (2 requests instead of 5) and the bytes are extracted with funnel shifts; the pair of threads that owns a band
  //    (even: rows 0-2, label; odd: rows 3-4, centre) needs 7 / 5 loads per pixel instead of 19;
  //  * the pixels of step s+1 are known before step s ends (the permutation is an input) and do not depend on the weights:
  //    the index is read at the start of step s, the words are requested right after this CTA's arrival at the first grid barrier, so they arrive
  //    under the reduction / Adam phase, and are normalised into X at the head of step s+1 (table of the 256 quotients
  //    float32(m)/MSB.max(): same correctly-rounded values as the reference's division, LBDRNdataset.py:120).
  // Border pixels (reflected columns) and windows at the very ends of the buffer are fetched byte by byte into the same
  // registers; every other configuration takes the plain gather.
  constexpr int NSH_ = THREADS / NPIX;
  constexpr int PN = 5, PD = 2;              // the window the prefetch path is specialised for (D = 2: the paper's setting)
  const bool pf_enabled = a.mode == TRAIN_FUSED && NSH_ == 8 && net.nco == 0 && net.ncol != 0 && n == PN && C <= 4 &&
                          !net.msb_u16 && !net.lsb_u16;
  uint32_t pf_w[3][2], pf_aux = 0u;          // [row][word]; aux: label code (even thread) or centre byte (odd thread)
  bool pf_have = false;
  long long gp_idx = 0, gp_pos = -1;                 // generic gather: permutation entry loaded one chunk ahead
  __shared__ uint32_t s_ctrw[NPIX];                  // generic gather, interleaved planes: the pixel's own word (centres)
  int pf_state = 0;                          // 1: aligned words in registers, 2: window bytes already extracted (border
                                             // pixel: reflected per-byte loads), 3: padding lane (beyond the batch)
  int pf_gy = 0, pf_gx = 0;                  // pixel the registers belong to
  // TMA variant (state 5): ONE cp.async.bulk.tensor per interior pixel lands the 32 x 5 x C byte box around its window
  // (16 B-aligned superset) in shared memory, a second one the 16 x 1 x C box holding its labels; 64 + 64 asynchronous
  // copies per step (four lanes of every warp issue) replace ~3 500 scattered word loads whose outstanding misses stalled every warp
  // for ~7k cycles at 8192^2 (DRAM-latency x MLP bound).  Border pixels keep the register path.
  const bool tma_on = pf_enabled && a.tmap_msb != nullptr;
  // TMA destinations must be 128 B-aligned in the shared window; the dynamic region itself is only 16 B-aligned
  uint8_t* const pfbox = reinterpret_cast<uint8_t*>(smem4) + a.pf_off +
                         ((128u - ((smem_u32(smem4) + (uint32_t)a.pf_off) & 127u)) & 127u);
  const int pf_lsb_off = a.pf_stride - 128;  // the label box sits in the last 128 B of a pixel's slot
  __shared__ __align__(8) uint64_t s_pf_mbar;
  uint32_t pf_phase = 0u;
  bool pf_tma_pending = false;               // uniform: copies of the next chunk are in flight
  if (tma_on && tid == 0) mbar_init(smem_u32(&s_pf_mbar), NPIX);
  bool pf_none = true;                       // this CTA has no chunk in the next step (uniform)
  __shared__ int s_ny[NPIX], s_nx[NPIX];     // next step's pixel coordinates (one division per pixel, not per thread)
  __shared__ float s_quot[256];
  if (!net.msb_u16)                          // value / max for every 8-bit value: one correctly rounded division each, once
    for (int i = tid; i < 256; i += THREADS) s_quot[i] = __fdiv_rn((float)i, maxv);
  const size_t pf_total = (size_t)C * net.buf_rows * net.W;     // bytes in the MSB buffer
  const bool idx32 = (long long)net.H * net.W < (1ll << 31);
  // stage A (start of step s): coordinates of the pixels of step s+1
  // the permutation entry itself is read one step earlier still (its global-memory latency was exposed at the head of every step)
  long long pf_idx = 0;
  auto prefetch_perm = [&](int s_nn) {
    if (!pf_enabled || s_nn >= a.n_steps || tid >= NPIX) return;
    const long long o = (long long)s_nn * a.bs + (long long)blockIdx.x * NPIX + tid;
    if (o < a.n_perm) pf_idx = __ldg(a.perm + o);
  };
  prefetch_perm(1);
  auto prefetch_index = [&](int s_next) {
    pf_none = true;
    if (!pf_enabled || s_next >= a.n_steps) return;
    const long long b0n = (long long)s_next * a.bs;
    const long long remn = a.n_perm - b0n;
    const int Bn = (int)(remn < a.bs ? remn : a.bs);
    const int ch = blockIdx.x;
    if (Bn <= 0 || ch * NPIX >= Bn) return;                       // this CTA has no chunk in the next step
    pf_none = false;
    if (tid < NPIX) {
      int y = -1, x = 0;                                           // y = -1: padding lane
      if (ch * NPIX + tid < Bn) {
        const long long idx = pf_idx;                              // = a.perm[b0n + ch * NPIX + tid], requested a step ago
        if (idx32) { y = (int)((unsigned)idx / (unsigned)net.W); x = (int)((unsigned)idx - (unsigned)y * (unsigned)net.W); }
        else { y = (int)(idx / net.W); x = (int)(idx - (long long)y * net.W); }
      }
      s_ny[tid] = y; s_nx[tid] = x;
    }
  };
  // stage A' (head of step s, once the coordinates are in shared memory): L2 prefetch hints for the same bytes.  At 8192^2
  // the planes (2 x 268 MB) do not fit in L2, every window row of a random pixel is a DRAM (and TLB) miss, and the register
  // loads of stage B were bound by the number of misses an SM can keep in flight (issue phase 7k cycles vs 3k at 2048^2).
  // The hints hold no registers or scoreboard entries and have a whole step to land.
  auto prefetch_l2 = [&]() {
    if (pf_none || tma_on || !a.l2_hints) return;
    const int pp = tid & (NPIX - 1), share = tid / NPIX, c = share >> 1, odd = share & 1;
    const int gy = s_ny[pp], gx = s_nx[pp];
    if (gy < 0 || c >= C) return;
    const size_t plane = (size_t)c * net.buf_rows;
    const uint8_t* base = (const uint8_t*)a.msb;
    const int x0 = gx >= PD ? gx - PD : 0, x1 = gx + PD < net.W ? gx + PD : net.W - 1;
    const int r0 = odd ? 3 : 0, r1 = odd ? PN : 3;
    if (!odd) {
      const uint8_t* q = (const uint8_t*)a.lsb + (plane + (gy - net.buf_row0)) * net.W + gx;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int dy = r0 + q;
      if (dy < r1) {
        const uint8_t* row = base + (plane + (reflect_clamp(gy + dy - PD, net.H) - net.buf_row0)) * net.W;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(row + x0));
        if ((reinterpret_cast<uintptr_t>(row + x0) ^ reinterpret_cast<uintptr_t>(row + x1)) >> 5)     // window spans two sectors
          asm volatile("prefetch.global.L2 [%0];" ::"l"(row + x1));
      }
    }
  };
  // stage B (after arriving at the first grid barrier of step s): request the words.  Rows of the window this thread owns: band = share / 2;
  // even thread rows [0, min(3, n)), odd thread rows [3, n)
  auto prefetch_issue = [&]() {
    pf_have = !pf_none;
    if (pf_none) return;
    const int pp = tid & (NPIX - 1), share = tid / NPIX, c = share >> 1, odd = share & 1;
    pf_gy = s_ny[pp]; pf_gx = s_nx[pp];
    if (tma_on) {
      pf_tma_pending = false;
      if ((tid & 7) == 0) {                                          // 4 lanes of every warp: thread 8p issues for pixel p
        const int p = tid >> 3, y = s_ny[p], x = s_nx[p];
        const bool in = y >= PD && y + PD < net.H && x >= PD && x + PD < net.W;
        // one arrival per pixel (the barrier counts NPIX): with the bytes of its two copies, or with none for a border /
        // padding pixel -- the commit always waits for exactly one phase per prefetched chunk
        mbar_expect_tx(smem_u32(&s_pf_mbar), in ? (uint32_t)(32 * PN * C + 16 * C) : 0u);
        if (in) {
          const uint32_t dst = smem_u32(pfbox + (size_t)p * a.pf_stride);
          tma_load_3d(dst, a.tmap_msb, (x - PD) & ~15, y - PD, 0, smem_u32(&s_pf_mbar));
          tma_load_3d(dst + pf_lsb_off, a.tmap_lsb, x & ~15, y, 0, smem_u32(&s_pf_mbar));
        }
      }
      pf_tma_pending = true;
    }
    pf_state = 3;
    if (pf_gy < 0 || c >= C) return;
    const int gy = pf_gy, gx = pf_gx;
    if (tma_on && gy >= PD && gy + PD < net.H && gx >= PD && gx + PD < net.W) { pf_state = 5; return; }
    const size_t plane = (size_t)c * net.buf_rows;
    const size_t off = (plane + (gy - net.buf_row0)) * net.W + gx;
    const uint8_t* base = (const uint8_t*)a.msb;
    const int r0 = odd ? 3 : 0, r1 = odd ? PN : 3;
    pf_aux = odd ? (uint32_t)base[off] : (uint32_t)((const uint8_t*)a.lsb)[off];
    // the two aligned words of a row may extend 3 bytes before / 6 bytes after the window: stay inside the buffer, and the
    // window must not need reflected columns
    bool fast = gx >= PD && gx + PD < net.W;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int dy = r0 + q;
      if (dy < r1) {
        const size_t first = (plane + (reflect_clamp(gy + dy - PD, net.H) - net.buf_row0)) * net.W + gx - PD;
        fast = fast && first >= 4 && first + 12 <= pf_total;
      }
    }
    pf_state = fast ? 1 : 2;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int dy = r0 + q;
      if (dy < r1) {
        const size_t rowoff = (plane + (reflect_clamp(gy + dy - PD, net.H) - net.buf_row0)) * net.W;
        if (fast) {
          const uint8_t* p = base + rowoff + gx - PD;
          const uint32_t* w = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
          pf_w[q][0] = w[0];
          pf_w[q][1] = w[1];
        } else {                                                     // rare: reflected columns, byte by byte
          uint32_t lo4 = 0u, hi4 = 0u;
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            {
              const uint32_t b = base[rowoff + reflect_clamp(gx + dx - PD, net.W)];
              if (dx < 4) lo4 |= b << (8 * dx); else hi4 = b;
            }
          }
          pf_w[q][0] = lo4; pf_w[q][1] = hi4;
        }
      }
    }
  };
  // ---- interleaved variant (a.imsb): a window row of ALL bands is 5 consecutive 32-bit words, i.e. inside two aligned 16 B
  // loads -- 11 requests per pixel (5 rows x 2 + labels) + 4 centre words instead of ~48, and a third of the DRAM sectors.
  // Thread share = tid / NPIX: shares 0-4 own window row dy = share (8 words + the centre word), share 5 the labels.
  const bool il = pf_enabled && a.imsb != nullptr && !tma_on;
  const bool gil = a.imsb != nullptr && net.ncol != 0 && net.nco == 0 && (n == 3 || n == 5 || n == 7);   // generic gather from the interleaved copies
  // registers shared with the CHW variant: pf_w[3][2] + il_w6, il_w7 = the aligned 32 bytes holding the row (border pixel: the
  // 5 words, packed); pf_aux = centre pixel (rows != D) or label word (share 5)
  uint32_t il_w6 = 0u, il_w7 = 0u;
  int il_o = 0;                                           // word offset of the window inside the 8 words
  auto il_row_ptr = [&](int gy, int dy) {
    return a.imsb + (size_t)(reflect_clamp(gy + dy - PD, net.H) - net.buf_row0) * net.W;
  };
  auto prefetch_l2_il = [&]() {
    if (pf_none || !a.l2_hints) return;
    const int pp = tid & (NPIX - 1), share = tid / NPIX;
    const int gy = s_ny[pp], gx = s_nx[pp];
    if (gy < 0 || share > PN) return;
    if (share == PN) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ilsb + (size_t)(gy - net.buf_row0) * net.W + gx));
      return;
    }
    const uint32_t* row = il_row_ptr(gy, share);
    const int x0 = gx >= PD ? gx - PD : 0, x1 = gx + PD < net.W ? gx + PD : net.W - 1;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(row + x0));
    if ((reinterpret_cast<uintptr_t>(row + x0) ^ reinterpret_cast<uintptr_t>(row + x1)) >> 5)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(row + x1));
  };
  auto prefetch_issue_il = [&]() {
    pf_have = !pf_none;
    if (pf_none) return;
    const int pp = tid & (NPIX - 1), share = tid / NPIX;
    pf_gy = s_ny[pp]; pf_gx = s_nx[pp];
    pf_state = 3;
    if (pf_gy < 0 || share > PN) return;
    const int gy = pf_gy, gx = pf_gx;
    pf_state = 1;
    if (share == PN) {                                              // labels of all bands: one word
      pf_aux = a.ilsb[(size_t)(gy - net.buf_row0) * net.W + gx];
      return;
    }
    const uint32_t* row = il_row_ptr(gy, share);
    if (share != PD) pf_aux = a.imsb[(size_t)(gy - net.buf_row0) * net.W + gx];
    if (gx >= PD && gx + PD < net.W) {
      // the aligned 32 bytes around the 20-byte window (the buffers carry 32 bytes of slack behind the last pixel)
      const uint32_t* p = row + gx - PD;
      const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)15);
      il_o = (int)((reinterpret_cast<uintptr_t>(p) >> 2) & 3);
      const uint4 qa = q[0], qb = q[1];
      pf_w[0][0] = qa.x; pf_w[0][1] = qa.y; pf_w[1][0] = qa.z; pf_w[1][1] = qa.w;
      pf_w[2][0] = qb.x; pf_w[2][1] = qb.y; il_w6 = qb.z; il_w7 = qb.w;
    } else {                                                        // reflected columns: word by word
      il_o = 0;
      pf_w[0][0] = row[reflect_clamp(gx - 2, net.W)]; pf_w[0][1] = row[reflect_clamp(gx - 1, net.W)]; pf_w[1][0] = row[gx];
      pf_w[1][1] = row[reflect_clamp(gx + 1, net.W)]; pf_w[2][0] = row[reflect_clamp(gx + 2, net.W)];
    }
  };
  auto prefetch_commit_il = [&]() {
    const int pp = tid & (NPIX - 1), share = tid / NPIX;
    const bool ok = pf_state == 1;
    if (tid < NPIX) s_valid[tid] = pf_gy >= 0;
    if (share == PN) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < C) Tl[c * LDP + pp] = ok ? __fdiv_rn((float)((pf_aux >> (8 * c)) & 255u), net.qmax) : 0.f;
    } else if (share < PN) {
      // window words v[dx] = w[il_o + dx] without indexing registers dynamically
      const uint32_t w[8] = {pf_w[0][0], pf_w[0][1], pf_w[1][0], pf_w[1][1], pf_w[2][0], pf_w[2][1], il_w6, il_w7};
      uint32_t v[PN];
#pragma unroll
      for (int dx = 0; dx < PN; ++dx)
        v[dx] = il_o == 0 ? w[dx] : (il_o == 1 ? w[dx + 1] : (il_o == 2 ? w[dx + 2] : w[dx + 3]));
      const uint32_t cw = share == PD ? v[PD] : pf_aux;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < C) {
          const float ctr = net.relative ? s_quot[(cw >> (8 * c)) & 255u] : 0.f;
          float* d = X + pp + (size_t)((c * PN + share) * PN) * LDP;
#pragma unroll
          for (int dx = 0; dx < PN; ++dx) d[(size_t)dx * LDP] = ok ? s_quot[(v[dx] >> (8 * c)) & 255u] - ctr : 0.f;
        }
      }
    }
  };
  // head of step s+1: registers -> X / Tl / s_valid (the values the plain gather would write)
  auto prefetch_commit = [&]() {
    const int pp = tid & (NPIX - 1), share = tid / NPIX, c = share >> 1, odd = share & 1;
    const bool ok = pf_state == 1 || pf_state == 2 || pf_state == 5;
    if (pf_tma_pending) {                                            // uniform
      mbar_wait(smem_u32(&s_pf_mbar), pf_phase, 7, (int)blockIdx.x);
      pf_phase ^= 1u;
      pf_tma_pending = false;
    }
    if (tid < NPIX) s_valid[tid] = pf_gy >= 0;
    if (c < C) {
      const int gy = ok ? pf_gy : 0, gx = ok ? pf_gx : 0;
      const size_t plane = (size_t)c * net.buf_rows;
      const int r0 = odd ? 3 : 0, r1 = odd ? PN : 3;
      if (pf_state == 5) {                         // this pixel's boxes are in shared memory
        const uint8_t* slot = pfbox + (size_t)pp * a.pf_stride;
        const int o = (gx - PD) & 15;              // window start inside the 16 B-aligned box row
        pf_aux = odd ? (uint32_t)slot[(c * PN + PD) * 32 + o + PD] : (uint32_t)slot[pf_lsb_off + c * 16 + (gx & 15)];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int dy = r0 + q;
          if (dy < r1) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(slot + (c * PN + dy) * 32 + (o & ~3));
            const uint32_t sh = (uint32_t)(o & 3) * 8u;
            pf_w[q][0] = __funnelshift_r(w[0], w[1], sh);
            pf_w[q][1] = w[1] >> sh;
          }
        }
      }
      uint32_t ctr_raw = odd ? pf_aux : 0u;      // the odd thread loaded the centre itself; row D <= 2 belongs to the even one
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int dy = r0 + q;
        if (dy < r1) {
          if (pf_state == 1) {
            const uintptr_t addr = reinterpret_cast<uintptr_t>((const uint8_t*)a.msb) +
                                   (plane + (reflect_clamp(gy + dy - PD, net.H) - net.buf_row0)) * net.W + gx - PD;
            const uint32_t sh = (uint32_t)(addr & 3) * 8u;
            const uint32_t lo4 = __funnelshift_r(pf_w[q][0], pf_w[q][1], sh);      // window bytes 0..3
            const uint32_t hi4 = pf_w[q][1] >> sh;                                 // window byte 4 (n = 5)
            pf_w[q][0] = lo4; pf_w[q][1] = hi4;
          }
          if (!odd && dy == PD) ctr_raw = (pf_w[q][0] >> (8 * PD)) & 255u;
        }
      }
      const float ctr = net.relative ? s_quot[ctr_raw & 255u] : 0.f;
      if (!odd) Tl[c * LDP + pp] = ok ? __fdiv_rn((float)pf_aux, net.qmax) : 0.f;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int dy = r0 + q;
        if (dy < r1) {
          float* d = X + pp + (size_t)((c * PN + dy) * PN) * LDP;
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            {
              const uint32_t b = dx < 4 ? (pf_w[q][0] >> (8 * dx)) & 255u : pf_w[q][1] & 255u;
              d[(size_t)dx * LDP] = ok ? s_quot[b] - ctr : 0.f;
            }
          }
        }
      }
    }
    if (tma_on) fence_async_smem();      // our reads of the landing boxes are ordered before the next TMA writes into them
  };

  // TC5: features -> split operand image.  A warp covers 8 pixels x 4 feature pairs = 8 x 16 contiguous bytes per store
  // instruction (conflict-free), reading X at banks 8*pair + pixel (conflict-free).
  auto t5_split_x = [&]() {
    for (int i = tid; i < KP0 * (NPIX / 2); i += THREADS) {
      const int kp = i & 3, pp = ((i >> 2) & 7) + 8 * ((i >> 5) & 7), k = 8 * (i >> 8) + 2 * kp;
      const float x0 = k < net.dim_in ? X[(size_t)k * LDP + pp] : 0.f;
      const float x1 = k + 1 < net.dim_in ? X[(size_t)(k + 1) * LDP + pp] : 0.f;
      uint32_t hi, lo;
      split_h2(x0, x1, hi, lo);
      const uint32_t o = (uint32_t)img_off(NPIX, pp, k);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(t5_x + o), "r"(hi) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(t5_x + t5_xbytes + o), "r"(lo) : "memory");
    }
  };

  auto tx_split_x = [&]() {       // TCX: features -> split operand image with 32 rows (a warp: 8 pixels x 4 feature pairs)
    for (int i = tid; i < KP0 * (NPIX / 2); i += THREADS) {
      const int kp = i & 3, pp = ((i >> 2) & 7) + 8 * ((i >> 5) & 3), k = 8 * (i >> 7) + 2 * kp;
      const float x0 = k < net.dim_in ? X[(size_t)k * LDP + pp] : 0.f;
      const float x1 = k + 1 < net.dim_in ? X[(size_t)(k + 1) * LDP + pp] : 0.f;
      uint32_t hi, lo;
      split_h2(x0, x1, hi, lo);
      const uint32_t o = (uint32_t)img_off(NPIX, pp, k);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(tx_x + o), "r"(hi) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(tx_x + tx_xbytes + o), "r"(lo) : "memory");
    }
  };

  for (int s = 0; s < a.n_steps; ++s) {
    if (tid == 0) s_sse = 0.f;
    prefetch_index(s + 1);
    prefetch_perm(s + 2);
    if (tid == 32 && a.mode == TRAIN_FUSED) {
      if (a.adam_tab) {
        const float2 t2 = a.adam_tab[s];
        s_adam[0] = t2.x; s_adam[1] = t2.y;
      } else {
        double t = (double)(a.adam_t0 + s + 1);
        double bc1 = 1.0 - pow(a.beta1, t), bc2 = 1.0 - pow(a.beta2, t);
        s_adam[0] = (float)(a.lr / bc1);              // step_size
        s_adam[1] = (float)sqrt(bc2);                 // bias_correction2_sqrt
      }
    }
    // The features of this step's chunk were requested during step s-1 and do not depend on the weights: they are committed
    // to shared memory BEFORE waiting for the other CTAs' Adam phase (the second half of the split barrier).
    const bool early = pf_have;                       // uniform: a function of (step, CTA) only
    LBDRN_PHASE(21)  // step head: next coordinates, permutation prefetch, Adam table
    if (early) {
      if (il) prefetch_commit_il(); else prefetch_commit();
      pf_have = false;
    }
    LBDRN_PHASE(13)   // commit of the prefetched neighbourhoods
    if (s > 0 && tid == 0) gbar_wait(a.gbar + 1, (unsigned)s * gridDim.x);
    __syncthreads();
    LBDRN_PHASE(7)    // wait for the Adam phase of every CTA (second barrier of the previous step)
    if (il) prefetch_l2_il(); else prefetch_l2();
    // ---- (re)load weights ----------------------------------------------------------------------------
    if constexpr (TCX) {
      // fp32 part (hidden biases, W_o, b_o) and the 16-row W_o image; the hidden weights stay in the global image
      const int nlite = L * BC + C * BC + C;
      for (int i = tid; i < nlite; i += THREADS) {
        const int nb = L * BC, nwo = C * BC;
        const int src = i < nb ? net.boff[i / BC] + i % BC : (i < nb + nwo ? net.woff[L] + (i - nb) : net.boff[L] + (i - nb - nwo));
        wsm[i] = __ldcg(a.params + src);
      }
      const uint8_t* wo_img = tx_wimg(L);
      for (int i = tid; i < (int)(2u * tx_wobytes >> 4); i += THREADS)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tx_wo + 16u * i), "l"(wo_img + 16 * (size_t)i) : "memory");
      cp_async_commit();
      cp_async_wait_group<0>();
      fence_async_smem();
    }
    if (H2 || TC5) {
      const bool from_image = s > 0 && a.wimg != nullptr;           // the Adam phase of step s-1 wrote every weight's halves
      // fp32 part (hidden biases, output layer): requested FIRST (behind 49 KB of copies per SM its round trip tripled),
      // stored after the copies have been issued
      const int nb = L * BC, nwo = C * BC;
      auto lite_src = [&](int i) {
        return i < nb ? net.boff[i / BC] + i % BC : (i < nb + nwo ? net.woff[L] + (i - nb) : net.boff[L] + (i - nb - nwo));
      };
      float lite0 = 0.f;
      if (tid < nb + nwo + C) lite0 = __ldcg(a.params + lite_src(tid));
      if (from_image) {
        const uint32_t img0 = TC5 ? t5_base : wh_base;
        const int n16 = (int)(((TC5 ? t5_wend : wh_of(L)) - img0) >> 4);   // the image is the shared layout, byte for byte
        for (int i = tid; i < n16; i += THREADS)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(img0 + 16u * i),
                       "l"(reinterpret_cast<const uint4*>(a.wimg) + i) : "memory");
      }
      if (tid < nb + nwo + C) wsm[tid] = lite0;
      for (int i = tid + THREADS; i < nb + nwo + C; i += THREADS) wsm[i] = __ldcg(a.params + lite_src(i));
      // first step of a launch: hidden weights from the master copy, natural [unit][input] layout of the reference ->
      // scaled hi | lo halves, two inputs per thread
      for (int l = 0; l < (from_image ? 0 : L); ++l) {
        const int K = l == 0 ? net.dim_in : BC, KH = (l == 0 ? KP0 : BC) >> 1, ldw = ldw_of(l);
        const float* src = a.params + net.woff[l];
        const uint32_t w_hi = TC5 ? t5_w(l) : wh_of(l), w_lo = w_hi + (TC5 ? t5_wbytes(l) : (uint32_t)BC * ldw * 2u);
#pragma unroll 4
        for (int i = tid; i < BC * KH; i += THREADS) {
          const int u = i / KH, k = 2 * (i - u * KH);
          const float x0 = k < K ? __ldcg(src + (size_t)u * K + k) : 0.f;
          const float x1 = k + 1 < K ? __ldcg(src + (size_t)u * K + k + 1) : 0.f;
          uint32_t hi, lo;
          split_h2(x0 * kWScale, x1 * kWScale, hi, lo);
          const uint32_t o = TC5 ? (uint32_t)img_off(BC, u, k) : (uint32_t)(u * ldw + k) * 2u;
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(w_hi + o), "r"(hi) : "memory");
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(w_lo + o), "r"(lo) : "memory");
        }
      }
    } else if (WSMEM) {
      // asynchronous copies: their L2 latency hides under the commit of the prefetched neighbourhoods (or the gather),
      // neither of which reads weights; cp_async_wait_all() sits in front of the barrier that precedes the forward pass
      const int P4 = P >> 2;
      for (int i = tid; i < P4; i += THREADS)
        cp_async16(reinterpret_cast<float4*>(wsm) + i, reinterpret_cast<const float4*>(a.wpack) + i);
      for (int i = (P4 << 2) + tid; i < P; i += THREADS) wsm[i] = __ldcg(a.wpack + i);
      for (int l = 1; l < L; ++l) {
        const float4* src = reinterpret_cast<const float4*>(a.params + net.woff[l]);
        float4* dst = reinterpret_cast<float4*>(wnat_sm + (size_t)(l - 1) * BC * BC);
        for (int i = tid; i < BC * BC / 4; i += THREADS) cp_async16(dst + i, src + i);
      }
    }
    if (H2 && early) {
      // features -> split 16-bit operand arrays, under the latency of the weight copies
      for (int i = tid; i < KP0 * (NPIX / 2); i += THREADS) {
        const int k = i / (NPIX / 2), pp = 2 * (i - k * (NPIX / 2));
        float2 x = make_float2(0.f, 0.f);
        if (k < net.dim_in) x = *reinterpret_cast<const float2*>(X + (size_t)k * LDP + pp);
        uint32_t hi, lo;
        split_h2(x.x, x.y, hi, lo);
        const uint32_t o = (uint32_t)(k * kLDH + pp) * 2u;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(xh_hi + o), "r"(hi) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(xh_lo + o), "r"(lo) : "memory");
      }
    }
    if (TC5 && early) t5_split_x();
    if (WSMEM) cp_async_wait_all();
    if (TC5) fence_async_smem();      // weights and features were written through the generic proxy; the MMAs read them
    __syncthreads();
    LBDRN_PHASE(0)   // weight reload

    const long long b0 = a.mode == TRAIN_FUSED ? (long long)s * a.bs : 0;
    long long rem = a.n_perm - b0;
    const int B = (int)(rem < a.bs ? rem : a.bs);                 // last batch of an epoch may be partial
    const int Bglobal = a.mode == TRAIN_FUSED ? B : a.n_global;
    const float gscale = 2.0f / ((float)Bglobal * (float)C);       // d(mean((y-t)^2))/dy = 2 (y-t) / (B*C)
    const int n_chunks = (B + NPIX - 1) / NPIX;
    float* mypart = a.partial + (size_t)blockIdx.x * a.pstride;
    bool first = true;

    for (int ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
      __syncthreads();
      if (H2 && tid < kMaxLayers) s_dzmax[tid] = 0u;
      // ---- gather: pixel coordinates, centres, labels, features (LBDRNdataset.py:104-131,151-155) --------
      const int nvalid = min(NPIX, B - ch * NPIX);
      const bool staged = early && ch == (int)blockIdx.x;            // committed before the barrier wait above
      if (!staged) {
      if (tid < NPIX) {
        const long long pos = b0 + (long long)ch * NPIX + tid;
        long long idx = 0;
        if (tid < nvalid) idx = gp_pos == pos ? gp_idx : a.perm[pos];
        int y, x;
        if (idx32) { y = (int)((unsigned)idx / (unsigned)net.W); x = (int)((unsigned)idx - (unsigned)y * (unsigned)net.W); }
        else { y = (int)(idx / net.W); x = (int)(idx - (long long)y * net.W); }
        if (tid >= nvalid) { y = net.row0; x = 0; }      // padding lane of a partial chunk: a pixel that is inside the buffer
        s_py[tid] = y;
        s_px[tid] = x;
        s_valid[tid] = tid < nvalid;
        // the entry this thread needs next (this CTA's next chunk, or its first chunk of the next batch): the
        // load has a whole chunk of arithmetic to land under
        const bool more = ch + (int)gridDim.x < n_chunks;
        gp_pos = more ? pos + (long long)gridDim.x * NPIX
                      : (a.mode == TRAIN_FUSED ? b0 + a.bs + (long long)blockIdx.x * NPIX + tid : -1);
        if (gp_pos >= 0 && gp_pos < a.n_perm && (more || gp_pos < b0 + 2ll * a.bs)) gp_idx = a.perm[gp_pos];
        else gp_pos = -1;
        if (gil) {
          // interleaved planes: the pixel's own word holds the centre of every band, its LSB word every label
          // (padding lanes of a partial chunk read nothing: row 0 need not be inside the plane buffer)
          uint32_t lw = 0u, cw = 0u;
          if (tid < nvalid) {
            const size_t off = (size_t)(y - net.buf_row0) * net.W + x;
            lw = a.ilsb[off];
            cw = a.imsb[off];
          }
          s_ctrw[tid] = cw;
          for (int c = 0; c < C; ++c) Tl[c * LDP + tid] = __fdiv_rn((float)((lw >> (8 * c)) & 255u), net.qmax);
        }
      }
      __syncthreads();
      LBDRN_PHASE(22)  // pixel coordinates of the chunk
      if (gil) {
        // thread = (pixel, window row): the row of ALL bands is n consecutive words = two or three aligned 16-byte loads
        // (the gather is bound by L1 wavefronts: one request per lane and row instead of one per lane, band and row)
        auto gather_il = [&](auto nc) {
          constexpr int N_ = decltype(nc)::value;
          constexpr int D_ = N_ / 2;
          for (int it = tid; it < NPIX * N_; it += THREADS) {
            const int pp = it & (NPIX - 1), dy = it / NPIX;
            const int gy = s_py[pp], gx = s_px[pp];
            const bool ok = s_valid[pp] != 0;
            const uint32_t* row = a.imsb + (size_t)(reflect_clamp(gy + dy - D_, net.H) - net.buf_row0) * net.W;
            uint32_t v[N_];
            if (!ok) {
#pragma unroll
              for (int dx = 0; dx < N_; ++dx) v[dx] = 0u;
            } else if (gx >= D_ && gx + D_ < net.W) {
              const uint32_t* p0 = row + (gx - D_);
              const int o = (int)((reinterpret_cast<uintptr_t>(p0) >> 2) & 3);
              const uint4* q = reinterpret_cast<const uint4*>(p0 - o);
              const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
              uint4 q2 = make_uint4(0u, 0u, 0u, 0u);
              if (o + N_ > 8) q2 = __ldg(q + 2);
              const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
              for (int dx = 0; dx < N_; ++dx)
                v[dx] = o == 0 ? w[dx] : o == 1 ? w[dx + 1] : o == 2 ? w[dx + 2] : w[dx + 3];
            } else {
#pragma unroll
              for (int dx = 0; dx < N_; ++dx) v[dx] = row[reflect_clamp(gx + dx - D_, net.W)];
            }
            const uint32_t cw = net.relative ? s_ctrw[pp] : 0u;
            for (int c = 0; c < C; ++c) {
              const float ctr = net.relative ? s_quot[(cw >> (8 * c)) & 255u] : 0.f;
              float* d = X + pp + (size_t)(net.nco + (c * N_ + dy) * N_) * LDP;
#pragma unroll
              for (int dx = 0; dx < N_; ++dx) d[(size_t)dx * LDP] = ok ? s_quot[(v[dx] >> (8 * c)) & 255u] - ctr : 0.f;
            }
          }
        };
        switch (n) {
          case 3: gather_il(std::integral_constant<int, 3>{}); break;
          case 5: gather_il(std::integral_constant<int, 5>{}); break;
          default: gather_il(std::integral_constant<int, 7>{}); break;
        }
      } else
      {
        // thread = (pixel, share); a share takes (band, window-row) items straight from the resident planes
        constexpr int NSH = THREADS / NPIX;
        const int pp = tid & (NPIX - 1), share = tid / NPIX;
        const int gy = s_py[pp], gx = s_px[pp];
        const bool ok = s_valid[pp] != 0;
        float* dst = X + pp;
        if (net.nco && share == NSH - 1) {
          const float* trow = a.tab + (size_t)gy * net.tabw;
          const float* tcol = a.tab + (size_t)(net.H + gx) * net.tabw;
          for (int i = 0; i < net.tabw; ++i) {
            dst[(size_t)i * LDP] = ok ? trow[i] : 0.f;
            dst[(size_t)(net.tabw + i) * LDP] = ok ? tcol[i] : 0.f;
          }
        }
        const int n_items = net.ncol ? C * n : C;
        // usual window sizes: a thread takes a band (or half of its rows when there are shares to spare) and reads
        // each window row as two or three aligned 8-byte words: the gather is bound by L1 wavefronts (every lane of
        // a load hits another sector), so fewer, wider requests are what shortens it
        auto gather_n = [&](auto nc) {
          constexpr int N_ = decltype(nc)::value;
          constexpr int D_ = N_ / 2;
          const int RG = NSH >= 2 * C ? 2 : 1;
          const int rows_per = (N_ + RG - 1) / RG;
          for (int it = share; it < C * RG; it += NSH) {
            const int c = it / RG, g = it - c * RG;
            const int dy0 = g * rows_per, dy1 = min(N_, dy0 + rows_per);
            if (!ok) {                       // padding lane of a partial chunk: zeros, no loads
              if (g == 0) Tl[c * LDP + pp] = 0.f;
              for (int dy = dy0; dy < dy1; ++dy)
#pragma unroll
                for (int dx = 0; dx < N_; ++dx) dst[(size_t)(net.nco + (c * N_ + dy) * N_ + dx) * LDP] = 0.f;
              continue;
            }
            const size_t plane = (size_t)c * net.buf_rows;
            const size_t off = (plane + (gy - net.buf_row0)) * net.W + gx;
            if (g == 0) {
              const uint32_t code = net.lsb_u16 ? (uint32_t)((const uint16_t*)a.lsb)[off] : (uint32_t)((const uint8_t*)a.lsb)[off];
              Tl[c * LDP + pp] = __fdiv_rn((float)code, net.qmax);          // label = LSB/(2^K-1)
            }
            float ctr = 0.f;
            if (net.relative) ctr = net.msb_u16 ? load_msb_norm(a.msb, 1, off, maxv) : s_quot[((const uint8_t*)a.msb)[off]];
#pragma unroll
            for (int dy = 0; dy < N_; ++dy) {
              if (dy < dy0 || dy >= dy1) continue;
              const size_t rowoff = (plane + (reflect_clamp(gy + dy - D_, net.H) - net.buf_row0)) * net.W;
              gather_row_wide<N_, LDP>(a.msb, pf_total << net.msb_u16, net.msb_u16, rowoff, gx, net, maxv, s_quot, ctr, ok,
                                       dst + (size_t)(net.nco + (c * N_ + dy) * N_) * LDP);
            }
          }
        };
        bool batched = net.ncol != 0;
        switch (net.ncol ? n : 0) {
          case 3: gather_n(std::integral_constant<int, 3>{}); break;
          case 5: gather_n(std::integral_constant<int, 5>{}); break;
          case 7: gather_n(std::integral_constant<int, 7>{}); break;
          default: batched = false;
        }
        if (!batched)
        for (int it = share; it < n_items; it += NSH) {
          const int c = net.ncol ? it / n : it, dy = net.ncol ? it - c * n : 0;
          const size_t plane = (size_t)c * net.buf_rows;
          const size_t off = (plane + (gy - net.buf_row0)) * net.W + gx;
          if (dy == 0) {
            const uint32_t code = net.lsb_u16 ? (uint32_t)((const uint16_t*)a.lsb)[off] : (uint32_t)((const uint8_t*)a.lsb)[off];
            Tl[c * LDP + pp] = __fdiv_rn((float)code, net.qmax);          // label = LSB/(2^K-1)
          }
          if (net.ncol) {
            const float ctr = net.relative ? load_msb_norm(a.msb, net.msb_u16, off, maxv) : 0.f;
            float* d = dst + (size_t)(net.nco + (c * n + dy) * n) * LDP;
            const size_t rowoff = (plane + (reflect_clamp(gy + dy - D, net.H) - net.buf_row0)) * net.W;
            switch (n) {
              case 1: gather_row<1, LDP>(a.msb, net.msb_u16, rowoff, gx, net, maxv, ctr, ok, d); break;
              case 3: gather_row<3, LDP>(a.msb, net.msb_u16, rowoff, gx, net, maxv, ctr, ok, d); break;
              case 5: gather_row<5, LDP>(a.msb, net.msb_u16, rowoff, gx, net, maxv, ctr, ok, d); break;
              case 7: gather_row<7, LDP>(a.msb, net.msb_u16, rowoff, gx, net, maxv, ctr, ok, d); break;
              default:
                for (int dx = 0; dx < n; ++dx, d += LDP) {
                  float v = load_msb_norm(a.msb, net.msb_u16, rowoff + reflect_clamp(gx + dx - D, net.W), maxv) - ctr;
                  *d = ok ? v : 0.f;
                }
            }
          }
        }
      }
      __syncthreads();
      LBDRN_PHASE(23)  // labels, centres, window rows
      if (H2) {
        // features -> split 16-bit operand arrays (rows >= dim_in are the zero padding of the last k = 16 step)
        for (int i = tid; i < KP0 * (NPIX / 2); i += THREADS) {
          const int k = i / (NPIX / 2), pp = 2 * (i - k * (NPIX / 2));
          float2 x = make_float2(0.f, 0.f);
          if (k < net.dim_in) x = *reinterpret_cast<const float2*>(X + (size_t)k * LDP + pp);
          uint32_t hi, lo;
          split_h2(x.x, x.y, hi, lo);
          const uint32_t o = (uint32_t)(k * kLDH + pp) * 2u;
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(xh_hi + o), "r"(hi) : "memory");
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(xh_lo + o), "r"(lo) : "memory");
        }
        __syncthreads();
      }
      if (TC5) {
        t5_split_x();
        fence_async_smem();
        __syncthreads();
      }
      if (TCX) {
        tx_split_x();
        fence_async_smem();
        __syncthreads();
      }
      }
      LBDRN_PHASE(1)   // gather

      // ---- forward (LBDRNmodel.py:79-82), keeping h_l and act'(z_l) per layer ------------------------------
      float h[TM][TN];
      const int warp = tid >> 5, lane = tid & 31;
      if constexpr (TC5) {
        // ================= chunk forward + backward on tcgen05 (M = 64 pixels, accumulators in tensor memory) ==============
        // Every product is hi*hi + hi*lo + lo*hi of fp16 split operands accumulated in fp32 (as the MMA = 2 kernel), issued by
        // one elected lane; completion comes back through one mbarrier.  Thread <-> accumulator element as in a warp-level
        // m16n8 fragment: sub-partition sp = warp % 4 holds pixels 16 sp .. +15 (rows g, g + 8), column group cg = warp / 4
        // holds 16 columns (two 8-column tiles j; columns 2t, 2t + 1 of each).
        const int sp = warp & 3, cg = warp >> 2, g = lane >> 2, t = lane & 3;
        const uint32_t tm_lane = t5_tmem + ((uint32_t)(32 * sp) << 16);
        const uint32_t mbar = smem_u32(&s_t5_mbar);
        // the feature staging X shares its bytes with the hidden-output / dz images: put back the constant blocks behind the
        // hidden outputs (1, 0, .., 0 per pixel in hi; zero in lo) that X's rows ran over.  Every reader of X is past the CTA
        // barrier in front of this block; the first reader of a constant block is a weight-gradient GEMM, several fenced
        // barriers from here.
        for (int i = tid; i < 2 * L * NPIX; i += THREADS) {
          const int l = i / (2 * NPIX), r = i - l * 2 * NPIX, half = r / NPIX, pp = r - half * NPIX;
          const uint32_t addr = t5_h(l) + (uint32_t)half * t5_hbytes + (uint32_t)BC * NPIX * 2u + 16u * pp;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(addr), "r"(half == 0 ? 0x00003c00u : 0u), "r"(0u) : "memory");
        }
        auto mma3 = [&](uint32_t d_col, uint32_t a_img, uint32_t a_half, int a_rows, int a_mn, uint32_t b_img, uint32_t b_half,
                        int b_rows, int b_mn, int N, int ksteps) {
          const uint32_t idesc = umma_idesc_f16_major(64, N, a_mn, b_mn);
          for (int i = 0; i < ksteps; ++i) {       // small terms first
            umma_f16(t5_tmem + d_col, img_desc(a_img + a_half, a_rows, a_mn, i), img_desc(b_img, b_rows, b_mn, i), idesc, i > 0);
            umma_f16(t5_tmem + d_col, img_desc(a_img, a_rows, a_mn, i), img_desc(b_img + b_half, b_rows, b_mn, i), idesc, 1u);
            umma_f16(t5_tmem + d_col, img_desc(a_img, a_rows, a_mn, i), img_desc(b_img, b_rows, b_mn, i), idesc, 1u);
          }
        };
        auto t5_wait = [&]() {
          mbar_wait(mbar, t5_phase, 8, (int)blockIdx.x);
          t5_phase ^= 1u;
          tc_fence_after();
        };
        // ---- forward (LBDRNmodel.py:79-82) ----------------------------------------------------------------------------------
        for (int l = 0; l < L; ++l) {
          const int Kp = l == 0 ? KP0 : BC;
          if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
              mma3((uint32_t)(l & 1) * BC, l == 0 ? t5_x : t5_h(l - 1), l == 0 ? t5_xbytes : t5_hbytes, NPIX, 0, t5_w(l),
                   t5_wbytes(l), BC, 0, BC, Kp >> 4);
              umma_commit(mbar);
            }
            __syncwarp();
          }
          t5_wait();
          LBDRN_PHASE(15)   // fwd: GEMMs
          uint32_t r[8];
          tmem_ld16x256_x2(tm_lane + (uint32_t)(l & 1) * BC + 16u * cg, r);
          tmem_ld_wait8(r);
          float* Gl = Gbuf + (size_t)l * GSZ;
          float hv[2][4], gv[2][4], amax = 0.f;
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int u = 16 * cg + 8 * j + 2 * t + (e & 1);
              const float z = __uint_as_float(r[4 * j + e]) * kWInv + wsm[l * BC + u];
              if (net.relu) {
                hv[j][e] = fmaxf(z, 0.f);
                gv[j][e] = z > 0.f ? 1.f : 0.f;
              } else if (a.fast_sine) {
                // MUFU.SIN / MUFU.COS (opt-in, LBDRN_TRAIN_FASTSIN=1): the unit reduces the argument itself, abs. error
                // ~ 6e-8 |arg| + 4e-7 instead of 7e-8
                const float arg = net.w0 * z;
                hv[j][e] = __sinf(arg);
                gv[j][e] = __cosf(arg) * net.w0;
              } else {
                const float arg = net.w0 * z;
                float sn, cs;
                sincos_cw_core(arg, sn, cs);
                amax = fmaxf(amax, fabsf(arg));
                hv[j][e] = sn;
                gv[j][e] = cs * net.w0;
              }
            }
          if (__builtin_expect(amax > 48000.0f, 0)) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int u = 16 * cg + 8 * j + 2 * t + (e & 1);
                const float arg = net.w0 * (__uint_as_float(r[4 * j + e]) * kWInv + wsm[l * BC + u]);
                if (fabsf(arg) > 48000.0f) { hv[j][e] = sin_slow(arg); gv[j][e] = cos_slow(arg) * net.w0; }
              }
          }
          const uint32_t o_hi = t5_h(l), o_lo = o_hi + t5_hbytes;
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int pixel = 16 * sp + g + 8 * e2, u = 16 * cg + 8 * j + 2 * t;
              Gl[(size_t)u * LDP + pixel] = gv[j][2 * e2];              // bank = g + 8t: conflict-free
              Gl[(size_t)(u + 1) * LDP + pixel] = gv[j][2 * e2 + 1];
              uint32_t hi, lo;
              split_h2(hv[j][2 * e2], hv[j][2 * e2 + 1], hi, lo);
              const uint32_t o = (uint32_t)img_off(NPIX, pixel, u);    // 8 pixels x 16 contiguous bytes per store instruction
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(o_hi + o), "r"(hi) : "memory");
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(o_lo + o), "r"(lo) : "memory");
            }
          if (l == L - 1) {
            // output layer (LBDRNmodel.py:76-77) from registers: partial sums over this thread's 4 units of each of its 2 pixels,
            // then over the 4 lanes t of the pixel row (fixed order); lane t publishes band t (and t + 4) of its column group
            float part[2][CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) {
              part[0][c] = 0.f; part[1][c] = 0.f;
              if (c < C) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const float2 w2 = *reinterpret_cast<const float2*>(wo_p + c * BC + 16 * cg + 8 * j + 2 * t);
                  part[0][c] = fmaf(w2.x, hv[j][0], part[0][c]); part[0][c] = fmaf(w2.y, hv[j][1], part[0][c]);
                  part[1][c] = fmaf(w2.x, hv[j][2], part[1][c]); part[1][c] = fmaf(w2.y, hv[j][3], part[1][c]);
                }
              }
            }
#pragma unroll
            for (int c = 0; c < CP; ++c)
#pragma unroll
              for (int e2 = 0; e2 < 2; ++e2) {
                part[e2][c] += __shfl_xor_sync(0xffffffffu, part[e2][c], 1);
                part[e2][c] += __shfl_xor_sync(0xffffffffu, part[e2][c], 2);
              }
#pragma unroll
            for (int c = 0; c < CP; ++c)
              if ((c & 3) == t && c < C) {
                Pp5[(cg * CP + c) * LDP + 16 * sp + g] = part[0][c];
                Pp5[(cg * CP + c) * LDP + 16 * sp + g + 8] = part[1][c];
              }
          }
          fence_async_smem();
          tc_fence_before();
          __syncthreads();
          LBDRN_PHASE(16)   // fwd: bias + sine / cosine + stores
        }
        LBDRN_PHASE(2)
        // ---- output layer + loss (LBDRNloss.py:9): one thread per (band, pixel) sums the four column groups --------------------
        // d(mean((y-t)^2))/dz_o = gscale (y-t) y (1-y): in fp32 for dh_L and db_o; for dW_o = h_L^T . dz_o on the tensor core it is
        // stored as (y-t) y (1-y) 2^16 (|.| <= 2^14: fp16 range, no underflow of small errors), gscale 2^-16 comes off at read-out
        const float inv_o = gscale * (1.0f / 65536.0f);
        float sse = 0.f;
        if (tid < C * NPIX) {
          const int c = tid / NPIX, pp = tid - c * NPIX;
          const float z = (Pp5[(0 * CP + c) * LDP + pp] + Pp5[(1 * CP + c) * LDP + pp]) +
                          (Pp5[(2 * CP + c) * LDP + pp] + Pp5[(3 * CP + c) * LDP + pp]);
          const bool ok = s_valid[pp] != 0;
          const float y = sigmoidf_rn(z + bo_p[c]);
          const float d = y - Tl[c * LDP + pp];
          const float q = (1.0f - y) * y;
          dZo5[c * LDP + pp] = ok ? (gscale * d) * q : 0.f;                 // mse backward then sigmoid backward
          uint16_t hi, lo;
          split_h1(ok ? (d * q) * 65536.0f : 0.f, hi, lo);
          const uint32_t o = (uint32_t)img_off(NPIX, pp, c);
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(t5_dzo + o), "h"(hi) : "memory");
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(t5_dzo + t5_dzobytes + o), "h"(lo) : "memory");
          if (ok) sse = d * d;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, off);
        if (lane == 0) s_red[warp] = sse;
        if (tid < kMaxLayers) s_dzmax[tid] = 0u;
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
          float tsum = 0.f;
          for (int i = 0; i < THREADS / 32; ++i) tsum += s_red[i];
          s_sse += tsum;
        }
        if (tid < C) {
          const float gsum = row_sum<NPIX>(dZo5 + tid * LDP);
          float* dd = mypart + net.boff[L] + tid;
          *dd = first ? gsum : *dd + gsum;
        }
        LBDRN_PHASE(3)   // output layer + loss
        // ---- backward -----------------------------------------------------------------------------------------------------------
        // dW_o = h_L^T . dz_o (both images read MN-major) goes to the tensor core now and is collected with the last commit
        if (warp == 0) {
          tc_fence_after();
          if (elect_one()) mma3(T5_DWO, t5_h(L - 1), t5_hbytes, NPIX, 1, t5_dzo, t5_dzobytes, NPIX, 1, 16, NPIX >> 4);
          __syncwarp();
        }
        float inv_next = 1.f;
        for (int l = L - 1; l >= 0; --l) {
          const float* Gl = Gbuf + (size_t)l * GSZ;
          float dz[2][4], mx = 0.f;
          if (l == L - 1) {
            // dh_L[u][p] = sum_c W_o[c][u] dz_o[c][p] from registers (4 bands: not worth a round trip through the tensor core)
            float dh[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int e = 0; e < 4; ++e) dh[j][e] = 0.f;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
              if (c < C) {
                const float d0 = dZo5[c * LDP + 16 * sp + g], d1 = dZo5[c * LDP + 16 * sp + g + 8];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const float2 w2 = *reinterpret_cast<const float2*>(wo_p + c * BC + 16 * cg + 8 * j + 2 * t);
                  dh[j][0] = fmaf(w2.x, d0, dh[j][0]); dh[j][1] = fmaf(w2.y, d0, dh[j][1]);
                  dh[j][2] = fmaf(w2.x, d1, dh[j][2]); dh[j][3] = fmaf(w2.y, d1, dh[j][3]);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int pixel = 16 * sp + g + (e >> 1) * 8, u = 16 * cg + 8 * j + 2 * t + (e & 1);
                dz[j][e] = dh[j][e] * Gl[(size_t)u * LDP + pixel];
                mx = fmaxf(mx, fabsf(dz[j][e]));
              }
          } else {
            t5_wait();
            uint32_t r[8];
            tmem_ld16x256_x2(tm_lane + (uint32_t)(l & 1) * BC + 16u * cg, r);
            tmem_ld_wait8(r);
            const float sc = kWInv * inv_next;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int pixel = 16 * sp + g + (e >> 1) * 8, u = 16 * cg + 8 * j + 2 * t + (e & 1);
                dz[j][e] = (__uint_as_float(r[4 * j + e]) * sc) * Gl[(size_t)u * LDP + pixel];
                mx = fmaxf(mx, fabsf(dz[j][e]));
              }
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
          if (lane == 0) atomicMax(&s_dzmax[l], __float_as_uint(mx));     // order-independent: the step stays deterministic
          tc_fence_before();
          __syncthreads();
          LBDRN_PHASE(8)    // bwd: dh (MMA wait or registers) * act' + chunk maximum
          float S, inv;
          dz_scale(s_dzmax[l], S, inv);
          const uint32_t o_hi = t5_dz(l), o_lo = o_hi + t5_dzbytes;
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int pixel = 16 * sp + g + 8 * e2, u = 16 * cg + 8 * j + 2 * t;
              uint32_t hi, lo;
              split_h2(dz[j][2 * e2] * S, dz[j][2 * e2 + 1] * S, hi, lo);
              const uint32_t o = (uint32_t)img_off(NPIX, pixel, u);
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(o_hi + o), "r"(hi) : "memory");
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(o_lo + o), "r"(lo) : "memory");
            }
          if (tid == 0) s_t5_inv[l] = inv;
          fence_async_smem();
          __syncthreads();
          LBDRN_PHASE(9 + 2 * (l > 0 ? 1 : 0))     // bwd: dh + dz (9: layer 0, 11: layer >= 1)
          if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
              const int Kp = l == 0 ? KP0 : BC;
              if (l > 0) {
                // dh_{l-1} = dz_l . W_l (W_l image read MN-major)
                mma3((uint32_t)((l - 1) & 1) * BC, t5_dz(l), t5_dzbytes, NPIX, 0, t5_w(l), t5_wbytes(l), BC, 1, BC, BC >> 4);
                umma_commit(mbar);
              }
              // [dW_l | db_l] = dz_l^T . [in_l | 1] (both images read MN-major; the constant block follows the input's hi half)
              mma3(t5_dwcol(l), t5_dz(l), t5_dzbytes, NPIX, 1, l == 0 ? t5_x : t5_h(l - 1), l == 0 ? t5_xbytes : t5_hbytes, NPIX,
                   1, Kp + 8, NPIX >> 4);
              if (l == 0) umma_commit(mbar);          // covers every gradient GEMM issued before it
            }
            __syncwarp();
          }
          inv_next = inv;
        }
        // ---- gradient accumulators -> this CTA's partial (rows = units: sub-partition sp holds units 16 sp + g, + 8) ---------
        // every load of the thread is issued before the one wait (the read-out is latency, not bandwidth): layer 0's columns
        // in four 32-column slices (one per column group; columns >= KP0 + 8 of the last slice belong to the next accumulator
        // and are ignored), a later layer's 64 weight columns in 16-column slices, its bias column and W_o by column group 0
        t5_wait();
        LBDRN_PHASE(10)   // bwd: wait for the last weight-gradient GEMM
        {
          uint32_t r0[16], r1[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, rb[4] = {0u, 0u, 0u, 0u}, ro[4] = {0u, 0u, 0u, 0u};
          tmem_ld16x256_x4(tm_lane + t5_dwcol(0) + 32u * cg, r0);
          if (L > 1) tmem_ld16x256_x2(tm_lane + t5_dwcol(1) + 16u * cg, r1);
          if (cg == 0) {
            if (L > 1) tmem_ld16x256_x1(tm_lane + t5_dwcol(1) + (uint32_t)BC, rb);
            tmem_ld16x256_x1(tm_lane + T5_DWO, ro);
          }
          tmem_ld_wait16(r0);
          tmem_ld_wait8(r1);
          tmem_ld_wait4(rb);
          tmem_ld_wait4(ro);
          auto put_pair = [&](float* dst, int Kin, bool pair_ok, int u, int q, float v0, float v1) {
            if (pair_ok && q + 1 < Kin) {
              float2* d2 = reinterpret_cast<float2*>(dst + (size_t)u * Kin + q);
              float2 v = make_float2(v0, v1);
              if (!first) { const float2 o = *d2; v.x += o.x; v.y += o.y; }
              *d2 = v;
            } else {
              if (q < Kin) { float* d1 = dst + (size_t)u * Kin + q; *d1 = first ? v0 : *d1 + v0; }
              if (q + 1 < Kin) { float* d1 = dst + (size_t)u * Kin + q + 1; *d1 = first ? v1 : *d1 + v1; }
            }
          };
          auto put_bias = [&](int l, int u, float v) {
            float* dd = mypart + net.boff[l] + u;
            *dd = first ? v : *dd + v;
          };
          {
            const float inv = s_t5_inv[0];
            float* dst = mypart + net.woff[0];
            const bool pair_ok = (net.dim_in & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
            for (int base = 0; base < KP0 + 8; base += 128) {      // first layers wider than 120 inputs: further rounds of 128 columns
              if (base > 0) {
                tmem_ld16x256_x4(tm_lane + t5_dwcol(0) + (uint32_t)base + 32u * cg, r0);
                tmem_ld_wait16(r0);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e2 = 0; e2 < 2; ++e2) {
                  const int u = 16 * sp + g + 8 * e2, q = base + 32 * cg + 8 * j + 2 * t;
                  const float v0 = __uint_as_float(r0[4 * j + 2 * e2]) * inv, v1 = __uint_as_float(r0[4 * j + 2 * e2 + 1]) * inv;
                  if (q == KP0) put_bias(0, u, v0);                // the constant block's column: bias gradient
                  else put_pair(dst, net.dim_in, pair_ok, u, q, v0, v1);
                }
            }
          }
          if (L > 1) {
            const float inv = s_t5_inv[1];
            float* dst = mypart + net.woff[1];
            const bool pair_ok = (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int e2 = 0; e2 < 2; ++e2) {
                const int u = 16 * sp + g + 8 * e2, q = 16 * cg + 8 * j + 2 * t;
                put_pair(dst, BC, pair_ok, u, q, __uint_as_float(r1[4 * j + 2 * e2]) * inv, __uint_as_float(r1[4 * j + 2 * e2 + 1]) * inv);
              }
            if (cg == 0 && t == 0) {
              put_bias(1, 16 * sp + g, __uint_as_float(rb[0]) * inv);
              put_bias(1, 16 * sp + g + 8, __uint_as_float(rb[2]) * inv);
            }
          }
          if (cg == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int u = 16 * sp + g + (e >> 1) * 8, c = 2 * t + (e & 1);
              if (c < C) {
                float* d1 = mypart + net.woff[L] + c * BC + u;
                const float v = __uint_as_float(ro[e]) * inv_o;
                *d1 = first ? v : *d1 + v;
              }
            }
          }
        }
        // layers beyond the second (nl >= 3): tile by tile
        for (int l = 2; l < L; ++l) {
          const int ntile = (BC + 8) >> 3;
          const float inv = s_t5_inv[l];
          float* dst = mypart + net.woff[l];
          const bool pair_ok = (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
          for (int tt = cg; tt < ntile; tt += 4) {
            uint32_t r[4];
            tmem_ld16x256_x1(tm_lane + t5_dwcol(l) + 8u * tt, r);
            tmem_ld_wait4(r);
            const int q = 8 * tt + 2 * t;
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int u = 16 * sp + g + 8 * e2;
              const float v0 = __uint_as_float(r[2 * e2]) * inv, v1 = __uint_as_float(r[2 * e2 + 1]) * inv;
              if (q == BC) {
                float* dd = mypart + net.boff[l] + u;
                *dd = first ? v0 : *dd + v0;
              } else if (pair_ok) {
                float2* d2 = reinterpret_cast<float2*>(dst + (size_t)u * BC + q);
                float2 v = make_float2(v0, v1);
                if (!first) { const float2 o = *d2; v.x += o.x; v.y += o.y; }
                *d2 = v;
              } else {
                float* d1 = dst + (size_t)u * BC + q;
                d1[0] = first ? v0 : d1[0] + v0;
                d1[1] = first ? v1 : d1[1] + v1;
              }
            }
          }
        }
        tc_fence_before();       // the accumulators are read: the next chunk / step may overwrite them after its barrier
        first = false;
        LBDRN_PHASE(4)   // backward (remainder)
        continue;
      }

      if constexpr (TCX) {
        // ============ chunk forward + backward on tcgen05 for wide layers (bc 256): TRANSPOSED GEMMs ========================
        // z^T[unit][pixel] = W . x^T: the weights are the A operand (M = 128 units per block, full tensor rate), the 32-pixel
        // activation / gradient images the B operand (N = 32), so accumulators are 32 columns per block and all 16 warps own
        // accumulator rows: thread (sub-partition sp, block mb, pixel half ph) <-> unit u = 128 mb + 32 sp + lane, 16 pixels.
        // Hidden weights (475 KB as hi + lo at D = 3) cannot be resident: every GEMM that contracts with them streams one K
        // step (16 KB: hi 8 KB | lo 8 KB) at a time from the global image into a two-stage ring with per-thread cp.async and
        // issues its MMAs as the stages land; dh = W^T dz reads the same image transposed (the stage is gathered as 16-row
        // pieces and consumed through an MN-major descriptor).  Products are hi*hi + hi*lo + lo*hi in fp32 as in MMA = 3.
        constexpr int NB128 = BC / 128;
        const int sp = warp & 3, qd = warp >> 2, mb = qd / 2 % NB128, ph = qd & 1;
        const int un = 128 * mb + 32 * sp + lane;                       // this thread's unit (accumulator row)
        const uint32_t tm_lane = t5_tmem + ((uint32_t)(32 * sp) << 16);
        const uint32_t mbar = smem_u32(&s_t5_mbar);
        const uint32_t ring = smem_u32(smem4);
        constexpr uint32_t STG = 16384u, HALF = 8192u;
        static_assert(BC == 256, "stage geometry (one 16-byte piece per thread and half) is written for bc = 256");
        auto tx_wait = [&]() {
          mbar_wait(mbar, t5_phase, 8, (int)blockIdx.x);
          t5_phase ^= 1u;
          tc_fence_after();
        };
        auto mma3x = [&](uint32_t d_col, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, uint32_t acc) {
          umma_f16(t5_tmem + d_col, a_lo, b_hi, idesc, acc);        // small terms first
          umma_f16(t5_tmem + d_col, a_hi, b_lo, idesc, 1u);
          umma_f16(t5_tmem + d_col, a_hi, b_hi, idesc, 1u);
        };
        // D[slot][unit][pixel] = sum_k W-stage[unit][k] * B-image[pixel][k], nk K steps; TRANSPOSED: the stage holds 16 rows
        // (= K step) x BC columns of the weight image, i.e. W^T, read MN-major.  Ends with a commit to the chunk barrier.
        auto stream_gemm = [&](int nk, const uint8_t* w_hi, uint32_t w_half, bool transposed, uint32_t b_img, uint32_t b_half,
                               uint32_t d_col) {
          auto load_stage = [&](int ks, uint32_t stg) {
            // forward: K step ks of the K-major image = BC*32 contiguous bytes; transposed: rows 16 ks .. +15 of every 8-column
            // group (BC/8 pieces of 256 bytes); either way thread tid moves bytes [16 tid, 16 tid + 16) of each half
            const size_t src = transposed ? ((size_t)(tid >> 4) * BC + 16 * ks + (tid & 15)) * 16 : (size_t)ks * (BC * 32) + 16 * (size_t)tid;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stg + 16u * tid), "l"(w_hi + src) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stg + HALF + 16u * tid), "l"(w_hi + w_half + src) : "memory");
            cp_async_commit();
          };
          const uint32_t idesc = umma_idesc_f16_major(128, NPIX, transposed ? 1 : 0, 0);
          // four stages: two over the feature staging buffer, two over the image of the LAST hidden layer's output, which is
          // dead while a streamed GEMM runs (written by the last forward epilogue after its GEMM, last read by dW_o, which
          // the caller waits for before the transposed stream starts)
          auto stage_addr = [&](int st) { return st < 2 ? ring + (uint32_t)st * STG : tx_h(L - 1) + (uint32_t)(st - 2) * STG; };
          constexpr int NS = 4;
          for (int p = 0; p < NS && p < nk; ++p) load_stage(p, stage_addr(p));      // every stage is free at the start
          for (int ks = 0; ks < nk; ++ks) {
            const int st = ks & (NS - 1);
            // K step ks + 2 goes into the stage K step ks - 2 used: its MMAs were committed two iterations ago, so this wait
            // does not stall, and the copy has two iterations (> the L2 latency) to land
            if (ks >= 2 && ks + 2 < nk) {
              const int sr = (ks + 2) & (NS - 1);
              mbar_wait(smem_u32(&s_tx_free[sr]), (tx_fpar >> sr) & 1u, 10, (int)blockIdx.x);
              tx_fpar ^= 1u << sr;
              load_stage(ks + 2, stage_addr(sr));
            }
            // groups that may stay in flight behind K step ks: those of ks + 1, ks + 2 (the first two iterations: stricter)
            if (ks + 2 < nk) cp_async_wait_group<2>();
            else if (ks + 1 < nk) cp_async_wait_group<1>();
            else cp_async_wait_group<0>();
            fence_async_smem();
            __syncthreads();
            if (warp == 0) {
              tc_fence_after();
              if (elect_one()) {
                const uint32_t sa = stage_addr(st);
#pragma unroll
                for (int b = 0; b < NB128; ++b) {
                  const uint64_t a_hi = transposed ? umma_desc(sa + (uint32_t)b * 4096u, 128u, 256u) : umma_desc(sa + (uint32_t)b * 2048u, BC * 16u, 128u);
                  const uint64_t a_lo = transposed ? umma_desc(sa + HALF + (uint32_t)b * 4096u, 128u, 256u)
                                                   : umma_desc(sa + HALF + (uint32_t)b * 2048u, BC * 16u, 128u);
                  mma3x(d_col + 32u * b, a_hi, a_lo, img_desc(b_img, NPIX, 0, ks), img_desc(b_img + b_half, NPIX, 0, ks), idesc, ks > 0);
                }
                if (ks + NS < nk) umma_commit(smem_u32(&s_tx_free[st]));   // refilled with K step ks + NS at iteration ks + 2
                if (ks + 1 == nk) umma_commit(mbar);
              }
              __syncwarp();
            }
          }
        };
        // ---- forward (LBDRNmodel.py:79-82) ----------------------------------------------------------------------------------
        for (int l = 0; l < L; ++l) {
          const int Kp = l == 0 ? KP0 : BC;
          stream_gemm(Kp >> 4, tx_wimg(l), (uint32_t)Kp * BC * 2u, false, l == 0 ? tx_x : tx_h(l - 1), l == 0 ? tx_xbytes : tx_hbytes,
                      (uint32_t)(l & 1) * TX_SLOT);
          tx_wait();
          LBDRN_PHASE(15)   // fwd: GEMMs
          float acc[16];
          tmem_ld16(tm_lane + (uint32_t)(l & 1) * TX_SLOT + 32u * mb + 16u * ph, acc);
          float* Gl = tx_g(l);
          const float bias = wsm[l * BC + un];
          const uint32_t o_hi = tx_h(l), o_lo = o_hi + tx_hbytes;
          float amax = 0.f;
          float hv[16], gv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float z = acc[i] * kWInv + bias;
            if (net.relu) {
              hv[i] = fmaxf(z, 0.f);
              gv[i] = z > 0.f ? 1.f : 0.f;
            } else {
              const float arg = net.w0 * z;
              float sn, cs;
              sincos_cw_core(arg, sn, cs);
              amax = fmaxf(amax, fabsf(arg));
              hv[i] = sn;
              gv[i] = cs * net.w0;
            }
          }
          if (__builtin_expect(amax > 48000.0f, 0)) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float arg = net.w0 * (acc[i] * kWInv + bias);
              if (fabsf(arg) > 48000.0f) { hv[i] = sin_slow(arg); gv[i] = cos_slow(arg) * net.w0; }
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int pixel = 16 * ph + i;
            Gl[pixel * BC + un] = gv[i];                                // lanes = consecutive units: conflict-free
            uint16_t hi, lo;
            split_h1(hv[i], hi, lo);
            const uint32_t o = (uint32_t)img_off(NPIX, pixel, un);
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(o_hi + o), "h"(hi) : "memory");
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(o_lo + o), "h"(lo) : "memory");
          }
          fence_async_smem();
          tc_fence_before();
          __syncthreads();
          LBDRN_PHASE(16)   // fwd: bias + sine / cosine + stores
        }
        LBDRN_PHASE(2)
        // ---- output layer + loss (LBDRNloss.py:9) on the tensor core: z_o[pixel][c] = h_L . W_o^T, M = 64 (rows >= 32 unused),
        // N = 16 (rows >= C of the W_o image are zero) ---------------------------------------------------------------------------
        if (warp == 0) {
          tc_fence_after();
          if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16_major(64, 16, 0, 0);
            for (int ks = 0; ks < (BC >> 4); ++ks)
              mma3x(TX_ZO, img_desc(tx_h(L - 1), NPIX, 0, ks), img_desc(tx_h(L - 1) + tx_hbytes, NPIX, 0, ks),
                    img_desc(tx_wo, 16, 0, ks), img_desc(tx_wo + tx_wobytes, 16, 0, ks), idesc, ks > 0);
            umma_commit(mbar);
          }
          __syncwarp();
        }
        tx_wait();
        const float inv_o = gscale * (1.0f / 65536.0f);
        float sse = 0.f;
        float* const dZo5 = Pp;                                          // fp32 dz_o [CP][LDP] (the Pp slot of this variant)
        if (warp < 2) {                                                  // M = 64 accumulator: pixels 16 sp + g, + 8; bands 2t, 2t + 1
          const int g = lane >> 2, t = lane & 3;
          uint32_t r[4];
          tmem_ld16x256_x1(tm_lane + TX_ZO, r);
          tmem_ld_wait4(r);
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int pixel = 16 * sp + g + 8 * e2;
            const bool ok = s_valid[pixel] != 0;
            float v[2] = {0.f, 0.f};
#pragma unroll
            for (int e1 = 0; e1 < 2; ++e1) {
              const int c = 2 * t + e1;
              if (c < C) {
                const float y = sigmoidf_rn(__uint_as_float(r[2 * e2 + e1]) * kWInv + bo_p[c]);
                const float d = y - Tl[c * LDP + pixel];
                const float qq = (1.0f - y) * y;
                dZo5[c * LDP + pixel] = ok ? (gscale * d) * qq : 0.f;
                if (ok) { v[e1] = (d * qq) * 65536.0f; sse += d * d; }
              }
            }
            uint32_t hi, lo;
            split_h2(v[0], v[1], hi, lo);
            const uint32_t o = (uint32_t)img_off(NPIX, pixel, 2 * t);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(tx_dzo + o), "r"(hi) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(tx_dzo + tx_dzobytes + o), "r"(lo) : "memory");
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, off);
        if (lane == 0) s_red[warp] = sse;
        if (tid < kMaxLayers) s_dzmax[tid] = 0u;
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
          float tsum = 0.f;
          for (int i = 0; i < THREADS / 32; ++i) tsum += s_red[i];
          s_sse += tsum;
        }
        if (tid < C) {
          const float gsum = row_sum<NPIX>(dZo5 + tid * LDP);
          float* dd = mypart + net.boff[L] + tid;
          *dd = first ? gsum : *dd + gsum;
        }
        LBDRN_PHASE(3)   // output layer + loss
        // ---- backward: dW_o^T = h_L^T . dz_o first (small; it must have read h_L's image before the transposed stream reuses
        // that region as ring stages), then dh_L^T = W_o^T . dz_o^T (W_o image read MN-major, K = 16 bands); ONE commit for both:
        // two arrivals on the same barrier in quick succession could both land before a slow thread has observed the first,
        // and a parity wait that misses a phase never returns
        if (warp == 0) {
          tc_fence_after();
          if (elect_one()) {
            const uint32_t id1 = umma_idesc_f16_major(128, NPIX, 1, 0), id2 = umma_idesc_f16_major(128, 16, 1, 1);
            for (int ks = 0; ks < (NPIX >> 4); ++ks)
#pragma unroll
              for (int b = 0; b < NB128; ++b)
                mma3x(TX_DWO + 16u * b, umma_desc(tx_h(L - 1) + (uint32_t)b * (16u * NPIX * 16u) + (uint32_t)ks * 256u, 128u, NPIX * 16u),
                      umma_desc(tx_h(L - 1) + tx_hbytes + (uint32_t)b * (16u * NPIX * 16u) + (uint32_t)ks * 256u, 128u, NPIX * 16u),
                      img_desc(tx_dzo, NPIX, 1, ks), img_desc(tx_dzo + tx_dzobytes, NPIX, 1, ks), id2, ks > 0);
#pragma unroll
            for (int b = 0; b < NB128; ++b)
              mma3x((uint32_t)((L - 1) & 1) * TX_SLOT + 32u * b, umma_desc(tx_wo + (uint32_t)b * 4096u, 128u, 256u),
                    umma_desc(tx_wo + tx_wobytes + (uint32_t)b * 4096u, 128u, 256u), img_desc(tx_dzo, NPIX, 0, 0),
                    img_desc(tx_dzo + tx_dzobytes, NPIX, 0, 0), id1, 0u);
            umma_commit(mbar);
          }
          __syncwarp();
        }
        float inv_next = inv_o;
        for (int l = L - 1; l >= 0; --l) {
          tx_wait();
          float acc[16];
          tmem_ld16(tm_lane + (uint32_t)(l & 1) * TX_SLOT + 32u * mb + 16u * ph, acc);
          const float* Gl = tx_g(l);
          const float sc = kWInv * inv_next;
          float dz[16], mx = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            dz[i] = (acc[i] * sc) * Gl[(16 * ph + i) * BC + un];
            mx = fmaxf(mx, fabsf(dz[i]));
          }
          // db_l[u] = sum over the chunk's pixels of dz_l: this thread's 16 pixels in a fixed order, the two halves below
          s_txdb[ph][un] = (((dz[0] + dz[1]) + (dz[2] + dz[3])) + ((dz[4] + dz[5]) + (dz[6] + dz[7]))) +
                           (((dz[8] + dz[9]) + (dz[10] + dz[11])) + ((dz[12] + dz[13]) + (dz[14] + dz[15])));
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
          if (lane == 0) atomicMax(&s_dzmax[l], __float_as_uint(mx));
          tc_fence_before();
          __syncthreads();                                               // every act' of the layer has been read: overwrite
          if (tid < BC) {
            const float dbv = s_txdb[0][tid] + s_txdb[1][tid];
            float* dd = mypart + net.boff[l] + tid;
            if (first) *dd = dbv; else atomicAdd(dd, dbv);
          }
          LBDRN_PHASE(8)    // bwd: dh * act' + chunk maximum
          float S, inv;
          dz_scale(s_dzmax[l], S, inv);
          const uint32_t o_hi = tx_dz(l), o_lo = o_hi + tx_dzbytes;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            uint16_t hi, lo;
            split_h1(dz[i] * S, hi, lo);
            const uint32_t o = (uint32_t)img_off(NPIX, 16 * ph + i, un);
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(o_hi + o), "h"(hi) : "memory");
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(o_lo + o), "h"(lo) : "memory");
          }
          if (tid == 0) s_t5_inv[l] = inv;
          fence_async_smem();
          __syncthreads();
          LBDRN_PHASE(9 + 2 * (l > 0 ? 1 : 0))     // bwd: dz image
          if (l > 0)      // dh_{l-1}^T = W_l^T . dz_l^T: the weight image streamed transposed
            stream_gemm(BC >> 4, tx_wimg(l), (uint32_t)BC * BC * 2u, true, tx_dz(l), tx_dzbytes, (uint32_t)((l - 1) & 1) * TX_SLOT);
          inv_next = inv;
        }
        // ---- weight gradients: per layer and block of 128 INPUT features one pass D[k][unit] = in_l^T . dz_l (both images read
        // MN-major, contraction over the chunk's 32 pixels, N = all BC units) into TMEM columns TX_DWP..; rows beyond the layer's
        // K multiply whatever follows the image and are never read.  Thread <-> input feature k (the accumulator row), so a warp
        // store writes 32 consecutive k of one unit's gradient row: 128 contiguous bytes (the [unit][k] orientation wrote 8 rows
        // x 32 bytes per instruction: 9k cycles per pass, bound by L1 wavefronts) ---------------------------------------------
        {
          const int g = lane >> 2, t = lane & 3;
          for (int l = L - 1; l >= 0; --l) {
            const int Kin = l == 0 ? net.dim_in : BC, Kp = l == 0 ? KP0 : BC, nblk = (Kp + 127) >> 7;
            const uint32_t in_img = l == 0 ? tx_x : tx_h(l - 1), in_half = l == 0 ? tx_xbytes : tx_hbytes;
            const float inv = s_t5_inv[l];
            float* dst = mypart + net.woff[l];
            for (int b = 0; b < nblk; ++b) {
              if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
                  const uint32_t idw = umma_idesc_f16_major(128, BC, 1, 1);
                  for (int ks = 0; ks < (NPIX >> 4); ++ks)
                    mma3x(TX_DWP, umma_desc(in_img + (uint32_t)b * (16u * NPIX * 16u) + (uint32_t)ks * 256u, 128u, NPIX * 16u),
                          umma_desc(in_img + in_half + (uint32_t)b * (16u * NPIX * 16u) + (uint32_t)ks * 256u, 128u, NPIX * 16u),
                          img_desc(tx_dz(l), NPIX, 1, ks), img_desc(tx_dz(l) + tx_dzbytes, NPIX, 1, ks), idw, ks > 0);
                  umma_commit(mbar);
                }
                __syncwarp();
              }
              tx_wait();
              // warp (sp, qd): k = 128 b + 32 sp + lane, units 64 qd .. + 63.  A later chunk of the same step ADDS to what this
              // thread stored for the first one with reductions that return nothing (the same thread owns the address in
              // every chunk, so the additions happen in program order and the sum stays deterministic)
              const int k = 128 * b + 32 * sp + lane;
              float* dk = dst + k;
#pragma unroll 1
              for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t r0[16], r1[16];
                tmem_ld16_issue(tm_lane + TX_DWP + 64u * qd + c0, r0);
                tmem_ld16_issue(tm_lane + TX_DWP + 64u * qd + c0 + 16u, r1);
                tmem_ld16_wait(r0);
                tmem_ld16_wait(r1);
                if (k < Kin) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    float* d1 = dk + (size_t)(64 * qd + c0 + j) * Kin;
                    const float v = __uint_as_float(r0[j]) * inv;
                    if (first) *d1 = v; else atomicAdd(d1, v);
                  }
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    float* d1 = dk + (size_t)(64 * qd + c0 + 16 + j) * Kin;
                    const float v = __uint_as_float(r1[j]) * inv;
                    if (first) *d1 = v; else atomicAdd(d1, v);
                  }
                }
              }
              tc_fence_before();
              __syncthreads();                                           // the pass' columns are read: the next pass may overwrite
            }
          }
          // W_o gradient (issued long ago, complete by any later commit): D[unit][c]
          if (qd < NB128) {
            for (int h = 0; h < 2; ++h) {
              uint32_t r[4];
              tmem_ld16x256_x1(tm_lane + ((uint32_t)(16 * h) << 16) + TX_DWO + 16u * qd, r);
              tmem_ld_wait4(r);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int u = 128 * qd + 32 * sp + 16 * h + g + (e >> 1) * 8, c = 2 * t + (e & 1);
                if (c < C) {
                  float* d1 = mypart + net.woff[L] + c * BC + u;
                  const float v = __uint_as_float(r[e]) * inv_o;
                  if (first) *d1 = v; else atomicAdd(d1, v);
                }
              }
            }
          }
        }
        tc_fence_before();
        first = false;
        LBDRN_PHASE(4)   // backward (remainder)
        continue;
      }
      for (int l = 0; l < L; ++l) {
        const int K = l == 0 ? net.dim_in : BC;
        const float* in = l == 0 ? X : Hbuf + (size_t)(l - 1) * BC * LDP;
        float* Hl = Hbuf + (size_t)l * BC * LDP;
        float* Gl = Gbuf + (size_t)l * GSZ;
        if constexpr (H2) {
          constexpr int NT = BC / 32;
          float acc[NT][4];
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
          const int Kp = l == 0 ? KP0 : BC, ldw = ldw_of(l);
          const uint32_t in_hi = l == 0 ? xh_hi : hh_base + (uint32_t)(l - 1) * BC * kLDH * 4u;
          const uint32_t in_lo = in_hi + (uint32_t)Kp * kLDH * 2u;
          const uint32_t w_hi = wh_of(l), w_lo = w_hi + (uint32_t)BC * ldw * 2u;
          gemm_px_unit_h2<NT, false>(acc, in_hi, in_lo, w_hi, w_lo, ldw, Kp >> 4, warp, lane);
          LBDRN_PHASE(15)   // fwd: GEMMs
          const int g = lane >> 2, t = lane & 3;
          const uint32_t out_hi = hh_base + (uint32_t)l * BC * kLDH * 4u, out_lo = out_hi + (uint32_t)BC * kLDH * 2u;
          float hv[NT][4], gv[NT][4], amax = 0.f;
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int u = 8 * ((warp >> 2) + 4 * j) + 2 * t + (e & 1);
              const float z = acc[j][e] * kWInv + wsm[l * BC + u];      // the power-of-two weight scale comes off exactly
              if (net.relu) {
                hv[j][e] = fmaxf(z, 0.f);
                gv[j][e] = z > 0.f ? 1.f : 0.f;
              } else {
                const float arg = net.w0 * z;
                float sn, cs;
                sincos_cw_core(arg, sn, cs);
                amax = fmaxf(amax, fabsf(arg));
                hv[j][e] = sn;
                gv[j][e] = cs * net.w0;
              }
            }
          if (__builtin_expect(amax > 48000.0f, 0)) {               // one large-argument test per thread, not per value
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int u = 8 * ((warp >> 2) + 4 * j) + 2 * t + (e & 1);
                const float arg = net.w0 * (acc[j][e] * kWInv + wsm[l * BC + u]);
                if (fabsf(arg) > 48000.0f) { hv[j][e] = sin_slow(arg); gv[j][e] = cos_slow(arg) * net.w0; }
              }
          }
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {                          // e = 2 e2, 2 e2 + 1: units u, u+1 of one pixel
              const int pixel = 16 * (warp & 3) + g + e2 * 8, u = 8 * ((warp >> 2) + 4 * j) + 2 * t;
              Gl[(size_t)u * LDP + pixel] = gv[j][2 * e2];              // bank = g + 8t: conflict-free
              Gl[(size_t)(u + 1) * LDP + pixel] = gv[j][2 * e2 + 1];
              if (l == L - 1) {
                Hbuf[(size_t)u * LDP + pixel] = hv[j][2 * e2];
                Hbuf[(size_t)(u + 1) * LDP + pixel] = hv[j][2 * e2 + 1];
              } else {
                uint32_t hi, lo;
                split_h2(hv[j][2 * e2], hv[j][2 * e2 + 1], hi, lo);
                const uint32_t o = (uint32_t)(u * kLDH + pixel) * 2u;
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(out_hi + o), "h"((uint16_t)(hi & 0xffffu)) : "memory");
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(out_hi + o + kLDH * 2u), "h"((uint16_t)(hi >> 16)) : "memory");
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(out_lo + o), "h"((uint16_t)(lo & 0xffffu)) : "memory");
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(out_lo + o + kLDH * 2u), "h"((uint16_t)(lo >> 16)) : "memory");
              }
            }
          LBDRN_PHASE(16)   // fwd: bias + sine / cosine + stores
          __syncthreads();
          continue;
        }
        if (MMA) {
          constexpr int NT = BC / 8 / NGW;
          float acc[NT][4];
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
          gemm_px_unit_mma<BC, LDP, NT, PT>(acc, in, w + net.woff[l], K, warp, lane);
          LBDRN_PHASE(15)   // fwd: GEMMs
          const int g = lane >> 2, t = lane & 3;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int pixel = 16 * (warp % PT) + g + (e >> 1) * 8, u = 8 * ((warp / PT) + NGW * j) + 2 * t + (e & 1);
              const float z = acc[j][e] + w[net.boff[l] + u];
              float hv, gv;
              if (net.relu) {
                hv = fmaxf(z, 0.f);
                gv = z > 0.f ? 1.f : 0.f;
              } else {
                float sn, cs;
                sincos_cw(net.w0 * z, sn, cs);
                hv = sn;
                gv = cs * net.w0;
              }
              Hl[(size_t)u * LDP + pixel] = hv;          // bank = g + 8t: conflict-free
              Gl[(size_t)u * LDP + pixel] = gv;
            }
          }
          LBDRN_PHASE(16)   // fwd: bias + sine / cosine + stores
          __syncthreads();
          continue;
        }
        float acc[TM][TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          float b = w[net.boff[l] + unit(j)];
#pragma unroll
          for (int i = 0; i < TM; ++i) acc[i][j] = b;
        }
        gemm_kmajor_v<TM, TN, VEC, BC>(acc, in, w + net.woff[l], K, pg, ubase);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          float g4[TM];
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            if (net.relu) {
              h[i][j] = fmaxf(acc[i][j], 0.f);
              g4[i] = acc[i][j] > 0.f ? 1.f : 0.f;
            } else {
              float sn, cs;
              sincos_cw(net.w0 * acc[i][j], sn, cs);   // d sin(w0 z)/dz = w0 cos(w0 z)
              h[i][j] = sn;
              g4[i] = cs * net.w0;
            }
          }
          size_t o = (size_t)unit(j) * LDP + pg * TM;
          float hv[TM];
#pragma unroll
          for (int i = 0; i < TM; ++i) hv[i] = h[i][j];
          store_vec<TM>(Hl + o, hv);
          store_vec<TM>(Gl + o, g4);
        }
        layer_sync();
      }

      LBDRN_PHASE(2)   // hidden layers forward
      // ---- output layer + loss (LBDRNloss.py:9) ---------------------------------------------------------
      const float* wo = wo_p;
      float sse = 0.f;
      if (MMA) {
        // z[c][p] = b[c] + sum_u W_o[c][u] h_L[u][p] from shared memory: one thread per (band, pixel)
        const float* HL = Hbuf + (size_t)(H2 ? 0 : L - 1) * BC * LDP;
        if (tid < C * NPIX) {
          const int c = tid / NPIX, pp = tid - c * NPIX;
          float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;
#pragma unroll 4
          for (int u = 0; u < BC; u += 4) {
            z0 = fmaf(wo[c * BC + u], HL[(size_t)u * LDP + pp], z0);
            z1 = fmaf(wo[c * BC + u + 1], HL[(size_t)(u + 1) * LDP + pp], z1);
            z2 = fmaf(wo[c * BC + u + 2], HL[(size_t)(u + 2) * LDP + pp], z2);
            z3 = fmaf(wo[c * BC + u + 3], HL[(size_t)(u + 3) * LDP + pp], z3);
          }
          const bool ok = s_valid[pp] != 0;
          const float y = sigmoidf_rn(((z0 + z1) + (z2 + z3)) + bo_p[c]);
          const float d = y - Tl[c * LDP + pp];
          dZo[c * LDP + pp] = ok ? (gscale * d) * ((1.0f - y) * y) : 0.f;   // mse backward then sigmoid backward
          if (ok) sse = d * d;
        }
      }
      float part[TM * CP];
      if (!MMA) {
#pragma unroll
      for (int c = 0; c < CP; ++c) {
#pragma unroll
        for (int i = 0; i < TM; ++i) part[i * CP + c] = 0.f;
        if (c < C) {
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            float wv = wo[c * BC + unit(j)];
#pragma unroll
            for (int i = 0; i < TM; ++i) part[i * CP + c] = fmaf(wv, h[i][j], part[i * CP + c]);
          }
        }
      }
      group8_allreduce<TM * CP>(part);
      if (US > 1) {
        // cross-split sum through smem: lane tn == i publishes pixel i of its group
#pragma unroll
        for (int i = 0; i < TM; ++i)
          if (i == tn)
#pragma unroll
            for (int c = 0; c < CP; ++c) Pp[(us * CP + c) * LDP + pg * TM + i] = part[i * CP + c];
        __syncthreads();
      }
      }
      if (!MMA && us == 0) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          if (i == tn) {
            int pp = pg * TM + i;
            bool ok = s_valid[pp] != 0;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
              if (c < C) {
                float z = part[i * CP + c];
                if (US > 1) {
                  z = 0.f;
#pragma unroll
                  for (int u2 = 0; u2 < US; ++u2) z += Pp[(u2 * CP + c) * LDP + pp];
                }
                float y = sigmoidf_rn(z + bo_p[c]);
                float d = y - Tl[c * LDP + pp];
                float dz = ok ? (gscale * d) * ((1.0f - y) * y) : 0.f;   // mse backward then sigmoid backward
                dZo[c * LDP + pp] = dz;
                if (ok) sse += d * d;
              }
            }
          }
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, off);
      if ((tid & 31) == 0) s_red[tid >> 5] = sse;
      __syncthreads();
      if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < THREADS / 32; ++i) t += s_red[i];
        s_sse += t;
      }

      LBDRN_PHASE(3)   // output layer + loss
      // ---- backward --------------------------------------------------------------------------------------
      // output layer: dW_o[c][n] = sum_p dz_o[c][p] h_L[n][p];  db_o[c] = sum_p dz_o[c][p]
      {
        const float* HL = Hbuf + (size_t)(H2 ? 0 : L - 1) * BC * LDP;
        for (int o = tid; o < C * BC; o += THREADS) {
          int c = o / BC, u = o - c * BC;
          float g = row_dot<NPIX>(dZo + c * LDP, HL + (size_t)u * LDP);
          float* d = mypart + net.woff[L] + o;
          *d = first ? g : *d + g;
        }
        if (tid < C) {
          float g = row_sum<NPIX>(dZo + tid * LDP);
          float* d = mypart + net.boff[L] + tid;
          *d = first ? g : *d + g;
        }
      }
      LBDRN_PHASE(8)    // bwd: output-layer grads
      float inv_next = 1.f;                                            // H2: 1 / scale of dz_{l+1}
      for (int l = L - 1; l >= 0; --l) {
        float* Gl = Gbuf + (size_t)l * GSZ;
        if constexpr (H2) {
          const uint32_t g_hi = smem_u32(Gl), g_lo = g_hi + (uint32_t)BC * kLDH * 2u;
          constexpr int NT = BC / 32, UPT = BC / 64;
          const int g = lane >> 2, t = lane & 3, pc = tid & 7;
          float dzo[UPT][8];          // l == L-1: thread = (unit (tid >> 3) + 64 i, pixels 4 pc .. +3 and 32 + 4 pc .. +3)
          float dzr[NT][4];           // l <  L-1: the accumulator fragment positions
          float mx = 0.f;
          if (l == L - 1) {
            // dh_L[u][p] = sum_c W_o[c][u] dz_o[c][p];  dz_L = dh_L * act'(z_L);  db_L[u] = sum_p dz_L[u][p]
#pragma unroll
            for (int i = 0; i < UPT; ++i) {
              const int u = (tid >> 3) + 64 * i;
              const float4 a0 = *reinterpret_cast<const float4*>(Gl + (size_t)u * LDP + 4 * pc);
              const float4 a1 = *reinterpret_cast<const float4*>(Gl + (size_t)u * LDP + 32 + 4 * pc);
              float d[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) d[k] = 0.f;
#pragma unroll
              for (int c = 0; c < CP; ++c) {
                if (c < C) {
                  const float wv = wo[c * BC + u];
                  const float4 v0 = *reinterpret_cast<const float4*>(dZo + c * LDP + 4 * pc);
                  const float4 v1 = *reinterpret_cast<const float4*>(dZo + c * LDP + 32 + 4 * pc);
                  d[0] = fmaf(wv, v0.x, d[0]); d[1] = fmaf(wv, v0.y, d[1]); d[2] = fmaf(wv, v0.z, d[2]); d[3] = fmaf(wv, v0.w, d[3]);
                  d[4] = fmaf(wv, v1.x, d[4]); d[5] = fmaf(wv, v1.y, d[5]); d[6] = fmaf(wv, v1.z, d[6]); d[7] = fmaf(wv, v1.w, d[7]);
                }
              }
              dzo[i][0] = a0.x * d[0]; dzo[i][1] = a0.y * d[1]; dzo[i][2] = a0.z * d[2]; dzo[i][3] = a0.w * d[3];
              dzo[i][4] = a1.x * d[4]; dzo[i][5] = a1.y * d[5]; dzo[i][6] = a1.z * d[6]; dzo[i][7] = a1.w * d[7];
#pragma unroll
              for (int k = 0; k < 8; ++k) mx = fmaxf(mx, fabsf(dzo[i][k]));
              float sum = ((dzo[i][0] + dzo[i][1]) + (dzo[i][2] + dzo[i][3])) + ((dzo[i][4] + dzo[i][5]) + (dzo[i][6] + dzo[i][7]));
              sum += __shfl_xor_sync(0xffffffffu, sum, 1);
              sum += __shfl_xor_sync(0xffffffffu, sum, 2);
              sum += __shfl_xor_sync(0xffffffffu, sum, 4);
              if (pc == 0) {
                float* dd = mypart + net.boff[l] + u;
                *dd = first ? sum : *dd + sum;
              }
            }
          } else {
            // dh_l[u][p] = sum_m W_{l+1}[m][u] dz_{l+1}[m][p] on the tensor cores; dz_l = dh_l * act'(z_l) at the fragment positions
            float dacc[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
              for (int e = 0; e < 4; ++e) dacc[j][e] = 0.f;
            const uint32_t gn_hi = smem_u32(Gbuf + (size_t)(l + 1) * GSZ), gn_lo = gn_hi + (uint32_t)BC * kLDH * 2u;
            const int ldw = ldw_of(l + 1);
            const uint32_t w_hi = wh_of(l + 1), w_lo = w_hi + (uint32_t)BC * ldw * 2u;
            gemm_px_unit_h2<NT, true>(dacc, gn_hi, gn_lo, w_hi, w_lo, ldw, BC >> 4, warp, lane);
            const float sc = kWInv * inv_next;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int pixel = 16 * (warp & 3) + g + (e >> 1) * 8, u = 8 * ((warp >> 2) + 4 * j) + 2 * t + (e & 1);
                dzr[j][e] = (dacc[j][e] * sc) * Gl[(size_t)u * LDP + pixel];
                mx = fmaxf(mx, fabsf(dzr[j][e]));
              }
              // bias-gradient partial of this warp's 16 pixels for units 2t, 2t+1 of tile j: fixed-order tree over g
              float s0 = dzr[j][0] + dzr[j][2], s1 = dzr[j][1] + dzr[j][3];
#pragma unroll
              for (int off = 4; off < 32; off <<= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, off);
                s1 += __shfl_xor_sync(0xffffffffu, s1, off);
              }
              if (g == 0) {
                const int u = 8 * ((warp >> 2) + 4 * j) + 2 * t;
                s_db[l & 1][warp & 3][u] = s0;
                s_db[l & 1][warp & 3][u + 1] = s1;
              }
            }
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
          if (lane == 0) atomicMax(&s_dzmax[l], __float_as_uint(mx));     // non-negative floats order like their bit patterns
          __syncthreads();                                               // every act' of the layer has been read: overwrite
          float S, inv;
          dz_scale(s_dzmax[l], S, inv);
          if (l == L - 1) {
#pragma unroll
            for (int i = 0; i < UPT; ++i) {
              const int u = (tid >> 3) + 64 * i;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t hi0, lo0, hi1, lo1;
                split_h2(dzo[i][4 * h] * S, dzo[i][4 * h + 1] * S, hi0, lo0);
                split_h2(dzo[i][4 * h + 2] * S, dzo[i][4 * h + 3] * S, hi1, lo1);
                const uint32_t o = (uint32_t)(u * kLDH + 32 * h + 4 * pc) * 2u;
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(g_hi + o), "r"(hi0), "r"(hi1) : "memory");
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(g_lo + o), "r"(lo0), "r"(lo1) : "memory");
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
              for (int e2 = 0; e2 < 2; ++e2) {
                const int pixel = 16 * (warp & 3) + g + e2 * 8, u = 8 * ((warp >> 2) + 4 * j) + 2 * t;
                uint32_t hi, lo;
                split_h2(dzr[j][2 * e2] * S, dzr[j][2 * e2 + 1] * S, hi, lo);
                const uint32_t o = (uint32_t)(u * kLDH + pixel) * 2u;
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(g_hi + o), "h"((uint16_t)(hi & 0xffffu)) : "memory");
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(g_hi + o + kLDH * 2u), "h"((uint16_t)(hi >> 16)) : "memory");
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(g_lo + o), "h"((uint16_t)(lo & 0xffffu)) : "memory");
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(g_lo + o + kLDH * 2u), "h"((uint16_t)(lo >> 16)) : "memory");
              }
          }
          __syncthreads();
          LBDRN_PHASE(9 + 2 * (l > 0 ? 1 : 0))     // bwd: dh + dz (9: layer 0, 11: layer >= 1)
          // dW_l = dz_l . in_l^T on the tensor cores
          const int Kp = l == 0 ? KP0 : BC;
          const uint32_t in_hi = l == 0 ? xh_hi : hh_base + (uint32_t)(l - 1) * BC * kLDH * 4u;
          const uint32_t in_lo = in_hi + (uint32_t)Kp * kLDH * 2u;
          grad_weight_h2<BC>(g_hi, g_lo, in_hi, in_lo, l == 0 ? net.dim_in : BC, Kp >> 3, inv, mypart + net.woff[l], first,
                             warp, lane);
          if (l < L - 1)
            for (int u = tid; u < BC; u += THREADS) {
              const float sum = (s_db[l & 1][0][u] + s_db[l & 1][1][u]) + (s_db[l & 1][2][u] + s_db[l & 1][3][u]);
              float* dd = mypart + net.boff[l] + u;
              *dd = first ? sum : *dd + sum;
            }
          inv_next = inv;
          LBDRN_PHASE(10 + 2 * (l > 0 ? 1 : 0))    // bwd: dW + db (10: layer 0, 12: layer >= 1)
          continue;
        }
        float acc[TM][TN];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
        if (l == L - 1) {
          // dh_L[u][p] = sum_c W_o[c][u] dz_o[c][p]
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            if (c < C) {
              float dv[TM];
              load_vec<TM>(dZo + c * LDP + pg * TM, dv);
#pragma unroll
              for (int j = 0; j < TN; ++j) {
                float wv = wo[c * BC + unit(j)];
#pragma unroll
                for (int i = 0; i < TM; ++i) acc[i][j] = fmaf(wv, dv[i], acc[i][j]);
              }
            }
          }
        } else if (MMA) {
          // dh_l[u][p] = sum_m W_{l+1}[m][u] dz_{l+1}[m][p] on the tensor cores; dz_l = dh_l * act'(z_l) at the fragment positions
          constexpr int NT = BC / 8 / NGW;
          float dacc[NT][4];
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) dacc[j][e] = 0.f;
          gemm_px_unit_mma<BC, LDP, NT, PT>(dacc, Gbuf + (size_t)(l + 1) * GSZ, wnat(l + 1), BC, warp, lane);
          const int g = lane >> 2, t = lane & 3;
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int pixel = 16 * (warp % PT) + g + (e >> 1) * 8, u = 8 * ((warp / PT) + NGW * j) + 2 * t + (e & 1);
              Gl[(size_t)u * LDP + pixel] *= dacc[j][e];
            }
        } else {
          // dh_l[u][p] = sum_m W_{l+1}[m][u] dz_{l+1}[m][p]   (k-major in m on both operands)
          gemm_kmajor_v<TM, TN, VEC, BC>(acc, Gbuf + (size_t)(l + 1) * GSZ, wnat(l + 1), BC, pg, ubase);
        }
        // dz_l = dh_l * act'(z_l): thread-private read-modify-write of its own (unit, pixel) entries
        if (!(MMA && l < L - 1))
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          float* gp = Gl + (size_t)unit(j) * LDP + pg * TM;
          float g[TM];
          load_vec<TM>(gp, g);
#pragma unroll
          for (int i = 0; i < TM; ++i) g[i] *= acc[i][j];
          store_vec<TM>(gp, g);
        }
        __syncthreads();
        LBDRN_PHASE(9 + 2 * (l > 0 ? 1 : 0))     // bwd: dh + dz (9: layer 0, 11: layer >= 1)
        // dW_l = dz_l . in_l^T ; db_l = row sums
        const int K = l == 0 ? net.dim_in : BC;
        const float* in = l == 0 ? X : Hbuf + (size_t)(l - 1) * BC * LDP;
        if (MMA) grad_weight_mma<BC, LDP, NPIX>(Gl, in, K, l == 0 ? a.dimpad : BC, mypart + net.woff[l], first, warp, lane);
        else grad_weight_nt_t<BC, THREADS, TM>(Gl, in, K, l == 0 ? a.dimpad : BC, mypart + net.woff[l], first);
        for (int u = tid; u < BC; u += THREADS) {
          float g = row_sum<NPIX>(Gl + (size_t)u * LDP);
          float* d = mypart + net.boff[l] + u;
          *d = first ? g : *d + g;
        }
        LBDRN_PHASE(10 + 2 * (l > 0 ? 1 : 0))    // bwd: dW + db (10: layer 0, 12: layer >= 1)
      }
      first = false;
      LBDRN_PHASE(4)   // backward (remainder)
    }
    __syncthreads();
    if (tid == 0) {
      if (!first) mypart[P] = s_sse;
      gbar_arrive(a.gbar);                   // first barrier, arrive: this CTA's gradient partial is complete
    }
    // next step's neighbourhood loads, between the two halves of the barrier: in flight while the other CTAs finish and
    // during the reduction / Adam phase (itself L2-latency-bound).  Only thread 0 fences, before its own loads are issued
    // (a membar waits for the thread's outstanding loads).
    if (il) prefetch_issue_il(); else prefetch_issue();
    LBDRN_PHASE(14)   // issue of the next step's neighbourhood loads
    if (tid == 0) gbar_wait(a.gbar, (unsigned)(s + 1) * gridDim.x);
    __syncthreads();
    LBDRN_PHASE(5)    // first barrier, wait

    // ---- fixed-order reduction over the CTAs that produced partials, then Adam ----------------------------
    const int n_act = min((int)gridDim.x, n_chunks);
    {
      // Every partial row is read as whole 128-byte lines (8 lanes x float4 = 32 consecutive parameters of one row) and ALL
      // loads of the phase are in flight at once: the CTA owns `per_cta` float4 columns; 16 warps = 4 column blocks of 8 x
      // 16 row groups, a thread sums rows rg, rg + 16, ... (10 loads: up to 160 rows) of its column, the 16 row-group sums of a
      // column meet in shared memory (over the feature buffer, dead here) and thread (column, component) adds them in a fixed
      // order and owns that parameter's Adam update.  History: 4 lanes per parameter with 32-bit loads touched four lines per
      // warp instruction and used a quarter of each (7.3k cycles, bound by the L1 wavefront queue); 3 warps x 2 batches of 16
      // float4 loads were bound by latency x bytes in flight (5k).  Deterministic: the order of every sum is fixed by the code.
      const int n4 = (P + 4) >> 2;                                    // float4 columns of a partial row (P + 1 floats)
      const int per_cta = (((n4 + (int)gridDim.x - 1) / (int)gridDim.x) + 7) & ~7;   // multiple of 8: whole lines
      const size_t ps4 = (size_t)(a.pstride >> 2);
      constexpr int NCB = 4, RW = 16, NRG = 16, NCOL = 8 * NCB;       // 4 column blocks of 8 x 4 groups of 4 row groups = 16 warps
      constexpr int NLD = 10;                                         // rows per thread in one batch: 160 rows (grids up to 148)
      static_assert(THREADS / 32 >= RW, "reduction layout needs 16 warps");
      float4* const rbuf = reinterpret_cast<float4*>(smem4);          // [NRG][NCOL]
      const int r_lane = tid & 31, r_warp = tid >> 5;
      const int cblk = r_warp % NCB, rg = (r_warp / NCB) * 4 + (r_lane >> 3);
      const bool loader = r_warp < RW;
      for (int cb = 0; cb < per_cta; cb += NCOL) {
        // owner of (column cb + tid / 4, component tid % 4): its parameter's state is requested first
        const int oc = cb + (tid >> 2), i = 4 * (blockIdx.x * per_cta + oc) + (tid & 3);
        const bool in = tid < 4 * NCOL && oc < per_cta && i <= P;
        const bool upd = in && a.mode != TRAIN_GRAD_ONLY && i < P;
        float p_old = 0.f, m_old = 0.f, v_old = 0.f;
        if (upd) { p_old = __ldcg(a.params + i); m_old = __ldcg(a.m + i); v_old = __ldcg(a.v + i); }
        if (loader) {
          const int col = cb + 8 * cblk + (r_lane & 7), i4 = blockIdx.x * per_cta + col;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col < per_cta && i4 < n4) {
            // plain loads: ordered after the other CTAs' stores by the acquire in gbar_wait + the CTA barrier behind it
            const float4* pp = reinterpret_cast<const float4*>(a.partial) + i4;
            float4 v[NLD];
#pragma unroll
            for (int j = 0; j < NLD; ++j) {
              const int r = rg + NRG * j;
              v[j] = r < n_act ? pp[(size_t)r * ps4] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            acc.x = (((v[0].x + v[1].x) + (v[2].x + v[3].x)) + ((v[4].x + v[5].x) + (v[6].x + v[7].x))) + (v[8].x + v[9].x);
            acc.y = (((v[0].y + v[1].y) + (v[2].y + v[3].y)) + ((v[4].y + v[5].y) + (v[6].y + v[7].y))) + (v[8].y + v[9].y);
            acc.z = (((v[0].z + v[1].z) + (v[2].z + v[3].z)) + ((v[4].z + v[5].z) + (v[6].z + v[7].z))) + (v[8].z + v[9].z);
            acc.w = (((v[0].w + v[1].w) + (v[2].w + v[3].w)) + ((v[4].w + v[5].w) + (v[6].w + v[7].w))) + (v[8].w + v[9].w);
            for (int r = rg + NRG * NLD; r < n_act; r += NRG) {        // more rows than NLD per group
              const float4 t4 = pp[(size_t)r * ps4];
              acc.x += t4.x; acc.y += t4.y; acc.z += t4.z; acc.w += t4.w;
            }
          }
          rbuf[rg * NCOL + 8 * cblk + (r_lane & 7)] = acc;
        }
        __syncthreads();
        LBDRN_PHASE(17)   // reduce: row loads + exchange
        float g = 0.f;
        if (tid < 4 * NCOL) {
          const float* rb = reinterpret_cast<const float*>(rbuf) + tid;       // column tid / 4, component tid % 4
          float t[NRG];
#pragma unroll
          for (int q = 0; q < NRG; ++q) t[q] = rb[q * 4 * NCOL];
#pragma unroll
          for (int w = 1; w < NRG; w <<= 1)
#pragma unroll
            for (int q = 0; q < NRG; q += 2 * w) t[q] += t[q + w];
          g = t[0];
        }
        LBDRN_PHASE(18)   // reduce: row-group sums
        if (in) {
          if (a.mode == TRAIN_GRAD_ONLY) {
            a.grad_out[i] = g;
          } else if (i == P) {
            a.losses[s] = g / ((float)B * (float)C);
          } else {
            float p = p_old, m = m_old, v = v_old;
            adam_update(p, m, v, g, a.omb1, a.omb2, a.beta2f, a.eps, s_adam[0], s_adam[1]);
            a.params[i] = p; a.m[i] = m; a.v[i] = v;
            if (TC5 || TCX) {
              if (a.wimg != nullptr) {
                uint32_t img = 0u;                                    // halves before layer l's block
                bool done = false;
                for (int l = 0; l < L && !done; ++l) {
                  const int K = l == 0 ? net.dim_in : BC, Kp = l == 0 ? KP0 : BC, o = i - net.woff[l];
                  if (o >= 0 && o < K * BC) {
                    const int u = o / K, k = o - u * K;
                    uint16_t hi, lo;
                    split_h1(p * kWScale, hi, lo);
                    const uint32_t e = img + (uint32_t)(img_off(BC, u, k) >> 1);
                    a.wimg[e] = hi;
                    a.wimg[e + (uint32_t)Kp * BC] = lo;
                    done = true;
                  }
                  img += 2u * (uint32_t)Kp * BC;
                }
                if (TCX && !done) {
                  const int oo = i - net.woff[L];
                  if (oo >= 0 && oo < C * BC) {                       // W_o image, 16 rows
                    const int c = oo / BC, u = oo - c * BC;
                    uint16_t hi, lo;
                    split_h1(p * kWScale, hi, lo);
                    const uint32_t e = img + (uint32_t)(img_off(16, c, u) >> 1);
                    a.wimg[e] = hi;
                    a.wimg[e + 16u * BC] = lo;
                  }
                }
              }
            } else if (H2) {
              if (a.wimg != nullptr) {
                uint32_t img = 0u;                                    // halves before layer l's block
                for (int l = 0; l < L; ++l) {
                  const int K = l == 0 ? net.dim_in : BC, ldw = ldw_of(l), o = i - net.woff[l];
                  if (o >= 0 && o < K * BC) {
                    const int u = o / K, k = o - u * K;
                    uint16_t hi, lo;
                    split_h1(p * kWScale, hi, lo);
                    a.wimg[img + (uint32_t)(u * ldw + k)] = hi;
                    a.wimg[img + (uint32_t)(BC * ldw + u * ldw + k)] = lo;
                    break;
                  }
                  img += 2u * BC * ldw;
                }
              }
            } else {
              a.wpack[packed_index(net, i)] = p;
            }
          }
        }
        if (cb + NCOL < per_cta) __syncthreads();    // the exchange buffer is reused by the next column block
      }
    }
    LBDRN_PHASE(6)   // reduce + Adam
    __syncthreads();
    LBDRN_PHASE(19)  // slowest warp of the CTA in reduce + Adam
    if (tid == 0) gbar_arrive(a.gbar + 1);   // second barrier, arrive; the wait is at the head of the next step
    LBDRN_PHASE(20)  // fence + arrive 2
  }
  if constexpr (TC5 || TCX) {
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t5_tmem), "r"(512) : "memory");
  }
#undef LBDRN_PHASE
}

// Adam on an externally reduced gradient (data-parallel mode) + packed-copy refresh.
static __global__ void adam_apply_kernel(Net net, const float* __restrict__ grad, float* params, float* wpack, float* m,
                                  float* v, float omb1, float omb2, float beta2, float eps, float step_size,
                                  float bc2_sqrt) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < net.P; i += gridDim.x * blockDim.x) {
    float p = params[i], mm = m[i], vv = v[i];
    adam_update(p, mm, vv, grad[i], omb1, omb2, beta2, eps, step_size, bc2_sqrt);
    params[i] = p; m[i] = mm; v[i] = vv;
    wpack[packed_index(net, i)] = p;
  }
}

}  // namespace lbdrn
