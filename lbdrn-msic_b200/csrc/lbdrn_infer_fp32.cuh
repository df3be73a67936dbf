// lbdrn_infer_fp32.cuh -- "PRECISE" inference kernels: fused tile+halo staging, on-the-fly features, fp32 FFMA MLP,
// sigmoid, inverse quantisation and integer write (decode), or raw network output (predict), or squared error
// against the LSB labels (eval).  Replaces reference decode.py:77-134 / encode.py:105-108 on the device.
//
// One persistent CTA (128 threads) loops over TH x 16 pixel tiles:
//   1. stage the normalised (tile + D halo) of every band in smem (each MSB byte is read ~once from HBM/L2)
//   2. expand the dim_in feature rows of the tile into act[k][pixel]            (never leaves the SM)
//   3. hidden layers: register-tiled GEMM act x W^T (k-major operands, LDS.128), sin(w0 .) written back in place
//   4. output layer from registers + 8-lane shuffle reduction, sigmoid, epilogue
// Roofline: FFMA issue-bound (2*(dim_in*bc+(nl-1)bc^2+bc*C) flop/pixel on the fp32 pipe); HBM traffic is the
// algorithmic C*(sizeof(msb)+2) bytes/pixel.
#pragma once
#include "lbdrn_common.cuh"

namespace lbdrn {

enum { MODE_DECODE = 0, MODE_PREDICT = 1, MODE_SSE = 2 };

struct InferArgs {
  Net net;
  const void* msb;
  const void* lsb;       // MODE_SSE only
  const float* wpack;    // packed parameters: hidden W transposed to [K][bc], everything else as in `params`
  const float* tab;      // coordinate tables or nullptr
  void* out;             // uint16 image (DECODE) / float y (PREDICT)
  double* partials;      // MODE_SSE: [gridDim.x]
  unsigned int* counter; // MODE_SSE: zero on entry, zero on exit
  double* sse_out;       // MODE_SSE
  int tiles_x, n_tiles;
  int kmax;              // max(dim_in, bc): rows of the activation buffer
  const int* skip_flag;  // optional device flag: when non-null and non-zero the kernel exits at once (the tensor-core
                         // kernel launched before it already decoded the scene)
};

// Transpose hidden-layer weights W_l [bc][K_l] -> [K_l][bc]; copy biases and the output layer unchanged.
static __global__ void pack_params_kernel(Net net, const float* __restrict__ params, float* __restrict__ wpack) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < net.P; i += gridDim.x * blockDim.x) {
    int dst = i;
    for (int l = 0; l < net.nl; ++l) {
      int K = l == 0 ? net.dim_in : net.bc;
      int o = i - net.woff[l];
      if (o >= 0 && o < K * net.bc) {
        int n = o / K, k = o - n * K;
        dst = net.woff[l] + k * net.bc + n;
      }
    }
    wpack[dst] = params[i];
  }
}

template <int TM>
__host__ __device__ constexpr int ldp_of() { return 16 * TM + 4; }

template <int BC, int TM, int CP, bool WSMEM, int MODE>
__global__ void __launch_bounds__(kThreads) infer_fp32_kernel(const InferArgs a) {
  constexpr int NPIX = 16 * TM, LDP = ldp_of<TM>(), TH = TM, TW = 16, TN = BC / 8;
  const Net& net = a.net;
  const int tid = threadIdx.x, tn = tid & 7, pg = tid >> 3;
  const int D = net.D, n = net.n, C = net.C;
  const int trows = TH + 2 * D, twp = TW + 2 * D;

  if (a.skip_flag != nullptr && *a.skip_flag != 0) return;
  const float maxv = net_maxv(net);

  extern __shared__ float4 smem4[];
  float* act = reinterpret_cast<float*>(smem4);
  float* tile = act + (size_t)a.kmax * LDP;
  float* wsm = tile + round4(C * trows * twp);
  __shared__ double red[kThreads / 32];

  if (WSMEM) {
    const int P4 = net.P >> 2;
    for (int i = tid; i < P4; i += kThreads)
      reinterpret_cast<float4*>(wsm)[i] = reinterpret_cast<const float4*>(a.wpack)[i];
    for (int i = (P4 << 2) + tid; i < net.P; i += kThreads) wsm[i] = a.wpack[i];
  }
  const float* w = WSMEM ? wsm : a.wpack;
  double sse_local = 0.0;

  for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const int ty0 = net.row0 + (t / a.tiles_x) * TH, tx0 = (t % a.tiles_x) * TW;

    // ---- 1. normalised tile + halo (reflect at true image borders only); one warp per tile row ---------------
    if (net.ncol) {
      const int lane = tid & 31, warp = tid >> 5;
      for (int c = 0; c < C; ++c) {
        for (int r = warp; r < trows; r += kThreads / 32) {
          const int gy = reflect_clamp(ty0 - D + r, net.H);
          const size_t rowoff = ((size_t)c * net.buf_rows + (gy - net.buf_row0)) * net.W;
          for (int x = lane; x < twp; x += 32)
            tile[(c * trows + r) * twp + x] =
                load_msb_norm(a.msb, net.msb_u16, rowoff + reflect_clamp(tx0 - D + x, net.W), maxv);
        }
      }
    }
    __syncthreads();   // also orders the previous tile's compute before act[] is overwritten below

    // ---- 2. feature rows (LBDRNdataset.py:104-130): [coords | colours(c, dy, dx)]; thread = (pixel, band share) ----
    {
      constexpr int SHARE = kThreads / NPIX;               // threads cooperating on one pixel (1 or 2)
      const int p = tid % NPIX, share = tid / NPIX;
      const int r = p >> 4, x = p & 15;
      float* dst = act + p;
      if (net.nco && share == 0) {
        const int gy = min(ty0 + r, net.H - 1), gx = min(tx0 + x, net.W - 1);
        const float* trow = a.tab + (size_t)gy * net.tabw;
        const float* tcol = a.tab + (size_t)(net.H + gx) * net.tabw;
        for (int i = 0; i < net.tabw; ++i) {
          dst[(size_t)i * LDP] = trow[i];
          dst[(size_t)(net.tabw + i) * LDP] = tcol[i];
        }
      }
      if (net.ncol) {
        for (int c = share; c < C; c += SHARE) {
          const float* tc = tile + (c * trows + r) * twp + x;
          const float ctr = net.relative ? tc[D * twp + D] : 0.f;
          float* d = dst + (size_t)(net.nco + c * n * n) * LDP;
          for (int dy = 0; dy < n; ++dy)
            for (int dx = 0; dx < n; ++dx, d += LDP) *d = tc[dy * twp + dx] - ctr;
        }
      }
    }
    __syncthreads();

    // ---- 3. hidden layers ---------------------------------------------------------------------------
    float h[TM][TN];
    for (int l = 0; l < net.nl; ++l) {
      const int K = l == 0 ? net.dim_in : BC;
      float acc[TM][TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        float b = w[net.boff[l] + unit_of<TN>(j, tn)];
#pragma unroll
        for (int i = 0; i < TM; ++i) acc[i][j] = b;
      }
      gemm_kmajor<TM, TN, BC>(acc, act, LDP, w + net.woff[l], K, pg, tn);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) h[i][j] = net.relu ? fmaxf(acc[i][j], 0.f) : act_sine(acc[i][j], net.w0);
      if (l + 1 < net.nl) {
        __syncwarp();   // a pixel group's columns are private to its warp: warp-level ordering is enough
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          float* dst = act + (size_t)unit_of<TN>(j, tn) * LDP + pg * TM;
#pragma unroll
          for (int i = 0; i < TM; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(h[i][j], h[i + 1][j], h[i + 2][j], h[i + 3][j]);
        }
        __syncwarp();
      }
    }

    // ---- 4. output layer: partial dot over this lane's units, reduced across the 8 lanes of the group ----
    float part[TM * CP];
    const float* wo = w + net.woff[net.nl];
#pragma unroll
    for (int c = 0; c < CP; ++c) {
#pragma unroll
      for (int i = 0; i < TM; ++i) part[i * CP + c] = 0.f;
      if (c < C) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          float wv = wo[c * BC + unit_of<TN>(j, tn)];
#pragma unroll
          for (int i = 0; i < TM; ++i) part[i * CP + c] = fmaf(wv, h[i][j], part[i * CP + c]);
        }
      }
    }
    group8_allreduce<TM * CP>(part);

    // lane tn finishes pixel i == tn of its group
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      if (i == tn) {
        int p = pg * TM + i;
        int gy = ty0 + (p >> 4), gx = tx0 + (p & 15);
        if (gy < net.row1 && gx < net.W) {
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            if (c < C) {
              float y = sigmoidf_rn(part[i * CP + c] + w[net.boff[net.nl] + c]);
              size_t off = ((size_t)c * net.buf_rows + (gy - net.buf_row0)) * net.W + gx;
              if (MODE == MODE_DECODE) {
                // residual = round_half_even(y*(2^K-1)); image = (base << K) + residual  (decode.py:131-134)
                uint32_t m = load_msb_int(a.msb, net.msb_u16, off);
                int res = (int)rintf(y * net.qmax);
                reinterpret_cast<uint16_t*>(a.out)[off] = (uint16_t)((m << net.K) + (uint32_t)res);
              } else if (MODE == MODE_PREDICT) {
                reinterpret_cast<float*>(a.out)[((size_t)(gy - net.row0) * net.W + gx) * C + c] = y;
              } else {
                uint32_t code = net.lsb_u16 ? (uint32_t)((const uint16_t*)a.lsb)[off] : (uint32_t)((const uint8_t*)a.lsb)[off];
                float d = y - __fdiv_rn((float)code, net.qmax);   // label = LSB/(2^K-1)  (LBDRNdataset.py:97)
                sse_local += (double)(d * d);
              }
            }
          }
        }
      }
    }
  }

  if (MODE == MODE_SSE) {
    // deterministic reduction: lanes -> warp -> CTA partial -> last CTA sums partials in index order
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sse_local += __shfl_xor_sync(0xffffffffu, sse_local, off);
    if ((tid & 31) == 0) red[tid >> 5] = sse_local;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int i = 0; i < kThreads / 32; ++i) s += red[i];
      a.partials[blockIdx.x] = s;
      __threadfence();
      unsigned int done = atomicAdd(a.counter, 1u);
      if (done == gridDim.x - 1) {
        __threadfence();
        double tot = 0.0;
        for (unsigned int i = 0; i < gridDim.x; ++i) tot += ((volatile double*)a.partials)[i];
        *a.sse_out = tot;
        *a.counter = 0u;
      }
    }
  }
}

}  // namespace lbdrn
