/* lbdrn.h -- C ABI of liblbdrn_b200.so: the B200 (sm_100a) implementation of LBDRN's per-pixel
 * bit-depth-recovery hot path.
 *
 * This is the drop-in boundary.  The reference (lidq92/LBDRN-MSIC) is pure Python and has no FFI of its own;
 * each entry point below replaces the reference code cited next to it (paths relative to the reference repo),
 * and `INTEGRATION.md` shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ or torch types cross the boundary.
 *   - every `*_dev` pointer is DEVICE memory owned by the caller (e.g. a torch tensor's data_ptr());
 *     images are contiguous CHW planes; parameters are one flat float32 vector in the reference's
 *     `state_dict` order (net.{i}.linear.weight [bc,in] row-major, net.{i}.linear.bias [bc], ...,
 *     last_layer.linear.weight [C,bc], last_layer.linear.bias [C]) -- the layout `encode.py:123-128`
 *     flattens and `decode.py:114-120` slices.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls are asynchronous on that
 *     stream unless stated; the library never synchronises except in lbdrn_train_create/_destroy.
 *   - return value: 0 (LBDRN_OK) or a negative LBDRN_E_* code; the message of the last failure on the
 *     calling thread is available from lbdrn_last_error().  No exception crosses the boundary.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns LBDRN_E_CUDA.
 *   - concurrency: lbdrn_decode / lbdrn_predict / lbdrn_eval_sse keep a small per-device scratch (packed weights, the
 *     partials of the squared-error reduction, a ring of TMA descriptors) that is NOT keyed by stream: on one device,
 *     queue these calls on ONE stream at a time (a LbdrnTrain handle owns its own state and may run on another stream
 *     concurrently -- FusedTrainer trains on one stream and evaluates on a second one).
 *   - hidden widths: the kernels are built for bc = 32, 64, 128, 256; any other width returns LBDRN_E_UNSUPPORTED
 *     (the .bin header can describe any power of two: such streams are refused, not decoded wrongly).
 */
#ifndef LBDRN_H_
#define LBDRN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: optional device-side msb_max in LbdrnDesc; 3: lbdrn_randperm added; 4: LBDRN_PATH_TENSOR_FASTSIN2 and the
 * lbdrn_fpz_* nn sub-stream codec added; 5: lbdrn_selftest_tc_gemm3, lbdrn_host_randperm[32]; 6: lbdrn_host_randperm32_progress (all additive: older callers
 * are unaffected) */
#define LBDRN_ABI_VERSION 6

enum {
  LBDRN_OK = 0,
  LBDRN_E_INVALID = -1,      /* bad argument / inconsistent descriptor */
  LBDRN_E_UNSUPPORTED = -2,  /* configuration outside what the kernels are built for */
  LBDRN_E_CUDA = -3,         /* CUDA runtime error (message in lbdrn_last_error) */
  LBDRN_E_NOMEM = -4
};

/* feature-set flags: the module globals of reference constants.py:3-14 */
enum {
  LBDRN_USE_COORDINATES = 1u << 0,
  LBDRN_EMBEDDING = 1u << 1,
  LBDRN_USE_COLORS = 1u << 2,
  LBDRN_RELATIVE = 1u << 3,
  LBDRN_ACT_RELU = 1u << 4   /* hidden activation ReLU instead of Sine (the commented alternative at encode.py:75) */
};

enum { LBDRN_U8 = 0, LBDRN_U16 = 1 };

/* kernel selection for lbdrn_decode: PRECISE = fp32 FFMA everywhere (closest to the reference's fp32 path);
 * TENSOR = tcgen05 split-precision tensor-core path where the configuration supports it. AUTO picks TENSOR
 * when available, else PRECISE. */
enum { LBDRN_PATH_AUTO = 0, LBDRN_PATH_PRECISE = 1, LBDRN_PATH_TENSOR = 2,
       LBDRN_PATH_TENSOR_FASTSIN = 3,  /* TENSOR with MUFU.SIN after an exact range reduction (abs err ~4e-7) */
       LBDRN_PATH_TENSOR_FASTSIN2 = 4  /* TENSOR with MUFU.SIN alone: the unit reduces the argument itself after one
                                          multiplication by 1/(2 pi) rounded toward zero (adds <= |w0 z| * 1.2e-7) */ };

/* One scene (or one row stripe of it) + network shape.  Rows are GLOBAL image rows: a rank that decodes
 * stripe [row0,row1) passes a buffer holding rows [buf_row0, buf_row0+buf_rows) which must cover
 * [row0-D, row1+D) clipped to the image (reflect padding is applied only at true image borders). */
typedef struct LbdrnDesc {
  int32_t C, H, W;        /* bands, full image height, width */
  int32_t K, D;           /* dropped LSBs, neighbourhood radius */
  int32_t bc, nl;         /* hidden width, hidden layers */
  uint32_t flags;         /* LBDRN_USE_* | LBDRN_RELATIVE | LBDRN_ACT_RELU */
  float w0;               /* 30.0 in the reference (LBDRNmodel.py:58-59) */
  int32_t n_freq;         /* N_FREQ (12); coordinate table width per axis = 2*n_freq*EMBEDDING + 1 */
  uint32_t msb_max;       /* global max of the MSB image over all bands/pixels (LBDRNdataset.py:120) */
  int32_t msb_dtype;      /* LBDRN_U8 or LBDRN_U16 (LBDRNdataset.py:100) */
  int32_t row0, row1;     /* output rows [row0,row1) */
  int32_t buf_row0;       /* first image row present in the msb/lsb/out buffers */
  int32_t buf_rows;       /* rows present in the buffers (plane stride = buf_rows*W) */
  int32_t path;           /* LBDRN_PATH_* */
  int32_t reserved[3];
  const uint32_t* msb_max_dev; /* optional DEVICE word holding MSB.max(): when non-NULL the kernels read the normaliser
                                  from it (no host round trip after a device-side reduction / all-reduce) and `msb_max`
                                  only has to be an upper bound (it still selects uint8/uint16-safe kernels) */
} LbdrnDesc;

/* Optimiser / loop settings of the encoder (encode.py:84-85, -lr -bs -e flags). */
typedef struct LbdrnTrainCfg {
  int32_t batch_size;     /* -bs */
  int32_t world_size;     /* data-parallel ranks sharing every batch (1 = single GPU) */
  int32_t rank;
  int32_t reserved0;
  double beta1, beta2, eps; /* Adam: 0.9, 0.999, 1e-8 (torch defaults used by encode.py:84); doubles like torch's
                               python floats, so 1-beta rounds to fp32 exactly as in torch/optim/adam.py */
  int32_t reserved[4];
} LbdrnTrainCfg;

typedef struct LbdrnTrain LbdrnTrain; /* opaque: Adam state, gradient partials, staging */

/* ---- introspection ------------------------------------------------------------------------------------ */
int32_t lbdrn_version(void);
const char* lbdrn_last_error(void);
/* dim_in / parameter count of the network the descriptor implies (LBDRNdataset.py:104-106, LBDRNmodel.py:62-77) */
int32_t lbdrn_dim_in(const LbdrnDesc* d);
int64_t lbdrn_param_count(const LbdrnDesc* d);
/* 1 if lbdrn_decode would run the tcgen05 tensor-core kernel for this descriptor, else 0 */
int32_t lbdrn_has_tensor_path(const LbdrnDesc* d);
/* number of kernel launches the library has issued on this process since load (for bench accounting) */
int64_t lbdrn_launch_count(void);

/* ---- a1: MSB/LSB split on device (LBDRNdataset.py:93-100) ------------------------------------------------
 * img_dev: CHW uint16 [C*H*W]; msb_dev: uint8 or uint16 per msb_dtype; lsb_dev: integer LSB codes, uint8 if
 * K<=8 else uint16 (label = code/(2^K-1) is formed in-kernel).  n = C*H*W. */
int32_t lbdrn_split(const uint16_t* img_dev, int64_t n, int32_t K, int32_t msb_dtype, void* msb_dev,
                    void* lsb_dev, void* stream);
/* global max of a uint16 CHW image shifted right by K (so the caller can pick msb_dtype / msb_max);
 * result written to *max_dev (device uint32, must be zero-initialised by the caller). */
int32_t lbdrn_max_shifted(const uint16_t* img_dev, int64_t n, int32_t K, uint32_t* max_dev, void* stream);

/* ---- a10: batch sampling on the device (encode.py:69-70: one uniform random permutation of the N pixels per epoch) ------
 * out_dev[i], i < n, is a pseudo-random permutation of 0..n-1 chosen by `seed`: a 6-round Feistel network over the
 * ceil(log2 n) index bits (murmur3-finalizer round function keyed from the seed by splitmix64) with cycle walking for
 * indices that land >= n.  A bijection by construction (every pixel exactly once per epoch, like the DataLoader's
 * RandomSampler); the ORDER is not torch's -- callers that need the reference's exact batches pass the DataLoader's own
 * permutation to lbdrn_train_steps instead.  8 bytes written per element, no sort, no scratch memory. */
int32_t lbdrn_randperm(int64_t n, uint64_t seed, int64_t* out_dev, void* stream);

/* The REFERENCE's batch order (encode.py:69-70: DataLoader(shuffle=True) -> RandomSampler -> torch.randperm(n, generator=g)
 * with g.manual_seed(seed)), written to HOST memory: bit-identical to torch's CPU randperm for n < 2^32/20 (forward
 * Fisher-Yates on MT19937 outputs; larger n return LBDRN_E_UNSUPPORTED -- use torch), about 3x faster because the draws run
 * ahead of the swaps and the lines they will touch are prefetched.  Host-only; thread-safe; no device is needed. */
int32_t lbdrn_host_randperm(int64_t n, uint64_t seed, int64_t* out_host);
int32_t lbdrn_host_randperm32(int64_t n, uint64_t seed, int32_t* out_host);   /* same order as 32-bit indices (half the traffic) */
/* The 32-bit order with its progress published (ABI 6): *progress_host (aligned int64 in host memory, written with
 * release stores about every 2^18 steps) = number of LEADING entries of out_host that are final, n when the call
 * returns.  The shuffle is a forward Fisher-Yates, so a consumer on another thread may read out_host[0 .. *progress_host)
 * while the call is still running: the trainer starts the first epoch on the head of its order. */
int32_t lbdrn_host_randperm32_progress(int64_t n, uint64_t seed, int32_t* out_host, int64_t* progress_host);

/* ---- a16: quality read-out (decode.py:216): *sse_dev (device uint64, zero-initialised by the caller) += sum over n
 * elements of (a-b)^2 for two uint16 images.  Integer accumulation: exact and order-independent. */
int32_t lbdrn_sse_u16(const uint16_t* a_dev, const uint16_t* b_dev, int64_t n, uint64_t* sse_dev, void* stream);

/* ---- a14+a15: fused decode (decode.py:77-134) -----------------------------------------------------------
 * msb_dev: base layer, CHW per desc (rows buf_row0..).  params_dev: flat fp32 [P].  coord_tab_dev: float32
 * [(H + W) * tabw] row table then column table, tabw = 2*n_freq*EMBEDDING+1, or NULL when USE_COORDINATES is
 * off.  out_dev: uint16 CHW with the same row window / plane stride as the msb buffer; rows [row0,row1) are
 * written: out = (base << K) + round_half_even(sigmoid(...) * (2^K-1)). */
int32_t lbdrn_decode(const LbdrnDesc* d, const void* msb_dev, const float* params_dev,
                     const float* coord_tab_dev, uint16_t* out_dev, void* stream);

/* Diagnostic: D[128][64] (fp32) = A[128][K] * B[64][K]^T with fp16 row-major A, B through the same tcgen05 descriptors,
 * shared-memory operand layout and TMEM read-back as the tensor decode kernel (K multiple of 16, <= 256). */
int32_t lbdrn_selftest_tc_gemm(const void* a_dev, const void* b_dev, float* d_dev, int32_t K, void* stream);

/* Diagnostic: D[128][N] = A[128][K] * B[N][K]^T (fp16 row-major inputs) with operand A and/or B staged in the MN-major
 * "[group of 8][k][8]" shared-memory layout and consumed through MN-major UMMA descriptors (a_mn / b_mn != 0). */
int32_t lbdrn_selftest_tc_gemm2(const void* a_dev, const void* b_dev, float* d_dev, int32_t N, int32_t K, int32_t a_mn,
                                int32_t b_mn, void* stream);

/* Diagnostic: D[64][N] = A . B^T with the M = 64 instruction shape of the fused training step (modified_ignite_engine.py:18-27:
 * its chunk GEMMs), operands passed as ready-made fp16 "images" (element (r, k) of a [rows][K] array at byte
 * ((k>>3)*rows + r)*16 + (k&7)*2) and read through the K-major (x_mn = 0) or the MN-major (x_mn = 1: the image's k index is
 * the M / N index, its row index is contracted) descriptor; ksteps = contraction length / 16.  raw_dev (optional): all 128
 * TMEM lanes x N columns. */
int32_t lbdrn_selftest_tc_gemm3(const void* a_img_dev, int32_t a_bytes, const void* b_img_dev, int32_t b_bytes, float* d_dev,
                                float* raw_dev, int32_t N, int32_t ksteps, int32_t a_mn, int32_t a_rows, int32_t b_mn,
                                int32_t b_rows, void* stream);

/* network output y [n_rows*W, C] float32 (pixel-major, like model(x) in decode.py:130) for rows [row0,row1) */
int32_t lbdrn_predict(const LbdrnDesc* d, const void* msb_dev, const float* params_dev,
                      const float* coord_tab_dev, float* y_dev, void* stream);

/* ---- a11: full-scene squared error for best-epoch selection (encode.py:105-108, LBDRNperformance.py) -----
 * *sse_dev (device double) = sum over rows [row0,row1), all bands, of (y - code/(2^K-1))^2. Deterministic. */
int32_t lbdrn_eval_sse(const LbdrnDesc* d, const void* msb_dev, const void* lsb_dev, const float* params_dev,
                       const float* coord_tab_dev, double* sse_dev, void* stream);

/* ---- a7-a10: fused training (modified_ignite_engine.py:18-27, encode.py:84-85) ----------------------------
 * The handle owns parameters, Adam moments and gradient staging on the current device. msb/lsb buffers stay
 * caller-owned and must outlive the handle's use. */
int32_t lbdrn_train_create(const LbdrnDesc* d, const LbdrnTrainCfg* cfg, LbdrnTrain** out);
int32_t lbdrn_train_destroy(LbdrnTrain* t);
int32_t lbdrn_train_set_params(LbdrnTrain* t, const float* params_dev, void* stream);
int32_t lbdrn_train_get_params(LbdrnTrain* t, float* params_dev, void* stream);
/* Run `n_steps` consecutive optimiser steps in one persistent launch. Step s uses pixels
 * perm_dev[s*bs .. min((s+1)*bs, n_perm)) (flat pixel index y*W+x, int64 as produced by torch.randperm);
 * the last batch of an epoch may be partial (DataLoader drop_last=False, encode.py:69-70). `lr` is the
 * epoch's learning rate; losses_dev[s] receives the batch MSE (the value trainer.state.output holds).
 * `adam_t0` = optimiser steps already taken (bias correction uses t = adam_t0 + s + 1). */
int32_t lbdrn_train_steps(LbdrnTrain* t, const void* msb_dev, const void* lsb_dev, const float* coord_tab_dev,
                          const int64_t* perm_dev, int64_t n_perm, int32_t n_steps, int64_t adam_t0, double lr,
                          float* losses_dev, void* stream);
/* Data-parallel mode (cfg.world_size > 1): run forward/backward of ONE step on this rank's share of the batch
 * and leave the UNREDUCED local gradient sum in grad_dev [P] plus the local squared-error sum in
 * grad_dev[P]; the caller all-reduces that vector (NCCL) and then calls lbdrn_train_apply. */
int32_t lbdrn_train_grad(LbdrnTrain* t, const void* msb_dev, const void* lsb_dev, const float* coord_tab_dev,
                         const int64_t* batch_dev, int32_t n_local, int32_t n_global, float* grad_dev,
                         void* stream);
int32_t lbdrn_train_apply(LbdrnTrain* t, const float* grad_dev, int64_t adam_t, double lr, void* stream);

/* ---- a12/a13: nn sub-stream codec (HOST code: no device needed) ------------------------------------------------------
 * Replaces `fpzip.compress(params, precision=prec, order='C')` (encode.py:129) and `fpzip.decompress(bytes, order='C')`
 * (decode.py:113) when the fpzip package is not installed: the published fpzip algorithm for a flat float32 vector (PCmap
 * value map keeping the top `precision` bits, one-dimensional Lorenzo predictor, adaptive range coder).  Decoded values
 * are the inputs with their low 32-precision bits cleared, exactly as with fpzip (precision 0 or 32 = lossless); byte
 * compatibility of the payload with the real library is unverified (the library is not available to this build).
 * All pointers are HOST memory.  lbdrn_fpz_bound(n) = output capacity that always suffices for n values. */
int64_t lbdrn_fpz_bound(int64_t n);
int32_t lbdrn_fpz_compress(const float* data, int64_t n, int32_t precision, uint8_t* out, int64_t out_cap, int64_t* out_bytes);
int32_t lbdrn_fpz_header(const uint8_t* in, int64_t in_bytes, int64_t* n_out, int32_t* precision_out);
int32_t lbdrn_fpz_decompress(const uint8_t* in, int64_t in_bytes, float* out, int64_t out_cap);

#ifdef __cplusplus
}
#endif
#endif /* LBDRN_H_ */
