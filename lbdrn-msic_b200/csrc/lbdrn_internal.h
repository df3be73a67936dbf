// lbdrn_internal.h -- declarations shared by the translation units of liblbdrn_b200 (not part of the ABI).
#pragma once
#include <cuda.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <string>

#include "lbdrn_common.cuh"

namespace lbdrn {

int fail(int code, const char* fmt, ...);          // records the thread-local message, returns `code`
extern std::atomic<long long> g_launches;

#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess) return ::lbdrn::fail(LBDRN_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// per-device scratch (packed weights, SSE partials); grow-only, lives until process exit
struct Scratch {
  float* wpack = nullptr;
  size_t wpack_n = 0;
  double* partials = nullptr;
  unsigned int* counter = nullptr;
  int sms = 0, max_smem = 0;
};

int get_scratch(int P, Scratch*& out);   // per-device scratch of the calling thread's current device

struct InferArgs;
// fp32 inference launchers, one translation unit per mode (MODE_DECODE / MODE_PREDICT / MODE_SSE)
int infer_fp32_decode(InferArgs& a, const Scratch& sc, cudaStream_t st);
int infer_fp32_predict(InferArgs& a, const Scratch& sc, cudaStream_t st);
int infer_fp32_sse(InferArgs& a, const Scratch& sc, cudaStream_t st);
void launch_pack_params(const Net& n, const float* params, float* wpack, cudaStream_t st);

// fused training: kernel selection + launch (lbdrn_train_fp32.cu)
struct TrainPlan {
  void* kernel = nullptr;
  size_t smem = 0;
  bool wsmem = true;
  size_t wimg_bytes = 0;              // h2: bytes of the global image of the split weight operands
  bool h2 = false;                    // fp16 hi+lo split operands + ldmatrix variant selected
  bool tcx = false;                   // streamed tcgen05 variant (bc 256): the global weight image is (re)built before every launch
  int grid = 0, dimpad = 0, pstride = 0;
  int pf_off = 0, pf_stride = 0;      // TMA neighbourhood boxes: offset inside dynamic smem, bytes per pixel (0: not planned)
};
struct TrainArgs;
int train_fp32_plan(const Net& n, int batch_size, int sms, int max_smem, TrainPlan& plan);
int train_fp32_launch(const TrainPlan& plan, TrainArgs& a, cudaStream_t st);
void launch_tcx_wimg(const Net& n, const float* params, uint16_t* wimg, cudaStream_t st);
void launch_interleave_u8(const void* planes, int C, size_t npix, uint32_t* out, int sms, cudaStream_t st);
void launch_adam_apply(const Net& n, const float* grad, float* params, float* wpack, float* m, float* v, float omb1,
                       float omb2, float beta2, float eps, float step_size, float bc2_sqrt, cudaStream_t st);

// TMA descriptor of CHW planes, uploaded into a per-device ring (lbdrn_tc.cu); *out = nullptr if TMA cannot address them
int make_tensor_map_3d(const void* base, int es, int W, int rows, int C, int box_w, int box_h, int box_c, int dev,
                       cudaStream_t st, const CUtensorMap** out);

// tcgen05 tensor-core decode (lbdrn_tc.cu)
bool tc_supported(const Net& n);
int tc_decode(const Net& n, const void* msb, const float* params, const float* tab, uint16_t* out, int fast_sine,
              cudaStream_t st);
int tc_eval_sse(const Net& n, const void* msb, const void* lsb, const float* params, const float* tab, double* sse_out,
                cudaStream_t st);
// wide (bc 128/256) tcgen05 decode (lbdrn_tcw.cu)
bool tcw_supported(const Net& n);
int tcw_decode(const Net& n, const void* msb, const float* params, uint16_t* out, int fast_sine, const int** exact_flag_out,
               cudaStream_t st);
int tcw_eval_sse(const Net& n, const void* msb, const void* lsb, const float* params, double* sse_out, cudaStream_t st);
int tc_selftest2(const void* a_dev, const void* b_dev, float* d_dev, int N, int K, int a_mn, int b_mn, cudaStream_t st);
int tc_selftest3(const void* a_img, int a_bytes, const void* b_img, int b_bytes, float* d_dev, float* raw_dev, int N,
                 int ksteps, int a_mn, int a_rows, int b_mn, int b_rows, cudaStream_t st);
int tc_selftest(const void* a_dev, const void* b_dev, float* d_dev, int K, cudaStream_t st);

}  // namespace lbdrn
