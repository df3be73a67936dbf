// lbdrn_infer_fp32.cu -- instantiations + launch dispatch of the fp32 inference kernel for ONE mode.
// Compiled three times (-DLBDRN_INFER_MODE=0/1/2) so the 48 heavy instantiations build in parallel.
#include "lbdrn_infer_fp32.cuh"
#include "lbdrn_internal.h"

#ifndef LBDRN_INFER_MODE
#error "compile with -DLBDRN_INFER_MODE=0|1|2"
#endif

namespace lbdrn {

constexpr int MODE = LBDRN_INFER_MODE;

template <int BC, int TM, int CP, bool WSMEM>
static int launch_t(const InferArgs& a, size_t smem, int sms, cudaStream_t st) {
  auto kern = infer_fp32_kernel<BC, TM, CP, WSMEM, MODE>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
  if (occ < 1) return fail(LBDRN_E_UNSUPPORTED, "fp32 inference kernel does not fit (smem %zu B)", smem);
  int grid = sms * occ;                       // persistent: a whole number of CTAs per SM
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (MODE == MODE_SSE && grid > 4096) grid = 4096;
  kern<<<grid, kThreads, smem, st>>>(a);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return LBDRN_OK;
}

template <int BC, int TM>
static int launch_bc(InferArgs& a, const Scratch& sc, cudaStream_t st) {
  const Net& n = a.net;
  constexpr int LDP = ldp_of<TM>(), TH = TM, TW = 16;
  a.kmax = n.dim_in > BC ? n.dim_in : BC;
  a.tiles_x = (n.W + TW - 1) / TW;
  a.n_tiles = a.tiles_x * ((n.row1 - n.row0 + TH - 1) / TH);
  const size_t base = ((size_t)a.kmax * LDP + round4(n.C * (TH + 2 * n.D) * (TW + 2 * n.D))) * sizeof(float);
  const size_t with_w = base + (size_t)round4(n.P) * sizeof(float);
  const bool wsmem = with_w <= (size_t)sc.max_smem;
  if (!wsmem && base > (size_t)sc.max_smem)
    return fail(LBDRN_E_UNSUPPORTED, "activation tile does not fit in shared memory");
  if (n.C <= 4)
    return wsmem ? launch_t<BC, TM, 4, true>(a, with_w, sc.sms, st) : launch_t<BC, TM, 4, false>(a, base, sc.sms, st);
  return wsmem ? launch_t<BC, TM, 8, true>(a, with_w, sc.sms, st) : launch_t<BC, TM, 8, false>(a, base, sc.sms, st);
}

static int dispatch(InferArgs& a, const Scratch& sc, cudaStream_t st) {
  switch (a.net.bc) {
    case 32: return launch_bc<32, 8>(a, sc, st);
    case 64: return launch_bc<64, 8>(a, sc, st);
    case 128: return launch_bc<128, 4>(a, sc, st);
    default: return launch_bc<256, 4>(a, sc, st);
  }
}

#if LBDRN_INFER_MODE == 0
int infer_fp32_decode(InferArgs& a, const Scratch& sc, cudaStream_t st) { return dispatch(a, sc, st); }
void launch_pack_params(const Net& n, const float* params, float* wpack, cudaStream_t st) {
  pack_params_kernel<<<(n.P + 255) / 256, 256, 0, st>>>(n, params, wpack);
  ++g_launches;
}
#elif LBDRN_INFER_MODE == 1
int infer_fp32_predict(InferArgs& a, const Scratch& sc, cudaStream_t st) { return dispatch(a, sc, st); }
#else
int infer_fp32_sse(InferArgs& a, const Scratch& sc, cudaStream_t st) { return dispatch(a, sc, st); }
#endif

}  // namespace lbdrn
