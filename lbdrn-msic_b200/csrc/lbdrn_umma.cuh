// lbdrn_umma.cuh -- shared pieces of the tcgen05 kernels: packed weight-block layout, UMMA descriptors, PTX wrappers.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "lbdrn_common.cuh"

namespace lbdrn {

constexpr int TC_THREADS = 128;       // one warpgroup: thread t <-> TMEM lane t <-> pixel t of the tile
constexpr int TC_BC = 64;             // hidden width this kernel is instantiated for
constexpr int TC_TH = 8, TC_TW = 16;  // 128-pixel tile
constexpr int TC_TMEM_COLS = 64;      // fp32 accumulator: bc columns (power of two >= 32)
constexpr int TC_MAX_K1 = 400;        // C*(2D+1)^2 <= 400 (8 bands, D=3 -> 392)

// ---- packed weight block (global scratch and, copied verbatim, shared memory) ------------------------------------
struct TcHeader {
  int exact;            // 1: every hidden weight is exactly representable as fp16 after its power-of-two scale
  int k1, k1pad;        // layer-1 K (= dim_in) and K rounded up to 16
  int nl;
  float scale[kMaxLayers];   // per hidden layer: multiply the fp32 accumulator by this (2^-s, and /max for layer 0)
  int off_bias, off_w3, off_b[kMaxLayers], total;   // byte offsets inside the block
  int off_blo[kMaxLayers];   // low-order fp16 term of the weights (W*2^s = hi + lo); all zero when `exact`
  int hi_bytes;              // block size without the lo operands (what the exact-weights kernel copies)
  int guard_mask;            // bit l set: layer l's sine needs the large-argument guard (|w0 z| may exceed 20000, or unknown)
  int klayout;               // order of layer 0's contraction index (lbdrn_tc.cu: plan_block)
  int pad[5];
};

__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

// byte offset of element (row, k) of a K-major, no-swizzle UMMA operand with `rows` rows:
// 8x8 core matrices of 128 contiguous bytes; consecutive 8-row groups are contiguous (SBO = 128 B), consecutive
// 8-element K chunks are rows*16 B apart (LBO).
__host__ __device__ inline int umma_off(int rows, int row, int k) { return ((k >> 3) * rows + row) * 16 + (k & 7) * 2; }

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout NONE [61,64)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// cute::UMMA::InstrDescriptor, kind::f16: D=f32 (bit 4), A=B=f16 (0), both K-major (0), N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

static __device__ int g_tc_timeout[4];   // diagnostics: {which barrier, tile, block, count}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity, int what = 0, int tile = -1, int no_trap = 0) {
  // bounded: a lost arrival traps (surfacing as a CUDA error) instead of hanging the GPU
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(mbar), "r"(parity), "r"(20000u)      // suspend-time hint (ns): sleep in hardware instead of re-polling
        : "memory");
    if (ok) return;
  }
  if (threadIdx.x == 0) {
    g_tc_timeout[0] = what; g_tc_timeout[1] = tile; g_tc_timeout[2] = blockIdx.x;
    atomicAdd(&g_tc_timeout[3], 1);
    __threadfence_system();
  }
  if (no_trap) return;      // LBDRN_DEBUG: let the kernel finish (garbage output) so the host can read the diagnostics
  __nanosleep(1000000);
  __trap();
}

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {      // release semantics at CTA scope (PTX default)
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}

// TMA: one 3-D box (x, y, band) of the MSB planes -> shared memory; out-of-range elements are zero-filled
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(mbar)
      : "memory");
}

// One lane of a converged warp.  The control warps run their loops warp-uniformly (every lane waits on the barriers) and
// only the tcgen05 / bulk-copy instructions sit under this predicate: operands then live in uniform registers.  Issuing
// from inside `if (lane == 0)` instead makes the compiler wrap every such instruction in an R2UR.BROADCAST / ELECT
// loop (~80 dependent instructions per K step: 300 cycles on the one thread the whole CTA's tensor work goes through).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.ld split into issue and wait so that the next block's load is in flight while this block's values are
// consumed.  The wait names the destination registers as read-write operands: nothing that uses them can be scheduled
// above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Packed fp32 pairs (sm_100: FFMA2 / FADD2, two IEEE fp32 operations per issue slot; each lane rounds exactly like the
// scalar instruction).  The kernels are bound by instruction issue, not by the FMA pipe, so pairing halves their cost.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  unsigned long long a, b, c, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long a, b, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

__device__ __forceinline__ uint32_t umma_idesc_f16_major(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// MN-major, no swizzle: element (mn, k) at ((mn/8)*K + k)*16 + (mn%8)*2 : 8 contiguous MN elements per 16 B, the 8 k-rows
// of a core matrix are 16 B apart, next k-group +128 B (LBO), next MN-group +K*16 B (SBO).
__host__ __device__ inline int umma_off_mn(int K, int mn, int k) { return ((mn >> 3) * K + k) * 16 + (mn & 7) * 2; }


// ---- operand images shared by the forward and the backward GEMMs of the tcgen05 training step ------------------------------
// An "image" is a [rows][K] fp16 array in the canonical no-swizzle layout: element (r, k) at ((k>>3)*rows + r)*16 + (k&7)*2.
//  * K-major view  (r indexes M or N, k is contracted):  LBO = rows*16, SBO = 128, one K = 16 step advances rows*32 bytes;
//  * MN-major view (k indexes M or N, r is contracted; the contraction length is `rows`): LBO = 128, SBO = rows*16, one
//    step of 16 contracted rows advances 256 bytes.
// So ONE copy of an activation / weight / gradient block serves x.W^T, dz.W and dz^T.x alike.
__host__ __device__ inline int img_off(int rows, int r, int k) { return ((k >> 3) * rows + r) * 16 + (k & 7) * 2; }
__device__ __forceinline__ uint64_t img_desc(uint32_t saddr, int rows, int mn_major, int kstep) {
  return mn_major ? umma_desc(saddr + (uint32_t)kstep * 256u, 128u, (uint32_t)rows * 16u)
                  : umma_desc(saddr + (uint32_t)kstep * (uint32_t)rows * 32u, (uint32_t)rows * 16u, 128u);
}

// tcgen05.ld 16x256b: the warp reads 16 lanes (its sub-partition's lanes 0-15: where an M = 64 accumulator keeps rows
// 16*(warp%4) .. +15) x 8 columns per repetition; register 4j+0/1 = row lane/4, columns 8j + 2*(lane%4) + {0,1};
// 4j+2/3 = row lane/4 + 8, same columns -- the accumulator fragment of a warp-level m16n8 MMA.
__device__ __forceinline__ void tmem_ld16x256_x1(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16x256_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16x256_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
// the wait names the destination registers as read-write operands so that no use of them is scheduled above it
__device__ __forceinline__ void tmem_ld_wait4(uint32_t (&r)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) : : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) : : "memory");
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) : : "memory");
}

// 32x32b.x8: thread <-> lane (row), 8 consecutive columns
__device__ __forceinline__ void tmem_ld32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}

}  // namespace lbdrn
