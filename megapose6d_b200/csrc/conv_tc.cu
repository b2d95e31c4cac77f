// Implicit-GEMM convolution for sm_100a: TMA im2col (activations) + TMA tiled (weights) ->
// 128B-swizzled shared memory -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> fused epilogue
// (folded-BN bias, residual add, ReLU) -> bf16 NHWC.
//
// Replaces the cuDNN convolutions issued by every nn.Conv2d + BatchNorm2d + ReLU (+ residual) of the
// reference backbone (reference: src/megapose/models/torchvision_resnet.py:74-120 BasicBlock.forward,
// :298-311 ResNet._forward_impl).
//
// GEMM view:  D[M, N] = A[M, K] * B[N, K]^T
//   M = n_img * P * Q output pixels (flattened NHW, what TMA im2col walks natively)
//   N = C_out
//   K = R * S * C_in, ordered (r, s, c);  one K-block = 64 channels of one filter tap
//
// Persistent, warp-specialised CTA (256 threads):
//   warp 0   : TMA producer (one elected lane)
//   warp 1   : tcgen05.mma issuer (one elected lane)
//   warp 2   : TMEM allocator / deallocator
//   warps 4-7: epilogue (TMEM -> registers -> global), one TMEM lane quarter each
// Pipelines: smem ring full/empty (TMA <-> MMA) and a 2-deep TMEM accumulator ring (MMA <-> epilogue).
#include <cuda.h>
#include <vector>
#include "mpx_common.cuh"

namespace mpx {

#ifdef MPX_ACT_BF16
#define kTmaActType CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#else
#define kTmaActType CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#endif

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {  // ~4 s at 2 GHz
      printf("mpx conv: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(void* smem, const CUtensorMap* map, uint64_t* bar,
                                                   int c, int w, int h, int n, uint16_t off_w,
                                                   uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers fp16 and bf16 inputs (InstrDescriptor a/b format) with fp32 accumulation.
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-uniform form: executed by ALL lanes of the issuing warp with warp-uniform operands; one lane is elected inside the
// instruction.  With `if (lane == 0) tc_mma_f16(...)` ptxas cannot prove that the descriptors / TMEM address are uniform and
// wraps every UTCHMMA in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop (three dependent uniform-datapath round trips per
// instruction: ~150 cycles per MMA and issuing thread, tools/gpu_mma_probe.py).
__device__ __forceinline__ void tc_mma_f16_u(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_u(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

// elected-lane forms of the remaining single-thread operations of the producer / issuer warps (see tc_mma_f16_u)
#define MPX_ELECT_PRED "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
__device__ __forceinline__ void mbar_expect_tx_u(uint64_t* bar, uint32_t bytes) {
  asm volatile(MPX_ELECT_PRED "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_u(uint64_t* bar) {
  asm volatile(MPX_ELECT_PRED "@e mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_u(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(MPX_ELECT_PRED
               "@e cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4}], [%2];\n\t}" ::"r"(smem_u32(smem)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_u(void* smem, const CUtensorMap* map, uint64_t* bar, int c, int w, int h,
                                                     int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(MPX_ELECT_PRED
               "@e cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};\n\t}" ::"r"(smem_u32(smem)),
               "l"(map), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
               : "memory");
}
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart).
// Field layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor (start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version [46,48) = 1, layout_type [61,64) = 2 for SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(1) << 16;   // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(64) << 32;  // SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;   // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
  return d;
}

// ---------------------------------------------------------------------------------------------
// Epilogue of one accumulator row (one thread = one TMEM lane = one output pixel): TMEM -> registers ->
// +bias (shared memory, broadcast) -> +residual -> ReLU -> bf16 -> 16-byte global stores.  The residual of the
// next 32-column chunk is requested before the current chunk is processed, and the caller requests the first
// chunk before it waits for the accumulator, so that the global-load latency hides behind the MMAs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_res_chunk(const act_t* res_row, int c, uint4 (&r)[4]) {
  const uint4* r4 = reinterpret_cast<const uint4*>(res_row + c);
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = __ldg(r4 + i);
}

// +bias, +residual, ReLU, conversion and the four 16-byte stores of one 32-column chunk
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], act_t* out_c, const float* bias_c, bool has_res,
                                               const uint4 (&res)[4], int relu) {
  uint4* o4 = reinterpret_cast<uint4*>(out_c);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float f[8];
    const float4 b0 = *reinterpret_cast<const float4*>(bias_c + 8 * i);
    const float4 b1 = *reinterpret_cast<const float4*>(bias_c + 8 * i + 4);
    f[0] = __uint_as_float(v[8 * i + 0]) + b0.x;
    f[1] = __uint_as_float(v[8 * i + 1]) + b0.y;
    f[2] = __uint_as_float(v[8 * i + 2]) + b0.z;
    f[3] = __uint_as_float(v[8 * i + 3]) + b0.w;
    f[4] = __uint_as_float(v[8 * i + 4]) + b1.x;
    f[5] = __uint_as_float(v[8 * i + 5]) + b1.y;
    f[6] = __uint_as_float(v[8 * i + 6]) + b1.z;
    f[7] = __uint_as_float(v[8 * i + 7]) + b1.w;
    if (has_res) {
      const uint32_t rr[4] = {res[i].x, res[i].y, res[i].z, res[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_act2(rr[j]);
        f[2 * j] += t.x;
        f[2 * j + 1] += t.y;
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    uint4 o;
    o.x = pack_act2(f[0], f[1]);
    o.y = pack_act2(f[2], f[3]);
    o.z = pack_act2(f[4], f[5]);
    o.w = pack_act2(f[6], f[7]);
    o4[i] = o;
  }
}

// The TMEM reads are double-buffered: the tcgen05.ld of chunk c + 1 is in flight while chunk c is converted and stored (r02
// ncu: with one load + wait per chunk a quarter of the epilogue warps' samples sat on the first use of the loaded registers,
// and the N = 64 kernels -- whose epilogue per MMA cycle is the largest -- ran at 37-54% tensor-pipe activity).
template <int NCOLS>
__device__ __forceinline__ void epilogue_row(uint32_t taddr, bool valid, act_t* out_row,
                                             const act_t* res_row, const float* bias_s, int relu,
                                             uint4 (&res_cur)[4]) {
  const bool has_res = valid && res_row != nullptr;
  uint32_t va[32], vb[32];
  tc_ld_32x32(taddr, va);
#pragma unroll
  for (int cc = 0; cc < NCOLS / 32; ++cc) {
    const int c = cc * 32;
    uint4 res_nxt[4];
    if (has_res && c + 32 < NCOLS) load_res_chunk(res_row, c + 32, res_nxt);
    tc_wait_ld();  // chunk cc has landed (the only load in flight)
    if ((cc & 1) == 0) {
      if (c + 32 < NCOLS) tc_ld_32x32(taddr + static_cast<uint32_t>(c + 32), vb);
      if (valid) epilogue_chunk(va, out_row + c, bias_s + c, has_res, res_cur, relu);
    } else {
      if (c + 32 < NCOLS) tc_ld_32x32(taddr + static_cast<uint32_t>(c + 32), va);
      if (valid) epilogue_chunk(vb, out_row + c, bias_s + c, has_res, res_cur, relu);
    }
    if (has_res && c + 32 < NCOLS) {
#pragma unroll
      for (int i = 0; i < 4; ++i) res_cur[i] = res_nxt[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Staged epilogue of a 32-row x 64-column accumulator block (one warp), r02.  The row-per-thread form above stores 16 bytes
// per lane into 32 different 128-byte lines per instruction (and loads the residual the same way): 32 L1 wavefronts per
// instruction, 256 (512 with a residual) per warp and tile -- ncu on layer1: l1tex 67% / 86% busy, the epilogue warps never
// waiting for an accumulator, the MMA issuers waiting for free TMEM.  Here the block goes through a 4 KB shared-memory tile
// of the warp (rows of 128 B, 16-byte chunk c of row r at chunk position c ^ (r & 7): conflict-free both for "lane = row"
// and for "8 lanes = one row" access), so that global loads and stores are whole 128-byte lines, 4 per instruction:
//   stage_residual64   8 coalesced 16-byte loads per lane -> staging tile          (before the accumulator is waited for)
//   epilogue_compute64 TMEM -> +bias -> +residual (staging) -> ReLU -> act16 -> staging tile (in place)
//   store_staged64     staging tile -> 8 coalesced 16-byte stores per lane        (after the accumulator has been released)
// `pix` = index of the lane's output pixel (row of the [M, 64] output matrix), -1 for rows that are not stored.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
constexpr int kStageTileBytes = 32 * 128;  // per epilogue warp

__device__ __forceinline__ void stage_residual64(uint32_t stg, int lane, int pix, const act_t* residual) {
  const int c = lane & 7;
  uint4 t[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = 4 * j + (lane >> 3);
    const int pix_r = __shfl_sync(0xffffffffu, pix, r);
    t[j] = pix_r >= 0 ? __ldg(reinterpret_cast<const uint4*>(residual + static_cast<size_t>(pix_r) * 64) + c)
                      : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = 4 * j + (lane >> 3);
    sts128(stg + static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)), t[j]);
  }
  __syncwarp();
}

// bias: 64 floats in shared memory (`bias_smem`, a shared-window address) or, when `bias_g` is given, in global memory (read
// through the read-only path with one address per instruction: a broadcast that stays in L1)
__device__ __forceinline__ void epilogue_compute64(uint32_t taddr, uint32_t stg, int lane, bool has_res, uint32_t bias_smem,
                                                   int relu, const float* __restrict__ bias_g = nullptr) {
  uint32_t va[32], vb[32];
  tc_ld_32x32(taddr, va);
  tc_ld_32x32(taddr + 32u, vb);
  tc_wait_ld();
  const uint32_t row_base = stg + static_cast<uint32_t>(lane * 128);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t* v = (i < 4) ? (va + 8 * i) : (vb + 8 * (i - 4));
    const float4 b0 = bias_g ? __ldg(reinterpret_cast<const float4*>(bias_g) + 2 * i) : lds_f4(bias_smem + static_cast<uint32_t>(32 * i));
    const float4 b1 = bias_g ? __ldg(reinterpret_cast<const float4*>(bias_g) + 2 * i + 1)
                             : lds_f4(bias_smem + static_cast<uint32_t>(32 * i + 16));
    const uint32_t slot = row_base + static_cast<uint32_t>((i ^ (lane & 7)) << 4);
    float f[8];
    f[0] = __uint_as_float(v[0]) + b0.x;
    f[1] = __uint_as_float(v[1]) + b0.y;
    f[2] = __uint_as_float(v[2]) + b0.z;
    f[3] = __uint_as_float(v[3]) + b0.w;
    f[4] = __uint_as_float(v[4]) + b1.x;
    f[5] = __uint_as_float(v[5]) + b1.y;
    f[6] = __uint_as_float(v[6]) + b1.z;
    f[7] = __uint_as_float(v[7]) + b1.w;
    if (has_res) {
      const uint4 rv = lds128(slot);
      const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_act2(rr[j]);
        f[2 * j] += t.x;
        f[2 * j + 1] += t.y;
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    uint4 o;
    o.x = pack_act2(f[0], f[1]);
    o.y = pack_act2(f[2], f[3]);
    o.z = pack_act2(f[4], f[5]);
    o.w = pack_act2(f[6], f[7]);
    sts128(slot, o);
  }
}

template <int ROW_ELEMS = 64>  // channels per output row (the staged tile holds 64 of them, at `out`)
__device__ __forceinline__ void store_staged64(uint32_t stg, int lane, int pix, act_t* out) {
  const int c = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = 4 * j + (lane >> 3);
    const int pix_r = __shfl_sync(0xffffffffu, pix, r);
    const uint4 o = lds128(stg + static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)));
    if (pix_r >= 0) *(reinterpret_cast<uint4*>(out + static_cast<size_t>(pix_r) * ROW_ELEMS) + c) = o;
  }
  __syncwarp();  // the tile is free for the next residual
}

// Fused 3x3 / stride-2 / pad-1 max-pool of a staged tile (the ReLU'd stem output is consumed by nothing else): instead of
// storing its 32 pixels the warp max-reduces them into the POOLED tensor with 16-byte vector reductions
// (REDG.E.MAX.F16x8).  The maximum is exact, idempotent and order-independent, the inputs are >= 0 and the pooled tensor is
// zeroed before the launch, so partial windows combine to exactly what maxpool3x3s2_kernel computes from the stored tensor.
// Within the tile the three columns of a pooled pixel are combined first: an even column x is the centre of pooled column
// x/2 and takes its neighbours from the rows before / after it in the staging tile; an odd column only emits on its own when
// its centre lies in another warp's tile (first / last row of the tile).  ~0.8 reductions per stored 16 bytes instead of 2.25.
//   info: bit 0 centre, bit 1 orphan, bit 2 left neighbour in tile, bit 3 right neighbour in tile, bit 4 second pooled row
__device__ __forceinline__ void red_max_act8(act_t* dst, const uint4& v) {
#ifdef MPX_ACT_BF16
  asm volatile("red.global.max.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
  asm volatile("red.global.max.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#endif
}
__device__ __forceinline__ uint32_t max_act2(uint32_t a, uint32_t b) {
  const act_t2 r = __hmax2(*reinterpret_cast<const act_t2*>(&a), *reinterpret_cast<const act_t2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
// Rows with work (centres and orphans: every other row of the tile, ~17 of 32) are compacted first -- rank by ballot / popc,
// row index into a 32-byte list of the warp in shared memory (`list`) -- so that the 8-lane groups walk only those: ~5 fully
// populated iterations instead of 8 half-empty ones (r02 ncu: the epilogue warps of the stem were busy 100% of the time, 56% of
// it in this function, the tensor pipe waiting for them at 62%).
__device__ __forceinline__ void pool_staged64(uint32_t stg, uint32_t list, int lane, int info, int base, int row_elems,
                                              act_t* pool) {
  const unsigned work = __ballot_sync(0xffffffffu, (info & 3) != 0);
  const int count = __popc(work);
  if ((info & 3) != 0) {
    const uint32_t slot = list + static_cast<uint32_t>(__popc(work & ((1u << lane) - 1u)));
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(slot), "r"(lane) : "memory");
  }
  __syncwarp();
  const int c = lane & 7, g = lane >> 3;
  const int n_it = (count + 3) >> 2;
  for (int it = 0; it < n_it; ++it) {
    const int s = 4 * it + g;
    const bool active = s < count;
    uint32_t r = 0;
    if (active) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(list + static_cast<uint32_t>(s)));
    const int info_r = __shfl_sync(0xffffffffu, info, static_cast<int>(r));
    const int base_r = __shfl_sync(0xffffffffu, base, static_cast<int>(r));
    if (!active) continue;
    uint4 o = lds128(stg + (r * 128u + ((static_cast<uint32_t>(c) ^ (r & 7u)) << 4)));
    if (info_r & 4) {
      const uint4 t = lds128(stg + ((r - 1u) * 128u + ((static_cast<uint32_t>(c) ^ ((r - 1u) & 7u)) << 4)));
      o.x = max_act2(o.x, t.x); o.y = max_act2(o.y, t.y); o.z = max_act2(o.z, t.z); o.w = max_act2(o.w, t.w);
    }
    if (info_r & 8) {
      const uint4 t = lds128(stg + ((r + 1u) * 128u + ((static_cast<uint32_t>(c) ^ ((r + 1u) & 7u)) << 4)));
      o.x = max_act2(o.x, t.x); o.y = max_act2(o.y, t.y); o.z = max_act2(o.z, t.z); o.w = max_act2(o.w, t.w);
    }
    act_t* dst = pool + static_cast<size_t>(base_r) * 64 + c * 8;
    red_max_act8(dst, o);
    if (info_r & 16) red_max_act8(dst + row_elems, o);
  }
  __syncwarp();  // the tile (and the list) is free for the next block
}

// the same with the row length of the output matrix as a run-time value (conv_igemm2_kernel: C_out = 256 | 512)
__device__ __forceinline__ void store_staged64_rt(uint32_t stg, int lane, long long row, act_t* out, int row_elems) {
  const int c = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = 4 * j + (lane >> 3);
    const long long row_r = __shfl_sync(0xffffffffu, row, r);
    const uint4 o = lds128(stg + static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)));
    if (row_r >= 0) *(reinterpret_cast<uint4*>(out + static_cast<size_t>(row_r) * row_elems) + c) = o;
  }
  __syncwarp();
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Split-K over a thread-block cluster (small batches: a handful of output tiles, up to 72 serial k-blocks whose TMA
// issue rate bounds a single CTA).  The S CTAs of a cluster compute the same output tile over disjoint k-ranges; each
// parks its fp32 accumulator tile in its own shared memory (the drained pipeline stages), the cluster synchronises,
// and CTA `rank` finishes rows [rank*128/S, (rank+1)*128/S): it sums the S partial rows through distributed shared
// memory in rank order (deterministic), applies bias / residual / ReLU and stores bf16.  No global scratch, no
// second kernel; the reduction costs two cluster barriers and one round of DSMEM reads.
// ---------------------------------------------------------------------------------------------
template <int NCOLS>
struct SplitKTile {
  static constexpr int kPitch = NCOLS * 4 + 16;  // bytes per row; +16 keeps the row-owner 16-byte stores conflict-free
  static constexpr int kBytes = 128 * kPitch;
};

// this thread's accumulator row -> own shared memory
template <int NCOLS>
__device__ __forceinline__ void splitk_park_row(uint32_t taddr, uint32_t tile_smem, int row) {
  const uint32_t base = tile_smem + static_cast<uint32_t>(row * SplitKTile<NCOLS>::kPitch);
#pragma unroll 1
  for (int c = 0; c < NCOLS; c += 32) {
    uint32_t v[32];
    tc_ld_32x32(taddr + static_cast<uint32_t>(c), v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + static_cast<uint32_t>((c + 4 * i) * 4)),
                   "r"(v[4 * i]), "r"(v[4 * i + 1]), "r"(v[4 * i + 2]), "r"(v[4 * i + 3])
                   : "memory");
  }
}

// rows [rank*128/S, (rank+1)*128/S) of the tile: sum over the cluster, epilogue, store.  All threads of the CTA.
template <int NCOLS>
__device__ __forceinline__ void splitk_reduce_slice(uint32_t tile_smem, int splits, int rank, long long m0,
                                                    int M_total, int C_out, int n0, act_t* __restrict__ out,
                                                    const act_t* __restrict__ residual, const float* bias_s,
                                                    int relu) {
  constexpr int kVecPerRow = NCOLS / 4;
  const int rows_per = 128 / splits;
  for (int u = threadIdx.x; u < rows_per * kVecPerRow; u += blockDim.x) {
    const int r = rank * rows_per + u / kVecPerRow;
    const int c = (u % kVecPerRow) * 4;
    const long long m = m0 + r;
    if (m >= M_total) continue;
    const uint32_t local = tile_smem + static_cast<uint32_t>(r * SplitKTile<NCOLS>::kPitch + c * 4);
    float4 part[8];
#pragma unroll
    for (int sp = 0; sp < 8; ++sp) {
      if (sp < splits) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(sp));
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(part[sp].x), "=f"(part[sp].y), "=f"(part[sp].z), "=f"(part[sp].w)
                     : "r"(remote)
                     : "memory");
      }
    }
    float4 a = part[0];
#pragma unroll
    for (int sp = 1; sp < 8; ++sp)
      if (sp < splits) { a.x += part[sp].x; a.y += part[sp].y; a.z += part[sp].z; a.w += part[sp].w; }
    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c);
    float f[4] = {a.x + b4.x, a.y + b4.y, a.z + b4.z, a.w + b4.w};
    const size_t off = static_cast<size_t>(m) * C_out + n0 + c;
    if (residual != nullptr) {
      const uint2 rr = __ldg(reinterpret_cast<const uint2*>(residual + off));
      const float2 t0 = unpack_act2(rr.x), t1 = unpack_act2(rr.y);
      f[0] += t0.x; f[1] += t0.y; f[2] += t1.x; f[3] += t1.y;
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    uint2 o;
    o.x = pack_act2(f[0], f[1]);
    o.y = pack_act2(f[2], f[3]);
    *reinterpret_cast<uint2*>(out + off) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// optional per-launch timing (bench.py roofline): CUDA events around every conv launch
// ---------------------------------------------------------------------------------------------
struct ProfileSlot {
  cudaEvent_t e0, e1;
  double flops;
};
static bool g_profile = false;
static std::vector<ProfileSlot> g_slots;
static size_t g_slots_used = 0;

static ProfileSlot* profile_begin(cudaStream_t stream) {
  if (!g_profile) return nullptr;
  if (g_slots_used == g_slots.size()) {
    ProfileSlot s;
    if (cudaEventCreate(&s.e0) != cudaSuccess || cudaEventCreate(&s.e1) != cudaSuccess) return nullptr;
    s.flops = 0;
    g_slots.push_back(s);
  }
  ProfileSlot* slot = &g_slots[g_slots_used++];
  cudaEventRecord(slot->e0, stream);
  return slot;
}
static void profile_end(ProfileSlot* slot, cudaStream_t stream, double flops) {
  if (!slot) return;
  slot->flops = flops;
  cudaEventRecord(slot->e1, stream);
}
bool conv_profile_enabled() { return g_profile; }
void conv_profile_enable(int on) {
  g_profile = on != 0;
  g_slots_used = 0;
}
// Synchronises the device; sums the recorded launches since conv_profile_enable(1) and resets.
int conv_profile_summary(double* total_ms, double* total_flops, long long* launches) {
  MPX_CHECK_CUDA(cudaDeviceSynchronize());
  double ms = 0, fl = 0;
  for (size_t i = 0; i < g_slots_used; ++i) {
    float t = 0.f;
    MPX_CHECK_CUDA(cudaEventElapsedTime(&t, g_slots[i].e0, g_slots[i].e1));
    ms += t;
    fl += g_slots[i].flops;
  }
  *total_ms = ms;
  *total_flops = fl;
  *launches = static_cast<long long>(g_slots_used);
  g_slots_used = 0;
  return MPX_OK;
}

struct ConvParams {
  int M_total;  // n_img * P * Q
  int P, Q;     // output height / width
  int C_out;
  int S;         // filter width (taps are ordered r-major)
  int stride;
  int pad_h, pad_w;  // lower padding
  int cblocks;       // C_in / 64
  int num_k_blocks;  // R * S * cblocks
  int m_tiles, n_tiles;
  int relu;
  const float* bias;                // [C_out] folded BN shift
  const act_t* residual;    // [M_total, C_out] or nullptr
  act_t* out;               // [M_total, C_out]
  // split-K: the `splits` (1, 2, 4 or 8) CTAs of a cluster share one output tile, each accumulating a contiguous
  // range of k-blocks; reduction through distributed shared memory (see SplitKTile)
  int splits;
  int staged;  // conv_igemm2_kernel<256>: staged epilogue, one 128-channel half at a time (mode bit 25)
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int kBTileBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BLOCK_N;  // double-buffered accumulator
  static constexpr int kSmemBytes =
      kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 2048 /*bias, C_out <= 512*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const ConvParams p) {
  using Cfg = ConvCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kATileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  float* bias_s = reinterpret_cast<float*>(bars + 32);  // [C_out] folded-BN bias, read by broadcast in the epilogue

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction (role dispatch, UR operands)
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;  // work items: (tile, k-split)
  for (int i = threadIdx.x; i < p.C_out; i += blockDim.x) bias_s[i] = p.bias[i];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // activations / residual of the previous kernel are complete from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      int stage = 0;
      uint32_t phase = 0;
      const int pq = p.P * p.Q;
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int tile = item / p.splits;
        const int split = item - tile * p.splits;
        const int kb_begin = split * p.num_k_blocks / p.splits;
        const int kb_end = (split + 1) * p.num_k_blocks / p.splits;
        const int m_tile = tile / p.n_tiles;
        const int n_tile = tile - m_tile * p.n_tiles;
        const int m0 = m_tile * kBlockM;
        const int img = m0 / pq;
        const int rem = m0 - img * pq;
        const int p0 = rem / p.Q;
        const int q0 = rem - p0 * p.Q;
        const int base_w = q0 * p.stride - p.pad_w;
        const int base_h = p0 * p.stride - p.pad_h;
        int tap = kb_begin / p.cblocks, cb = kb_begin - tap * p.cblocks;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx_u(&full_bar[stage], Cfg::kStageBytes);
          const int r = tap / p.S;
          const int s = tap - r * p.S;
          tma_load_im2col_4d_u(smem_a + stage * kATileBytes, &map_a, &full_bar[stage], cb * kBlockK,
                             base_w, base_h, img, static_cast<uint16_t>(s),
                             static_cast<uint16_t>(r));
          tma_load_2d_u(smem_b + stage * Cfg::kBTileBytes, &map_b, &full_bar[stage], kb * kBlockK,
                      n_tile * BLOCK_N);
          if (++cb == p.cblocks) {
            cb = 0;
            ++tap;
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      // InstrDescriptor (cute/arch/mma_sm100_desc.hpp): c_format F32 [4,6)=1, a/b format
      // [7,10), [10,13) = kIdescAB (0 = F16, 1 = BF16), a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
      constexpr uint32_t idesc = (1u << 4) | kIdescAB |
                                 (static_cast<uint32_t>(BLOCK_N >> 3) << 17) |
                                 (static_cast<uint32_t>(kBlockM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x, ++local) {
        const int split = item % p.splits;
        const int kb_begin = split * p.num_k_blocks / p.splits;
        const int kb_end = (split + 1) * p.num_k_blocks / p.splits;
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sw128_desc(smem_u32(smem_a + stage * kATileBytes));
          const uint64_t db = make_sw128_desc(smem_u32(smem_b + stage * Cfg::kBTileBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the swizzle atom: +2 in the (addr >> 4) field
            tc_mma_f16_u(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                        idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          tc_commit_u(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit_u(&tmem_full[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q4 = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q4 * 32 + lane;
    int local = 0;
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x, ++local) {
      const int tile = item / p.splits;
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const long long m = static_cast<long long>(m_tile) * kBlockM + row;
      const bool valid = m < p.M_total;
      const int n0 = n_tile * BLOCK_N;
      const size_t off = static_cast<size_t>(valid ? m : 0) * p.C_out + n0;
      const act_t* res_row = p.residual ? p.residual + off : nullptr;
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
      if (p.splits == 1) {
        uint4 res_cur[4];
        if (valid && res_row) load_res_chunk(res_row, 0, res_cur);  // in flight while the MMAs finish
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        epilogue_row<BLOCK_N>(taddr, valid, p.out + off, res_row, bias_s + n0, p.relu, res_cur);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      } else {
        // one item per CTA: park the fp32 accumulator row in the (drained) pipeline stage memory
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        splitk_park_row<BLOCK_N>(taddr, smem_u32(smem), row);
        tc_fence_before();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.splits > 1) {
    cluster_sync_all();  // every CTA of the cluster has parked its partial tile
    const int item = blockIdx.x;
    const int tile = item / p.splits;
    const int m_tile = tile / p.n_tiles;
    const int n_tile = tile - m_tile * p.n_tiles;
    const int n0 = n_tile * BLOCK_N;
    splitk_reduce_slice<BLOCK_N>(smem_u32(smem), p.splits, static_cast<int>(cluster_ctarank()),
                                 static_cast<long long>(m_tile) * kBlockM, p.M_total, p.C_out, n0, p.out, p.residual,
                                 bias_s + n0, p.relu);
    cluster_sync_all();  // peers may still be reading this CTA's shared memory
  }
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(Cfg::kTmemCols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                     cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill);

static PFN_encodeTiled g_encode_tiled = nullptr;
static PFN_encodeIm2col g_encode_im2col = nullptr;

// libcuda is reached through the runtime so the library loads (and its symbols can be listed) on a
// machine without a driver.
static int load_driver_entry_points() {
  if (g_encode_tiled && g_encode_im2col) return MPX_OK;
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  MPX_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  MPX_REQUIRE(f != nullptr && q == cudaDriverEntryPointSuccess,
              "cuTensorMapEncodeTiled not available from the driver");
  g_encode_tiled = reinterpret_cast<PFN_encodeTiled>(f);
  f = nullptr;
  MPX_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q));
  MPX_REQUIRE(f != nullptr && q == cudaDriverEntryPointSuccess,
              "cuTensorMapEncodeIm2col not available from the driver");
  g_encode_im2col = reinterpret_cast<PFN_encodeIm2col>(f);
  return MPX_OK;
}

template <int BLOCK_N>
static int launch_conv(const CUtensorMap& ma, const CUtensorMap& mb, const ConvParams& p,
                       cudaStream_t stream, int max_ctas) {
  using Cfg = ConvCfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    MPX_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  int grid = p.m_tiles * p.n_tiles * p.splits;
  int cap = max_ctas > 0 ? max_ctas : sm_count();
  ProfileSlot* slot = profile_begin(stream);
  // split-K: one work item per CTA, the k-splits of a tile form a cluster
  if (p.splits == 1 && grid > cap) grid = cap;
  MPX_CHECK_CUDA(launch_pdl(conv_igemm_kernel<BLOCK_N>, dim3(grid), dim3(256), Cfg::kSmemBytes, stream, p.splits, ma, mb, p));
  MPX_CHECK_CUDA(cudaGetLastError());
  ++g_launches;
  profile_end(slot, stream, 2.0 * p.M_total * p.C_out * p.num_k_blocks * kBlockK);
  return MPX_OK;
}

static int conv_window_try(const ConvDesc& d, const void* x, const void* w, const float* bias, const void* residual,
                           void* out, int max_ctas, cudaStream_t stream);
static int conv_window2_try(const ConvDesc& d, const void* x, const void* w, const float* bias, const void* residual,
                            void* out, int max_ctas, cudaStream_t stream);
static int conv_window2q_try(const ConvDesc& d, const void* x, const void* w, const float* bias, const void* residual,
                             void* out, int max_ctas, cudaStream_t stream);
static int conv_windowq_try(const ConvDesc& d, const void* x, const void* w, const float* bias, const void* residual,
                            void* out, int max_ctas, cudaStream_t stream);
static int conv_windows_try(const ConvDesc& d, const void* x, const void* w, const float* bias, const void* residual,
                            void* out, int max_ctas, cudaStream_t stream);
static int conv_mode();
struct ConvParams;
template <int BLOCK_N>
static int launch_conv2(const CUtensorMap& ma, const CUtensorMap& mb, const ConvParams& p, cudaStream_t stream,
                        int max_ctas);

int conv_out_dim(int in, int pad_lo, int pad_hi, int k, int stride) {
  return (in + pad_lo + pad_hi - k) / stride + 1;
}

// x: [n_img, H, W, C_in] bf16; w: [C_out, R*S*C_in] bf16 ((r,s,c) ordered); bias fp32 [C_out];
// residual/out: [n_img, P, Q, C_out] bf16.
int conv_forward(const ConvDesc& d, const void* x, const void* w, const float* bias,
                 const void* residual, void* out, int block_n_override, int max_ctas,
                 cudaStream_t stream, int splitk) {
  MPX_REQUIRE(d.C_in % 64 == 0 && d.C_in >= 64, "conv: C_in=%d must be a multiple of 64", d.C_in);
  MPX_REQUIRE(d.C_out % 64 == 0 && d.C_out <= 512, "conv: C_out=%d must be a multiple of 64, at most 512", d.C_out);
  MPX_REQUIRE(d.stride == 1 || d.stride == 2, "conv: stride %d unsupported", d.stride);
  MPX_REQUIRE(d.R >= 1 && d.R <= 8 && d.S >= 1 && d.S <= 8, "conv: filter %dx%d unsupported", d.R,
              d.S);
  int rc = load_driver_entry_points();
  if (rc != MPX_OK) return rc;
  if (d.pool) {  // fused max-pool epilogue: only the CTA-pair window kernels have one; the caller falls back to conv + max-pool
    if (block_n_override != 0) return MPX_ERR_UNSUPPORTED;
    rc = conv_windows_try(d, x, w, bias, residual, out, max_ctas, stream);
    if (rc != MPX_ERR_UNSUPPORTED) return rc;
    return conv_windowq_try(d, x, w, bias, residual, out, max_ctas, stream);
  }
  if (block_n_override == 0) {  // auto: the window kernel serves the 64 -> 64 stride-1 layers
    rc = conv_windows_try(d, x, w, bias, residual, out, max_ctas, stream);  // CTA pairs, sliding window (bit 23)
    if (rc != MPX_ERR_UNSUPPORTED) return rc;
    rc = conv_windowq_try(d, x, w, bias, residual, out, max_ctas, stream);  // CTA pairs (bit 15, default)
    if (rc != MPX_ERR_UNSUPPORTED) return rc;
    rc = conv_window_try(d, x, w, bias, residual, out, max_ctas, stream);
    if (rc != MPX_ERR_UNSUPPORTED) return rc;
    rc = conv_window2q_try(d, x, w, bias, residual, out, max_ctas, stream);  // CTA pairs (bit 14, default)
    if (rc != MPX_ERR_UNSUPPORTED) return rc;
    rc = conv_window2_try(d, x, w, bias, residual, out, max_ctas, stream);
    if (rc != MPX_ERR_UNSUPPORTED) return rc;
  }

  const int P = conv_out_dim(d.H, d.pad_lo_h, d.pad_hi_h, d.R, d.stride);
  const int Q = conv_out_dim(d.W, d.pad_lo_w, d.pad_hi_w, d.S, d.stride);
  MPX_REQUIRE(P > 0 && Q > 0, "conv: empty output");
  const long long M_total = static_cast<long long>(d.n_img) * P * Q;
  MPX_REQUIRE(M_total > 0 && M_total < (1LL << 31), "conv: M=%lld out of range", M_total);

  int block_n = block_n_override;
  if (block_n <= 0) {
    // wide tiles amortise the activation loads; with only a handful of output tiles (refiner: one sample) the
    // conv is bound by the serial K loop of a single tile instead, so narrower tiles spread it over more SMs
    const long long m_tiles_est = (M_total + kBlockM - 1) / kBlockM;
    block_n = d.C_out >= 256 ? 256 : d.C_out;
    while (block_n > 64 && m_tiles_est * (d.C_out / block_n) < 32) block_n /= 2;
  }
  MPX_REQUIRE((block_n == 64 || block_n == 128 || block_n == 256) && d.C_out % block_n == 0,
              "conv: BLOCK_N=%d invalid for C_out=%d", block_n, d.C_out);
  // measured on B200 (tools/gpu_probe_pair.py): the pair kernel wins for BLOCK_N = 256 (layer3 1184 -> 1314, layer4 1382 ->
  // 1488 TFLOP/s) and is on par or slightly behind for 128, so it serves the 256-wide tiles only (bit 2 forces 128 too)
  const bool use_pair = (conv_mode() & 2) != 0 && block_n_override == 0 &&
                        (block_n == 256 || (block_n == 128 && (conv_mode() & 4) != 0));

  // --- activation map (im2col). Dims are innermost-first: {C, W, H, N}.
  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.C_in), static_cast<cuuint64_t>(d.W),
                          static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.n_img)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d.C_in) * 2,
                             static_cast<cuuint64_t>(d.W) * d.C_in * 2,
                             static_cast<cuuint64_t>(d.H) * d.W * d.C_in * 2};
    // Bounding box of the filter's *base* pixel (CUTLASS: lower = -pad_lo,
    // upper = pad_hi - (filter-1)*dilation; cutlass/conv/collective/detail.hpp).
    int lower[2] = {-d.pad_lo_w, -d.pad_lo_h};
    int upper[2] = {d.pad_hi_w - (d.S - 1), d.pad_hi_h - (d.R - 1)};
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(d.stride), static_cast<cuuint32_t>(d.stride), 1};
    CUresult r = g_encode_im2col(&map_a, kTmaActType, 4, const_cast<void*>(x),
                                 dims, strides, lower, upper, /*channelsPerPixel=*/kBlockK,
                                 /*pixelsPerColumn=*/kBlockM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d)", static_cast<int>(r));
    // Driver quirk (<= 13.1) for im2col maps over tensors smaller than 128 KiB: same fix-up as
    // cute/atom/copy_traits_sm90_im2col.hpp.
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const size_t bytes = static_cast<size_t>(d.n_img) * d.H * d.W * d.C_in * 2;
    if (drv <= 13010 && bytes < 131072) {
      reinterpret_cast<uint64_t*>(&map_a)[1] &= ~(1ull << 21);
    }
  }
  {
    const cuuint64_t K_total = static_cast<cuuint64_t>(d.R) * d.S * d.C_in;
    cuuint64_t dims[2] = {K_total, static_cast<cuuint64_t>(d.C_out)};
    cuuint64_t strides[1] = {K_total * 2};
    // the CTA-pair kernel loads half of the weight tile per CTA
    cuuint32_t box[2] = {kBlockK, static_cast<cuuint32_t>(use_pair ? block_n / 2 : block_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&map_b, kTmaActType, 2, const_cast<void*>(w), dims,
                                strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  }

  ConvParams p{};
  p.M_total = static_cast<int>(M_total);
  p.P = P;
  p.Q = Q;
  p.C_out = d.C_out;
  p.S = d.S;
  p.stride = d.stride;
  p.pad_h = d.pad_lo_h;
  p.pad_w = d.pad_lo_w;
  p.cblocks = d.C_in / kBlockK;
  p.num_k_blocks = d.R * d.S * p.cblocks;
  p.m_tiles = static_cast<int>((M_total + kBlockM - 1) / kBlockM);
  p.n_tiles = d.C_out / block_n;
  p.relu = d.relu;
  p.bias = bias;
  p.residual = reinterpret_cast<const act_t*>(residual);
  p.out = reinterpret_cast<act_t*>(out);
  p.splits = 1;
  if (splitk != 0 && !use_pair) {
    // Few output tiles and a long K loop (deep layers at batch 1: one 80-row tile, 72 k-blocks): split K over a cluster.
    const int tiles = p.m_tiles * p.n_tiles;
    const int sms = max_ctas > 0 ? max_ctas : sm_count();
    int want = splitk > 0 ? splitk : 1;
    if (splitk < 0 && tiles * 2 <= sms && p.num_k_blocks >= 8) {
      want = p.num_k_blocks / 4;
      if (want > sms / tiles) want = sms / tiles;
    }
    // bits 18 / 19 of the mode cap the automatic split at 2 / 1: fewer, longer CTAs -- less SM time per layer at a higher
    // latency, the better trade when another frame's kernels fill the device beside this one (frame_pipeline.py)
    const int cap = (splitk < 0 && (conv_mode() & 524288)) ? 1 : ((splitk < 0 && (conv_mode() & 262144)) ? 2 : 8);
    int splits = 1;
    while (splits * 2 <= want && splits * 2 <= cap && splits * 2 <= p.num_k_blocks) splits *= 2;
    p.splits = splits;
  }

  if (use_pair) {
    if (block_n == 128) return launch_conv2<128>(map_a, map_b, p, stream, max_ctas);
    return launch_conv2<256>(map_a, map_b, p, stream, max_ctas);
  }
  switch (block_n) {
    case 64:
      return launch_conv<64>(map_a, map_b, p, stream, max_ctas);
    case 128:
      return launch_conv<128>(map_a, map_b, p, stream, max_ctas);
    default:
      return launch_conv<256>(map_a, map_b, p, stream, max_ctas);
  }
}

// ---------------------------------------------------------------------------------------------
// "Window" convolution for the 64 -> 64 channel stride-1 layers (space-to-depth stem, layer1).
//
// The im2col kernel above re-reads every activation once per filter tap from L2 (9 x 16 KB of A plus 9 x 8 KB of
// B per 128x64 output tile): those layers are L2->SM bandwidth bound at ~22% tensor-pipe activity
// (profiles/r01_ncu_summary.md).  Here the M dimension walks the *zero-padded* image linearly,
// q = (img*Hp + y)*Wp + x, so that the input of tap (r, s) for output row q is simply row q + r*Wp + s of one
// contiguous window.  TMA im2col synthesises that padded window on the fly from the unpadded NHWC tensor (bounding
// box = image + padding ring, zero OOB fill, offsets 0), one load of 128 + (R-1)*Wp + (S-1) rows per tile, and all
// R*S taps issue their tcgen05.mma from row-shifted shared-memory descriptors into the same window (the 128B
// swizzle is a function of the absolute shared-memory address, so a start address shifted by whole 128-byte rows
// addresses the swizzled rows correctly; verified by mpx_debug_umma_rowshift).  The R*S weight tiles stay resident
// in shared memory for the lifetime of the CTA.  L2 traffic per tile drops from 216 KB to 38 KB; ring rows of
// the padded space are computed and discarded (6% for 60x80).
// ---------------------------------------------------------------------------------------------
struct WinParams {
  int Hp, Wp;          // padded image size (H + pl_h + ph_h, W + pl_w + ph_w)
  int H, W;            // output (= input) spatial size
  int pl_h, pl_w;      // top / left padding
  int n_img;
  int taps_per_win;    // rg * S filter taps served by one window
  int n_windows;       // R / rg
  int S;
  int rg;              // filter rows per window
  int win_rows;        // rows of one window actually needed
  int chunk_rows;      // rows per TMA issue (<= 256, multiple of 8)
  int n_chunks;
  int win_bytes;       // chunk_rows * n_chunks * 128
  long long M_pad;     // n_img * Hp * Wp
  long long q_base;    // first padded-linear index that can be a valid output
  int m_tiles;
  int relu;
  int mma_issuers;     // 1 or 2 issuing threads (tiles alternate)
  int observers_arrive;  // 1: a stage is refilled only after EVERY issuer has seen its fill (empty count = issuers)
  int row_epilogue;      // conv_windowq_kernel: 1 = row-per-thread epilogue (mode bit 20), 0 = staged / coalesced
  unsigned idesc;        // conv_windowq_kernel: the MMA instruction descriptor, as a parameter so that it lives in one uniform register
  unsigned long long kskip;  // bit 4 * tap + ks set: K step ks (16 channels) of filter tap `tap` has all-zero weights
  const float* bias;
  const act_t* residual;
  act_t* out;
  int pool;            // conv_windowq_kernel: 1 = `out` is the zero-initialised [n, pool_H, pool_W, 64] max-pooled tensor (3x3/s2/p1)
  int pool_H, pool_W;
};

constexpr int kWinN = 64;          // C_out
constexpr int kWinBTile = 64 * 128;  // one tap's weights: 64 rows x 64 channels bf16

// Epilogue organisation: one epilogue warp per scheduler cannot hide its own latencies (tcgen05.ld, the dependent
// fp32 math, the stores): ~2 us per 128 x 64 tile against 0.6 us of MMAs.  EPI_SETS groups of four warps therefore take
// tiles round-robin (set = tile % EPI_SETS) over kWinAccBufs TMEM accumulators, so several tiles are drained at once.
constexpr int kWinAccBufs = 4;
template <int EPI_SETS>
__global__ void __launch_bounds__(128 + 128 * EPI_SETS, 1)
conv_window_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const WinParams p, int stages, int n_taps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_b = smem;                                  // n_taps resident weight tiles
  uint8_t* smem_a = smem + n_taps * kWinBTile;             // `stages` windows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + static_cast<size_t>(stages) * p.win_bytes);
  uint64_t* full_bar = bars;             // [stages]
  uint64_t* empty_bar = bars + 8;        // [stages]
  uint64_t* tmem_full = bars + 16;       // [kWinAccBufs]
  uint64_t* tmem_empty = bars + 20;      // [kWinAccBufs]
  uint64_t* b_full = bars + 24;          // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 25);
  float* bias_s = reinterpret_cast<float*>(bars + 32);  // [64]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction (role dispatch, UR operands)
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < kWinN) bias_s[threadIdx.x] = p.bias[threadIdx.x];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], p.observers_arrive ? static_cast<uint32_t>(p.mma_issuers) : 1u);
    }
    for (int i = 0; i < kWinAccBufs; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)),
                 "r"(kWinAccBufs * kWinN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int hpwp = p.Hp * p.Wp;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      // resident weights: one [64 x 64] tile per tap
      mbar_expect_tx_u(b_full, static_cast<uint32_t>(n_taps) * kWinBTile);
      for (int t = 0; t < n_taps; ++t) tma_load_2d_u(smem_b + t * kWinBTile, &map_b, b_full, t * kBlockK, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        const long long q0 = p.q_base + static_cast<long long>(tile) * kBlockM;
        for (int wi = 0; wi < p.n_windows; ++wi) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx_u(&full_bar[stage], static_cast<uint32_t>(p.win_bytes));
          // window start in padded-linear space: rows of tap row-group wi start rg*wi rows of Wp further down
          long long qs = q0 - (static_cast<long long>(p.pl_h) * p.Wp + p.pl_w) + static_cast<long long>(wi) * p.rg * p.Wp;
          for (int ch = 0; ch < p.n_chunks; ++ch) {
            const long long q = qs + static_cast<long long>(ch) * p.chunk_rows;
            const int img = static_cast<int>(q / hpwp);
            const int rem = static_cast<int>(q - static_cast<long long>(img) * hpwp);
            const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
            tma_load_im2col_4d_u(smem_a + static_cast<size_t>(stage) * p.win_bytes + static_cast<size_t>(ch) * p.chunk_rows * 128,
                               &map_a, &full_bar[stage], 0, xp - p.pl_w, yp - p.pl_h, img, 0, 0);
          }
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 1 && warp <= 3) {
    // ===================== MMA issuers =====================
    // Two issuing threads (warps 1 and 3) take tiles alternately: every tile's MMAs are still issued in order by one
    // thread into its own accumulator (deterministic), but the per-instruction issue cost of the short N = 64 MMAs --
    // which, not the tensor pipe, bounds a single issuer -- overlaps between two tiles.  mode bit 5 (32): warp 1 only.
    const int n_issuers = p.mma_issuers;
    const int which = warp == 1 ? 0 : (warp == 3 ? 1 : 2);  // warp 2 (TMEM allocator) doubles as the third issuer
    if (which < n_issuers) {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      constexpr uint32_t idesc = (1u << 4) | kIdescAB | (static_cast<uint32_t>(kWinN >> 3) << 17) |
                                 (static_cast<uint32_t>(kBlockM >> 4) << 24);
      mbar_wait(b_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++local) {
        if (local % n_issuers != which) {
          // The other issuer's tile: only observe its window fills, so that this thread never runs more than one phase
          // ahead of the ring (mbarrier parity waits cannot tell fill i from fill i + 2).  The converse -- this thread
          // falling two fills of one stage BEHIND -- would need the producer's whole TMA round trip for a later window
          // (it starts only when this thread's own previous window has retired) to beat the few instructions between
          // that commit and this wait; with observers_arrive the refill additionally waits for this thread's arrival,
          // which rules the case out by construction (always on since r02: no measurable cost, 13.42 vs 13.50 ms per step).
          for (int wi = 0; wi < p.n_windows; ++wi) {
            mbar_wait(&full_bar[stage], phase);
            if (p.observers_arrive) mbar_arrive_u(&empty_bar[stage]);
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          continue;
        }
        const int acc = local % kWinAccBufs;
        const uint32_t acc_phase = (local / kWinAccBufs) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kWinN);
        uint32_t first = 1;
        for (int wi = 0; wi < p.n_windows; ++wi) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // descriptors advance by plain 64-bit adds on the (address >> 4) field: one 128-byte row = 8, one weight
          // tile = 512; the issuing thread's instruction count per MMA bounds the rate of these short (N = 64) MMAs
          const uint64_t da_win = make_sw128_desc(smem_u32(smem_a + static_cast<size_t>(stage) * p.win_bytes));
          uint64_t db = make_sw128_desc(smem_u32(smem_b + wi * p.taps_per_win * kWinBTile));
          uint64_t da_row = da_win;
          int tap = wi * p.taps_per_win;
          for (int r = 0; r < p.rg; ++r) {
            uint64_t da = da_row;
            for (int s = 0; s < p.S; ++s) {
              const unsigned sk = static_cast<unsigned>(p.kskip >> (4 * tap)) & 15u;
              if (sk == 0u) {
                tc_mma_f16_u(tmem_d, da, db, idesc, first ? 0u : 1u);
                tc_mma_f16_u(tmem_d, da + 2, db + 2, idesc, 1u);
                tc_mma_f16_u(tmem_d, da + 4, db + 4, idesc, 1u);
                tc_mma_f16_u(tmem_d, da + 6, db + 6, idesc, 1u);
                first = 0;
              } else {  // structurally zero weight slices (7x7 stem inside its 8x8 space-to-depth footprint): not issued
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (((sk >> ks) & 1u) == 0u) {
                    tc_mma_f16_u(tmem_d, da + 2 * ks, db + 2 * ks, idesc, first ? 0u : 1u);
                    first = 0;
                  }
              }
              ++tap;
              da += 8;
              db += kWinBTile / 16;
            }
            da_row += static_cast<uint64_t>(p.Wp) * 8;
          }
          tc_commit_u(&empty_bar[stage]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit_u(&tmem_full[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q4 = warp & 3;
    const int set = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++local) {
      if (local % EPI_SETS != set) continue;
      const int acc = local % kWinAccBufs;
      const uint32_t acc_phase = (local / kWinAccBufs) & 1;
      const long long q = p.q_base + static_cast<long long>(tile) * kBlockM + row;
      bool valid = q < p.M_pad;
      size_t off = 0;
      if (valid) {
        const int img = static_cast<int>(q / hpwp);
        const int rem = static_cast<int>(q - static_cast<long long>(img) * hpwp);
        const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
        const int y = yp - p.pl_h, x = xp - p.pl_w;
        valid = (y >= 0) && (y < p.H) && (x >= 0) && (x < p.W);
        off = valid ? ((static_cast<size_t>(img) * p.H + y) * p.W + x) * kWinN : 0;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(acc * kWinN);
      const act_t* res_row = p.residual ? p.residual + off : nullptr;
      uint4 res_cur[4];
      if (valid && res_row) load_res_chunk(res_row, 0, res_cur);  // in flight while the MMAs finish
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_row<kWinN>(taddr, valid, p.out + off, res_row, bias_s, p.relu, res_cur);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kWinAccBufs * kWinN)
                 : "memory");
  }
}

// bit 0: shared-memory window kernel for the 64 -> 64 stride-1 layers; bit 1: CTA-pair (cta_group::2) kernel for
// C_out >= 128; 0 = single-CTA im2col kernel everywhere
static int g_conv_mode = 60866571;  // 11 + CTA-pair window kernels for layer2 (16384) and stem / layer1 (32768) + max-pool in the stem's epilogue
                                    // (2097152) + sliding window / TMA residual in the 64 -> 64 pair kernel (8388608) + staged epilogues /
                                    // TMA residual in the layer2 (16777216) and layer3-4 (33554432) pair kernels
static int conv_mode() { return g_conv_mode; }
int conv_get_mode() { return g_conv_mode; }
void conv_set_mode(int mode) { g_conv_mode = mode; }

// The 7x7 / stride-2 stem as a 4x4 convolution over the space-to-depth input (backbone.py: _stem_s2d): tap (by, bx), K step
// ks = dy * 2 + dx holds w[:, :, 2 by + dy - 1, 2 bx + dx - 1], which lies outside the 7x7 filter for by = 0, dy = 0 and for
// bx = 0, dx = 0: 15 of the 64 (tap, K step) slices are zero by construction and need no MMA (K = 784 instead of 1024).
// Only valid when every K step is one sub-pixel, i.e. C_in = 4 * 16.
static unsigned long long stem_kskip(const ConvDesc& d) {
  if (!d.s2d_stem || d.R != 4 || d.S != 4 || d.C_in != 64 || (g_conv_mode & 131072) != 0) return 0ull;
  unsigned long long m = 0;
  for (int by = 0; by < 4; ++by)
    for (int bx = 0; bx < 4; ++bx)
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx)
          if ((by == 0 && dy == 0) || (bx == 0 && dx == 0)) m |= 1ull << (4 * (by * 4 + bx) + dy * 2 + dx);
  return m;
}

// Returns MPX_ERR_UNSUPPORTED (without setting an error) when the shape does not fit the window kernel.
static int conv_window_try(const ConvDesc& d, const void* x, const void* w, const float* bias, const void* residual,
                           void* out, int max_ctas, cudaStream_t stream) {
  if ((g_conv_mode & 1) == 0) return MPX_ERR_UNSUPPORTED;
  if (d.stride != 1 || d.C_in != 64 || d.C_out != 64) return MPX_ERR_UNSUPPORTED;
  if (d.R > 4 || d.S > 4 || d.R * d.S > 16) return MPX_ERR_UNSUPPORTED;
  const int P = conv_out_dim(d.H, d.pad_lo_h, d.pad_hi_h, d.R, 1), Q = conv_out_dim(d.W, d.pad_lo_w, d.pad_hi_w, d.S, 1);
  if (P != d.H || Q != d.W) return MPX_ERR_UNSUPPORTED;  // "same" convolutions only
  WinParams p{};
  p.Hp = d.H + d.pad_lo_h + d.pad_hi_h;
  p.Wp = d.W + d.pad_lo_w + d.pad_hi_w;
  p.H = d.H;
  p.W = d.W;
  p.pl_h = d.pad_lo_h;
  p.pl_w = d.pad_lo_w;
  p.n_img = d.n_img;
  p.S = d.S;
  const int n_taps = d.R * d.S;
  const int b_bytes = n_taps * kWinBTile;
  // two epilogue warp sets hide the residual read of conv2 (0.352 -> 0.291 ms at batch 576, tools/gpu_probe_epilogue.py);
  // mode bit 4 (16) selects a single set
  const int epi_sets = (g_conv_mode & 16) != 0 ? 1 : ((g_conv_mode & 128) != 0 ? 3 : 2);
  const int smem_limit = 227 * 1024 - 1024 /*align*/ - 1024 /*barriers, bias*/;
  // Choose the number of filter rows per window.  The two MMA issuers work on alternating tiles, so what matters is how
  // many TILES' worth of windows fit the ring (2 = both issuers always have a resident tile); ties go to the larger row
  // group (fewer halo rows re-loaded).  layer1: rg = 3, one 38 KB window per tile, 4 stages.  Coarse stem (16 resident
  // weight tiles = 128 KB): rg = 2 fits 2 stages = 1 tile, rg = 1 fits 5 windows of 17 KB = 1.25 tiles.
  int rg = 0, stages = 0;
  double best = 0.0;
  for (int cand = d.R; cand >= 1; --cand) {
    if (d.R % cand) continue;
    const int rows = kBlockM + (cand - 1) * p.Wp + (d.S - 1);
    const int n_chunks = (rows + 255) / 256;
    const int chunk = ((rows + n_chunks - 1) / n_chunks + 7) & ~7;
    const int win_bytes = chunk * n_chunks * 128;
    int st = (smem_limit - b_bytes) / win_bytes;
    if (st > 8) st = 8;
    if (st < 2) continue;
    double tiles = static_cast<double>(st) / (d.R / cand);
    if (tiles > 2.0) tiles = 2.0;
    if (tiles > best + 1e-9) {
      best = tiles;
      rg = cand;
      stages = st;
      p.win_rows = rows;
      p.n_chunks = n_chunks;
      p.chunk_rows = chunk;
      p.win_bytes = win_bytes;
    }
  }
  if ((g_conv_mode & 1024) != 0 && d.R % 2 == 0) {  // diagnostic (mode bit 10): force the pre-r01 choice rg = R / 2
    const int cand = d.R / 2;
    const int rows = kBlockM + (cand - 1) * p.Wp + (d.S - 1);
    const int n_chunks = (rows + 255) / 256;
    const int chunk = ((rows + n_chunks - 1) / n_chunks + 7) & ~7;
    const int win_bytes = chunk * n_chunks * 128;
    const int st = (smem_limit - b_bytes) / win_bytes;
    if (st >= 2) {
      rg = cand; stages = st > 8 ? 8 : st;
      p.win_rows = rows; p.n_chunks = n_chunks; p.chunk_rows = chunk; p.win_bytes = win_bytes;
    }
  }
  if (rg < 1 || stages < 2) return MPX_ERR_UNSUPPORTED;
  p.rg = rg;
  p.n_windows = d.R / rg;
  p.taps_per_win = rg * d.S;
  p.M_pad = static_cast<long long>(d.n_img) * p.Hp * p.Wp;
  p.q_base = static_cast<long long>(p.pl_h) * p.Wp + p.pl_w;
  const long long m_tiles = (p.M_pad - p.q_base + kBlockM - 1) / kBlockM;
  if (m_tiles <= 0 || m_tiles >= (1LL << 31)) return MPX_ERR_UNSUPPORTED;
  p.m_tiles = static_cast<int>(m_tiles);
  p.relu = d.relu;
  p.mma_issuers = (g_conv_mode & 32) != 0 ? 1 : ((g_conv_mode & 64) != 0 ? 3 : 2);
  p.observers_arrive = 1;
  p.row_epilogue = 0;
  p.kskip = stem_kskip(d);
  p.bias = bias;
  p.residual = reinterpret_cast<const act_t*>(residual);
  p.out = reinterpret_cast<act_t*>(out);

  int rc = load_driver_entry_points();
  if (rc != MPX_OK) return rc;
  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.n_img)};
    cuuint64_t strides[3] = {128, static_cast<cuuint64_t>(d.W) * 128, static_cast<cuuint64_t>(d.H) * d.W * 128};
    int lower[2] = {-d.pad_lo_w, -d.pad_lo_h};
    int upper[2] = {d.pad_hi_w, d.pad_hi_h};  // the base pixel walks the whole padded image
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_im2col(&map_a, kTmaActType, 4, const_cast<void*>(x), dims, strides, lower,
                                 upper, kBlockK, static_cast<cuuint32_t>(p.chunk_rows), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MPX_ERR_UNSUPPORTED;
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const size_t bytes = static_cast<size_t>(d.n_img) * d.H * d.W * 128;
    if (drv <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&map_a)[1] &= ~(1ull << 21);
  }
  {
    const cuuint64_t K_total = static_cast<cuuint64_t>(n_taps) * 64;
    cuuint64_t dims[2] = {K_total, 64};
    cuuint64_t strides[1] = {K_total * 2};
    cuuint32_t box[2] = {kBlockK, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&map_b, kTmaActType, 2, const_cast<void*>(w), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  }
  const int smem_bytes = 1024 + b_bytes + stages * p.win_bytes + 1024;
  int grid = p.m_tiles;
  const int cap = max_ctas > 0 ? max_ctas : sm_count();
  if (grid > cap) grid = cap;
  ProfileSlot* slot = profile_begin(stream);
  rc = MPX_OK;
#define MPX_WIN_LAUNCH(E)                                                                                        \
  do {                                                                                                           \
    static bool attr_set = false;                                                                                \
    if (!attr_set) {                                                                                             \
      MPX_CHECK_CUDA(cudaFuncSetAttribute(conv_window_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                          227 * 1024));                                                          \
      attr_set = true;                                                                                           \
    }                                                                                                            \
    MPX_CHECK_CUDA(launch_pdl(conv_window_kernel<E>, dim3(grid), dim3(128 + 128 * E), smem_bytes, stream, 1, map_a, map_b, \
                              p, stages, n_taps));                                                             \
  } while (0)
  if (epi_sets == 1) MPX_WIN_LAUNCH(1);
  else if (epi_sets == 2) MPX_WIN_LAUNCH(2);
  else MPX_WIN_LAUNCH(3);
  MPX_CHECK_CUDA(cudaGetLastError());
  ++g_launches;
  profile_end(slot, stream, 2.0 * d.n_img * d.H * d.W * 64.0 * n_taps * 64.0);
  return MPX_OK;
}

// ---------------------------------------------------------------------------------------------
// "Window" convolution for the 128 -> 128 channel 3x3 stride-1 layers (layer2, 7 of its 9 convolutions).
//
// These layers are bound by L2 -> SM bandwidth in the im2col kernel (13.7 TB/s measured, the LTS cap): per 128 x 128
// output tile it re-reads the activations once per filter tap (9 x 32 KB) and the whole 295 KB weight matrix.  Here
//   * a CTA works on a SUPER-TILE of 256 consecutive padded-linear output rows (two 128-row MMA tiles);
//   * the activations of a super-tile are loaded ONCE as a contiguous window (256 + 2*Wp + 2 rows, synthesised from
//     the unpadded tensor by TMA im2col exactly like conv_window_kernel), split into two channel PANELS of 64
//     (each panel is a 128B-swizzled K-major operand; every tap is a row-shifted descriptor into it);
//   * the weights stream through an 8-stage ring of [128 c_out x 64 c_in] tiles, one pass per super-tile, and every
//     ring stage feeds BOTH tiles: two MMA-issuing threads (one per tile, each with its own TMEM accumulator) consume
//     it, which also interleaves two accumulate chains on the tensor pipe;
//   * K order is panel-major (all 9 taps of channels 0-63, then of channels 64-127), so that panel 0 of the next
//     super-tile is loaded while panel 1 of the current one is consumed (and vice versa): double buffering of the
//     window at half granularity with no extra shared memory.
// L2 -> SM bytes per 128 output rows: 576 KB -> 44 + 144 KB.
// Warps: 0 window producer, 1 and 3 MMA issuers (tile 0 / 1), 2 TMEM allocation + weight producer, 4-7 and 8-11 epilogue
// of tile 0 / 1.  TMEM: 2 (super-tiles in flight) x 2 (tiles) accumulators of 128 columns.
// ---------------------------------------------------------------------------------------------
struct Win2Params {
  int Hp, Wp;          // padded image size
  int H, W;
  int n_img;
  int chunk_rows;      // rows per TMA issue (<= 256, multiple of 8)
  int n_chunks;
  int panel_bytes;     // chunk_rows * n_chunks * 128
  long long M_pad;     // n_img * Hp * Wp
  long long q_base;    // Wp + 1
  int n_super;
  int relu;
  const float* bias;
  const act_t* residual;
  act_t* out;
  int staged;      // conv_window2q_kernel: 1 = staged epilogue (coalesced stores, residual rows by TMA), mode bit 24
  int b_stages;    // conv_window2q_kernel: depth of the weight ring (<= kW2qBStages)
};

constexpr int kW2N = 128;                 // C_out = C_in
constexpr int kW2BStages = 8;
constexpr int kW2BTile = kW2N * 128;      // [128 c_out][64 c_in] bf16 = 16 KB
constexpr int kW2Taps = 9;

__global__ void __launch_bounds__(384, 1)
conv_window2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const Win2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_b = smem;                                         // weight ring
  uint8_t* smem_a = smem + kW2BStages * kW2BTile;                 // two window panels
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + 2 * static_cast<size_t>(p.panel_bytes));
  uint64_t* b_full = bars;                // [8]
  uint64_t* b_empty = bars + 8;           // [8]  two arrivals (one commit per issuer)
  uint64_t* a_full = bars + 16;           // [2]
  uint64_t* a_empty = bars + 18;          // [2]  two arrivals
  uint64_t* tmem_full = bars + 20;        // [4]
  uint64_t* tmem_empty = bars + 24;       // [4]  four arrivals (epilogue warps of the owning set)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 28);
  float* bias_s = reinterpret_cast<float*>(bars + 32);  // [128]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction (role dispatch, UR operands)
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < kW2N) bias_s[threadIdx.x] = p.bias[threadIdx.x];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kW2BStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 2);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)),
                 "r"(4 * kW2N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int hpwp = p.Hp * p.Wp;
  pdl_wait();

  if (warp == 0) {
    // ===================== window producer =====================
    {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      int local = 0;
      for (int st = blockIdx.x; st < p.n_super; st += gridDim.x, ++local) {
        const long long qs = p.q_base + static_cast<long long>(st) * 256 - (p.Wp + 1);  // first window row
        const uint32_t par = static_cast<uint32_t>(local & 1);
        for (int c = 0; c < 2; ++c) {
          mbar_wait(&a_empty[c], par ^ 1u);
          mbar_expect_tx_u(&a_full[c], static_cast<uint32_t>(p.panel_bytes));
          for (int ch = 0; ch < p.n_chunks; ++ch) {
            const long long q = qs + static_cast<long long>(ch) * p.chunk_rows;
            const int img = static_cast<int>(q / hpwp);
            const int rem = static_cast<int>(q - static_cast<long long>(img) * hpwp);
            const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
            tma_load_im2col_4d_u(smem_a + static_cast<size_t>(c) * p.panel_bytes + static_cast<size_t>(ch) * p.chunk_rows * 128,
                               &map_a, &a_full[c], c * 64, xp - 1, yp - 1, img, 0, 0);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== weight producer =====================
    {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      int stage = 0;
      uint32_t phase = 0;
      for (int st = blockIdx.x; st < p.n_super; st += gridDim.x) {
        for (int c = 0; c < 2; ++c) {
          for (int t = 0; t < kW2Taps; ++t) {
            mbar_wait(&b_empty[stage], phase ^ 1u);
            mbar_expect_tx_u(&b_full[stage], kW2BTile);
            tma_load_2d_u(smem_b + stage * kW2BTile, &map_b, &b_full[stage], t * 128 + c * 64, 0);
            if (++stage == kW2BStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers: warp 1 -> rows 0-127 of the super-tile, warp 3 -> rows 128-255 ==========
    {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      const int ti = warp == 1 ? 0 : 1;
      constexpr uint32_t idesc = (1u << 4) | kIdescAB | (static_cast<uint32_t>(kW2N >> 3) << 17) |
                                 (static_cast<uint32_t>(kBlockM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int st = blockIdx.x; st < p.n_super; st += gridDim.x, ++local) {
        const int buf = (local & 1) * 2 + ti;
        const uint32_t apar = static_cast<uint32_t>(local & 1);
        mbar_wait(&tmem_empty[buf], (static_cast<uint32_t>(local >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * kW2N);
        uint32_t first = 1;
        for (int c = 0; c < 2; ++c) {
          mbar_wait(&a_full[c], apar);
          tc_fence_after();
          // this tile's rows start 128 rows (16 KB) into the panel; taps add r*Wp + s rows (8 per row in addr>>4 units)
          const uint64_t da_tile = make_sw128_desc(smem_u32(smem_a + static_cast<size_t>(c) * p.panel_bytes) +
                                                   static_cast<uint32_t>(ti) * 128u * 128u);
          uint64_t da_row = da_tile;
          for (int r = 0; r < 3; ++r) {
            uint64_t da = da_row;
            for (int s = 0; s < 3; ++s) {
              mbar_wait(&b_full[stage], phase);
              tc_fence_after();
              const uint64_t db = make_sw128_desc(smem_u32(smem_b + stage * kW2BTile));
              tc_mma_f16_u(tmem_d, da, db, idesc, first ? 0u : 1u);
              tc_mma_f16_u(tmem_d, da + 2, db + 2, idesc, 1u);
              tc_mma_f16_u(tmem_d, da + 4, db + 4, idesc, 1u);
              tc_mma_f16_u(tmem_d, da + 6, db + 6, idesc, 1u);
              first = 0;
              tc_commit_u(&b_empty[stage]);
              if (++stage == kW2BStages) {
                stage = 0;
                phase ^= 1u;
              }
              da += 8;
            }
            da_row += static_cast<uint64_t>(p.Wp) * 8;
          }
          tc_commit_u(&a_empty[c]);  // this issuer is done with panel c
        }
        tc_commit_u(&tmem_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: warps 4-7 tile 0, warps 8-11 tile 1 =====================
    const int q4 = warp & 3;
    const int ti = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;
    int local = 0;
    for (int st = blockIdx.x; st < p.n_super; st += gridDim.x, ++local) {
      const int buf = (local & 1) * 2 + ti;
      const long long q = p.q_base + static_cast<long long>(st) * 256 + ti * 128 + row;
      bool valid = q < p.M_pad;
      size_t off = 0;
      if (valid) {
        const int img = static_cast<int>(q / hpwp);
        const int rem = static_cast<int>(q - static_cast<long long>(img) * hpwp);
        const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
        const int y = yp - 1, x = xp - 1;
        valid = (y >= 0) && (y < p.H) && (x >= 0) && (x < p.W);
        off = valid ? ((static_cast<size_t>(img) * p.H + y) * p.W + x) * kW2N : 0;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(buf * kW2N);
      const act_t* res_row = p.residual ? p.residual + off : nullptr;
      uint4 res_cur[4];
      if (valid && res_row) load_res_chunk(res_row, 0, res_cur);  // in flight while the MMAs finish
      mbar_wait(&tmem_full[buf], static_cast<uint32_t>(local >> 1) & 1u);
      tc_fence_after();
      epilogue_row<kW2N>(taddr, valid, p.out + off, res_row, bias_s, p.relu, res_cur);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(4 * kW2N) : "memory");
  }
}

// Returns MPX_ERR_UNSUPPORTED (without setting an error) when the shape does not fit.
static int conv_window2_try(const ConvDesc& d, const void* x, const void* w, const float* bias, const void* residual,
                            void* out, int max_ctas, cudaStream_t stream) {
  if ((g_conv_mode & 1) == 0 || (g_conv_mode & 256) != 0) return MPX_ERR_UNSUPPORTED;
  if (d.stride != 1 || d.C_in != 128 || d.C_out != 128 || d.R != 3 || d.S != 3) return MPX_ERR_UNSUPPORTED;
  if (d.pad_lo_h != 1 || d.pad_lo_w != 1 || d.pad_hi_h != 1 || d.pad_hi_w != 1) return MPX_ERR_UNSUPPORTED;
  Win2Params p{};
  p.Hp = d.H + 2;
  p.Wp = d.W + 2;
  p.H = d.H;
  p.W = d.W;
  p.n_img = d.n_img;
  const int rows = 256 + 2 * p.Wp + 2;
  p.n_chunks = (rows + 255) / 256;
  p.chunk_rows = ((rows + p.n_chunks - 1) / p.n_chunks + 7) & ~7;
  p.panel_bytes = p.chunk_rows * p.n_chunks * 128;
  const int smem_bytes = 1024 + kW2BStages * kW2BTile + 2 * p.panel_bytes + 1024;
  if (smem_bytes > 227 * 1024) return MPX_ERR_UNSUPPORTED;
  p.M_pad = static_cast<long long>(d.n_img) * p.Hp * p.Wp;
  p.q_base = p.Wp + 1;
  const long long n_super = (p.M_pad - p.q_base + 255) / 256;
  if (n_super <= 0 || n_super >= (1LL << 31)) return MPX_ERR_UNSUPPORTED;
  // a handful of super-tiles (small batches) leaves most SMs idle: the im2col kernel with split-K is the better fit
  if (n_super * 2 < (max_ctas > 0 ? max_ctas : sm_count())) return MPX_ERR_UNSUPPORTED;
  p.n_super = static_cast<int>(n_super);
  p.relu = d.relu;
  p.bias = bias;
  p.residual = reinterpret_cast<const act_t*>(residual);
  p.out = reinterpret_cast<act_t*>(out);

  int rc = load_driver_entry_points();
  if (rc != MPX_OK) return rc;
  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[4] = {128, static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.n_img)};
    cuuint64_t strides[3] = {256, static_cast<cuuint64_t>(d.W) * 256, static_cast<cuuint64_t>(d.H) * d.W * 256};
    int lower[2] = {-1, -1};
    int upper[2] = {1, 1};  // the base pixel walks the whole padded image
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_im2col(&map_a, kTmaActType, 4, const_cast<void*>(x), dims, strides, lower,
                                 upper, kBlockK, static_cast<cuuint32_t>(p.chunk_rows), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MPX_ERR_UNSUPPORTED;
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const size_t bytes = static_cast<size_t>(d.n_img) * d.H * d.W * 256;
    if (drv <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&map_a)[1] &= ~(1ull << 21);
  }
  {
    const cuuint64_t K_total = static_cast<cuuint64_t>(kW2Taps) * 128;
    cuuint64_t dims[2] = {K_total, 128};
    cuuint64_t strides[1] = {K_total * 2};
    cuuint32_t box[2] = {kBlockK, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&map_b, kTmaActType, 2, const_cast<void*>(w), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  }
  static bool attr_set = false;
  if (!attr_set) {
    MPX_CHECK_CUDA(cudaFuncSetAttribute(conv_window2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int grid = p.n_super;
  const int cap = max_ctas > 0 ? max_ctas : sm_count();
  if (grid > cap) grid = cap;
  ProfileSlot* slot = profile_begin(stream);
  MPX_CHECK_CUDA(launch_pdl(conv_window2_kernel, dim3(grid), dim3(384), smem_bytes, stream, 1, map_a, map_b, p));
  MPX_CHECK_CUDA(cudaGetLastError());
  ++g_launches;
  profile_end(slot, stream, 2.0 * d.n_img * d.H * d.W * 128.0 * kW2Taps * 128.0);
  return MPX_OK;
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant of the im2col kernel (tcgen05.mma.cta_group::2): two CTAs of a cluster compute one 256 x BLOCK_N
// tile; each loads its own 128 activation rows and HALF of the weight tile, the leader issues the MMAs that read
// both CTAs' shared memory and write both CTAs' TMEM.  Per CTA and K-block the shared-memory fill drops from
// 16 KB + BLOCK_N*128 B to 16 KB + BLOCK_N*64 B -- the wide layers (C_out >= 128) are L2->SM bandwidth bound.
// Barrier protocol (DeepGEMM / CUTLASS sm100 2-SM pattern):
//   full[s]       lives in the leader; the leader arms expect_tx for both CTAs' bytes, the peer arrives remotely;
//                 both CTAs' TMA loads complete_tx on the leader's barrier (peer bit of the address cleared)
//   empty[s]      per CTA; the leader's tcgen05.commit multicasts the arrive to both
//   tmem_full[a]  per CTA, multicast commit; tmem_empty[a] in the leader, 4 warps of each CTA arrive (peer remotely)
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address
constexpr uint64_t kTmaCacheHintNormal = 0x1000000000000000ull;

__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta_rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta_rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(kTmaCacheHintNormal)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c, int w,
                                                    int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;" ::"r"(smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h),
      "l"(kTmaCacheHintNormal)
      : "memory");
}
__device__ __forceinline__ void tc2_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

__device__ __forceinline__ void mbar_arrive_remote_u(uint64_t* bar, uint32_t cta_rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta_rank));
  asm volatile(MPX_ELECT_PRED "@e mbarrier.arrive.shared::cluster.b64 _, [%0];\n\t}" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d_u(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(MPX_ELECT_PRED
               "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
               " [%0], [%1, {%3, %4}], [%2], %5;\n\t}" ::"r"(smem_u32(smem)),
               "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(kTmaCacheHintNormal)
               : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d_u(void* smem, const CUtensorMap* map, uint64_t* bar, int c, int w, int h,
                                                      int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(MPX_ELECT_PRED
               "@e cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
               " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;\n\t}" ::"r"(smem_u32(smem)),
               "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h),
               "l"(kTmaCacheHintNormal)
               : "memory");
}
__device__ __forceinline__ void tc2_mma_f16_u(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Advance a shared-memory descriptor by `inc` 16-byte units without touching its high word (SBO, version, layout): the start
// address field never carries out of the low word (shared memory is < 256 KB), and a 32-bit add on the low register of the
// uniform pair is half of the 64-bit add-with-carry the compiler emits for `desc + inc`.
__device__ __forceinline__ uint64_t desc_add_lo(uint64_t d, uint32_t inc) {
  return (d & 0xFFFFFFFF00000000ull) | static_cast<uint64_t>(static_cast<uint32_t>(d) + inc);
}
__device__ __forceinline__ void tc2_commit_mc_u(uint64_t* bar) {
  asm volatile(MPX_ELECT_PRED
               "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

template <int BLOCK_N>
struct Conv2Cfg {
  static constexpr int kBHalfBytes = (BLOCK_N / 2) * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBHalfBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 6 : 8;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + 2048;
  // staged epilogue (BLOCK_N = 256): 32 KB of staging (128 rows x 128 channels) between the pipeline stages and the barriers
  static constexpr int kStgBytes = 2 * kBlockM * 128;
  static constexpr int kSmemBytesStaged = kStages * kStageBytes + kStgBytes + 1024 + 512;  // bias read from global memory
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
conv_igemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_r, const ConvParams p) {
  using Cfg = Conv2Cfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kATileBytes;
  const bool staged = BLOCK_N == 256 && p.staged != 0;
  const bool has_res = p.residual != nullptr;
  uint8_t* smem_stg = smem + kStages * Cfg::kStageBytes;  // staged epilogue: two 64-channel panels of 128 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + (staged ? Cfg::kStgBytes : 0));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  float* bias_s = reinterpret_cast<float*>(bars + 32);  // [BLOCK_N <= 256]
  uint64_t* res_full = bars + 2 * kStages + 6;   // staged epilogue: the residual rows of one half tile landed (TMA)
  uint64_t* res_empty = bars + 2 * kStages + 7;  // ... and the four epilogue warps are done with the staging tile

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction (role dispatch, UR operands)
  const int lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;  // = %cluster_ctarank for clusters of two along x; from blockIdx so that the compiler knows it is uniform
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int m_pair_tiles = (p.m_tiles + 1) >> 1;  // 256-row tiles
  const int total_tiles = m_pair_tiles * p.n_tiles;
  if (!staged)  // (the staged epilogue reads the bias from global memory: its 32 KB tile takes the room)
    for (int i = threadIdx.x; i < p.C_out; i += blockDim.x) bias_s[i] = p.bias[i];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 2);   // leader's expect_tx arrive + peer's remote arrive
      mbar_init(&empty_bar[i], 1);  // multicast commit from the leader
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);  // 4 epilogue warps x 2 CTAs
    }
    if (staged) {
      mbar_init(res_full, 1);
      mbar_init(res_empty, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      int stage = 0;
      uint32_t phase = 0;
      const int pq = p.P * p.Q;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const int m_pair = tile / p.n_tiles;
        const int n_tile = tile - m_pair * p.n_tiles;
        const long long m0 = static_cast<long long>(m_pair) * 256 + static_cast<long long>(rank) * kBlockM;
        const int img = static_cast<int>(m0 / pq);
        const int rem = static_cast<int>(m0 - static_cast<long long>(img) * pq);
        const int p0 = rem / p.Q;
        const int q0 = rem - p0 * p.Q;
        const int base_w = q0 * p.stride - p.pad_w;
        const int base_h = p0 * p.stride - p.pad_h;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (leader) mbar_expect_tx_u(&full_bar[stage], 2u * Cfg::kStageBytes);
          else mbar_arrive_remote_u(&full_bar[stage], 0);
          const int r = tap / p.S;
          const int s = tap - r * p.S;
          tma2_load_im2col_4d_u(smem_a + stage * kATileBytes, &map_a, &full_bar[stage], cb * kBlockK, base_w, base_h, img,
                              static_cast<uint16_t>(s), static_cast<uint16_t>(r));
          tma2_load_2d_u(smem_b + stage * Cfg::kBHalfBytes, &map_b, &full_bar[stage], kb * kBlockK,
                       n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2));
          if (++cb == p.cblocks) {
            cb = 0;
            ++tap;
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      constexpr uint32_t idesc = (1u << 4) | kIdescAB | (static_cast<uint32_t>(BLOCK_N >> 3) << 17) |
                                 (static_cast<uint32_t>(256 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sw128_desc(smem_u32(smem_a + stage * kATileBytes));
          const uint64_t db = make_sw128_desc(smem_u32(smem_b + stage * Cfg::kBHalfBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            tc2_mma_f16_u(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc2_commit_mc_u(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc2_commit_mc_u(&tmem_full[acc]);
      }
    }
  } else if (warp == 3) {
    // ===================== residual producer (staged epilogue, both CTAs) =====================
    // One half tile (128 rows x 128 channels of the [M, C_out] residual matrix, two 64-channel panels) at a time into the
    // single staging tile: the epilogue frees it after the stores of the previous half.
    if (staged && has_res) {
      uint32_t use = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const int m_pair = tile / p.n_tiles;
        const int n_tile = tile - m_pair * p.n_tiles;
        const int m0 = m_pair * 256 + static_cast<int>(rank) * kBlockM;
        for (int h = 0; h < 2; ++h, ++use) {
          mbar_wait(res_empty, (use & 1u) ^ 1u);
          mbar_expect_tx_u(res_full, static_cast<uint32_t>(Cfg::kStgBytes));
          for (int c = 0; c < 2; ++c)
            tma_load_2d_u(smem_stg + c * kBlockM * 128, &map_r, res_full, n_tile * BLOCK_N + h * 128 + c * 64, m0);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    int local = 0;
    uint32_t use = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++local) {
      const int m_pair = tile / p.n_tiles;
      const int n_tile = tile - m_pair * p.n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const long long m = static_cast<long long>(m_pair) * 256 + static_cast<long long>(rank) * kBlockM + row;
      const bool valid = m < p.M_total;
      const int n0 = n_tile * BLOCK_N;
      if (staged) {
        // 128 channels at a time through the staging tile (residual rows put there by TMA, overwritten in place, stored as
        // whole 128-byte lines); the accumulator is released after the second half has been read
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
        const uint32_t stg = smem_u32(smem_stg) + static_cast<uint32_t>(q4 * kStageTileBytes);
        for (int h = 0; h < 2; ++h, ++use) {
          if (has_res) mbar_wait(res_full, use & 1u);
          if (h == 0) {
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
          }
          const float* bias_h = p.bias + n0 + h * 128;
          epilogue_compute64(taddr0 + static_cast<uint32_t>(h * 128), stg, lane, has_res, 0u, p.relu, bias_h);
          epilogue_compute64(taddr0 + static_cast<uint32_t>(h * 128 + 64), stg + kBlockM * 128u, lane, has_res, 0u, p.relu, bias_h + 64);
          if (h == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (leader) mbar_arrive(&tmem_empty[acc]);
              else mbar_arrive_remote(&tmem_empty[acc], 0);
            }
          }
          act_t* out_h = p.out + n0 + h * 128;
          store_staged64_rt(stg, lane, valid ? m : -1, out_h, p.C_out);
          store_staged64_rt(stg + kBlockM * 128u, lane, valid ? m : -1, out_h + 64, p.C_out);
          if (has_res) {  // generic-proxy writes into the tile before the next TMA write
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(res_empty);
          }
        }
        continue;
      }
      const size_t off = static_cast<size_t>(valid ? m : 0) * p.C_out + n0;
      const act_t* res_row = p.residual ? p.residual + off : nullptr;
      uint4 res_cur[4];
      if (valid && res_row) load_res_chunk(res_row, 0, res_cur);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
      epilogue_row<BLOCK_N>(taddr, valid, p.out + off, res_row, bias_s + n0, p.relu, res_cur);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_remote(&tmem_empty[acc], 0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's shared memory / TMEM must stay alive until the leader's MMAs have retired
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols)
                 : "memory");
  }
}

template <int BLOCK_N>
static int launch_conv2(const CUtensorMap& ma, const CUtensorMap& mb, const ConvParams& p_in, cudaStream_t stream,
                        int max_ctas) {
  using Cfg = Conv2Cfg<BLOCK_N>;
  ConvParams p = p_in;
  // staged epilogue (mode bit 25): pays off when a tile's MMAs (num_k_blocks x 512 cycles) cover the two serial half-tile passes;
  // the 1x1 downsample convolutions (2-4 k-blocks) keep the row-per-thread form (measured: 0.048 -> 0.058 ms staged)
  p.staged = (BLOCK_N == 256 && (g_conv_mode & 33554432) != 0 && p.num_k_blocks >= 12) ? 1 : 0;
  const int smem_bytes = p.staged ? Cfg::kSmemBytesStaged : Cfg::kSmemBytes;
  static bool attr_set = false;
  if (!attr_set) {
    MPX_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm2_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytesStaged > Cfg::kSmemBytes ? Cfg::kSmemBytesStaged : Cfg::kSmemBytes));
    attr_set = true;
  }
  CUtensorMap mr = ma;  // residual rows as 2-D tiles of the [M, C_out] matrix (only read by the staged epilogue)
  if (p.staged && p.residual != nullptr) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(p.C_out), static_cast<cuuint64_t>(p.M_total)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(p.C_out) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(kBlockM)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&mr, kTmaActType, 2, const_cast<act_t*>(p.residual), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (residual) failed (%d)", static_cast<int>(r));
  }
  const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  int cap = (max_ctas > 0 ? max_ctas : sm_count()) / 2;
  if (cap < 1) cap = 1;
  const int pairs = pair_tiles < cap ? pair_tiles : cap;
  ProfileSlot* slot = profile_begin(stream);
  conv_igemm2_kernel<BLOCK_N><<<2 * pairs, 256, smem_bytes, stream>>>(ma, mb, mr, p);
  MPX_CHECK_CUDA(cudaGetLastError());
  ++g_launches;
  profile_end(slot, stream, 2.0 * p.M_total * p.C_out * p.num_k_blocks * kBlockK);
  return MPX_OK;
}

// ---------------------------------------------------------------------------------------------
// Mode bit 14 = 16384 (default since r02: layer2 0.167 -> 0.153 ms, with residual 0.206 -> 0.185 ms at batch 576): conv_window2_kernel on CTA PAIRS.  Same work
// decomposition per CTA as conv_window2_kernel (a super-tile of 256 padded-linear rows = two MMA tiles, the window loaded
// once as two 64-channel panels, weights streamed through a ring that feeds both tiles, two MMA-issuing threads), but
// the two CTAs of a cluster run their super-tiles in lockstep and the LEADER's two issuers drive tcgen05.mma.cta_group::2:
// tile ti of the pair = rows ti*128.. of BOTH CTAs (M = 256), each CTA holds only HALF of every weight tile (64 of the 128
// output channels, 8 KB).  Per SM and MMA the tensor core then reads 4 KB (A) + 2 KB (B) of shared memory instead of
// 4 + 4 KB -- the operand rate, not the tensor pipe, bounds conv_window2_kernel -- and the weight
// ring shrinks from 128 to 96 KB at 12 stages.  Barrier protocol as in conv_igemm2_kernel: *_full in the leader (two
// arrivals, both CTAs' bytes), *_empty and tmem_full per CTA through multicast commits (two issuers -> count 2 on the
// empties), tmem_empty in the leader (4 epilogue warps x 2 CTAs).
// ---------------------------------------------------------------------------------------------
constexpr int kW2qBStages = 12;
constexpr int kW2qBHalf = (kW2N / 2) * 128;  // [64 c_out][64 c_in] bf16 = 8 KB per CTA and stage

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
conv_window2q_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_r, const Win2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_b = smem;                                         // weight ring (this CTA's half tiles)
  const int b_stages = p.b_stages;
  uint8_t* smem_a = smem + b_stages * kW2qBHalf;                  // two window panels (this CTA's 256 rows + halo)
  // staged epilogue (p.staged): per tile of the super-tile two 64-channel panels of 128 rows x 128 B, 128B-swizzled -- the layout
  // TMA writes the residual rows in and the epilogue overwrites in place before storing whole 128-byte lines
  uint8_t* smem_stg = smem_a + 2 * static_cast<size_t>(p.panel_bytes);
  const bool staged = p.staged != 0;
  const bool has_res = p.residual != nullptr;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + (staged ? 2 * 2 * kBlockM * 128 : 0));
  uint64_t* b_full = bars;                 // [12] leader
  uint64_t* b_empty = bars + 12;           // [12] per CTA, two arrivals (one multicast commit per issuer)
  uint64_t* a_full = bars + 24;            // [2]  leader
  uint64_t* a_empty = bars + 26;           // [2]  per CTA, two arrivals
  uint64_t* tmem_full = bars + 28;         // [4]  per CTA
  uint64_t* tmem_empty = bars + 32;        // [4]  leader, eight arrivals
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 36);
  uint64_t* res_full = bars + 37;          // [2]  per CTA: residual rows of tile ti landed (staged epilogue)
  float* bias_s = reinterpret_cast<float*>(bars + 40);  // [128]
  uint64_t* res_empty = bars + 104;        // [2]  per CTA, the four warps of tile ti

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction (role dispatch, UR operands)
  const int lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;  // = %cluster_ctarank for clusters of two along x; from blockIdx so that the compiler knows it is uniform
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int n_items = (p.n_super + 1) >> 1;  // an item = two super-tiles, one per CTA
  if (threadIdx.x < kW2N) bias_s[threadIdx.x] = p.bias[threadIdx.x];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < b_stages; ++i) {
      mbar_init(&b_full[i], 2);
      mbar_init(&b_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&res_empty[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 2);
      mbar_init(&a_empty[i], 2);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(4 * kW2N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int hpwp = p.Hp * p.Wp;

  if (warp == 0) {
    // ===================== window producer (both CTAs: own super-tile) =====================
    {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      int local = 0;
      for (int item = pair; item < n_items; item += n_pairs, ++local) {
        const long long st = 2LL * item + rank;
        const long long qs = p.q_base + st * 256 - (p.Wp + 1);  // first window row (rows past the tensor load as zeros)
        const uint32_t par = static_cast<uint32_t>(local & 1);
        for (int c = 0; c < 2; ++c) {
          mbar_wait(&a_empty[c], par ^ 1u);
          if (leader) mbar_expect_tx_u(&a_full[c], 2u * static_cast<uint32_t>(p.panel_bytes));
          else mbar_arrive_remote_u(&a_full[c], 0);
          for (int ch = 0; ch < p.n_chunks; ++ch) {
            const long long q = qs + static_cast<long long>(ch) * p.chunk_rows;
            const int img = static_cast<int>(q / hpwp);
            const int rem = static_cast<int>(q - static_cast<long long>(img) * hpwp);
            const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
            tma2_load_im2col_4d_u(smem_a + static_cast<size_t>(c) * p.panel_bytes + static_cast<size_t>(ch) * p.chunk_rows * 128,
                                &map_a, &a_full[c], c * 64, xp - 1, yp - 1, img, 0, 0);
          }
        }
        if (staged && has_res) {
          // residual rows of the item's two tiles: the same 128 padded-linear positions of the residual tensor (ring
          // positions zero-filled), both 64-channel panels, into the staging tiles the epilogue of the previous item has freed
          for (int t2 = 0; t2 < 2; ++t2) {
            const long long q = p.q_base + st * 256 + t2 * 128;
            const int img = static_cast<int>(q / hpwp);
            const int rem = static_cast<int>(q - static_cast<long long>(img) * hpwp);
            const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
            mbar_wait(&res_empty[t2], par ^ 1u);
            mbar_expect_tx_u(&res_full[t2], 2u * kBlockM * 128u);
            for (int c = 0; c < 2; ++c)
              tma_load_im2col_4d_u(smem_stg + static_cast<size_t>(t2 * 2 + c) * kBlockM * 128, &map_r, &res_full[t2], c * 64,
                                   xp - 1, yp - 1, img, 0, 0);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== weight producer (both CTAs: own 64 output channels) =====================
    {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      int stage = 0;
      uint32_t phase = 0;
      for (int item = pair; item < n_items; item += n_pairs) {
        for (int c = 0; c < 2; ++c) {
          for (int t = 0; t < kW2Taps; ++t) {
            mbar_wait(&b_empty[stage], phase ^ 1u);
            if (leader) mbar_expect_tx_u(&b_full[stage], 2u * kW2qBHalf);
            else mbar_arrive_remote_u(&b_full[stage], 0);
            tma2_load_2d_u(smem_b + stage * kW2qBHalf, &map_b, &b_full[stage], t * 128 + c * 64,
                         static_cast<int>(rank) * (kW2N / 2));
            if (++stage == b_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers (leader only): warp 1 -> rows 0-127 of both CTAs, warp 3 -> rows 128-255 ======
    if (leader) {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      const int ti = warp == 1 ? 0 : 1;
      constexpr uint32_t idesc = (1u << 4) | kIdescAB | (static_cast<uint32_t>(kW2N >> 3) << 17) |
                                 (static_cast<uint32_t>(256 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int item = pair; item < n_items; item += n_pairs, ++local) {
        const int buf = (local & 1) * 2 + ti;
        const uint32_t apar = static_cast<uint32_t>(local & 1);
        mbar_wait(&tmem_empty[buf], (static_cast<uint32_t>(local >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * kW2N);
        uint32_t first = 1;
        for (int c = 0; c < 2; ++c) {
          mbar_wait(&a_full[c], apar);
          tc_fence_after();
          const uint64_t da_tile = make_sw128_desc(smem_u32(smem_a + static_cast<size_t>(c) * p.panel_bytes) +
                                                   static_cast<uint32_t>(ti) * 128u * 128u);
          uint64_t da_row = da_tile;
          for (int r = 0; r < 3; ++r) {
            uint64_t da = da_row;
            for (int s = 0; s < 3; ++s) {
              mbar_wait(&b_full[stage], phase);
              tc_fence_after();
              const uint64_t db = make_sw128_desc(smem_u32(smem_b + stage * kW2qBHalf));
              tc2_mma_f16_u(tmem_d, da, db, idesc, first ? 0u : 1u);
              tc2_mma_f16_u(tmem_d, da + 2, db + 2, idesc, 1u);
              tc2_mma_f16_u(tmem_d, da + 4, db + 4, idesc, 1u);
              tc2_mma_f16_u(tmem_d, da + 6, db + 6, idesc, 1u);
              first = 0;
              tc2_commit_mc_u(&b_empty[stage]);
              if (++stage == b_stages) {
                stage = 0;
                phase ^= 1u;
              }
              da += 8;
            }
            da_row += static_cast<uint64_t>(p.Wp) * 8;
          }
          tc2_commit_mc_u(&a_empty[c]);  // this issuer is done with panel c (in both CTAs)
        }
        tc2_commit_mc_u(&tmem_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): warps 4-7 tile 0, warps 8-11 tile 1 of the own super-tile ==========
    const int q4 = warp & 3;
    const int ti = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;
    int local = 0;
    for (int item = pair; item < n_items; item += n_pairs, ++local) {
      const int buf = (local & 1) * 2 + ti;
      const long long q = p.q_base + (2LL * item + rank) * 256 + ti * 128 + row;
      bool valid = q < p.M_pad;
      size_t off = 0;
      if (valid) {
        const int img = static_cast<int>(q / hpwp);
        const int rem = static_cast<int>(q - static_cast<long long>(img) * hpwp);
        const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
        const int y = yp - 1, x = xp - 1;
        valid = (y >= 0) && (y < p.H) && (x >= 0) && (x < p.W);
        off = valid ? ((static_cast<size_t>(img) * p.H + y) * p.W + x) * kW2N : 0;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(buf * kW2N);
      if (staged) {
        const int pix = valid ? static_cast<int>(off / kW2N) : -1;
        const uint32_t stg = smem_u32(smem_stg) + static_cast<uint32_t>(ti * 2 * kBlockM * 128 + q4 * kStageTileBytes);
        if (has_res) mbar_wait(&res_full[ti], static_cast<uint32_t>(local & 1));
        mbar_wait(&tmem_full[buf], static_cast<uint32_t>(local >> 1) & 1u);
        tc_fence_after();
        epilogue_compute64(taddr, stg, lane, has_res, smem_u32(bias_s), p.relu);
        epilogue_compute64(taddr + 64u, stg + kBlockM * 128u, lane, has_res, smem_u32(bias_s + 64), p.relu);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tmem_empty[buf]);
          else mbar_arrive_remote(&tmem_empty[buf], 0);
        }
        store_staged64<kW2N>(stg, lane, pix, p.out);
        store_staged64<kW2N>(stg + kBlockM * 128u, lane, pix, p.out + 64);
        if (has_res) {  // generic-proxy writes into the tile before the next TMA write
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&res_empty[ti]);
        }
        continue;
      }
      const act_t* res_row = p.residual ? p.residual + off : nullptr;
      uint4 res_cur[4];
      if (valid && res_row) load_res_chunk(res_row, 0, res_cur);
      mbar_wait(&tmem_full[buf], static_cast<uint32_t>(local >> 1) & 1u);
      tc_fence_after();
      epilogue_row<kW2N>(taddr, valid, p.out + off, res_row, bias_s, p.relu, res_cur);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[buf]);
        else mbar_arrive_remote(&tmem_empty[buf], 0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's shared memory / TMEM must stay alive until the leader's MMAs have retired
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(4 * kW2N) : "memory");
  }
}

// Returns MPX_ERR_UNSUPPORTED (without setting an error) when the shape does not fit or the mode bit is off.
static int conv_window2q_try(const ConvDesc& d, const void* x, const void* w, const float* bias, const void* residual,
                             void* out, int max_ctas, cudaStream_t stream) {
  if ((g_conv_mode & 16384) == 0) return MPX_ERR_UNSUPPORTED;
  if (d.stride != 1 || d.C_in != 128 || d.C_out != 128 || d.R != 3 || d.S != 3) return MPX_ERR_UNSUPPORTED;
  if (d.pad_lo_h != 1 || d.pad_lo_w != 1 || d.pad_hi_h != 1 || d.pad_hi_w != 1) return MPX_ERR_UNSUPPORTED;
  Win2Params p{};
  p.Hp = d.H + 2;
  p.Wp = d.W + 2;
  p.H = d.H;
  p.W = d.W;
  p.n_img = d.n_img;
  const int rows = 256 + 2 * p.Wp + 2;
  p.n_chunks = (rows + 255) / 256;
  p.chunk_rows = ((rows + p.n_chunks - 1) / p.n_chunks + 7) & ~7;
  p.panel_bytes = p.chunk_rows * p.n_chunks * 128;
  // mode bit 24: staged epilogue (64 KB of staging tiles; the weight ring gives up stages for them)
  p.staged = (g_conv_mode & 16777216) ? 1 : 0;
  p.b_stages = kW2qBStages;
  const int stg_bytes = p.staged ? 2 * 2 * kBlockM * 128 : 0;
  while (p.b_stages > 6 && 1024 + p.b_stages * kW2qBHalf + 2 * p.panel_bytes + stg_bytes + 1024 > 227 * 1024) --p.b_stages;
  const int smem_bytes = 1024 + p.b_stages * kW2qBHalf + 2 * p.panel_bytes + stg_bytes + 1024;
  if (smem_bytes > 227 * 1024) return MPX_ERR_UNSUPPORTED;
  p.M_pad = static_cast<long long>(d.n_img) * p.Hp * p.Wp;
  p.q_base = p.Wp + 1;
  const long long n_super = (p.M_pad - p.q_base + 255) / 256;
  const int sms = max_ctas > 0 ? max_ctas : sm_count();
  if (n_super <= 0 || n_super >= (1LL << 30) || n_super * 2 < sms) return MPX_ERR_UNSUPPORTED;
  p.n_super = static_cast<int>(n_super);
  p.relu = d.relu;
  p.bias = bias;
  p.residual = reinterpret_cast<const act_t*>(residual);
  p.out = reinterpret_cast<act_t*>(out);

  int rc = load_driver_entry_points();
  if (rc != MPX_OK) return rc;
  CUtensorMap map_a, map_b, map_r;
  for (int which = 0; which < 2; ++which) {  // activation windows (chunk_rows pixels per load); residual tiles (128 pixels)
    CUtensorMap& m = which == 0 ? map_a : map_r;
    const void* base = which == 0 ? x : (residual != nullptr ? residual : x);
    cuuint64_t dims[4] = {128, static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.n_img)};
    cuuint64_t strides[3] = {256, static_cast<cuuint64_t>(d.W) * 256, static_cast<cuuint64_t>(d.H) * d.W * 256};
    int lower[2] = {-1, -1};
    int upper[2] = {1, 1};  // the base pixel walks the whole padded image
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_im2col(&m, kTmaActType, 4, const_cast<void*>(base), dims, strides, lower, upper, kBlockK,
                                 static_cast<cuuint32_t>(which == 0 ? p.chunk_rows : kBlockM), estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MPX_ERR_UNSUPPORTED;
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const size_t bytes = static_cast<size_t>(d.n_img) * d.H * d.W * 256;
    if (drv <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
  }
  {
    const cuuint64_t K_total = static_cast<cuuint64_t>(kW2Taps) * 128;
    cuuint64_t dims[2] = {K_total, 128};
    cuuint64_t strides[1] = {K_total * 2};
    cuuint32_t box[2] = {kBlockK, kW2N / 2};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&map_b, kTmaActType, 2, const_cast<void*>(w), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  }
  static bool attr_set = false;
  if (!attr_set) {
    MPX_CHECK_CUDA(cudaFuncSetAttribute(conv_window2q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int items = (p.n_super + 1) / 2;
  int cap = sms / 2;
  if (cap < 1) cap = 1;
  const int pairs = items < cap ? items : cap;
  ProfileSlot* slot = profile_begin(stream);
  conv_window2q_kernel<<<2 * pairs, 384, smem_bytes, stream>>>(map_a, map_b, map_r, p);
  MPX_CHECK_CUDA(cudaGetLastError());
  ++g_launches;
  profile_end(slot, stream, 2.0 * d.n_img * d.H * d.W * 128.0 * kW2Taps * 128.0);
  return MPX_OK;
}

// ---------------------------------------------------------------------------------------------
// Mode bit 15 = 32768 (default since r02: layer1 0.255 -> 0.227 ms, with residual 0.342 -> 0.305 ms at batch 576): conv_window_kernel (64 -> 64 channels: s2d stem,
// layer1) on CTA PAIRS.  A pair tile is 256 padded-linear rows, 128 per CTA; every CTA loads its own windows and keeps
// HALF of the resident weights (32 of the 64 output channels per tap: 4 KB instead of 8 KB, so the stem's 16 taps take 64 KB
// and a deeper window ring fits); the leader's two issuers take pair tiles alternately and issue
// tcgen05.mma.cta_group::2 (M = 256, N = 64).  Per SM and MMA the tensor core reads 4 KB (A) + 1 KB (B) of shared memory
// instead of 4 + 2 KB, and one instruction feeds both SMs' tensor cores -- both the operand-rate floor (77 -> 64 cycles)
// and the per-instruction accumulate-chain cost (halved per SM) of DESIGN.md section 3 move.  Barriers as in
// conv_igemm2_kernel (full / b_full / tmem_empty in the leader, empty / tmem_full per CTA through multicast commits); the
// issuers observe each other's window fills exactly as in conv_window_kernel.  Launched with the cluster and the
// programmatic-dependent-launch attributes (launch_pdl), so it has no __cluster_dims__.
// ---------------------------------------------------------------------------------------------
constexpr int kWinqBHalf = (kWinN / 2) * 128;  // one tap's weights for this CTA: 32 rows x 64 channels bf16 = 4 KB

__global__ void __launch_bounds__(384, 1)
conv_windowq_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const WinParams p, int stages, int n_taps) {
  constexpr int kEpiSets = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_b = smem;                                  // n_taps resident half weight tiles
  uint8_t* smem_a = smem + n_taps * kWinqBHalf;            // `stages` windows
  uint8_t* smem_stg = smem_a + static_cast<size_t>(stages) * p.win_bytes;  // 8 epilogue warps x 4 KB staging tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + 8 * kStageTileBytes);
  uint64_t* full_bar = bars;             // [stages <= 8] leader
  uint64_t* empty_bar = bars + 8;        // [stages]      per CTA
  uint64_t* tmem_full = bars + 16;       // [4]           per CTA
  uint64_t* tmem_empty = bars + 20;      // [4]           leader, eight arrivals
  uint64_t* b_full = bars + 24;          // [1]           leader
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 25);
  float* bias_s = reinterpret_cast<float*>(bars + 32);  // [64]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction (role dispatch, UR operands)
  const int lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;  // = %cluster_ctarank for clusters of two along x; from blockIdx so that the compiler knows it is uniform
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int n_ptiles = (p.m_tiles + 1) >> 1;
  if (threadIdx.x < kWinN) bias_s[threadIdx.x] = p.bias[threadIdx.x];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kWinAccBufs; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    mbar_init(b_full, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(kWinAccBufs * kWinN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int hpwp = p.Hp * p.Wp;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own windows, own half of the weights) =====================
    {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      if (leader) mbar_expect_tx_u(b_full, 2u * static_cast<uint32_t>(n_taps) * kWinqBHalf);
      else mbar_arrive_remote_u(b_full, 0);
      for (int t = 0; t < n_taps; ++t)
        tma2_load_2d_u(smem_b + t * kWinqBHalf, &map_b, b_full, t * kBlockK, static_cast<int>(rank) * (kWinN / 2));
      int stage = 0;
      uint32_t phase = 0;
      // Window wi of a tile starts at padded-linear row (2 * tile + rank) * 128 + wi * rg * Wp; its position (img, y, x) in
      // the padded image space is carried from tile to tile by additions (r02: with two 64-bit divisions per TMA
      // instruction this warp executed ~120 instructions per window and never waited -- it paced the whole kernel)
      const long long q_first = (2LL * pair + rank) * kBlockM;
      int t_img = static_cast<int>(q_first / hpwp);
      int t_y = static_cast<int>((q_first - static_cast<long long>(t_img) * hpwp) / p.Wp);
      int t_x = static_cast<int>(q_first - static_cast<long long>(t_img) * hpwp - static_cast<long long>(t_y) * p.Wp);
      const long long q_step = 2LL * n_pairs * kBlockM;
      const int s_img = static_cast<int>(q_step / hpwp);
      const int s_y = static_cast<int>((q_step - static_cast<long long>(s_img) * hpwp) / p.Wp);
      const int s_x = static_cast<int>(q_step - static_cast<long long>(s_img) * hpwp - static_cast<long long>(s_y) * p.Wp);
      const int c_y = p.chunk_rows / p.Wp, c_x = p.chunk_rows - c_y * p.Wp;
      for (int tile = pair; tile < n_ptiles; tile += n_pairs) {
        int w_img = t_img, w_y = t_y;
        for (int wi = 0; wi < p.n_windows; ++wi) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (leader) mbar_expect_tx_u(&full_bar[stage], 2u * static_cast<uint32_t>(p.win_bytes));
          else mbar_arrive_remote_u(&full_bar[stage], 0);
          int img = w_img, yp = w_y, xp = t_x;
          for (int ch = 0; ch < p.n_chunks; ++ch) {
            tma2_load_im2col_4d_u(smem_a + static_cast<size_t>(stage) * p.win_bytes + static_cast<size_t>(ch) * p.chunk_rows * 128,
                                &map_a, &full_bar[stage], 0, xp - p.pl_w, yp - p.pl_h, img, 0, 0);
            xp += c_x;
            if (xp >= p.Wp) {
              xp -= p.Wp;
              ++yp;
            }
            yp += c_y;
            while (yp >= p.Hp) {
              yp -= p.Hp;
              ++img;
            }
          }
          w_y += p.rg;
          while (w_y >= p.Hp) {
            w_y -= p.Hp;
            ++w_img;
          }
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        t_x += s_x;
        if (t_x >= p.Wp) {
          t_x -= p.Wp;
          ++t_y;
        }
        t_y += s_y;
        while (t_y >= p.Hp) {
          t_y -= p.Hp;
          ++t_img;
        }
        t_img += s_img;
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers (leader only), pair tiles alternately =====================
    const int which = warp == 1 ? 0 : 1;
    if (leader) {  // whole warp, warp-uniform operands; one lane is elected inside each instruction
      const uint32_t idesc = p.idesc;
      mbar_wait(b_full, 0);
      // Every per-MMA operand below is a function of kernel parameters and loop counters only (tmem_base is known to be 0,
      // checked above), so ptxas keeps the whole issue loop in uniform registers: ~5 instructions per MMA instead of ~10
      // (r02 ncu: the two issuers executed 14-19 instructions per MMA and were busy 85-100% of the time).
      const uint32_t row_step = static_cast<uint32_t>(p.Wp) * 8u - 8u * static_cast<uint32_t>(p.S);
      const uint64_t b_base = make_sw128_desc(smem_u32(smem_b));
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = pair; tile < n_ptiles; tile += n_pairs, ++local) {
        if ((local & 1) != which) {
          // the other issuer's tile: only observe its window fills (see conv_window_kernel)
          for (int wi = 0; wi < p.n_windows; ++wi) {
            mbar_wait(&full_bar[stage], phase);
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          continue;
        }
        const int acc = local % kWinAccBufs;
        const uint32_t acc_phase = (local / kWinAccBufs) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = static_cast<uint32_t>(acc * kWinN);
        uint32_t first = 1;
        uint64_t db = b_base;  // walks the resident half weight tiles: +2 per K step of 16 channels, 256 per tap
        int tap = 0;
        for (int wi = 0; wi < p.n_windows; ++wi) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          uint64_t da = make_sw128_desc(smem_u32(smem_a + static_cast<size_t>(stage) * p.win_bytes));
          for (int r = 0; r < p.rg; ++r) {
            for (int s = 0; s < p.S; ++s) {
              const unsigned sk = static_cast<unsigned>(p.kskip >> (4 * tap)) & 15u;
              if (sk == 0u) {
                tc2_mma_f16_u(tmem_d, da, db, idesc, first ? 0u : 1u);
                da = desc_add_lo(da, 2u);
                db = desc_add_lo(db, 2u);
                tc2_mma_f16_u(tmem_d, da, db, idesc, 1u);
                da = desc_add_lo(da, 2u);
                db = desc_add_lo(db, 2u);
                tc2_mma_f16_u(tmem_d, da, db, idesc, 1u);
                da = desc_add_lo(da, 2u);
                db = desc_add_lo(db, 2u);
                tc2_mma_f16_u(tmem_d, da, db, idesc, 1u);
                da = desc_add_lo(da, 2u);                      // next tap: one pixel (128 B) further
                db = desc_add_lo(db, kWinqBHalf / 16 - 6u);
                first = 0;
              } else {  // structurally zero weight slices (7x7 stem inside its 8x8 space-to-depth footprint): not issued
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (((sk >> ks) & 1u) == 0u) {
                    tc2_mma_f16_u(tmem_d, desc_add_lo(da, 2u * ks), desc_add_lo(db, 2u * ks), idesc, first ? 0u : 1u);
                    first = 0;
                  }
                da = desc_add_lo(da, 8u);
                db = desc_add_lo(db, kWinqBHalf / 16);
              }
              ++tap;
            }
            da = desc_add_lo(da, row_step);  // next filter row: Wp pixels further, minus the S taps already walked
          }
          tc2_commit_mc_u(&empty_bar[stage]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc2_commit_mc_u(&tmem_full[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows), two warp sets, tiles round-robin =====================
    // The MMA issuers address the accumulators from column 0 (uniform-register arithmetic only): one allocation by a CTA
    // that owns its SM (227 KB of shared memory) starts there; checked here, where tmem_base is a per-thread value anyway.
    if (tmem_base != 0u) __trap();
    const int q4 = warp & 3;
    const int set = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;
    // index of this lane's output pixel (row of the [M, 64] output matrix) in a pair tile, -1 for ring / out-of-range rows
    const unsigned hpwp_u = static_cast<unsigned>(hpwp), Wp_u = static_cast<unsigned>(p.Wp);
    int pool_info = 0, pool_base = 0;  // fused max-pool (p.pool): this lane's row as seen by pool_staged64
    auto pix_of = [&](int tile) -> int {
      pool_info = 0;
      const long long q = p.q_base + (2LL * tile + rank) * kBlockM + row;
      if (q >= p.M_pad) return -1;
      const unsigned qu = static_cast<unsigned>(q);  // M_pad < 2^31 (checked on the host)
      const unsigned img = qu / hpwp_u;
      const unsigned rem = qu - img * hpwp_u;
      const unsigned yp = rem / Wp_u, xp = rem - yp * Wp_u;
      const int y = static_cast<int>(yp) - p.pl_h, x = static_cast<int>(xp) - p.pl_w;
      if (y < 0 || y >= p.H || x < 0 || x >= p.W) return -1;
      if (p.pool) {
        int j = x >> 1;
        if ((x & 1) == 0) {  // centre of pooled column x / 2
          pool_info = 1 | ((x > 0 && lane > 0) ? 4 : 0) | ((x + 1 < p.W && lane < 31) ? 8 : 0);
        } else if (lane == 0) {  // its centre x - 1 is the last row of the previous warp tile
          pool_info = 2;
        } else if (lane == 31 && x + 1 < p.W) {  // its centre x + 1 is the first row of the next warp tile
          pool_info = 2;
          j += 1;
        }
        if ((y & 1) && (y >> 1) + 1 < p.pool_H) pool_info |= 16;
        pool_base = (static_cast<int>(img) * p.pool_H + (y >> 1)) * p.pool_W + j;  // pooled pixel index
      }
      return (static_cast<int>(img) * p.H + y) * p.W + x;
    };
    int local = 0;
    for (int tile = pair; tile < n_ptiles; tile += n_pairs, ++local) {
      if (local % kEpiSets != set) continue;
      const int acc = local % kWinAccBufs;
      const uint32_t acc_phase = (local / kWinAccBufs) & 1;
      const int pix = pix_of(tile);
      const bool valid = pix >= 0;
      const size_t off = valid ? static_cast<size_t>(pix) * kWinN : 0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(acc * kWinN);
      if (p.row_epilogue) {  // mode bit 20: one output row per thread, 16-byte accesses scattered over 32 lines (r01 form)
        const act_t* res_row = p.residual ? p.residual + off : nullptr;
        uint4 res_cur[4];
        if (valid && res_row) load_res_chunk(res_row, 0, res_cur);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        epilogue_row<kWinN>(taddr, valid, p.out + off, res_row, bias_s, p.relu, res_cur);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tmem_empty[acc]);
          else mbar_arrive_remote(&tmem_empty[acc], 0);
        }
        continue;
      }
      const uint32_t stg = smem_u32(smem_stg) + static_cast<uint32_t>((warp - 4) * kStageTileBytes);
      if (p.residual) {
        // the residual rows of this warp set's NEXT tile are pulled into L2 now (ncu: 40% of the epilogue warps' samples
        // sat on the first use of the residual loads below), this tile's are staged while its MMAs run
        const int tile_n = tile + kEpiSets * n_pairs;
        const int pix_n = tile_n < n_ptiles ? pix_of(tile_n) : -1;
        if (pix_n >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + static_cast<size_t>(pix_n) * kWinN));
        stage_residual64(stg, lane, pix, p.residual);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_compute64(taddr, stg, lane, p.residual != nullptr, smem_u32(bias_s), p.relu);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // the accumulator is free before the global stores are issued
        if (leader) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_remote(&tmem_empty[acc], 0);
      }
      if (p.pool) pool_staged64(stg, smem_u32(bars + 64) + static_cast<uint32_t>((warp - 4) * 32), lane, pool_info, pool_base, p.pool_W * kWinN, p.out);
      else store_staged64(stg, lane, pix, p.out);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kWinAccBufs * kWinN)
                 : "memory");
  }
}

// Returns MPX_ERR_UNSUPPORTED (without setting an error) when the shape does not fit or the mode bit is off.
static int conv_windowq_try(const ConvDesc& d, const void* x, const void* w, const float* bias, const void* residual,
                            void* out, int max_ctas, cudaStream_t stream) {
  if ((g_conv_mode & 32768) == 0) return MPX_ERR_UNSUPPORTED;
  if (d.stride != 1 || d.C_in != 64 || d.C_out != 64) return MPX_ERR_UNSUPPORTED;
  if (d.R > 4 || d.S > 4 || d.R * d.S > 16) return MPX_ERR_UNSUPPORTED;
  if (d.pool && (residual != nullptr || !d.relu || (g_conv_mode & 1048576) != 0)) return MPX_ERR_UNSUPPORTED;
  const int P = conv_out_dim(d.H, d.pad_lo_h, d.pad_hi_h, d.R, 1), Q = conv_out_dim(d.W, d.pad_lo_w, d.pad_hi_w, d.S, 1);
  if (P != d.H || Q != d.W) return MPX_ERR_UNSUPPORTED;
  WinParams p{};
  p.Hp = d.H + d.pad_lo_h + d.pad_hi_h;
  p.Wp = d.W + d.pad_lo_w + d.pad_hi_w;
  p.H = d.H;
  p.W = d.W;
  p.pl_h = d.pad_lo_h;
  p.pl_w = d.pad_lo_w;
  p.n_img = d.n_img;
  p.S = d.S;
  const int n_taps = d.R * d.S;
  const int b_bytes = n_taps * kWinqBHalf;
  const int smem_limit = 227 * 1024 - 1024 /*align*/ - 1024 /*barriers, bias*/;
  // same choice as conv_window_try: the row group with the most tiles' worth of windows in the ring (at most two)
  int rg = 0, stages = 0;
  double best = 0.0;
  for (int cand = d.R; cand >= 1; --cand) {
    if (d.R % cand) continue;
    const int rows = kBlockM + (cand - 1) * p.Wp + (d.S - 1);
    const int n_chunks = (rows + 255) / 256;
    const int chunk = ((rows + n_chunks - 1) / n_chunks + 7) & ~7;
    const int win_bytes = chunk * n_chunks * 128;
    int st = (smem_limit - b_bytes - 8 * kStageTileBytes) / win_bytes;
    if (st > 8) st = 8;
    if (st < 2) continue;
    double tiles = static_cast<double>(st) / (d.R / cand);
    if (tiles > 2.0) tiles = 2.0;
    if (tiles > best + 1e-9) {
      best = tiles;
      rg = cand;
      stages = st;
      p.win_rows = rows;
      p.n_chunks = n_chunks;
      p.chunk_rows = chunk;
      p.win_bytes = win_bytes;
    }
  }
  if (rg < 1 || stages < 2) return MPX_ERR_UNSUPPORTED;
  p.rg = rg;
  p.n_windows = d.R / rg;
  p.taps_per_win = rg * d.S;
  p.M_pad = static_cast<long long>(d.n_img) * p.Hp * p.Wp;
  p.q_base = static_cast<long long>(p.pl_h) * p.Wp + p.pl_w;
  const long long m_tiles = (p.M_pad - p.q_base + kBlockM - 1) / kBlockM;
  if (m_tiles < 2 || m_tiles >= (1LL << 30)) return MPX_ERR_UNSUPPORTED;
  if (p.M_pad + 2LL * kBlockM * sm_count() >= (1LL << 31)) return MPX_ERR_UNSUPPORTED;  // padded-linear row and pixel indices are 32-bit
  p.m_tiles = static_cast<int>(m_tiles);
  p.relu = d.relu;
  p.mma_issuers = 2;
  p.observers_arrive = 0;
  p.row_epilogue = (g_conv_mode & 1048576) ? 1 : 0;
  p.idesc = (1u << 4) | kIdescAB | (static_cast<unsigned>(kWinN >> 3) << 17) | (static_cast<unsigned>(256 >> 4) << 24);
  p.kskip = stem_kskip(d);
  p.bias = bias;
  p.residual = reinterpret_cast<const act_t*>(residual);
  p.out = reinterpret_cast<act_t*>(out);
  p.pool = d.pool ? 1 : 0;
  p.pool_H = (d.H - 1) / 2 + 1;
  p.pool_W = (d.W - 1) / 2 + 1;

  int rc = load_driver_entry_points();
  if (rc != MPX_OK) return rc;
  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.n_img)};
    cuuint64_t strides[3] = {128, static_cast<cuuint64_t>(d.W) * 128, static_cast<cuuint64_t>(d.H) * d.W * 128};
    int lower[2] = {-d.pad_lo_w, -d.pad_lo_h};
    int upper[2] = {d.pad_hi_w, d.pad_hi_h};  // the base pixel walks the whole padded image
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_im2col(&map_a, kTmaActType, 4, const_cast<void*>(x), dims, strides, lower,
                                 upper, kBlockK, static_cast<cuuint32_t>(p.chunk_rows), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MPX_ERR_UNSUPPORTED;
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const size_t bytes = static_cast<size_t>(d.n_img) * d.H * d.W * 128;
    if (drv <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&map_a)[1] &= ~(1ull << 21);
  }
  {
    const cuuint64_t K_total = static_cast<cuuint64_t>(n_taps) * 64;
    cuuint64_t dims[2] = {K_total, 64};
    cuuint64_t strides[1] = {K_total * 2};
    cuuint32_t box[2] = {kBlockK, kWinN / 2};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&map_b, kTmaActType, 2, const_cast<void*>(w), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  }
  const int smem_bytes = 1024 + b_bytes + stages * p.win_bytes + 8 * kStageTileBytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MPX_CHECK_CUDA(cudaFuncSetAttribute(conv_windowq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int n_ptiles = (p.m_tiles + 1) / 2;
  int cap = (max_ctas > 0 ? max_ctas : sm_count()) / 2;
  if (cap < 1) cap = 1;
  const int pairs = n_ptiles < cap ? n_ptiles : cap;
  ProfileSlot* slot = profile_begin(stream);
  MPX_CHECK_CUDA(launch_pdl(conv_windowq_kernel, dim3(2 * pairs), dim3(384), smem_bytes, stream, 2, map_a, map_b, p, stages,
                            n_taps));
  MPX_CHECK_CUDA(cudaGetLastError());
  ++g_launches;
  profile_end(slot, stream, 2.0 * d.n_img * d.H * d.W * 64.0 * n_taps * 64.0);
  return MPX_OK;
}

// ---------------------------------------------------------------------------------------------
// Mode bit 23 = 8388608: the 64 -> 64 pair window kernel with a SLIDING window.  conv_windowq_kernel reloads the whole
// window -- 128 output rows + (R-1) * Wp + S - 1 halo rows: 620 rows (79 KB) for the stem, 294 (38 KB) for layer1 -- for every
// tile because its tiles are dealt out round-robin; at the N = 64 tensor rate that is 37 B/clk/SM of L2 -> SM traffic for the
// stem plus the stores: beyond the ~43 B/clk/SM the L2 sustains chip-wide (B300_MICROARCH.md: LTS throughput cap), which --
// not HBM and not the tensor pipe -- is what bounded the stem and layer1.  Here every CTA owns a CONTIGUOUS run of T blocks
// of 128 padded-linear output rows; the activations live in a ring of 16 KB chunks (128 rows each) and block i reads chunks
// i .. i + ncw - 1, so that each block costs ONE new chunk (16 KB) instead of a window.  The ring is addressed linearly by
// the operand descriptors: an operand of 128 rows that starts in the last slot runs on into a shadow copy of slot 0 kept
// behind it (chunks destined for slot 0 are loaded twice: + 1/slots traffic).  The two MMA issuers still alternate blocks; a
// chunk is last read by blocks i and i - 1 (one per issuer), so its empty barrier takes two arrivals -- block i's commit
// arrives on the barriers of chunk i and chunk i + 1.  Everything else (half weight tiles per CTA, accumulator ring, staged /
// pooled epilogue) is conv_windowq_kernel's.  Results are bit-identical to it (same MMAs in the same order per output row).
// ---------------------------------------------------------------------------------------------
constexpr int kWsChunkBytes = kBlockM * 128;  // 128 rows x 64 channels x 2 B = 16 KB
constexpr int kWsChunkUnits = kWsChunkBytes / 16;

__global__ void __launch_bounds__(384, 1)
conv_windows_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_r, const WinParams p, int slots, int ncw, int T, int n_taps) {
  constexpr int kEpiSets = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_b = smem;                                   // n_taps resident half weight tiles
  uint8_t* smem_ring = smem + n_taps * kWinqBHalf;          // `slots` chunks + the shadow of slot 0
  // staging: 8 epilogue warps x 4 KB -- or, with a residual, four 16 KB blocks that TMA fills with the residual rows of blocks
  // i, i+1, i+2, i+3 (same 128B-swizzled layout as the staging tiles) and the epilogue then overwrites in place
  uint8_t* smem_stg = smem_ring + static_cast<size_t>(slots + 1) * kWsChunkBytes;
  const bool has_res = p.residual != nullptr;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + (has_res ? kWinAccBufs * kWsChunkBytes : 8 * kStageTileBytes));
  uint64_t* full_bar = bars;             // [slots <= 8] leader
  uint64_t* empty_bar = bars + 8;        // [slots]      per CTA, two arrivals
  uint64_t* tmem_full = bars + 16;       // [4]          per CTA
  uint64_t* tmem_empty = bars + 20;      // [4]          leader, eight arrivals
  uint64_t* b_full = bars + 24;          // [1]          leader
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 25);
  uint64_t* res_full = bars + 26;        // [4]          per CTA: residual rows of block i landed in buffer i % 4
  float* bias_s = reinterpret_cast<float*>(bars + 32);  // [64]
  uint64_t* res_empty = bars + 64;       // [4]          per CTA, the four warps of the set that owned the buffer

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  if (threadIdx.x < kWinN) bias_s[threadIdx.x] = p.bias[threadIdx.x];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < slots; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 2);
    }
    for (int i = 0; i < kWinAccBufs; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
      mbar_init(&res_full[i], 1);
      mbar_init(&res_empty[i], 4);
    }
    mbar_init(b_full, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(kWinAccBufs * kWinN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int hpwp = p.Hp * p.Wp;
  const long long run_start = (2LL * pair + rank) * T;  // first block of this CTA's run
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own chunks, own half of the weights) =====================
    if (leader) mbar_expect_tx_u(b_full, 2u * static_cast<uint32_t>(n_taps) * kWinqBHalf);
    else mbar_arrive_remote_u(b_full, 0);
    for (int t = 0; t < n_taps; ++t)
      tma2_load_2d_u(smem_b + t * kWinqBHalf, &map_b, b_full, t * kBlockK, static_cast<int>(rank) * (kWinN / 2));
    mbar_arrive_u(&empty_bar[0]);  // stands in for "the block before block 0" (see the issuer's double commit)
    const long long l_first = run_start * kBlockM;  // padded-linear input row of chunk 0
    int img = static_cast<int>(l_first / hpwp);
    int yp = static_cast<int>((l_first - static_cast<long long>(img) * hpwp) / p.Wp);
    int xp = static_cast<int>(l_first - static_cast<long long>(img) * hpwp - static_cast<long long>(yp) * p.Wp);
    int slot = 0;
    uint32_t phase = 0;
    const int n_chunks = T + ncw - 1;
    for (int j = 0; j < n_chunks; ++j) {
      mbar_wait(&empty_bar[slot], phase ^ 1u);
      const uint32_t bytes = slot == 0 ? 2u * kWsChunkBytes : static_cast<uint32_t>(kWsChunkBytes);
      if (leader) mbar_expect_tx_u(&full_bar[slot], 2u * bytes);
      else mbar_arrive_remote_u(&full_bar[slot], 0);
      tma2_load_im2col_4d_u(smem_ring + static_cast<size_t>(slot) * kWsChunkBytes, &map_a, &full_bar[slot], 0, xp - p.pl_w,
                            yp - p.pl_h, img, 0, 0);
      if (slot == 0)
        tma2_load_im2col_4d_u(smem_ring + static_cast<size_t>(slots) * kWsChunkBytes, &map_a, &full_bar[slot], 0, xp - p.pl_w,
                              yp - p.pl_h, img, 0, 0);
      xp += kBlockM;
      while (xp >= p.Wp) {
        xp -= p.Wp;
        if (++yp == p.Hp) {
          yp = 0;
          ++img;
        }
      }
      if (++slot == slots) {
        slot = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers (leader only), blocks alternately =====================
    const int which = warp == 1 ? 0 : 1;
    if (leader) {
      const uint32_t idesc = p.idesc;
      mbar_wait(b_full, 0);
      const uint32_t ring_units = static_cast<uint32_t>(slots) * kWsChunkUnits;
      const uint32_t row_units = static_cast<uint32_t>(p.Wp) * 8u;
      const uint64_t ring_desc = make_sw128_desc(smem_u32(smem_ring));
      const uint64_t b_base = make_sw128_desc(smem_u32(smem_b));
      const int R = n_taps / p.S;
      const unsigned ks_row0 = static_cast<unsigned>(p.kskip) & 15u;          // dead slices of the taps (0, s > 0) ...
      const unsigned ks_col0 = static_cast<unsigned>(p.kskip >> 16) & 15u;    // ... and of the taps (r > 0, 0): see conv_windows_try
      for (int i = which; i < T; i += 2) {
        const int acc = i % kWinAccBufs;
        const uint32_t acc_phase = (i / kWinAccBufs) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        for (int k = 0; k < ncw; ++k) {  // the chunks of this block (all but the newest were awaited for earlier blocks)
          const int j = i + k;
          mbar_wait(&full_bar[j % slots], static_cast<uint32_t>(j / slots) & 1u);
        }
        tc_fence_after();
        const uint32_t tmem_d = static_cast<uint32_t>(acc * kWinN);
        uint32_t first = 1;
        uint64_t db = b_base;
        uint32_t pr = static_cast<uint32_t>(i % slots) * kWsChunkUnits;  // ring position (16-byte units) of filter row r
        for (int r = 0; r < R; ++r) {
          uint32_t pt = pr;
          for (int s = 0; s < p.S; ++s) {
            const uint64_t da = desc_add_lo(ring_desc, pt);
            // structurally zero K slices of the space-to-depth stem: the (dy, dx) sub-pixels that fall into the filter's
            // padding on its first row / first column -- a function of (r == 0, s == 0) only, kept in uniform registers
            // (the per-tap 64-bit shift of conv_windowq_kernel ran on the vector pipe: the two issuers were busy all the time
            // at ~12 instructions x ~10 cycles per MMA, and the stem -- 44% of whose taps take the slow path -- was issue-bound)
            const unsigned sk = (r == 0 ? ks_row0 : 0u) | (s == 0 ? ks_col0 : 0u);
            if (sk == 0u) {
              tc2_mma_f16_u(tmem_d, da, db, idesc, first ? 0u : 1u);
              tc2_mma_f16_u(tmem_d, desc_add_lo(da, 2u), desc_add_lo(db, 2u), idesc, 1u);
              tc2_mma_f16_u(tmem_d, desc_add_lo(da, 4u), desc_add_lo(db, 4u), idesc, 1u);
              tc2_mma_f16_u(tmem_d, desc_add_lo(da, 6u), desc_add_lo(db, 6u), idesc, 1u);
              first = 0;
            } else {  // structurally zero weight slices of the space-to-depth stem: not issued
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (((sk >> ks) & 1u) == 0u) {
                  tc2_mma_f16_u(tmem_d, desc_add_lo(da, 2u * ks), desc_add_lo(db, 2u * ks), idesc, first ? 0u : 1u);
                  first = 0;
                }
            }
            db = desc_add_lo(db, kWinqBHalf / 16);
            pt += 8u;  // next tap: one pixel (128 B) further
            if (pt >= ring_units) pt -= ring_units;
          }
          pr += row_units;  // next filter row: Wp pixels further
          if (pr >= ring_units) pr -= ring_units;
        }
        tc2_commit_mc_u(&empty_bar[i % slots]);        // chunk i: last read by this block and by block i - 1
        tc2_commit_mc_u(&empty_bar[(i + 1) % slots]);  // chunk i + 1: last read by block i + 1 and by this one
        tc2_commit_mc_u(&tmem_full[acc]);
      }
    }
  } else if (warp == 2) {
    // ===================== residual producer (both CTAs): TMA, four blocks ahead of the epilogue =====================
    // The residual pixel of output row q sits at padded-linear position q of the residual tensor seen through the same
    // image + padding-ring box as the activations: 128 consecutive positions = the block's rows, ring positions zero-filled.
    // (Loaded by the epilogue warps with LDG + STS, as conv_windowq_kernel does, the global latency sat in the epilogue's
    // dependent chain: ncu tensor-pipe activity 47% with a residual against 71% without.)
    if (has_res) {
      const long long q_first = p.q_base + run_start * kBlockM;
      int img = static_cast<int>(q_first / hpwp);
      int yp = static_cast<int>((q_first - static_cast<long long>(img) * hpwp) / p.Wp);
      int xp = static_cast<int>(q_first - static_cast<long long>(img) * hpwp - static_cast<long long>(yp) * p.Wp);
      for (int i = 0; i < T; ++i) {
        const int b = i % kWinAccBufs;
        mbar_wait(&res_empty[b], ((static_cast<uint32_t>(i) / kWinAccBufs) & 1u) ^ 1u);
        mbar_expect_tx_u(&res_full[b], static_cast<uint32_t>(kWsChunkBytes));
        tma_load_im2col_4d_u(smem_stg + static_cast<size_t>(b) * kWsChunkBytes, &map_r, &res_full[b], 0, xp - p.pl_w, yp - p.pl_h,
                             img, 0, 0);
        xp += kBlockM;
        while (xp >= p.Wp) {
          xp -= p.Wp;
          if (++yp == p.Hp) {
            yp = 0;
            ++img;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows), two warp sets, blocks round-robin =====================
    if (tmem_base != 0u) __trap();
    const int q4 = warp & 3;
    const int set = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;
    const unsigned hpwp_u = static_cast<unsigned>(hpwp), Wp_u = static_cast<unsigned>(p.Wp);
    int pool_info = 0, pool_base = 0;
    auto pix_of = [&](int i) -> int {
      pool_info = 0;
      const long long q = p.q_base + (run_start + i) * kBlockM + row;
      if (q >= p.M_pad) return -1;
      const unsigned qu = static_cast<unsigned>(q);
      const unsigned img = qu / hpwp_u;
      const unsigned rem = qu - img * hpwp_u;
      const unsigned yp = rem / Wp_u, xp = rem - yp * Wp_u;
      const int y = static_cast<int>(yp) - p.pl_h, x = static_cast<int>(xp) - p.pl_w;
      if (y < 0 || y >= p.H || x < 0 || x >= p.W) return -1;
      if (p.pool) {  // see conv_windowq_kernel
        int j = x >> 1;
        if ((x & 1) == 0) {
          pool_info = 1 | ((x > 0 && lane > 0) ? 4 : 0) | ((x + 1 < p.W && lane < 31) ? 8 : 0);
        } else if (lane == 0) {
          pool_info = 2;
        } else if (lane == 31 && x + 1 < p.W) {
          pool_info = 2;
          j += 1;
        }
        if ((y & 1) && (y >> 1) + 1 < p.pool_H) pool_info |= 16;
        pool_base = (static_cast<int>(img) * p.pool_H + (y >> 1)) * p.pool_W + j;
      }
      return (static_cast<int>(img) * p.H + y) * p.W + x;
    };
    for (int i = set; i < T; i += kEpiSets) {
      const int acc = i % kWinAccBufs;
      const uint32_t acc_phase = (i / kWinAccBufs) & 1;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(acc * kWinN);
      const uint32_t stg = smem_u32(smem_stg) + (has_res ? static_cast<uint32_t>(acc * kWsChunkBytes + q4 * kStageTileBytes)
                                                         : static_cast<uint32_t>((warp - 4) * kStageTileBytes));
      const int pix = pix_of(i);
      if (has_res) mbar_wait(&res_full[acc], acc_phase);  // written through the async proxy, complete_tx on the barrier
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_compute64(taddr, stg, lane, has_res, smem_u32(bias_s), p.relu);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_remote(&tmem_empty[acc], 0);
      }
      if (p.pool) pool_staged64(stg, smem_u32(bars + 72) + static_cast<uint32_t>((warp - 4) * 32), lane, pool_info, pool_base, p.pool_W * kWinN, p.out);
      else store_staged64(stg, lane, pix, p.out);
      if (has_res) {  // the buffer goes back to the residual producer: generic-proxy writes before the next TMA write
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&res_empty[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kWinAccBufs * kWinN)
                 : "memory");
  }
}

// Returns MPX_ERR_UNSUPPORTED (without setting an error) when the shape does not fit or the mode bit is off.
static int conv_windows_try(const ConvDesc& d, const void* x, const void* w, const float* bias, const void* residual,
                            void* out, int max_ctas, cudaStream_t stream) {
  if ((g_conv_mode & 8388608) == 0 || (g_conv_mode & 32768) == 0 || (g_conv_mode & 1048576) != 0) return MPX_ERR_UNSUPPORTED;
  if (d.stride != 1 || d.C_in != 64 || d.C_out != 64) return MPX_ERR_UNSUPPORTED;
  if (d.R > 4 || d.S > 4 || d.R * d.S > 16) return MPX_ERR_UNSUPPORTED;
  if (d.pool && (residual != nullptr || !d.relu)) return MPX_ERR_UNSUPPORTED;
  const int P = conv_out_dim(d.H, d.pad_lo_h, d.pad_hi_h, d.R, 1), Q = conv_out_dim(d.W, d.pad_lo_w, d.pad_hi_w, d.S, 1);
  if (P != d.H || Q != d.W) return MPX_ERR_UNSUPPORTED;
  WinParams p{};
  p.Hp = d.H + d.pad_lo_h + d.pad_hi_h;
  p.Wp = d.W + d.pad_lo_w + d.pad_hi_w;
  p.H = d.H;
  p.W = d.W;
  p.pl_h = d.pad_lo_h;
  p.pl_w = d.pad_lo_w;
  p.n_img = d.n_img;
  p.S = d.S;
  const int n_taps = d.R * d.S;
  const int b_bytes = n_taps * kWinqBHalf;
  const int halo = (d.R - 1) * p.Wp + (d.S - 1);
  const int ncw = (kBlockM + halo + kBlockM - 1) / kBlockM;
  if (ncw < 2) return MPX_ERR_UNSUPPORTED;  // no halo, nothing to slide; also: block j + slots - 1 must need chunk j + slots (the
                                            // release of slot j's previous occupant orders the arrivals of consecutive phases)
  const int smem_limit = 227 * 1024 - 1024 /*align*/ - 1024 /*barriers, bias*/;
  const int stg_bytes = residual != nullptr ? kWinAccBufs * kWsChunkBytes : 8 * kStageTileBytes;
  int slots = (smem_limit - b_bytes - stg_bytes) / kWsChunkBytes - 1;
  if (slots > 8) slots = 8;
  if (slots < ncw + 1) return MPX_ERR_UNSUPPORTED;
  if (p.Wp * 8 >= slots * kWsChunkUnits) return MPX_ERR_UNSUPPORTED;
  p.M_pad = static_cast<long long>(d.n_img) * p.Hp * p.Wp;
  p.q_base = static_cast<long long>(p.pl_h) * p.Wp + p.pl_w;
  const long long blocks = (p.M_pad - p.q_base + kBlockM - 1) / kBlockM;
  int cap = (max_ctas > 0 ? max_ctas : sm_count()) / 2;
  if (cap < 1) cap = 1;
  const long long T_ll = (blocks + 2LL * cap - 1) / (2LL * cap);
  if (T_ll < 8) return MPX_ERR_UNSUPPORTED;  // short runs reload most of their window anyway: conv_windowq_kernel
  const int T = static_cast<int>(T_ll);
  const int pairs = static_cast<int>((blocks + 2LL * T - 1) / (2LL * T));
  if (p.M_pad + (2LL * pairs * T + ncw + 2) * kBlockM >= (1LL << 31)) return MPX_ERR_UNSUPPORTED;
  p.m_tiles = static_cast<int>(blocks);
  p.relu = d.relu;
  p.idesc = (1u << 4) | kIdescAB | (static_cast<unsigned>(kWinN >> 3) << 17) | (static_cast<unsigned>(256 >> 4) << 24);
  {
    // the kernel takes the zero slices in structured form: dead (dy, dx) slices of the first filter row (tap (0, 1)) and of
    // the first filter column (tap (1, 0)); any other pattern goes to conv_windowq_kernel
    const unsigned long long full = stem_kskip(d);
    unsigned long long packed = 0;
    if (full != 0) {
      const unsigned row0 = static_cast<unsigned>(full >> (4 * 1)) & 15u, col0 = static_cast<unsigned>(full >> (4 * d.S)) & 15u;
      unsigned long long expect = 0;
      for (int r = 0; r < d.R; ++r)
        for (int sx = 0; sx < d.S; ++sx)
          expect |= static_cast<unsigned long long>((r == 0 ? row0 : 0u) | (sx == 0 ? col0 : 0u)) << (4 * (r * d.S + sx));
      if (expect != full) return MPX_ERR_UNSUPPORTED;
      packed = row0 | (static_cast<unsigned long long>(col0) << 16);
    }
    p.kskip = packed;
  }
  p.bias = bias;
  p.residual = reinterpret_cast<const act_t*>(residual);
  p.out = reinterpret_cast<act_t*>(out);
  p.pool = d.pool ? 1 : 0;
  p.pool_H = (d.H - 1) / 2 + 1;
  p.pool_W = (d.W - 1) / 2 + 1;

  int rc = load_driver_entry_points();
  if (rc != MPX_OK) return rc;
  CUtensorMap map_a, map_b, map_r;
  for (int which = 0; which < 2; ++which) {  // activations; residual (same geometry: both are [n, H, W, 64])
    CUtensorMap& m = which == 0 ? map_a : map_r;
    const void* base = which == 0 ? x : (residual != nullptr ? residual : x);
    cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.n_img)};
    cuuint64_t strides[3] = {128, static_cast<cuuint64_t>(d.W) * 128, static_cast<cuuint64_t>(d.H) * d.W * 128};
    int lower[2] = {-d.pad_lo_w, -d.pad_lo_h};
    int upper[2] = {d.pad_hi_w, d.pad_hi_h};  // the base pixel walks the whole padded image
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_im2col(&m, kTmaActType, 4, const_cast<void*>(base), dims, strides, lower, upper, kBlockK,
                                 static_cast<cuuint32_t>(kBlockM), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return MPX_ERR_UNSUPPORTED;
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const size_t bytes = static_cast<size_t>(d.n_img) * d.H * d.W * 128;
    if (drv <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
  }
  {
    const cuuint64_t K_total = static_cast<cuuint64_t>(n_taps) * 64;
    cuuint64_t dims[2] = {K_total, 64};
    cuuint64_t strides[1] = {K_total * 2};
    cuuint32_t box[2] = {kBlockK, kWinN / 2};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&map_b, kTmaActType, 2, const_cast<void*>(w), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  }
  const int smem_bytes = 1024 + b_bytes + (slots + 1) * kWsChunkBytes + stg_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MPX_CHECK_CUDA(cudaFuncSetAttribute(conv_windows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  ProfileSlot* slot = profile_begin(stream);
  MPX_CHECK_CUDA(launch_pdl(conv_windows_kernel, dim3(2 * pairs), dim3(384), smem_bytes, stream, 2, map_a, map_b, map_r, p, slots,
                            ncw, T, n_taps));
  MPX_CHECK_CUDA(cudaGetLastError());
  ++g_launches;
  profile_end(slot, stream, 2.0 * d.n_img * d.H * d.W * 64.0 * n_taps * 64.0);
  return MPX_OK;
}

// ---------------------------------------------------------------------------------------------
// Bring-up probe: D[128,64] = A[r0 : r0+128, 0:64] * B[64,64]^T with the A descriptor started r0 rows
// into a 128B-swizzled tile that TMA wrote at a 1024-byte aligned address.  Answers whether a
// row-shifted start address (plus optional base_offset) addresses the swizzled rows correctly, the
// precondition for reusing one shared-memory halo tile across the filter taps.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_rowshift_probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                           int r0, int base_off, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;                 // 144 rows x 128 B
  uint8_t* smem_b = smem + 144 * 128 + 1024 - (144 * 128) % 1024;  // next 1024-aligned address
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + 64 * 128);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)),
                 "r"(64)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 144 * 128 + 64 * 128);
    tma_load_2d(smem_a, &map_a, &bars[0], 0, 0);
    tma_load_2d(smem_b, &map_b, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc = (1u << 4) | kIdescAB | (static_cast<uint32_t>(64 >> 3) << 17) |
                               (static_cast<uint32_t>(128 >> 4) << 24);
    const uint64_t da = make_sw128_desc(smem_u32(smem_a) + static_cast<uint32_t>(r0) * 128u) |
                        (static_cast<uint64_t>(base_off & 7) << 49);
    const uint64_t db = make_sw128_desc(smem_u32(smem_b));
#pragma unroll
    for (int k = 0; k < 4; ++k)
      tc_mma_f16(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                  k > 0 ? 1u : 0u);
    tc_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
#pragma unroll 1
  for (int c = 0; c < 64; c += 32) {
    uint32_t v[32];
    tc_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c), v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[row * 64 + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
  }
}

// a: [144, 64] bf16 row-major, b: [64, 64] bf16 row-major (K contiguous), out: [128, 64] fp32
int umma_rowshift_probe(const void* a, const void* b, int r0, int base_off, float* out, cudaStream_t stream) {
  MPX_REQUIRE(r0 >= 0 && r0 <= 16, "probe: r0 out of range");
  int rc = load_driver_entry_points();
  if (rc != MPX_OK) return rc;
  CUtensorMap map_a, map_b;
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {64, 144};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 144};
    CUresult r = g_encode_tiled(&map_a, kTmaActType, 2, const_cast<void*>(a), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPX_REQUIRE(r == CUDA_SUCCESS, "probe: encode A failed (%d)", static_cast<int>(r));
  }
  {
    cuuint64_t dims[2] = {64, 64};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 64};
    CUresult r = g_encode_tiled(&map_b, kTmaActType, 2, const_cast<void*>(b), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MPX_REQUIRE(r == CUDA_SUCCESS, "probe: encode B failed (%d)", static_cast<int>(r));
  }
  const int smem_bytes = 1024 + 144 * 128 + 1024 + 64 * 128 + 64;
  MPX_CHECK_CUDA(cudaFuncSetAttribute(umma_rowshift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      smem_bytes));
  umma_rowshift_probe_kernel<<<1, 128, smem_bytes, stream>>>(map_a, map_b, r0, base_off, out);
  MPX_CHECK_CUDA(cudaGetLastError());
  ++g_launches;
  return MPX_OK;
}

// ---------------------------------------------------------------------------------------------
// Tensor-pipe probe (tools/gpu_mma_probe.py): how many cycles one tcgen05.mma.kind::f16 of shape (128 x cta_group) x N x 16
// costs with both operands in shared memory, as a function of N, of the number of independent accumulate chains one
// issuing thread interleaves, and of the number of issuing threads.  Every SM (or SM pair) runs the same loop on its own
// zero-filled operands; the result is the mean over CTAs of elapsed cycles / MMAs issued.  These are the floors the
// convolution kernels are measured against in DESIGN.md.
// ---------------------------------------------------------------------------------------------
template <int CG>
__global__ void __launch_bounds__(192, 1)
mma_probe_kernel(int N, int chains, int issuers, int n_mma, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;               // 128 rows x 128 B
  uint8_t* smem_b = smem + 16384;       // N / CG rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + 32768);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zero fill -> async-proxy (tensor core) reads
  if (warp == 4) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const bool leader = CG == 1 || cluster_ctarank() == 0;
  const bool uniform = (issuers & 0x100) != 0;  // warp-uniform issue path (elect.sync inside the instruction), cta_group 1
  issuers &= 0xff;
  const int which = __shfl_sync(0xffffffffu, warp, 0);  // issuing warps 0 .. issuers-1 (one per SM sub-partition)
  if (CG == 1 && uniform) {
    if (which < issuers) {
      const uint32_t idesc = (1u << 4) | kIdescAB | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
      const uint64_t da = make_sw128_desc(smem_u32(smem_a));
      const uint64_t db = make_sw128_desc(smem_u32(smem_b));
      const uint32_t d0 = __shfl_sync(0xffffffffu, tmem_base, 0) + static_cast<uint32_t>(which * chains * N);
      const uint32_t cmask = static_cast<uint32_t>(chains - 1);
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 4) {
        const uint32_t d = d0 + ((static_cast<uint32_t>(i) >> 2) & cmask) * static_cast<uint32_t>(N);
        tc_mma_f16_u(d, da, db, idesc, 1u);
        tc_mma_f16_u(d, da + 2, db + 2, idesc, 1u);
        tc_mma_f16_u(d, da + 4, db + 4, idesc, 1u);
        tc_mma_f16_u(d, da + 6, db + 6, idesc, 1u);
      }
      const long long t_issue = clock64();
      tc_commit_u(&bars[which]);
      mbar_wait(&bars[which], 0);
      const long long t1 = clock64();
      if (lane == 0) {
        cycles[blockIdx.x * 8 + which] = t1 - t0;
        cycles[blockIdx.x * 8 + 4 + which] = t_issue - t0;
      }
    }
  } else if (leader && lane == 0 && which < issuers) {
    const uint32_t idesc = (1u << 4) | kIdescAB | (static_cast<uint32_t>(N >> 3) << 17) |
                           (static_cast<uint32_t>((128 * CG) >> 4) << 24);
    const uint64_t da = make_sw128_desc(smem_u32(smem_a));
    const uint64_t db = make_sw128_desc(smem_u32(smem_b));
    const uint32_t d0 = tmem_base + static_cast<uint32_t>(which * chains * N);
    const uint32_t cmask = static_cast<uint32_t>(chains - 1);  // chains is a power of two
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      // four K steps of one 64-wide k-block, as the convolution kernels issue them
      const uint32_t d = d0 + ((static_cast<uint32_t>(i) >> 2) & cmask) * static_cast<uint32_t>(N);
      if (CG == 1) {
        tc_mma_f16(d, da, db, idesc, 1u);
        tc_mma_f16(d, da + 2, db + 2, idesc, 1u);
        tc_mma_f16(d, da + 4, db + 4, idesc, 1u);
        tc_mma_f16(d, da + 6, db + 6, idesc, 1u);
      } else {
        tc2_mma_f16(d, da, db, idesc, 1u);
        tc2_mma_f16(d, da + 2, db + 2, idesc, 1u);
        tc2_mma_f16(d, da + 4, db + 4, idesc, 1u);
        tc2_mma_f16(d, da + 6, db + 6, idesc, 1u);
      }
    }
    const long long t_issue = clock64();
    if (CG == 1) tc_commit(&bars[which]);
    else tc2_commit_mc(&bars[which]);
    mbar_wait(&bars[which], 0);
    const long long t1 = clock64();
    cycles[(blockIdx.x / CG) * 8 + which] = t1 - t0;
    cycles[(blockIdx.x / CG) * 8 + 4 + which] = t_issue - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 4) {
    tc_fence_after();
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int mma_probe(int cta_group, int n, int chains, int issuers, int n_mma, double* cycles_per_mma, double* issue_cycles_per_mma) {
  MPX_REQUIRE(cta_group == 1 || cta_group == 2, "mma_probe: cta_group must be 1 or 2");
  MPX_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0, "mma_probe: N=%d", n);
  const int n_issuers = issuers & 0xff;
  MPX_REQUIRE(n_issuers >= 1 && n_issuers <= 4 && (chains == 1 || chains == 2 || chains == 4) && n_issuers * chains * n <= 512,
              "mma_probe: issuers in 1..4, chains in {1, 2, 4}, accumulators must fit the 512 TMEM columns");
  MPX_REQUIRE(n_mma >= 4 && n_mma % 4 == 0 && n_mma <= (1 << 22), "mma_probe: n_mma=%d", n_mma);
  const int ctas = (sm_count() / cta_group) * cta_group;
  long long* d_cycles = nullptr;
  MPX_CHECK_CUDA(cudaMalloc(&d_cycles, sizeof(long long) * 8 * ctas));
  MPX_CHECK_CUDA(cudaMemset(d_cycles, 0, sizeof(long long) * 8 * ctas));
  const int smem_bytes = 1024 + 16384 + 32768 + 64;
  if (cta_group == 1) {
    MPX_CHECK_CUDA(cudaFuncSetAttribute(mma_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    mma_probe_kernel<1><<<ctas, 192, smem_bytes>>>(n, chains, issuers, n_mma, d_cycles);
  } else {
    MPX_CHECK_CUDA(cudaFuncSetAttribute(mma_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MPX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mma_probe_kernel<2>, n, chains, issuers, n_mma, d_cycles));
  }
  MPX_CHECK_CUDA(cudaGetLastError());
  MPX_CHECK_CUDA(cudaDeviceSynchronize());
  std::vector<long long> h(8 * ctas);
  MPX_CHECK_CUDA(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * 8 * ctas, cudaMemcpyDeviceToHost));
  cudaFree(d_cycles);
  double sum = 0, sum_issue = 0;
  int cnt = 0;
  for (int g = 0; g < ctas / cta_group; ++g) {
    long long worst = 0, worst_issue = 0;
    for (int w = 0; w < n_issuers; ++w) {
      worst = h[8 * g + w] > worst ? h[8 * g + w] : worst;
      worst_issue = h[8 * g + 4 + w] > worst_issue ? h[8 * g + 4 + w] : worst_issue;
    }
    sum += static_cast<double>(worst) / (static_cast<double>(n_mma) * n_issuers);
    sum_issue += static_cast<double>(worst_issue) / static_cast<double>(n_mma);
    ++cnt;
  }
  *cycles_per_mma = cnt ? sum / cnt : 0.0;                 // SM time per MMA with all issuers running
  if (issue_cycles_per_mma) *issue_cycles_per_mma = cnt ? sum_issue / cnt : 0.0;  // one thread's time per instruction
  ++g_launches;
  return MPX_OK;
}

}  // namespace mpx
