// Tensor-core arithmetic of the AO-v0 step path (AOG_PRECISION_TENSOR), sm_100a only.
//
// The matrix Fourier transform F = M1 . E . M2 (reference AO_env.py:138 -> hcipy
// FraunhoferPropagator / MatrixFourierTransform) runs as two real-embedded complex GEMMs on the
// 5th-gen tensor cores (tcgen05.mma kind::f16, FP32 accumulators in TMEM), operands staged in
// shared memory by TMA (cp.async.bulk.tensor, SWIZZLE_64B) through a 3-stage mbarrier ring:
//
//   stage 1  [Tr ; Ti] (256 x 240) = [[M1r, -M1i], [M1i, M1r]] (256 x 480) . [Er ; Ei] (480 x 240)
//   stage 2  [Fr | Fi] (128 x 256) = [Tr | Ti] (128 x 480) . [[M2r, M2i], [-M2i, M2r]] (480 x 256)
//
// Precision: every operand is split x = hi + lo into two fp16 values (|x| <= 240, so the pair
// carries ~22 bits) and each product is issued as hi.hi + hi.lo + lo.hi ("3x split"), which
// keeps FP32-class accuracy at fp16 tensor throughput.  The unit-modulus twiddle tables are
// split once on the host from their FP64 values; the pupil field is split by the field kernel,
// the stage-1 product by the stage-1 epilogue.  The stage-2 epilogue never writes the focal
// plane: it projects it on the fibre modes (AO_env.py:471) straight out of TMEM.
#include "tensor_path.cuh"
#include "kernels_f64.cuh"

#include <cuda.h>
#include <algorithm>
#include <cmath>
#include <vector>

namespace {

constexpr int TC_NP = 240;          // pupil pixels per side this path is built for
constexpr int TC_NF = 128;          // focal pixels per side
constexpr int TC_K = 2 * TC_NP;     // real-embedded contraction length (480)
constexpr int KB = 32;              // K elements per pipeline stage (64 B rows, SWIZZLE_64B)
constexpr int NUM_KB = TC_K / KB;   // 15
constexpr int NUM_STAGES = 3;
constexpr int A_TILE = 128 * KB * 2;            // 8 KB: 128 rows x 64 B
constexpr int B_TILE = 256 * KB * 2;            // 16 KB slot (stage 1 uses 240 rows of it)
constexpr int STAGE_BYTES = 4 * A_TILE + 2 * B_TILE;   // 64 KB
constexpr int SMEM_BYTES = NUM_STAGES * STAGE_BYTES + 1024 /*align*/ + 4096 /*barriers + reduction scratch*/;
constexpr int TC_THREADS = 192;     // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue

struct TensorState {
  // operands (device)
  __half* A1_hi = nullptr; __half* A1_lo = nullptr;   // [256][480]  stage-1 constant
  __half* B2_hi = nullptr; __half* B2_lo = nullptr;   // [256][480]  stage-2 constant (N-major rows, K contiguous)
  __half* E_hi = nullptr; __half* E_lo = nullptr;     // [chunk][240][480]  pupil field, x-major
  __half* T_hi = nullptr; __half* T_lo = nullptr;     // [chunk][128][480]  stage-1 product
  float* lpw = nullptr;                               // [J][128][128] fibre modes * weight / max
  float* screensT = nullptr;                          // [B][Np][Np] FP32, x-major (transposed), ring-buffered
  float* modesT = nullptr;                            // [K][Np][Np] FP32 transposed
  float* apT = nullptr;                               // [Np][Np]
  double2* m2oT = nullptr;                            // [n][Np] transposed obs table
  int* err_flag = nullptr;                            // device flag set by a timed-out barrier wait
  CUtensorMap tmA1_hi, tmA1_lo, tmE_hi, tmE_lo, tmT_hi, tmT_lo, tmB2_hi, tmB2_lo;
  double pupil_weight = 0.0;                          // |M1| (grid weight folded in the table)
  double lpw_scale = 0.0;
  int num_sms = 148;
  bool have_m1 = false, have_m2 = false, have_lp = false, have_modes = false, have_ap = false, have_m2o = false;
};

TensorState* TS(aog_env* env) { return reinterpret_cast<TensorState*>(env->tensor_state); }

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// Bounded wait: a protocol bug must trap (the launch fails, the GPU stays usable), never hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {   // ~2 s
      atomicExch(err_flag, code);
      __threadfence_system();
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 64 B,
// 8-row groups 512 B apart (SBO), LBO = 1 (ignored for swizzled K-major), version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;   // LayoutType::SWIZZLE_64B
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, K-major both.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ----------------------------------------------------------------------------- the GEMM kernel
struct TcParams {
  int num_items;        // MODE 0: envs;  MODE 1: env pairs
  int num_envs;         // envs in this chunk
  // MODE 0 epilogue: stage-1 product as split fp16, [env][128][480]
  __half* T_hi; __half* T_lo;
  // MODE 1 epilogue: fibre projection
  const float* lpw;     // [J][128][128]
  double2* coef;        // [env][J]
  int J;
  double2 scale;        // norm * amp * pupil_weight * lpw_scale
  int* err_flag;
};

// MODE 0: stage 1 (A = constant twiddles, 2 row-halves = real / imaginary rows; B = one env's field)
// MODE 1: stage 2 (A = stage-1 product of 2 envs, one per row-half; B = constant twiddles)
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_mft_tc(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
         const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const TcParams p) {
  constexpr int N_TILE = (MODE == 0) ? TC_NP : 2 * TC_NF;            // 240 | 256
  constexpr uint32_t TX_BYTES = 4 * A_TILE + 2 * N_TILE * KB * 2;    // bytes landing per stage
  constexpr uint32_t IDESC = umma_idesc_f16(128, N_TILE);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(base + NUM_STAGES * STAGE_BYTES);
  uint64_t* empty = full + NUM_STAGES;
  uint64_t* tmem_full = empty + NUM_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  double* red = reinterpret_cast<double*>(base + NUM_STAGES * STAGE_BYTES + 128);   // [2][4][2*AOG_MAX_LP*2] max 2*4*32

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB_lo)) : "memory");
    for (int s = 0; s < NUM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: all 512 columns (two 128-lane x 256-column FP32 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int a_row0 = (MODE == 0) ? 0 : item * 256;
        const int b_row0 = (MODE == 0) ? item * TC_NP : 0;
        for (int kb = 0; kb < NUM_KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, p.err_flag, 1);
          mbar_expect_tx(&full[stage], TX_BYTES);
          const uint32_t s0 = smem_u32(base + stage * STAGE_BYTES);
          const int k0 = kb * KB;
          tma_load_2d(s0 + 0 * A_TILE, &tmA_hi, &full[stage], k0, a_row0);
          tma_load_2d(s0 + 1 * A_TILE, &tmA_hi, &full[stage], k0, a_row0 + 128);
          tma_load_2d(s0 + 2 * A_TILE, &tmA_lo, &full[stage], k0, a_row0);
          tma_load_2d(s0 + 3 * A_TILE, &tmA_lo, &full[stage], k0, a_row0 + 128);
          tma_load_2d(s0 + 4 * A_TILE, &tmB_hi, &full[stage], k0, b_row0);
          tma_load_2d(s0 + 4 * A_TILE + B_TILE, &tmB_lo, &full[stage], k0, b_row0);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        mbar_wait(tmem_empty, tphase ^ 1, p.err_flag, 2);   // epilogue has drained the accumulators
        tc_fence_after();
        for (int kb = 0; kb < NUM_KB; ++kb) {
          mbar_wait(&full[stage], phase, p.err_flag, 3);
          tc_fence_after();
          const uint32_t s0 = smem_u32(base + stage * STAGE_BYTES);
#pragma unroll
          for (int ks = 0; ks < KB / 16; ++ks) {
            const uint64_t b_hi = umma_desc_sw64(s0 + 4 * A_TILE + ks * 32);
            const uint64_t b_lo = umma_desc_sw64(s0 + 4 * A_TILE + B_TILE + ks * 32);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t a_hi = umma_desc_sw64(s0 + h * A_TILE + ks * 32);
              const uint64_t a_lo = umma_desc_sw64(s0 + (2 + h) * A_TILE + ks * 32);
              const uint32_t d = tmem_base + h * 256;
              tc_mma_f16(d, a_hi, b_hi, IDESC, (kb | ks) != 0);
              tc_mma_f16(d, a_hi, b_lo, IDESC, 1);
              tc_mma_f16(d, a_lo, b_hi, IDESC, 1);
            }
          }
          tc_commit(&empty[stage]);                 // frees the smem slot when these MMAs retire
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tmem_full);                       // accumulators complete
        tphase ^= 1;
      }
    }
  } else {
    // ===================== epilogue: 4 warps, TMEM lane group = warp % 4 =====================
    const int lg = warp & 3;
    const int row = lg * 32 + lane;                 // accumulator row (focal row v)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    uint32_t tphase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      mbar_wait(tmem_full, tphase, p.err_flag, 4);
      tc_fence_after();
      if constexpr (MODE == 0) {
        // stage-1 product -> split fp16, row-major [v][k] with k = (x | 240 + x)
        const size_t rbase = ((size_t)item * 128 + row) * TC_K;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
#pragma unroll 1
          for (int c = 0; c < TC_NP / 16; ++c) {
            float v[16];
            tc_ld16(lane_addr + h * 256 + c * 16, v);
            tc_wait_ld();
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __half h0 = __float2half_rn(v[2 * i]), h1 = __float2half_rn(v[2 * i + 1]);
              const __half l0 = __float2half_rn(v[2 * i] - __half2float(h0));
              const __half l1 = __float2half_rn(v[2 * i + 1] - __half2float(h1));
              hi[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
              lo[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            uint4* dh = reinterpret_cast<uint4*>(p.T_hi + rbase + h * TC_NP + c * 16);
            uint4* dl = reinterpret_cast<uint4*>(p.T_lo + rbase + h * TC_NP + c * 16);
            dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          }
        }
        tc_fence_before();
        mbar_arrive(tmem_empty);
      } else {
        // fibre projection: c_j = sum_{v,u} F[v][u] (mode_j w)[v][u], F = Fr + i Fi out of TMEM
        double acc[2][AOG_MAX_LP][2];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < AOG_MAX_LP; ++j) acc[h][j][0] = acc[h][j][1] = 0.0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (2 * item + h >= p.num_envs) continue;
#pragma unroll 1
          for (int c = 0; c < TC_NF / 16; ++c) {
            float fr[16], fi[16];
            tc_ld16(lane_addr + h * 256 + c * 16, fr);
            tc_ld16(lane_addr + h * 256 + TC_NF + c * 16, fi);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < AOG_MAX_LP; ++j) {
              if (j < p.J) {
                const float4* w4 = reinterpret_cast<const float4*>(p.lpw + ((size_t)j * TC_NF + row) * TC_NF + c * 16);
                float sr = 0.f, si = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 w = __ldg(w4 + q);
                  sr = fmaf(fr[4 * q + 0], w.x, sr); si = fmaf(fi[4 * q + 0], w.x, si);
                  sr = fmaf(fr[4 * q + 1], w.y, sr); si = fmaf(fi[4 * q + 1], w.y, si);
                  sr = fmaf(fr[4 * q + 2], w.z, sr); si = fmaf(fi[4 * q + 2], w.z, si);
                  sr = fmaf(fr[4 * q + 3], w.w, sr); si = fmaf(fi[4 * q + 3], w.w, si);
                }
                acc[h][j][0] += (double)sr;
                acc[h][j][1] += (double)si;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(tmem_empty);                    // TMEM is free: the next tile's MMAs may start
        double* rbuf = red + (it & 1) * (4 * 2 * AOG_MAX_LP * 2);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < AOG_MAX_LP; ++j)
            if (j < p.J) {
              const double a = warp_sum(acc[h][j][0]), b = warp_sum(acc[h][j][1]);
              if (lane == 0) {
                rbuf[((lg * 2 + h) * AOG_MAX_LP + j) * 2 + 0] = a;
                rbuf[((lg * 2 + h) * AOG_MAX_LP + j) * 2 + 1] = b;
              }
            }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane < 2 * p.J) {
          const int h = lane / p.J, j = lane - h * p.J;
          const int e = 2 * item + h;
          if (e < p.num_envs) {
            double sr = 0.0, si = 0.0;
            for (int g = 0; g < 4; ++g) {
              sr += rbuf[((g * 2 + h) * AOG_MAX_LP + j) * 2 + 0];
              si += rbuf[((g * 2 + h) * AOG_MAX_LP + j) * 2 + 1];
            }
            p.coef[(size_t)e * p.J + j] = make_double2(sr * p.scale.x - si * p.scale.y, sr * p.scale.y + si * p.scale.x);
          }
        }
      }
      tphase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ----------------------------------------------------------------------------- field kernel
// DM surface + pupil field (unit modulus, split fp16, x-major) + Strehl partial sums + obs-arm
// column products, FP32 phase / FP64 reductions.  Block = 4 pupil columns x 64 threads (60 active,
// 4 consecutive y each) looping over ET envs, so each mode value is loaded once per ET envs.
// grid (Np / 4, ceil(nB / ET)), block 256.
template <int ET>
__global__ void __launch_bounds__(256)
k_field_tc(const float* __restrict__ screensT, const double* __restrict__ act, const float* __restrict__ modesT,
           const float* __restrict__ apT, const double2* __restrict__ m1o, __half* __restrict__ E_hi,
           __half* __restrict__ E_lo, double2* __restrict__ R, double2* __restrict__ strehl_part, int K, int n,
           int env0, int nB, int col_origin, double turns_wfs_S, double turns_wfs_s, double turns_sci_S,
           double turns_sci_s, int do_strehl) {
  constexpr int Np = TC_NP;
  extern __shared__ float sh_act[];                      // [ET][K]
  __shared__ double red[ET][4][2][2 * AOG_MAX_OBS + 2];  // [env][column][warp of the column][obs re/im.., strehl re/im]
  const int e0 = blockIdx.y * ET;
  for (int i = threadIdx.x; i < ET * K; i += blockDim.x) {
    const int e = i / K, k = i - e * K;
    sh_act[i] = (e0 + e < nB) ? (float)act[(size_t)(env0 + e0 + e) * K + k] : 0.f;
  }
  __syncthreads();
  const int colw = threadIdx.x >> 6;                     // 0..3 column within the block
  const int t = threadIdx.x & 63;                        // 0..63, active < 60
  const int x = blockIdx.x * 4 + colw;                   // logical pupil column
  const int y0 = 4 * t;
  const bool active = t < Np / 4;
  int xp = x + col_origin;
  if (xp >= Np) xp -= Np;
  float4 ap = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) ap = *reinterpret_cast<const float4*>(apT + (size_t)x * Np + y0);
  const bool lit = active && (ap.x + ap.y + ap.z + ap.w) > 0.f;

  float s[ET][4];
#pragma unroll
  for (int e = 0; e < ET; ++e) s[e][0] = s[e][1] = s[e][2] = s[e][3] = 0.f;
  if (lit) {
    const float* mp = modesT + (size_t)x * Np + y0;
    for (int k = 0; k < K; ++k) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mp + (size_t)k * Np * Np));
#pragma unroll
      for (int e = 0; e < ET; ++e) {
        const float a = sh_act[e * K + k];
        s[e][0] = fmaf(m.x, a, s[e][0]);
        s[e][1] = fmaf(m.y, a, s[e][1]);
        s[e][2] = fmaf(m.z, a, s[e][2]);
        s[e][3] = fmaf(m.w, a, s[e][3]);
      }
    }
  }
  // obs-arm table rows for this thread's 4 pixels: m1o[v][y0..y0+3]
  const float apv[4] = {ap.x, ap.y, ap.z, ap.w};
  const int wcol = (threadIdx.x >> 5) & 1;               // which of the column's two warps
#pragma unroll
  for (int e = 0; e < ET; ++e) {
    const bool env_ok = e0 + e < nB;
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, sn[4] = {0.f, 0.f, 0.f, 0.f};
    double st_re = 0.0, st_im = 0.0;
    if (lit && env_ok) {
      const float4 S4 = *reinterpret_cast<const float4*>(screensT + ((size_t)(env0 + e0 + e) * Np + xp) * Np + y0);
      const float Sv[4] = {S4.x, S4.y, S4.z, S4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // phase in turns, reduced to [-1/2, 1/2] in FP64, then sin/cos(2 pi t) in FP32
        double tw = (double)Sv[i] * turns_wfs_S + (double)s[e][i] * turns_wfs_s;
        tw -= rint(tw);
        float sv, cv;
        sincospif((float)(2.0 * tw), &sv, &cv);
        cs[i] = cv * apv[i];
        sn[i] = sv * apv[i];
        if (do_strehl) {
          double ts = (double)Sv[i] * turns_sci_S + (double)s[e][i] * turns_sci_s;
          ts -= rint(ts);
          sincospif((float)(2.0 * ts), &sv, &cv);
          st_re += (double)(cv * apv[i]);
          st_im += (double)(sv * apv[i]);
        }
      }
    }
    if (active && env_ok) {
      // split fp16 operand rows: E[x][k = y] = re, E[x][k = 240 + y] = im
      uint32_t hr[2], lr[2], hi_[2], li_[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const __half a0 = __float2half_rn(cs[2 * i]), a1 = __float2half_rn(cs[2 * i + 1]);
        const __half b0 = __float2half_rn(cs[2 * i] - __half2float(a0)), b1 = __float2half_rn(cs[2 * i + 1] - __half2float(a1));
        const __half c0 = __float2half_rn(sn[2 * i]), c1 = __float2half_rn(sn[2 * i + 1]);
        const __half d0 = __float2half_rn(sn[2 * i] - __half2float(c0)), d1 = __float2half_rn(sn[2 * i + 1] - __half2float(c1));
        hr[i] = (uint32_t)__half_as_ushort(a0) | ((uint32_t)__half_as_ushort(a1) << 16);
        lr[i] = (uint32_t)__half_as_ushort(b0) | ((uint32_t)__half_as_ushort(b1) << 16);
        hi_[i] = (uint32_t)__half_as_ushort(c0) | ((uint32_t)__half_as_ushort(c1) << 16);
        li_[i] = (uint32_t)__half_as_ushort(d0) | ((uint32_t)__half_as_ushort(d1) << 16);
      }
      const size_t o = ((size_t)(e0 + e) * Np + x) * TC_K + y0;
      *reinterpret_cast<uint2*>(E_hi + o) = make_uint2(hr[0], hr[1]);
      *reinterpret_cast<uint2*>(E_lo + o) = make_uint2(lr[0], lr[1]);
      *reinterpret_cast<uint2*>(E_hi + o + Np) = make_uint2(hi_[0], hi_[1]);
      *reinterpret_cast<uint2*>(E_lo + o + Np) = make_uint2(li_[0], li_[1]);
    }
    // obs arm: r[v] = sum_y m1o[v][y] E[y][x] over this thread's pixels, then over the column
    for (int v = 0; v < n; ++v) {
      double re = 0.0, im = 0.0;
      if (lit && env_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 m = m1o[(size_t)v * Np + y0 + i];
          re += m.x * (double)cs[i] - m.y * (double)sn[i];
          im += m.x * (double)sn[i] + m.y * (double)cs[i];
        }
      }
      re = warp_sum(re);
      im = warp_sum(im);
      if ((threadIdx.x & 31) == 0) { red[e][colw][wcol][2 * v] = re; red[e][colw][wcol][2 * v + 1] = im; }
    }
    if (do_strehl) {
      st_re = warp_sum(st_re);
      st_im = warp_sum(st_im);
      if ((threadIdx.x & 31) == 0) { red[e][colw][wcol][2 * n] = st_re; red[e][colw][wcol][2 * n + 1] = st_im; }
    }
  }
  __syncthreads();
  // column sums -> R[env][x][v];  block Strehl partial -> strehl_part[env][blockIdx.x]
  for (int i = threadIdx.x; i < ET * 4 * n; i += blockDim.x) {
    const int e = i / (4 * n), r = i - e * 4 * n, c = r / n, v = r - c * n;
    if (e0 + e < nB) {
      const double re = red[e][c][0][2 * v] + red[e][c][1][2 * v];
      const double im = red[e][c][0][2 * v + 1] + red[e][c][1][2 * v + 1];
      R[((size_t)(e0 + e) * Np + blockIdx.x * 4 + c) * n + v] = make_double2(re, im);
    }
  }
  if (do_strehl && threadIdx.x < ET && e0 + threadIdx.x < nB) {
    double re = 0.0, im = 0.0;
    for (int c = 0; c < 4; ++c)
      for (int w = 0; w < 2; ++w) { re += red[threadIdx.x][c][w][2 * n]; im += red[threadIdx.x][c][w][2 * n + 1]; }
    strehl_part[(size_t)(e0 + threadIdx.x) * gridDim.x + blockIdx.x] = make_double2(re, im);
  }
}

// screens FP64 [B][y][xp] -> FP32 transposed [B][xp][y]   (full refresh)
__global__ void k_screens_to_T(const double* __restrict__ src, float* __restrict__ dst, int Np) {
  __shared__ float tile[32][33];
  const size_t b = blockIdx.z;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int y = y0 + j, x = x0 + threadIdx.x;
    if (y < Np && x < Np) tile[j][threadIdx.x] = (float)src[(b * Np + y) * Np + x];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int x = x0 + j, y = y0 + threadIdx.x;
    if (y < Np && x < Np) dst[(b * Np + x) * Np + y] = tile[threadIdx.x][j];
  }
}
// one physical column refresh after an extrusion
__global__ void k_column_to_T(const double* __restrict__ src, float* __restrict__ dst, int Np, int phys_col) {
  const size_t b = blockIdx.y;
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  if (y < Np) dst[(b * Np + phys_col) * Np + y] = (float)src[(b * Np + y) * Np + phys_col];
}

// ----------------------------------------------------------------------------- host helpers
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp16 row-major [rows][480] tensor, box = KB x box_rows, 64-byte swizzle.
int make_map(aog_env* env, CUtensorMap* map, const __half* ptr, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) AOG_FAIL(AOG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {(cuuint64_t)TC_K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)TC_K * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)KB, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) AOG_FAIL(AOG_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return AOG_OK;
}

template <typename T>
int talloc(aog_env* env, T** p, size_t count) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  AOG_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
  return AOG_OK;
}

void split_half(double x, __half* hi, __half* lo) {
  const __half h = __float2half_rn((float)x);
  *hi = h;
  *lo = __float2half_rn((float)(x - (double)__half2float(h)));
}

int upload_split(aog_env* env, const std::vector<double>& m, __half* d_hi, __half* d_lo) {
  std::vector<__half> hi(m.size()), lo(m.size());
  for (size_t i = 0; i < m.size(); ++i) split_half(m[i], &hi[i], &lo[i]);
  AOG_CUDA(cudaMemcpy(d_hi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice));
  AOG_CUDA(cudaMemcpy(d_lo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice));
  return AOG_OK;
}

}  // namespace

// =============================================================================================
int aog_tensor_create(aog_env* env) {
  const aog_config& c = env->cfg;
  if (c.num_pupil_pixels != TC_NP || c.num_focal_pixels != TC_NF)
    AOG_FAIL(AOG_ERR_UNSUPPORTED, "tensor precision path is built for num_pupil_pixels=240, num_focal_pixels=128; "
                                  "use precision='f64' for other grids");
  TensorState* ts = new TensorState();
  env->tensor_state = ts;
  cudaDeviceProp prop;
  AOG_CUDA(cudaGetDeviceProperties(&prop, c.device));
  ts->num_sms = prop.multiProcessorCount;
  const size_t ch = env->chunk, B = c.num_envs, P = env->P;
  int rc;
#define A(expr) if ((rc = (expr))) return rc
  A(talloc(env, &ts->A1_hi, (size_t)256 * TC_K));
  A(talloc(env, &ts->A1_lo, (size_t)256 * TC_K));
  A(talloc(env, &ts->B2_hi, (size_t)256 * TC_K));
  A(talloc(env, &ts->B2_lo, (size_t)256 * TC_K));
  A(talloc(env, &ts->E_hi, ch * TC_NP * TC_K));
  A(talloc(env, &ts->E_lo, ch * TC_NP * TC_K));
  // stage-2 reads env pairs: keep one spare env of rows so the last odd pair stays in bounds
  A(talloc(env, &ts->T_hi, (ch + 1) * 128 * TC_K));
  A(talloc(env, &ts->T_lo, (ch + 1) * 128 * TC_K));
  AOG_CUDA(cudaMemset(ts->T_hi, 0, (ch + 1) * 128 * TC_K * sizeof(__half)));
  AOG_CUDA(cudaMemset(ts->T_lo, 0, (ch + 1) * 128 * TC_K * sizeof(__half)));
  A(talloc(env, &ts->lpw, (size_t)c.num_lp_modes * env->NF2));
  A(talloc(env, &ts->screensT, B * P));
  AOG_CUDA(cudaMemset(ts->screensT, 0, B * P * sizeof(float)));
  A(talloc(env, &ts->modesT, (size_t)c.num_modes * P));
  A(talloc(env, &ts->apT, P));
  A(talloc(env, &ts->m2oT, (size_t)c.obs_dim * TC_NP));
  A(talloc(env, &ts->err_flag, 1));
  AOG_CUDA(cudaMemset(ts->err_flag, 0, sizeof(int)));
  A(make_map(env, &ts->tmA1_hi, ts->A1_hi, 256, 128));
  A(make_map(env, &ts->tmA1_lo, ts->A1_lo, 256, 128));
  A(make_map(env, &ts->tmE_hi, ts->E_hi, ch * TC_NP, TC_NP));
  A(make_map(env, &ts->tmE_lo, ts->E_lo, ch * TC_NP, TC_NP));
  A(make_map(env, &ts->tmT_hi, ts->T_hi, (ch + 1) * 128, 128));
  A(make_map(env, &ts->tmT_lo, ts->T_lo, (ch + 1) * 128, 128));
  A(make_map(env, &ts->tmB2_hi, ts->B2_hi, 256, 256));
  A(make_map(env, &ts->tmB2_lo, ts->B2_lo, 256, 256));
#undef A
  AOG_CUDA(cudaFuncSetAttribute(k_mft_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  AOG_CUDA(cudaFuncSetAttribute(k_mft_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return AOG_OK;
}

void aog_tensor_destroy(aog_env* env) {
  TensorState* ts = TS(env);
  if (!ts) return;
  void* ptrs[] = {ts->A1_hi, ts->A1_lo, ts->B2_hi, ts->B2_lo, ts->E_hi, ts->E_lo, ts->T_hi, ts->T_lo, ts->lpw,
                  ts->screensT, ts->modesT, ts->apT, ts->m2oT, ts->err_flag};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete ts;
  env->tensor_state = nullptr;
}

int aog_tensor_table_updated(aog_env* env, int which, const void* host) {
  TensorState* ts = TS(env);
  const aog_config& c = env->cfg;
  const int Np = TC_NP, Nf = TC_NF;
  if (which == AOG_TABLE_MFT_FIB_1) {
    // M1 [Nf][Np] complex, every entry of modulus w (the pupil grid weight): factor it out.
    const double* m = static_cast<const double*>(host);
    const double w = std::hypot(m[0], m[1]);
    ts->pupil_weight = w;
    std::vector<double> a((size_t)256 * TC_K);
    for (int v = 0; v < Nf; ++v)
      for (int y = 0; y < Np; ++y) {
        const double re = m[2 * ((size_t)v * Np + y)] / w, im = m[2 * ((size_t)v * Np + y) + 1] / w;
        a[(size_t)v * TC_K + y] = re;               // Tr rows:  [M1r | -M1i]
        a[(size_t)v * TC_K + Np + y] = -im;
        a[(size_t)(128 + v) * TC_K + y] = im;       // Ti rows:  [M1i |  M1r]
        a[(size_t)(128 + v) * TC_K + Np + y] = re;
      }
    int rc = upload_split(env, a, ts->A1_hi, ts->A1_lo);
    if (rc) return rc;
    ts->have_m1 = true;
  } else if (which == AOG_TABLE_MFT_FIB_2) {
    // M2 [Np][Nf] complex (unit modulus).  Operand rows n = (u | 128 + u), K = (x | 240 + x):
    //   Fr[u] = sum_x Tr[x] M2r[x][u] - Ti[x] M2i[x][u];   Fi[u] = sum_x Tr[x] M2i[x][u] + Ti[x] M2r[x][u]
    const double* m = static_cast<const double*>(host);
    std::vector<double> b((size_t)256 * TC_K);
    for (int x = 0; x < Np; ++x)
      for (int u = 0; u < Nf; ++u) {
        const double re = m[2 * ((size_t)x * Nf + u)], im = m[2 * ((size_t)x * Nf + u) + 1];
        b[(size_t)u * TC_K + x] = re;
        b[(size_t)u * TC_K + Np + x] = -im;
        b[(size_t)(128 + u) * TC_K + x] = im;
        b[(size_t)(128 + u) * TC_K + Np + x] = re;
      }
    int rc = upload_split(env, b, ts->B2_hi, ts->B2_lo);
    if (rc) return rc;
    ts->have_m2 = true;
  } else if (which == AOG_TABLE_LP_MODES_W) {
    const double* m = static_cast<const double*>(host);
    const size_t cnt = (size_t)c.num_lp_modes * env->NF2;
    double mx = 0.0;
    for (size_t i = 0; i < cnt; ++i) mx = std::max(mx, std::fabs(m[i]));
    if (mx == 0.0) mx = 1.0;
    ts->lpw_scale = mx;
    std::vector<float> f(cnt);
    for (size_t i = 0; i < cnt; ++i) f[i] = (float)(m[i] / mx);
    AOG_CUDA(cudaMemcpy(ts->lpw, f.data(), cnt * sizeof(float), cudaMemcpyHostToDevice));
    ts->have_lp = true;
  } else if (which == AOG_TABLE_DM_MODES) {
    const double* m = static_cast<const double*>(host);
    std::vector<float> f((size_t)c.num_modes * env->P);
    for (int k = 0; k < c.num_modes; ++k)
      for (int y = 0; y < Np; ++y)
        for (int x = 0; x < Np; ++x) f[((size_t)k * Np + x) * Np + y] = (float)m[((size_t)k * Np + y) * Np + x];
    AOG_CUDA(cudaMemcpy(ts->modesT, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
    ts->have_modes = true;
  } else if (which == AOG_TABLE_APERTURE) {
    const double* m = static_cast<const double*>(host);
    std::vector<float> f(env->P);
    for (int y = 0; y < Np; ++y)
      for (int x = 0; x < Np; ++x) f[(size_t)x * Np + y] = (float)m[(size_t)y * Np + x];
    AOG_CUDA(cudaMemcpy(ts->apT, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
    ts->have_ap = true;
  } else if (which == AOG_TABLE_MFT_OBS_2) {
    const double* m = static_cast<const double*>(host);   // [Np][n] complex -> [n][Np]
    const int n = c.obs_dim;
    std::vector<double> f((size_t)2 * n * Np);
    for (int x = 0; x < Np; ++x)
      for (int u = 0; u < n; ++u) {
        f[2 * ((size_t)u * Np + x)] = m[2 * ((size_t)x * n + u)];
        f[2 * ((size_t)u * Np + x) + 1] = m[2 * ((size_t)x * n + u) + 1];
      }
    AOG_CUDA(cudaMemcpy(ts->m2oT, f.data(), f.size() * sizeof(double), cudaMemcpyHostToDevice));
    ts->have_m2o = true;
  }
  return AOG_OK;
}

int aog_tensor_screens_updated(aog_env* env) {
  TensorState* ts = TS(env);
  const int Np = TC_NP;
  dim3 g(cdiv(Np, 32), cdiv(Np, 32), env->cfg.num_envs), b(32, 8);
  k_screens_to_T<<<g, b>>>(env->screens, ts->screensT, Np);
  AOG_LAUNCH_CHECK();
  AOG_CUDA(cudaDeviceSynchronize());
  return AOG_OK;
}

int aog_tensor_column_updated(aog_env* env, int phys_col, cudaStream_t st) {
  TensorState* ts = TS(env);
  dim3 g(cdiv(TC_NP, 128), env->cfg.num_envs);
  k_column_to_T<<<g, 128, 0, st>>>(env->screens, ts->screensT, TC_NP, phys_col);
  AOG_LAUNCH_CHECK();
  return AOG_OK;
}

// Debug read-back (tests): the split-fp16 intermediates of the LAST chunk processed, recombined
// hi + lo and returned in the FP64 path's conventions.
//   which = AOG_FIELD_TC_PUPIL : pupil field [y][x] complex (unit modulus x aperture)
//   which = AOG_FIELD_TC_STAGE1: stage-1 product [v][x] complex (unit-modulus twiddles)
int aog_tensor_get_field(aog_env* env, int which, int env_in_chunk, double* host_out, size_t count) {
  TensorState* ts = TS(env);
  if (!ts) AOG_FAIL(AOG_ERR_STATE, "handle has no tensor path");
  if (env_in_chunk < 0 || env_in_chunk >= env->chunk) AOG_FAIL(AOG_ERR_INVALID, "env index outside the chunk");
  const int rows = (which == AOG_FIELD_TC_PUPIL) ? TC_NP : 128;
  if (count != (size_t)2 * rows * TC_NP) AOG_FAIL(AOG_ERR_INVALID, "count");
  const __half* dh = (which == AOG_FIELD_TC_PUPIL) ? ts->E_hi : ts->T_hi;
  const __half* dl = (which == AOG_FIELD_TC_PUPIL) ? ts->E_lo : ts->T_lo;
  const size_t nel = (size_t)rows * TC_K;
  std::vector<__half> hi(nel), lo(nel);
  AOG_CUDA(cudaMemcpy(hi.data(), dh + (size_t)env_in_chunk * nel, nel * sizeof(__half), cudaMemcpyDeviceToHost));
  AOG_CUDA(cudaMemcpy(lo.data(), dl + (size_t)env_in_chunk * nel, nel * sizeof(__half), cudaMemcpyDeviceToHost));
  auto val = [&](size_t i) { return (double)__half2float(hi[i]) + (double)__half2float(lo[i]); };
  if (which == AOG_FIELD_TC_PUPIL) {          // stored [x][k = y | 240 + y]  ->  out [y][x]
    for (int x = 0; x < TC_NP; ++x)
      for (int y = 0; y < TC_NP; ++y) {
        host_out[2 * ((size_t)y * TC_NP + x)] = val((size_t)x * TC_K + y);
        host_out[2 * ((size_t)y * TC_NP + x) + 1] = val((size_t)x * TC_K + TC_NP + y);
      }
  } else {                                    // stored [v][k = x | 240 + x]  ->  out [v][x]
    for (int v = 0; v < 128; ++v)
      for (int x = 0; x < TC_NP; ++x) {
        host_out[2 * ((size_t)v * TC_NP + x)] = val((size_t)v * TC_K + x);
        host_out[2 * ((size_t)v * TC_NP + x) + 1] = val((size_t)v * TC_K + TC_NP + x);
      }
  }
  return AOG_OK;
}

int aog_tensor_check(aog_env* env) {
  TensorState* ts = TS(env);
  if (!ts) return AOG_OK;
  int flag = 0;
  AOG_CUDA(cudaMemcpy(&flag, ts->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) AOG_FAIL(AOG_ERR_CUDA, "tensor pipeline barrier timeout, code " + std::to_string(flag));
  return AOG_OK;
}

int aog_tensor_optics(aog_env* env, bool flat_dm, bool with_reward, const aog_outputs& out, cudaStream_t st) {
  TensorState* ts = TS(env);
  const aog_config& c = env->cfg;
  if (!(ts->have_m1 && ts->have_m2 && ts->have_lp && ts->have_modes && ts->have_ap && ts->have_m2o))
    AOG_FAIL(AOG_ERR_STATE, "tensor path tables incomplete");
  (void)flat_dm;
  const int Np = TC_NP, n = c.obs_dim, K = c.num_modes, J = c.num_lp_modes, B = c.num_envs;
  const bool strehl = with_reward && c.rew_type == AOG_REW_STREHL_RATIO;
  constexpr int ET = 8;
  const double two_pi = 6.283185307179586476925286766559;
  // phase [turns] = S / (lambda 2 pi) + s * 2 / lambda
  const double tw_S = 1.0 / (c.wavelength_wfs * two_pi), tw_s = 2.0 / c.wavelength_wfs;
  const double tsS = 1.0 / (c.wavelength_sci * two_pi), tss = 2.0 / c.wavelength_sci;
  const double2 norm = make_double2(c.mft_norm_re * c.amp_fiber, c.mft_norm_im * c.amp_fiber);
  for (int e0 = 0; e0 < B; e0 += env->chunk) {
    const int nB = std::min(env->chunk, B - e0);
    {
      dim3 g(Np / 4, cdiv(nB, ET));
      k_field_tc<ET><<<g, 256, ET * K * sizeof(float), st>>>(
          ts->screensT, env->act, ts->modesT, ts->apT, env->t_m1o, ts->E_hi, ts->E_lo, env->bufR, env->strehl_part, K,
          n, e0, nB, (int)env->cnt.column_origin, tw_S, tw_s, tsS, tss, strehl ? 1 : 0);
      AOG_LAUNCH_CHECK();
    }
    if (env->timing) { AOG_CUDA(cudaEventRecord(env->ev0, st)); }
    TcParams p{};
    p.num_envs = nB;
    p.err_flag = ts->err_flag;
    p.T_hi = ts->T_hi; p.T_lo = ts->T_lo;
    p.lpw = ts->lpw; p.coef = env->coef; p.J = J;
    const double sc = c.amp_fiber * ts->pupil_weight * ts->lpw_scale;
    p.scale = make_double2(c.mft_norm_re * sc, c.mft_norm_im * sc);
    p.num_items = nB;
    k_mft_tc<0><<<std::min(ts->num_sms, p.num_items), TC_THREADS, SMEM_BYTES, st>>>(ts->tmA1_hi, ts->tmA1_lo, ts->tmE_hi,
                                                                                    ts->tmE_lo, p);
    AOG_LAUNCH_CHECK();
    if (with_reward) {
      p.num_items = (nB + 1) / 2;
      k_mft_tc<1><<<std::min(ts->num_sms, p.num_items), TC_THREADS, SMEM_BYTES, st>>>(ts->tmT_hi, ts->tmT_lo, ts->tmB2_hi,
                                                                                      ts->tmB2_lo, p);
      AOG_LAUNCH_CHECK();
    }
    if (env->timing) { AOG_CUDA(cudaEventRecord(env->ev1, st)); env->ev_valid = true; }
    FinalizeArgs a{};
    a.R = env->bufR; a.m1o = ts->m2oT; a.coef = env->coef; a.lpphase = env->t_lpphase; a.lpgram = env->t_lpgram;
    a.strehl_part = env->strehl_part; a.strehl_blocks = Np / 4;
    a.Np = Np; a.n = n; a.J = J; a.rew_type = c.rew_type; a.has_thr = c.has_rew_threshold;
    a.compute_reward = with_reward ? 1 : 0;
    a.transpose_out = 1;     // R is [x][v] and the table is M2o^T: results come out as (u, v)
    a.thr = c.rew_threshold; a.obs_weight = c.obs_weight; a.strehl_scale = c.strehl_scale; a.ssim_peak = c.ssim_ref_peak;
    a.norm = norm;
    const size_t n2 = (size_t)env->n2;
    a.obs16 = out.obs_f16 ? out.obs_f16 + (size_t)e0 * n2 : nullptr;
    a.obs64 = out.obs_f64 ? out.obs_f64 + (size_t)e0 * n2 : nullptr;
    a.reward = out.reward ? out.reward + e0 : nullptr;
    a.power = out.power ? out.power + e0 : nullptr;
    a.strehl = out.strehl ? out.strehl + e0 : nullptr;
    a.ssim = out.ssim ? out.ssim + e0 : nullptr;
    k_finalize<<<nB, 64, 0, st>>>(a);
    AOG_LAUNCH_CHECK();
  }
  return AOG_OK;
}
