// tc_gemm.cu — C[M,N] (+)= A[M,K] · Bk[N,K]ᵀ for a LONG M and a weight-sized N = K ∈ {64, 128}, on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA) with the 3×TF32 split that keeps
// fp32-level accuracy:   A·B ≈ A_hi·B_hi + A_lo·B_hi + A_hi·B_lo,   x_hi = tf32(x) (round to nearest), x_lo = x − x_hi.
// The dropped lo·lo term and the truncation of the lo operands are ~2⁻²¹ relative per product — the fp32 SIMT GEMM of
// dense.cu is ~2⁻²⁴ — so the layer-level parity bar (1e-5 against the fp64 oracle) holds with room.
//
// This is the dense per-relation transform Z_i = S_i·W_i and the self transform H·W_s of the reference's layer
// (prediction/IDDGCN.py:70-76) and their data gradients, at sizes where they are a real contraction (M ≥ 32768 rows);
// at the reference's own 845-node graph the SIMT path of dense.cu is used.
//
// One persistent CTA per SM, 12 warps with fixed roles:
//   warp 0      TMA producer: streams A in k-blocks of 32 floats (128 rows × 128 B, SWIZZLE_128B) into a ring of stages
//   warp 1      MMA issuer: one thread issues 3 × 4 tcgen05.mma (128 × N × 8) per k-block; owns the TMEM allocation
//   warps 4-7   epilogue: tcgen05.ld the finished 128 × N accumulator (one row per thread), add C if asked, store
//   warps 8-11  splitter: rewrite a landed A k-block in place as A_hi and write A_lo beside it (same swizzled layout)
// B_hi / B_lo (pre-split by a small kernel, weight-sized) stay resident in shared memory for the CTA's lifetime.
// Two accumulators (2 × N TMEM columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <stdlib.h>
#include "dense.cuh"

namespace iddgcn {

namespace {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_KB = 32;                              // floats per k-block = one 128-byte swizzle row
constexpr int TC_THREADS = 384;
constexpr int TC_PREFETCH_TILES = 1;                   // L2 prefetch distance in tiles of this CTA (swept 0-4: 1 is best)
constexpr int TC_A_SLAB = TC_BLOCK_M * TC_KB * 4;      // 16 KB: one A k-block (hi or lo)

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spins on the phase with the given parity.  A single wait that lasts ~10 s of SM clock can only be a protocol bug
// (legitimate waits are microseconds, and a time-sliced GPU stretches them by slices, not by seconds): trap instead of
// hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 20000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// pull a box into L2 only (no shared-memory slot needed): shortens the latency of the real load issued later
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem] · B[smem], tf32 inputs, fp32 accumulate; issued by ONE thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes × 32 consecutive fp32 columns: thread = TMEM lane (accumulator row), v[j] = column c0 + j
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 256-bit global accesses (sm_100): one full 32-byte sector per lane
__device__ __forceinline__ void st_global_v8(float* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_v8(const float* p, float* o) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7])
               : "l"(p)
               : "memory");
}

// Shared-memory matrix descriptor of a K-major operand slab: rows of 128 bytes (32 tf32), SWIZZLE_128B, 8-row groups
// 1024 bytes apart.  Fields (sm_100 UMMA descriptor): [0,14) start address >> 4, [16,30) leading byte offset >> 4
// (unused for swizzled K-major: 1), [32,46) stride byte offset >> 4 (1024 >> 4 = 64), [46,48) version = 1,
// [61,64) layout type = 2 (SWIZZLE_128B).  Stepping 8 tf32 (32 bytes) along K inside the swizzle row adds 2 to the
// start-address field.
__device__ __forceinline__ uint64_t kmajor_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// Instruction descriptor, kind::tf32: [4,6) D format = 1 (f32), [7,10) A format = 2 (tf32), [10,13) B format = 2,
// [15] A major = 0 (K), [16] B major = 0 (K), [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t tf32_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- weight split (and optional transpose): hi = tf32(x), lo = x − hi ------------------------------------------------
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ float tf32_trunc(float x) {   // the 19 bits a kind::tf32 operand is read with
  return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

// in: [batch][rows][cols] row-major.  transpose == 0: hi/lo are [batch][rows][cols]; else [batch][cols][rows].
__global__ void split_tf32_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo,
                                  int rows, int cols, int64_t in_stride, int transpose) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i % cols;
  const float x = in[(int64_t)b * in_stride + (int64_t)r * cols + c];
  const float h = tf32_round(x);
  const int64_t o = (int64_t)b * rows * cols + (transpose ? (int64_t)c * rows + r : (int64_t)i);
  hi[o] = h;
  lo[o] = tf32_round(x - h);   // rounded here so that the tensor core's own truncation of the operand is a no-op
}

// ---- the GEMM kernel ---------------------------------------------------------------------------------------------------
template <int N_, int K_, int STAGES>
struct TcSmem {
  static constexpr int KBLOCKS = K_ / TC_KB;
  static constexpr int B_SLAB = N_ * TC_KB * 4;                       // one B k-block (hi or lo)
  static constexpr int B_BYTES = 2 * KBLOCKS * B_SLAB;                 // hi then lo
  static constexpr int A_BYTES = STAGES * 2 * TC_A_SLAB;               // per stage: hi slab, lo slab
  static constexpr int BAR_OFF = B_BYTES + A_BYTES;
  static constexpr int NBARS = 3 * STAGES + 4 + 1;                     // full, ready, empty; tmem full/empty x2; b_full
  static constexpr int BYTES = BAR_OFF + NBARS * 8 + 16;
  static constexpr int DYN_BYTES = BYTES + 1024;                       // slack to align the base to 1024
};

template <int N_, int K_, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                      const __grid_constant__ CUtensorMap map_blo, float* __restrict__ C, int64_t ldc, int64_t M,
                      int accumulate, int hw_trunc) {
  using L = TcSmem<N_, K_, STAGES>;
  constexpr int KBLOCKS = L::KBLOCKS;
  constexpr int TMEM_COLS = 2 * N_;
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B slabs need 1024-byte alignment
  uint8_t* base_ptr = tc_smem_raw + (base - smem_u32(tc_smem_raw));
  const uint32_t sB = base, sA = base + L::B_BYTES, sBar = base + L::BAR_OFF;
  auto bar_full = [&](int s) { return sBar + 8u * s; };
  auto bar_ready = [&](int s) { return sBar + 8u * (STAGES + s); };
  auto bar_empty = [&](int s) { return sBar + 8u * (2 * STAGES + s); };
  auto bar_tfull = [&](int a) { return sBar + 8u * (3 * STAGES + a); };
  auto bar_tempty = [&](int a) { return sBar + 8u * (3 * STAGES + 2 + a); };
  const uint32_t bar_b = sBar + 8u * (3 * STAGES + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::BAR_OFF + L::NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tiles = (M + TC_BLOCK_M - 1) / TC_BLOCK_M;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_bhi);
    tma_prefetch_desc(&map_blo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_ready(s), 128);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull(a), 1);
      mbar_init(bar_tempty(a), 128);
    }
    mbar_init(bar_b, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32((const void*)tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(bar_b, (uint32_t)L::B_BYTES);
      for (int kb = 0; kb < KBLOCKS; ++kb) {
        tma_load_2d(sB + kb * L::B_SLAB, &map_bhi, bar_b, kb * TC_KB, 0);
        tma_load_2d(sB + (KBLOCKS + kb) * L::B_SLAB, &map_blo, bar_b, kb * TC_KB, 0);
      }
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t ahead = tile + (int64_t)TC_PREFETCH_TILES * gridDim.x;   // this CTA's next tile: into L2 now
        if (ahead < tiles)
          for (int kb = 0; kb < KBLOCKS; ++kb) tma_prefetch_l2_2d(&map_a, kb * TC_KB, (int)(ahead * TC_BLOCK_M));
        for (int kb = 0; kb < KBLOCKS; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(bar_empty(s), ph ^ 1u);
          mbar_expect_tx(bar_full(s), (uint32_t)TC_A_SLAB);
          tma_load_2d(sA + s * 2 * TC_A_SLAB, &map_a, bar_full(s), kb * TC_KB, (int)(tile * TC_BLOCK_M));
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = tf32_idesc(TC_BLOCK_M, N_);
      mbar_wait(bar_b, 0);
      uint32_t it = 0, t_local = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++t_local) {
        const int a = t_local & 1;
        const uint32_t aph = (t_local >> 1) & 1u;
        mbar_wait(bar_tempty(a), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * N_);
        for (int kb = 0; kb < KBLOCKS; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(bar_ready(s), ph);
          tc_fence_after();
          const uint64_t a_hi = kmajor_sw128_desc(sA + s * 2 * TC_A_SLAB);
          const uint64_t a_lo = kmajor_sw128_desc(sA + s * 2 * TC_A_SLAB + TC_A_SLAB);
          const uint64_t b_hi = kmajor_sw128_desc(sB + kb * L::B_SLAB);
          const uint64_t b_lo = kmajor_sw128_desc(sB + (KBLOCKS + kb) * L::B_SLAB);
#pragma unroll
          for (int k = 0; k < TC_KB / 8; ++k) {               // 8 tf32 = 32 bytes per MMA along K
            const uint64_t adv = (uint64_t)(k * 2);
            umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
            umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
            umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
          }
          umma_commit(bar_empty(s));                          // stage s may be overwritten once these MMAs are done
        }
        umma_commit(bar_tfull(a));                            // accumulator a is complete
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue: warp q reads TMEM lanes [32q, 32q+32) =====
    const int q = warp & 3;
    uint32_t t_local = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++t_local) {
      const int a = t_local & 1;
      const uint32_t aph = (t_local >> 1) & 1u;
      mbar_wait(bar_tfull(a), aph);
      tc_fence_after();
      const int64_t row = tile * TC_BLOCK_M + q * 32 + lane;
      float* crow = C + row * ldc;
#pragma unroll 1
      for (int c0 = 0; c0 < N_; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * N_ + c0), v);
        if (row < M) {
          if (accumulate) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float old[8];
              ld_global_v8(crow + c0 + j, old);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[j + e] = __float_as_uint(__uint_as_float(v[j + e]) + old[e]);
            }
          }
#pragma unroll
#pragma unroll
          for (int j = 0; j < 32; j += 8) st_global_v8(crow + c0 + j, v + j);
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty(a));
    }
  } else if (warp >= 8) {
    // ===== splitter: A k-block (raw fp32, as landed by TMA) -> hi in place, lo in the neighbouring slab =====
    const int t = threadIdx.x - 256;                           // 0..127
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int kb = 0; kb < KBLOCKS; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        mbar_wait(bar_full(s), ph);
        float4* hi = reinterpret_cast<float4*>(base_ptr + L::B_BYTES + (size_t)s * 2 * TC_A_SLAB);
        float4* lo = hi + TC_A_SLAB / 16;
        if (hw_trunc) {
          // A_hi is what the tensor core reads from the raw fp32 block anyway (kind::tf32 ignores the low 13 mantissa
          // bits): leave the landed block untouched and write only A_lo = A - trunc(A) — half the splitter's stores
#pragma unroll
          for (int i = 0; i < TC_A_SLAB / 16 / 128; ++i) {
            const int o = i * 128 + t;
            const float4 x = hi[o];
            lo[o] = make_float4(tf32_round(x.x - tf32_trunc(x.x)), tf32_round(x.y - tf32_trunc(x.y)),
                                tf32_round(x.z - tf32_trunc(x.z)), tf32_round(x.w - tf32_trunc(x.w)));
          }
        } else {
#pragma unroll
          for (int i = 0; i < TC_A_SLAB / 16 / 128; ++i) {
            const int o = i * 128 + t;
            const float4 x = hi[o];
            const float4 h = make_float4(tf32_round(x.x), tf32_round(x.y), tf32_round(x.z), tf32_round(x.w));
            hi[o] = h;
            lo[o] = make_float4(tf32_round(x.x - h.x), tf32_round(x.y - h.y), tf32_round(x.z - h.z),
                                tf32_round(x.w - h.w));
          }
        }
        fence_proxy_async();                                   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(bar_ready(s));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// fp32 [rows, cols] row-major with leading dimension ld (floats); box = 32 columns × box_rows rows, SWIZZLE_128B
int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows,
             CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = encode_tiled_fn();
  IDD_CHECK(enc != nullptr, IDDGCN_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IDD_CHECK(r == CUDA_SUCCESS, IDDGCN_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return IDDGCN_OK;
}

template <int N_, int K_, int STAGES>
int configure_tc() {  // opt in to the large dynamic shared-memory carve-out, once per device
  using L = TcSmem<N_, K_, STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  IDD_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[dev]) {
    IDD_CUDA(cudaFuncSetAttribute(tc_gemm_nt_kernel<N_, K_, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  L::DYN_BYTES));
    if (dev < 64) configured[dev] = true;
  }
  return IDDGCN_OK;
}

template <int N_, int K_, int STAGES>
int launch_tc(const float* A, int64_t lda, const float* Bhi, const float* Blo, float* C, int64_t ldc, int64_t M,
              int accumulate, cudaStream_t st) {
  using L = TcSmem<N_, K_, STAGES>;
  static_assert(L::DYN_BYTES <= 227 * 1024, "shared-memory budget exceeded");
  auto kern = tc_gemm_nt_kernel<N_, K_, STAGES>;
  IDD_TRY((configure_tc<N_, K_, STAGES>()));
  CUtensorMap ma, mh, ml;
  IDD_TRY(make_map(&ma, A, M, K_, lda, TC_BLOCK_M));
  IDD_TRY(make_map(&mh, Bhi, N_, K_, K_, N_));
  IDD_TRY(make_map(&ml, Blo, N_, K_, K_, N_));
  const int64_t tiles = (M + TC_BLOCK_M - 1) / TC_BLOCK_M;
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  // A_hi from the tensor core's own truncation (default; IDDGCN_TC_AHI_TRUNC=0 restores the rounded in-place split):
  // 2.18 -> 2.10 ms at D=128, 1.08 -> 0.94 ms at D=64 for 9.4 M rows (profiles/r02_dense_transform_hw_trunc_ab.json)
  const char* ht = getenv("IDDGCN_TC_AHI_TRUNC");
  kern<<<grid, TC_THREADS, L::DYN_BYTES, st>>>(ma, mh, ml, C, ldc, M, accumulate, (ht && ht[0] == '0') ? 0 : 1);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

// per-stream scratch for the split weights (hi and lo, up to TC_WS_BATCH matrices of 128 × 128)
constexpr int TC_WS_BATCH = 8;
struct SplitWs { float* p = nullptr; };
std::mutex g_ws_mu;
std::unordered_map<uint64_t, SplitWs> g_ws;

int split_ws(cudaStream_t st, float** out) {
  int dev = 0;
  IDD_CUDA(cudaGetDevice(&dev));
  const uint64_t key = ((uint64_t)dev << 56) ^ (uint64_t)(uintptr_t)st;
  std::lock_guard<std::mutex> lk(g_ws_mu);
  SplitWs& w = g_ws[key];
  if (!w.p) IDD_CUDA(pool_alloc((void**)&w.p, (size_t)TC_WS_BATCH * 2 * 128 * 128 * sizeof(float)));
  *out = w.p;
  return IDDGCN_OK;
}


// =====================================================================================================================
// Weight-gradient product  C[Mo,No] = Σ_rows A[row, :Mo]ᵀ · B[row, :No]   (dW = Sᵀ·dZ, dW_self = Hᵀ·dP): both operands
// stream, the reduction runs over millions of rows, the output is weight-sized (Mo = No ∈ {64, 128}).
//
// Each CTA owns a contiguous range of rows.  Per k-block of 32 rows the producer lands A and B as 32-float-wide column
// atoms ([32 rows × 128 B], SWIZZLE_128B_ATOM_32B) — which is exactly the MN-major canonical layout of the tensor core: both
// operands are consumed "transposed" without any data movement.  The splitter turns both into hi/lo in place, the MMA
// thread issues 4 × 3 tcgen05.mma (128 × No × 8).  Tensor-core accumulation is not round-to-nearest, so the TMEM
// accumulator is only trusted for TN_FLUSH k-blocks (256 rows): the epilogue warps then pull it (tcgen05.ld) and add it
// to a register-resident fp32 total while the MMAs continue into the second TMEM buffer.  At the end every CTA writes
// its 128 × No total to the split-K workspace; dense.cu's fixed-order reduction finishes the sum (deterministic).
// For Mo = 64 the upper half of the A operand is shared memory zeroed once (the tensor core's M is 128).
// =====================================================================================================================
constexpr int TN_FLUSH = 8;            // k-blocks accumulated inside TMEM between two flushes
constexpr int TN_PREFETCH = 4;         // k-blocks of L2 prefetch distance
constexpr int TN_THREADS = 320;        // warps 0-3 epilogue, 4-7 splitter, 8 producer, 9 MMA
constexpr int TN_ATOM = 32 * 128;      // one column atom: 32 rows × 128 bytes

// MN-major tf32 operands have exactly one legal shared-memory layout: 128-byte rows whose 32-BYTE chunks are swizzled
// with (row % 4) (UMMA layout type 1, "128B swizzle, 32B atom" — what TMA's SWIZZLE_128B_ATOM_32B writes).  The swizzle
// atom is 4 rows × 128 bytes: leading byte offset = distance between 32-float column atoms (4096), stride byte offset =
// distance between 4-row groups (512); one MMA (K = 8) spans two groups.
__device__ __forceinline__ uint64_t mnmajor_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(TN_ATOM >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
__host__ __device__ constexpr uint32_t tf32_idesc_mn(int M, int N) { return tf32_idesc(M, N) | (1u << 15) | (1u << 16); }

template <int D_, int STAGES>
struct TnSmem {
  static constexpr int ATOMS = D_ / 32;                                // real column atoms of A and of B
  static constexpr int A_SLAB = 4 * TN_ATOM;                           // the A operand always spans M = 128
  static constexpr int B_SLAB = ATOMS * TN_ATOM;
  static constexpr int STAGE = 2 * A_SLAB + 2 * B_SLAB;                // A hi, A lo, B hi, B lo
  static constexpr int BAR_OFF = STAGES * STAGE;
  static constexpr int NBARS = 3 * STAGES + 4;
  static constexpr int BYTES = BAR_OFF + NBARS * 8 + 16;
  static constexpr int DYN_BYTES = BYTES + 1024;
  static constexpr uint32_t TX = (uint32_t)(2 * ATOMS * TN_ATOM);      // bytes one k-block brings in
};

template <int D_, int STAGES>
__global__ void __launch_bounds__(TN_THREADS, 1)
    tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      float* __restrict__ ws, int64_t rows, int kb_per_cta) {
  using L = TnSmem<D_, STAGES>;
  constexpr int ATOMS = L::ATOMS;
  constexpr int TMEM_COLS = 2 * D_;
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = tc_smem_raw + (base - smem_u32(tc_smem_raw));
  const uint32_t sBar = base + L::BAR_OFF;
  auto bar_full = [&](int s) { return sBar + 8u * s; };
  auto bar_ready = [&](int s) { return sBar + 8u * (STAGES + s); };
  auto bar_empty = [&](int s) { return sBar + 8u * (2 * STAGES + s); };
  auto bar_tfull = [&](int a) { return sBar + 8u * (3 * STAGES + a); };
  auto bar_tempty = [&](int a) { return sBar + 8u * (3 * STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::BAR_OFF + L::NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total_kb = (rows + 31) / 32;
  const int64_t kb0 = (int64_t)blockIdx.x * kb_per_cta;
  const int64_t kb1 = kb0 + kb_per_cta < total_kb ? kb0 + kb_per_cta : total_kb;
  const int nkb = kb1 > kb0 ? (int)(kb1 - kb0) : 0;
  const int ngroups = (nkb + TN_FLUSH - 1) / TN_FLUSH;

  if (ATOMS < 4) {  // rows Mo..127 of the A operand are never loaded: zero them once (hi and lo, every stage)
    for (int s = 0; s < STAGES; ++s)
      for (int half = 0; half < 2; ++half) {
        float4* z = reinterpret_cast<float4*>(base_ptr + (size_t)s * L::STAGE + half * L::A_SLAB + ATOMS * TN_ATOM);
        for (int i = threadIdx.x; i < (4 - ATOMS) * TN_ATOM / 16; i += TN_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    fence_proxy_async();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_ready(s), 128);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull(a), 1);
      mbar_init(bar_tempty(a), 128);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<TMEM_COLS>(smem_u32((const void*)tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1u;
        const int row = (int)((kb0 + i) * 32);
        if (i + TN_PREFETCH < nkb) {
#pragma unroll
          for (int c = 0; c < ATOMS; ++c) {
            tma_prefetch_l2_2d(&map_a, c * 32, row + TN_PREFETCH * 32);
            tma_prefetch_l2_2d(&map_b, c * 32, row + TN_PREFETCH * 32);
          }
        }
        mbar_wait(bar_empty(s), ph ^ 1u);
        mbar_expect_tx(bar_full(s), L::TX);
        const uint32_t st = base + (uint32_t)s * L::STAGE;
#pragma unroll
        for (int c = 0; c < ATOMS; ++c) {
          tma_load_2d(st + c * TN_ATOM, &map_a, bar_full(s), c * 32, row);
          tma_load_2d(st + 2 * L::A_SLAB + c * TN_ATOM, &map_b, bar_full(s), c * 32, row);
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = tf32_idesc_mn(128, D_);
      int i = 0;
      for (int gidx = 0; gidx < ngroups; ++gidx) {
        const int a = gidx & 1;
        const uint32_t aph = (uint32_t)(gidx >> 1) & 1u;
        mbar_wait(bar_tempty(a), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * D_);
        const int gend = (gidx + 1) * TN_FLUSH < nkb ? (gidx + 1) * TN_FLUSH : nkb;
        for (bool first = true; i < gend; ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1u;
          mbar_wait(bar_ready(s), ph);
          tc_fence_after();
          const uint32_t st = base + (uint32_t)s * L::STAGE;
          const uint64_t a_hi = mnmajor_sw128_desc(st), a_lo = mnmajor_sw128_desc(st + L::A_SLAB);
          const uint64_t b_hi = mnmajor_sw128_desc(st + 2 * L::A_SLAB);
          const uint64_t b_lo = mnmajor_sw128_desc(st + 2 * L::A_SLAB + L::B_SLAB);
#pragma unroll
          for (int k = 0; k < 4; ++k) {                       // 8 rows = one 1024-byte group per MMA
            const uint64_t adv = (uint64_t)(k * 64);
            umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, first ? 0u : 1u);
            first = false;
            umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
            umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
          }
          umma_commit(bar_empty(s));
        }
        umma_commit(bar_tfull(a));
      }
    }
  } else if (warp < 4) {
    // ===== epilogue: fp32 running total in registers, one output row per thread =====
    float total[D_];
#pragma unroll
    for (int j = 0; j < D_; ++j) total[j] = 0.f;
    for (int gidx = 0; gidx < ngroups; ++gidx) {
      const int a = gidx & 1;
      const uint32_t aph = (uint32_t)(gidx >> 1) & 1u;
      mbar_wait(bar_tfull(a), aph);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < D_; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * D_ + c0), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) total[c0 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      mbar_arrive(bar_tempty(a));
    }
    const int row = warp * 32 + lane;
    if (row < D_) {
      float* out = ws + ((size_t)blockIdx.x * D_ + row) * D_;
#pragma unroll
      for (int j = 0; j < D_; j += 8) {
        uint32_t v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(total[j + e]);
        st_global_v8(out + j, v);
      }
    }
  } else if (warp < 8) {
    // ===== splitter: both operands of the landed k-block -> hi in place, lo beside =====
    const int t = threadIdx.x - 128;                            // 0..127
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1u;
      mbar_wait(bar_full(s), ph);
      uint8_t* st = base_ptr + (size_t)s * L::STAGE;
#pragma unroll
      for (int op = 0; op < 2; ++op) {
        float4* hi = reinterpret_cast<float4*>(st + (op ? 2 * L::A_SLAB : 0));
        float4* lo = reinterpret_cast<float4*>(st + (op ? 2 * L::A_SLAB + L::B_SLAB : L::A_SLAB));
#pragma unroll
        for (int q = 0; q < ATOMS * TN_ATOM / 16 / 128; ++q) {
          const int o = q * 128 + t;
          const float4 x = hi[o];
          const float4 h = make_float4(tf32_round(x.x), tf32_round(x.y), tf32_round(x.z), tf32_round(x.w));
          hi[o] = h;
          lo[o] = make_float4(tf32_round(x.x - h.x), tf32_round(x.y - h.y), tf32_round(x.z - h.z),
                              tf32_round(x.w - h.w));
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_ready(s));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// fp32 [rows, cols] row-major, ld floats between rows; box = 32 columns × 32 rows (one column atom), 32-byte swizzle atom
int make_map_atom(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld) {
  return make_map(map, ptr, rows, cols, ld, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

template <int D_, int STAGES>
int configure_tn() {
  using L = TnSmem<D_, STAGES>;
  static_assert(L::DYN_BYTES <= 227 * 1024, "shared-memory budget exceeded");
  static bool configured[64] = {};
  int dev = 0;
  IDD_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !configured[dev]) {
    IDD_CUDA(cudaFuncSetAttribute(tc_gemm_tn_kernel<D_, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  L::DYN_BYTES));
    if (dev < 64) configured[dev] = true;
  }
  return IDDGCN_OK;
}

template <int D_, int STAGES>
int launch_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t rows, float* ws, int grid,
              int kb_per_cta, cudaStream_t st) {
  using L = TnSmem<D_, STAGES>;
  IDD_TRY((configure_tn<D_, STAGES>()));
  CUtensorMap ma, mb;
  IDD_TRY(make_map_atom(&ma, A, rows, D_, lda));
  IDD_TRY(make_map_atom(&mb, B, rows, D_, ldb));
  tc_gemm_tn_kernel<D_, STAGES><<<(unsigned)grid, TN_THREADS, L::DYN_BYTES, st>>>(ma, mb, ws, rows, kb_per_cta);
  IDD_LAUNCH_CHECK();
  return IDDGCN_OK;
}

}  // namespace

// Everything that may not happen inside a stream capture (allocation, function attributes): call once per stream
// before capturing a step that contains eligible GEMMs.
int tc_gemm_prepare(cudaStream_t st) {
  float* ws = nullptr;
  IDD_TRY(split_ws(st, &ws));
  IDD_TRY((configure_tc<128, 128, 3>()));
  IDD_TRY((configure_tc<64, 64, 5>()));
  IDD_TRY((configure_tn<128, 3>()));
  IDD_TRY((configure_tn<64, 4>()));
  IDD_CHECK(encode_tiled_fn() != nullptr, IDDGCN_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  return IDDGCN_OK;
}

// number of row ranges (= partial sums written to the workspace) the weight-gradient kernel uses for `rows` rows
int tc_gemm_tn_splits(int64_t rows) {
  const int64_t total_kb = (rows + 31) / 32;
  const int64_t want = total_kb < sm_count() ? total_kb : sm_count();
  const int64_t per = (total_kb + want - 1) / want;
  return (int)((total_kb + per - 1) / per);
}

bool tc_gemm_tn_eligible(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                         int64_t min_rows) {
  if (M != N || (M != 64 && M != 128) || K < min_rows || K >= ((int64_t)1 << 31) - 64) return false;
  if ((lda % 4) != 0 || (ldb % 4) != 0 || lda < M || ldb < N) return false;
  return ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15u) == 0;
}

// ws[s][M][N] (s < tc_gemm_tn_splits(K)) = partial sums of A[K,M]ᵀ·B[K,N]; the caller reduces them in fixed order
int gemm_tn_tf32x3_partials(int64_t M, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* ws,
                            cudaStream_t st) {
  const int grid = tc_gemm_tn_splits(K);
  const int64_t total_kb = (K + 31) / 32;
  const int per = (int)((total_kb + grid - 1) / grid);
  if (M == 128) return launch_tn<128, 3>(A, lda, B, ldb, K, ws, grid, per, st);
  return launch_tn<64, 4>(A, lda, B, ldb, K, ws, grid, per, st);
}

bool tc_gemm_enabled() {
  const char* e = getenv("IDDGCN_GEMM_TF32X3");   // read per call so tests can flip it
  return !(e && e[0] == '0');
}

bool tc_gemm_eligible(bool transA, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                      int64_t ldb, const float* C, int64_t ldc, int64_t min_rows) {
  if (transA || N != K || (N != 64 && N != 128) || M < min_rows || M >= ((int64_t)1 << 31) - TC_BLOCK_M) return false;
  // TMA wants 16-byte aligned rows of A; the epilogue moves C with 256-bit accesses, so C rows must be 32-byte aligned
  if (ldb != N || (lda % 4) != 0 || (ldc % 8) != 0 || lda < K || ldc < N) return false;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B);
  return (bits & 15u) == 0 && (reinterpret_cast<uintptr_t>(C) & 31u) == 0;
}

// C[b] (+)= A[b] · op(B[b]) with op(B) = B (B is [K,N] row-major) when !transB, Bᵀ (B is [N,K]) when transB.
int gemm_tf32x3(bool transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, int64_t strideA,
                const float* B, int64_t strideB, float* C, int64_t ldc, int64_t strideC, int batch, bool accumulate,
                cudaStream_t st) {
  IDD_CHECK(N == K && (N == 64 || N == 128), IDDGCN_EINVAL, "tf32x3 GEMM needs N == K in {64, 128}");
  float* ws = nullptr;
  IDD_TRY(split_ws(st, &ws));
  const int64_t W = N * K;
  for (int b0 = 0; b0 < batch; b0 += TC_WS_BATCH) {
    const int nb = batch - b0 < TC_WS_BATCH ? batch - b0 : TC_WS_BATCH;
    float* hi = ws;
    float* lo = ws + (size_t)TC_WS_BATCH * 128 * 128;
    // the kernel wants Bk = [N, K] with K contiguous: that is B itself when transB, Bᵀ otherwise
    dim3 sg((unsigned)((W + 255) / 256), (unsigned)nb);
    split_tf32_kernel<<<sg, 256, 0, st>>>(B + (int64_t)b0 * strideB, hi, lo, (int)(transB ? N : K), (int)(transB ? K : N),
                                          strideB, transB ? 0 : 1);
    IDD_LAUNCH_CHECK();
    for (int b = 0; b < nb; ++b) {
      const float* Ab = A + (int64_t)(b0 + b) * strideA;
      float* Cb = C + (int64_t)(b0 + b) * strideC;
      const int acc = accumulate ? 1 : 0;
      if (N == 128)
        IDD_TRY((launch_tc<128, 128, 3>(Ab, lda, hi + b * W, lo + b * W, Cb, ldc, M, acc, st)));
      else
        IDD_TRY((launch_tc<64, 64, 5>(Ab, lda, hi + b * W, lo + b * W, Cb, ldc, M, acc, st)));
    }
  }
  return IDDGCN_OK;
}

}  // namespace iddgcn
