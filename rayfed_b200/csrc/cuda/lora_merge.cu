// K3 (GEMM part): exact federated LoRA aggregation on the 5th-gen tensor cores.
//
// FedAvg of LoRA adapters is not "average A, average B": the quantity to
// average is the product, dW = sum_i w_i * B_i A_i.  Written as one GEMM,
//
//     dW[M x N] = [B_1 ... B_P] (M x P*r)  *  [A_1 ; ... ; A_P] (P*r x N),
//
// the K dimension is *spread over the P parties* (SURVEY §2.D K3).  This kernel
// fuses the all-gather with the GEMM: every CTA pulls its operand tiles straight
// from the peers' adapter buffers over NVLink (P2P 16-byte loads from the
// symmetric heap) into the canonical K-major SWIZZLE_128B shared-memory layout,
// one elected thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M=128,
// N=128, K=16 per instruction) accumulating in TMEM, and the epilogue reads the
// accumulator back with tcgen05.ld, applies  W <- W_base + scale * dW  and
// writes bf16 — or re-quantises to block-scaled e4m3 (UE8M0 per 32 along the
// row; one thread holds exactly one 32-element block of one row after
// tcgen05.ld.32x32b.x32, so the block scale needs no shuffles).
//
// Storage convention (ours, chosen for the hardware): B_i is [M, r] row-major
// and A_i is stored transposed, At_i = [N, r] row-major, so both operands are
// K-major and a party's slice of K is r contiguous elements.
#include <cuda.h>

#include "kernels.h"

namespace fedk {

constexpr int kLoraBM = 128;
constexpr int kLoraBN = 128;
constexpr int kLoraThreads = 128;

__device__ __forceinline__ uint32_t lm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1").
//   start address >> 4 | LBO(=1, unused for swizzled K-major) | SBO = 1024 B (8 rows x 128 B) | layout 2
__device__ __forceinline__ uint64_t lm_make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128.
__device__ __forceinline__ uint32_t lm_instr_desc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kLoraBN >> 3) << 17) | ((uint32_t)(kLoraBM >> 4) << 24);
}

__device__ __forceinline__ void lm_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void lm_tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Gather one operand tile (128 rows x K) from the P parties into swizzled smem.
//   src_i(row, c) : 16-byte chunk c of row `row` of party i's [rows_total, r] matrix
//   smem layout   : [K/64 atoms][128 rows][128 B], 16 B chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ void lm_gather_tile(unsigned char* smem, const PeerPtrs& src, int n_parties, int r,
                                               size_t row0, size_t rows_total, const float* party_w) {
  const int chunks_per_party = r >> 3;  // 16 B = 8 bf16
  const int per_party = kLoraBM * chunks_per_party;
  const int total = n_parties * per_party;
  constexpr int kBatch = 8;  // 8 independent 16 B requests in flight per thread before the first smem store
  for (int base = threadIdx.x; base < total; base += kLoraThreads * kBatch) {
    uint4 v[kBatch];
    int dst_off[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int idx = base + u * kLoraThreads;
      v[u] = make_uint4(0, 0, 0, 0);
      dst_off[u] = -1;
      if (idx < total) {
        // consecutive threads -> consecutive 16 B chunks of one party's contiguous [128 x r] block
        const int i = idx / per_party;
        const int rem = idx - i * per_party;
        const int row = rem / chunks_per_party;
        const int c = rem - row * chunks_per_party;
        if (row0 + row < rows_total)
          v[u] = ld_v4(reinterpret_cast<const char*>(src.p[i]) + ((row0 + row) * (size_t)r + (size_t)c * 8) * 2);
        const int k0 = i * r + c * 8;
        dst_off[u] = (k0 >> 6) * (kLoraBM * 128) + row * 128 + ((((k0 & 63) >> 3) ^ (row & 7)) << 4);
        if (party_w != nullptr && party_w[i] != 1.0f) {
          const float w = party_w[i];
          v[u].x = pack_bf16x2(bf16_lo(v[u].x) * w, bf16_hi(v[u].x) * w);
          v[u].y = pack_bf16x2(bf16_lo(v[u].y) * w, bf16_hi(v[u].y) * w);
          v[u].z = pack_bf16x2(bf16_lo(v[u].z) * w, bf16_hi(v[u].z) * w);
          v[u].w = pack_bf16x2(bf16_lo(v[u].w) * w, bf16_hi(v[u].w) * w);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u)
      if (dst_off[u] >= 0) *reinterpret_cast<uint4*>(smem + dst_off[u]) = v[u];
  }
}

__global__ void __launch_bounds__(kLoraThreads) fed_lora_merge_kernel(const __grid_constant__ LoraMergeArgs a) {
  extern __shared__ unsigned char lm_smem[];
  __shared__ __align__(8) uint64_t mma_done;
  __shared__ uint32_t tmem_base_slot;

  const int K = a.n_parties * a.r;
  const int atoms = (K + 63) >> 6;
  // SWIZZLE_128B atoms need a 1024 B aligned base; dynamic smem starts after the statics.
  const uint32_t raw = lm_smem_u32(lm_smem);
  unsigned char* tile0 = lm_smem + (((raw + 1023u) & ~1023u) - raw);
  unsigned char* sA = tile0;                                      // B_cat tile: rows of W  (M operand)
  unsigned char* sB = tile0 + (size_t)atoms * (kLoraBM * 128);    // At_cat tile: cols of W (N operand)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const size_t m0 = (size_t)blockIdx.y * kLoraBM;
  const size_t n0 = (size_t)blockIdx.x * kLoraBN;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     lm_smem_u32(&tmem_base_slot)),
                 "n"(kLoraBN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lm_smem_u32(&mma_done)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- all-gather of the operand tiles, fused: peers' memory -> my smem ------------------
  lm_gather_tile(sA, a.b, a.n_parties, a.r, m0, a.m, a.w);
  lm_gather_tile(sB, a.at, a.n_parties, a.r, n0, a.n, nullptr);
  // generic-proxy smem writes must be visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  // ---- one thread drives the tensor core ---------------------------------------------------
  if (threadIdx.x == 0) {
    const uint32_t idesc = lm_instr_desc();
    const uint32_t a_base = lm_smem_u32(sA), b_base = lm_smem_u32(sB);
    const int ksteps = K >> 4;
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint32_t off = (uint32_t)(ks >> 2) * (kLoraBM * 128) + (uint32_t)(ks & 3) * 32;  // atom + 32 B per K=16
      lm_mma(tmem_d, lm_make_desc(a_base + off), lm_make_desc(b_base + off), idesc, ks > 0 ? 1u : 0u);
    }
    // commit: mbarrier arrives when every MMA above has finished writing TMEM
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     lm_smem_u32(&mma_done))
                 : "memory");
  }

  // ---- epilogue: TMEM -> registers -> W = W_base + scale * dW -> global ---------------------
  {
    uint32_t done;
    do {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(lm_smem_u32(&mma_done)), "r"(0u)
          : "memory");
    } while (!done);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const size_t row = m0 + (size_t)warp * 32 + lane;  // warp w owns TMEM lanes 32w..32w+31
  const float scale = a.scale;
#pragma unroll 1
  for (int c = 0; c < kLoraBN / 32; ++c) {
    const size_t col = n0 + (size_t)c * 32;
    const bool in_range = row < a.m && col < a.n;
    const size_t e = row * a.n + col;
    uint4 wpre[4];
    const bool bf16_base = in_range && a.w_base && !a.base_fp8;
    if (bf16_base) {
      const uint4* wb = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.w_base) + e);
#pragma unroll
      for (int q = 0; q < 4; ++q) wpre[q] = ld_stream_v4(wb + q);
    }
    uint32_t v[32];
    lm_tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
    if (in_range) {
      float acc[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[k] = __uint_as_float(v[k]) * scale;
      if (a.w_base) {
        if (!a.base_fp8) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 x = wpre[q];
            acc[q * 8 + 0] += bf16_lo(x.x); acc[q * 8 + 1] += bf16_hi(x.x);
            acc[q * 8 + 2] += bf16_lo(x.y); acc[q * 8 + 3] += bf16_hi(x.y);
            acc[q * 8 + 4] += bf16_lo(x.z); acc[q * 8 + 5] += bf16_hi(x.z);
            acc[q * 8 + 6] += bf16_lo(x.w); acc[q * 8 + 7] += bf16_hi(x.w);
          }
        } else {
          const unsigned char* wq = reinterpret_cast<const unsigned char*>(a.w_base) + e;
          const float s = __uint_as_float((uint32_t)a.w_base_scales[e >> 5] << 23);
          const uint4* w4 = reinterpret_cast<const uint4*>(wq);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint4 x = w4[q];
            uint32_t ws[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(ws[j] & 0xffffu), __NV_E4M3);
              __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(ws[j] >> 16), __NV_E4M3);
              float2 f0 = __half22float2(*reinterpret_cast<__half2*>(&lo));
              float2 f1 = __half22float2(*reinterpret_cast<__half2*>(&hi));
              acc[q * 16 + j * 4 + 0] += f0.x * s; acc[q * 16 + j * 4 + 1] += f0.y * s;
              acc[q * 16 + j * 4 + 2] += f1.x * s; acc[q * 16 + j * 4 + 3] += f1.y * s;
            }
          }
        }
      }
      if (!a.out_fp8) {
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + e);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 y;
          y.x = pack_bf16x2(acc[q * 8 + 0], acc[q * 8 + 1]);
          y.y = pack_bf16x2(acc[q * 8 + 2], acc[q * 8 + 3]);
          y.z = pack_bf16x2(acc[q * 8 + 4], acc[q * 8 + 5]);
          y.w = pack_bf16x2(acc[q * 8 + 6], acc[q * 8 + 7]);
          o[q] = y;
        }
      } else {
        float amax = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) amax = fmaxf(amax, fabsf(acc[k]));
        float rr = amax * (1.0f / 448.0f);
        uint32_t bits = __float_as_uint(rr);
        uint32_t eb = (bits >> 23) & 0xffu;
        if (bits & 0x7fffffu) eb += 1;
        eb = eb < 1u ? 1u : (eb > 253u ? 253u : eb);
        const float inv = __uint_as_float((254u - eb) << 23);
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(a.out) + e);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t lo = __nv_cvt_float2_to_fp8x2(
                make_float2(acc[q * 16 + j * 4 + 0] * inv, acc[q * 16 + j * 4 + 1] * inv), __NV_SATFINITE, __NV_E4M3);
            uint32_t hi = __nv_cvt_float2_to_fp8x2(
                make_float2(acc[q * 16 + j * 4 + 2] * inv, acc[q * 16 + j * 4 + 3] * inv), __NV_SATFINITE, __NV_E4M3);
            w[j] = lo | (hi << 16);
          }
          o[q] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        a.out_scales[e >> 5] = (unsigned char)eb;
      }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kLoraBN) : "memory");
  }
}


// ---------------------------------------------------------------------------
// TMA-tiled variant (bf16 base/out): the W_base tile is fetched by the TMA engine
// (cp.async.bulk.tensor.2d, SWIZZLE_128B) while the SMs gather the operands from the
// peers, the epilogue updates the tile in shared memory (row-per-thread after
// tcgen05.ld, conflict-free thanks to the hardware swizzle) and one thread writes
// it back with a TMA store.  Global traffic is fully coalesced both ways.
struct LoraTmaMaps {
  CUtensorMap w;
  CUtensorMap out;
};

__device__ __forceinline__ void lm_tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(lm_smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(lm_smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void lm_tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
               "r"(c1), "r"(lm_smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void lm_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(lm_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

__global__ void __launch_bounds__(kLoraThreads)
fed_lora_merge_tma_kernel(const __grid_constant__ LoraMergeArgs a, const __grid_constant__ LoraTmaMaps maps) {
  extern __shared__ unsigned char lm_smem[];
  __shared__ __align__(8) uint64_t mma_done;
  __shared__ __align__(8) uint64_t w_full;
  __shared__ uint32_t tmem_base_slot;

  const int K = a.n_parties * a.r;
  const int atoms = (K + 63) >> 6;
  const uint32_t raw = lm_smem_u32(lm_smem);
  unsigned char* tile0 = lm_smem + (((raw + 1023u) & ~1023u) - raw);
  unsigned char* sW = tile0;                                   // [2 atoms][128 rows][128 B] output tile
  unsigned char* sA = tile0 + 2 * (kLoraBM * 128);
  unsigned char* sB = sA + (size_t)atoms * (kLoraBM * 128);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const size_t m0 = (size_t)blockIdx.y * kLoraBM;
  const size_t n0 = (size_t)blockIdx.x * kLoraBN;
  const bool has_base = a.w_base != nullptr;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     lm_smem_u32(&tmem_base_slot)),
                 "n"(kLoraBN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lm_smem_u32(&mma_done)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lm_smem_u32(&w_full)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (has_base) {
      // W_base tile by TMA: overlaps the operand gather below
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lm_smem_u32(&w_full)),
                   "r"(2u * kLoraBM * 128u)
                   : "memory");
      lm_tma_load_2d(sW, &maps.w, (int)n0, (int)m0, &w_full);
      lm_tma_load_2d(sW + kLoraBM * 128, &maps.w, (int)n0 + 64, (int)m0, &w_full);
    }
  }

  lm_gather_tile(sA, a.b, a.n_parties, a.r, m0, a.m, a.w);
  lm_gather_tile(sB, a.at, a.n_parties, a.r, n0, a.n, nullptr);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (threadIdx.x == 0) {
    const uint32_t idesc = lm_instr_desc();
    const uint32_t a_base = lm_smem_u32(sA), b_base = lm_smem_u32(sB);
    const int ksteps = K >> 4;
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint32_t off = (uint32_t)(ks >> 2) * (kLoraBM * 128) + (uint32_t)(ks & 3) * 32;
      lm_mma(tmem_d, lm_make_desc(a_base + off), lm_make_desc(b_base + off), idesc, ks > 0 ? 1u : 0u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     lm_smem_u32(&mma_done))
                 : "memory");
  }

  lm_mbar_wait(&mma_done, 0);
  if (has_base) lm_mbar_wait(&w_full, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const int r = warp * 32 + lane;  // tile row owned by this thread (TMEM lane)
  const float scale = a.scale;
#pragma unroll 1
  for (int c = 0; c < kLoraBN / 32; ++c) {
    uint32_t v[32];
    lm_tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
    unsigned char* row = sW + (size_t)(c >> 1) * (kLoraBM * 128) + (size_t)r * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = (c & 1) * 4 + j;  // logical 16 B chunk within the 128 B atom row
      uint4* slot = reinterpret_cast<uint4*>(row + ((q ^ (r & 7)) << 4));
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[j * 8 + k]) * scale;
      if (has_base) {
        const uint4 x = *slot;
        f[0] += bf16_lo(x.x); f[1] += bf16_hi(x.x); f[2] += bf16_lo(x.y); f[3] += bf16_hi(x.y);
        f[4] += bf16_lo(x.z); f[5] += bf16_hi(x.z); f[6] += bf16_lo(x.w); f[7] += bf16_hi(x.w);
      }
      uint4 y;
      y.x = pack_bf16x2(f[0], f[1]);
      y.y = pack_bf16x2(f[2], f[3]);
      y.z = pack_bf16x2(f[4], f[5]);
      y.w = pack_bf16x2(f[6], f[7]);
      *slot = y;
    }
  }
  // generic-proxy writes of the tile -> visible to the TMA engine
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    lm_tma_store_2d(&maps.out, (int)n0, (int)m0, sW);
    lm_tma_store_2d(&maps.out, (int)n0 + 64, (int)m0, sW + kLoraBM * 128);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem may be released after this
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kLoraBN) : "memory");
  }
}

// ---------------------------------------------------------------------------
// Staged, persistent, warp-specialised variant (default for bf16 base/out, K <= 128).
//
// The direct kernels above pull every operand tile from the peers once per *output
// tile*: a 4096 x 4096 merge moves 1024 x 64 KB = 64 MB of adapter data through the
// LSUs (56 MB of it over NVLink at 8 parties) to produce 32 MB of output.  Here the
// all-gather is hoisted into a tiny first phase — every peer slice crosses NVLink
// exactly once (2 MB) into K-major concatenated operands [M, Kpad] / [N, Kpad] that
// stay L2-resident — and the GEMM becomes a persistent TMA -> tcgen05 -> TMA pipeline:
//
//   warp 0      producer : TMA loads {W tile, A tile, B tile} into a 2-stage smem ring
//   warp 1      MMA      : tcgen05.mma into one of two 128-column TMEM accumulators
//   warps 2..9  epilogue : tcgen05.ld -> W += scale * dW in the swizzled smem tile
//                          -> TMA store; the next tile's loads and MMAs run underneath
//
// All hand-offs are mbarriers: full[s] (TMA bytes landed), tmem_full[s] (tcgen05.commit),
// tmem_empty[s] (8 epilogue warps drained the accumulator), empty[s] (TMA store finished
// reading the stage).
constexpr int kStgStages = 2;
constexpr int kStgThreads = 320;        // 10 warps
constexpr int kStgEpiThreads = 256;     // warps 2..9
constexpr int kAtomBytes = kLoraBM * 128;  // one 64-column SWIZZLE_128B atom of a 128-row tile

struct LoraStagedMaps {
  CUtensorMap w;
  CUtensorMap out;
  CUtensorMap bcat;   // [M, Kpad]
  CUtensorMap atcat;  // [N, Kpad]
};

__global__ void __launch_bounds__(256) fed_lora_gather_kernel(const __grid_constant__ LoraMergeArgs a,
                                                              __nv_bfloat16* __restrict__ bcat,
                                                              __nv_bfloat16* __restrict__ atcat, int kpad) {
  const int cpp = a.r >> 3;  // 16 B chunks per party row
  const size_t nb = (size_t)a.n_parties * a.m * cpp;
  const size_t na = (size_t)a.n_parties * a.n * cpp;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nb + na; idx += stride) {
    const bool is_b = idx < nb;
    const size_t j = is_b ? idx : idx - nb;
    const size_t rows = is_b ? a.m : a.n;
    // (party, row, chunk): consecutive threads read consecutive bytes of ONE peer
    const int i = (int)(j / (rows * cpp));
    const size_t rem = j - (size_t)i * rows * cpp;
    const size_t row = rem / cpp;
    const int c = (int)(rem - row * cpp);
    const char* src = reinterpret_cast<const char*>(is_b ? a.b.p[i] : a.at.p[i]);
    uint4 v = ld_v4(src + (row * (size_t)a.r + (size_t)c * 8) * 2);
    if (is_b && a.w[i] != 1.0f) {
      const float w = a.w[i];
      v.x = pack_bf16x2(bf16_lo(v.x) * w, bf16_hi(v.x) * w);
      v.y = pack_bf16x2(bf16_lo(v.y) * w, bf16_hi(v.y) * w);
      v.z = pack_bf16x2(bf16_lo(v.z) * w, bf16_hi(v.z) * w);
      v.w = pack_bf16x2(bf16_lo(v.w) * w, bf16_hi(v.w) * w);
    }
    __nv_bfloat16* dst = (is_b ? bcat : atcat) + row * (size_t)kpad + (size_t)i * a.r + (size_t)c * 8;
    *reinterpret_cast<uint4*>(dst) = v;  // default caching: the GEMM phase reads this from L2
  }
}

__device__ __forceinline__ void lm_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lm_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void lm_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(lm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void lm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lm_smem_u32(bar)), "r"(bytes)
               : "memory");
}

__global__ void __launch_bounds__(kStgThreads, 1)
fed_lora_gemm_kernel(const __grid_constant__ LoraMergeArgs a, const __grid_constant__ LoraStagedMaps maps,
                     int atoms, int tiles_n, int n_tiles) {
  extern __shared__ unsigned char lm_smem[];
  __shared__ __align__(8) uint64_t full_bar[kStgStages];
  __shared__ __align__(8) uint64_t empty_bar[kStgStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[kStgStages];
  __shared__ __align__(8) uint64_t tmem_empty_bar[kStgStages];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t raw = lm_smem_u32(lm_smem);
  unsigned char* ring = lm_smem + (((raw + 1023u) & ~1023u) - raw);
  const int stage_bytes = (2 + 2 * atoms) * kAtomBytes;  // [W: 2 atoms][A: atoms][B: atoms]
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool has_base = a.w_base != nullptr;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStgStages; ++s) {
      lm_mbar_init(&full_bar[s], 1);
      lm_mbar_init(&empty_bar[s], 1);
      lm_mbar_init(&tmem_full_bar[s], 1);
      lm_mbar_init(&tmem_empty_bar[s], kStgEpiThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     lm_smem_u32(&tmem_base_slot)),
                 "n"(kStgStages * kLoraBN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ---------------- producer: TMA loads ----------------------------------------------------
    if (lane == 0) {
      const uint32_t tx = (uint32_t)((has_base ? 2 : 0) + 2 * atoms) * kAtomBytes;
      for (int it = 0; it < my_tiles; ++it) {
        const int t = (int)blockIdx.x + it * (int)gridDim.x;
        const int m0 = (t / tiles_n) * kLoraBM, n0 = (t % tiles_n) * kLoraBN;
        const int s = it & 1;
        const uint32_t ph = (uint32_t)(it >> 1) & 1u;
        lm_mbar_wait(&empty_bar[s], ph ^ 1u);  // first pass over the ring falls through
        unsigned char* st = ring + (size_t)s * stage_bytes;
        lm_mbar_expect_tx(&full_bar[s], tx);
        if (has_base) {
          lm_tma_load_2d(st, &maps.w, n0, m0, &full_bar[s]);
          lm_tma_load_2d(st + kAtomBytes, &maps.w, n0 + 64, m0, &full_bar[s]);
        }
        unsigned char* sA = st + 2 * kAtomBytes;
        unsigned char* sB = sA + (size_t)atoms * kAtomBytes;
        for (int k = 0; k < atoms; ++k) {
          lm_tma_load_2d(sA + (size_t)k * kAtomBytes, &maps.bcat, k * 64, m0, &full_bar[s]);
          lm_tma_load_2d(sB + (size_t)k * kAtomBytes, &maps.atcat, k * 64, n0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer -----------------------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = lm_instr_desc();
      const int ksteps = (a.n_parties * a.r) >> 4;
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it & 1;
        const uint32_t ph = (uint32_t)(it >> 1) & 1u;
        lm_mbar_wait(&full_bar[s], ph);
        lm_mbar_wait(&tmem_empty_bar[s], ph ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        unsigned char* st = ring + (size_t)s * stage_bytes;
        const uint32_t a_base = lm_smem_u32(st + 2 * kAtomBytes);
        const uint32_t b_base = a_base + (uint32_t)atoms * kAtomBytes;
        const uint32_t tmem_d = tmem_base + (uint32_t)(s * kLoraBN);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * kAtomBytes + (uint32_t)(ks & 3) * 32;
          lm_mma(tmem_d, lm_make_desc(a_base + off), lm_make_desc(b_base + off), idesc, ks > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         lm_smem_u32(&tmem_full_bar[s]))
                     : "memory");
      }
    }
  } else {
    // ---------------- epilogue: 8 warps; warp%4 selects the TMEM lane quarter, two warps per quarter
    //                  split the 128 accumulator columns (= the two 64-column smem atoms) ----------
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const float scale = a.scale;
    const bool storer = threadIdx.x == 64;  // first epilogue thread
    for (int it = 0; it < my_tiles; ++it) {
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const int m0 = (t / tiles_n) * kLoraBM, n0 = (t % tiles_n) * kLoraBN;
      const int s = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      unsigned char* sW = ring + (size_t)s * stage_bytes;
      lm_mbar_wait(&full_bar[s], ph);       // W tile landed (and the stage is ours)
      lm_mbar_wait(&tmem_full_bar[s], ph);  // accumulator complete
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      unsigned char* row = sW + (size_t)h * kAtomBytes + (size_t)r * 128;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        lm_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kLoraBN + h * 64 + c * 32), v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = c * 4 + j;
          uint4* slot = reinterpret_cast<uint4*>(row + ((chunk ^ (r & 7)) << 4));
          float f[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[j * 8 + k]) * scale;
          if (has_base) {
            const uint4 x = *slot;
            f[0] += bf16_lo(x.x); f[1] += bf16_hi(x.x); f[2] += bf16_lo(x.y); f[3] += bf16_hi(x.y);
            f[4] += bf16_lo(x.z); f[5] += bf16_hi(x.z); f[6] += bf16_lo(x.w); f[7] += bf16_hi(x.w);
          }
          uint4 y;
          y.x = pack_bf16x2(f[0], f[1]);
          y.y = pack_bf16x2(f[2], f[3]);
          y.z = pack_bf16x2(f[4], f[5]);
          y.w = pack_bf16x2(f[6], f[7]);
          *slot = y;
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) lm_mbar_arrive(&tmem_empty_bar[s]);
      // tile complete in smem -> TMA store
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("barrier.sync 1, %0;" ::"n"(kStgEpiThreads) : "memory");
      if (storer) {
        lm_tma_store_2d(&maps.out, n0, m0, sW);
        lm_tma_store_2d(&maps.out, n0 + 64, m0, sW + kAtomBytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem of this stage is reusable
        lm_mbar_arrive(&empty_bar[s]);
      }
    }
    if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kStgStages * kLoraBN)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------
// "staged3": same warp roles as fed_lora_gemm_kernel, but the W tiles and the operand tiles
// live in SEPARATE rings with separate life times.  An operand stage is free the moment its
// MMAs have retired (tcgen05.commit -> op_empty), long before the epilogue of that tile ran;
// a W stage is free when its TMA store has finished reading it.  With 2 operand stages and 3 W
// stages the producer runs two to three tiles ahead, so the per-tile critical path shrinks from
// max(epilogue, load latency + MMA) towards the epilogue alone.  Shared memory at K = 128:
// 3 x 32 KB (W) + 2 x 64 KB (A|B) = 224 KB of the 227 KB a CTA may have.
// OPT-IN (impl="staged3"): written after this round's GPU budget was spent — not yet executed
// on hardware; fed_lora_gemm_kernel above stays the default until this one has been measured.
constexpr int kWStages3 = 3;
constexpr int kOpStages3 = 2;

__global__ void __launch_bounds__(kStgThreads, 1)
fed_lora_gemm3_kernel(const __grid_constant__ LoraMergeArgs a, const __grid_constant__ LoraStagedMaps maps,
                      int atoms, int tiles_n, int n_tiles) {
  extern __shared__ unsigned char lm_smem[];
  __shared__ __align__(8) uint64_t w_full[kWStages3];
  __shared__ __align__(8) uint64_t w_empty[kWStages3];
  __shared__ __align__(8) uint64_t op_full[kOpStages3];
  __shared__ __align__(8) uint64_t op_empty[kOpStages3];
  __shared__ __align__(8) uint64_t tmem_full_bar[kStgStages];
  __shared__ __align__(8) uint64_t tmem_empty_bar[kStgStages];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t raw = lm_smem_u32(lm_smem);
  unsigned char* w_ring = lm_smem + (((raw + 1023u) & ~1023u) - raw);  // [kWStages3][2 atoms]
  unsigned char* op_ring = w_ring + (size_t)kWStages3 * 2 * kAtomBytes;  // [kOpStages3][A: atoms | B: atoms]
  const int op_stage_bytes = 2 * atoms * kAtomBytes;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool has_base = a.w_base != nullptr;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kWStages3; ++s) {
      lm_mbar_init(&w_full[s], 1);
      lm_mbar_init(&w_empty[s], 1);
    }
#pragma unroll
    for (int s = 0; s < kOpStages3; ++s) {
      lm_mbar_init(&op_full[s], 1);
      lm_mbar_init(&op_empty[s], 1);
    }
#pragma unroll
    for (int s = 0; s < kStgStages; ++s) {
      lm_mbar_init(&tmem_full_bar[s], 1);
      lm_mbar_init(&tmem_empty_bar[s], kStgEpiThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     lm_smem_u32(&tmem_base_slot)),
                 "n"(kStgStages * kLoraBN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ---------------- producer -----------------------------------------------------------------
    if (lane == 0) {
      const uint32_t op_tx = (uint32_t)(2 * atoms) * kAtomBytes;
      for (int it = 0; it < my_tiles; ++it) {
        const int t = (int)blockIdx.x + it * (int)gridDim.x;
        const int m0 = (t / tiles_n) * kLoraBM, n0 = (t % tiles_n) * kLoraBN;
        // operands first: they gate the MMA, which is the head of the per-tile dependency chain
        const int os = it % kOpStages3;
        const uint32_t oph = (uint32_t)(it / kOpStages3) & 1u;
        lm_mbar_wait(&op_empty[os], oph ^ 1u);
        unsigned char* sA = op_ring + (size_t)os * op_stage_bytes;
        unsigned char* sB = sA + (size_t)atoms * kAtomBytes;
        lm_mbar_expect_tx(&op_full[os], op_tx);
        for (int k = 0; k < atoms; ++k) {
          lm_tma_load_2d(sA + (size_t)k * kAtomBytes, &maps.bcat, k * 64, m0, &op_full[os]);
          lm_tma_load_2d(sB + (size_t)k * kAtomBytes, &maps.atcat, k * 64, n0, &op_full[os]);
        }
        const int ws = it % kWStages3;
        const uint32_t wph = (uint32_t)(it / kWStages3) & 1u;
        lm_mbar_wait(&w_empty[ws], wph ^ 1u);
        unsigned char* sW = w_ring + (size_t)ws * 2 * kAtomBytes;
        if (has_base) {
          lm_mbar_expect_tx(&w_full[ws], 2u * kAtomBytes);
          lm_tma_load_2d(sW, &maps.w, n0, m0, &w_full[ws]);
          lm_tma_load_2d(sW + kAtomBytes, &maps.w, n0 + 64, m0, &w_full[ws]);
        } else {
          lm_mbar_arrive(&w_full[ws]);  // nothing to load: the stage is simply handed over
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer -----------------------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = lm_instr_desc();
      const int ksteps = (a.n_parties * a.r) >> 4;
      for (int it = 0; it < my_tiles; ++it) {
        const int os = it % kOpStages3;
        const uint32_t oph = (uint32_t)(it / kOpStages3) & 1u;
        const int ts = it & 1;
        const uint32_t tph = (uint32_t)(it >> 1) & 1u;
        lm_mbar_wait(&op_full[os], oph);
        lm_mbar_wait(&tmem_empty_bar[ts], tph ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_base = lm_smem_u32(op_ring + (size_t)os * op_stage_bytes);
        const uint32_t b_base = a_base + (uint32_t)atoms * kAtomBytes;
        const uint32_t tmem_d = tmem_base + (uint32_t)(ts * kLoraBN);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * kAtomBytes + (uint32_t)(ks & 3) * 32;
          lm_mma(tmem_d, lm_make_desc(a_base + off), lm_make_desc(b_base + off), idesc, ks > 0 ? 1u : 0u);
        }
        // both commits track every MMA issued above: accumulator ready / operand stage reusable
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         lm_smem_u32(&tmem_full_bar[ts]))
                     : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         lm_smem_u32(&op_empty[os]))
                     : "memory");
      }
    }
  } else {
    // ---------------- epilogue (identical tile math to fed_lora_gemm_kernel) -----------------------
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const float scale = a.scale;
    const bool storer = threadIdx.x == 64;
    for (int it = 0; it < my_tiles; ++it) {
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const int m0 = (t / tiles_n) * kLoraBM, n0 = (t % tiles_n) * kLoraBN;
      const int ws = it % kWStages3;
      const uint32_t wph = (uint32_t)(it / kWStages3) & 1u;
      const int ts = it & 1;
      const uint32_t tph = (uint32_t)(it >> 1) & 1u;
      unsigned char* sW = w_ring + (size_t)ws * 2 * kAtomBytes;
      lm_mbar_wait(&w_full[ws], wph);
      lm_mbar_wait(&tmem_full_bar[ts], tph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      unsigned char* row = sW + (size_t)h * kAtomBytes + (size_t)r * 128;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        lm_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ts * kLoraBN + h * 64 + c * 32), v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = c * 4 + j;
          uint4* slot = reinterpret_cast<uint4*>(row + ((chunk ^ (r & 7)) << 4));
          float f[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[j * 8 + k]) * scale;
          if (has_base) {
            const uint4 x = *slot;
            f[0] += bf16_lo(x.x); f[1] += bf16_hi(x.x); f[2] += bf16_lo(x.y); f[3] += bf16_hi(x.y);
            f[4] += bf16_lo(x.z); f[5] += bf16_hi(x.z); f[6] += bf16_lo(x.w); f[7] += bf16_hi(x.w);
          }
          uint4 y;
          y.x = pack_bf16x2(f[0], f[1]);
          y.y = pack_bf16x2(f[2], f[3]);
          y.z = pack_bf16x2(f[4], f[5]);
          y.w = pack_bf16x2(f[6], f[7]);
          *slot = y;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) lm_mbar_arrive(&tmem_empty_bar[ts]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("barrier.sync 1, %0;" ::"n"(kStgEpiThreads) : "memory");
      if (storer) {
        lm_tma_store_2d(&maps.out, n0, m0, sW);
        lm_tma_store_2d(&maps.out, n0 + 64, m0, sW + kAtomBytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // release the stage as early as possible (the TMA only has to have READ it): the
        // sooner W(i+3) can be issued, the more of epilogue(i+1), (i+2) its latency hides under
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        lm_mbar_arrive(&w_empty[ws]);
      }
    }
    if (storer) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kStgStages * kLoraBN)
                 : "memory");
  }
}


#define FEDK_PRELOAD(fn)                                                        \
  do {                                                                          \
    cudaFuncAttributes at_;                                                     \
    cudaError_t e_ = cudaFuncGetAttributes(&at_, fn);                           \
    if (e_ != cudaSuccess) return e_;                                           \
  } while (0)

extern "C" cudaError_t fedk_preload_lora() {
  FEDK_PRELOAD(fed_lora_merge_kernel);
  FEDK_PRELOAD(fed_lora_merge_tma_kernel);
  FEDK_PRELOAD(fed_lora_gather_kernel);
  FEDK_PRELOAD(fed_lora_gemm_kernel);
  FEDK_PRELOAD(fed_lora_gemm3_kernel);
  return cudaSuccess;
}

extern "C" cudaError_t fedk_launch_lora_merge_staged(const LoraMergeArgs* a, const void* maps, void* bcat, void* atcat,
                                                     int kpad, int sm_count, int deep, cudaStream_t stream) {
  const int K = a->n_parties * a->r;
  if (a->r % 16 != 0 || K > 128 || K < 16 || kpad % 64 != 0 || kpad < K || a->n % 8 != 0 || a->base_fp8 ||
      a->out_fp8)
    return cudaErrorInvalidValue;
  const int atoms = kpad / 64;
  // phase 1: all-gather of the adapter slices (each peer byte crosses NVLink once)
  {
    const size_t chunks = (size_t)a->n_parties * (a->m + a->n) * (a->r / 8);
    size_t blocks = (chunks + 255) / 256;
    const size_t cap = (size_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    fed_lora_gather_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*a, reinterpret_cast<__nv_bfloat16*>(bcat),
                                                                 reinterpret_cast<__nv_bfloat16*>(atcat), kpad);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  // phase 2: persistent GEMM + update
  const int tiles_m = (int)((a->m + kLoraBM - 1) / kLoraBM), tiles_n = (int)((a->n + kLoraBN - 1) / kLoraBN);
  const int n_tiles = tiles_m * tiles_n;
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  if (deep) {
    const int smem3 = (kWStages3 * 2 + kOpStages3 * 2 * atoms) * kAtomBytes + 1024;
    static int configured3 = 0;
    if (configured3 < smem3) {
      cudaError_t e = cudaFuncSetAttribute(fed_lora_gemm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3);
      if (e != cudaSuccess) return e;
      configured3 = smem3;
    }
    fed_lora_gemm3_kernel<<<grid, kStgThreads, smem3, stream>>>(*a, *reinterpret_cast<const LoraStagedMaps*>(maps),
                                                                 atoms, tiles_n, n_tiles);
    return cudaGetLastError();
  }
  const int smem = kStgStages * (2 + 2 * atoms) * kAtomBytes + 1024;
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(fed_lora_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  fed_lora_gemm_kernel<<<grid, kStgThreads, smem, stream>>>(*a, *reinterpret_cast<const LoraStagedMaps*>(maps), atoms,
                                                            tiles_n, n_tiles);
  return cudaGetLastError();
}

extern "C" cudaError_t fedk_launch_lora_merge_tma(const LoraMergeArgs* a, const void* maps, cudaStream_t stream) {
  const int K = a->n_parties * a->r;
  if (a->r % 16 != 0 || K > 256 || K < 16 || a->n % 64 != 0 || a->base_fp8 || a->out_fp8) return cudaErrorInvalidValue;
  const int atoms = (K + 63) / 64;
  const int smem = (2 + 2 * atoms) * kLoraBM * 128 + 1024;
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(fed_lora_merge_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  dim3 grid((unsigned)((a->n + kLoraBN - 1) / kLoraBN), (unsigned)((a->m + kLoraBM - 1) / kLoraBM));
  fed_lora_merge_tma_kernel<<<grid, kLoraThreads, smem, stream>>>(*a, *reinterpret_cast<const LoraTmaMaps*>(maps));
  return cudaGetLastError();
}

extern "C" cudaError_t fedk_launch_lora_merge(const LoraMergeArgs* a, cudaStream_t stream) {
  const int K = a->n_parties * a->r;
  if (a->r % 16 != 0 || K > 256 || K < 16) return cudaErrorInvalidValue;
  if (a->n % 32 != 0) return cudaErrorInvalidValue;
  const int atoms = (K + 63) / 64;
  const int smem = 2 * atoms * kLoraBM * 128 + 1024;  // + slack for the 1024 B alignment
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(fed_lora_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  dim3 grid((unsigned)((a->n + kLoraBN - 1) / kLoraBN), (unsigned)((a->m + kLoraBM - 1) / kLoraBM));
  fed_lora_merge_kernel<<<grid, kLoraThreads, smem, stream>>>(*a);
  return cudaGetLastError();
}

}  // namespace fedk
