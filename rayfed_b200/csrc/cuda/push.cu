// K1 fed_push / K4 fed_unpack — the cross-party object push, on the device.
//
// Replaces the reference hot path  D2H -> cloudpickle -> protobuf -> HTTP/2 ->
// event table -> unpickle -> H2D  (fed/proxy/grpc/grpc_proxy.py:193-220 sender,
// :272-342 receiver) with ONE kernel on the owner party's GPU that
// reads the tensor, optionally casts / scales / block-quantises it ("serialize")
// and stores it straight into the consumer party's receive arena in *peer HBM*
// over NVLink, then publishes a release flag.  The consumer's fed_unpack kernel
// acquires the flag on its own stream (K4: the wait is on the device, the host
// never blocks), undoes the wire format into the destination tensor and returns
// flow-control credit to the producer with a peer store.
//
//   fed_push_kernel        SM-issued 16-byte vector stores (all sizes, all modes)
//   fed_push_bulk_kernel   cp.async.bulk (TMA engine) gmem->smem->peer gmem,
//                          mbarrier pipelined; plain-copy mode, large messages
//   fed_unpack_kernel      flag acquire + unpack + credit return
#include "kernels.h"

namespace fedk {

// The 64-byte descriptor goes out as four 16-byte peer stores; the release store of the flag
// that follows (same thread) orders them, and the payload, before the signal.
__device__ __forceinline__ void publish_desc(const PushArgs& a) {
  if (!a.desc_dst) return;
  const uint4* d = reinterpret_cast<const uint4*>(&a.desc);
  uint4* o = reinterpret_cast<uint4*>(a.desc_dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) st_v4(o + i, d[i]);
}

// Does the descriptor in the arena match what the host frame announced?
__device__ __forceinline__ bool desc_matches(const PushDesc& want, const void* got_ptr) {
  const uint4* g = reinterpret_cast<const uint4*>(got_ptr);
  const uint4* w = reinterpret_cast<const uint4*>(&want);
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 x = ld_v4(g + i);
    ok = ok && x.x == w[i].x && x.y == w[i].y && x.z == w[i].z && x.w == w[i].w;
  }
  return ok;
}

// Grid-wide "last CTA publishes" epilogue.
__device__ __forceinline__ void publish_when_grid_done(const PushArgs& a, bool is_unpack) {
  fence_acq_rel_sys();  // my payload stores are visible before I count myself done
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned prev = atomicAdd(a.done_counter, 1u);
    if (prev == gridDim.x - 1) {
      *a.done_counter = 0;  // re-arm for the next launch on this stream
      fence_acq_rel_sys();
      if (!is_unpack) {
        publish_desc(a);
        if (a.flag) st_release_sys_u32(a.flag, a.flag_value);
      } else {
        if (a.credit) st_release_sys_u64(a.credit, a.credit_value);
      }
    }
  }
}

__device__ __forceinline__ uint32_t ue8m0_from_amax(float amax) {
  // smallest power of two s with amax / s <= 448 (e4m3 max)
  float r = amax * (1.0f / 448.0f);
  uint32_t bits = __float_as_uint(r);
  uint32_t e = (bits >> 23) & 0xffu;
  if (bits & 0x7fffffu) e += 1;
  e = e < 1u ? 1u : (e > 253u ? 253u : e);
  return e;
}
__device__ __forceinline__ float pow2_from_biased(uint32_t e) { return __uint_as_float(e << 23); }

__device__ __forceinline__ uint32_t quant4_e4m3(float a, float b, float c, float d) {
  uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}

__device__ __forceinline__ void dequant4_e4m3(uint32_t w, float s, float* out) {
  __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w & 0xffffu), __NV_E4M3);
  __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w >> 16), __NV_E4M3);
  float2 f0 = __half22float2(*reinterpret_cast<__half2*>(&lo));
  float2 f1 = __half22float2(*reinterpret_cast<__half2*>(&hi));
  out[0] = f0.x * s;
  out[1] = f0.y * s;
  out[2] = f1.x * s;
  out[3] = f1.y * s;
}

template <bool kIsUnpack>
__global__ void __launch_bounds__(512) fed_push_kernel(PushArgs a, int mode) {
  if (kIsUnpack) {
    // K4: device-side receive barrier. Only one lane polls; the CTA parks on bar.sync.
    __shared__ int ok;
    if (threadIdx.x == 0) {
      // a.flag == nullptr: the stream was already ordered after the flag by a stream memory
      // operation (cuStreamWaitValue32) — no SM spins
      ok = (a.flag == nullptr || spin_wait_ge(a.flag, a.flag_value, a.timeout_ns)) ? 1 : 0;
      // status may live in pinned host memory (the host polls it without a CUDA call)
      if (!ok && a.status) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.status), "r"(1u) : "memory");
    }
    __syncthreads();
    if (!ok) {
      // the payload never arrived: leave the destination untouched, still return credit
      publish_when_grid_done(a, true);
      return;
    }
  }
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;

  if (mode == kCopy) {
    const size_t nvec = a.n >> 4;
    const uint4* s = reinterpret_cast<const uint4*>(a.src);
    uint4* d = reinterpret_cast<uint4*>(a.dst);
    size_t i = tid;
    // 4 independent 16 B requests in flight per thread before the first store
    for (; i + 3 * stride < nvec; i += 4 * stride) {
      uint4 v0 = kIsUnpack ? ld_v4(s + i) : ld_stream_v4(s + i);
      uint4 v1 = kIsUnpack ? ld_v4(s + i + stride) : ld_stream_v4(s + i + stride);
      uint4 v2 = kIsUnpack ? ld_v4(s + i + 2 * stride) : ld_stream_v4(s + i + 2 * stride);
      uint4 v3 = kIsUnpack ? ld_v4(s + i + 3 * stride) : ld_stream_v4(s + i + 3 * stride);
      st_v4(d + i, v0);
      st_v4(d + i + stride, v1);
      st_v4(d + i + 2 * stride, v2);
      st_v4(d + i + 3 * stride, v3);
    }
    for (; i < nvec; i += stride) st_v4(d + i, kIsUnpack ? ld_v4(s + i) : ld_stream_v4(s + i));
    if (tid == 0) {
      const unsigned char* sb = reinterpret_cast<const unsigned char*>(a.src);
      unsigned char* db = reinterpret_cast<unsigned char*>(a.dst);
      for (size_t b = nvec << 4; b < a.n; ++b) db[b] = sb[b];
    }
  } else if (mode == kF32ToBF16) {
    // 8 floats (2 x 16 B loads) -> 8 bf16 (one 16 B store)
    const size_t nvec = a.n >> 3;
    const uint4* s = reinterpret_cast<const uint4*>(a.src);
    uint4* d = reinterpret_cast<uint4*>(a.dst);
    const float sc = a.scale;
    for (size_t i = tid; i < nvec; i += stride) {
      uint4 x = ld_stream_v4(s + 2 * i), y = ld_stream_v4(s + 2 * i + 1);
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(x.x) * sc, __uint_as_float(x.y) * sc);
      o.y = pack_bf16x2(__uint_as_float(x.z) * sc, __uint_as_float(x.w) * sc);
      o.z = pack_bf16x2(__uint_as_float(y.x) * sc, __uint_as_float(y.y) * sc);
      o.w = pack_bf16x2(__uint_as_float(y.z) * sc, __uint_as_float(y.w) * sc);
      st_v4(d + i, o);
    }
    if (tid == 0) {
      const float* sf = reinterpret_cast<const float*>(a.src);
      __nv_bfloat16* db = reinterpret_cast<__nv_bfloat16*>(a.dst);
      for (size_t e = nvec << 3; e < a.n; ++e) db[e] = __float2bfloat16(sf[e] * sc);
    }
  } else if (mode == kBF16ToF32) {
    const size_t nvec = a.n >> 3;
    const uint4* s = reinterpret_cast<const uint4*>(a.src);
    uint4* d = reinterpret_cast<uint4*>(a.dst);
    const float sc = a.scale;
    for (size_t i = tid; i < nvec; i += stride) {
      uint4 x = ld_v4(s + i);
      uint4 o0, o1;
      o0.x = __float_as_uint(bf16_lo(x.x) * sc);
      o0.y = __float_as_uint(bf16_hi(x.x) * sc);
      o0.z = __float_as_uint(bf16_lo(x.y) * sc);
      o0.w = __float_as_uint(bf16_hi(x.y) * sc);
      o1.x = __float_as_uint(bf16_lo(x.z) * sc);
      o1.y = __float_as_uint(bf16_hi(x.z) * sc);
      o1.z = __float_as_uint(bf16_lo(x.w) * sc);
      o1.w = __float_as_uint(bf16_hi(x.w) * sc);
      st_v4(d + 2 * i, o0);
      st_v4(d + 2 * i + 1, o1);
    }
    if (tid == 0) {
      const __nv_bfloat16* sb = reinterpret_cast<const __nv_bfloat16*>(a.src);
      float* df = reinterpret_cast<float*>(a.dst);
      for (size_t e = nvec << 3; e < a.n; ++e) df[e] = __bfloat162float(sb[e]) * sc;
    }
  } else if (mode == kBF16Scale) {
    const size_t nvec = a.n >> 3;
    const uint4* s = reinterpret_cast<const uint4*>(a.src);
    uint4* d = reinterpret_cast<uint4*>(a.dst);
    const float sc = a.scale;
    for (size_t i = tid; i < nvec; i += stride) {
      uint4 x = ld_stream_v4(s + i);
      uint4 o;
      o.x = pack_bf16x2(bf16_lo(x.x) * sc, bf16_hi(x.x) * sc);
      o.y = pack_bf16x2(bf16_lo(x.y) * sc, bf16_hi(x.y) * sc);
      o.z = pack_bf16x2(bf16_lo(x.z) * sc, bf16_hi(x.z) * sc);
      o.w = pack_bf16x2(bf16_lo(x.w) * sc, bf16_hi(x.w) * sc);
      st_v4(d + i, o);
    }
    if (tid == 0) {
      const __nv_bfloat16* sb = reinterpret_cast<const __nv_bfloat16*>(a.src);
      __nv_bfloat16* db = reinterpret_cast<__nv_bfloat16*>(a.dst);
      for (size_t e = nvec << 3; e < a.n; ++e) db[e] = __float2bfloat16(__bfloat162float(sb[e]) * sc);
    }
  } else if (mode == kBF16ToFP8) {
    // n % 32 == 0 (host guarantees). One lane = 8 elements; 4 lanes = one scale block.
    const size_t nvec = a.n >> 3;
    const uint4* s = reinterpret_cast<const uint4*>(a.src);
    uint2* d = reinterpret_cast<uint2*>(a.dst);
    unsigned char* sc_out = reinterpret_cast<unsigned char*>(a.dst_scales);
    const float pre = a.scale;
    const size_t nvec_warp_rounded = (nvec + 31) & ~(size_t)31;
    for (size_t i = tid; i < nvec_warp_rounded; i += stride) {
      const bool live = i < nvec;
      uint4 x = live ? ld_stream_v4(s + i) : make_uint4(0, 0, 0, 0);
      float f[8] = {bf16_lo(x.x) * pre, bf16_hi(x.x) * pre, bf16_lo(x.y) * pre, bf16_hi(x.y) * pre,
                    bf16_lo(x.z) * pre, bf16_hi(x.z) * pre, bf16_lo(x.w) * pre, bf16_hi(x.w) * pre};
      float amax = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) amax = fmaxf(amax, fabsf(f[k]));
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
      const uint32_t e = ue8m0_from_amax(amax);
      const float inv = pow2_from_biased(254u - e);  // 1 / 2^(e-127)
      if (live) {
        uint2 o;
        o.x = quant4_e4m3(f[0] * inv, f[1] * inv, f[2] * inv, f[3] * inv);
        o.y = quant4_e4m3(f[4] * inv, f[5] * inv, f[6] * inv, f[7] * inv);
        st_v2(d + i, o);
      }
      // gather the warp's 8 scale bytes into one 8-byte peer store
      const int lane = threadIdx.x & 31;
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) lo |= (__shfl_sync(0xffffffffu, e, k * 4) & 0xffu) << (8 * k);
#pragma unroll
      for (int k = 0; k < 4; ++k) hi |= (__shfl_sync(0xffffffffu, e, 16 + k * 4) & 0xffu) << (8 * k);
      const size_t warp_first = i - lane;
      if (lane == 0) {
        const size_t blk0 = warp_first >> 2;
        const size_t nblk = a.n >> 5;
        if (blk0 + 8 <= nblk) {
          st_v2(sc_out + blk0, make_uint2(lo, hi));
        } else {
          for (size_t b = blk0; b < nblk; ++b) {
            uint32_t k = (uint32_t)(b - blk0);
            sc_out[b] = (unsigned char)((k < 4 ? (lo >> (8 * k)) : (hi >> (8 * (k - 4)))) & 0xffu);
          }
        }
      }
    }
  } else if (mode == kFP8ToBF16) {
    const size_t nvec = a.n >> 3;
    const uint2* s = reinterpret_cast<const uint2*>(a.src);
    uint4* d = reinterpret_cast<uint4*>(a.dst);
    const unsigned char* sc_in = reinterpret_cast<const unsigned char*>(a.dst_scales);
    const float post = a.scale;
    for (size_t i = tid; i < nvec; i += stride) {
      uint2 x = ld_v2(s + i);
      const float sc = pow2_from_biased((uint32_t)sc_in[i >> 2]) * post;
      float f[8];
      dequant4_e4m3(x.x, sc, f);
      dequant4_e4m3(x.y, sc, f + 4);
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]);
      o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]);
      o.w = pack_bf16x2(f[6], f[7]);
      st_v4(d + i, o);
    }
  }
  publish_when_grid_done(a, kIsUnpack);
}

// ---------------------------------------------------------------------------
// Multi-tensor push.  Work items are 64 KiB slices of the concatenated segments; a CTA walks
// items blockIdx.x, +gridDim.x, ... (segment lookup is a linear scan over <= kMaxSegs entries
// in constant/param space, once per item).
constexpr unsigned long long kSegItem = 64 * 1024;

__global__ void __launch_bounds__(512) fed_push_multi_kernel(const __grid_constant__ PushArgs a,
                                                             const __grid_constant__ PushSegs segs) {
  unsigned long long total_items = 0;
  for (int i = 0; i < segs.n; ++i) total_items += (segs.s[i].nbytes + kSegItem - 1) / kSegItem;
  char* base = reinterpret_cast<char*>(a.dst);
  for (unsigned long long item = blockIdx.x; item < total_items; item += gridDim.x) {
    unsigned long long rem = item;
    int si = 0;
    for (; si < segs.n; ++si) {
      unsigned long long cnt = (segs.s[si].nbytes + kSegItem - 1) / kSegItem;
      if (rem < cnt) break;
      rem -= cnt;
    }
    const unsigned long long off = rem * kSegItem;
    const unsigned long long left = segs.s[si].nbytes - off;
    const unsigned long long nb = left < kSegItem ? left : kSegItem;
    const char* s = reinterpret_cast<const char*>(segs.s[si].src) + off;
    char* d = base + segs.s[si].dst_off + off;
    const unsigned nvec = (unsigned)(nb >> 4);
    const uint4* sv = reinterpret_cast<const uint4*>(s);
    uint4* dv = reinterpret_cast<uint4*>(d);
    unsigned i = threadIdx.x;
    for (; i + 3 * 512 < nvec; i += 4 * 512) {
      uint4 v0 = ld_stream_v4(sv + i), v1 = ld_stream_v4(sv + i + 512);
      uint4 v2 = ld_stream_v4(sv + i + 1024), v3 = ld_stream_v4(sv + i + 1536);
      st_v4(dv + i, v0);
      st_v4(dv + i + 512, v1);
      st_v4(dv + i + 1024, v2);
      st_v4(dv + i + 1536, v3);
    }
    for (; i < nvec; i += 512) st_v4(dv + i, ld_stream_v4(sv + i));
    if (threadIdx.x == 0)
      for (unsigned long long b = (unsigned long long)nvec << 4; b < nb; ++b) d[b] = s[b];
  }
  publish_when_grid_done(a, false);
}

// One warp: compare the arena descriptor with the announced one; raise status 3 on mismatch.
__global__ void fed_desc_check_kernel(const __grid_constant__ PushArgs a) {
  if (threadIdx.x == 0 && !desc_matches(a.desc, a.src)) {
    if (a.status) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.status), "r"(3u) : "memory");
  }
}

// Flow control: "block <value> of my region in your arena is free again" — one entry of the
// freed ring in the *producer's* control block, then the ring's count (release).
__global__ void fed_credit_kernel(unsigned long long* ring_slot, unsigned long long value,
                                  unsigned long long* count_ptr, unsigned long long count) {
  if (threadIdx.x == 0) {
    if (ring_slot) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(ring_slot), "l"(value) : "memory");
    st_release_sys_u64(count_ptr, count);
  }
}

// ---------------------------------------------------------------------------
// TMA-engine variant: cp.async.bulk global->shared (mbarrier complete_tx) then
// cp.async.bulk shared->peer global (bulk_group). One elected thread per CTA
// drives a kStages-deep ring; the SMs issue no per-byte instructions at all.
constexpr int kBulkStages = 4;
constexpr int kBulkChunk = 32 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

__global__ void __launch_bounds__(128) fed_push_bulk_kernel(PushArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[kBulkStages];
  const size_t nbytes = a.n;
  const size_t nchunks = (nbytes + kBulkChunk - 1) / kBulkChunk;
  const char* src = reinterpret_cast<const char*>(a.src);
  char* dst = reinterpret_cast<char*>(a.dst);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBulkStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    auto chunk_bytes = [&](size_t c) -> uint32_t {
      size_t off = c * (size_t)kBulkChunk;
      size_t rem = nbytes - off;
      return (uint32_t)(rem < (size_t)kBulkChunk ? rem : (size_t)kBulkChunk);
    };
    // my chunks: blockIdx.x, +gridDim.x, ...
    size_t my_n = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // prologue: fill the ring
    for (size_t it = 0; it < my_n && it < (size_t)kBulkStages; ++it) {
      size_t c = blockIdx.x + it * gridDim.x;
      uint32_t b = chunk_bytes(c);
      mbar_expect_tx(&full[it], b);
      bulk_g2s(smem + it * kBulkChunk, src + c * (size_t)kBulkChunk, b, &full[it]);
    }
    for (size_t it = 0; it < my_n; ++it) {
      const int s = (int)(it % kBulkStages);
      const uint32_t parity = (uint32_t)((it / kBulkStages) & 1);
      size_t c = blockIdx.x + it * gridDim.x;
      uint32_t b = chunk_bytes(c);
      mbar_wait(&full[s], parity);
      bulk_s2g(dst + c * (size_t)kBulkChunk, smem + s * kBulkChunk, b);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // refill the stage whose store was issued one iteration ago
      if (it >= 1) {
        size_t nxt = it - 1 + kBulkStages;
        if (nxt < my_n) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          const int ps = (int)((it - 1) % kBulkStages);
          size_t c2 = blockIdx.x + nxt * gridDim.x;
          uint32_t b2 = chunk_bytes(c2);
          mbar_expect_tx(&full[ps], b2);
          bulk_g2s(smem + ps * kBulkChunk, src + c2 * (size_t)kBulkChunk, b2, &full[ps]);
        }
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // writes complete, not just smem reads
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  publish_when_grid_done(a, false);
}

// ---------------------------------------------------------------------------
// K1b fed_push_mc: one multimem.st stream through the NVSwitch reaches every party
// bound to the multicast object (fed.get's owner -> all broadcast, weight
// redistribution).  Byte copy only; the release flag is then stored to each peer.
__global__ void __launch_bounds__(512) fed_push_mc_kernel(const __grid_constant__ PushArgs a,
                                                          const __grid_constant__ PeerPtrs flags, int n_flags) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nvec = a.n >> 4;  // host guarantees n % 16 == 0
  const uint4* s = reinterpret_cast<const uint4*>(a.src);
  char* d = reinterpret_cast<char*>(a.dst);  // multicast alias of the broadcast arena slot
  size_t i = tid;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    uint4 v0 = ld_stream_v4(s + i), v1 = ld_stream_v4(s + i + stride);
    uint4 v2 = ld_stream_v4(s + i + 2 * stride), v3 = ld_stream_v4(s + i + 3 * stride);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(d + (i << 4)), "r"(v0.x),
                 "r"(v0.y), "r"(v0.z), "r"(v0.w) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(d + ((i + stride) << 4)),
                 "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(d + ((i + 2 * stride) << 4)),
                 "r"(v2.x), "r"(v2.y), "r"(v2.z), "r"(v2.w) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(d + ((i + 3 * stride) << 4)),
                 "r"(v3.x), "r"(v3.y), "r"(v3.z), "r"(v3.w) : "memory");
  }
  for (; i < nvec; i += stride) {
    uint4 v = ld_stream_v4(s + i);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(d + (i << 4)), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
  fence_acq_rel_sys();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned prev = atomicAdd(a.done_counter, 1u);
    if (prev == gridDim.x - 1) {
      *a.done_counter = 0;
      if (a.desc_dst) {  // the block header, replicated by the switch like the payload
        const uint4* dd = reinterpret_cast<const uint4*>(&a.desc);
        char* o = reinterpret_cast<char*>(a.desc_dst);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + 16 * k), "r"(dd[k].x),
                       "r"(dd[k].y), "r"(dd[k].z), "r"(dd[k].w) : "memory");
      }
      fence_acq_rel_sys();
      for (int p = 0; p < n_flags; ++p) st_release_sys_u32(reinterpret_cast<uint32_t*>(flags.p[p]), a.flag_value);
    }
  }
}

// ---------------------------------------------------------------------------
extern "C" {

cudaError_t fedk_launch_push_mc(const PushArgs* a, const PeerPtrs* flags, int n_flags, int blocks,
                                cudaStream_t stream) {
  fed_push_mc_kernel<<<blocks, 512, 0, stream>>>(*a, *flags, n_flags);
  return cudaGetLastError();
}

cudaError_t fedk_launch_push(const PushArgs* a, int mode, int blocks, int use_bulk, cudaStream_t stream) {
  if (use_bulk && mode == kCopy) {
    const int smem = kBulkStages * kBulkChunk;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(fed_push_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    fed_push_bulk_kernel<<<blocks, 128, smem, stream>>>(*a);
  } else {
    fed_push_kernel<false><<<blocks, 512, 0, stream>>>(*a, mode);
  }
  return cudaGetLastError();
}


#define FEDK_PRELOAD(fn)                                                        \
  do {                                                                          \
    cudaFuncAttributes at_;                                                     \
    cudaError_t e_ = cudaFuncGetAttributes(&at_, fn);                           \
    if (e_ != cudaSuccess) return e_;                                           \
  } while (0)

cudaError_t fedk_preload_push() {
  FEDK_PRELOAD(fed_push_kernel<false>);
  FEDK_PRELOAD(fed_push_kernel<true>);
  FEDK_PRELOAD(fed_push_bulk_kernel);
  FEDK_PRELOAD(fed_push_mc_kernel);
  FEDK_PRELOAD(fed_push_multi_kernel);
  FEDK_PRELOAD(fed_desc_check_kernel);
  FEDK_PRELOAD(fed_credit_kernel);
  return cudaSuccess;
}

cudaError_t fedk_launch_push_multi(const PushArgs* a, const PushSegs* segs, int blocks, cudaStream_t stream) {
  fed_push_multi_kernel<<<blocks, 512, 0, stream>>>(*a, *segs);
  return cudaGetLastError();
}

cudaError_t fedk_launch_desc_check(const PushArgs* a, cudaStream_t stream) {
  fed_desc_check_kernel<<<1, 32, 0, stream>>>(*a);
  return cudaGetLastError();
}

cudaError_t fedk_launch_credit(unsigned long long* ring_slot, unsigned long long value, unsigned long long* count_ptr,
                               unsigned long long count, cudaStream_t stream) {
  fed_credit_kernel<<<1, 32, 0, stream>>>(ring_slot, value, count_ptr, count);
  return cudaGetLastError();
}

cudaError_t fedk_launch_unpack(const PushArgs* a, int mode, int blocks, cudaStream_t stream) {
  fed_push_kernel<true><<<blocks, 512, 0, stream>>>(*a, mode);
  return cudaGetLastError();
}

}  // extern "C"

}  // namespace fedk
