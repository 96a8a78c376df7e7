// K2 fed_aggregate_sgd / K3 fed_aggregate_adam(_fp8) / K5 fed_barrier.
//
// The federated aggregate.  In the reference this is user code composed from
// pushes: every party ships its weights to one party over gRPC, that party runs
// `mean`, and the result is shipped back (fed/tests/test_fed_get.py:45-47,71-78;
// SURVEY §2.D C4).  Here it is ONE kernel per party that fuses
//
//   cross-party all-reduce (two-shot: every party owns 1/N of the elements,
//   pulls that shard from all N parties over NVLink — P2P loads, or one
//   multimem.ld_reduce through the NVSwitch when a multicast mapping exists),
//   the 1/N (or n_k/n weighted) scale,
//   the server/optimizer update in registers (FedAvg, FedAvgM, SGD+momentum+wd,
//   Adam/AdamW, FedAdam) against a ZeRO-style *sharded* fp32 master + moments
//   kept only by the owning party, and
//   the all-gather of the new weights (P2P stores to every party, or one
//   multimem.st), optionally re-quantised to block-scaled fp8 e4m3 (UE8M0 / 32)
//
// with the cross-GPU synchronisation done inside the kernel on a signal pad
// (per-CTA flags, st.release.sys / ld.acquire.sys), so there is no host hop and
// no separate NCCL call or optimizer launch.
#include <cstdlib>

#include "kernels.h"

namespace fedk {

// All CTAs with the same blockIdx on every party rendezvous.  Entry: "my inputs
// are final" (stream order + release).  Exit: "my reads of your inputs and my
// writes of your outputs are done".
__device__ __forceinline__ void cta_party_barrier(const AggArgs& a, uint32_t value) {
  // bar.sync orders every thread's earlier accesses before the signalling threads' fence at CTA
  // scope; their fence.acq_rel.sys + st.release.sys then publishes them system-wide (the memory
  // model's causality order is cumulative) — one system fence per signalling thread instead of one
  // per thread of the CTA (a MEMBAR.SYS costs microseconds).
  __syncthreads();
  const int t = threadIdx.x;
  if (t < a.n_parties && t != a.my_rank) {
    fence_acq_rel_sys();
    uint32_t* remote = reinterpret_cast<uint32_t*>(a.sig_peers.p[t]) + (size_t)blockIdx.x * kMaxParties + a.my_rank;
    st_release_sys_u32(remote, value);
    const uint32_t* mine = a.sig + (size_t)blockIdx.x * kMaxParties + t;
    if (!spin_wait_ge(mine, value, a.timeout_ns) && a.status)
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.status), "r"(2u) : "memory");
    fence_acq_rel_sys();  // acquire side: peers' writes that preceded their signal are visible to the CTA
  }
  __syncthreads();
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& x, float* f) {
  f[0] = bf16_lo(x.x); f[1] = bf16_hi(x.x); f[2] = bf16_lo(x.y); f[3] = bf16_hi(x.y);
  f[4] = bf16_lo(x.z); f[5] = bf16_hi(x.z); f[6] = bf16_lo(x.w); f[7] = bf16_hi(x.w);
}

__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ uint4 multimem_ld_reduce_f32x4(const void* mc) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_v4(void* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_st_v2(void* mc, const uint2& v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};" ::"l"(mc), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ uint32_t agg_ue8m0(float amax) {
  float r = amax * (1.0f / 448.0f);
  uint32_t bits = __float_as_uint(r);
  uint32_t e = (bits >> 23) & 0xffu;
  if (bits & 0x7fffffu) e += 1;
  return e < 1u ? 1u : (e > 253u ? 253u : e);
}
__device__ __forceinline__ uint32_t agg_quant4(float a, float b, float c, float d) {
  uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}

// Optimizer/server update for 8 consecutive elements held in registers, then the
// all-gather store of the result to every party (or one multicast store).
// Optimizer state of one 8-element vector, fetched together with the contributions
// so that its HBM latency overlaps theirs.
struct StateRegs {
  float4 m0, m1, q0, q1, r0, r1;
};

__device__ __forceinline__ void load_state(const AggArgs& a, size_t v, StateRegs& st) {
  const size_t s0 = v << 3;
  st.m0 = *reinterpret_cast<const float4*>(a.master + s0);
  st.m1 = *reinterpret_cast<const float4*>(a.master + s0 + 4);
  st.q0 = *reinterpret_cast<const float4*>(a.mom + s0);
  st.q1 = *reinterpret_cast<const float4*>(a.mom + s0 + 4);
  if (a.kind == kAdamGrad || a.kind == kFedAdam) {
    st.r0 = *reinterpret_cast<const float4*>(a.var + s0);
    st.r1 = *reinterpret_cast<const float4*>(a.var + s0 + 4);
  }
}

template <bool kUseMc>
__device__ __forceinline__ void finish_vector(const AggArgs& a, size_t v, size_t nvec, bool live, float* acc,
                                              int lane, const StateRegs& st) {
  const int N = a.n_parties;
  const size_t e0 = a.shard_start + (v << 3);
  if (live && a.kind != kAvg) {
    const size_t s0 = v << 3;  // index into the local shard state
    const float4 m0 = st.m0, m1 = st.m1, q0 = st.q0, q1 = st.q1;
    float ms[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float mo[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    if (a.kind == kFedAvgM || a.kind == kSgdGrad) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float g = (a.kind == kFedAvgM) ? (ms[k] - acc[k]) : (acc[k] + a.weight_decay * ms[k]);
        mo[k] = a.momentum * mo[k] + g;
        ms[k] -= a.lr * mo[k];
        acc[k] = ms[k];
      }
    } else {
      const float4 r0 = st.r0, r1 = st.r1;
      float va[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float g = (a.kind == kFedAdam) ? (ms[k] - acc[k]) : acc[k];
        mo[k] = a.beta1 * mo[k] + (1.f - a.beta1) * g;
        va[k] = a.beta2 * va[k] + (1.f - a.beta2) * g * g;
        float mhat = mo[k] * a.bias_c1;  // 1 / (1 - beta1^t)
        float vhat = va[k] * a.bias_c2;  // 1 / (1 - beta2^t)
        ms[k] -= a.lr * (mhat / (sqrtf(vhat) + a.eps) + a.weight_decay * ms[k]);
        acc[k] = ms[k];
      }
      *reinterpret_cast<float4*>(a.var + s0) = make_float4(va[0], va[1], va[2], va[3]);
      *reinterpret_cast<float4*>(a.var + s0 + 4) = make_float4(va[4], va[5], va[6], va[7]);
    }
    *reinterpret_cast<float4*>(a.master + s0) = make_float4(ms[0], ms[1], ms[2], ms[3]);
    *reinterpret_cast<float4*>(a.master + s0 + 4) = make_float4(ms[4], ms[5], ms[6], ms[7]);
    *reinterpret_cast<float4*>(a.mom + s0) = make_float4(mo[0], mo[1], mo[2], mo[3]);
    *reinterpret_cast<float4*>(a.mom + s0 + 4) = make_float4(mo[4], mo[5], mo[6], mo[7]);
  }

  if (!a.out_fp8) {
    if (live) {
      uint4 o;
      o.x = pack_bf16x2(acc[0], acc[1]);
      o.y = pack_bf16x2(acc[2], acc[3]);
      o.z = pack_bf16x2(acc[4], acc[5]);
      o.w = pack_bf16x2(acc[6], acc[7]);
      if (kUseMc) {
        multimem_st_v4(reinterpret_cast<char*>(a.mc_out) + e0 * 2, o);
      } else {
        for (int j = 0; j < N; ++j) st_v4(reinterpret_cast<char*>(a.out.p[(a.my_rank + j) % N]) + e0 * 2, o);
      }
    }
  } else {
    // block-scaled e4m3: 4 lanes share one UE8M0 scale (32 elements)
    float amax = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) amax = fmaxf(amax, fabsf(acc[k]));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    const uint32_t e = agg_ue8m0(amax);
    const float inv = __uint_as_float((254u - e) << 23);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) lo |= (__shfl_sync(0xffffffffu, e, k * 4) & 0xffu) << (8 * k);
#pragma unroll
    for (int k = 0; k < 4; ++k) hi |= (__shfl_sync(0xffffffffu, e, 16 + k * 4) & 0xffu) << (8 * k);
    if (live) {
      uint2 o;
      o.x = agg_quant4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
      o.y = agg_quant4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
      for (int j = 0; j < N; ++j) st_v2(reinterpret_cast<char*>(a.out.p[(a.my_rank + j) % N]) + e0, o);
    }
    if (lane == 0 && (v - lane) < nvec) {
      const size_t blk0 = (a.shard_start >> 5) + ((v - lane) >> 2);
      const size_t nblk_here = min((size_t)8, ((nvec - (v - lane)) + 3) >> 2);
      for (int j = 0; j < N; ++j) {
        unsigned char* sc = reinterpret_cast<unsigned char*>(a.out_scales.p[(a.my_rank + j) % N]);
        if (nblk_here == 8 && ((blk0 & 7) == 0)) {
          st_v2(sc + blk0, make_uint2(lo, hi));
        } else {
          for (size_t b = 0; b < nblk_here; ++b)
            sc[blk0 + b] = (unsigned char)((b < 4 ? (lo >> (8 * b)) : (hi >> (8 * (b - 4)))) & 0xffu);
        }
      }
    }
  }
}

// kNP = parties whose 16-byte requests are preloaded per vector, kU = vectors per
// thread per trip; kNP * kU outstanding 16 B loads per thread (8 => 32 registers),
// all issued before the first use, so ~1 MB/SM is in flight against NVLink latency.
template <bool kUseMc, int kNP, int kU, bool kState>
__global__ void __launch_bounds__(512) fed_aggregate_kernel(const __grid_constant__ AggArgs a) {
  // graph-replayable epoch (kept in a register: writing to the parameter struct would
  // force a local-memory copy of all 700+ bytes of it)
  const uint32_t epoch = a.epoch_dev ? ld_relaxed_sys_u32(a.epoch_dev) : a.epoch;
  if (!a.entry_ordered) cta_party_barrier(a, epoch + 1);

  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nvec = a.shard_len >> 3;
  const size_t nvec_rounded = (nvec + 31) & ~(size_t)31;  // keep warps converged for the fp8 shuffles
  const int N = a.n_parties;
  const int lane = threadIdx.x & 31;

  if (a.in_fp32) {
    // fp32 contributions (gradient buckets): two 16 B requests per party per vector
    for (size_t v = tid; v < nvec_rounded; v += stride) {
      const bool live = v < nvec;
      const size_t e0 = a.shard_start + (v << 3);
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      StateRegs st;
      if (kState && live) load_state(a, v, st);
      if (live) {
        if (kUseMc) {
          const char* p = reinterpret_cast<const char*>(a.mc_in) + e0 * 4;
          uint4 s0 = multimem_ld_reduce_f32x4(p), s1 = multimem_ld_reduce_f32x4(p + 16);
          const float w0 = a.w[0];
          acc[0] = w0 * __uint_as_float(s0.x); acc[1] = w0 * __uint_as_float(s0.y);
          acc[2] = w0 * __uint_as_float(s0.z); acc[3] = w0 * __uint_as_float(s0.w);
          acc[4] = w0 * __uint_as_float(s1.x); acc[5] = w0 * __uint_as_float(s1.y);
          acc[6] = w0 * __uint_as_float(s1.z); acc[7] = w0 * __uint_as_float(s1.w);
        } else {
          for (int j0 = 0; j0 < N; j0 += 4) {
            uint4 x0[4], x1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j0 + j < N) {
                const char* p = reinterpret_cast<const char*>(a.in.p[(a.my_rank + j0 + j) % N]) + e0 * 4;
                x0[j] = ld_v4(p);
                x1[j] = ld_v4(p + 16);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j0 + j < N) {
                const float wj = a.w[(a.my_rank + j0 + j) % N];
                acc[0] += wj * __uint_as_float(x0[j].x); acc[1] += wj * __uint_as_float(x0[j].y);
                acc[2] += wj * __uint_as_float(x0[j].z); acc[3] += wj * __uint_as_float(x0[j].w);
                acc[4] += wj * __uint_as_float(x1[j].x); acc[5] += wj * __uint_as_float(x1[j].y);
                acc[6] += wj * __uint_as_float(x1[j].z); acc[7] += wj * __uint_as_float(x1[j].w);
              }
            }
          }
        }
      }
      finish_vector<kUseMc>(a, v, nvec, live, acc, lane, st);
    }
  } else {
    for (size_t v0 = tid; v0 < nvec_rounded; v0 += stride * kU) {
      uint4 x[kU][kNP];
      bool live[kU];
      StateRegs st[kState ? kU : 1];
      // ---- phase 1: issue every request of this trip ------------------------------
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const size_t v = v0 + (size_t)u * stride;
        live[u] = v < nvec;
        if (live[u]) {
          if (kState) load_state(a, v, st[u]);
          const size_t e0 = a.shard_start + (v << 3);
          if (kUseMc) {
            x[u][0] = multimem_ld_reduce_bf16x8(reinterpret_cast<const char*>(a.mc_in) + e0 * 2);
          } else {
#pragma unroll
            for (int j = 0; j < kNP; ++j)
              if (j < N) x[u][j] = ld_v4(reinterpret_cast<const char*>(a.in.p[(a.my_rank + j) % N]) + e0 * 2);
          }
        }
      }
      // ---- phase 2: reduce, update, all-gather ---------------------------------------
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const size_t v = v0 + (size_t)u * stride;
        if (v >= nvec_rounded) continue;  // (not `break`: keeps the unrolled arrays in registers)
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        if (live[u]) {
          if (kUseMc) {
            float f[8];
            unpack_bf16x8(x[u][0], f);
            const float w0 = a.w[0];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = w0 * f[k];
          } else {
#pragma unroll
            for (int j = 0; j < kNP; ++j) {
              if (j < N) {
                float f[8];
                unpack_bf16x8(x[u][j], f);
                const float wj = a.w[(a.my_rank + j) % N];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += wj * f[k];
              }
            }
            const size_t e0 = a.shard_start + (v << 3);
            for (int j = kNP; j < N; ++j) {  // more parties than preload slots
              uint4 y = ld_v4(reinterpret_cast<const char*>(a.in.p[(a.my_rank + j) % N]) + e0 * 2);
              float f[8];
              unpack_bf16x8(y, f);
              const float wj = a.w[(a.my_rank + j) % N];
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] += wj * f[k];
            }
          }
        }
        finish_vector<kUseMc>(a, v, nvec, live[u], acc, lane, st[kState ? u : 0]);
      }
    }
  }

  cta_party_barrier(a, epoch + 2);
  if (a.epoch_dev && threadIdx.x == 0) {
    unsigned prev = atomicAdd(a.agg_done, 1u);
    if (prev == gridDim.x - 1) {  // every CTA has read the epoch and passed both barriers
      *a.agg_done = 0;
      *a.epoch_dev = epoch + 2;
    }
  }
}

// ---------------------------------------------------------------------------
// Plain FedAvg of bf16 weights (the flagship path): a lean, software-pipelined version of the
// kernel above.  The generic kernel issues ALL of a thread's loads, then all of its stores — with
// a 6 MB shard per party that is one trip, so on every GPU the reduce phase (egress-heavy: the
// peers/switch pull my buffer) and the gather phase (ingress-heavy) run one after the other and
// each NVLink direction idles half of the time.  Here every thread keeps kD vectors in flight and
// re-issues the load of vector i+kD right before it reduces/stores vector i, with 2-3 CTAs per SM
// (40 registers): loads and stores of different trips overlap, both directions stay busy.
template <bool kUseMc, int kNP, int kD>
__global__ void __launch_bounds__(512, 2) fed_allreduce_avg_kernel(const __grid_constant__ AggArgs a) {
  const uint32_t epoch = a.epoch_dev ? ld_relaxed_sys_u32(a.epoch_dev) : a.epoch;
  unsigned long long* ts = a.phase_ts ? a.phase_ts + (size_t)blockIdx.x * 4 : nullptr;
  if (ts && threadIdx.x == 0) ts[0] = global_ns();
  // With the pre-launch rendezvous (FedAggregator._rendezvous) every party's inputs are final before
  // any party's kernel starts: the entry barrier has nothing left to order.
  if (!a.entry_ordered) cta_party_barrier(a, epoch + 1);
  if (ts && threadIdx.x == 0) ts[1] = global_ns();

  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nvec = a.shard_len >> 3;
  const int N = a.n_parties;
  const char* in_mc = reinterpret_cast<const char*>(a.mc_in) + a.shard_start * 2;
  char* out_mc = reinterpret_cast<char*>(a.mc_out) + a.shard_start * 2;

  uint4 x[kD][kNP];
  auto issue = [&](int d, size_t v) {
    if (kUseMc) {
      x[d][0] = multimem_ld_reduce_bf16x8(in_mc + (v << 4));
    } else {
#pragma unroll
      for (int j = 0; j < kNP; ++j)
        if (j < N)
          x[d][j] = ld_v4(reinterpret_cast<const char*>(a.in.p[(a.my_rank + j) % N]) + ((a.shard_start + (v << 3)) << 1));
    }
  };
#pragma unroll
  for (int d = 0; d < kD; ++d) {
    const size_t v = tid + (size_t)d * stride;
    if (v < nvec) issue(d, v);
  }
  for (size_t v0 = tid; v0 < nvec; v0 += stride * kD) {
#pragma unroll
    for (int d = 0; d < kD; ++d) {
      const size_t v = v0 + (size_t)d * stride;
      if (v >= nvec) continue;
      float acc[8];
      if (kUseMc) {
        float f[8];
        unpack_bf16x8(x[d][0], f);
        const float w0 = a.w[0];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = w0 * f[k];
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
        for (int j = 0; j < kNP; ++j) {
          if (j < N) {
            float f[8];
            unpack_bf16x8(x[d][j], f);
            const float wj = a.w[(a.my_rank + j) % N];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += wj * f[k];
          }
        }
      }
      // the slot is consumed: put the next trip's request on the wire before storing
      const size_t vn = v + stride * kD;
      if (vn < nvec) issue(d, vn);
      uint4 o;
      o.x = pack_bf16x2(acc[0], acc[1]);
      o.y = pack_bf16x2(acc[2], acc[3]);
      o.z = pack_bf16x2(acc[4], acc[5]);
      o.w = pack_bf16x2(acc[6], acc[7]);
      if (kUseMc) {
        multimem_st_v4(out_mc + (v << 4), o);
      } else {
        const size_t off = (a.shard_start + (v << 3)) << 1;
        for (int j = 0; j < N; ++j) st_v4(reinterpret_cast<char*>(a.out.p[(a.my_rank + j) % N]) + off, o);
      }
    }
  }
  if (ts && threadIdx.x == 0) ts[2] = global_ns();
  cta_party_barrier(a, epoch + 2);
  if (ts && threadIdx.x == 0) ts[3] = global_ns();
  if (a.epoch_dev && threadIdx.x == 0) {
    unsigned prev = atomicAdd(a.agg_done, 1u);
    if (prev == gridDim.x - 1) {
      *a.agg_done = 0;
      *a.epoch_dev = epoch + 2;
    }
  }
}

// "I have reached the collective": one st.release.sys per peer, into word [my_rank] of the peers'
// arrival arrays.  The launching stream then parks on its own arrival words with stream memory
// operations (cuStreamWaitValue32) — see FedAggregator.aggregate: the compute kernel, which spins in
// its barriers and fills the GPU, is not even launched before every party is there, so it can
// never starve a local kernel (a queued push) that a late party is still waiting for.
__global__ void fed_signal_kernel(const __grid_constant__ PeerPtrs arrive, int n_parties, int my_rank,
                                  uint32_t value) {
  const int t = threadIdx.x;
  fence_acq_rel_sys();
  if (t < n_parties && t != my_rank)
    st_release_sys_u32(reinterpret_cast<uint32_t*>(arrive.p[t]) + my_rank, value);
}

// K5: N-party device barrier (replaces the reference's ping loop,
// fed/proxy/barriers.py:497-523, once the symmetric heap is up).
__global__ void fed_barrier_kernel(const __grid_constant__ AggArgs a) { cta_party_barrier(a, a.epoch + 1); }

// Fused local optimizer step on flat buffers: SGD + momentum + weight decay,
// optional gradient clear in the same pass (saves a memset of the grad bucket).
__global__ void __launch_bounds__(512) fed_local_sgd_kernel(LocalSgdArgs a) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nvec = a.n >> 3;
  for (size_t v = tid; v < nvec; v += stride) {
    uint4 wv = *reinterpret_cast<const uint4*>(a.w + (v << 3));
    uint4 gv = *reinterpret_cast<const uint4*>(a.g + (v << 3));
    float4 q0 = *reinterpret_cast<const float4*>(a.mom + (v << 3));
    float4 q1 = *reinterpret_cast<const float4*>(a.mom + (v << 3) + 4);
    float w[8], g[8];
    float mo[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    unpack_bf16x8(wv, w);
    unpack_bf16x8(gv, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float gg = g[k] * a.grad_scale + a.weight_decay * w[k];
      mo[k] = a.momentum * mo[k] + gg;
      w[k] -= a.lr * mo[k];
    }
    uint4 o;
    o.x = pack_bf16x2(w[0], w[1]);
    o.y = pack_bf16x2(w[2], w[3]);
    o.z = pack_bf16x2(w[4], w[5]);
    o.w = pack_bf16x2(w[6], w[7]);
    *reinterpret_cast<uint4*>(a.w + (v << 3)) = o;
    *reinterpret_cast<float4*>(a.mom + (v << 3)) = make_float4(mo[0], mo[1], mo[2], mo[3]);
    *reinterpret_cast<float4*>(a.mom + (v << 3) + 4) = make_float4(mo[4], mo[5], mo[6], mo[7]);
    if (a.zero_grad) *reinterpret_cast<uint4*>(a.g + (v << 3)) = make_uint4(0, 0, 0, 0);
  }
}

extern "C" {

// Grid of the pipelined FedAvg kernel: every CTA of an SM resident at once (2 per SM with 512
// threads at <= 64 registers), never more CTAs than 16-byte vectors / 512, and never more than the
// signal pad has per-CTA barrier slots for (64 KiB / (kMaxParties * 4 B) = 1024).
static int lean_grid(const AggArgs* a) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const unsigned long long nvec = a->shard_len >> 3;
  long long want = (long long)((nvec + 511) / 512);
  // One CTA per SM by default: the kernel spins in cross-party barriers, and half of every SM's
  // registers / thread slots stay free for whatever else this GPU has queued (measured at N=2:
  // 1 and 2 CTAs per SM are both link-bound, 115 us).  RAYFED_B200_AGG_CTAS_PER_SM=2 fills the SMs.
  long long cap = (long long)sms;
  if (const char* e = getenv("RAYFED_B200_AGG_CTAS_PER_SM")) cap = (long long)sms * (atoi(e) > 0 ? atoi(e) : 1);
  // NVLS (multimem.ld_reduce / multimem.st) wants FEW CTAs: measured at N=8, 51 MB: 163 us with 32
  // CTAs, 169 with 16 or 64, 190 with 148 (the switch's reduction engines, not the SMs, do the work;
  // more CTAs only add contention and a longer per-CTA tail in the exit barrier).
  if (a->mc_in != nullptr && a->mc_out != nullptr && !getenv("RAYFED_B200_AGG_CTAS_PER_SM") && cap > 32) cap = 32;
  if (a->grid_cap > 0 && a->grid_cap < cap) cap = a->grid_cap;
  long long g = want < cap ? want : cap;
  if (g > 1024) g = 1024;
  return (int)(g < 1 ? 1 : g);
}

cudaError_t fedk_launch_aggregate(const AggArgs* a, int blocks, cudaStream_t stream) {
  const bool mc = a->mc_in != nullptr && a->mc_out != nullptr && !a->out_fp8;
  const int n = a->n_parties;
  const bool state = a->kind != kAvg;
  static const bool lean_off = [] {
    const char* e = getenv("RAYFED_B200_AGG_LEAN");
    return e && e[0] == '0';
  }();
  if (!state && !a->in_fp32 && !a->out_fp8 && !lean_off && !a->force_generic && n <= 8) {
    const int g = lean_grid(a);
    if (mc)          fed_allreduce_avg_kernel<true, 1, 4><<<g, 512, 0, stream>>>(*a);
    else if (n == 1) fed_allreduce_avg_kernel<false, 1, 4><<<g, 512, 0, stream>>>(*a);
    else if (n <= 2) fed_allreduce_avg_kernel<false, 2, 4><<<g, 512, 0, stream>>>(*a);
    else if (n <= 4) fed_allreduce_avg_kernel<false, 4, 2><<<g, 512, 0, stream>>>(*a);
    else             fed_allreduce_avg_kernel<false, 8, 1><<<g, 512, 0, stream>>>(*a);
    return cudaGetLastError();
  }
  // <multicast, preloaded parties per vector, vectors per trip>: kNP*kU 16-byte
  // requests (+ the optimizer state of kU vectors) are in flight per thread.
  if (mc) {
    if (state) fed_aggregate_kernel<true, 1, 2, true><<<blocks, 512, 0, stream>>>(*a);
    else       fed_aggregate_kernel<true, 1, 8, false><<<blocks, 512, 0, stream>>>(*a);
  } else if (n == 1) {
    if (state) fed_aggregate_kernel<false, 1, 2, true><<<blocks, 512, 0, stream>>>(*a);
    else       fed_aggregate_kernel<false, 1, 8, false><<<blocks, 512, 0, stream>>>(*a);
  } else if (n <= 2) {
    if (state) fed_aggregate_kernel<false, 2, 1, true><<<blocks, 512, 0, stream>>>(*a);
    else       fed_aggregate_kernel<false, 2, 4, false><<<blocks, 512, 0, stream>>>(*a);
  } else if (n <= 4) {
    if (state) fed_aggregate_kernel<false, 4, 1, true><<<blocks, 512, 0, stream>>>(*a);
    else       fed_aggregate_kernel<false, 4, 2, false><<<blocks, 512, 0, stream>>>(*a);
  } else {
    if (state) fed_aggregate_kernel<false, 8, 1, true><<<blocks, 512, 0, stream>>>(*a);
    else       fed_aggregate_kernel<false, 8, 1, false><<<blocks, 512, 0, stream>>>(*a);
  }
  return cudaGetLastError();
}


#define FEDK_PRELOAD(fn)                                                        \
  do {                                                                          \
    cudaFuncAttributes at_;                                                     \
    cudaError_t e_ = cudaFuncGetAttributes(&at_, fn);                           \
    if (e_ != cudaSuccess) return e_;                                           \
  } while (0)

cudaError_t fedk_preload_aggregate() {
  FEDK_PRELOAD((fed_aggregate_kernel<true, 1, 2, true>));
  FEDK_PRELOAD((fed_aggregate_kernel<true, 1, 8, false>));
  FEDK_PRELOAD((fed_aggregate_kernel<false, 1, 2, true>));
  FEDK_PRELOAD((fed_aggregate_kernel<false, 1, 8, false>));
  FEDK_PRELOAD((fed_aggregate_kernel<false, 2, 1, true>));
  FEDK_PRELOAD((fed_aggregate_kernel<false, 2, 4, false>));
  FEDK_PRELOAD((fed_aggregate_kernel<false, 4, 1, true>));
  FEDK_PRELOAD((fed_aggregate_kernel<false, 4, 2, false>));
  FEDK_PRELOAD((fed_aggregate_kernel<false, 8, 1, true>));
  FEDK_PRELOAD((fed_aggregate_kernel<false, 8, 1, false>));
  FEDK_PRELOAD((fed_allreduce_avg_kernel<true, 1, 4>));
  FEDK_PRELOAD((fed_allreduce_avg_kernel<false, 1, 4>));
  FEDK_PRELOAD((fed_allreduce_avg_kernel<false, 2, 4>));
  FEDK_PRELOAD((fed_allreduce_avg_kernel<false, 4, 2>));
  FEDK_PRELOAD((fed_allreduce_avg_kernel<false, 8, 1>));
  FEDK_PRELOAD(fed_signal_kernel);
  FEDK_PRELOAD(fed_barrier_kernel);
  FEDK_PRELOAD(fed_local_sgd_kernel);
  return cudaSuccess;
}

cudaError_t fedk_launch_signal(const PeerPtrs* arrive, int n_parties, int my_rank, uint32_t value,
                               cudaStream_t stream) {
  fed_signal_kernel<<<1, 32, 0, stream>>>(*arrive, n_parties, my_rank, value);
  return cudaGetLastError();
}

cudaError_t fedk_launch_barrier(const AggArgs* a, cudaStream_t stream) {
  fed_barrier_kernel<<<1, 32, 0, stream>>>(*a);
  return cudaGetLastError();
}

cudaError_t fedk_launch_local_sgd(const LocalSgdArgs* a, int blocks, cudaStream_t stream) {
  fed_local_sgd_kernel<<<blocks, 512, 0, stream>>>(*a);
  return cudaGetLastError();
}

}  // extern "C"

}  // namespace fedk
