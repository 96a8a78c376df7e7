// Argument blocks shared by the kernels (.cu) and the Python bindings.
#pragma once
#include "common.cuh"

namespace fedk {

enum PushMode : int {
  kCopy = 0,         // byte copy
  kF32ToBF16 = 1,    // cast (x * scale)
  kBF16ToFP8 = 2,    // block-scaled e4m3, one UE8M0 scale byte per 32 elements
  kBF16Scale = 3,    // bf16 * scale -> bf16
  kFP8ToBF16 = 4,    // unpack of mode 2
  kBF16ToF32 = 5,    // unpack of mode 1
};

// Fixed-layout 64-byte message descriptor, written by the push kernel into the first bytes of
// the message's block in the *consumer's* arena, before the release flag (replaces
// the protobuf SendDataRequest{data, upstream_seq_id, downstream_seq_id, job_name} of
// fed/grpc/fed.proto:9-14 on the device path).  The consumer validates it on its own stream
// (fed_desc_check / fed_unpack): a mismatch means diverged controllers or a foreign job and
// raises a status word instead of handing out foreign bytes.
struct alignas(16) PushDesc {
  uint32_t magic;      // kDescMagic
  uint32_t seq;        // k-th message of this (source -> destination) pair == release-flag value
  uint32_t job_hash;   // fnv1a32(job name): the 417 "JobName mis-match" check of the gRPC path
  uint32_t up_task;    // upstream fed task id   } the (upstream_seq_id, downstream_seq_id) key of
  uint32_t up_idx;     // upstream output index  } the reference's event table
  uint32_t down;       // downstream seq id      }
  unsigned long long offset;  // block offset inside the source's region of the arena
  unsigned long long nbytes;  // block bytes (header + payload, wire format)
  unsigned char dtype, mode, ndim, status;  // status: 1 = DATA, 2 = ERROR
  uint32_t n_tensors;  // tensors packed in this block
  uint32_t shape[4];   // leading dims of the first tensor (the host frame has the full metadata)
};
static_assert(sizeof(PushDesc) == 64, "PushDesc must be 64 bytes");
constexpr uint32_t kDescMagic = 0x46454450u;  // "FEDP"
constexpr int kBlockHeaderBytes = 256;        // descriptor + padding: payload stays 256-byte aligned

struct PushArgs {
  const void* src;
  void* dst;
  void* dst_scales;         // mode 2: peer scales region; mode 4: local scales region (src side)
  unsigned long long n;     // bytes for kCopy, elements otherwise
  float scale;
  uint32_t* flag;           // push: peer flag (release-stored at the end); unpack: local flag (acquired first)
  uint32_t flag_value;
  uint32_t* done_counter;   // local, zero between launches
  unsigned long long* credit;  // unpack: producer's credit word in peer memory (may be null)
  unsigned long long credit_value;
  uint32_t* status;         // local error word (timeout)
  long long timeout_ns;
  void* desc_dst;           // push: where the descriptor goes in peer memory (null: none)
  PushDesc desc;            // push: written before the flag; unpack/check: the expected descriptor
};

// Multi-tensor push: one launch moves every tensor of an object (a state dict, a pytree) into
// its block of the consumer's arena.  Segments ride in the kernel parameters (no table H2D).
constexpr int kMaxSegs = 48;
struct PushSegs {
  int n;
  int pad_;
  struct Seg {
    const void* src;
    unsigned long long dst_off;  // relative to PushArgs::dst
    unsigned long long nbytes;
  } s[kMaxSegs];
};

enum AggKind : int {
  kAvg = 0,        // w <- mean_i x_i
  kFedAvgM = 1,    // d = master - mean x ; m = mu m + d ; master -= lr m
  kSgdGrad = 2,    // g = mean x + wd master ; m = mu m + g ; master -= lr m
  kAdamGrad = 3,   // AdamW on g = mean x
  kFedAdam = 4,    // AdamW on d = master - mean x
};

struct AggArgs {
  PeerPtrs in;             // contribution buffers of every party (peer-mapped)
  PeerPtrs out;            // weight buffers of every party (may alias `in`)
  PeerPtrs out_scales;     // fp8 output: UE8M0 scale bytes of every party
  void* mc_in;             // multicast alias of `in`  (nullptr -> P2P loads)
  void* mc_out;            // multicast alias of `out` (nullptr -> P2P stores)
  float* master;           // local fp32 master shard, indexed from 0
  float* mom;              // local fp32 first moment shard
  float* var;              // local fp32 second moment shard (Adam)
  unsigned long long shard_start;  // elements, multiple of 32
  unsigned long long shard_len;    // elements, multiple of 8 (32 for fp8 out)
  int n_parties;
  int my_rank;
  int kind;
  int in_fp32;             // contributions are fp32 (else bf16)
  int out_fp8;             // write block-scaled e4m3 (else bf16)
  float w[kMaxParties];    // aggregation weights (1/N or n_k/n)
  float lr, momentum, weight_decay, beta1, beta2, eps, bias_c1, bias_c2;
  uint32_t* sig;           // my signal pad
  PeerPtrs sig_peers;      // every party's signal pad
  uint32_t epoch;          // this launch uses epoch+1 (entry) and epoch+2 (exit)
  uint32_t* epoch_dev;     // if set: epoch is read from this device word and advanced by 2 by the
                           // last CTA, so an identical launch can be replayed from a CUDA graph
  uint32_t* agg_done;      // device counter for "last CTA" detection (zero between launches)
  uint32_t* status;
  long long timeout_ns;
  int entry_ordered;             // 1: a pre-launch rendezvous already ordered the parties (skip the entry barrier)
  int force_generic;             // experiments: 1 = never take the pipelined FedAvg kernel
  int grid_cap;                  // experiments: upper bound on the pipelined kernel's grid (0: 2 CTAs per SM)
  unsigned long long* phase_ts;  // debug: per CTA {start, entry barrier done, loop done, exit barrier done}
                                 // in %globaltimer ns (nullptr: off)
};

struct LocalSgdArgs {
  __nv_bfloat16* w;
  __nv_bfloat16* g;
  float* mom;
  unsigned long long n;  // multiple of 8
  float lr, momentum, weight_decay, grad_scale;
  int zero_grad;
};


// K3 GEMM: W <- W_base + scale * sum_i w_i * B_i A_i with operands gathered from peers.
struct LoraMergeArgs {
  PeerPtrs b;    // B_i  [m, r] bf16, row-major, one per party (peer-mapped)
  PeerPtrs at;   // A_i^T [n, r] bf16, row-major, one per party (peer-mapped)
  const void* w_base;                  // [m, n] bf16, or e4m3 when base_fp8 (nullptr: dW only)
  const unsigned char* w_base_scales;  // UE8M0 per 32 along n (base_fp8)
  void* out;                           // [m, n] bf16, or e4m3 when out_fp8
  unsigned char* out_scales;
  unsigned long long m, n;
  int r, n_parties;
  float scale;
  float w[kMaxParties];  // per-party weights applied to B_i while gathering (1.0 = none)
  int base_fp8, out_fp8;
};

extern "C" {
cudaError_t fedk_launch_lora_merge(const LoraMergeArgs* a, cudaStream_t stream);
cudaError_t fedk_launch_lora_merge_tma(const LoraMergeArgs* a, const void* tma_maps, cudaStream_t stream);
cudaError_t fedk_launch_lora_merge_staged(const LoraMergeArgs* a, const void* staged_maps, void* bcat, void* atcat,
                                          int kpad, int sm_count, int deep, cudaStream_t stream);
cudaError_t fedk_launch_push(const PushArgs* a, int mode, int blocks, int use_bulk, cudaStream_t stream);
cudaError_t fedk_launch_push_mc(const PushArgs* a, const PeerPtrs* flags, int n_flags, int blocks, cudaStream_t stream);
cudaError_t fedk_launch_unpack(const PushArgs* a, int mode, int blocks, cudaStream_t stream);
cudaError_t fedk_launch_push_multi(const PushArgs* a, const PushSegs* segs, int blocks, cudaStream_t stream);
cudaError_t fedk_launch_desc_check(const PushArgs* a, cudaStream_t stream);
cudaError_t fedk_launch_credit(unsigned long long* ring_slot, unsigned long long value, unsigned long long* count_ptr,
                               unsigned long long count, cudaStream_t stream);
cudaError_t fedk_preload_push();       // force the (lazy) load of every kernel of a translation unit:
cudaError_t fedk_preload_aggregate();  // a first launch must never happen while a stream wait is parked
cudaError_t fedk_preload_lora();
cudaError_t fedk_launch_aggregate(const AggArgs* a, int blocks, cudaStream_t stream);
cudaError_t fedk_launch_signal(const PeerPtrs* arrive, int n_parties, int my_rank, uint32_t value, cudaStream_t stream);
cudaError_t fedk_launch_barrier(const AggArgs* a, cudaStream_t stream);
cudaError_t fedk_launch_local_sgd(const LocalSgdArgs* a, int blocks, cudaStream_t stream);
}

}  // namespace fedk
