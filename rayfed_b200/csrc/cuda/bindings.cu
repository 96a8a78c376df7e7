// _fedcuda: pybind11 surface of the sm_100a kernels and the symmetric heap.
//
// Pointers and streams cross the boundary as integers (torch gives us
// tensor.data_ptr() and stream.cuda_stream), so this module has no libtorch
// dependency and compiles in seconds.  cudart is linked statically and the
// driver API (cuMem*/cuMulticast* for the VMM + NVLS path) is resolved at run
// time through cudaGetDriverEntryPoint, so importing the module on a machine
// without a GPU or libcuda works — only launching needs a device.
#include <cuda.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../core/device_hooks.h"
#include "kernels.h"

namespace py = pybind11;

using namespace fedk;

namespace {

void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

template <typename T>
T* P(uint64_t v) {
  return reinterpret_cast<T*>(static_cast<uintptr_t>(v));
}

void fill_peers(PeerPtrs& pp, const std::vector<uint64_t>& v) {
  if (v.size() > (size_t)kMaxParties) throw std::runtime_error("too many parties");
  for (int i = 0; i < kMaxParties; ++i) pp.p[i] = i < (int)v.size() ? P<void>(v[i]) : nullptr;
}

// ---- driver entry points -----------------------------------------------------
template <typename Fn>
Fn drv(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr)
    throw std::runtime_error(std::string("driver entry point not available: ") + name);
  return reinterpret_cast<Fn>(fn);
}

void cu_check(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) {
    const char* s = nullptr;
    try {
      auto f = drv<CUresult (*)(CUresult, const char**)>("cuGetErrorString");
      f(r, &s);
    } catch (...) {
    }
    throw std::runtime_error(std::string(what) + ": CUresult " + std::to_string((int)r) + " " + (s ? s : ""));
  }
}

size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

void fill_desc(PushArgs& a, uint64_t desc_dst, const std::string& desc) {
  a.desc_dst = P<void>(desc_dst);
  if (!desc.empty()) {
    if (desc.size() != sizeof(PushDesc)) throw std::runtime_error("descriptor must be 64 bytes");
    std::memcpy(&a.desc, desc.data(), sizeof(PushDesc));
  }
}

// ---- the scheduler's CUDA calls (csrc/core/device_hooks.h) ---------------------------------
const fedcore::DeviceHooks* cuda_device_hooks() {
  static fedcore::DeviceHooks h = {
      [](uint64_t* ev) -> int {
        cudaEvent_t e = nullptr;
        cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        *ev = (uint64_t)(uintptr_t)e;
        return (int)r;
      },
      [](uint64_t ev) -> int { return (int)cudaEventDestroy((cudaEvent_t)(uintptr_t)ev); },
      [](uint64_t ev, uint64_t stream) -> int {
        return (int)cudaEventRecord((cudaEvent_t)(uintptr_t)ev, (cudaStream_t)(uintptr_t)stream);
      },
      [](uint64_t stream, uint64_t ev) -> int {
        return (int)cudaStreamWaitEvent((cudaStream_t)(uintptr_t)stream, (cudaEvent_t)(uintptr_t)ev, 0);
      },
      [](uint64_t ev) -> int { return (int)cudaEventSynchronize((cudaEvent_t)(uintptr_t)ev); },
      [](uint64_t ev) -> int {
        cudaError_t r = cudaEventQuery((cudaEvent_t)(uintptr_t)ev);
        if (r == cudaErrorNotReady) {
          cudaGetLastError();
          return 1;
        }
        return r == cudaSuccess ? 0 : (int)r;
      },
  };
  return &h;
}

// [rows, cols] row-major bf16 matrix -> TMA descriptor with a {64 cols, 128 rows} box, SWIZZLE_128B.
CUtensorMap make_bf16_tile_map(uint64_t ptr, uint64_t rows, uint64_t cols) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = drv<EncodeFn>("cuTensorMapEncodeTiled");
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  cu_check(encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, P<void>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE),
           "cuTensorMapEncodeTiled");
  return m;
}

}  // namespace

PYBIND11_MODULE(_fedcuda, m) {
  m.doc() = "rayfed_b200 sm_100a kernels + symmetric heap";
  m.attr("MAX_PARTIES") = (int)kMaxParties;

  // ---- cudaIpc symmetric memory (tier A: works everywhere P2P works) ---------
  m.def("ipc_alloc", [](size_t nbytes, bool zero) {
    void* p = nullptr;
    check(cudaMalloc(&p, nbytes), "cudaMalloc");
    if (zero) check(cudaMemset(p, 0, nbytes), "cudaMemset");
    cudaIpcMemHandle_t h;
    check(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
    return py::make_tuple((uint64_t)(uintptr_t)p, py::bytes(reinterpret_cast<const char*>(&h), sizeof(h)));
  }, py::arg("nbytes"), py::arg("zero") = true);
  m.def("ipc_open", [](py::bytes handle) {
    std::string s = handle;
    if (s.size() != sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("bad ipc handle size");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, s.data(), sizeof(h));
    void* p = nullptr;
    check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return (uint64_t)(uintptr_t)p;
  });
  m.def("ipc_close", [](uint64_t p) { check(cudaIpcCloseMemHandle(P<void>(p)), "cudaIpcCloseMemHandle"); });
  m.def("dev_free", [](uint64_t p) { check(cudaFree(P<void>(p)), "cudaFree"); });
  m.def("dev_malloc", [](size_t nbytes, bool zero) {
    void* p = nullptr;
    check(cudaMalloc(&p, nbytes), "cudaMalloc");
    if (zero) check(cudaMemset(p, 0, nbytes), "cudaMemset");
    return (uint64_t)(uintptr_t)p;
  }, py::arg("nbytes"), py::arg("zero") = true);
  m.def("enable_peer_access", [](int peer) {
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
      cudaGetLastError();
      return true;
    }
    return e == cudaSuccess;
  });
  m.def("can_access_peer", [](int dev, int peer) {
    int ok = 0;
    check(cudaDeviceCanAccessPeer(&ok, dev, peer), "cudaDeviceCanAccessPeer");
    return ok != 0;
  });
  m.def("set_device", [](int d) { check(cudaSetDevice(d), "cudaSetDevice"); });
  m.def("current_device", [] {
    int d = -1;
    check(cudaGetDevice(&d), "cudaGetDevice");
    return d;
  });
  m.def("device_synchronize", [] {
    py::gil_scoped_release rel;
    check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  });
  m.def("sm_count", [](int dev) {
    int n = 0;
    check(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
    return n;
  });
  // small synchronous host<->device words (credit polling, flag reset)
  m.def("read_u64", [](uint64_t p) {
    unsigned long long v = 0;
    check(cudaMemcpy(&v, P<void>(p), 8, cudaMemcpyDeviceToHost), "cudaMemcpy");
    return (uint64_t)v;
  });
  m.def("read_u32", [](uint64_t p) {
    uint32_t v = 0;
    check(cudaMemcpy(&v, P<void>(p), 4, cudaMemcpyDeviceToHost), "cudaMemcpy");
    return v;
  });
  m.def("write_u32", [](uint64_t p, uint32_t v) {
    check(cudaMemcpy(P<void>(p), &v, 4, cudaMemcpyHostToDevice), "cudaMemcpy");
  });
  m.def("memset", [](uint64_t p, int byte, size_t n) { check(cudaMemset(P<void>(p), byte, n), "cudaMemset"); });

  // ---- streams, events, stream memory operations ---------------------------------
  m.def("device_hooks", [] {
    return py::capsule(const_cast<fedcore::DeviceHooks*>(cuda_device_hooks()), "rayfed_b200.device_hooks");
  });
  m.def("preload_kernels", [] {
    check(fedk_preload_push(), "preload push kernels");
    check(fedk_preload_aggregate(), "preload aggregate kernels");
    check(fedk_preload_lora(), "preload lora kernels");
  }, "Force the (lazy) load of every kernel of this module: a first launch must never happen "
     "while a stream wait is parked (it would block the launching thread until the wait ends).");
  m.def("module_loading_mode", [] {
    using Fn = CUresult (*)(CUmoduleLoadingMode*);
    CUmoduleLoadingMode mode;
    cu_check(drv<Fn>("cuModuleGetLoadingMode")(&mode), "cuModuleGetLoadingMode");
    return std::string(mode == CU_MODULE_EAGER_LOADING ? "eager" : "lazy");
  });
  m.def("stream_create", [](int priority) {
    cudaStream_t s = nullptr;
    check(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, priority), "cudaStreamCreateWithPriority");
    return (uint64_t)(uintptr_t)s;
  }, py::arg("priority") = 0);
  m.def("stream_destroy", [](uint64_t s) { check(cudaStreamDestroy((cudaStream_t)(uintptr_t)s), "cudaStreamDestroy"); });
  m.def("stream_synchronize", [](uint64_t s) {
    py::gil_scoped_release rel;
    check(cudaStreamSynchronize((cudaStream_t)(uintptr_t)s), "cudaStreamSynchronize");
  });
  m.def("event_create", [](bool timing) {
    cudaEvent_t e = nullptr;
    check(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming), "cudaEventCreate");
    return (uint64_t)(uintptr_t)e;
  }, py::arg("timing") = false);
  m.def("event_destroy", [](uint64_t e) { check(cudaEventDestroy((cudaEvent_t)(uintptr_t)e), "cudaEventDestroy"); });
  m.def("event_record", [](uint64_t e, uint64_t s) {
    check(cudaEventRecord((cudaEvent_t)(uintptr_t)e, (cudaStream_t)(uintptr_t)s), "cudaEventRecord");
  });
  m.def("event_synchronize", [](uint64_t e) {
    py::gil_scoped_release rel;
    check(cudaEventSynchronize((cudaEvent_t)(uintptr_t)e), "cudaEventSynchronize");
  });
  m.def("event_query", [](uint64_t e) {
    cudaError_t r = cudaEventQuery((cudaEvent_t)(uintptr_t)e);
    if (r == cudaErrorNotReady) {
      cudaGetLastError();
      return false;
    }
    check(r, "cudaEventQuery");
    return true;
  });
  m.def("event_elapsed_ms", [](uint64_t a, uint64_t b) {
    float ms = 0.f;
    check(cudaEventElapsedTime(&ms, (cudaEvent_t)(uintptr_t)a, (cudaEvent_t)(uintptr_t)b), "cudaEventElapsedTime");
    return ms;
  });
  m.def("stream_wait_event", [](uint64_t s, uint64_t e) {
    check(cudaStreamWaitEvent((cudaStream_t)(uintptr_t)s, (cudaEvent_t)(uintptr_t)e, 0), "cudaStreamWaitEvent");
  });
  // K4: order a stream after a 32-bit word reaches `value` (cyclic >=) — a stream memory
  // operation: no SM is occupied and no kernel spins while the producer is still pushing.
  m.def("stream_wait_value32", [](uint64_t s, uint64_t addr, uint32_t value) {
    using Fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static Fn fn = drv<Fn>("cuStreamWaitValue32");
    cu_check(fn((CUstream)(uintptr_t)s, (CUdeviceptr)addr, value, CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32");
  });
  m.def("stream_write_value32", [](uint64_t s, uint64_t addr, uint32_t value) {
    using Fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static Fn fn = drv<Fn>("cuStreamWriteValue32");
    cu_check(fn((CUstream)(uintptr_t)s, (CUdeviceptr)addr, value, CU_STREAM_WRITE_VALUE_DEFAULT), "cuStreamWriteValue32");
  });
  // n u64 words, copied on `stream` (never the legacy default stream: that one serialises with
  // whatever the user's main thread is doing)
  m.def("read_u64s", [](uint64_t p, size_t n, uint64_t stream) {
    std::vector<unsigned long long> v(n);
    {
      py::gil_scoped_release rel;
      check(cudaMemcpyAsync(v.data(), P<void>(p), n * 8, cudaMemcpyDeviceToHost, (cudaStream_t)(uintptr_t)stream),
            "cudaMemcpyAsync");
      check(cudaStreamSynchronize((cudaStream_t)(uintptr_t)stream), "cudaStreamSynchronize");
    }
    return std::vector<uint64_t>(v.begin(), v.end());
  });
  m.def("read_bytes", [](uint64_t p, size_t n) {
    std::string out(n, '\0');
    check(cudaMemcpy(out.data(), P<void>(p), n, cudaMemcpyDeviceToHost), "cudaMemcpy");
    return py::bytes(out);
  });

  // ---- CUDA VMM + NVLS multicast (tier B) -------------------------------------
  m.def("vmm_info", [](int dev) {
    py::dict d;
    int v = 0;
    auto attr = drv<CUresult (*)(int*, CUdevice_attribute, CUdevice)>("cuDeviceGetAttribute");
    cu_check(attr(&v, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev), "attr vmm");
    d["vmm"] = v;
    cu_check(attr(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev), "attr fd");
    d["posix_fd"] = v;
    cu_check(attr(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev), "attr mc");
    d["multicast"] = v;
    CUresult r = attr(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev);
    d["fabric"] = r == CUDA_SUCCESS ? v : -1;
    return d;
  });
  // Allocate `nbytes` of device memory with an exportable POSIX-fd handle, map it
  // locally, return (ptr, fd, mapped_size, mem_handle).
  m.def("vmm_alloc", [](size_t nbytes, int dev) {
    auto getGran = drv<CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags)>(
        "cuMemGetAllocationGranularity");
    auto create = drv<CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*,
                                   unsigned long long)>("cuMemCreate");
    auto exportH = drv<CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                    unsigned long long)>("cuMemExportToShareableHandle");
    auto reserve = drv<CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long)>(
        "cuMemAddressReserve");
    auto map = drv<CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long)>(
        "cuMemMap");
    auto setAccess = drv<CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t)>("cuMemSetAccess");
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    cu_check(getGran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
    size_t size = round_up(nbytes, gran);
    CUmemGenericAllocationHandle h;
    cu_check(create(&h, size, &prop, 0), "cuMemCreate");
    int fd = -1;
    cu_check(exportH(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
    CUdeviceptr va = 0;
    cu_check(reserve(&va, size, gran, 0, 0), "cuMemAddressReserve");
    cu_check(map(va, size, 0, h, 0), "cuMemMap");
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    cu_check(setAccess(va, size, &acc, 1), "cuMemSetAccess");
    check(cudaMemset(reinterpret_cast<void*>(va), 0, size), "cudaMemset");
    return py::make_tuple((uint64_t)va, fd, size, (uint64_t)h);
  });
  // Import a peer's fd and map it with read/write access for `dev`.
  m.def("vmm_import", [](int fd, size_t size, int dev) {
    auto importH = drv<CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType)>(
        "cuMemImportFromShareableHandle");
    auto reserve = drv<CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long)>(
        "cuMemAddressReserve");
    auto map = drv<CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long)>(
        "cuMemMap");
    auto setAccess = drv<CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t)>("cuMemSetAccess");
    CUmemGenericAllocationHandle h;
    cu_check(importH(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
             "cuMemImportFromShareableHandle");
    CUdeviceptr va = 0;
    cu_check(reserve(&va, size, 0, 0, 0), "cuMemAddressReserve");
    cu_check(map(va, size, 0, h, 0), "cuMemMap");
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    cu_check(setAccess(va, size, &acc, 1), "cuMemSetAccess");
    return py::make_tuple((uint64_t)va, (uint64_t)h);
  });
  // Multicast object: rank 0 creates + exports; everyone imports, adds its device,
  // binds its own physical allocation and maps the multicast VA.
  m.def("mc_granularity", [](size_t nbytes, int ndev) {
    auto gran = drv<CUresult (*)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags)>(
        "cuMulticastGetGranularity");
    CUmulticastObjectProp prop{};
    prop.numDevices = ndev;
    prop.size = nbytes;
    prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t g = 0;
    cu_check(gran(&g, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
    return g;
  });
  m.def("mc_create", [](size_t size, int ndev) {
    auto create = drv<CUresult (*)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*)>("cuMulticastCreate");
    auto exportH = drv<CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                    unsigned long long)>("cuMemExportToShareableHandle");
    CUmulticastObjectProp prop{};
    prop.numDevices = ndev;
    prop.size = size;
    prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle h;
    cu_check(create(&h, &prop), "cuMulticastCreate");
    int fd = -1;
    cu_check(exportH(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle(mc)");
    return py::make_tuple((uint64_t)h, fd);
  });
  m.def("mc_import", [](int fd) {
    auto importH = drv<CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType)>(
        "cuMemImportFromShareableHandle");
    CUmemGenericAllocationHandle h;
    cu_check(importH(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
             "cuMemImportFromShareableHandle(mc)");
    return (uint64_t)h;
  });
  m.def("mc_add_device", [](uint64_t mc, int dev) {
    auto add = drv<CUresult (*)(CUmemGenericAllocationHandle, CUdevice)>("cuMulticastAddDevice");
    cu_check(add((CUmemGenericAllocationHandle)mc, dev), "cuMulticastAddDevice");
  });
  m.def("mc_bind_and_map", [](uint64_t mc, uint64_t mem_handle, size_t size, int dev) {
    auto bind = drv<CUresult (*)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                                 unsigned long long)>("cuMulticastBindMem");
    auto reserve = drv<CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long)>(
        "cuMemAddressReserve");
    auto map = drv<CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long)>(
        "cuMemMap");
    auto setAccess = drv<CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t)>("cuMemSetAccess");
    cu_check(bind((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)mem_handle, 0, size, 0),
             "cuMulticastBindMem");
    CUdeviceptr va = 0;
    cu_check(reserve(&va, size, 0, 0, 0), "cuMemAddressReserve(mc)");
    cu_check(map(va, size, 0, (CUmemGenericAllocationHandle)mc, 0), "cuMemMap(mc)");
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    cu_check(setAccess(va, size, &acc, 1), "cuMemSetAccess(mc)");
    return (uint64_t)va;
  });

  // ---- kernels -------------------------------------------------------------------
  m.def(
      "push",
      [](uint64_t src, uint64_t dst, uint64_t dst_scales, uint64_t n, float scale, uint64_t flag, uint32_t flag_value,
         uint64_t done_counter, int mode, int blocks, bool use_bulk, uint64_t stream, uint64_t desc_dst,
         const std::string& desc) {
        PushArgs a{};
        a.src = P<const void>(src);
        a.dst = P<void>(dst);
        a.dst_scales = P<void>(dst_scales);
        a.n = n;
        a.scale = scale;
        a.flag = P<uint32_t>(flag);
        a.flag_value = flag_value;
        a.done_counter = P<uint32_t>(done_counter);
        fill_desc(a, desc_dst, desc);
        check(fedk_launch_push(&a, mode, blocks, use_bulk ? 1 : 0, (cudaStream_t)(uintptr_t)stream), "fed_push launch");
      },
      py::arg("src"), py::arg("dst"), py::arg("dst_scales"), py::arg("n"), py::arg("scale"), py::arg("flag"),
      py::arg("flag_value"), py::arg("done_counter"), py::arg("mode"), py::arg("blocks"), py::arg("use_bulk"),
      py::arg("stream"), py::arg("desc_dst") = 0, py::arg("desc") = py::bytes());
  m.def(
      "push_multi",
      [](const std::vector<std::tuple<uint64_t, uint64_t, uint64_t>>& segs, uint64_t dst, uint64_t flag,
         uint32_t flag_value, uint64_t done_counter, int blocks, uint64_t stream, uint64_t desc_dst,
         const std::string& desc) {
        if (segs.empty() || segs.size() > (size_t)kMaxSegs) throw std::runtime_error("push_multi: 1..48 segments");
        PushSegs sg{};
        sg.n = (int)segs.size();
        for (int i = 0; i < sg.n; ++i) {
          sg.s[i].src = P<const void>(std::get<0>(segs[i]));
          sg.s[i].dst_off = std::get<1>(segs[i]);
          sg.s[i].nbytes = std::get<2>(segs[i]);
          if ((std::get<0>(segs[i]) | std::get<1>(segs[i])) & 15) throw std::runtime_error("push_multi: 16-byte alignment");
        }
        PushArgs a{};
        a.dst = P<void>(dst);
        a.flag = P<uint32_t>(flag);
        a.flag_value = flag_value;
        a.done_counter = P<uint32_t>(done_counter);
        fill_desc(a, desc_dst, desc);
        check(fedk_launch_push_multi(&a, &sg, blocks, (cudaStream_t)(uintptr_t)stream), "fed_push_multi launch");
      },
      py::arg("segs"), py::arg("dst"), py::arg("flag"), py::arg("flag_value"), py::arg("done_counter"),
      py::arg("blocks"), py::arg("stream"), py::arg("desc_dst") = 0, py::arg("desc") = py::bytes());
  m.def(
      "desc_check",
      [](uint64_t arena_desc, const std::string& want, uint64_t status, uint64_t stream) {
        PushArgs a{};
        a.src = P<const void>(arena_desc);
        a.status = P<uint32_t>(status);
        fill_desc(a, 0, want);
        check(fedk_launch_desc_check(&a, (cudaStream_t)(uintptr_t)stream), "fed_desc_check launch");
      },
      py::arg("arena_desc"), py::arg("want"), py::arg("status"), py::arg("stream"));
  m.def(
      "credit",
      [](uint64_t ring_slot, uint64_t value, uint64_t count_ptr, uint64_t count, uint64_t stream) {
        check(fedk_launch_credit(P<unsigned long long>(ring_slot), value, P<unsigned long long>(count_ptr), count,
                                 (cudaStream_t)(uintptr_t)stream),
              "fed_credit launch");
      },
      py::arg("ring_slot"), py::arg("value"), py::arg("count_ptr"), py::arg("count"), py::arg("stream"));
  m.def(
      "push_mc",
      [](uint64_t src, uint64_t mc_dst, uint64_t nbytes, const std::vector<uint64_t>& flag_ptrs, uint32_t flag_value,
         uint64_t done_counter, int blocks, uint64_t stream, uint64_t desc_dst, const std::string& desc) {
        if (nbytes % 16) throw std::runtime_error("push_mc needs a multiple of 16 bytes");
        PushArgs a{};
        a.src = P<const void>(src);
        a.dst = P<void>(mc_dst);
        a.n = nbytes;
        a.flag_value = flag_value;
        a.done_counter = P<uint32_t>(done_counter);
        fill_desc(a, desc_dst, desc);
        PeerPtrs flags{};
        fill_peers(flags, flag_ptrs);
        check(fedk_launch_push_mc(&a, &flags, (int)flag_ptrs.size(), blocks, (cudaStream_t)(uintptr_t)stream),
              "fed_push_mc launch");
      },
      py::arg("src"), py::arg("mc_dst"), py::arg("nbytes"), py::arg("flag_ptrs"), py::arg("flag_value"),
      py::arg("done_counter"), py::arg("blocks"), py::arg("stream"), py::arg("desc_dst") = 0,
      py::arg("desc") = py::bytes());
  m.def(
      "unpack",
      [](uint64_t src, uint64_t dst, uint64_t scales, uint64_t n, float scale, uint64_t flag, uint32_t flag_value,
         uint64_t done_counter, uint64_t credit, uint64_t credit_value, uint64_t status, long long timeout_ns,
         int mode, int blocks, uint64_t stream) {
        PushArgs a{};
        a.src = P<const void>(src);
        a.dst = P<void>(dst);
        a.dst_scales = P<void>(scales);
        a.n = n;
        a.scale = scale;
        a.flag = P<uint32_t>(flag);
        a.flag_value = flag_value;
        a.done_counter = P<uint32_t>(done_counter);
        a.credit = P<unsigned long long>(credit);
        a.credit_value = credit_value;
        a.status = P<uint32_t>(status);
        a.timeout_ns = timeout_ns;
        check(fedk_launch_unpack(&a, mode, blocks, (cudaStream_t)(uintptr_t)stream), "fed_unpack launch");
      },
      py::arg("src"), py::arg("dst"), py::arg("scales"), py::arg("n"), py::arg("scale"), py::arg("flag"),
      py::arg("flag_value"), py::arg("done_counter"), py::arg("credit"), py::arg("credit_value"), py::arg("status"),
      py::arg("timeout_ns"), py::arg("mode"), py::arg("blocks"), py::arg("stream"));

  m.def(
      "aggregate",
      [](const std::vector<uint64_t>& in, const std::vector<uint64_t>& out, const std::vector<uint64_t>& out_scales,
         uint64_t mc_in, uint64_t mc_out, uint64_t master, uint64_t mom, uint64_t var, uint64_t shard_start,
         uint64_t shard_len, int my_rank, int kind, bool in_fp32, bool out_fp8, const std::vector<float>& weights,
         float lr, float momentum, float weight_decay, float beta1, float beta2, float eps, float bias_c1,
         float bias_c2, uint64_t sig, const std::vector<uint64_t>& sig_peers, uint32_t epoch, uint64_t status,
         long long timeout_ns, int blocks, uint64_t stream) {
        AggArgs a{};
        fill_peers(a.in, in);
        fill_peers(a.out, out);
        fill_peers(a.out_scales, out_scales);
        a.mc_in = P<void>(mc_in);
        a.mc_out = P<void>(mc_out);
        a.master = P<float>(master);
        a.mom = P<float>(mom);
        a.var = P<float>(var);
        a.shard_start = shard_start;
        a.shard_len = shard_len;
        a.n_parties = (int)in.size();
        a.my_rank = my_rank;
        a.kind = kind;
        a.in_fp32 = in_fp32;
        a.out_fp8 = out_fp8;
        for (int i = 0; i < kMaxParties; ++i) a.w[i] = i < (int)weights.size() ? weights[i] : 0.f;
        a.lr = lr;
        a.momentum = momentum;
        a.weight_decay = weight_decay;
        a.beta1 = beta1;
        a.beta2 = beta2;
        a.eps = eps;
        a.bias_c1 = bias_c1;
        a.bias_c2 = bias_c2;
        a.sig = P<uint32_t>(sig);
        fill_peers(a.sig_peers, sig_peers);
        a.epoch = epoch;
        a.status = P<uint32_t>(status);
        a.timeout_ns = timeout_ns;
        check(fedk_launch_aggregate(&a, blocks, (cudaStream_t)(uintptr_t)stream), "fed_aggregate launch");
      },
      py::arg("inputs"), py::arg("outputs"), py::arg("out_scales"), py::arg("mc_in"), py::arg("mc_out"),
      py::arg("master"), py::arg("mom"), py::arg("var"), py::arg("shard_start"), py::arg("shard_len"),
      py::arg("my_rank"), py::arg("kind"), py::arg("in_fp32"), py::arg("out_fp8"), py::arg("weights"), py::arg("lr"),
      py::arg("momentum"), py::arg("weight_decay"), py::arg("beta1"), py::arg("beta2"), py::arg("eps"),
      py::arg("bias_c1"), py::arg("bias_c2"), py::arg("sig"), py::arg("sig_peers"), py::arg("epoch"),
      py::arg("status"), py::arg("timeout_ns"), py::arg("blocks"), py::arg("stream"));

  // Pre-baked argument block: per-round launches only patch the scalars that change.
  struct AggPlan {
    AggArgs a;
    int blocks;
  };
  py::class_<AggPlan>(m, "AggPlan")
      .def(py::init([](const std::vector<uint64_t>& in, const std::vector<uint64_t>& out,
                       const std::vector<uint64_t>& out_scales, uint64_t mc_in, uint64_t mc_out, uint64_t master,
                       uint64_t mom, uint64_t var, uint64_t shard_start, uint64_t shard_len, int my_rank, int kind,
                       bool in_fp32, bool out_fp8, const std::vector<float>& weights, float lr, float momentum,
                       float weight_decay, float beta1, float beta2, float eps, uint64_t sig,
                       const std::vector<uint64_t>& sig_peers, uint64_t epoch_dev, uint64_t agg_done,
                       uint64_t status, long long timeout_ns, int blocks) {
             AggPlan p{};
             AggArgs& a = p.a;
             fill_peers(a.in, in);
             fill_peers(a.out, out);
             fill_peers(a.out_scales, out_scales);
             a.mc_in = P<void>(mc_in);
             a.mc_out = P<void>(mc_out);
             a.master = P<float>(master);
             a.mom = P<float>(mom);
             a.var = P<float>(var);
             a.shard_start = shard_start;
             a.shard_len = shard_len;
             a.n_parties = (int)in.size();
             a.my_rank = my_rank;
             a.kind = kind;
             a.in_fp32 = in_fp32;
             a.out_fp8 = out_fp8;
             for (int i = 0; i < kMaxParties; ++i) a.w[i] = i < (int)weights.size() ? weights[i] : 0.f;
             a.lr = lr;
             a.momentum = momentum;
             a.weight_decay = weight_decay;
             a.beta1 = beta1;
             a.beta2 = beta2;
             a.eps = eps;
             a.bias_c1 = 1.f;
             a.bias_c2 = 1.f;
             a.sig = P<uint32_t>(sig);
             fill_peers(a.sig_peers, sig_peers);
             a.epoch_dev = P<uint32_t>(epoch_dev);
             a.agg_done = P<uint32_t>(agg_done);
             a.status = P<uint32_t>(status);
             a.timeout_ns = timeout_ns;
             p.blocks = blocks;
             return p;
           }),
           py::arg("inputs"), py::arg("outputs"), py::arg("out_scales"), py::arg("mc_in"), py::arg("mc_out"),
           py::arg("master"), py::arg("mom"), py::arg("var"), py::arg("shard_start"), py::arg("shard_len"),
           py::arg("my_rank"), py::arg("kind"), py::arg("in_fp32"), py::arg("out_fp8"), py::arg("weights"),
           py::arg("lr"), py::arg("momentum"), py::arg("weight_decay"), py::arg("beta1"), py::arg("beta2"),
           py::arg("eps"), py::arg("sig"), py::arg("sig_peers"), py::arg("epoch_dev"), py::arg("agg_done"),
           py::arg("status"), py::arg("timeout_ns"), py::arg("blocks"))
      .def("set_entry_ordered", [](AggPlan& p, bool v) { p.a.entry_ordered = v ? 1 : 0; })
      .def("set_force_generic", [](AggPlan& p, bool v) { p.a.force_generic = v ? 1 : 0; })
      .def("set_grid_cap", [](AggPlan& p, int n) { p.a.grid_cap = n; })
      .def("set_phase_ts", [](AggPlan& p, uint64_t ptr) { p.a.phase_ts = P<unsigned long long>(ptr); },
           "debug: device buffer of 4 x u64 per CTA receiving %globaltimer stamps {start, entry barrier "
           "done, loop done, exit barrier done}; 0 turns it off")
      .def("launch",
           [](AggPlan& p, uint32_t epoch, float lr, float bias_c1, float bias_c2, uint64_t stream) {
             p.a.epoch = epoch;
             p.a.lr = lr;
             p.a.bias_c1 = bias_c1;
             p.a.bias_c2 = bias_c2;
             check(fedk_launch_aggregate(&p.a, p.blocks, (cudaStream_t)(uintptr_t)stream), "fed_aggregate launch");
           },
           py::arg("epoch"), py::arg("lr"), py::arg("bias_c1"), py::arg("bias_c2"), py::arg("stream"));

  m.def(
      "signal_peers",
      [](const std::vector<uint64_t>& arrive_ptrs, int my_rank, uint32_t value, uint64_t stream) {
        PeerPtrs pp{};
        fill_peers(pp, arrive_ptrs);
        check(fedk_launch_signal(&pp, (int)arrive_ptrs.size(), my_rank, value, (cudaStream_t)(uintptr_t)stream),
              "fed_signal launch");
      },
      py::arg("arrive_ptrs"), py::arg("my_rank"), py::arg("value"), py::arg("stream"));
  m.def(
      "barrier",
      [](int n_parties, int my_rank, uint64_t sig, const std::vector<uint64_t>& sig_peers, uint32_t epoch,
         uint64_t status, long long timeout_ns, uint64_t stream) {
        AggArgs a{};
        a.n_parties = n_parties;
        a.my_rank = my_rank;
        a.sig = P<uint32_t>(sig);
        fill_peers(a.sig_peers, sig_peers);
        a.epoch = epoch;
        a.status = P<uint32_t>(status);
        a.timeout_ns = timeout_ns;
        check(fedk_launch_barrier(&a, (cudaStream_t)(uintptr_t)stream), "fed_barrier launch");
      },
      py::arg("n_parties"), py::arg("my_rank"), py::arg("sig"), py::arg("sig_peers"), py::arg("epoch"),
      py::arg("status"), py::arg("timeout_ns"), py::arg("stream"));

  m.def(
      "lora_merge",
      [](const std::vector<uint64_t>& b, const std::vector<uint64_t>& at, uint64_t w_base, uint64_t w_base_scales,
         uint64_t out, uint64_t out_scales, uint64_t m_rows, uint64_t n_cols, int r, float scale,
         const std::vector<float>& weights, bool base_fp8, bool out_fp8, uint64_t stream, bool use_tma,
         uint64_t bcat, uint64_t atcat, int kpad, bool deep) {
        LoraMergeArgs a{};
        fill_peers(a.b, b);
        fill_peers(a.at, at);
        a.w_base = P<const void>(w_base);
        a.w_base_scales = P<const unsigned char>(w_base_scales);
        a.out = P<void>(out);
        a.out_scales = P<unsigned char>(out_scales);
        a.m = m_rows;
        a.n = n_cols;
        a.r = r;
        a.n_parties = (int)b.size();
        a.scale = scale;
        for (int i = 0; i < kMaxParties; ++i) a.w[i] = i < (int)weights.size() ? weights[i] : 1.f;
        a.base_fp8 = base_fp8;
        a.out_fp8 = out_fp8;
        if (bcat && atcat) {
          // staged persistent path: gather once, then TMA -> tcgen05 -> TMA
          if (base_fp8 || out_fp8 || n_cols % 8 || out % 16 || w_base % 16 || bcat % 16 || atcat % 16)
            throw std::runtime_error("lora_merge(staged): needs bf16 base/out, n % 8 == 0 and aligned buffers");
          struct alignas(64) Maps {
            CUtensorMap w;
            CUtensorMap out;
            CUtensorMap bcat;
            CUtensorMap atcat;
          } maps;
          maps.out = make_bf16_tile_map(out, m_rows, n_cols);
          maps.w = w_base ? make_bf16_tile_map(w_base, m_rows, n_cols) : maps.out;
          maps.bcat = make_bf16_tile_map(bcat, m_rows, (uint64_t)kpad);
          maps.atcat = make_bf16_tile_map(atcat, n_cols, (uint64_t)kpad);
          int dev = 0, sms = 0;
          check(cudaGetDevice(&dev), "cudaGetDevice");
          check(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "sm count");
          check(fedk_launch_lora_merge_staged(&a, &maps, P<void>(bcat), P<void>(atcat), kpad, sms, deep ? 1 : 0,
                                              (cudaStream_t)(uintptr_t)stream),
                "fed_lora_merge_staged launch");
        } else if (use_tma && !base_fp8 && !out_fp8 && n_cols % 64 == 0 && (out % 16 == 0) && (w_base % 16 == 0)) {
          struct alignas(64) Maps {
            CUtensorMap w;
            CUtensorMap out;
          } maps;
          maps.out = make_bf16_tile_map(out, m_rows, n_cols);
          maps.w = w_base ? make_bf16_tile_map(w_base, m_rows, n_cols) : maps.out;
          check(fedk_launch_lora_merge_tma(&a, &maps, (cudaStream_t)(uintptr_t)stream), "fed_lora_merge_tma launch");
        } else {
          check(fedk_launch_lora_merge(&a, (cudaStream_t)(uintptr_t)stream), "fed_lora_merge launch");
        }
      },
      py::arg("b"), py::arg("at"), py::arg("w_base"), py::arg("w_base_scales"), py::arg("out"),
      py::arg("out_scales"), py::arg("m"), py::arg("n"), py::arg("r"), py::arg("scale"), py::arg("weights"),
      py::arg("base_fp8"), py::arg("out_fp8"), py::arg("stream"), py::arg("use_tma") = true, py::arg("bcat") = 0,
      py::arg("atcat") = 0, py::arg("kpad") = 0, py::arg("deep") = false);

  m.def(
      "local_sgd",
      [](uint64_t w, uint64_t g, uint64_t mom, uint64_t n, float lr, float momentum, float weight_decay,
         float grad_scale, bool zero_grad, int blocks, uint64_t stream) {
        LocalSgdArgs a{};
        a.w = P<__nv_bfloat16>(w);
        a.g = P<__nv_bfloat16>(g);
        a.mom = P<float>(mom);
        a.n = n;
        a.lr = lr;
        a.momentum = momentum;
        a.weight_decay = weight_decay;
        a.grad_scale = grad_scale;
        a.zero_grad = zero_grad ? 1 : 0;
        check(fedk_launch_local_sgd(&a, blocks, (cudaStream_t)(uintptr_t)stream), "fed_local_sgd launch");
      },
      py::arg("w"), py::arg("g"), py::arg("mom"), py::arg("n"), py::arg("lr"), py::arg("momentum"),
      py::arg("weight_decay"), py::arg("grad_scale"), py::arg("zero_grad"), py::arg("blocks"), py::arg("stream"));
}
