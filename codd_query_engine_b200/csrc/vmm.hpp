// vmm.hpp -- growable device arrays on CUDA virtual memory management.
//
// The corpus (up to ~77 GB for 100M x 384 bf16 on one B200) must grow without ever being copied
// and without its base pointer changing, so that queries already in flight keep scanning a stable
// prefix while the indexer appends (BASELINE.json config 5).  Each array reserves virtual address
// space for `max_rows` once and maps physical 2 MB-granular chunks behind it on demand.
//
// The driver entry points are resolved through cudaGetDriverEntryPoint so that libcoddvec.so has no
// link-time dependency on libcuda.so.1 (it must dlopen on a machine without a driver, where every
// compute entry point then fails loudly with CV_ERR_CUDA).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace cv {

struct DriverApi {
    CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*memAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill) = nullptr;
    bool ok = false;
    std::string why;
};

inline const DriverApi& driver_api() {
    static DriverApi api = [] {
        DriverApi a;
        auto load = [&](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
            if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || *fn == nullptr) {
                if (a.why.empty()) a.why = std::string("driver entry point unavailable: ") + name;
                (void)cudaGetLastError();
                return false;
            }
            return true;
        };
        bool ok = true;
        ok &= load("cuMemAddressReserve", reinterpret_cast<void**>(&a.memAddressReserve));
        ok &= load("cuMemAddressFree", reinterpret_cast<void**>(&a.memAddressFree));
        ok &= load("cuMemCreate", reinterpret_cast<void**>(&a.memCreate));
        ok &= load("cuMemRelease", reinterpret_cast<void**>(&a.memRelease));
        ok &= load("cuMemMap", reinterpret_cast<void**>(&a.memMap));
        ok &= load("cuMemUnmap", reinterpret_cast<void**>(&a.memUnmap));
        ok &= load("cuMemSetAccess", reinterpret_cast<void**>(&a.memSetAccess));
        ok &= load("cuMemGetAllocationGranularity", reinterpret_cast<void**>(&a.memGetAllocationGranularity));
        ok &= load("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&a.tensorMapEncodeTiled));
        a.ok = ok;
        return a;
    }();
    return api;
}

class VmmArray {
public:
    VmmArray() = default;
    VmmArray(const VmmArray&) = delete;
    VmmArray& operator=(const VmmArray&) = delete;
    ~VmmArray() { destroy(); }

    // Reserve address space for max_bytes on `device`.  Returns "" on success, else an error text.
    std::string reserve(int device, size_t max_bytes) {
        const DriverApi& d = driver_api();
        if (!d.ok) return d.why;
        device_ = device;
        prop_ = CUmemAllocationProp{};
        prop_.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop_.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop_.location.id = device;
        CUresult r = d.memGetAllocationGranularity(&gran_, &prop_, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
        if (r != CUDA_SUCCESS || gran_ == 0) return "cuMemGetAllocationGranularity failed: " + std::to_string(r);
        reserved_ = round_up(max_bytes ? max_bytes : gran_, gran_);
        r = d.memAddressReserve(&base_, reserved_, 0, 0, 0);
        if (r != CUDA_SUCCESS) return "cuMemAddressReserve(" + std::to_string(reserved_) + ") failed: " + std::to_string(r);
        return "";
    }

    // Make [0, bytes) backed by physical memory.  Never moves existing data.
    // Returns "" / "OOM: ..." / other error text.
    std::string ensure(size_t bytes) {
        if (bytes <= mapped_) return "";
        if (bytes > reserved_) return "capacity exceeded: index was created with a smaller max_rows";
        const DriverApi& d = driver_api();
        // geometric growth, capped at +1 GiB per step, so small collections stay small
        size_t grow = mapped_ < (size_t(1) << 30) ? mapped_ : (size_t(1) << 30);
        size_t target = bytes > mapped_ + grow ? bytes : mapped_ + grow;
        target = round_up(target, gran_);
        if (target > reserved_) target = reserved_;
        while (true) {
            size_t chunk = target - mapped_;
            CUmemGenericAllocationHandle h;
            CUresult r = d.memCreate(&h, chunk, &prop_, 0);
            if (r == CUDA_ERROR_OUT_OF_MEMORY && target > round_up(bytes, gran_)) {
                target = round_up(bytes, gran_);  // retry with the exact need
                continue;
            }
            if (r != CUDA_SUCCESS) return (r == CUDA_ERROR_OUT_OF_MEMORY ? "OOM: cuMemCreate(" : "cuMemCreate(") + std::to_string(chunk) + ") failed: " + std::to_string(r);
            r = d.memMap(base_ + mapped_, chunk, 0, h, 0);
            if (r != CUDA_SUCCESS) { d.memRelease(h); return "cuMemMap failed: " + std::to_string(r); }
            CUmemAccessDesc acc{};
            acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            acc.location.id = device_;
            acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            r = d.memSetAccess(base_ + mapped_, chunk, &acc, 1);
            if (r != CUDA_SUCCESS) { d.memUnmap(base_ + mapped_, chunk); d.memRelease(h); return "cuMemSetAccess failed: " + std::to_string(r); }
            chunks_.push_back({h, mapped_, chunk});
            mapped_ += chunk;
            return "";
        }
    }

    void destroy() {
        if (!base_) return;
        const DriverApi& d = driver_api();
        for (auto& c : chunks_) {
            d.memUnmap(base_ + c.offset, c.size);
            d.memRelease(c.handle);
        }
        chunks_.clear();
        d.memAddressFree(base_, reserved_);
        base_ = 0;
        mapped_ = reserved_ = 0;
    }

    template <typename T> T* ptr() const { return reinterpret_cast<T*>(base_); }
    size_t mapped_bytes() const { return mapped_; }
    size_t reserved_bytes() const { return reserved_; }

private:
    static size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }
    struct Chunk { CUmemGenericAllocationHandle handle; size_t offset; size_t size; };
    int device_ = 0;
    size_t gran_ = 0;
    size_t reserved_ = 0;
    size_t mapped_ = 0;
    CUdeviceptr base_ = 0;
    CUmemAllocationProp prop_{};
    std::vector<Chunk> chunks_;
};

}  // namespace cv
