// Virtual-memory-management allocations and NVSwitch multicast objects for the peer-memory communicator.
//
// Why: the in-switch reduction (NVLS) that NCCL uses for /root/reference/cifar_example_ddp.py:83,106 needs the
// gradient / parameter buffers of all ranks bound to ONE multicast object, which only works for memory created
// with cuMemCreate (not cudaMalloc + cudaIpc).  With it, `multimem.ld_reduce.add.v4.f32` on the multicast
// address returns the sum over all ranks' copies (the switch does the W reads and the adds), and `multimem.st`
// writes one value into all W replicas.
//
// Every process creates its buffers with cuMemCreate (shareable as POSIX file descriptors), maps them, exports
// the descriptors to its peers (host side: Unix-domain sockets + SCM_RIGHTS, parallel.py), imports the peers'
// descriptors and maps them too (plain peer pointers for the flag pads and the non-NVLS fallback).  Rank 0
// additionally creates the multicast object; every rank adds its device, binds its local allocation at offset 0
// and maps the object: that mapping is the multicast pointer.
//
// The driver entry points are resolved at run time (cudaGetDriverEntryPoint): the library does not link libcuda.
#include <cuda.h>
#include <unistd.h>
#include <vector>

#include "ops.cuh"

namespace b2 {

struct Drv {
    bool ready = false;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
    CUresult (*MemRelease)(CUmemGenericAllocationHandle);
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
    CUresult (*MemAddressFree)(CUdeviceptr, size_t);
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    CUresult (*MemUnmap)(CUdeviceptr, size_t);
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
    CUresult (*MemExport)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
    CUresult (*MemImport)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType);
    CUresult (*MemGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags);
    CUresult (*McCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *);
    CUresult (*McAddDevice)(CUmemGenericAllocationHandle, CUdevice);
    CUresult (*McBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
    CUresult (*McGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags);
    CUresult (*DeviceGet)(CUdevice *, int);
    CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice);
    CUresult (*GetErrorString)(CUresult, const char **);
};
static Drv g_drv;

template <typename F>
static int load(F *slot, const char *name) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    B2_CUDA(cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q));
    B2_CHECK(fn && q == cudaDriverEntryPointSuccess, B2_ERR_UNSUPPORTED, "driver entry point %s not available", name);
    *slot = (F)fn;
    return 0;
}
static int drv_init() {
    if (g_drv.ready) return 0;
    B2_CUDA(cudaFree(0));
    B2_TRY(load(&g_drv.MemCreate, "cuMemCreate"));
    B2_TRY(load(&g_drv.MemRelease, "cuMemRelease"));
    B2_TRY(load(&g_drv.MemAddressReserve, "cuMemAddressReserve"));
    B2_TRY(load(&g_drv.MemAddressFree, "cuMemAddressFree"));
    B2_TRY(load(&g_drv.MemMap, "cuMemMap"));
    B2_TRY(load(&g_drv.MemUnmap, "cuMemUnmap"));
    B2_TRY(load(&g_drv.MemSetAccess, "cuMemSetAccess"));
    B2_TRY(load(&g_drv.MemExport, "cuMemExportToShareableHandle"));
    B2_TRY(load(&g_drv.MemImport, "cuMemImportFromShareableHandle"));
    B2_TRY(load(&g_drv.MemGranularity, "cuMemGetAllocationGranularity"));
    B2_TRY(load(&g_drv.McCreate, "cuMulticastCreate"));
    B2_TRY(load(&g_drv.McAddDevice, "cuMulticastAddDevice"));
    B2_TRY(load(&g_drv.McBindMem, "cuMulticastBindMem"));
    B2_TRY(load(&g_drv.McGranularity, "cuMulticastGetGranularity"));
    B2_TRY(load(&g_drv.DeviceGet, "cuDeviceGet"));
    B2_TRY(load(&g_drv.DeviceGetAttribute, "cuDeviceGetAttribute"));
    B2_TRY(load(&g_drv.GetErrorString, "cuGetErrorString"));
    g_drv.ready = true;
    return 0;
}
static const char *drv_err(CUresult r) {
    const char *s = nullptr;
    if (g_drv.GetErrorString && g_drv.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "unknown driver error";
}
#define B2_DRV(expr)                                                                             \
    do {                                                                                         \
        CUresult _r = (expr);                                                                    \
        if (_r != CUDA_SUCCESS) {                                                                \
            b2::set_error("%s:%d: %s -> %d (%s)", __FILE__, __LINE__, #expr, (int)_r, drv_err(_r)); \
            return B2_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

struct VmmAlloc { CUdeviceptr ptr; size_t bytes; CUmemGenericAllocationHandle handle; bool owns_handle; };
static std::vector<VmmAlloc> g_vmm;        // everything mapped by this process (own, imported, multicast)
static size_t g_align = 0;                 // granularity reported by b2_vmm_query: every reservation is aligned to it

static CUmemAllocationProp alloc_prop(int device) {
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}
static int map_handle(CUmemGenericAllocationHandle h, size_t bytes, size_t align, int device, bool owns, void **out) {
    CUdeviceptr va = 0;
    B2_DRV(g_drv.MemAddressReserve(&va, bytes, align ? align : g_align, 0, 0));
    B2_DRV(g_drv.MemMap(va, bytes, 0, h, 0));
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = device; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    B2_DRV(g_drv.MemSetAccess(va, bytes, &acc, 1));
    g_vmm.push_back(VmmAlloc{va, bytes, h, owns});
    *out = (void *)va;
    return 0;
}

}  // namespace b2

using namespace b2;

extern "C" {

int b2_vmm_query(int device, int *multicast_supported, int64_t *granularity) {
    B2_TRY(drv_init());
    CUdevice dev;
    B2_DRV(g_drv.DeviceGet(&dev, device));
    int mc = 0;
    B2_DRV(g_drv.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
    if (multicast_supported) *multicast_supported = mc;
    CUmemAllocationProp prop = alloc_prop(device);
    size_t g = 0;
    B2_DRV(g_drv.MemGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    if (mc) {
        CUmulticastObjectProp mp{};
        mp.numDevices = 2; mp.size = g; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        size_t mg = 0;
        if (g_drv.McGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
    }
    g_align = g;
    if (granularity) *granularity = (int64_t)g;
    return 0;
}

/* zero-filled device allocation of `bytes` (a multiple of the granularity b2_vmm_query reports), mapped read/write
 * on `device`; *handle_out identifies it for b2_mc_bind; *fd_out is a POSIX file descriptor a peer process imports. */
int b2_vmm_alloc(int device, int64_t bytes, void **ptr, int64_t *handle_out, int *fd_out) {
    B2_TRY(drv_init());
    B2_CHECK(ptr && handle_out && fd_out && bytes > 0, B2_ERR_ARG, "vmm_alloc: bad arguments");
    CUmemAllocationProp prop = alloc_prop(device);
    CUmemGenericAllocationHandle h;
    B2_DRV(g_drv.MemCreate(&h, (size_t)bytes, &prop, 0));
    B2_TRY(map_handle(h, (size_t)bytes, 0, device, true, ptr));
    B2_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    B2_CUDA(cudaDeviceSynchronize());
    int fd = -1;
    B2_DRV(g_drv.MemExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *handle_out = (int64_t)h; *fd_out = fd;
    return 0;
}

/* map a peer's allocation (received file descriptor) read/write on `device` */
int b2_vmm_import(int device, int fd, int64_t bytes, void **peer_ptr) {
    B2_TRY(drv_init());
    B2_CHECK(peer_ptr && fd >= 0 && bytes > 0, B2_ERR_ARG, "vmm_import: bad arguments");
    CUmemGenericAllocationHandle h;
    B2_DRV(g_drv.MemImport(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    return map_handle(h, (size_t)bytes, 0, device, true, peer_ptr);
}

/* rank 0: create the multicast object for `world` devices, `bytes` per device */
int b2_mc_create(int world, int64_t bytes, int64_t *mc_handle, int *fd_out) {
    B2_TRY(drv_init());
    B2_CHECK(mc_handle && fd_out && world >= 2 && bytes > 0, B2_ERR_ARG, "mc_create: bad arguments");
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)world; mp.size = (size_t)bytes; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle h;
    B2_DRV(g_drv.McCreate(&h, &mp));
    int fd = -1;
    B2_DRV(g_drv.MemExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *mc_handle = (int64_t)h; *fd_out = fd;
    return 0;
}
int b2_mc_import(int fd, int64_t *mc_handle) {
    B2_TRY(drv_init());
    B2_CHECK(mc_handle && fd >= 0, B2_ERR_ARG, "mc_import: bad arguments");
    CUmemGenericAllocationHandle h;
    B2_DRV(g_drv.MemImport(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    *mc_handle = (int64_t)h;
    return 0;
}
int b2_mc_add_device(int64_t mc_handle, int device) {
    B2_TRY(drv_init());
    CUdevice dev;
    B2_DRV(g_drv.DeviceGet(&dev, device));
    B2_DRV(g_drv.McAddDevice((CUmemGenericAllocationHandle)mc_handle, dev));
    return 0;
}
/* after EVERY rank has added its device: bind the local allocation and map the object; *mc_ptr is the multicast address */
int b2_mc_bind_and_map(int64_t mc_handle, int64_t mem_handle, int64_t bytes, int device, void **mc_ptr) {
    B2_TRY(drv_init());
    B2_CHECK(mc_ptr, B2_ERR_ARG, "mc_bind_and_map: NULL out");
    B2_DRV(g_drv.McBindMem((CUmemGenericAllocationHandle)mc_handle, 0, (CUmemGenericAllocationHandle)mem_handle, 0, (size_t)bytes, 0));
    return map_handle((CUmemGenericAllocationHandle)mc_handle, (size_t)bytes, 0, device, false, mc_ptr);
}
int b2_vmm_close_fd(int fd) { if (fd >= 0) close(fd); return 0; }

/* unmap and release everything this process mapped through the calls above (call after the last kernel) */
int b2_vmm_release_all(void) {
    if (!g_drv.ready) return 0;
    cudaDeviceSynchronize();
    for (auto it = g_vmm.rbegin(); it != g_vmm.rend(); ++it) {
        g_drv.MemUnmap(it->ptr, it->bytes);
        g_drv.MemAddressFree(it->ptr, it->bytes);
    }
    for (auto &a : g_vmm) if (a.owns_handle) g_drv.MemRelease(a.handle);
    // multicast handles: one release per mapping that did not own (every rank holds one reference)
    for (auto &a : g_vmm) if (!a.owns_handle) g_drv.MemRelease(a.handle);
    g_vmm.clear();
    return 0;
}

}  // extern "C"
