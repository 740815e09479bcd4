// Symmetric heap over the CUDA virtual-memory-management API, with an NVLS multicast mapping.
//
// Every rank of a tensor-parallel group creates ONE physical allocation of the same size (cuMemCreate, POSIX-fd shareable),
// maps its own and every peer's allocation into its address space (cuMemImportFromShareableHandle + cuMemMap), and — when the
// fabric supports it — binds all of them to one multicast object (cuMulticastCreate / AddDevice / BindMem) that is mapped a
// second time: a store to the multicast address lands in EVERY rank's copy, a `multimem.ld_reduce` returns the SUM over all
// copies computed inside the NVSwitch.  Offsets are symmetric: the same byte offset names "the same" buffer on every rank, so
// a kernel needs only (local base, peer bases[], multicast base).
//
// The file descriptors travel between the processes over Unix-domain sockets (SCM_RIGHTS) from Python
// (parallel/symm_heap.py); this file owns the driver calls.  Reference: the Neuron runtime's collective buffers are opaque to
// the reference; SURVEY §5.8 asks for the B200 equivalent (peer-mapped + multicast symmetric memory).
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

namespace nxdi {

namespace {

template <typename Fn>
Fn drv(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr)
    throw std::runtime_error(std::string("symm_heap: driver entry point not found: ") + name);
  return reinterpret_cast<Fn>(p);
}

void ck(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) throw std::runtime_error(std::string("symm_heap: ") + what + " failed with CUresult " + std::to_string((int)r));
}

struct Heap {
  int device = 0;
  size_t size = 0;       // rounded up to the allocation / multicast granularity
  CUmemGenericAllocationHandle local = 0;
  CUdeviceptr local_va = 0;
  std::vector<CUdeviceptr> peer_va;                         // by rank; own rank -> local_va
  std::vector<CUmemGenericAllocationHandle> peer_handles;   // imported handles (0 for own rank)
  CUmemGenericAllocationHandle mc = 0;
  CUdeviceptr mc_va = 0;
  bool mc_bound = false;
};

std::vector<Heap*>& heaps() {
  static std::vector<Heap*> v;
  return v;
}
Heap& heap(long long h) {
  if (h < 0 || h >= (long long)heaps().size() || heaps()[h] == nullptr) throw std::runtime_error("symm_heap: bad handle");
  return *heaps()[h];
}

CUmemAllocationProp alloc_prop(int device) {
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

CUdeviceptr map_rw(CUmemGenericAllocationHandle handle, size_t size, int device, size_t align) {
  auto reserve = drv<CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long)>("cuMemAddressReserve");
  auto map = drv<CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long)>("cuMemMap");
  auto set_access = drv<CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t)>("cuMemSetAccess");
  CUdeviceptr va = 0;
  ck(reserve(&va, size, align, 0, 0), "cuMemAddressReserve");
  ck(map(va, size, 0, handle, 0), "cuMemMap");
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  ck(set_access(va, size, &acc, 1), "cuMemSetAccess");
  return va;
}

}  // namespace

// -> handle; the allocation is zero-initialised
long long symm_heap_create(long long bytes, int device, int world, int rank) {
  auto granularity = drv<CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags)>("cuMemGetAllocationGranularity");
  auto create = drv<CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long)>("cuMemCreate");
  cudaSetDevice(device);
  cudaFree(nullptr);   // make sure the primary context exists
  auto* hp = new Heap();
  hp->device = device;
  CUmemAllocationProp prop = alloc_prop(device);
  size_t gran = 0;
  ck(granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  // the multicast object binds the whole allocation: its (minimum) granularity must divide the size too
  size_t mc_gran = 0;
  {
    auto mgran = drv<CUresult (*)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags)>("cuMulticastGetGranularity");
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)world;
    mp.size = (size_t)bytes;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (mgran(&mc_gran, &mp, CU_MULTICAST_GRANULARITY_MINIMUM) != CUDA_SUCCESS) mc_gran = 0;
  }
  size_t round_to = std::max<size_t>(gran, (size_t)32 << 20);
  if (mc_gran > round_to) round_to = mc_gran;
  if (mc_gran != 0 && round_to % mc_gran != 0) round_to = round_to / mc_gran * mc_gran + mc_gran;
  hp->size = ((size_t)bytes + round_to - 1) / round_to * round_to;
  ck(create(&hp->local, hp->size, &prop, 0), "cuMemCreate");
  hp->local_va = map_rw(hp->local, hp->size, device, round_to);
  hp->peer_va.assign(world, 0);
  hp->peer_handles.assign(world, 0);
  hp->peer_va[rank] = hp->local_va;
  if (cudaMemset(reinterpret_cast<void*>(hp->local_va), 0, hp->size) != cudaSuccess) throw std::runtime_error("symm_heap: memset failed");
  cudaDeviceSynchronize();
  heaps().push_back(hp);
  return (long long)heaps().size() - 1;
}

long long symm_heap_size(long long h) { return (long long)heap(h).size; }
long long symm_heap_local_va(long long h) { return (long long)heap(h).local_va; }

// file descriptor of this rank's allocation (the caller sends it to the peers and closes it)
int symm_heap_export_fd(long long h) {
  auto exp = drv<CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long)>("cuMemExportToShareableHandle");
  int fd = -1;
  ck(exp(&fd, heap(h).local, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
  return fd;
}

// map the allocation behind `fd` (received from rank `peer`) -> device address in this process
long long symm_heap_import_peer(long long h, int peer, int fd) {
  auto imp = drv<CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType)>("cuMemImportFromShareableHandle");
  Heap& hp = heap(h);
  CUmemGenericAllocationHandle handle = 0;
  ck(imp(&handle, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "cuMemImportFromShareableHandle");
  hp.peer_handles[peer] = handle;
  hp.peer_va[peer] = map_rw(handle, hp.size, hp.device, (size_t)32 << 20);
  return (long long)hp.peer_va[peer];
}

// ---- NVLS multicast ------------------------------------------------------------------------------------------------------
bool symm_heap_multicast_supported(int device) {
  int v = 0;
  auto get = drv<CUresult (*)(int*, CUdevice_attribute, CUdevice)>("cuDeviceGetAttribute");
  if (get(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device) != CUDA_SUCCESS) return false;
  return v != 0;
}

// rank 0: create the multicast object for `world` devices -> fd to hand to every rank
int symm_heap_mc_create(long long h, int world) {
  auto create = drv<CUresult (*)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*)>("cuMulticastCreate");
  auto gran = drv<CUresult (*)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags)>("cuMulticastGetGranularity");
  auto exp = drv<CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long)>("cuMemExportToShareableHandle");
  Heap& hp = heap(h);
  CUmulticastObjectProp prop{};
  prop.numDevices = (unsigned)world;
  prop.size = hp.size;
  prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t g = 0;
  ck(gran(&g, &prop, CU_MULTICAST_GRANULARITY_MINIMUM), "cuMulticastGetGranularity");
  if (hp.size % g != 0) throw std::runtime_error("symm_heap: heap size is not a multiple of the multicast granularity");
  ck(create(&hp.mc, &prop), "cuMulticastCreate");
  int fd = -1;
  ck(exp(&fd, hp.mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle(multicast)");
  return fd;
}

void symm_heap_mc_import(long long h, int fd) {
  auto imp = drv<CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType)>("cuMemImportFromShareableHandle");
  ck(imp(&heap(h).mc, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
     "cuMemImportFromShareableHandle(multicast)");
}

// every rank, after all ranks hold the object: join with this rank's device (collective: the caller barriers afterwards)
void symm_heap_mc_add_device(long long h) {
  auto add = drv<CUresult (*)(CUmemGenericAllocationHandle, CUdevice)>("cuMulticastAddDevice");
  Heap& hp = heap(h);
  ck(add(hp.mc, hp.device), "cuMulticastAddDevice");
}

// every rank, after ALL devices were added: bind the local physical memory and map the multicast address range
long long symm_heap_mc_bind_map(long long h) {
  auto bind = drv<CUresult (*)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long)>("cuMulticastBindMem");
  Heap& hp = heap(h);
  ck(bind(hp.mc, 0, hp.local, 0, hp.size, 0), "cuMulticastBindMem");
  hp.mc_bound = true;
  hp.mc_va = map_rw(hp.mc, hp.size, hp.device, (size_t)32 << 20);
  return (long long)hp.mc_va;
}

long long symm_heap_mc_va(long long h) { return (long long)heap(h).mc_va; }

void symm_heap_destroy(long long h) {
  Heap& hp = heap(h);
  auto unmap = drv<CUresult (*)(CUdeviceptr, size_t)>("cuMemUnmap");
  auto afree = drv<CUresult (*)(CUdeviceptr, size_t)>("cuMemAddressFree");
  auto release = drv<CUresult (*)(CUmemGenericAllocationHandle)>("cuMemRelease");
  cudaDeviceSynchronize();
  if (hp.mc_va) { unmap(hp.mc_va, hp.size); afree(hp.mc_va, hp.size); }
  if (hp.mc) release(hp.mc);
  for (size_t r = 0; r < hp.peer_va.size(); ++r) {
    if (hp.peer_va[r] && hp.peer_va[r] != hp.local_va) { unmap(hp.peer_va[r], hp.size); afree(hp.peer_va[r], hp.size); }
    if (hp.peer_handles[r]) release(hp.peer_handles[r]);
  }
  if (hp.local_va) { unmap(hp.local_va, hp.size); afree(hp.local_va, hp.size); }
  if (hp.local) release(hp.local);
  delete heaps()[h];
  heaps()[h] = nullptr;
}

}  // namespace nxdi
