// vf_api.cu — library-wide plumbing of libvfuse: error string, launch counter, tensor-map encoding.
#include "vf_common.cuh"
#include "../../include/vfuse.h"

#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>

namespace vf {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return VF_OK;
  set_last_error("%s failed: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return VF_ERR_NO_DEVICE;
  return VF_ERR_CUDA;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    // on by default (VF_PDL=0 disables): with the folded LayerNorms the step has 24 tiny statistics kernels between its
    // GEMMs and overlapping every kernel's prologue (barrier init, TMEM allocation, tensor-map prefetch) with its
    // predecessor's tail is worth 0.12 ms of the 12.9 ms step (12.99 / 12.96 -> 12.88 / 12.83 ms, same box); it was
    // neutral before that, and 2-7 % for CUDA-graph latency at batch 1-4 (1.069 -> 0.998 ms at batch 1)
    const char* e = getenv("VF_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

int device_sm_count() {
  static std::atomic<int> sms[64];   // zero-initialised; indexed by device ordinal
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int n = sms[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    sms[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, uint32_t rank, const void* base,
                const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  VF_REQUIRE(fn != nullptr, VF_ERR_NO_DEVICE,
             "cuTensorMapEncodeTiled is unavailable (no CUDA driver / no GPU on this machine)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(out, dt, rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %u, dims %llu/%llu, box %u/%u)",
                   (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   box[0], rank > 1 ? box[1] : 0);
    return VF_ERR_CUDA;
  }
  return VF_OK;
}

}  // namespace vf

extern "C" int vf_version(void) { return VF_VERSION; }
extern "C" const char* vf_last_error(void) { return vf::g_err; }
extern "C" int64_t vf_launch_count(void) { return vf::g_launches.load(); }
extern "C" void vf_launch_count_reset(void) { vf::g_launches.store(0); }
