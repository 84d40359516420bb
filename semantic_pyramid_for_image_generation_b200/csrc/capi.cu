// C-ABI plumbing: thread-local error string, launch counter, TMA descriptor encode.
#include "common.cuh"
#include "../../include/spyramid_b200.h"
#include <stdarg.h>
#include <atomic>

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void spyr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void spyr_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_precision{0};
bool spyr_split() { return g_precision.load(std::memory_order_relaxed) == 1; }

// which tensor-core kernel the last spyr_conv2d_fprop / spyr_conv2d_wgrad of this thread launched (bench.py keys its
// roofline entries by the real kernel name)
static thread_local int g_last_kernel = -1;
void spyr_note_kernel(int id) { g_last_kernel = id; }
extern "C" const char* spyr_last_conv_kernel(void) {
  static const char* names[] = {"conv_halo2_kernel", "conv_halo_kernel", "conv_fprop_kernel", "wgrad_halo_kernel",
                                "conv_wgrad_kernel", "conv_stack3_kernel"};
  return (g_last_kernel >= 0 && g_last_kernel < 6) ? names[g_last_kernel] : "";
}

extern "C" const char* spyr_last_error(void) { return g_err; }
extern "C" int spyr_set_precision(int mode) {
  SPYR_REQUIRE(mode == 0 || mode == 1, "spyr_set_precision: mode %d (0 = BF16 operands, 1 = split BF16 hi+lo)", mode);
  g_precision.store(mode);
  return 0;
}
extern "C" int spyr_get_precision(void) { return g_precision.load(); }
extern "C" int spyr_version(void) { return 200; }
extern "C" long long spyr_launch_count(void) { return g_launches.load(); }
extern "C" void spyr_launch_count_reset(void) { g_launches.store(0); }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int spyr_tmap_encode(CUtensorMap* map, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle128) {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || sym == nullptr || qres != cudaDriverEntryPointSuccess) {
      spyr_set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return 1;
    }
    fn = (PFN_encodeTiled)sym;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(gptr), gdims, gstr, gbox,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : (swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    spyr_set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu %llu %llu box %u %u %u)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                   box[0], box[1], rank > 2 ? box[2] : 0);
    return 1;
  }
  return 0;
}
