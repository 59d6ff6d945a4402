// g2vlm_b200 — host-side runtime shared by all entry points: error slot, TMA tensor-map encoding.
#include <string.h>

#include <mutex>
#include <string>

#include "common.cuh"

namespace g2 {

static thread_local std::string g_last_error;

void set_last_error(const char* file, int line, const char* msg) {
  char buf[768];
  const char* base = strrchr(file, '/');
  snprintf(buf, sizeof(buf), "%s:%d: %s", base ? base + 1 : file, line, msg);
  g_last_error = buf;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The driver entry point is resolved through the runtime so the library has no link-time
// dependency on libcuda.so (which does not exist on the GPU-less build container).
static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled driver entry point unavailable");
    return G2VLM_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_pitch_bytes & 15) || box_cols * 2 > 128 ||
      box_rows > 256 || rows == 0 || cols == 0) {
    set_last_error(__FILE__, __LINE__,
                   "tensor map: base/pitch must be 16-byte aligned, box <= 64 x 256");
    return G2VLM_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[128];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    set_last_error(__FILE__, __LINE__, msg);
    return G2VLM_ERR_CUDA;
  }
  return G2VLM_OK;
}

int make_tmap_2d_f32_box32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled driver entry point unavailable");
    return G2VLM_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_pitch_bytes & 15) || rows == 0 || cols == 0) {
    set_last_error(__FILE__, __LINE__, "tensor map (fp32): base/pitch must be 16-byte aligned");
    return G2VLM_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {32, 32};   // 32 fp32 = 128 bytes: one swizzle row
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[128];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
    set_last_error(__FILE__, __LINE__, msg);
    return G2VLM_ERR_CUDA;
  }
  return G2VLM_OK;
}

static constexpr int kMaxDevices = 64;

int num_sms() {
  static int n[kMaxDevices] = {0};   // per device (a process may drive several GPUs)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  }
  return n[dev];
}

int ensure_dyn_smem(const void* func, int bytes) {
  struct Entry { const void* func; unsigned long long devices; };
  static Entry table[64];
  static int n_entries = 0;
  static std::mutex mu;
  int dev = 0;
  G2_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) {
    G2_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return G2VLM_OK;
  }
  std::lock_guard<std::mutex> lock(mu);
  Entry* e = nullptr;
  for (int i = 0; i < n_entries; ++i)
    if (table[i].func == func) { e = &table[i]; break; }
  if (e == nullptr && n_entries < 64) { e = &table[n_entries++]; e->func = func; e->devices = 0; }
  if (e != nullptr && (e->devices >> dev) & 1ull) return G2VLM_OK;
  G2_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (e != nullptr) e->devices |= 1ull << dev;
  return G2VLM_OK;
}

}  // namespace g2

extern "C" int g2vlm_abi_version(void) { return G2VLM_ABI_VERSION; }
extern "C" const char* g2vlm_last_error(void) { return g2::g_last_error.c_str(); }
