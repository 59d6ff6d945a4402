// g2vlm_b200 — view-sharded K|V exchange for hosts that own a raw NCCL communicator ("g2vlm_sp_*").
// NCCL is resolved at run time from the library the process already uses (torch bundles libnccl.so.2): no link-time
// dependency, and a box without NCCL still loads the kernel library.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace g2 {

// the four NCCL entry points used here (nccl.h: ncclResult_t = int, ncclSuccess = 0, ncclBfloat16 = 9)
typedef int (*nccl_group_fn)();
typedef int (*nccl_send_fn)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_recv_fn)(void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_err_fn)(int);
constexpr int kNcclBfloat16 = 9;

struct NcclApi {
  nccl_group_fn group_start = nullptr, group_end = nullptr;
  nccl_send_fn send = nullptr;
  nccl_recv_fn recv = nullptr;
  nccl_err_fn err = nullptr;
  bool ok = false;
};

static const NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the process already loaded (torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.group_start = reinterpret_cast<nccl_group_fn>(dlsym(h, "ncclGroupStart"));
    api.group_end = reinterpret_cast<nccl_group_fn>(dlsym(h, "ncclGroupEnd"));
    api.send = reinterpret_cast<nccl_send_fn>(dlsym(h, "ncclSend"));
    api.recv = reinterpret_cast<nccl_recv_fn>(dlsym(h, "ncclRecv"));
    api.err = reinterpret_cast<nccl_err_fn>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.group_start && api.group_end && api.send && api.recv;
  });
  return api;
}

}  // namespace g2

#define G2_NCCL_OK(expr)                                                                        \
  do {                                                                                          \
    const int _r = (expr);                                                                      \
    if (_r != 0) {                                                                              \
      char _m[192];                                                                             \
      snprintf(_m, sizeof(_m), "NCCL: %s failed: %s", #expr, api.err ? api.err(_r) : "error"); \
      set_last_error(__FILE__, __LINE__, _m);                                                   \
      return G2VLM_ERR_CUDA;                                                                    \
    }                                                                                           \
  } while (0)

extern "C" int g2vlm_sp_kv_exchange(void* nccl_comm, int32_t rank, int32_t world, const int64_t* rank_rows, int64_t kvw,
                                    const void* kv_send, void* kv_remote, void* stream) {
  using namespace g2;
  G2_REQUIRE(nccl_comm && rank_rows && kv_send && kv_remote, "sp_kv_exchange: null argument");
  G2_REQUIRE(world >= 1 && rank >= 0 && rank < world && kvw > 0, "sp_kv_exchange: bad rank / world / row width");
  const NcclApi& api = nccl_api();
  G2_REQUIRE(api.ok, "sp_kv_exchange: libnccl.so.2 (ncclGroupStart / ncclSend / ncclRecv) is not available in this process");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t mine = static_cast<size_t>(rank_rows[rank]) * static_cast<size_t>(kvw);
  G2_NCCL_OK(api.group_start());
  long long off = 0;
  for (int j = 0; j < world; ++j) {
    if (j == rank) continue;
    G2_REQUIRE(rank_rows[j] >= 0, "sp_kv_exchange: negative row count");
    const size_t theirs = static_cast<size_t>(rank_rows[j]) * static_cast<size_t>(kvw);
    if (mine) G2_NCCL_OK(api.send(kv_send, mine, kNcclBfloat16, j, nccl_comm, st));
    if (theirs)
      G2_NCCL_OK(api.recv(reinterpret_cast<__nv_bfloat16*>(kv_remote) + off * kvw, theirs, kNcclBfloat16, j, nccl_comm, st));
    off += rank_rows[j];
  }
  G2_NCCL_OK(api.group_end());
  return G2VLM_OK;
}
