// The one collective of the RTM job: sum of the per-GPU stacked images.
// Contexts living in one process (one host thread per GPU) are reduced with a single
// ncclReduce to the first context's GPU.  NCCL is loaded at run time (dlopen) so that the
// library has no link-time dependency on it; multi-process launchers (torchrun) reduce the
// buffers exposed by rtm_stack_device() with their own communicator instead.
#include "../../include/rtm_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

int rtm_fail(int code, const char* fmt, ...);

namespace {
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
enum { ncclFloat = 7, ncclSum = 0 };
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl* load_nccl(std::string& why)
{
    static Nccl n;
    if (n.h) return &n;
    std::vector<std::string> names;
    if (const char* e = std::getenv("RTM_NCCL_LIB")) names.push_back(e);
    names.push_back("libnccl.so.2");
    names.push_back("libnccl.so");
    for (auto& nm : names) {
        n.h = dlopen(nm.c_str(), RTLD_NOW | RTLD_GLOBAL);
        if (n.h) break;
    }
    if (!n.h) { why = "libnccl.so.2 not found (set RTM_NCCL_LIB)"; return nullptr; }
#define SYM(field, name)                                             \
    n.field = (decltype(n.field))dlsym(n.h, name);                   \
    if (!n.field) { why = std::string("missing symbol ") + name; n.h = nullptr; return nullptr; }
    SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Reduce, "ncclReduce")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return &n;
}
}  // namespace

extern "C" int rtm_ctx_device(rtm_ctx* ctx);
int rtm_stack_reduce_p2p(rtm_ctx** ctxs, int nctx);  // rtm_engine.cu
static const char* g_backend = "none";
extern "C" const char* rtm_stack_reduce_backend(void) { return g_backend; }

extern "C" int rtm_stack_reduce(rtm_ctx** ctxs, int nctx, float* up_sum, float* down_sum, int* nshots)
{
    if (!ctxs || nctx < 1) return rtm_fail(RTM_ERR_ARG, "rtm_stack_reduce: no contexts");
    int total = 0;
    if (nctx == 1) return rtm_stack_get(ctxs[0], up_sum, down_sum, nshots);
    std::vector<int> devs(nctx);
    std::vector<void*> buf(nctx);
    size_t nfl = 0;
    for (int i = 0; i < nctx; ++i) {
        int ns = 0;
        devs[i] = rtm_ctx_device(ctxs[i]);
        if (int rc = rtm_stack_device(ctxs[i], &buf[i], &nfl, &ns)) return rc;
        total += ns;
    }
    std::string why;
    const char* force = std::getenv("RTM_REDUCE");
    Nccl* n = (force && std::string(force) == "p2p") ? nullptr : load_nccl(why);
    if (!n) {
        // NCCL not loadable (stand-alone executable without the library on its path): the same
        // sum over NVLink peer copies, gathered and added on the first context's GPU.
        if (force && std::string(force) == "nccl") return rtm_fail(RTM_ERR_NCCL, "rtm_stack_reduce: %s", why.c_str());
        int rc = rtm_stack_reduce_p2p(ctxs, nctx);
        if (rc) return rc;
        g_backend = "p2p";
    } else {
    std::vector<ncclComm_t> comms(nctx);
    ncclResult_t r = n->CommInitAll(comms.data(), nctx, devs.data());
    if (r) return rtm_fail(RTM_ERR_NCCL, "ncclCommInitAll: %s", n->GetErrorString(r));
    // in-place reduce into rank 0's stack (a copy of it is what the caller reads back)
    n->GroupStart();
    for (int i = 0; i < nctx && !r; ++i) {
        cudaSetDevice(devs[i]);
        r = n->Reduce(buf[i], buf[i], nfl, ncclFloat, ncclSum, 0, comms[i], 0);
    }
    ncclResult_t r2 = n->GroupEnd();
    if (!r) r = r2;
    for (int i = 0; i < nctx; ++i) {
        cudaSetDevice(devs[i]);
        cudaStreamSynchronize(0);
    }
    for (auto c : comms) n->CommDestroy(c);
    if (r) return rtm_fail(RTM_ERR_NCCL, "ncclReduce: %s", n->GetErrorString(r));
    g_backend = "nccl";
    }
    cudaSetDevice(devs[0]);
    const size_t ncell = nfl / 2;
    if (up_sum && cudaMemcpy(up_sum, buf[0], ncell * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return rtm_fail(RTM_ERR_CUDA, "rtm_stack_reduce: copy back failed");
    if (down_sum && cudaMemcpy(down_sum, (float*)buf[0] + ncell, ncell * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return rtm_fail(RTM_ERR_CUDA, "rtm_stack_reduce: copy back failed");
    if (nshots) *nshots = total;
    return RTM_OK;
}
