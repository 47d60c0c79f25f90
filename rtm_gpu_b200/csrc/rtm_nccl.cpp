// The one collective of the RTM job: sum of the per-GPU stacked images.
// Contexts living in one process (one host thread per GPU) are reduced with a single
// ncclReduce into a scratch buffer on the first context's GPU (the per-context stacks are not
// modified: the call is repeatable); if NCCL cannot be loaded or fails, peer copies do the same
// sum.  NCCL is loaded at run time (dlopen) so that the library has no link-time dependency on it;
// multi-process launchers (torchrun) reduce the buffers exposed by rtm_stack_device() with their
// own communicator instead.
#include "../../include/rtm_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
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

// One communicator set per process, created once (ncclCommInitAll takes seconds) and reused by every reduce over the
// same devices; rtm_stack_reduce_prepare lets a caller create it ahead of time, next to its other start-up work.
struct CommCache {
    std::mutex mu;
    std::vector<int> devs;
    std::vector<ncclComm_t> comms;
};
CommCache& cache() { static CommCache c; return c; }
// returns "" on success, the reason otherwise; comms valid while the cache lock is held by the caller
std::string ensure_comms(Nccl* n, const std::vector<int>& devs, std::vector<ncclComm_t>** out)
{
    CommCache& c = cache();
    if (c.devs != devs || c.comms.empty()) {
        for (auto cm : c.comms) if (cm) n->CommDestroy(cm);
        c.comms.assign(devs.size(), nullptr);
        c.devs.clear();
        const ncclResult_t r = n->CommInitAll(c.comms.data(), (int)devs.size(), devs.data());
        if (r) { c.comms.clear(); return std::string("ncclCommInitAll: ") + n->GetErrorString(r); }
        c.devs = devs;
    }
    *out = &c.comms;
    return "";
}
}  // namespace

extern "C" int rtm_stack_reduce_prepare(const int* devices, int n)
{
    if (!devices || n < 2) return RTM_OK;
    std::string why;
    Nccl* nc = load_nccl(why);
    if (!nc) return rtm_fail(RTM_ERR_NCCL, "rtm_stack_reduce_prepare: %s", why.c_str());
    std::lock_guard<std::mutex> l(cache().mu);
    std::vector<ncclComm_t>* comms = nullptr;
    why = ensure_comms(nc, std::vector<int>(devices, devices + n), &comms);
    if (!why.empty()) return rtm_fail(RTM_ERR_NCCL, "rtm_stack_reduce_prepare: %s", why.c_str());
    return RTM_OK;
}

extern "C" int rtm_ctx_device(rtm_ctx* ctx);
int rtm_stack_reduce_p2p(rtm_ctx** ctxs, int nctx, float* out);  // rtm_engine.cu
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
        size_t n_i = 0;
        devs[i] = rtm_ctx_device(ctxs[i]);
        if (int rc = rtm_stack_device(ctxs[i], &buf[i], &n_i, &ns)) return rc;
        if (i > 0 && n_i != nfl) return rtm_fail(RTM_ERR_ARG, "rtm_stack_reduce: contexts differ in image size");
        nfl = n_i;
        total += ns;
    }
    // The sum lands in a scratch buffer on the first context's GPU: the per-context stacks stay
    // untouched, so the call may be repeated and a failed NCCL attempt can fall back to peer copies.
    if (cudaSetDevice(devs[0]) != cudaSuccess) return rtm_fail(RTM_ERR_CUDA, "rtm_stack_reduce: cudaSetDevice(%d) failed", devs[0]);
    float* red = nullptr;
    if (cudaMalloc(&red, nfl * 4) != cudaSuccess) return rtm_fail(RTM_ERR_CUDA, "rtm_stack_reduce: scratch allocation of %zu bytes failed", nfl * 4);
    std::string why;
    const char* force = std::getenv("RTM_REDUCE");
    const bool want_p2p = force && std::string(force) == "p2p", want_nccl = force && std::string(force) == "nccl";
    bool done = false;
    Nccl* n = want_p2p ? nullptr : load_nccl(why);
    if (n) {   // one ncclReduce over NVLink (ncclCommInitAll: all contexts live in this process)
        std::lock_guard<std::mutex> lock(cache().mu);
        std::vector<ncclComm_t>* pc = nullptr;
        why = ensure_comms(n, devs, &pc);
        if (why.empty()) {
            std::vector<ncclComm_t>& comms = *pc;
            ncclResult_t r = 0;
            n->GroupStart();
            for (int i = 0; i < nctx && !r; ++i) {
                cudaSetDevice(devs[i]);
                r = n->Reduce(buf[i], i == 0 ? (void*)red : buf[i] /* ignored on non-root ranks */, nfl, ncclFloat, ncclSum, 0, comms[i], 0);
            }
            const ncclResult_t r2 = n->GroupEnd();
            if (!r) r = r2;
            for (int i = 0; i < nctx; ++i) {
                cudaSetDevice(devs[i]);
                if (cudaStreamSynchronize(0) != cudaSuccess && !r) r = 1;
            }
            if (r) {   // a failed communicator is not reused
                why = std::string("ncclReduce: ") + n->GetErrorString(r);
                for (auto c : comms) if (c) n->CommDestroy(c);
                cache().comms.clear(); cache().devs.clear();
            }
            else { done = true; g_backend = "nccl"; }
        }
    }
    if (!done) {
        // NCCL not loadable (stand-alone executable without the library on its path) or failed (no
        // /dev/shm, version mismatch, two contexts on one device ...): the same sum over NVLink peer
        // copies, added in context order on the first context's GPU.
        if (want_nccl) { cudaSetDevice(devs[0]); cudaFree(red); return rtm_fail(RTM_ERR_NCCL, "rtm_stack_reduce: %s", why.c_str()); }
        cudaGetLastError();   // a failed NCCL attempt may have left a sticky-free error behind
        if (int rc = rtm_stack_reduce_p2p(ctxs, nctx, red)) { cudaSetDevice(devs[0]); cudaFree(red); return rc; }
        g_backend = "p2p";
    }
    cudaSetDevice(devs[0]);
    const size_t ncell = nfl / 2;
    int rc = RTM_OK;
    if (up_sum && cudaMemcpy(up_sum, red, ncell * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = rtm_fail(RTM_ERR_CUDA, "rtm_stack_reduce: copy back failed");
    if (!rc && down_sum && cudaMemcpy(down_sum, red + ncell, ncell * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = rtm_fail(RTM_ERR_CUDA, "rtm_stack_reduce: copy back failed");
    cudaFree(red);
    if (nshots) *nshots = total;
    return rc;
}
