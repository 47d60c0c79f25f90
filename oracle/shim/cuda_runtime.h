/* oracle/shim/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * A stand-in for <cuda_runtime.h> that lets the reference's single translation
 * unit (/root/reference/kernel.cu) be compiled by g++ and executed on the host,
 * so that the reference's *own* kernels and main() become the CPU oracle
 * (SURVEY.md section 4.2).  Nothing here is copied from the reference; it only
 * supplies the handful of CUDA names the reference uses:
 *   __global__, dim3, blockIdx/threadIdx, cudaSetDevice, cudaMalloc, cudaMemcpy,
 *   cudaFree, cudaMemcpyHostToDevice/DeviceToHost   (kernel.cu:527,758-786,...)
 * and a launch macro that replaces `K<<<grid,block>>>(args)` (the build recipe
 * rewrites the 18 launch sites with one sed expression, see oracle/Makefile).
 *
 * Exactness: no reference kernel uses shared memory, __syncthreads, atomics or
 * device math, and no launch has an intra-launch read-after-write hazard, so a
 * serial sweep over (block, thread) indices computes exactly what the GPU does,
 * up to FMA contraction (controlled by -ffp-contract on the g++ command line).
 *
 * Optional taps (environment variables, read once):
 *   RTM_SHIM_DUMP_D2H=<path>   append every device->host copy to <path>
 *                              (per shot: slot NT-1, slot NT-2, Drel1, Drel2)
 *   RTM_SHIM_GATHER=<prefix>   the injected shim_forward_step_hook() records the
 *                              forward field at the data positions per step and
 *                              writes <prefix><shot>.bin  (n traces x NT floats)
 *   RTM_SHIM_SNAP=<prefix>,<k1>,<k2>,...  full-grid forward snapshots at steps k
 */
#ifndef RTM_ORACLE_SHIM_CUDA_RUNTIME_H
#define RTM_ORACLE_SHIM_CUDA_RUNTIME_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define __global__
#define __device__
#define __host__

struct shim_uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static shim_uint3 blockIdx, threadIdx;

enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
typedef int cudaError_t;

static inline cudaError_t cudaSetDevice(int) { return 0; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
    *p = (T *)calloc(n ? n : 1, 1);
    return *p ? 0 : 2;
}
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind kind) {
    memcpy(d, s, n);
    if (kind == cudaMemcpyDeviceToHost) {
        static const char *path = getenv("RTM_SHIM_DUMP_D2H");
        if (path && *path) {
            FILE *f = fopen(path, "ab");
            if (f) { fwrite(s, 1, n, f); fclose(f); }
        }
    }
    return 0;
}

/* K<<<g,b>>>(args);  ->  SHIM_LAUNCH(g,b) K(args);   (a single statement) */
#define SHIM_LAUNCH(g, b)                                              \
    for (blockIdx.x = 0; blockIdx.x < (g).x; blockIdx.x++)             \
    for (blockIdx.y = 0; blockIdx.y < (g).y; blockIdx.y++)             \
    for (threadIdx.y = 0; threadIdx.y < (b).y; threadIdx.y++)          \
    for (threadIdx.x = 0; threadIdx.x < (b).x; threadIdx.x++)

/* Injected after the reference's Hybrid3 launch (kernel.cu:819): P = blended
 * slot k of the forward field.  Records the "gather" the reference never writes
 * (gather[j][k] = slot_k[s_z][s_l + j*ds]). */
static inline void shim_forward_step_hook(const float *P, int k, int NT, int NZ, int NX,
                                          int s_z, int s_l, int ds, int n, int shot) {
    static const char *gpre = getenv("RTM_SHIM_GATHER");
    static const char *snap = getenv("RTM_SHIM_SNAP");
    static float *g = 0;
    static int gshot = -1;
    if (gpre && *gpre) {
        if (gshot != shot) { free(g); g = (float *)calloc((size_t)n * NT, sizeof(float)); gshot = shot; }
        for (int j = 0; j < n; j++) g[(size_t)j * NT + k] = P[(size_t)s_z * NX + s_l + j * ds];
        if (k == NT - 1) {
            char name[512];
            snprintf(name, sizeof name, "%s%d.bin", gpre, shot);
            FILE *f = fopen(name, "wb");
            if (f) { fwrite(g, sizeof(float), (size_t)n * NT, f); fclose(f); }
        }
    }
    if (snap && *snap) {
        char buf[512];
        strncpy(buf, snap, sizeof buf - 1); buf[sizeof buf - 1] = 0;
        char *save = 0, *tok = strtok_r(buf, ",", &save);
        const char *pre = tok;
        while ((tok = strtok_r(0, ",", &save))) {
            if (atoi(tok) == k) {
                char name[600];
                snprintf(name, sizeof name, "%s%d_%d.bin", pre, shot, k);
                FILE *f = fopen(name, "wb");
                if (f) { fwrite(P, sizeof(float), (size_t)NZ * NX, f); fclose(f); }
            }
        }
    }
}
#endif
