// greenctx_probe.cu -- can the engine carve a few SMs out for its copy/prep streams with CUDA green
// contexts, while still using the runtime API for launches, copies and events?
// Checks: driver entry points through cudaGetDriverEntryPoint (no -lcuda), SM split, streams of the
// two partitions, runtime kernel launches on them (which SMs do they land on?), runtime events
// recorded on one partition and waited for on the other, cudaMemcpyAsync on a partition stream, and
// whether a small kernel on the service partition starts while a machine-filling kernel runs.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/_bin/greenctx_probe scripts/greenctx_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <set>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("FAIL %s: CUresult %d\n", #x, (int)r_); return 1; } } while (0)

template <class F> bool entry(const char* name, F& fn)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        printf("no driver entry point %s\n", name);
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

__global__ void where(unsigned* smids, long long spin)
{
    unsigned id; asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    if (threadIdx.x == 0) smids[blockIdx.x] = id;
    const long long t0 = clock64();
    while (clock64() - t0 < spin) { }
}

int main()
{
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    decltype(&cuDeviceGetDevResource) pGetRes; decltype(&cuDevSmResourceSplitByCount) pSplit;
    decltype(&cuDevResourceGenerateDesc) pDesc; decltype(&cuGreenCtxCreate) pCreate;
    decltype(&cuGreenCtxStreamCreate) pStream; decltype(&cuGreenCtxDestroy) pDestroy;
    if (!entry("cuDeviceGetDevResource", pGetRes) || !entry("cuDevSmResourceSplitByCount", pSplit) ||
        !entry("cuDevResourceGenerateDesc", pDesc) || !entry("cuGreenCtxCreate", pCreate) ||
        !entry("cuGreenCtxStreamCreate", pStream) || !entry("cuGreenCtxDestroy", pDestroy)) return 1;
    CUdevResource all{}, svc{}, rest{};
    CU(pGetRes(0, &all, CU_DEV_RESOURCE_TYPE_SM));
    printf("device SMs: %u\n", all.sm.smCount);
    for (unsigned want : {8u, 16u, 32u}) {
        unsigned nb = 1;
        CUdevResource g{}, r{};
        CUresult rc = pSplit(&g, &nb, &all, &r, 0, want);
        printf("split min %u: rc %d groups %u group SMs %u remaining %u\n", want, (int)rc, nb, g.sm.smCount, r.sm.smCount);
        if (want == 32u) { svc = g; rest = r; }
    }
    CUdevResourceDesc dsvc, drest;
    CU(pDesc(&dsvc, &svc, 1)); CU(pDesc(&drest, &rest, 1));
    CUgreenCtx gsvc, grest;
    CU(pCreate(&gsvc, dsvc, 0, CU_GREEN_CTX_DEFAULT_STREAM));
    CU(pCreate(&grest, drest, 0, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream ssvc, srest;
    CU(pStream(&ssvc, gsvc, CU_STREAM_NON_BLOCKING, 0));
    CU(pStream(&srest, grest, CU_STREAM_NON_BLOCKING, 0));
    cudaStream_t a = (cudaStream_t)ssvc, b = (cudaStream_t)srest;

    unsigned *d1, *d2, *h1, *h2;
    const int NB = 2048;
    CK(cudaMalloc((void**)&d1, NB * 4)); CK(cudaMalloc((void**)&d2, NB * 4));
    CK(cudaHostAlloc((void**)&h1, NB * 4, 0)); CK(cudaHostAlloc((void**)&h2, NB * 4, 0));
    where<<<NB, 64, 0, a>>>(d1, 1000);  CK(cudaGetLastError());
    where<<<NB, 64, 0, b>>>(d2, 1000);  CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h1, d1, NB * 4, cudaMemcpyDeviceToHost, a));
    CK(cudaMemcpyAsync(h2, d2, NB * 4, cudaMemcpyDeviceToHost, b));
    CK(cudaStreamSynchronize(a)); CK(cudaStreamSynchronize(b));
    std::set<unsigned> s1(h1, h1 + NB), s2(h2, h2 + NB);
    bool disjoint = true; for (unsigned x : s1) if (s2.count(x)) disjoint = false;
    printf("service partition kernel ran on %zu SMs, rest on %zu SMs, disjoint %d\n", s1.size(), s2.size(), (int)disjoint);

    // runtime events across partitions
    cudaEvent_t e0, e1, e2, e3;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
    CK(cudaEventRecord(e0, b));
    where<<<148 * 32, 64, 0, b>>>(d2, 2000000);          // ~1 ms per block, several waves: fills its partition
    CK(cudaEventRecord(e1, b));
    CK(cudaEventRecord(e2, a));
    where<<<64, 64, 0, a>>>(d1, 1000);                    // small kernel on the service partition
    CK(cudaEventRecord(e3, a));
    CK(cudaStreamWaitEvent(a, e1, 0));                    // cross-partition wait
    where<<<1, 64, 0, a>>>(d1, 1000);
    CK(cudaStreamSynchronize(a)); CK(cudaStreamSynchronize(b));
    float big, small, start;
    CK(cudaEventElapsedTime(&big, e0, e1)); CK(cudaEventElapsedTime(&small, e2, e3)); CK(cudaEventElapsedTime(&start, e0, e3));
    printf("big kernel %.3f ms; small kernel on the service partition took %.3f ms and finished %.3f ms after the big one started\n", big, small, start);

    // the same small kernel on a plain high-priority stream (no partition) for comparison
    int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t p, q; CK(cudaStreamCreateWithPriority(&p, cudaStreamNonBlocking, hi)); CK(cudaStreamCreateWithPriority(&q, cudaStreamNonBlocking, lo));
    CK(cudaEventRecord(e0, q));
    where<<<148 * 32, 64, 0, q>>>(d2, 2000000);
    CK(cudaEventRecord(e1, q));
    CK(cudaStreamWaitEvent(p, e0, 0));
    CK(cudaEventRecord(e2, p));
    where<<<64, 64, 0, p>>>(d1, 1000);
    CK(cudaEventRecord(e3, p));
    CK(cudaStreamSynchronize(p)); CK(cudaStreamSynchronize(q));
    CK(cudaEventElapsedTime(&big, e0, e1)); CK(cudaEventElapsedTime(&small, e2, e3)); CK(cudaEventElapsedTime(&start, e0, e3));
    printf("plain streams: big %.3f ms; small high-priority kernel took %.3f ms and finished %.3f ms after the big one started\n", big, small, start);
    // H2D copy on a partition stream
    void* hb; void* db; CK(cudaHostAlloc(&hb, 64 << 20, 0)); CK(cudaMalloc(&db, 64 << 20));
    CK(cudaEventRecord(e0, a)); CK(cudaMemcpyAsync(db, hb, 64 << 20, cudaMemcpyHostToDevice, a)); CK(cudaEventRecord(e1, a));
    CK(cudaStreamSynchronize(a)); CK(cudaEventElapsedTime(&big, e0, e1));
    printf("H2D 64 MB on the service stream: %.3f ms (%.1f GB/s)\n", big, 67.1 / big);
    CK(cudaStreamDestroy(a)); CK(cudaStreamDestroy(b));
    CU(pDestroy(gsvc)); CU(pDestroy(grest));
    printf("done\n");
    return 0;
}
