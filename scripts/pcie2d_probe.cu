// pcie2d_probe.cu -- how fast can the result / descriptor COLUMNS of 72-byte SeqPair records cross
// PCIe without moving the whole record?  Compares, for 1M records in page-locked host memory:
//   full     contiguous cudaMemcpyAsync of the records (what the direct route does today)
//   2d       cudaMemcpy2DAsync of the 24 result bytes (D2H) / the first 40 bytes (H2D), pitch 72
//   mapped   a kernel storing the 24 result bytes into / loading 40 bytes from mapped host memory
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/_bin/pcie2d_probe scripts/pcie2d_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct Rec { int64_t idr, idq, id; int32_t len1, len2, h0, seqid, regid, score, tle, gtle, qle, gscore, max_off, pad; };
static_assert(sizeof(Rec) == 72, "layout");

// one thread per record: 24 result bytes (offset 44..68) as 4 + 8 + 8 + 4 byte stores
__global__ void store_results_thread(Rec* host, const int4* res, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 v = res[i];
        Rec& r = host[i];
        r.score = v.x & 0xffff; r.tle = v.y & 0xffff; r.gtle = v.y >> 16; r.qle = v.x >> 16; r.gscore = v.z & 0xffff; r.max_off = v.z >> 16;
    }
}
// warp-cooperative: lane l of 6 consecutive lanes writes word l of the record's result block
__global__ void store_results_words(Rec* host, const int4* res, int n)
{
    const long long total = (long long)n * 6;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / 6), k = (int)(t % 6);
        const int4 v = res[i];
        int val;
        switch (k) {
            case 0: val = v.x & 0xffff; break; case 1: val = v.y & 0xffff; break; case 2: val = v.y >> 16; break;
            case 3: val = v.x >> 16; break;    case 4: val = v.z & 0xffff; break; default: val = v.z >> 16; break;
        }
        reinterpret_cast<int*>(host + i)[11 + k] = val;
    }
}
// loads of the descriptor fields from mapped host memory: idr, idq (16 B) + len1, len2, h0 (12 B)
__global__ void load_desc_thread(const Rec* host, int4* out, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Rec& r = host[i];
        const longlong2 a = *reinterpret_cast<const longlong2*>(&r.idr + 0);   // records are 8-aligned: two 8-byte loads
        out[i] = make_int4((int)a.x, (int)a.y, r.len2 | (r.len1 << 16), r.h0);
    }
}
// warp-cooperative load of whole records through shared memory (contiguous 128-byte reads over PCIe)
__global__ void load_full_coop(const uint32_t* host, uint32_t* out, long long words)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < words; t += (long long)gridDim.x * blockDim.x)
        out[t] = host[t];
}

int main()
{
    const int n = 1000000;
    Rec* h = nullptr; Rec* d = nullptr; int4* dres = nullptr; uint8_t* dcol = nullptr;
    CK(cudaHostAlloc((void**)&h, sizeof(Rec) * n, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(h, 1, sizeof(Rec) * n);
    CK(cudaMalloc((void**)&d, sizeof(Rec) * n));
    CK(cudaMalloc((void**)&dres, sizeof(int4) * n));
    CK(cudaMalloc((void**)&dcol, 40 * (size_t)n));
    CK(cudaMemset(dres, 0, sizeof(int4) * n));
    Rec* hm = nullptr; CK(cudaHostGetDevicePointer((void**)&hm, h, 0));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char* name, double mb, auto fn) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0, s1); fn(); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        cudaError_t e = cudaGetLastError();
        printf("%-46s %8.3f ms  %7.1f GB/s payload%s\n", name, best, mb / best, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    timeit("D2H full records 72 MB", 72.0, [&] { cudaMemcpyAsync(h, d, sizeof(Rec) * n, cudaMemcpyDeviceToHost, s1); });
    timeit("H2D full records 72 MB", 72.0, [&] { cudaMemcpyAsync(d, h, sizeof(Rec) * n, cudaMemcpyHostToDevice, s1); });
    timeit("D2H 2D 24 B of 72 (dense src) 24 MB", 24.0, [&] { cudaMemcpy2DAsync((char*)h + 44, 72, dcol, 24, 24, n, cudaMemcpyDeviceToHost, s1); });
    timeit("D2H 2D 24 B of 72 (pitched src) 24 MB", 24.0, [&] { cudaMemcpy2DAsync((char*)h + 44, 72, (char*)d + 44, 72, 24, n, cudaMemcpyDeviceToHost, s1); });
    timeit("H2D 2D 40 B of 72 (dense dst) 40 MB", 40.0, [&] { cudaMemcpy2DAsync(dcol, 40, h, 72, 40, n, cudaMemcpyHostToDevice, s1); });
    timeit("mapped store, thread per record 24 MB", 24.0, [&] { store_results_thread<<<148 * 8, 256, 0, s1>>>(hm, dres, n); });
    timeit("mapped store, word per lane 24 MB", 24.0, [&] { store_results_words<<<148 * 8, 256, 0, s1>>>(hm, dres, n); });
    timeit("mapped load, thread per record 28(40) MB", 40.0, [&] { load_desc_thread<<<148 * 8, 256, 0, s1>>>(hm, (int4*)dcol, n); });
    timeit("mapped load, full records coalesced 72 MB", 72.0, [&] { load_full_coop<<<148 * 8, 256, 0, s1>>>((const uint32_t*)hm, (uint32_t*)d, (long long)n * 18); });
    // the same stores while an H2D stream is busy on the other stream (full duplex)
    uint8_t* hbig = nullptr; uint8_t* dbig = nullptr; const size_t big = 256u << 20;
    CK(cudaHostAlloc((void**)&hbig, big, cudaHostAllocDefault)); CK(cudaMalloc((void**)&dbig, big));
    timeit("H2D 256 MB alone", 256.0 * 1.048576, [&] { cudaMemcpyAsync(dbig, hbig, big, cudaMemcpyHostToDevice, s1); });
    timeit("H2D 256 MB + concurrent D2H full 72 MB x3", 256.0 * 1.048576, [&] {
        for (int k = 0; k < 3; ++k) cudaMemcpyAsync(h, d, sizeof(Rec) * n, cudaMemcpyDeviceToHost, s2);
        cudaMemcpyAsync(dbig, hbig, big, cudaMemcpyHostToDevice, s1); });
    cudaStreamSynchronize(s2);
    timeit("H2D 256 MB + concurrent mapped stores x3", 256.0 * 1.048576, [&] {
        for (int k = 0; k < 3; ++k) store_results_thread<<<148 * 8, 256, 0, s2>>>(hm, dres, n);
        cudaMemcpyAsync(dbig, hbig, big, cudaMemcpyHostToDevice, s1); });
    cudaStreamSynchronize(s2);
    timeit("H2D 256 MB + concurrent 2D D2H x3", 256.0 * 1.048576, [&] {
        for (int k = 0; k < 3; ++k) cudaMemcpy2DAsync((char*)h + 44, 72, dcol, 24, 24, n, cudaMemcpyDeviceToHost, s2);
        cudaMemcpyAsync(dbig, hbig, big, cudaMemcpyHostToDevice, s1); });
    cudaStreamSynchronize(s2);
    printf("done\n");
    return 0;
}
