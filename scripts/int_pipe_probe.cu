// int_pipe_probe.cu -- issue-rate microbenchmarks of the integer instructions the bsw kernels are
// made of (sm_100a).  Each test keeps 8 independent dependency chains per thread, so a warp always
// has an instruction ready; results are warp-instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipe_probe int_pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int OP>
__global__ void probe(int* out, int iters, int seed)
{
    int a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) a[k] = seed + threadIdx.x + k;
    int b = seed | 1, c = seed - 7;
    unsigned ub = (unsigned)b;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < CHAINS; ++k) {
                if (OP == 0) a[k] = __viaddmax_s32(a[k], b, c);                        // VIADDMNMX
                if (OP == 1) a[k] = __vimax3_s32(a[k], b, c);                          // VIMNMX3
                if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c));   // LOP3
                if (OP == 3) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));       // IMAD
                if (OP == 4) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));                      // IADD3 / VIADD
                if (OP == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));   // SHF
                if (OP == 6) asm volatile("max.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));                      // VIMNMX
                if (OP == 7) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));         // PRMT
                if (OP == 8) {                                                        // VIADDMNMX + IMAD alternating
                    if (k & 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
                    else a[k] = __viaddmax_s32(a[k], b, c);
                }
                if (OP == 9) {                                                        // LOP3 + IMAD alternating
                    if (k & 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c));
                }
                if (OP == 10) {                                                       // VIADDMNMX + LOP3 alternating
                    if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c));
                    else a[k] = __viaddmax_s32(a[k], b, c);
                }
                if (OP == 11) a[k] = __viaddmax_s16x2_relu((unsigned)a[k], ub, (unsigned)c);   // VIADDMNMX.S16x2.RELU
                if (OP == 12) asm volatile("{.reg .pred p; setp.ne.s32 p, %0, %1; selp.s32 %0, %1, %2, p;}" : "+r"(a[k]) : "r"(b), "r"(c));  // ISETP+SEL
                if (OP == 13) asm volatile("shl.b32 %0, %0, 3;" : "+r"(a[k]));        // shift by immediate (IMAD.SHL or SHF)
                if (OP == 14) {                                                       // VIADDMNMX + IADD alternating
                    if (k & 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
                    else a[k] = __viaddmax_s32(a[k], b, c);
                }
                if (OP == 16) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));                  // IMAD.HI.U32
                if (OP == 17) {                                                       // IMAD.HI + VIADDMNMX alternating
                    if (k & 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
                    else a[k] = __viaddmax_s32(a[k], b, c);
                }
                if (OP == 18) asm volatile("mad.lo.s32 %0, %0, -65536, %1;" : "+r"(a[k]) : "r"(c));           // IMAD with immediate
                if (OP == 19) asm volatile("shr.u32 %0, %0, 16;" : "+r"(a[k]));                              // shift right by immediate
                if (OP == 20) {                                                       // 2 IMAD : 1 IMAD.HI : 3 ALU (the planned cell mix)
                    if (k == 0 || k == 4) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
                    else if (k == 1 || k == 5) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
                    else if (k == 2) a[k] = __viaddmax_s32(a[k], b, c);
                    else if (k == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c));
                    else if (k == 6) a[k] = __vimax3_s32(a[k], b, c);
                    else asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
                }
                if (OP == 15) {                                                       // 1 IMAD : 3 mixed ALU
                    if ((k & 3) == 0) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
                    else if ((k & 3) == 1) a[k] = __viaddmax_s32(a[k], b, c);
                    else if ((k & 3) == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c));
                    else a[k] = __vimax3_s32(a[k], b, c);
                }
            }
        }
    }
    int x = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) x ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// shared-memory load/store rate next to ALU work: LDS + STS of one word per thread (bank == lane)
__global__ void probe_lds(int* out, int iters, int seed)
{
    extern __shared__ int sm[];
    int* p = sm + threadIdx.x;
    for (int k = 0; k < 32; ++k) p[k * blockDim.x] = seed + k;
    int x = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            int v = p[k * blockDim.x];
            x += v;
            p[k * blockDim.x] = x;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// 128-bit shared loads/stores, one 4-word group per lane, lane stride 4*odd words (conflict-free)
__global__ void probe_lds128(int* out, int iters, int seed, int stride_words)
{
    extern __shared__ int sm[];
    int4* p = reinterpret_cast<int4*>(sm + threadIdx.x * stride_words);
    for (int k = 0; k < 8; ++k) p[k] = make_int4(seed + k, seed, k, 1);
    int x = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int4 v = p[k];
            x += v.x + v.w;
            v.y = x;
            p[k] = v;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

static const char* names[] = {"VIADDMNMX", "VIMNMX3", "LOP3", "IMAD", "IADD(VIADD)", "SHF", "VIMNMX", "PRMT",
                              "VIADDMNMX+IMAD", "LOP3+IMAD", "VIADDMNMX+LOP3", "VIADDMNMX.S16x2.RELU",
                              "ISETP+SEL (2 instr)", "SHL imm", "VIADDMNMX+IADD", "IMAD+VIADDMNMX+LOP3+VIMNMX3",
                              "IMAD.HI.U32", "IMAD.HI+VIADDMNMX", "IMAD imm", "SHR imm", "cell mix 2HI:3IMAD:3ALU"};

template <int OP>
void run(int* d_out, int sms, double clk_ghz)
{
    const int threads = 256, blocks = sms * 8, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<OP><<<blocks, threads>>>(d_out, iters, 12345 + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    double mult = OP == 12 ? 2.0 : 1.0;
    const double winst = (double)threads / 32 * blocks * (double)iters * 4 * CHAINS * mult;
    const double per_clk_sm = winst / (best * 1e-3) / (clk_ghz * 1e9) / sms;
    printf("%-32s %8.3f ms  %6.3f warp-inst/clk/SM  (%5.1f lanes/clk/SM)  %.3e lane-ops/s\n", names[OP], best,
           per_clk_sm, per_clk_sm * 32, winst * 32 / (best * 1e-3));
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    printf("%s  SMs=%d  clock=%.3f GHz (max)\n", prop.name, prop.multiProcessorCount, ghz);
    int* d_out;
    cudaMalloc(&d_out, sizeof(int) * 256 * prop.multiProcessorCount * 8);
    const int sms = prop.multiProcessorCount;
    run<0>(d_out, sms, ghz); run<1>(d_out, sms, ghz); run<2>(d_out, sms, ghz); run<3>(d_out, sms, ghz);
    run<4>(d_out, sms, ghz); run<5>(d_out, sms, ghz); run<6>(d_out, sms, ghz); run<7>(d_out, sms, ghz);
    run<8>(d_out, sms, ghz); run<9>(d_out, sms, ghz); run<10>(d_out, sms, ghz); run<11>(d_out, sms, ghz);
    run<12>(d_out, sms, ghz); run<13>(d_out, sms, ghz); run<14>(d_out, sms, ghz); run<15>(d_out, sms, ghz);
    run<16>(d_out, sms, ghz); run<17>(d_out, sms, ghz); run<18>(d_out, sms, ghz); run<19>(d_out, sms, ghz); run<20>(d_out, sms, ghz);
    {
        const int threads = 64, blocks = sms * 16, iters = 2048;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            probe_lds<<<blocks, threads, 32 * threads * 4>>>(d_out, iters, rep);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        const double pairs = (double)threads / 32 * blocks * (double)iters * 32;
        printf("LDS+IADD+STS dependent triple      %8.3f ms  %6.3f triples/clk/SM\n", best, pairs / (best * 1e-3) / (ghz * 1e9) / sms);
    }
    for (int stride : {36, 44, 32, 33}) {
        const int threads = 64, blocks = sms * 8, iters = 2048;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            probe_lds128<<<blocks, threads, (threads * stride + 64) * 4>>>(d_out, iters, rep, stride);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        const double pairs = (double)threads / 32 * blocks * (double)iters * 8;
        printf("LDS.128+STS.128 pair, lane stride %2d words  %8.3f ms  %6.3f pairs/clk/SM\n", stride, best, pairs / (best * 1e-3) / (ghz * 1e9) / sms);
    }
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
