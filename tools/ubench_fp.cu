// Microbenchmark: per-SM issue throughput (in SM cycles, clock64) of candidate FIR inner-loop instructions, sm_100a.
// One 512-thread CTA per SM (4 warps per SMSP), 16 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp ubench_fp.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITERS 8192
#define NACC 8
typedef unsigned long long u64;

template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, long long *cyc, float a, float b, u64 nz, u64 one)
{
    float acc[NACC * 2];
    for (int i = 0; i < NACC * 2; i++) acc[i] = threadIdx.x * 1e-9f + i;
    u64 pk[NACC];
    for (int i = 0; i < NACC; i++) pk[i] = ((u64)__float_as_uint(acc[2*i+1]) << 32) | __float_as_uint(acc[2*i]);
    u64 ab = ((u64)__float_as_uint(a) << 32) | __float_as_uint(a);
    u64 bb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b);
    __half2 h[NACC * 2];
    for (int i = 0; i < NACC * 2; i++) h[i] = __floats2half2_rn(acc[i], acc[i]);
    __half2 ha = __floats2half2_rn(a, a), hb = __floats2half2_rn(b, b);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {            // FFMA 3-reg
#pragma unroll
            for (int i = 0; i < NACC * 2; i++) acc[i] = fmaf(acc[i], a, b);
        } else if (MODE == 1) {     // FMUL + FADD (exact scalar path)
#pragma unroll
            for (int i = 0; i < NACC * 2; i++) acc[i] = __fadd_rn(__fmul_rn(acc[i], a), b);
        } else if (MODE == 2) {     // FFMA2
#pragma unroll
            for (int i = 0; i < NACC; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pk[i]) : "l"(ab), "l"(bb));
        } else if (MODE == 4) {     // HFMA2
#pragma unroll
            for (int i = 0; i < NACC * 2; i++) h[i] = __hfma2(h[i], ha, hb);
        } else if (MODE == 5) {     // FFMA, immediate multiplier
#pragma unroll
            for (int i = 0; i < NACC * 2; i++) acc[i] = fmaf(acc[i], 0.999f, b);
        } else if (MODE == 6) {     // exact packed: FFMA2(a,b,-0) ; FFMA2(p,1,c)
#pragma unroll
            for (int i = 0; i < NACC; i++) {
                u64 p;
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(pk[i]), "l"(ab), "l"(nz));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pk[i]) : "l"(p), "l"(one), "l"(bb));
            }
        } else if (MODE == 7) {     // FMUL only
#pragma unroll
            for (int i = 0; i < NACC * 2; i++) acc[i] = __fmul_rn(acc[i], a);
        } else if (MODE == 8) {     // FADD only
#pragma unroll
            for (int i = 0; i < NACC * 2; i++) acc[i] = __fadd_rn(acc[i], b);
        } else if (MODE == 9) {     // FMUL with constant-bank operand + FADD
#pragma unroll
            for (int i = 0; i < NACC * 2; i++) acc[i] = __fadd_rn(__fmul_rn(acc[i], 0.999f), b);
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < NACC * 2; i++) s += acc[i] + __low2float(h[i]);
    for (int i = 0; i < NACC; i++) s += __uint_as_float((unsigned)pk[i]) + __uint_as_float((unsigned)(pk[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, double instr_per_iter, double laneops_per_instr, float *d, long long *dc)
{
    int sms = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, 512>>>(d, dc, 0.999f, 1e-3f, 0x8000000080000000ull, 0x3f8000003f800000ull);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 10; r++) k<MODE><<<sms, 512>>>(d, dc, 0.999f, 1e-3f, 0x8000000080000000ull, 0x3f8000003f800000ull);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    long long hc[148]; cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
    double cmax = 0; for (int i = 0; i < sms; i++) if (hc[i] > cmax) cmax = hc[i];
    double winstr = 16.0 * ITERS * instr_per_iter;           // warp-instructions per SM
    printf("%-34s %7.3f ms  %9.0f cyc  => %5.2f warp-instr/clk/SM  %6.1f lane-ops/clk/SM   (implied clock %.0f MHz) %s\n", name, ms, cmax,
           winstr / cmax, winstr * 32 * laneops_per_instr / cmax, cmax / (ms * 1e3), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *d; cudaMalloc(&d, 148 * 512 * 4);
    long long *dc; cudaMalloc(&dc, 148 * 8);
    for (int rep = 0; rep < 2; rep++) {
        run<0>("FFMA 3-reg", 16, 1, d, dc);
        run<5>("FFMA imm", 16, 1, d, dc);
        run<7>("FMUL", 16, 1, d, dc);
        run<8>("FADD", 16, 1, d, dc);
        run<1>("FMUL+FADD (exact scalar)", 32, 1, d, dc);
        run<9>("FMUL imm + FADD", 32, 1, d, dc);
        run<2>("FFMA2", 8, 2, d, dc);
        run<6>("FFMA2 x2 (exact packed mul,add)", 16, 2, d, dc);
        run<4>("HFMA2", 16, 2, d, dc);
    }
    return 0;
}
