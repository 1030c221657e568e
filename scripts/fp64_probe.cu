// fp64_probe.cu -- developer micro-benchmarks of the B200 FP64 pipe (how many independent chains / warps saturate it, what a
// division costs in a dependent chain).  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false
// -o gpurun_out/fp64_probe scripts/fp64_probe.cu ; run on the GPU box, prints a table.
#include <cstdio>
#include <cuda_runtime.h>

#define KT_FAST 0
#include "../raytrace_cpu_b200/csrc/kt_device.cuh"

using namespace kt::exact;

// C independent DFMA chains per thread, each `iters * 32` deep
template <int C>
__global__ void k_chain(double* out, long long* cyc, int iters)
{
    double a[C];
#pragma unroll
    for (int c = 0; c < C; c++) a[c] = threadIdx.x * 1e-9 + c;
    const double b = 1.0000001, d = 1e-7;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 32; u++) {
#pragma unroll
            for (int c = 0; c < C; c++) a[c] = fma(a[c], b, d);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; c++) s += a[c];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// same with an ALU instruction between FP64 instructions (integer work that does not depend on the chain)
template <int C>
__global__ void k_chain_mixed(double* out, long long* cyc, int iters)
{
    double a[C];
    unsigned z = threadIdx.x;
#pragma unroll
    for (int c = 0; c < C; c++) a[c] = threadIdx.x * 1e-9 + c;
    const double b = 1.0000001, d = 1e-7;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 32; u++) {
#pragma unroll
            for (int c = 0; c < C; c++) {
                a[c] = fma(a[c], b, d);
                z = z * 1664525u + 1013904223u;
            }
        }
    }
    const long long t1 = clock64();
    double s = z;
#pragma unroll
    for (int c = 0; c < C; c++) s += a[c];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// dependent chain of other FP64 instructions: KIND 0 DMUL, 1 DADD, 2 DFMA followed by an ALU op on the high word (pipe crossing),
// 3 DFMA with the multiplicand in the constant bank / immediate replaced by a register operand
template <int KIND>
__global__ void k_other(double* out, long long* cyc, int iters, int mask, double breg)
{
    double a = 1.0 + threadIdx.x * 1e-9;
    const double b = 1.0000001, d = 1e-7;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 32; u++) {
            if (KIND == 0) a = a * breg;
            else if (KIND == 1) a = a + breg;
            else if (KIND == 2) { a = fma(a, b, d); a = __hiloint2double(__double2hiint(a) ^ mask, __double2loint(a)); }
            else a = fma(a, breg, breg);
        }
    }
    const long long t1 = clock64();
    if (a == 12345.678) out[0] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// dependent divisions: MODE 0 operator '/', 1 xdiv (guard deferred), 2 xdiv without the second Newton step; C chains
template <int MODE, int C>
__global__ void k_div(double* out, long long* cyc, int iters)
{
    double a[C], b[C];
#pragma unroll
    for (int c = 0; c < C; c++) { a[c] = 1.0 + threadIdx.x * 1e-3 + c; b[c] = 1.0000001 + 1e-9 * threadIdx.x + 1e-3 * c; }
    bool ok = true;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int c = 0; c < C; c++) {
                if (MODE == 0) a[c] = a[c] / b[c];
                else if (MODE == 1) a[c] = xdiv(a[c], b[c], ok);
                else {
                    double y;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b[c]));
                    y = __hiloint2double(__double2hiint(y), 1);
                    double e = fma(-b[c], y, 1.0);
                    e = fma(e, e, e);
                    y = fma(y, e, y);
                    double q = a[c] * y;
                    const double r = fma(-b[c], q, a[c]);
                    a[c] = fma(y, r, q);
                }
            }
        }
    }
    const long long t1 = clock64();
    double s = ok ? 0.0 : 1.0;
#pragma unroll
    for (int c = 0; c < C; c++) s += a[c];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <class K>
static void run(const char* name, K kern, int warps_per_smsp, int ops_per_iter, int iters, double* d_out, long long* d_cyc, int chains)
{
    const int threads = 128 * warps_per_smsp;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<sms, threads>>>(d_out, d_cyc, 8);
    cudaEventRecord(e0);
    kern<<<sms, threads>>>(d_out, d_cyc, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double per_warp_op = (double)cyc / ((double)iters * ops_per_iter);            // cycles between dependent ops of one chain
    const double per_smsp_instr = (double)cyc / ((double)iters * ops_per_iter * chains * warps_per_smsp);   // cycles per warp-instruction issued per SMSP
    printf("%-22s warps/SMSP %d chains %d : %7.2f cycles per dependent op, %6.2f cycles per op per SMSP, %.3f ms %s\n", name,
           warps_per_smsp, chains, per_warp_op, per_smsp_instr, ms, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 64); cudaMalloc(&d_cyc, 64);
    const int W[] = {1, 2, 4, 6, 8};
    for (int w : W) {
        run("dfma chain", k_chain<1>, w, 32, 4096, d_out, d_cyc, 1);
        run("dfma chain", k_chain<2>, w, 32, 4096, d_out, d_cyc, 2);
        run("dfma chain", k_chain<3>, w, 32, 4096, d_out, d_cyc, 3);
        run("dfma chain", k_chain<4>, w, 32, 4096, d_out, d_cyc, 4);
    }
    for (int w : W) {
        run("dfma+imad chain", k_chain_mixed<1>, w, 32, 4096, d_out, d_cyc, 1);
        run("dfma+imad chain", k_chain_mixed<2>, w, 32, 4096, d_out, d_cyc, 2);
        run("dfma+imad chain", k_chain_mixed<4>, w, 32, 4096, d_out, d_cyc, 4);
    }
    for (int w : W) {
        auto r2 = [&](const char* name, void (*kern)(double*, long long*, int, int, double), int ops) {
            int sms = 0;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
            kern<<<sms, 128 * w>>>(d_out, d_cyc, 8, 0, 1.0000001);
            kern<<<sms, 128 * w>>>(d_out, d_cyc, 4096, 0, 1.0000001);
            cudaDeviceSynchronize();
            long long cyc = 0;
            cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
            printf("%-22s warps/SMSP %d : %7.2f cycles per iteration (%d dependent instructions)\n", name, w, (double)cyc / (4096.0 * 32), ops);
        };
        r2("dmul chain", k_other<0>, 1);
        r2("dadd chain", k_other<1>, 1);
        r2("dfma -> lop3 chain", k_other<2>, 2);
        r2("dfma reg operands", k_other<3>, 1);
    }
    for (int w : W) {
        run("div operator", k_div<0, 1>, w, 8, 2048, d_out, d_cyc, 1);
        run("div operator", k_div<0, 3>, w, 8, 2048, d_out, d_cyc, 3);
        run("div xdiv", k_div<1, 1>, w, 8, 2048, d_out, d_cyc, 1);
        run("div xdiv", k_div<1, 3>, w, 8, 2048, d_out, d_cyc, 3);
        run("div 6-op", k_div<2, 1>, w, 8, 2048, d_out, d_cyc, 1);
        run("div 6-op", k_div<2, 3>, w, 8, 2048, d_out, d_cyc, 3);
    }
    return 0;
}
