// issue_probe.cu -- developer micro-benchmark: how FP64 instructions with REGISTER operands share an SMSP's issue bandwidth with
// other instructions.  Per thread: C independent DFMA chains a = fma(a, b, d) (b, d in registers), each DFMA followed by N
// independent integer instructions with register operands (LOP3 / IADD3 on their own chains).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/issue_probe scripts/issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C, int N, int REGOPS>   // REGOPS: number of register operands of the DFMA besides the accumulator... 2 = (reg,reg,reg), 1 = (reg,reg,imm), 0 = (reg,imm,imm)
__global__ void k_mix(double* out, long long* cyc, const double* in, const unsigned* iin, int iters)
{
    double a[C], b[C], d[C];
    unsigned z[C][3], w[C], v[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        a[c] = in[c] + threadIdx.x * 1e-9;
        b[c] = REGOPS >= 1 ? in[8 + c] : 1.0000001;
        d[c] = REGOPS >= 2 ? in[16 + c] : 1e-7;
        w[c] = iin[c] + threadIdx.x; v[c] = iin[8 + c];
        for (int j = 0; j < 3; j++) z[c][j] = iin[16 + c + j];
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int c = 0; c < C; c++) {
                a[c] = fma(a[c], b[c], d[c]);
#pragma unroll
                for (int j = 0; j < N; j++) z[c][j] = (z[c][j] ^ w[c]) + v[c];
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    unsigned zz = 0;
#pragma unroll
    for (int c = 0; c < C; c++) { s += a[c]; for (int j = 0; j < 3; j++) zz += z[c][j]; }
    if (s == 12345.678 || zz == 0x12345u) out[0] = s + zz;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <class K>
static void run(const char* name, K kern, int warps, int chains, int n_alu, double* d_out, long long* d_cyc, const double* d_in, const unsigned* d_iin)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 2048;
    kern<<<sms, 128 * warps>>>(d_out, d_cyc, d_in, d_iin, 8);
    kern<<<sms, 128 * warps>>>(d_out, d_cyc, d_in, d_iin, iters);
    cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double per_dfma = (double)cyc / ((double)iters * 16 * chains * warps);
    // one "alu" = LOP3 + IADD3 = two instructions
    printf("%-26s warps/SMSP %d, %d x (LOP3 + IADD3) per DFMA : %6.2f cycles per DFMA group per SMSP   [2 F + O = %d, max(2 F, F + O) = %d]  %s\n",
           name, warps, n_alu, per_dfma, 2 + 2 * n_alu, (1 + 2 * n_alu) > 2 ? 1 + 2 * n_alu : 2, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double* d_out; long long* d_cyc; double* d_in; unsigned* d_iin;
    cudaMalloc(&d_out, 64); cudaMalloc(&d_cyc, 64); cudaMalloc(&d_in, 64 * 8); cudaMalloc(&d_iin, 64 * 4);
    double h_in[64]; unsigned h_iin[64];
    for (int i = 0; i < 64; i++) { h_in[i] = 1.0 + 1e-7 * i; h_iin[i] = 12345u * (i + 1); }
    for (int i = 8; i < 16; i++) h_in[i] = 1.0000001 + 1e-9 * i;
    for (int i = 16; i < 24; i++) h_in[i] = 1e-7 * (1 + i);
    cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    cudaMemcpy(d_iin, h_iin, sizeof(h_iin), cudaMemcpyHostToDevice);
    for (int w : {5, 8}) {
        run("dfma(reg,reg,reg)", k_mix<4, 0, 2>, w, 4, 0, d_out, d_cyc, d_in, d_iin);
        run("dfma(reg,reg,reg)+alu", k_mix<4, 1, 2>, w, 4, 1, d_out, d_cyc, d_in, d_iin);
        run("dfma(reg,reg,reg)+2alu", k_mix<4, 2, 2>, w, 4, 2, d_out, d_cyc, d_in, d_iin);
        run("dfma(reg,reg,reg)+3alu", k_mix<4, 3, 2>, w, 4, 3, d_out, d_cyc, d_in, d_iin);
        run("dfma(reg,reg,imm)", k_mix<4, 0, 1>, w, 4, 0, d_out, d_cyc, d_in, d_iin);
        run("dfma(reg,reg,imm)+alu", k_mix<4, 1, 1>, w, 4, 1, d_out, d_cyc, d_in, d_iin);
        run("dfma(reg,reg,imm)+2alu", k_mix<4, 2, 1>, w, 4, 2, d_out, d_cyc, d_in, d_iin);
        run("dfma(reg,imm,imm)", k_mix<4, 0, 0>, w, 4, 0, d_out, d_cyc, d_in, d_iin);
        run("dfma(reg,imm,imm)+alu", k_mix<4, 1, 0>, w, 4, 1, d_out, d_cyc, d_in, d_iin);
        run("dfma(reg,imm,imm)+2alu", k_mix<4, 2, 0>, w, 4, 2, d_out, d_cyc, d_in, d_iin);
        run("dfma(reg,imm,imm)+3alu", k_mix<4, 3, 0>, w, 4, 3, d_out, d_cyc, d_in, d_iin);
    }
    return 0;
}
