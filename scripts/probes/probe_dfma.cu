// DFMA throughput of one SM sub-partition as a function of resident warps and of the number of
// independent chains per thread (what occupancy does an fp64-pipe-bound kernel need on B200?)
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, int iters, double a, double b)
{
    double r[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) r[i] = fma(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += r[i];
    if (s == 1.2345) out[0] = s;
}
template <int CH>
void run(int warps_per_sm, int sms, double *out)
{
    // one CTA per SM with warps_per_sm warps (spread over the 4 sub-partitions)
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<CH><<<sms, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    const double fl = 2.0 * CH * (double)iters * warps_per_sm * 32 * sms;
    printf("chains %2d warps/SM %2d (%.2f per SMSP): %7.2f TFLOP/s\n", CH, warps_per_sm, warps_per_sm / 4.0, fl / (best * 1e-3) / 1e12);
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, 8);
    for (int w : {4, 8, 12, 16, 24, 32}) { run<8>(w, sms, out); run<16>(w, sms, out); run<32>(w, sms, out); }
    return 0;
}
