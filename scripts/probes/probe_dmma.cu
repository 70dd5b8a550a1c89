// fp64 throughput probes on B200: (1) DFMA alone, (2) DMMA (mma.sync f64) alone in its m8n8k4 and
// m16n8k16 shapes, (3) DFMA and DMMA interleaved in one warp, (4) DFMA warps next to DMMA warps.
// Question answered: is the fp64 tensor path a separate pipe (a hybrid kernel could beat the DFMA
// roofline) or the same units (nothing to gain)?
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// mode 0: DFMA only (16 chains); 1: m8n8k4 only (8 accumulator pairs); 2: m16n8k16 only (4 accumulator quads);
// 3: interleave 16 DFMA + 8 m8n8k4 per iteration; 4: even warps DFMA, odd warps m8n8k4
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
    double r[16], c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { r[i] = threadIdx.x + i; c[i] = 0.0; }
    const bool tensor_warp = (MODE == 1 || MODE == 2 || MODE == 3) || (MODE == 4 && ((threadIdx.x >> 5) & 1));
    const bool fma_warp = (MODE == 0 || MODE == 3) || (MODE == 4 && !((threadIdx.x >> 5) & 1));
    double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) bv[i] = b + i;
    for (int it = 0; it < iters; ++it) {
        if (fma_warp) {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = fma(r[i], a, b);
        }
        if (tensor_warp) {
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { double (&cc)[4] = *reinterpret_cast<double (*)[4]>(&c[4 * i]); dmma16816(cc, av, bv); }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i] + c[i];
    if (s == 1.2345) out[0] = s;
}

template <int MODE>
void run(const char *what, int sms, double *out, double fma_per_thread_iter, double mma_flop_per_warp_iter, double fma_warp_frac, double mma_warp_frac)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8192, blocks = sms * 4;
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    const double warps = 8.0 * blocks;
    const double fl_fma = 2.0 * fma_per_thread_iter * 32 * warps * fma_warp_frac * iters;
    const double fl_mma = mma_flop_per_warp_iter * warps * mma_warp_frac * iters;
    printf("%-46s %8.3f ms  DFMA %6.2f + DMMA %6.2f = %6.2f TFLOP/s\n", what, best, fl_fma / (best * 1e-3) / 1e12, fl_mma / (best * 1e-3) / 1e12,
           (fl_fma + fl_mma) / (best * 1e-3) / 1e12);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
}

// DMMA throughput against resident warps per SM and independent accumulator chains per warp
template <int CH>
__global__ void kocc(double *out, int iters, double a, double b)
{
    double c[2 * CH];
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) c[i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) s += c[i];
    if (s == 1.2345) out[0] = s;
}
template <int CH>
void occ(int warps_per_sm, int sms, double *out)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        kocc<CH><<<sms, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    const double fl = 512.0 * CH * (double)iters * warps_per_sm * sms;
    printf("DMMA chains %2d warps/SM %2d (%.2f per SMSP): %7.2f TFLOP/s\n", CH, warps_per_sm, warps_per_sm / 4.0, fl / (best * 1e-3) / 1e12);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, 8);
    run<0>("DFMA only, 16 chains", sms, out, 16, 0, 1, 0);
    run<1>("DMMA m8n8k4 only, 8 accumulators", sms, out, 0, 8 * 512.0, 0, 1);
    run<2>("DMMA m16n8k16 only, 4 accumulators", sms, out, 0, 4 * 4096.0, 0, 1);
    run<3>("DFMA + DMMA m8n8k4 interleaved per warp", sms, out, 16, 8 * 512.0, 1, 1);
    run<4>("DFMA warps beside DMMA m8n8k4 warps", sms, out, 16, 8 * 512.0, 0.5, 0.5);
    for (int w : {4, 8, 12, 16, 32}) { occ<1>(w, sms, out); occ<2>(w, sms, out); occ<4>(w, sms, out); occ<8>(w, sms, out); occ<16>(w, sms, out); }
    return 0;
}
