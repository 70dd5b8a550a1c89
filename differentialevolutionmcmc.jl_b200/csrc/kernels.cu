// kernels.cu -- CUDA (sm_100a) implementation of backend.h: every kernel of libdemcmc_b200.
//
//   k_propose      one warp per particle: DE / snooker / mutation proposal, kappa and block masks,
//                  bounds, prior, snooker adjustment                       (HBM-bound, ~5 d-vectors)
//   k_xdot         cross term sum_i sum_k x'_ik m'_pk of the expanded sum of squares for a tile of
//                  particles against slices of observations: the likelihood of the isotropic MVN
//                  and the hierarchical normal models                      (fp64-pipe-bound)
//   k_ll_pointwise per-observation log densities (Gaussian, LNR, LBA) for a tile of particles
//   k_accept       one warp per particle: fixed-order reduction of the partial sums, Metropolis
//                  accept, state-row write (replaces store_samples!)       (HBM-bound)
//   k_mig_*        migration picks and the cyclic shift
//   k_history      by-slot rows -> the reference's samples[n_rows, d, P] layout
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <string>

#include "backend.h"
#include "de_particle.h"

namespace de {
namespace be {

static thread_local std::string g_be_err;
static int64_t g_launches = 0;
static int g_dev = 0;
static cudaStream_t g_stream[64] = { nullptr };
static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;

static int cu_fail(cudaError_t e, const char *what)
{
    g_be_err = std::string(what) + ": " + cudaGetErrorString(e);
    return -1;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cu_fail(e_, #call); } while (0)
#define LAUNCHED(name) do { ++g_launches; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return cu_fail(e_, name); } while (0)

static cudaStream_t stream()
{
    if (!g_stream[g_dev]) cudaStreamCreateWithFlags(&g_stream[g_dev], cudaStreamNonBlocking);
    return g_stream[g_dev];
}

const char *name() { return "cuda-sm100a"; }
const char *last_error() { return g_be_err.c_str(); }
int device_count()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cu_fail(e, "cudaGetDeviceCount"); return 0; }
    return n;
}
int set_device(int dev)
{
    if (dev < 0 || dev >= 64) { g_be_err = "device ordinal out of range"; return -1; }
    CU(cudaSetDevice(dev));
    g_dev = dev;
    return 0;
}
void *dmalloc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 8);
    if (e != cudaSuccess) { cu_fail(e, "cudaMalloc"); return nullptr; }
    return p;
}
void dfree(void *p) { if (p) cudaFree(p); }
void *hmalloc_pinned(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 8);
    if (e != cudaSuccess) { cu_fail(e, "cudaMallocHost"); return nullptr; }
    return p;
}
void hfree_pinned(void *p) { if (p) cudaFreeHost(p); }
int h2d(void *dst, const void *src, size_t bytes) { if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream())); return 0; }
int d2h(void *dst, const void *src, size_t bytes)
{
    if (!bytes) return 0;
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream()));
    CU(cudaStreamSynchronize(stream()));
    return 0;
}
int d2d(void *dst, const void *src, size_t bytes) { if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream())); return 0; }
int dzero(void *dst, size_t bytes) { if (bytes) CU(cudaMemsetAsync(dst, 0, bytes, stream())); return 0; }
int sync() { CU(cudaStreamSynchronize(stream())); return 0; }
void *event_create()
{
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    return (void *)e;
}
void event_destroy(void *ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
int event_record(void *ev) { CU(cudaEventRecord((cudaEvent_t)ev, stream())); return 0; }
int event_wait(void *ev) { CU(cudaEventSynchronize((cudaEvent_t)ev)); return 0; }
void *tevent_create()
{
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    return (void *)e;
}
void tevent_destroy(void *ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
int tevent_elapsed(void *a, void *b, double *ms)
{
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, (cudaEvent_t)a, (cudaEvent_t)b));
    *ms = f;
    return 0;
}
int dfill(void *dst, int byte, size_t bytes) { if (bytes) CU(cudaMemsetAsync(dst, byte, bytes, stream())); return 0; }
int timer_start()
{
    if (!g_t0) { CU(cudaEventCreate(&g_t0)); CU(cudaEventCreate(&g_t1)); }
    CU(cudaEventRecord(g_t0, stream()));
    return 0;
}
int timer_stop(double *ms)
{
    CU(cudaEventRecord(g_t1, stream()));
    CU(cudaEventSynchronize(g_t1));
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, g_t0, g_t1));
    *ms = f;
    return 0;
}
int64_t launch_count() { return g_launches; }

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct WarpLanes {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int width() const { return 32; }
    __device__ __forceinline__ double sum(double x) const { return warp_sum(x); }
    __device__ __forceinline__ bool all(bool b) const { return __all_sync(0xffffffffu, b) != 0; }
    __device__ __forceinline__ int min_int(int x) const { return __reduce_min_sync(0xffffffffu, x); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};

struct ksum_t { double s, c; };
__device__ __forceinline__ void kadd(ksum_t &k, double x)
{
    const double t = k.s + x;
    if (isfinite(t)) { if (fabs(k.s) >= fabs(x)) k.c += (k.s - t) + x; else k.c += (x - t) + k.s; }
    k.s = t;
}
__device__ __forceinline__ double kval(const ksum_t &k) { return isfinite(k.s) ? k.s + k.c : k.s; }

// ------------------------------------------------------------------------------------------------
// propose / accept: one warp per particle of the level
// ------------------------------------------------------------------------------------------------
constexpr int PA_THREADS = 128;
static double *mT_buffer(size_t doubles);                 // staging of centred means for k_xdot (below)
static size_t mT_doubles(const ModelDev &m, int n);

__global__ void __launch_bounds__(PA_THREADS) k_propose(ConfigDev cfg, ModelDev m, Level lv, double *mT)
{
    const int wi = (blockIdx.x * PA_THREADS + threadIdx.x) >> 5;
    if (wi >= lv.n) return;
    const uint32_t e = (uint32_t)lv.order[wi];
    const SweepCtx ctx = lv.ctxs[e >> LV_SLOT_SHIFT];
    const int p = (int)(e & LV_POS_MASK);
    propose_particle(WarpLanes(), cfg, m, ctx, p);
    if (mT) {
        // leave the centred means of this proposal where k_xdot wants them: mT[tile][k][64]
        __syncwarp();
        const double *prop = ctx.prop_theta + (size_t)p * cfg.d;
        const size_t tile = (size_t)(wi / SSD_TP);
        const int pi = wi % SSD_TP;
        for (int k = threadIdx.x & 31; k < m.ssd_k; k += 32) mT[(tile * m.ssd_k + k) * SSD_TP + pi] = centred_mean(m, prop, k);
    }
}

__global__ void __launch_bounds__(PA_THREADS) k_accept(ConfigDev cfg, ModelDev m, Level lv)
{
    const int wi = (blockIdx.x * PA_THREADS + threadIdx.x) >> 5;
    if (wi >= lv.n) return;
    const uint32_t e = (uint32_t)lv.order[wi];
    const SweepCtx ctx = lv.ctxs[e >> LV_SLOT_SHIFT];
    accept_particle(WarpLanes(), cfg, m, ctx, (int)(e & LV_POS_MASK));
}

int launch_propose(const ConfigDev &cfg, const ModelDev &m, const Level &lv)
{
    const int blocks = (lv.n * 32 + PA_THREADS - 1) / PA_THREADS;
    double *mT = nullptr;
    if (m.kind == M_MVNORMAL || m.kind == M_HIER) {
        // sized once for the handle's whole population so it never grows inside a run
        mT = mT_buffer(mT_doubles(m, std::max(lv.n, cfg.G_local * cfg.Np)));
        if (!mT) return -1;
    }
    k_propose<<<blocks, PA_THREADS, 0, stream()>>>(cfg, m, lv, mT);
    LAUNCHED("k_propose");
    return 0;
}

int launch_accept(const ConfigDev &cfg, const ModelDev &m, const Level &lv)
{
    const int blocks = (lv.n * 32 + PA_THREADS - 1) / PA_THREADS;
    k_accept<<<blocks, PA_THREADS, 0, stream()>>>(cfg, m, lv);
    LAUNCHED("k_accept");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// select_base preparation (crossover.jl:282-289) on the sweep-start weights, one block per group:
// theta = exp.(w)/sum(exp.(w)); NaN anywhere => the raw weights; running sums for the cumulative walk
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_base_prep(ConfigDev cfg, const double *w, double *th, double *cw, double *tot)
{
    const int g = blockIdx.x, Np = cfg.Np;
    const double *wg = w + (size_t)g * Np;
    double *tg = th + (size_t)g * Np, *cg = cw + (size_t)g * Np;
    __shared__ double s_tot;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) tg[i] = exp(wg[i]);
    __syncthreads();
    if (threadIdx.x == 0) { ksum_t k = { 0.0, 0.0 }; for (int i = 0; i < Np; ++i) kadd(k, tg[i]); s_tot = kval(k); }
    __syncthreads();
    bool bad = false;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) { const double v = tg[i] / s_tot; tg[i] = v; bad |= (v != v); }
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    if (threadIdx.x == 0) {
        const double *src = any_bad ? wg : tg;
        ksum_t k = { 0.0, 0.0 };
        for (int i = 0; i < Np; ++i) kadd(k, src[i]);
        tot[g] = kval(k);
        double c = src[0];
        cg[0] = c;
        for (int i = 1; i < Np; ++i) { c += src[i]; cg[i] = c; }
    }
}

int launch_base_prep(const ConfigDev &cfg, const double *w, double *th, double *cw, double *tot)
{
    k_base_prep<<<cfg.G_local, 128, 0, stream()>>>(cfg, w, th, cw, tot);
    LAUNCHED("k_base_prep");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// pointwise likelihood kernels: block = PW_TP particles x one observation split
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(PW_THREADS) k_ll_pointwise(ModelDev m, const double *theta, Level lv, double *part)
{
    constexpr int NPAR = MAX_ACC + 4;
    __shared__ double par[PW_TP][NPAR];
    __shared__ double red[PW_THREADS / 32][PW_TP];
    const int tile = blockIdx.x, split = blockIdx.y, tid = threadIdx.x;
    const int nt = min(PW_TP, lv.n - tile * PW_TP);
    const int n_split = m.n_osplit * m.n_ksplit;
    // stage the tile's parameters (+ per-particle constants)
    if (tid < nt) {
        const int p = lv.order ? (int)((uint32_t)lv.order[tile * PW_TP + tid] & LV_POS_MASK) : tile * PW_TP + tid;
        const double *th = theta + (size_t)p * m.d;
        if (KIND == M_GAUSSIAN) { par[tid][0] = th[0]; par[tid][1] = th[1]; par[tid][2] = log(th[1]); }
        else if (KIND == M_LNR) { for (int r = 0; r <= m.n_dim; ++r) par[tid][r] = th[r]; }
        else {
            double pneg = 1.0;
            for (int r = 0; r < m.n_dim; ++r) { par[tid][r] = th[r]; pneg *= norm_cdf(-th[r]); }
            par[tid][m.n_dim] = th[m.n_dim]; par[tid][m.n_dim + 1] = th[m.n_dim + 1]; par[tid][m.n_dim + 2] = th[m.n_dim + 2];
            par[tid][m.n_dim + 3] = 1.0 / (1.0 - pneg);
        }
    }
    __syncthreads();
    double acc[PW_TP];
#pragma unroll
    for (int t = 0; t < PW_TP; ++t) acc[t] = 0.0;
    const int64_t i0 = (int64_t)split * m.split_len;
    const int64_t i1 = min(m.n_obs, i0 + (int64_t)m.split_len);
    const double *sg = m.has_sigma ? m.sigma_acc : nullptr;
    for (int64_t i = i0 + tid; i < i1; i += PW_THREADS) {
        const double x = m.x[i];
        const int c = (KIND == M_GAUSSIAN) ? 0 : m.choice[i] - 1;
#pragma unroll
        for (int t = 0; t < PW_TP; ++t) {
            if (t < nt) {
                if (KIND == M_GAUSSIAN) acc[t] += gaussian_obs(par[t], x);
                else if (KIND == M_LNR) acc[t] += lnr_obs(par[t], m.n_dim, sg, x, c);
                else acc[t] += lba_obs(par[t], m.n_dim, par[t][m.n_dim + 3], m.lba_floor, x, c);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < PW_TP; ++t) {
        const double v = warp_sum(acc[t]);
        if ((tid & 31) == 0) red[tid >> 5][t] = v;
    }
    __syncthreads();
    if (tid < nt) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < PW_THREADS / 32; ++w) v += red[w][tid];
        const int p = lv.order ? (int)((uint32_t)lv.order[tile * PW_TP + tid] & LV_POS_MASK) : tile * PW_TP + tid;
        part[(size_t)p * n_split + split] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// MVN / hierarchical likelihood kernel.  With centred data x' and centred means m',
//   sum_i sum_k (x_ik - m_pk)^2 = sum x'^2 - 2 B_p + n sum_k m'_pk^2,   B_p = sum_i sum_k x'_ik m'_pk
// and only B_p needs the O(N d) pass: ONE DFMA per (observation, dimension, particle), every
// observation streamed for every particle (no sufficient-statistic shortcut).  Zero padding
// contributes nothing to B, so there is no ragged-tile path.
//
// CTA = 128 threads, tile 64 particles x 64 observations, thread tile 4 particles x 8 observations
// (32 independent DFMA chains).  The proposal kernel leaves the centred means of every particle
// tile of the level in the layout this kernel wants (mT[tile][k][64]); one TMA bulk copy brings
// them to shared memory where they stay resident, while observation tiles stream through a
// 3-stage ring of [32 dims][64 obs] filled by TMA bulk copies (cp.async.bulk, one 512 B row per
// lane of warp 0) that complete on per-stage mbarriers; consumer warps release a stage through an
// "empty" mbarrier, so there is no CTA-wide barrier and no address arithmetic in the inner loop.  Shared-memory reads are conflict-free: the 8 lanes of an
// observation group read one contiguous 128 B row segment (the 4 particle groups of the warp
// broadcast), the 4 particle groups read 4 x 32 B of one mean row.
// The launch is ONE wave: every (particle tile, dimension split) gets C = slots / items CTAs, each
// taking a contiguous, balanced range of observation SLICES and writing one partial per slice, so
// the set of partial sums (and the summation order) is fixed by the model alone, independent of
// the level size and of the GPU count.
// ------------------------------------------------------------------------------------------------
constexpr int XD_THREADS = 128;
constexpr int XD_STAGES = 3;
constexpr int XD_CTAS_PER_SM = 3;

// ---- TMA bulk copy + mbarrier helpers (sm_90+; SASS: UBLKCP / SYNCS) -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // try_wait sleeps in hardware between polls; the bound turns a protocol bug into a trap, not a hang
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

static size_t xdot_smem_bytes(int klen)
{
    return sizeof(double) * ((size_t)klen * SSD_TP + (size_t)XD_STAGES * SSD_KC * SSD_TN) + sizeof(int) * SSD_TP +
           sizeof(uint64_t) * (2 * XD_STAGES + 1);
}

// staging buffer of centred means, [tile][ssd_k][64], one per device, grown on demand
static double *g_mT[64] = { nullptr };
static size_t g_mT_cap[64] = { 0 };
static double *mT_buffer(size_t doubles)
{
    if (doubles > g_mT_cap[g_dev]) {
        cudaStreamSynchronize(stream());
        if (g_mT[g_dev]) cudaFree(g_mT[g_dev]);
        g_mT[g_dev] = nullptr; g_mT_cap[g_dev] = 0;
        if (cudaMalloc(&g_mT[g_dev], sizeof(double) * doubles) != cudaSuccess) { g_be_err = "cudaMalloc(mean staging)"; return nullptr; }
        g_mT_cap[g_dev] = doubles;
    }
    return g_mT[g_dev];
}
static size_t mT_doubles(const ModelDev &m, int n) { return (size_t)((n + SSD_TP - 1) / SSD_TP) * m.ssd_k * SSD_TP; }

__global__ void __launch_bounds__(XD_THREADS, XD_CTAS_PER_SM) k_xdot(ModelDev m, const double *mT, Level lv, double *part, int C)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, to = tid & 7, tp = tid >> 3, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x / C, c_in = blockIdx.x - tile * C, ksplit = blockIdx.y;
    const int k_begin = ksplit * m.ksplit_len, k_end = min(m.ssd_k, k_begin + m.ksplit_len), klen = k_end - k_begin;
    double *ms = reinterpret_cast<double *>(smem_raw);                 // [klen][64] centred means
    double *xs = ms + (size_t)m.ksplit_len * SSD_TP;                   // [XD_STAGES][SSD_KC][SSD_TN]
    int *s_p = reinterpret_cast<int *>(xs + XD_STAGES * SSD_KC * SSD_TN);
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_p + SSD_TP);       // full[STAGES], empty[STAGES], means
    uint64_t *full = bars, *empty = bars + XD_STAGES, *bar_ms = bars + 2 * XD_STAGES;
    const int nt = min(SSD_TP, lv.n - tile * SSD_TP);
    const int n_split = m.n_osplit * m.n_ksplit;
    const int n_tiles = (int)(m.ssd_ld / SSD_TN), tps = m.ssd_tps;
    const int slice0 = (int)((int64_t)c_in * m.n_osplit / C), slice1 = (int)((int64_t)(c_in + 1) * m.n_osplit / C);
    const int T0 = slice0 * tps, T1 = min(n_tiles, slice1 * tps);
    const int n_kc = (klen + SSD_KC - 1) / SSD_KC;
    const int n_steps = (T1 - T0) * n_kc;
    if (n_steps <= 0) return;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < XD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], XD_THREADS / 32); }
        mbar_init(bar_ms, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < SSD_TP) s_p[tid] = tid < nt ? (lv.order ? (int)((uint32_t)lv.order[tile * SSD_TP + tid] & LV_POS_MASK) : tile * SSD_TP + tid) : -1;
    __syncthreads();

    // producer = warp 0: one bulk copy per row of the stage ([kc] rows of 64 observations = 512 B)
    auto produce = [&](int q) {
        if (q >= n_steps) return;
        const int st = q % XD_STAGES, use = q / XD_STAGES;
        mbar_wait(&empty[st], (use & 1) ^ 1);               // every warp released the previous use of the stage
        const int tt = T0 + q / n_kc, c = q - (q / n_kc) * n_kc;
        const int kc = min(SSD_KC, klen - c * SSD_KC);
        if (lane == 0) mbar_expect_tx(&full[st], (uint32_t)(kc * SSD_TN * sizeof(double)));
        __syncwarp();
        if (lane < kc)
            bulk_g2s(xs + ((size_t)st * SSD_KC + lane) * SSD_TN,
                     m.xT + (size_t)(k_begin + c * SSD_KC + lane) * m.ssd_ld + (size_t)tt * SSD_TN, SSD_TN * sizeof(double), &full[st]);
    };
    if (warp == 0) {
        if (lane == 0) {                                    // the tile's [klen][64] block of centred means
            mbar_expect_tx(bar_ms, (uint32_t)(klen * SSD_TP * sizeof(double)));
            bulk_g2s(ms, mT + ((size_t)tile * m.ssd_k + k_begin) * SSD_TP, (uint32_t)(klen * SSD_TP * sizeof(double)), bar_ms);
        }
        produce(0);
        produce(1);
    }

    double acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;

    mbar_wait(bar_ms, 0);
    int tt = T0, c = 0;                                      // (observation tile, dimension chunk) of step q
    for (int q = 0; q < n_steps; ++q) {
        const int st = q % XD_STAGES;
        if (warp == 0) produce(q + 2);
        mbar_wait(&full[st], (q / XD_STAGES) & 1);
        const int kc = min(SSD_KC, klen - c * SSD_KC);
        const double *xb = xs + (size_t)st * SSD_KC * SSD_TN + to * 2;
        const double *mb = ms + (size_t)(c * SSD_KC) * SSD_TP + tp * 4;
#pragma unroll 4
        for (int kk = 0; kk < kc; ++kk) {
            const double2 x0 = *reinterpret_cast<const double2 *>(xb + kk * SSD_TN);
            const double2 x1 = *reinterpret_cast<const double2 *>(xb + kk * SSD_TN + 16);
            const double2 x2 = *reinterpret_cast<const double2 *>(xb + kk * SSD_TN + 32);
            const double2 x3 = *reinterpret_cast<const double2 *>(xb + kk * SSD_TN + 48);
            const double2 m0 = *reinterpret_cast<const double2 *>(mb + kk * SSD_TP);
            const double2 m1 = *reinterpret_cast<const double2 *>(mb + kk * SSD_TP + 2);
            const double xv[8] = { x0.x, x0.y, x1.x, x1.y, x2.x, x2.y, x3.x, x3.y };
            const double mv[4] = { m0.x, m0.y, m1.x, m1.y };
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fma(xv[b], mv[a], acc[a][b]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);              // this warp is done reading the stage
        // end of a slice: reduce the 8 observation columns and the 8 lanes of the particle group,
        // write the slice's partial, restart the accumulators
        if (c == n_kc - 1) {
            if ((tt + 1) % tps == 0 || tt + 1 == T1) {
                const int slice = tt / tps;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    double v = ((acc[a][0] + acc[a][1]) + (acc[a][2] + acc[a][3])) + ((acc[a][4] + acc[a][5]) + (acc[a][6] + acc[a][7]));
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (to == 0) {
                        const int p = s_p[tp * 4 + a];
                        if (p >= 0) part[(size_t)p * n_split + (size_t)slice * m.n_ksplit + ksplit] = v;
                    }
#pragma unroll
                    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;
                }
            }
            c = 0; ++tt;
        } else {
            ++c;
        }
    }
}

// centred means of arbitrary parameter vectors in the k_xdot layout (demcmc_eval, initial weights)
__global__ void __launch_bounds__(PA_THREADS) k_stage_means(ModelDev m, const double *theta, int64_t n, double *mT)
{
    const int64_t wi = ((int64_t)blockIdx.x * PA_THREADS + threadIdx.x) >> 5;
    if (wi >= n) return;
    const double *th = theta + (size_t)wi * m.d;
    const size_t tile = (size_t)(wi / SSD_TP);
    const int pi = (int)(wi % SSD_TP);
    for (int k = threadIdx.x & 31; k < m.ssd_k; k += 32) mT[(tile * m.ssd_k + k) * SSD_TP + pi] = centred_mean(m, th, k);
}

static int n_sms()
{
    static int sms[64] = { 0 };
    if (!sms[g_dev]) cudaDeviceGetAttribute(&sms[g_dev], cudaDevAttrMultiProcessorCount, g_dev);
    return sms[g_dev] > 0 ? sms[g_dev] : 148;
}

static int launch_xdot(const ModelDev &m, const double *mT, const Level &lv, double *ll_part)
{
    static bool attr_set[64] = { false };
    const size_t smem = xdot_smem_bytes(m.ksplit_len);
    if (!attr_set[g_dev]) {
        CU(cudaFuncSetAttribute(k_xdot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xdot_smem_bytes(SSD_KS)));
        attr_set[g_dev] = true;
    }
    const int n_pt = (lv.n + SSD_TP - 1) / SSD_TP;
    const int64_t items = (int64_t)n_pt * m.n_ksplit, slots = (int64_t)XD_CTAS_PER_SM * n_sms();
    int C = (int)(slots / items);
    C = C < 1 ? 1 : (C > m.n_osplit ? m.n_osplit : C);
    dim3 grid((unsigned)(n_pt * C), (unsigned)m.n_ksplit);
    k_xdot<<<grid, XD_THREADS, smem, stream()>>>(m, mT, lv, ll_part, C);
    LAUNCHED("k_xdot");
    return 0;
}

int launch_loglik(const ConfigDev &cfg, const ModelDev &m, const double *theta, const Level &lv, double *ll_part)
{
    (void)cfg;
    if (lv.n <= 0 || m.kind == M_BINOMIAL) return 0;
    if (m.kind == M_MVNORMAL || m.kind == M_HIER) {
        // the proposal kernel of this level has staged the centred means
        if (!g_mT[g_dev] || g_mT_cap[g_dev] < mT_doubles(m, lv.n)) { g_be_err = "mean staging buffer missing"; return -1; }
        return launch_xdot(m, g_mT[g_dev], lv, ll_part);
    }
    dim3 grid((lv.n + PW_TP - 1) / PW_TP, m.n_osplit);
    if (m.kind == M_GAUSSIAN) k_ll_pointwise<M_GAUSSIAN><<<grid, PW_THREADS, 0, stream()>>>(m, theta, lv, ll_part);
    else if (m.kind == M_LNR) k_ll_pointwise<M_LNR><<<grid, PW_THREADS, 0, stream()>>>(m, theta, lv, ll_part);
    else if (m.kind == M_LBA) k_ll_pointwise<M_LBA><<<grid, PW_THREADS, 0, stream()>>>(m, theta, lv, ll_part);
    else { g_be_err = "no kernel for this model kind"; return -1; }
    LAUNCHED("k_ll_pointwise");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// data packing for k_xdot: x[n][k] (MVN, observation-major) or y[k][n] (hierarchical,
// subject-major) -> centred xT[k][ld] zero padded, center[k], and sum of the squared centred data.
// One block per dimension, fixed reduction order (deterministic).
// ------------------------------------------------------------------------------------------------
__device__ double block_sum_256(double v, double *red)
{
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) { for (int w = 0; w < 8; ++w) s += red[w]; red[8] = s; }
    __syncthreads();
    s = red[8];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(256) k_pack_center(const double *x, double *xT, double *center, double *colsq, int64_t n, int k,
                                                     int64_t ld, int obs_major)
{
    __shared__ double red[9];
    const int kk = blockIdx.x;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) s += obs_major ? x[i * k + kk] : x[(int64_t)kk * n + i];
    const double c = n > 0 ? block_sum_256(s, red) / (double)n : 0.0;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < ld; i += 256) {
        double v = 0.0;
        if (i < n) { v = (obs_major ? x[i * k + kk] : x[(int64_t)kk * n + i]) - c; q += v * v; }
        xT[(int64_t)kk * ld + i] = v;
    }
    q = block_sum_256(q, red);
    if (threadIdx.x == 0) { center[kk] = c; colsq[kk] = q; }
}

int launch_pack_ssd(const double *x_in, int in_on_device, ModelDev *m)
{
    const size_t bytes = sizeof(double) * (size_t)m->ssd_n * m->ssd_k;
    const double *src = x_in;
    double *tmp = nullptr, *colsq = (double *)dmalloc(sizeof(double) * m->ssd_k);
    if (!colsq) return -1;
    if (!in_on_device) {
        tmp = (double *)dmalloc(bytes);
        if (!tmp) { dfree(colsq); return -1; }
        if (h2d(tmp, x_in, bytes)) { dfree(tmp); dfree(colsq); return -1; }
        src = tmp;
    }
    k_pack_center<<<m->ssd_k, 256, 0, stream()>>>(src, const_cast<double *>(m->xT), const_cast<double *>(m->center), colsq, m->ssd_n,
                                                  m->ssd_k, m->ssd_ld, m->kind == M_MVNORMAL ? 1 : 0);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    double *h = new double[m->ssd_k];
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, colsq, sizeof(double) * m->ssd_k, cudaMemcpyDeviceToHost, stream());
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream());
    double xx = 0.0;
    for (int k = 0; k < m->ssd_k; ++k) xx += h[k];
    m->ssd_xx = xx;
    delete[] h;
    dfree(tmp); dfree(colsq);
    if (e != cudaSuccess) return cu_fail(e, "k_pack_center");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// evaluation of arbitrary parameter vectors (init_particle weights, demcmc_eval)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PA_THREADS) k_eval_finish(ConfigDev cfg, ModelDev m, const double *theta, int64_t n,
                                                             const double *part, double *ll, double *prior, double *w)
{
    const int64_t wi = ((int64_t)blockIdx.x * PA_THREADS + threadIdx.x) >> 5;
    if (wi >= n) return;
    const WarpLanes co;
    const double *th = theta + (size_t)wi * cfg.d;
    bool inb; double pr;
    bounds_and_prior(co, cfg, m, th, inb, pr);
    const int n_split = m.n_osplit * m.n_ksplit;
    double s = 0.0;
    if (m.kind != M_BINOMIAL) for (int q = co.lane(); q < n_split; q += 32) s += part[(size_t)wi * n_split + q];
    s = co.sum(s);
    const double l = finalize_ll(m, th, s, mean_sq(co, m, th));
    if (co.lane() == 0) {
        if (ll) ll[wi] = l;
        if (prior) prior[wi] = inb ? pr : -inf();
        if (w) w[wi] = inb ? add(pr, l) : -inf();
    }
}

int launch_eval(const ConfigDev &cfg, const ModelDev &m, const double *theta, int64_t n, double *ll, double *prior,
                double *w, double *scratch_part)
{
    if (n <= 0) return 0;
    Level lv; lv.order = nullptr; lv.n = (int32_t)n; lv.ctxs = nullptr;
    const int blocks = (int)((n * 32 + PA_THREADS - 1) / PA_THREADS);
    if (m.kind == M_MVNORMAL || m.kind == M_HIER) {
        double *mT = mT_buffer(mT_doubles(m, (int)n));
        if (!mT) return -1;
        k_stage_means<<<blocks, PA_THREADS, 0, stream()>>>(m, theta, n, mT);
        LAUNCHED("k_stage_means");
        if (launch_xdot(m, mT, lv, scratch_part)) return -1;
    } else if (launch_loglik(cfg, m, theta, lv, scratch_part)) return -1;
    k_eval_finish<<<blocks, PA_THREADS, 0, stream()>>>(cfg, m, theta, n, scratch_part, ll, prior, w);
    LAUNCHED("k_eval_finish");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// migration (migration.jl:11-116)
// ------------------------------------------------------------------------------------------------
// select_particle (migration.jl:89-95): p ~ exp(-w)/sum(exp(-w)); NaN => findmin(w), no draw
__device__ int select_particle_block(const double *w, int Np, double u, double *th /*smem Np*/)
{
    __shared__ double s_tot;
    __shared__ int s_res;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) th[i] = exp(-w[i]);
    __syncthreads();
    if (threadIdx.x == 0) { ksum_t k = { 0.0, 0.0 }; for (int i = 0; i < Np; ++i) kadd(k, th[i]); s_tot = kval(k); }
    __syncthreads();
    bool bad = false;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) { const double v = th[i] / s_tot; th[i] = v; bad |= (v != v); }
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    if (threadIdx.x == 0) {
        int r = 0;
        if (any_bad) {
            for (int i = 0; i < Np; ++i) { if (w[i] != w[i]) { r = i; break; } if (w[i] < w[r]) r = i; }
        } else {
            ksum_t k = { 0.0, 0.0 };
            for (int i = 0; i < Np; ++i) kadd(k, th[i]);
            const double t = u * kval(k);
            double cw = th[0];
            while (cw < t && r < Np - 1) { ++r; cw += th[r]; }
        }
        s_res = r;
    }
    __syncthreads();
    return s_res;
}

// select_base (crossover.jl:282-289), single block; used by demcmc_op_select
__device__ int select_base_block(const double *w, int Np, double u, double *th)
{
    __shared__ double s_tot;
    __shared__ int s_res;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) th[i] = exp(w[i]);
    __syncthreads();
    if (threadIdx.x == 0) { ksum_t k = { 0.0, 0.0 }; for (int i = 0; i < Np; ++i) kadd(k, th[i]); s_tot = kval(k); }
    __syncthreads();
    bool bad = false;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) { const double v = th[i] / s_tot; th[i] = v; bad |= (v != v); }
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    if (threadIdx.x == 0) {
        const double *src = any_bad ? w : th;
        ksum_t k = { 0.0, 0.0 };
        for (int i = 0; i < Np; ++i) kadd(k, src[i]);
        const double t = u * kval(k);
        int r = 0;
        double cw = src[0];
        while (cw < t && r < Np - 1) { ++r; cw += src[r]; }
        s_res = r;
    }
    __syncthreads();
    return s_res;
}

__global__ void __launch_bounds__(128) k_mig_pick(ConfigDev cfg, MigArgs a, const double *w, int32_t *picks)
{
    extern __shared__ double th[];
    const int i = blockIdx.x, gl = a.groups[i] - cfg.group_begin;
    if (gl < 0 || gl >= cfg.G_local) { if (threadIdx.x == 0) picks[i] = -1; return; }
    const int r = select_particle_block(w + (size_t)gl * cfg.Np, cfg.Np, a.u_pick[i], th);
    if (threadIdx.x == 0) picks[i] = r;
}

// The whole Particle object migrates (theta, weight, id and its accept/lp history): the staging row
// is {theta[d], weight, id, accept flag of the row being edited}
__global__ void __launch_bounds__(128) k_mig_gather(ConfigDev cfg, MigArgs a, const int32_t *picks, const double *theta,
                                                    const double *w, const int32_t *id, const uint8_t *acc, double *stage)
{
    const int i = blockIdx.x, gl = a.groups[i] - cfg.group_begin;
    if (gl < 0 || gl >= cfg.G_local) return;
    const size_t p = (size_t)gl * cfg.Np + picks[i];
    double *row = stage + (size_t)i * (cfg.d + 3);
    for (int k = threadIdx.x; k < cfg.d; k += blockDim.x) row[k] = theta[p * cfg.d + k];
    if (threadIdx.x == 0) { row[cfg.d] = w[p]; row[cfg.d + 1] = (double)id[p]; row[cfg.d + 2] = (double)acc[p]; }
}

// shift_particles! (migration.jl:109-116): position i receives the particle picked at position i-1
__global__ void __launch_bounds__(128) k_mig_scatter(ConfigDev cfg, MigArgs a, const int32_t *picks, const double *stage,
                                                     double *theta, double *w, int32_t *id, uint8_t *acc)
{
    const int i = blockIdx.x, gl = a.groups[i] - cfg.group_begin;
    if (gl < 0 || gl >= cfg.G_local) return;
    const size_t p = (size_t)gl * cfg.Np + picks[i];
    const double *row = stage + (size_t)((i + a.n - 1) % a.n) * (cfg.d + 3);
    for (int k = threadIdx.x; k < cfg.d; k += blockDim.x) theta[p * cfg.d + k] = row[k];
    if (threadIdx.x == 0) { w[p] = row[cfg.d]; id[p] = (int32_t)row[cfg.d + 1]; acc[p] = (uint8_t)row[cfg.d + 2]; }
}

int launch_mig_pick(const ConfigDev &cfg, const MigArgs &a, const double *w, int32_t *picks)
{
    k_mig_pick<<<a.n, 128, sizeof(double) * cfg.Np, stream()>>>(cfg, a, w, picks);
    LAUNCHED("k_mig_pick");
    return 0;
}
int launch_mig_gather(const ConfigDev &cfg, const MigArgs &a, const int32_t *picks, const double *theta, const double *w,
                      const int32_t *id, const uint8_t *acc, double *stage)
{
    k_mig_gather<<<a.n, 128, 0, stream()>>>(cfg, a, picks, theta, w, id, acc, stage);
    LAUNCHED("k_mig_gather");
    return 0;
}
int launch_mig_scatter(const ConfigDev &cfg, const MigArgs &a, const int32_t *picks, const double *stage, double *theta,
                       double *w, int32_t *id, uint8_t *acc)
{
    k_mig_scatter<<<a.n, 128, 0, stream()>>>(cfg, a, picks, stage, theta, w, id, acc);
    LAUNCHED("k_mig_scatter");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// history: rows by slot -> samples[P][d][n_rows] / lp[P][n_rows] / accept[P][n_rows] by particle id
// (the memory order of Julia's Array{T,3}(n_rows, d, P), utilities.jl:34).  Lanes run along rows so
// the writes coalesce.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_history(const double *rt, const double *rw, const uint8_t *ra, const int32_t *rid,
                                                 int64_t n_rows_dev, int64_t row0, int64_t n_rows_out, int P, int d, int id_base,
                                                 double *samples, double *lp, uint8_t *accept)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int slot = blockIdx.y;
    if (r >= n_rows_dev) return;
    const int id = rid[r * P + slot] - id_base;
    if (id < 0 || id >= P) return;
    const int64_t ro = row0 + r;
    if (samples) for (int k = 0; k < d; ++k) samples[((int64_t)id * d + k) * n_rows_out + ro] = rt[(r * P + slot) * d + k];
    if (lp) lp[(int64_t)id * n_rows_out + ro] = rw[r * P + slot];
    if (accept) accept[(int64_t)id * n_rows_out + ro] = ra[r * P + slot];
}

int launch_history_by_id(const double *rows_theta, const double *rows_w, const uint8_t *rows_acc, const int32_t *rows_id,
                         int64_t n_rows_dev, int64_t row0, int64_t n_rows_out, int32_t P, int32_t d, int32_t id_base,
                         double *samples, double *lp, uint8_t *accept)
{
    dim3 grid((unsigned)((n_rows_dev + 255) / 256), (unsigned)P);
    k_history<<<grid, 256, 0, stream()>>>(rows_theta, rows_w, rows_acc, rows_id, n_rows_dev, row0, n_rows_out, P, d, id_base, samples, lp, accept);
    LAUNCHED("k_history");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// particle algebra ops (single warp)
// ------------------------------------------------------------------------------------------------
__global__ void k_op_project(const double *p1, const double *p2, int d, double *out)
{
    const int lane = threadIdx.x;
    double v1 = 0.0, v2 = 0.0;
    for (int k = lane; k < d; k += 32) { v1 = add(v1, mul(p1[k], p2[k])); v2 = add(v2, mul(p2[k], p2[k])); }
    v1 = warp_sum(v1); v2 = warp_sum(v2);
    const double r = v1 / v2;
    for (int k = lane; k < d; k += 32) out[k] = mul(p2[k], r);
}
__global__ void k_op_snooker(const double *pt, const double *pz, const double *pm, const double *pn, double g,
                             const double *b, int d, double *out, double *log_adj)
{
    const int lane = threadIdx.x;
    double v1m = 0.0, v1n = 0.0, v2 = 0.0;
    for (int k = lane; k < d; k += 32) {
        const double pd = sub(pt[k], pz[k]);
        v1m = add(v1m, mul(pm[k], pd)); v1n = add(v1n, mul(pn[k], pd)); v2 = add(v2, mul(pd, pd));
    }
    v1m = warp_sum(v1m); v1n = warp_sum(v1n); v2 = warp_sum(v2);
    const double r1 = v1m / v2, r2 = v1n / v2;
    double sq1 = 0.0, sq2 = 0.0;
    for (int k = lane; k < d; k += 32) {
        const double v = snooker_elem(pt[k], pz[k], r1, r2, g, b[k]);
        out[k] = v;
        const double a = sub(v, pz[k]), c = sub(pt[k], pz[k]);
        sq1 = add(sq1, mul(a, a)); sq2 = add(sq2, mul(c, c));
    }
    sq1 = warp_sum(sq1); sq2 = warp_sum(sq2);
    if (lane == 0) *log_adj = adjust_loglike(sq1, sq2, d);
}
__global__ void k_op_de(const double *pt, const double *pm, const double *pn, const double *pb, double g1, double g2,
                        const double *b, int d, double *out)
{
    for (int k = threadIdx.x; k < d; k += 32) out[k] = de_elem(pt[k], pm[k], pn[k], pb ? pb[k] : pt[k], g1, g2, pb != nullptr, b[k]);
}
__global__ void k_op_reset(const double *prop, const double *pt, const uint8_t *mask, int d, double *out)
{
    for (int k = threadIdx.x; k < d; k += 32) out[k] = mask[k] ? prop[k] : pt[k];
}
__global__ void k_op_accept(const double *wp, const double *wc, const double *adj, const double *u, int n, uint8_t *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = accept(wp[i], wc[i], adj[i], u[i]) ? 1 : 0;
}
__global__ void __launch_bounds__(128) k_op_select(const double *w, int n, double u, int32_t *base_idx, int32_t *mig_idx)
{
    extern __shared__ double th[];
    const int b = select_base_block(w, n, u, th);
    __syncthreads();
    const int q = select_particle_block(w, n, u, th);
    if (threadIdx.x == 0) { *base_idx = b; *mig_idx = q; }
}

int launch_op_project(const double *p1, const double *p2, int d, double *out) { k_op_project<<<1, 32, 0, stream()>>>(p1, p2, d, out); LAUNCHED("k_op_project"); return 0; }
int launch_op_snooker(const double *pt, const double *pz, const double *pm, const double *pn, double g, const double *b, int d, double *out, double *log_adj)
{ k_op_snooker<<<1, 32, 0, stream()>>>(pt, pz, pm, pn, g, b, d, out, log_adj); LAUNCHED("k_op_snooker"); return 0; }
int launch_op_de(const double *pt, const double *pm, const double *pn, const double *pb, double g1, double g2, const double *b, int d, double *out)
{ k_op_de<<<1, 32, 0, stream()>>>(pt, pm, pn, pb, g1, g2, b, d, out); LAUNCHED("k_op_de"); return 0; }
int launch_op_reset(const double *prop, const double *pt, const uint8_t *mask, int d, double *out) { k_op_reset<<<1, 32, 0, stream()>>>(prop, pt, mask, d, out); LAUNCHED("k_op_reset"); return 0; }
int launch_op_accept(const double *wp, const double *wc, const double *adj, const double *u, int n, uint8_t *out)
{ k_op_accept<<<(n + 127) / 128, 128, 0, stream()>>>(wp, wc, adj, u, n, out); LAUNCHED("k_op_accept"); return 0; }
int launch_op_select(const double *w, int n, double u, int32_t *base_idx, int32_t *mig_idx)
{ k_op_select<<<1, 128, sizeof(double) * n, stream()>>>(w, n, u, base_idx, mig_idx); LAUNCHED("k_op_select"); return 0; }

// ------------------------------------------------------------------------------------------------
// roofline probes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b)
{
    double r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fma(r[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 123.456) out[0] = s;
}

int fp64_peak(double *tflops)
{
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g_dev));
    double *out = (double *)dmalloc(8);
    if (!out) return -1;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int iters = 1 << 14, blocks = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CU(cudaEventRecord(e0, stream()));
        k_dfma_peak<<<blocks, 256, 0, stream()>>>(out, iters, 0.999999, 1e-9);
        ++g_launches;
        CU(cudaEventRecord(e1, stream()));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
        if (rep > 0) best = fmax(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    dfree(out);
    *tflops = best;
    return 0;
}

int copy_peak(double *gbs)
{
    const size_t n = (size_t)1 << 30;
    void *a = dmalloc(n), *b = dmalloc(n);
    if (!a || !b) { dfree(a); dfree(b); return -1; }
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CU(cudaEventRecord(e0, stream()));
        CU(cudaMemcpyAsync(b, a, n, cudaMemcpyDeviceToDevice, stream()));
        CU(cudaEventRecord(e1, stream()));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = fmax(best, 2.0 * (double)n / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    dfree(a); dfree(b);
    *gbs = best;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// cross-rank migration: NCCL send/recv over NVLink, resolved at run time from the process's libnccl
// ------------------------------------------------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.lib) return 0;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
    if (!g_nccl.lib) { g_be_err = std::string("cannot load libnccl: ") + dlerror(); return -1; }
#define SYM(field, sym) do { *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, sym); if (!g_nccl.field) { g_be_err = std::string("libnccl misses ") + sym; g_nccl.lib = nullptr; return -1; } } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return 0;
}
#define NC(call) do { int r_ = (call); if (r_ != 0) { g_be_err = std::string(#call) + ": " + g_nccl.GetErrorString(r_); return -1; } } while (0)

int comm_unique_id(uint8_t id[128])
{
    if (nccl_load()) return -1;
    nccl_uid u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return 0;
}
int comm_init(const uint8_t id[128], int rank, int n_ranks, void **comm)
{
    if (nccl_load()) return -1;
    nccl_uid u;
    memcpy(u.internal, id, 128);
    nccl_comm c = nullptr;
    NC(g_nccl.CommInitRank(&c, n_ranks, u, rank));
    *comm = c;
    return 0;
}
int comm_destroy(void *comm)
{
    if (comm && g_nccl.lib) g_nccl.CommDestroy((nccl_comm)comm);
    return 0;
}
int comm_exchange(void *comm, int rank, int n, const int *src_rank, const int *dst_rank, double *stage_send,
                  double *stage_recv, int row_len)
{
    // position i consumes row r = i-1 (cyclic), produced on src_rank[i], consumed on dst_rank[i]
    NC(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
        if (src_rank[i] == dst_rank[i]) continue;
        const int r = (i + n - 1) % n;
        if (rank == src_rank[i]) NC(g_nccl.Send(stage_send + (size_t)r * row_len, (size_t)row_len, 8 /* ncclFloat64 */, dst_rank[i], (nccl_comm)comm, stream()));
        if (rank == dst_rank[i]) NC(g_nccl.Recv(stage_recv + (size_t)r * row_len, (size_t)row_len, 8 /* ncclFloat64 */, src_rank[i], (nccl_comm)comm, stream()));
    }
    NC(g_nccl.GroupEnd());
    return 0;
}

} // namespace be
} // namespace de
