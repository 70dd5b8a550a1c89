// kernels.cu -- CUDA (sm_100a) implementation of backend.h: every kernel of libdemcmc_b200.
//
//   k_propose      one warp per particle: DE / snooker / mutation proposal, kappa and block masks,
//                  bounds, prior, snooker adjustment                       (HBM-bound, ~5 d-vectors)
//   k_xdot         cross term sum_i sum_k x'_ik m'_pk of the expanded sum of squares for a tile of
//                  particles against a range of observation tiles, on the fp64 tensor path
//                  (DMMA m8n8k4): the likelihood of the isotropic MVN and the hierarchical
//                  normal models                                           (fp64-pipe-bound)
//   k_ll_pointwise per-observation log densities (Gaussian, LNR, LBA) for a tile of particles
//   k_accept       one warp per particle: fixed-order reduction of the partial sums, Metropolis
//                  accept, state-row write (replaces store_samples!)       (HBM-bound)
//   k_mig_*        migration picks and the cyclic shift
//   k_history      by-slot rows -> the reference's samples[n_rows, d, P] layout
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "backend.h"
#include "de_particle.h"

namespace de {
namespace be {

static thread_local std::string g_be_err;
static int64_t g_launches = 0;
static int g_dev = 0;
// Lanes: independent sets of groups run as concurrent kernel chains, one stream per lane, so the
// likelihood kernel of one lane covers the propose / accept latency of the other (engine.cpp).
// Lane 0 is the engine stream; every copy, event and one-off kernel runs there.
static int g_lane = 0;
static cudaStream_t g_stream[64][MAX_LANES] = { { nullptr } };
static cudaEvent_t g_lane_ev[64][MAX_LANES] = { { nullptr } };
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
    if (!g_stream[g_dev][g_lane]) cudaStreamCreateWithFlags(&g_stream[g_dev][g_lane], cudaStreamNonBlocking);
    return g_stream[g_dev][g_lane];
}

// Kernels of one level are chained with programmatic dependent launch: the next kernel's CTAs may
// become resident and run their state-independent prologue while the previous kernel drains; each
// kernel executes griddepcontrol.wait before it touches anything an earlier kernel wrote.
// DEMCMC_NO_PDL=1 falls back to plain stream order (A/B measurements).
static bool use_pdl()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("DEMCMC_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream();
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = use_pdl() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- debug timeline (DEMCMC_TIMELINE=<levels> DEMCMC_TIMELINE_FILE=<csv>): every kernel of a level
// stamps %globaltimer into one 16-word slot, so the gaps between the kernels of a PDL chain can be
// read without events (which would break the chain) and without a profiler (which serialises it)
enum { TL_P0 = 0, TL_P1, TL_X0, TL_X0MAX, TL_XWAIT, TL_XFIRST, TL_XLOOP0, TL_XLOOP1, TL_X1, TL_A0, TL_A1, TL_N, TL_WORDS = 16 };
static unsigned long long *g_tl = nullptr;
static int g_tl_cap = -1, g_tl_level = 0, g_tl_cta_level = -1;
constexpr int TL_CTA_WORDS = 4, TL_CTA_MAX = 4096;
static unsigned long long *tl_cta()      // per-CTA stamps of the one level named by DEMCMC_TIMELINE_CTA
{
    return (g_tl_cap > 0 && g_tl && g_tl_level == g_tl_cta_level) ? g_tl + (size_t)TL_WORDS * g_tl_cap : nullptr;
}
static unsigned long long *tl_slot()
{
    if (g_tl_cap < 0) {
        const char *e = getenv("DEMCMC_TIMELINE");
        g_tl_cap = e ? atoi(e) : 0;
        const size_t words = (size_t)TL_WORDS * g_tl_cap + (size_t)TL_CTA_WORDS * TL_CTA_MAX;
        if (g_tl_cap > 0 && (cudaMalloc(&g_tl, sizeof(unsigned long long) * words) != cudaSuccess ||
                             cudaMemset(g_tl, 0, sizeof(unsigned long long) * words) != cudaSuccess)) g_tl_cap = 0;
        if (const char *c = getenv("DEMCMC_TIMELINE_CTA")) g_tl_cta_level = atoi(c);
    }
    return (g_tl_cap > 0 && g_tl_level < g_tl_cap) ? g_tl + (size_t)TL_WORDS * g_tl_level : nullptr;
}
void timeline_dump()
{
    if (g_tl_cap <= 0 || !g_tl) return;
    const char *path = getenv("DEMCMC_TIMELINE_FILE");
    if (!path) return;
    cudaDeviceSynchronize();
    const int n = g_tl_level < g_tl_cap ? g_tl_level : g_tl_cap;
    std::vector<unsigned long long> h((size_t)TL_WORDS * (n > 0 ? n : 1));
    if (n <= 0 || cudaMemcpy(h.data(), g_tl, sizeof(unsigned long long) * TL_WORDS * n, cudaMemcpyDeviceToHost) != cudaSuccess) return;
    FILE *f = fopen(path, "w");
    if (!f) return;
    fprintf(f, "level,n,propose_start,propose_end,xdot_start,xdot_last_start,xdot_wait_done,xdot_first_data,xdot_loop_end_min,xdot_loop_end_max,xdot_end,accept_start,accept_end\n");
    const unsigned long long t0 = ~h[TL_P0];
    for (int i = 0; i < n; ++i) {
        const unsigned long long *w = h.data() + (size_t)TL_WORDS * i;
        auto rel = [&](unsigned long long v) { return (double)((long long)(v - t0)) * 1e-3; };
        fprintf(f, "%d,%llu,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f\n", i, w[TL_N], rel(~w[TL_P0]), rel(w[TL_P1]), rel(~w[TL_X0]), rel(w[TL_X0MAX]),
                rel(w[TL_XWAIT]), rel(w[TL_XFIRST]), rel(~w[TL_XLOOP0]), rel(w[TL_XLOOP1]), rel(w[TL_X1]), rel(~w[TL_A0]), rel(w[TL_A1]));
    }
    fclose(f);
    if (g_tl_cta_level >= 0) {
        std::vector<unsigned long long> c((size_t)TL_CTA_WORDS * TL_CTA_MAX);
        if (cudaMemcpy(c.data(), g_tl + (size_t)TL_WORDS * g_tl_cap, sizeof(unsigned long long) * c.size(), cudaMemcpyDeviceToHost) == cudaSuccess) {
            std::string p2 = std::string(path) + ".cta";
            FILE *g = fopen(p2.c_str(), "w");
            if (g) {
                fprintf(g, "cta,start,wait_done,loop_end,tiles\n");
                unsigned long long c0 = ~0ull;
                for (int i = 0; i < TL_CTA_MAX; ++i) if (c[(size_t)i * 4]) c0 = c[(size_t)i * 4] < c0 ? c[(size_t)i * 4] : c0;
                for (int i = 0; i < TL_CTA_MAX; ++i) if (c[(size_t)i * 4])
                    fprintf(g, "%d,%.3f,%.3f,%.3f,%llu\n", i, (double)(c[(size_t)i * 4] - c0) * 1e-3, (double)(c[(size_t)i * 4 + 1] - c0) * 1e-3, (double)(c[(size_t)i * 4 + 2] - c0) * 1e-3, c[(size_t)i * 4 + 3]);
                fclose(g);
            }
        }
    }
}
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tl_min(unsigned long long *tl, int w) { if (tl) atomicMax(tl + w, ~gtime()); }
__device__ __forceinline__ void tl_max(unsigned long long *tl, int w) { if (tl) atomicMax(tl + w, gtime()); }

void set_lane(int lane) { g_lane = (lane >= 0 && lane < MAX_LANES) ? lane : 0; }
static cudaEvent_t lane_event(int lane)
{
    if (!g_lane_ev[g_dev][lane]) cudaEventCreateWithFlags(&g_lane_ev[g_dev][lane], cudaEventDisableTiming);
    return g_lane_ev[g_dev][lane];
}
int lane_fork(int n_lanes)
{
    if (n_lanes < 2) return 0;
    const int keep = g_lane;
    g_lane = 0;
    cudaStream_t s0 = stream();
    CU(cudaEventRecord(lane_event(0), s0));
    for (int l = 1; l < n_lanes && l < MAX_LANES; ++l) { g_lane = l; CU(cudaStreamWaitEvent(stream(), lane_event(0), 0)); }
    g_lane = keep;
    return 0;
}
int lane_join(int n_lanes)
{
    if (n_lanes < 2) return 0;
    const int keep = g_lane;
    g_lane = 0;
    cudaStream_t s0 = stream();
    for (int l = 1; l < n_lanes && l < MAX_LANES; ++l) {
        g_lane = l;
        CU(cudaEventRecord(lane_event(l), stream()));
        CU(cudaStreamWaitEvent(s0, lane_event(l), 0));
    }
    g_lane = keep;
    return 0;
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
// Device memory comes from the stream-ordered allocator with an unbounded release threshold: a
// handle that is destroyed leaves its blocks in the pool, so creating the next one (one handle per
// sample() call) costs microseconds instead of a cudaMalloc / cudaFree pair per buffer.
static void pool_setup()
{
    static bool done[64] = { false };
    if (done[g_dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, g_dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[g_dev] = true;
}
void *dmalloc(size_t bytes)
{
    pool_setup();
    void *p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 8, stream());
    if (e != cudaSuccess) { cu_fail(e, "cudaMallocAsync"); return nullptr; }
    return p;
}
void dfree(void *p) { if (p) cudaFreeAsync(p, stream()); }
// pinned blocks are recycled through a small free list for the same reason
struct PinnedBlock { void *p; size_t bytes; bool busy; };
static std::vector<PinnedBlock> g_pinned;
void *hmalloc_pinned(size_t bytes)
{
    if (!bytes) bytes = 8;
    for (auto &b : g_pinned)
        if (!b.busy && b.bytes >= bytes && b.bytes <= 2 * bytes + 4096) { b.busy = true; return b.p; }
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) { cu_fail(e, "cudaMallocHost"); return nullptr; }
    g_pinned.push_back({ p, bytes, true });
    return p;
}
void hfree_pinned(void *p)
{
    if (!p) return;
    for (auto &b : g_pinned) if (b.p == p) { b.busy = false; return; }
    cudaFreeHost(p);
}
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
    // programmatic dependent launch: everything before this point touches only data that is constant
    // for the whole chunk (schedule, tape, Philox); the state written by earlier kernels comes after
    __device__ __forceinline__ void dependency_wait() const { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
};
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

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

// per-device staging for k_xdot: the proposals' centred means as DMMA B fragments,
// bfrag[octet of the level][dimension split][k-step j][lane], and the fixed-point magic constant of
// every particle of the level, magic[level order]
struct XdStage { double *bfrag = nullptr; double *magic = nullptr; size_t cap = 0; };
static XdStage g_xs[64][MAX_LANES];
static XdStage *xd_stage(const ModelDev &m, int n);

// leaves one parameter vector's centred means where k_xdot wants them, together with the particle's
// fixed-point scale (de_math.h: xd_scale); wi = rank of the particle in the launch
__device__ __forceinline__ void stage_scale(const ModelDev &m, double msq, int64_t wi, double *magic, long long *acc, double *q, double *msq_out)
{
    if ((threadIdx.x & 31) == 0) {
        const XdScale sc = xd_scale(msq, m.ssd_rowmax, m.ssd_qbits);
        magic[wi] = sc.magic;
        *q = sc.q;
        *acc = 0;
        if (msq_out) *msq_out = msq;
    }
}
__device__ __forceinline__ size_t bfrag_index(const ModelDev &m, int64_t wi, int k)
{
    const int64_t oct = wi / SSD_OCT;
    const int n = (int)(wi % SSD_OCT);
    const int ks = k / m.ksplit_len, kl = k - ks * m.ksplit_len;
    return (((size_t)oct * m.n_ksplit + ks) * m.ssd_nj + (kl >> 2)) * 32 + n * 4 + (kl & 3);
}
__device__ __forceinline__ void stage_bfrag(const ModelDev &m, const double *theta, int64_t wi, double *bfrag, double *magic,
                                            long long *acc, double *q, double *msq_out)
{
    const int lane = threadIdx.x & 31;
    double msq = 0.0;
    for (int k = lane; k < m.ssd_k; k += 32) {
        const double v = centred_mean(m, theta, k);
        msq += v * v;
        bfrag[bfrag_index(m, wi, k)] = v;
    }
    stage_scale(m, warp_sum(msq), wi, magic, acc, q, msq_out);
}

// the MVN model's means ARE proposal elements: stage them while they are still in registers
// (the data centre of the lane's first elements is fetched before the dependency wait)
struct StageSink {
    const ModelDev &m;
    double *bfrag;
    int64_t wi;
    bool on;
    double cen[PROP_PRE];
    double msq;
    __device__ __forceinline__ void prefetch(int q, int k) { if (on && k < m.ssd_k) cen[q] = m.center[k]; }
    __device__ __forceinline__ void elem(int q, int k, double v)
    {
        if (!on || k >= m.ssd_k) return;
        const double c = v - (q < PROP_PRE ? cen[q] : m.center[k]);
        msq += c * c;
        bfrag[bfrag_index(m, wi, k)] = c;
    }
};

__global__ void __launch_bounds__(PA_THREADS) k_propose(ConfigDev cfg, ModelDev m, Level lv, double *bfrag, double *magic, unsigned long long *tl)
{
    pdl_launch_dependents();
    if ((threadIdx.x & 31) == 0) tl_min(tl, TL_P0);
    const int wi = (blockIdx.x * PA_THREADS + threadIdx.x) >> 5;
    if (wi >= lv.n) { pdl_wait(); return; }
    const uint32_t e = (uint32_t)lv.order[wi];
    const SweepCtx ctx = lv.ctxs[e >> LV_SLOT_SHIFT];
    const int p = (int)(e & LV_POS_MASK);
    StageSink sink = { m, bfrag, wi, bfrag != nullptr && m.kind == M_MVNORMAL, { 0.0 }, 0.0 };
    propose_particle(WarpLanes(), cfg, m, ctx, p, sink);
    if (bfrag) {
        if (sink.on) stage_scale(m, warp_sum(sink.msq), wi, magic, ctx.ll_acc + p, ctx.ll_q + p, ctx.prop_msq + p);
        else {
            __syncwarp();
            stage_bfrag(m, ctx.prop_theta + (size_t)p * cfg.d, wi, bfrag, magic, ctx.ll_acc + p, ctx.ll_q + p, ctx.prop_msq + p);
        }
    }
    if ((threadIdx.x & 31) == 0) { tl_max(tl, TL_P1); if (tl && wi == 0) tl[TL_N] = (unsigned long long)lv.n; }
}

__global__ void __launch_bounds__(PA_THREADS) k_accept(ConfigDev cfg, ModelDev m, Level lv, unsigned long long *tl)
{
    pdl_launch_dependents();
    if ((threadIdx.x & 31) == 0) tl_min(tl, TL_A0);
    const int wi = (blockIdx.x * PA_THREADS + threadIdx.x) >> 5;
    if (wi >= lv.n) { pdl_wait(); return; }
    const uint32_t e = (uint32_t)lv.order[wi];
    const SweepCtx ctx = lv.ctxs[e >> LV_SLOT_SHIFT];
    accept_particle(WarpLanes(), cfg, m, ctx, (int)(e & LV_POS_MASK));
    if ((threadIdx.x & 31) == 0) tl_max(tl, TL_A1);
}

int launch_propose(const ConfigDev &cfg, const ModelDev &m, const Level &lv)
{
    const int blocks = (lv.n * 32 + PA_THREADS - 1) / PA_THREADS;
    XdStage *xs = nullptr;
    if (m.kind == M_MVNORMAL || m.kind == M_HIER) {
        // sized once for the handle's whole population so it never grows inside a run
        xs = xd_stage(m, std::max(lv.n, cfg.G_local * cfg.Np));
        if (!xs) return -1;
    }
    CU(launch_chained(k_propose, dim3(blocks), dim3(PA_THREADS), 0, cfg, m, lv, xs ? xs->bfrag : nullptr, xs ? xs->magic : nullptr, tl_slot()));
    LAUNCHED("k_propose");
    return 0;
}

int launch_accept(const ConfigDev &cfg, const ModelDev &m, const Level &lv)
{
    const int blocks = (lv.n * 32 + PA_THREADS - 1) / PA_THREADS;
    CU(launch_chained(k_accept, dim3(blocks), dim3(PA_THREADS), 0, cfg, m, lv, tl_slot()));
    LAUNCHED("k_accept");
    if (g_tl_cap > 0) ++g_tl_level;
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
    constexpr int NPAR = MAX_ACC + 5;
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
            par[tid][m.n_dim + 4] = 1.0 / th[m.n_dim];
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
// and only B_p needs the O(N d) pass: one multiply-add per (observation, dimension, particle), every
// observation streamed for every particle (no sufficient-statistic shortcut).
//
// B = X' M is a GEMM whose rows are summed away, so it runs on the fp64 tensor path: DMMA m8n8k4
// measures 37.1 TFLOP/s on B200 against 33.9 for DFMA (scripts/probes/probe_dmma.cu; the two share
// one pipe, they do not add), and it needs 8x fewer issue slots and no accumulator tile.
//   A (8 observations x 4 dimensions)  = packed centred data, streamed
//   B (4 dimensions x 8 particles)     = centred means of one particle octet, held in REGISTERS for
//                                        the whole kernel (<= 13 k-steps x 4 octets per thread)
//   C (8 observations x 8 particles)   = per-row-pair cross terms; a chain runs over the k-steps
//                                        of ONE observation tile, then is rounded to the
//                                        particle's fixed-point grid (de_math.h: xd_scale) and added
//                                        as an integer, which makes the total independent of how
//                                        observation tiles were dealt to CTAs
// CTA = 4 warps x one particle tile (32 particles = 4 octets).  Warp w owns row pair w (16 of the 64
// observations) of every observation tile in the CTA's range: its operand stream is contiguous in
// the packed layout, so each warp runs a PRIVATE 4-stage ring of TMA bulk copies (cp.async.bulk,
// one copy per stage, completing on the warp's own mbarriers) and the kernel has no CTA-wide
// barrier and no cross-warp wait at all.  Inner step: one LDS.128 (A fragments of two row tiles)
// feeds 8 DMMAs (2 row tiles x 4 octets) on 4 accumulator chains per warp (one per octet; the two
// row tiles of the pair add into the same chain, 4 DMMAs apart, so a tile ends with 8 conversions).
// The launch is one wave: CTAs are dealt to particle tiles in proportion to their octets (the last
// tile of a level may hold 1..4), each taking a balanced contiguous range of observation tiles;
// padding costs at most 7 particles per level.
// ------------------------------------------------------------------------------------------------
constexpr int XD_THREADS = 128;
constexpr int XD_STAGES = 4;
constexpr int XD_CTAS_PER_SM = 2;
constexpr int XD_MIN_TILES = 4;          // observation tiles a CTA should at least stream (amortises its prologue)
constexpr int XD_WAVE_TILES = 48;        // observation tiles per CTA when a level needs several waves

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
// D(8x8) += A(8x4) * B(4x8) in fp64 on the tensor path (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

static size_t xdot_smem_bytes(int nj)
{
    return (size_t)4 * XD_STAGES * nj * 64 * sizeof(double) + sizeof(uint64_t) * 4 * XD_STAGES;
}

static size_t xd_bfrag_doubles(const ModelDev &m, int n) { return (size_t)((n + SSD_OCT - 1) / SSD_OCT) * m.n_ksplit * m.ssd_nj * 32; }
static XdStage *xd_stage(const ModelDev &m, int n)
{
    XdStage &x = g_xs[g_dev][g_lane];
    const size_t need = xd_bfrag_doubles(m, n) + (size_t)((n + SSD_OCT - 1) / SSD_OCT) * SSD_OCT;
    if (need > x.cap) {
        cudaStreamSynchronize(stream());
        if (x.bfrag) cudaFree(x.bfrag);
        x.bfrag = nullptr; x.magic = nullptr; x.cap = 0;
        if (cudaMalloc(&x.bfrag, sizeof(double) * need) != cudaSuccess) { g_be_err = "cudaMalloc(mean staging)"; return nullptr; }
        x.cap = need;
    }
    // the dimension slots that pad a split to whole k-steps must read as zero, and the geometry may
    // change with the model: clear whenever the split between the two arrays moves
    double *magic = x.bfrag + (x.cap - (size_t)((n + SSD_OCT - 1) / SSD_OCT) * SSD_OCT);
    if (magic != x.magic) {
        if (cudaMemsetAsync(x.bfrag, 0, sizeof(double) * x.cap, stream()) != cudaSuccess) { g_be_err = "cudaMemset(mean staging)"; return nullptr; }
        x.magic = magic;
    }
    return &x;
}

// how the CTAs of one launch are dealt to the particle tiles of a level
// n_hi tiles of oct_hi octets with c_hi CTAs each, then n_lo tiles of oct_lo octets with c_lo CTAs each
struct XdGrid { int32_t n_hi, oct_hi, c_hi, n_lo, oct_lo, c_lo; };

template <int NOCT, int NJC>            // NJC: the model's k-steps when known at compile time (13), else 0
__device__ __forceinline__ void xdot_body(const ModelDev &m, const double *bfrag, const double *magic, const Level &lv,
                                          long long *ll_acc, int oct0, int T0, int T1, unsigned char *smem_raw, unsigned long long *tl, unsigned long long *tlc)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, ks = blockIdx.y;
    const int nj = NJC ? NJC : m.ssd_nj;
    const int n_tiles = (int)(m.ssd_ld / SSD_TN);
    const uint32_t stage_doubles = (uint32_t)nj * 64, stage_bytes = stage_doubles * (uint32_t)sizeof(double);
    double *ring = reinterpret_cast<double *>(smem_raw) + (size_t)warp * XD_STAGES * stage_doubles;
    uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(smem_raw) + (size_t)4 * XD_STAGES * stage_doubles) + warp * XD_STAGES;
    const double *src = m.xT + (((size_t)(ks * 4 + warp) * n_tiles + T0) * nj) * 64;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < XD_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
#pragma unroll
        for (int s = 0; s < XD_STAGES; ++s)
            if (T0 + s < T1) {
                mbar_expect_tx(&full[s], stage_bytes);
                bulk_g2s(ring + (size_t)s * stage_doubles, src + (size_t)s * stage_doubles, stage_bytes, &full[s]);
            }
    }
    __syncwarp();
    pdl_wait();                                              // the packed data are constant; the means are not
    if (tid == 0) tl_max(tl, TL_XWAIT);
    if (tlc && tid == 0 && blockIdx.x < TL_CTA_MAX) { tlc[blockIdx.x * 4 + 1] = gtime(); tlc[blockIdx.x * 4 + 3] = (unsigned long long)(T1 - T0) * 10 + NOCT; }

    // this tile's centred means as B fragments, and the particles' fixed-point constants
    double b[SSD_NJ][NOCT];
    {
#pragma unroll
        for (int pt = 0; pt < NOCT; ++pt) {
            const double *bf = bfrag + (((size_t)(oct0 + pt) * m.n_ksplit + ks) * nj) * 32 + lane;
#pragma unroll
            for (int j = 0; j < SSD_NJ; ++j) b[j][pt] = j < nj ? bf[j * 32] : 0.0;
        }
    }
    double mg[NOCT][2];
    unsigned long long isum[NOCT][2];
    // one chain per octet (both row tiles of the pair add into it), and TWO sets of them used by
    // alternate observation tiles: a tile's chains are rounded to the fixed-point grid one k-step
    // into the NEXT tile, when they have drained on their own, so the tile boundary costs neither a
    // pipe drain nor a block of conversions
    double accA[NOCT][2], accB[NOCT][2];
#pragma unroll
    for (int pt = 0; pt < NOCT; ++pt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            mg[pt][e] = magic[(size_t)(oct0 + pt) * SSD_OCT + 2 * (lane & 3) + e];
            isum[pt][e] = 0ull;
            accA[pt][e] = 0.0; accB[pt][e] = 0.0;
        }
    auto convert = [&](double (&acc)[NOCT][2]) {
#pragma unroll
        for (int pt = 0; pt < NOCT; ++pt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                isum[pt][e] += (unsigned long long)xd_bits(__dadd_rn(acc[pt][e], mg[pt][e]));
                acc[pt][e] = 0.0;
            }
    };
    // one observation tile: its DMMAs go to `acc`; `prev` (the other set) is converted after the
    // first k-step when it holds the previous tile
    auto tile = [&](int t, double (&acc)[NOCT][2], double (&prev)[NOCT][2], bool have_prev) {
        const int it = t - T0, st = it & (XD_STAGES - 1);
        mbar_wait(&full[st], (uint32_t)(it / XD_STAGES) & 1u);
        if (tl && tid == 0 && t == T0) tl_max(tl, TL_XFIRST);
        const double2 *xa = reinterpret_cast<const double2 *>(ring + (size_t)st * stage_doubles) + lane;
        double2 a = xa[0];
#pragma unroll
        for (int j = 0; j < SSD_NJ; ++j) {
            if (j >= nj) break;
            double2 an = a;
            if (j + 1 < nj) an = xa[(j + 1) * 32];
#pragma unroll
            for (int pt = 0; pt < NOCT; ++pt) dmma884(acc[pt][0], acc[pt][1], a.x, b[j][pt]);
#pragma unroll
            for (int pt = 0; pt < NOCT; ++pt) dmma884(acc[pt][0], acc[pt][1], a.y, b[j][pt]);
            a = an;
            if (j == 0 && have_prev) convert(prev);
        }
        __syncwarp();                                        // every lane's reads of the stage have landed
        if (lane == 0 && t + XD_STAGES < T1) {
            mbar_expect_tx(&full[st], stage_bytes);
            bulk_g2s(ring + (size_t)st * stage_doubles, src + (size_t)(it + XD_STAGES) * stage_doubles, stage_bytes, &full[st]);
        }
    };
    int t = T0;
    tile(t++, accA, accB, false);
    for (; t + 1 < T1; t += 2) {
        tile(t, accB, accA, true);
        tile(t + 1, accA, accB, true);
    }
    if (t < T1) { tile(t, accB, accA, true); convert(accB); }
    else convert(accA);

    if (tid == 0) { tl_min(tl, TL_XLOOP0); tl_max(tl, TL_XLOOP1); }
    if (tlc && tid == 0 && blockIdx.x < TL_CTA_MAX) tlc[blockIdx.x * 4 + 2] = gtime();
    // remove the magic offsets (one conversion per observation tile), sum the 8 rows held by the
    // lanes of each column group, and add the CTA's share to the particles' accumulators
    const unsigned long long n_conv = (unsigned long long)(T1 - T0);
#pragma unroll
    for (int pt = 0; pt < NOCT; ++pt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            unsigned long long v = isum[pt][e] - n_conv * (unsigned long long)xd_bits(mg[pt][e]);
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int idx = (oct0 + pt) * SSD_OCT + 2 * lane + e;
            if (lane < 4 && idx < lv.n) {
                const int p = lv.order ? (int)((uint32_t)lv.order[idx] & LV_POS_MASK) : idx;
                atomicAdd(reinterpret_cast<unsigned long long *>(ll_acc) + p, v);
            }
        }
    if (tid == 0) tl_max(tl, TL_X1);
}

__global__ void __launch_bounds__(XD_THREADS, XD_CTAS_PER_SM) k_xdot(ModelDev m, const double *bfrag, const double *magic, Level lv,
                                                                     long long *ll_acc, XdGrid g, unsigned long long *tl, unsigned long long *tlc)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pdl_launch_dependents();
    if (threadIdx.x == 0) { tl_min(tl, TL_X0); tl_max(tl, TL_X0MAX); if (tlc && blockIdx.x < TL_CTA_MAX) tlc[blockIdx.x * 4] = gtime(); }
    int oct0, c_in, C, noct;
    const int n_in_hi = g.n_hi * g.c_hi;
    if ((int)blockIdx.x < n_in_hi) { const int t = blockIdx.x / g.c_hi; c_in = blockIdx.x - t * g.c_hi; C = g.c_hi; noct = g.oct_hi; oct0 = t * g.oct_hi; }
    else { const int r = blockIdx.x - n_in_hi, t = r / g.c_lo; c_in = r - t * g.c_lo; C = g.c_lo; noct = g.oct_lo; oct0 = g.n_hi * g.oct_hi + t * g.oct_lo; }
    const int n_tiles = (int)(m.ssd_ld / SSD_TN);
    const int T0 = (int)((int64_t)c_in * n_tiles / C), T1 = (int)((int64_t)(c_in + 1) * n_tiles / C);
    if (T1 <= T0) { pdl_wait(); return; }
#define XD_CALL(NO, NJC) xdot_body<NO, NJC>(m, bfrag, magic, lv, ll_acc, oct0, T0, T1, smem_raw, tl, tlc)
    if (m.ssd_nj == SSD_NJ) {
        switch (noct) { case 4: XD_CALL(4, SSD_NJ); break; case 3: XD_CALL(3, SSD_NJ); break; case 2: XD_CALL(2, SSD_NJ); break; default: XD_CALL(1, SSD_NJ); break; }
    } else {
        switch (noct) { case 4: XD_CALL(4, 0); break; case 3: XD_CALL(3, 0); break; case 2: XD_CALL(2, 0); break; default: XD_CALL(1, 0); break; }
    }
#undef XD_CALL
}

// centred means of arbitrary parameter vectors in the k_xdot layout (demcmc_eval, initial weights)
__global__ void __launch_bounds__(PA_THREADS) k_stage_means(ModelDev m, const double *theta, int64_t n, double *bfrag, double *magic,
                                                            long long *acc, double *q)
{
    const int64_t wi = ((int64_t)blockIdx.x * PA_THREADS + threadIdx.x) >> 5;
    if (wi >= n) return;
    stage_bfrag(m, theta + (size_t)wi * m.d, wi, bfrag, magic, acc + wi, q + wi, nullptr);
}

static int n_sms()
{
    static int sms[64] = { 0 };
    if (!sms[g_dev]) cudaDeviceGetAttribute(&sms[g_dev], cudaDevAttrMultiProcessorCount, g_dev);
    return sms[g_dev] > 0 ? sms[g_dev] : 148;
}

// A level's octets are dealt to particle tiles as evenly as possible (tiles of 3 and 4 octets rather
// than 4,4,..,1: a one-octet CTA has two DMMA chains and starves next to eight-chain warps), and
// the resident CTA slots of one wave are dealt to the tiles in proportion to their octets; a level
// too large for that is cut into many short CTAs instead.
static XdGrid xdot_grid(const ModelDev &m, int n, int slots)
{
    XdGrid g;
    const int octets = (n + SSD_OCT - 1) / SSD_OCT;
    const int nt = (octets + 3) / 4;
    g.oct_lo = octets / nt; g.oct_hi = g.oct_lo + 1;
    g.n_hi = octets - g.oct_lo * nt; g.n_lo = nt - g.n_hi;
    const int n_tiles = (int)(m.ssd_ld / SSD_TN);
    const int c_max = std::max(1, n_tiles / XD_MIN_TILES);
    const int per_split = std::max(1, slots / std::max(1, m.n_ksplit));
    if (2 * nt <= per_split) {                                     // one wave
        g.c_lo = std::min(c_max, std::max(1, per_split * g.oct_lo / octets));
        g.c_hi = g.n_hi ? std::min(c_max, std::max(1, per_split * g.oct_hi / octets)) : 1;
        // spend what the rounding left over on whichever class is slower
        for (;;) {
            const int left = per_split - (g.n_hi * g.c_hi + g.n_lo * g.c_lo);
            const bool hi_slower = g.n_hi && (int64_t)g.oct_hi * g.c_lo > (int64_t)g.oct_lo * g.c_hi;
            if (hi_slower && left >= g.n_hi && g.c_hi < c_max) ++g.c_hi;
            else if (left >= g.n_lo && g.c_lo < c_max) ++g.c_lo;
            else if (g.n_hi && left >= g.n_hi && g.c_hi < c_max) ++g.c_hi;
            else break;
        }
    } else {
        // several waves: many short CTAs, the hardware scheduler balances them as slots free up
        g.c_hi = g.c_lo = std::min(c_max, std::max(1, n_tiles / XD_WAVE_TILES));
    }
    return g;
}

static int launch_xdot(const ModelDev &m, const XdStage &xs, const Level &lv, long long *ll_acc)
{
    static bool attr_set[64] = { false };
    const size_t smem = xdot_smem_bytes(m.ssd_nj);
    if (!attr_set[g_dev]) {
        CU(cudaFuncSetAttribute(k_xdot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xdot_smem_bytes(SSD_NJ)));
        attr_set[g_dev] = true;
    }
    const XdGrid g = xdot_grid(m, lv.n, XD_CTAS_PER_SM * n_sms());
    dim3 grid((unsigned)(g.n_hi * g.c_hi + g.n_lo * g.c_lo), (unsigned)m.n_ksplit);
    CU(launch_chained(k_xdot, grid, dim3(XD_THREADS), smem, m, (const double *)xs.bfrag, (const double *)xs.magic, lv, ll_acc, g, lv.ctxs ? tl_slot() : (unsigned long long *)nullptr, lv.ctxs ? tl_cta() : (unsigned long long *)nullptr));
    LAUNCHED("k_xdot");
    return 0;
}

int launch_loglik(const ConfigDev &cfg, const ModelDev &m, const double *theta, const Level &lv, double *ll_part, long long *ll_acc)
{
    (void)cfg;
    if (lv.n <= 0 || m.kind == M_BINOMIAL || m.kind == M_RASTRIGIN) return 0;
    if (m.kind == M_MVNORMAL || m.kind == M_HIER) {
        // the proposal kernel of this level has staged the centred means and cleared the accumulators
        const XdStage &xs = g_xs[g_dev][g_lane];
        if (!xs.bfrag || !xs.magic) { g_be_err = "mean staging buffer missing"; return -1; }
        return launch_xdot(m, xs, lv, ll_acc);
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
// subject-major) -> center[k], then the centred data as DMMA A fragments (de_types.h:
// ssd_pack_index), zero padded; also sum x'^2 and max_i |x'_i|.  Fixed reduction order.
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

__global__ void __launch_bounds__(256) k_col_center(const double *x, double *center, int64_t n, int k, int obs_major)
{
    __shared__ double red[9];
    const int kk = blockIdx.x;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) s += obs_major ? x[i * k + kk] : x[(int64_t)kk * n + i];
    const double c = n > 0 ? block_sum_256(s, red) / (double)n : 0.0;
    if (threadIdx.x == 0) center[kk] = c;
}

// one thread per observation: writes its centred row into the packed layout; per block the sum and
// the maximum of the squared row norms
__global__ void __launch_bounds__(256) k_pack_rows(const double *x, const double *center, double *xp, double *blk_sq, double *blk_max,
                                                   int64_t n, int k, int ksplit_len, int nj, int64_t n_tiles, int obs_major)
{
    __shared__ double red[9];
    __shared__ double redm[8];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double q = 0.0;
    if (i < n)
        for (int kk = 0; kk < k; ++kk) {
            const double v = (obs_major ? x[i * k + kk] : x[(int64_t)kk * n + i]) - center[kk];
            q += v * v;
            xp[ssd_pack_index(i, kk, ksplit_len, nj, n_tiles)] = v;
        }
    double mx = q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) redm[threadIdx.x >> 5] = mx;
    const double s = block_sum_256(q, red);
    if (threadIdx.x == 0) {
        double mm = redm[0];
        for (int w = 1; w < 8; ++w) mm = fmax(mm, redm[w]);
        blk_sq[blockIdx.x] = s;
        blk_max[blockIdx.x] = mm;
    }
}

size_t pack_ssd_doubles(const ModelDev &m) { return (size_t)m.n_ksplit * (size_t)(m.ssd_ld / SSD_TN) * m.ssd_nj * 256; }

int launch_pack_ssd(const double *x_in, int in_on_device, ModelDev *m)
{
    const size_t bytes = sizeof(double) * (size_t)m->ssd_n * m->ssd_k;
    const int n_blk = (int)std::max<int64_t>(1, (m->ssd_n + 255) / 256);
    const double *src = x_in;
    double *tmp = nullptr, *blk = (double *)dmalloc(sizeof(double) * 2 * n_blk);
    if (!blk) return -1;
    if (!in_on_device) {
        tmp = (double *)dmalloc(bytes);
        if (!tmp) { dfree(blk); return -1; }
        if (h2d(tmp, x_in, bytes)) { dfree(tmp); dfree(blk); return -1; }
        src = tmp;
    }
    const int obs_major = m->kind == M_MVNORMAL ? 1 : 0;
    cudaError_t e = cudaMemsetAsync(const_cast<double *>(m->xT), 0, sizeof(double) * pack_ssd_doubles(*m), stream());
    if (e == cudaSuccess) e = cudaMemsetAsync(blk, 0, sizeof(double) * 2 * n_blk, stream());
    if (e == cudaSuccess) {
        k_col_center<<<m->ssd_k, 256, 0, stream()>>>(src, const_cast<double *>(m->center), m->ssd_n, m->ssd_k, obs_major);
        k_pack_rows<<<n_blk, 256, 0, stream()>>>(src, m->center, const_cast<double *>(m->xT), blk, blk + n_blk, m->ssd_n, m->ssd_k,
                                                 m->ksplit_len, m->ssd_nj, m->ssd_ld / SSD_TN, obs_major);
        g_launches += 2;
        e = cudaGetLastError();
    }
    std::vector<double> h(2 * (size_t)n_blk);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), blk, sizeof(double) * 2 * n_blk, cudaMemcpyDeviceToHost, stream());
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream());
    double xx = 0.0, mx = 0.0;
    for (int b = 0; b < n_blk; ++b) { xx += h[b]; mx = std::max(mx, h[n_blk + b]); }
    m->ssd_xx = xx;
    m->ssd_rowmax = 2.0 * sqrt(mx);                          // a chain of k_xdot sums the terms of TWO observation rows
    dfree(tmp); dfree(blk);
    if (e != cudaSuccess) return cu_fail(e, "k_pack_rows");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// evaluation of arbitrary parameter vectors (init_particle weights, demcmc_eval)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PA_THREADS) k_eval_finish(ConfigDev cfg, ModelDev m, const double *theta, int64_t n,
                                                             const double *part, const long long *acc, const double *q,
                                                             double *ll, double *prior, double *w)
{
    const int64_t wi = ((int64_t)blockIdx.x * PA_THREADS + threadIdx.x) >> 5;
    if (wi >= n) return;
    const WarpLanes co;
    const double *th = theta + (size_t)wi * cfg.d;
    bool inb; double pr;
    bounds_and_prior(co, cfg, m, th, inb, pr);
    const int n_split = m.n_osplit * m.n_ksplit;
    double s = 0.0;
    if (m.kind == M_MVNORMAL || m.kind == M_HIER) s = (double)acc[wi] * q[wi];
    else {
        if (m.kind != M_BINOMIAL && m.kind != M_RASTRIGIN) for (int c = co.lane(); c < n_split; c += 32) s += part[(size_t)wi * n_split + c];
        s = co.sum(s);
    }
    const double l = finalize_ll(m, th, s, mean_sq(co, m, th));
    if (co.lane() == 0) {
        if (ll) ll[wi] = l;
        if (prior) prior[wi] = inb ? pr : -inf();
        if (w) w[wi] = cfg.fitness == FITNESS_FUN ? (inb ? l : (cfg.update == UPDATE_MAXIMIZE ? -inf() : inf())) : (inb ? add(pr, l) : -inf());
    }
}

int launch_eval(const ConfigDev &cfg, const ModelDev &m, const double *theta, int64_t n, double *ll, double *prior,
                double *w, double *scratch_part)
{
    if (n <= 0) return 0;
    Level lv; lv.order = nullptr; lv.n = (int32_t)n; lv.ctxs = nullptr;
    const int blocks = (int)((n * 32 + PA_THREADS - 1) / PA_THREADS);
    long long *acc = nullptr;
    double *q = nullptr;
    if (m.kind == M_MVNORMAL || m.kind == M_HIER) {
        XdStage *xs = xd_stage(m, (int)n);
        if (!xs) return -1;
        acc = (long long *)dmalloc(sizeof(long long) * n);
        q = (double *)dmalloc(sizeof(double) * n);
        if (!acc || !q) { dfree(acc); dfree(q); return -1; }
        k_stage_means<<<blocks, PA_THREADS, 0, stream()>>>(m, theta, n, xs->bfrag, xs->magic, acc, q);
        LAUNCHED("k_stage_means");
        if (launch_xdot(m, *xs, lv, acc)) { dfree(acc); dfree(q); return -1; }
    } else if (launch_loglik(cfg, m, theta, lv, scratch_part, nullptr)) return -1;
    k_eval_finish<<<blocks, PA_THREADS, 0, stream()>>>(cfg, m, theta, n, scratch_part, acc, q, ll, prior, w);
    LAUNCHED("k_eval_finish");
    if (acc) { cudaStreamSynchronize(stream()); dfree(acc); dfree(q); }
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
                                                     double *theta, double *w, int32_t *id, uint8_t *acc, int32_t *pos)
{
    const int i = blockIdx.x, gl = a.groups[i] - cfg.group_begin;
    if (gl < 0 || gl >= cfg.G_local) return;
    const size_t p = (size_t)gl * cfg.Np + picks[i];
    const double *row = stage + (size_t)((i + a.n - 1) % a.n) * (cfg.d + 3);
    for (int k = threadIdx.x; k < cfg.d; k += blockDim.x) theta[p * cfg.d + k] = row[k];
    if (threadIdx.x == 0) {
        w[p] = row[cfg.d]; id[p] = (int32_t)row[cfg.d + 1]; acc[p] = (uint8_t)row[cfg.d + 2];
        if (pos) pos[(int32_t)row[cfg.d + 1] - cfg.group_begin * cfg.Np] = (int32_t)p;
    }
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
                       double *w, int32_t *id, uint8_t *acc, int32_t *pos)
{
    k_mig_scatter<<<a.n, 128, 0, stream()>>>(cfg, a, picks, stage, theta, w, id, acc, pos);
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

// bundle_samples (main.jl:222-250) on the device: chains[c][k][row] for rows [row0, row0+n_rows) of
// the history; parameter columns k < d of chain c are the draws of particle id c, the columns
// "acceptance" (d) and "lp" (d+1) belong to the particle sitting at final position c
__global__ void __launch_bounds__(256) k_chain_pos(const int32_t *final_id, int P, int id_base, int32_t *pos_of_id)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P) return;
    const int id = final_id[c] - id_base;
    if (id >= 0 && id < P) pos_of_id[id] = c;
}
__global__ void __launch_bounds__(256) k_chains(const double *rt, const double *rw, const uint8_t *ra, const int32_t *rid,
                                                const int32_t *pos_of_id, int64_t row0, int64_t n_rows, int P, int d, int id_base, double *out)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int slot = blockIdx.y;
    if (r >= n_rows) return;
    const int64_t row = row0 + r;
    const int id = rid[row * P + slot] - id_base;
    if (id < 0 || id >= P) return;
    const double *src = rt + (row * P + slot) * d;
    double *dst = out + (int64_t)id * (d + 2) * n_rows + r;
    for (int k = 0; k < d; ++k) dst[(int64_t)k * n_rows] = src[k];
    double *dq = out + (int64_t)pos_of_id[id] * (d + 2) * n_rows + r;
    dq[(int64_t)d * n_rows] = (double)ra[row * P + slot];
    dq[(int64_t)(d + 1) * n_rows] = rw[row * P + slot];
}

int launch_chains(const double *rows_theta, const double *rows_w, const uint8_t *rows_acc, const int32_t *rows_id,
                  const int32_t *final_id, int32_t *pos_scratch, int64_t row0, int64_t n_rows, int32_t P, int32_t d, int32_t id_base, double *out)
{
    k_chain_pos<<<(P + 255) / 256, 256, 0, stream()>>>(final_id, P, id_base, pos_scratch);
    LAUNCHED("k_chain_pos");
    if (n_rows <= 0) return 0;
    dim3 grid((unsigned)((n_rows + 255) / 256), (unsigned)P);
    k_chains<<<grid, 256, 0, stream()>>>(rows_theta, rows_w, rows_acc, rows_id, pos_scratch, row0, n_rows, P, d, id_base, out);
    LAUNCHED("k_chains");
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

// the fp64 tensor path: 8 independent DMMA m8n8k4 chains per warp
__global__ void __launch_bounds__(256) k_dmma_peak(double *out, int iters, double a, double b)
{
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

// DFMA loop and DMMA loop, best of 5 timed repetitions each, TFLOP/s
int fp64_peaks(double *dfma_tflops, double *dmma_tflops)
{
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g_dev));
    double *out = (double *)dmalloc(8);
    if (!out) return -1;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int blocks = sms * 8;
    double best[2] = { 0.0, 0.0 };
    for (int which = 0; which < 2; ++which)
        for (int rep = 0; rep < 6; ++rep) {
            const int iters = which == 0 ? 1 << 14 : 1 << 12;
            CU(cudaEventRecord(e0, stream()));
            if (which == 0) k_dfma_peak<<<blocks, 256, 0, stream()>>>(out, iters, 0.999999, 1e-9);
            else k_dmma_peak<<<blocks, 256, 0, stream()>>>(out, iters, 0.999999, 1e-9);
            ++g_launches;
            CU(cudaEventRecord(e1, stream()));
            CU(cudaEventSynchronize(e1));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            const double fl = which == 0 ? 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks
                                         : 8.0 * 512.0 * (double)iters * 8.0 * (double)blocks;
            if (rep > 0) best[which] = fmax(best[which], fl / (ms * 1e-3) / 1e12);
        }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    dfree(out);
    if (dfma_tflops) *dfma_tflops = best[0];
    if (dmma_tflops) *dmma_tflops = best[1];
    return 0;
}

int fp64_peak(double *tflops)
{
    double a = 0.0, b = 0.0;
    if (fp64_peaks(&a, &b)) return -1;
    *tflops = fmax(a, b);
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
// One communicator per (device, rank, n_ranks) is kept for the life of the process: every handle
// of a job that is sharded the same way reuses it (creating one costs about a second), and the
// unique id of later calls is ignored -- all ranks take the same branch, so this stays collective.
struct CommSlot { int rank = -1, n = 0; nccl_comm c = nullptr; };
static CommSlot g_comm[64];

int comm_init(const uint8_t id[128], int rank, int n_ranks, void **comm)
{
    if (nccl_load()) return -1;
    CommSlot &slot = g_comm[g_dev];
    if (slot.c && slot.rank == rank && slot.n == n_ranks) { *comm = slot.c; return 0; }
    nccl_uid u;
    memcpy(u.internal, id, 128);
    nccl_comm c = nullptr;
    NC(g_nccl.CommInitRank(&c, n_ranks, u, rank));
    *comm = c;
    // NCCL opens point-to-point connections lazily: exchange one row with every peer now, so the
    // first migration that crosses ranks does not pay for the connection setup
    if (n_ranks > 1) {
        double *buf = (double *)dmalloc(sizeof(double) * 2 * n_ranks);
        if (!buf) return -1;
        NC(g_nccl.GroupStart());
        for (int r = 0; r < n_ranks; ++r) {
            if (r == rank) continue;
            NC(g_nccl.Send(buf + r, 1, 8 /* ncclFloat64 */, r, c, stream()));
            NC(g_nccl.Recv(buf + n_ranks + r, 1, 8, r, c, stream()));
        }
        NC(g_nccl.GroupEnd());
        CU(cudaStreamSynchronize(stream()));
        dfree(buf);
    }
    if (slot.c && g_nccl.lib) g_nccl.CommDestroy(slot.c);
    slot.rank = rank; slot.n = n_ranks; slot.c = c;
    return 0;
}
int comm_destroy(void *comm)
{
    (void)comm;                                              // owned by the per-device slot above
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
