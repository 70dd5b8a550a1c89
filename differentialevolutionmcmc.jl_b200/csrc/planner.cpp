// planner.cpp -- see planner.h
#include "planner.h"
#include <algorithm>
#include <cstdlib>
#include <thread>
#include "de_math.h"

namespace de {

void plan_chunk(const PlanInput &in, uint32_t sweep0, int32_t n_sweeps, const bool *base_dependency, ChunkPlan &out)
{
    const int Np = in.Np, G = in.G_local, P = Np * G;
    const int stride = in.P_stride > 0 ? in.P_stride : P;
    out.n_sweeps = n_sweeps;
    out.mutate.assign((size_t)n_sweeps * G, 0);
    out.order.resize((size_t)n_sweeps * P);
    std::vector<int32_t> level((size_t)n_sweeps * P, 0), prev(P, -1), cur(P, 0);
    // dependencies of every update as update indices (sweep * P + position), for the shaping pass below
    std::vector<int32_t> deps((size_t)n_sweeps * P * 4, -1);
    // Groups never read each other inside a chunk, so their level assignments are independent: large plans (the Philox
    // draws of plan_particle dominate: ~40 ns per update, 5 ms for 16 sweeps of 8192 particles) are split over a few host
    // threads by group range.  The result does not depend on the split.
    int n_thr = 1;
    if ((size_t)n_sweeps * P >= 16384 && G >= 2) {
        static const int hw = [] { const char *e = getenv("DEMCMC_PLAN_THREADS"); const int v = e ? atoi(e) : (int)std::thread::hardware_concurrency() / 2; return std::max(1, std::min(v, 8)); }();
        n_thr = std::min(hw, G);
    }
    std::vector<int> max_level_thr(n_thr, 0);
    auto plan_groups = [&](int tix, int g_lo, int g_hi) {
    int max_level = 0;
    for (int s = 0; s < n_sweeps; ++s) {
        const uint32_t sweep = sweep0 + (uint32_t)s * (uint32_t)in.sweep_stride;
        const uint8_t *tk = in.t_kind ? in.t_kind + (size_t)s * stride + in.pos_offset : nullptr;
        const int32_t *ti = in.t_idx ? in.t_idx + ((size_t)s * stride + in.pos_offset) * 3 : nullptr;
        for (int g = g_lo; g < g_hi; ++g) {
            const int gg = in.group_begin + g;
            bool mutate;
            if (tk) mutate = tk[g * Np] == KIND_MUTATION;
            else mutate = uniform2(in.seed, ST_MUT, sweep, (uint32_t)gg, 0).a <= in.beta;   // main.jl:200
            out.mutate[(size_t)s * G + g] = mutate ? 1 : 0;
            int32_t *lc = cur.data() + g * Np;
            const int32_t *lp = prev.data() + g * Np;
            for (int j = 0; j < Np; ++j) {
                int l = lp[j] + 1;                                   // its own previous update
                int32_t *du = deps.data() + ((size_t)s * P + g * Np + j) * 4;
                if (s > 0) du[0] = (int32_t)((size_t)(s - 1) * P + g * Np + j);
                if (!mutate) {
                    int dep[3], nd = 0;
                    if (in.resample) {
                        // only select_base still reads the current group (replay, burn-in)
                        if (tk && tk[g * Np + j] == KIND_DE && base_dependency && base_dependency[s]) dep[nd++] = ti[(size_t)(g * Np + j) * 3];
                    } else if (tk) {
                        const int32_t *ix = ti + (size_t)(g * Np + j) * 3;
                        if (tk[g * Np + j] == KIND_SNOOKER) { dep[nd++] = ix[0]; dep[nd++] = ix[1]; dep[nd++] = ix[2]; }
                        else { dep[nd++] = ix[1]; dep[nd++] = ix[2]; if (base_dependency && base_dependency[s]) dep[nd++] = ix[0]; }
                    } else {
                        const Plan p = plan_particle(in.seed, sweep, (uint32_t)(gg * Np + j), j, Np, false, in.theta_snooker);
                        if (p.kind == KIND_SNOOKER) dep[nd++] = p.i0;
                        dep[nd++] = p.i1; dep[nd++] = p.i2;
                    }
                    for (int q = 0; q < nd; ++q) {
                        const int k = dep[q];
                        if (k < 0 || k == j) continue;
                        l = std::max(l, (k < j ? lc[k] : lp[k]) + 1);
                        if (k < j) du[1 + q] = (int32_t)((size_t)s * P + g * Np + k);
                        else if (s > 0) du[1 + q] = (int32_t)((size_t)(s - 1) * P + g * Np + k);
                    }
                }
                lc[j] = l;
                level[(size_t)s * P + g * Np + j] = l;
                max_level = std::max(max_level, l);
            }
        }
        std::copy(cur.begin() + (size_t)g_lo * Np, cur.begin() + (size_t)g_hi * Np, prev.begin() + (size_t)g_lo * Np);
    }
    max_level_thr[tix] = max_level;
    };
    if (n_thr == 1) plan_groups(0, 0, G);
    else {
        std::vector<std::thread> thr;
        for (int t = 0; t < n_thr; ++t) thr.emplace_back(plan_groups, t, (int)((int64_t)t * G / n_thr), (int)((int64_t)(t + 1) * G / n_thr));
        for (auto &t : thr) t.join();
    }
    int max_level = *std::max_element(max_level_thr.begin(), max_level_thr.end());
    // Capacity: levels of at most level_cap updates.  (sweep, slot) order is a topological order of the dependencies
    // (own previous update, donors with a smaller slot of this sweep, donors with a larger slot of the previous one),
    // so every update can simply take the first level after its dependencies that still has room.
    if (in.level_cap > 0) {
        std::vector<int32_t> cnt;
        max_level = 0;
        for (size_t u = 0; u < level.size(); ++u) {
            int e = 0;
            for (int q = 0; q < 4; ++q) { const int32_t v = deps[u * 4 + q]; if (v >= 0) e = std::max(e, level[v] + 1); }
            while (e < (int)cnt.size() && cnt[e] >= in.level_cap) ++e;
            if (e >= (int)cnt.size()) cnt.resize(e + 1, 0);
            ++cnt[e];
            level[u] = e;
            max_level = std::max(max_level, e);
        }
    }
    // Shaping: the likelihood kernel pads every level to whole octets of particles (DMMA n = 8), on
    // average 3.5 idle columns per level.  An update whose dependents all sit two or more levels
    // later may run one level later at no cost, so each level hands its remainder (n mod 8) to the
    // next one where it can.  Any schedule that respects the dependencies gives the same chain.
    if (in.shape_octets && max_level > 0) {
        const size_t U = level.size();
        std::vector<int32_t> latest(U, max_level);
        for (size_t v = U; v-- > 0;)                                   // dependents come later in (sweep, slot) order
            for (int q = 0; q < 4; ++q) {
                const int32_t u = deps[v * 4 + q];
                if (u >= 0) latest[u] = std::min(latest[u], level[v] - 1);
            }
        std::vector<std::vector<int32_t>> members(max_level + 1);
        for (size_t u = 0; u < U; ++u) members[level[u]].push_back((int32_t)u);
        for (int l = 0; l < max_level; ++l) {
            int r = (int)(members[l].size() % (size_t)in.shape_octets);
            if (r == 0 || (int)members[l].size() < in.shape_octets) continue;
            // later entries first: they keep the level's (sweep, slot) order intact for the rest
            for (size_t i = members[l].size(); i-- > 0 && r > 0;) {
                const int32_t u = members[l][i];
                if (latest[u] <= l) continue;
                // a dependent that was itself moved gives even more room; the bound stays valid
                level[u] = l + 1;
                members[l + 1].push_back(u);
                members[l][i] = -1;
                --r;
            }
        }
    }
    // stable counting sort of all (sweep, position) entries by level
    out.n_levels = max_level + 1;
    out.level_off.assign(out.n_levels + 1, 0);
    for (size_t e = 0; e < level.size(); ++e) out.level_off[level[e] + 1]++;
    for (int l = 0; l < out.n_levels; ++l) out.level_off[l + 1] += out.level_off[l];
    std::vector<int32_t> cursor(out.level_off.begin(), out.level_off.end() - 1);
    for (int s = 0; s < n_sweeps; ++s)
        for (int p = 0; p < P; ++p)
            out.order[cursor[level[(size_t)s * P + p]]++] = (int32_t)(((uint32_t)s << ENTRY_SLOT_SHIFT) | (uint32_t)(in.pos_offset + p));
}

void plan_migration(uint64_t seed, uint32_t iter0, int32_t G, double alpha, MigSchedule &out)
{
    out.groups.clear(); out.u_pick.clear(); out.n = 0; out.migrate = false;
    const dbl2 u = uniform2(seed, ST_MIG, iter0, 0, 0);
    out.u_mig = u.a;
    if (G < 2 || !(u.a <= alpha)) return;                    // main.jl:85
    out.migrate = true;
    const int N = 2 + rand_index(u.b, G - 1);                // rand(2:n_groups), migration.jl:57
    std::vector<int32_t> arr(G);
    for (int i = 0; i < G; ++i) arr[i] = i;
    for (int i = 0; i < N; ++i) {                            // ordered subset without replacement
        const dbl2 v = uniform2(seed, ST_MIG, iter0, 0, (uint32_t)(1 + i));
        const int j = i + rand_index(v.a, G - i);
        std::swap(arr[i], arr[j]);
        out.groups.push_back(arr[i]);
        out.u_pick.push_back(v.b);
    }
    out.n = N;
}

} // namespace de
