// planner.cpp -- see planner.h
#include "planner.h"
#include <algorithm>
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
    int max_level = 0;
    for (int s = 0; s < n_sweeps; ++s) {
        const uint32_t sweep = sweep0 + (uint32_t)s;
        const uint8_t *tk = in.t_kind ? in.t_kind + (size_t)s * stride + in.pos_offset : nullptr;
        const int32_t *ti = in.t_idx ? in.t_idx + ((size_t)s * stride + in.pos_offset) * 3 : nullptr;
        for (int g = 0; g < G; ++g) {
            const int gg = in.group_begin + g;
            bool mutate;
            if (tk) mutate = tk[g * Np] == KIND_MUTATION;
            else mutate = uniform2(in.seed, ST_MUT, sweep, (uint32_t)gg, 0).a <= in.beta;   // main.jl:200
            out.mutate[(size_t)s * G + g] = mutate ? 1 : 0;
            int32_t *lc = cur.data() + g * Np;
            const int32_t *lp = prev.data() + g * Np;
            for (int j = 0; j < Np; ++j) {
                int l = lp[j] + 1;                                   // its own previous update
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
                    }
                }
                lc[j] = l;
                level[(size_t)s * P + g * Np + j] = l;
                max_level = std::max(max_level, l);
            }
        }
        prev = cur;
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
