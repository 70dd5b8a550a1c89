// planner.h -- host-side, state-independent scheduling of a sweep.
//
// The reference updates the particles of a group one after another, in place
// (crossover.jl:12-17, utilities.jl:201-210), so a donor with a smaller slot than the target has
// already taken this sweep's value.  Which slots are donors does not depend on the state, so the
// host can compute, ahead of the device, the dependency LEVEL of every particle:
//   level(j) = 1 + max{ level(k) : k a donor of j, k < j }   (0 without such a donor)
// Particles of one level are mutually independent and are updated by one launch; the levels of a
// sweep replay the sequential semantics exactly.  Mutation sweeps (mutation.jl:13-25) have a
// single level.
#pragma once
#include <stdint.h>
#include <vector>

namespace de {

struct PlanInput {
    uint64_t seed;
    int32_t Np, G_local, group_begin, G_total;
    int32_t proposal;          // 0 random_gamma
    double beta, theta_snooker;
    // replay: slices of this sweep for the LOCAL shard, else nullptr
    const uint8_t *t_kind;     // [P_local]
    const int32_t *t_idx;      // [P_local][3]
    bool base_dependency;      // replay + exact_base + random_gamma + burn-in: idx[.][0] is a donor too
};

struct SweepPlan {
    std::vector<uint8_t> mutate;      // [G_local]
    std::vector<int32_t> order;       // [P_local] local positions sorted by level (stable)
    std::vector<int32_t> level_off;   // [n_levels + 1]
    int32_t n_levels = 0;
};

void plan_sweep(const PlanInput &in, uint32_t sweep, SweepPlan &out);

// migration! schedule (migration.jl:56-60): u <= alpha, N = rand(2:G), ordered subset of N groups,
// and the uniform handed to select_particle for each position
struct MigSchedule { bool migrate; int32_t n; std::vector<int32_t> groups; std::vector<double> u_pick; double u_mig; };
void plan_migration(uint64_t seed, uint32_t iter0, int32_t G_total, double alpha, MigSchedule &out);

} // namespace de
