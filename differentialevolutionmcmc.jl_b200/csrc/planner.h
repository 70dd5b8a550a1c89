// planner.h -- host-side, state-independent scheduling of the population step.
//
// The reference updates the particles of a group one after another, in place
// (crossover.jl:12-17, utilities.jl:201-210): a donor with a smaller slot than the target already
// holds this sweep's value, a donor with a larger slot still holds the previous sweep's.  Which
// slots are donors does not depend on the state, so the host computes, ahead of the device, a
// dependency LEVEL for every (sweep, particle) update of a CHUNK of consecutive sweeps:
//   level(t, j) = 1 + max( level(t-1, j),                          its own previous update
//                          level(t,   k) for donors k < j,         this sweep's value
//                          level(t-1, k) for donors k > j )        the previous sweep's value
// Updates of one level are mutually independent and run in one launch; running the levels in
// order replays the sequential semantics exactly.  State rows are write-once (sweep t reads row
// t-1 and writes row t), so there are no anti-dependencies, and the tail levels of sweep t share
// launches with the head levels of sweep t+1 (about 4.7 levels per sweep instead of 8 at Np=256).
// A chunk ends wherever the state must be complete: before a migration, and before every sweep
// whose select_base needs the sweep-start weights.
#pragma once
#include <stdint.h>
#include <vector>

namespace de {

constexpr int ENTRY_SLOT_SHIFT = 24;                 // level entry = (sweep slot << 24) | local position
constexpr uint32_t ENTRY_POS_MASK = (1u << ENTRY_SLOT_SHIFT) - 1;
constexpr int MAX_CHUNK = 16;                        // sweeps planned and launched together

struct PlanInput {
    uint64_t seed;
    int32_t Np, G_local, group_begin, G_total;   // the groups this plan covers: G_local of them from global group group_begin
    int32_t pos_offset = 0;    // local position of the first particle covered (entries and tape slices are offset by it)
    int32_t P_stride = 0;      // particles per sweep in the tape slices (0: Np * G_local)
    int32_t sweep_stride = 1;  // Philox sweep coordinate of consecutive sweeps of the chunk: sweep0 + s * sweep_stride
                               // (unblocked iterations of a model with B parameter blocks: B)
    int32_t proposal;          // 0 random_gamma
    double beta, theta_snooker;
    int32_t shape_octets = 0;  // > 0: hand each level's remainder modulo this many updates to the next level where the
                               // dependencies allow (the DMMA likelihood kernel pads levels to whole octets of particles,
                               // and its particle tiles are most efficient with four octets: 8 or 32)
    int32_t level_cap = 0;     // > 0: at most this many updates per level (list scheduling in (sweep, slot) order: an update
                               // goes to the first level after its dependencies that still has room).  The persistent chunk
                               // kernel wants levels of one size: its lanes hide each other's accept -> propose chain only
                               // when their levels take about as long as that chain
    bool resample;             // donors come from stored rows (crossover.jl:113-124): no donor dependencies inside a sweep
    // replay: tape slices [sweep][P_local] of the chunk's FIRST sweep onwards, else nullptr
    const uint8_t *t_kind;     // [n_sweeps][P_local]
    const int32_t *t_idx;      // [n_sweeps][P_local][3]
};

struct ChunkPlan {
    int32_t n_sweeps = 0, n_levels = 0;
    std::vector<uint8_t> mutate;      // [n_sweeps][G_local]
    std::vector<int32_t> order;       // [n_sweeps * P_local] entries sorted by level (stable)
    std::vector<int32_t> level_off;   // [n_levels + 1]
};

// base_dependency[s]: idx[.][0] of sweep s is a donor too (replay, random_gamma, burn-in)
void plan_chunk(const PlanInput &in, uint32_t sweep0, int32_t n_sweeps, const bool *base_dependency, ChunkPlan &out);

// migration! schedule (migration.jl:56-60): u <= alpha, N = rand(2:G), ordered subset of N groups,
// and the uniform handed to select_particle for each position
struct MigSchedule { bool migrate; int32_t n; std::vector<int32_t> groups; std::vector<double> u_pick; double u_mig; };
void plan_migration(uint64_t seed, uint32_t iter0, int32_t G_total, double alpha, MigSchedule &out);

} // namespace de
