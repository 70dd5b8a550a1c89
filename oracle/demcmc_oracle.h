/*
 * demcmc_oracle.h -- CPU restatement of the DE-MCMC population step of
 * itsdfish/DifferentialEvolutionMCMC.jl (v0.7.10).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / the reported CPU baseline.
 *
 * PARITY STATUS: the reference is pure Julia and no Julia toolchain exists in this
 * image, so the reference itself cannot be executed here and it ships no golden
 * chain or golden log-density.  What IS pinned (tests/test_oracle_known_answers.py):
 * the reference's portable known-answer vectors for this path -- projection
 * (test/utility_tests.jl:71-93), reset! (:42-69), particle algebra (:161-199) and the
 * cyclic-shift property of migration (:95-154) -- and every density against
 * scipy/mpmath.  Chain-level (seed-exact) parity with a Julia run is UNPINNED.
 *
 * Every function cites the reference file:line it restates.
 */
#ifndef DEMCMC_ORACLE_H
#define DEMCMC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* registered likelihood kernels */
enum { ORC_GAUSSIAN = 0, ORC_MVNORMAL = 1, ORC_BINOMIAL = 2, ORC_LNR = 3, ORC_LBA = 4, ORC_HIER_NORMAL = 5,
       ORC_RASTRIGIN = 6 /* the objective of test/optimization_tests.jl:15-23; no data */,
       ORC_MVN_FULL = 7  /* MvNormal(mu, sigma^2 * Sigma) with a known covariance Sigma (not a model of the reference's
                            examples: Distributions.logpdf(MvNormal(mu, Sigma), data) by its published definition) */ };
/* de.update_particle! (utilities.jl:201-226) and de.evaluate_fitness! (utilities.jl:92-120) */
enum { ORC_UPDATE_MH = 0, ORC_UPDATE_MAXIMIZE = 1, ORC_UPDATE_MINIMIZE = 2 };
enum { ORC_FITNESS_POSTERIOR = 0, ORC_FITNESS_FUN = 1 };
/* registered prior specs, one per flattened parameter element */
enum { ORC_PRIOR_FLAT = 0, ORC_PRIOR_NORMAL = 1, ORC_PRIOR_HALFCAUCHY = 2, ORC_PRIOR_UNIFORM = 3,
       ORC_PRIOR_BETA = 4, ORC_PRIOR_NORMAL_REF = 5 };
/* generate_proposal (src/crossover.jl:154-226) */
enum { ORC_RANDOM_GAMMA = 0, ORC_FIXED_GAMMA = 1, ORC_VARIABLE_GAMMA = 2 };
/* per-particle update kind recorded on the tape */
enum { ORC_KIND_DE = 0, ORC_KIND_SNOOKER = 1, ORC_KIND_MUTATION = 2 };

typedef struct {
    int32_t kind;      /* ORC_PRIOR_* */
    int32_t ref;       /* ORC_PRIOR_NORMAL_REF: flattened index of the sd parameter */
    double a, b;       /* NORMAL(mean a, sd b); HALFCAUCHY(loc a, scale b) truncated to [0,inf);
                          UNIFORM(a,b); BETA(a,b); NORMAL_REF(mean a, sd = theta[ref]) */
} orc_prior;

typedef struct {
    int32_t kind;          /* ORC_GAUSSIAN ... */
    int32_t d;             /* flattened parameter count */
    int64_t n_obs;         /* observations / trials / (binomial: unused) */
    int32_t n_dim;         /* MVNORMAL: data dimension; LNR/LBA: accumulators; HIER: subjects */
    int32_t n_per;         /* HIER: observations per subject */
    const double *x;       /* GAUSSIAN x[n_obs]; MVNORMAL x[n_obs][n_dim]; LNR/LBA rt[n_obs];
                              HIER y[n_dim][n_per]; BINOMIAL {N, k} as doubles */
    const int32_t *choice; /* LNR/LBA: 1-based winner per trial */
    const double *sigma;   /* LNR: sd per accumulator [n_dim] (NULL => 1) */
    double lba_floor;      /* LBA: density floor (SequentialSamplingModels uses 1e-10); 0 disables */
    const orc_prior *prior;/* [d] */
    const double *cov;     /* MVN_FULL: covariance [n_dim][n_dim], symmetric positive definite */
} orc_model;

typedef struct {
    int32_t n_groups, Np, d;
    int32_t burnin, n_initial;
    double alpha, beta, eps, sigma, kappa, theta_snooker;
    int32_t proposal;        /* ORC_RANDOM_GAMMA ... */
    int32_t n_blocks;        /* 0 => blocking_on(de) == false */
    const uint8_t *blocks;   /* [n_blocks][d], 1 = updated in this block */
    const double *lo, *hi;   /* [d] bounds expanded per element (utilities.jl:70-78) */
    int32_t base_snapshot;   /* 0 = reference semantics.  1 = select_base reads the weights AND
                                the base particle's theta as they were at sweep start (the
                                documented native-mode deviation of the B200 path, burn-in only) */
    int32_t n_threads;       /* >1: one OpenMP task per group (main.jl:135-148) */
    uint64_t seed;           /* Philox key for generated draws */
    int32_t resample;        /* 1: de.sample = resample (crossover.jl:113-124): donors are n distinct
                                (row, id) cells of de.samples[1:de.iter-1, :, :] (DE-MCz); needs
                                n_initial > 0 and the caller's prior rows in `samples` */
    int32_t update;          /* ORC_UPDATE_*: mh_update!, maximize!, minimize! (the optimize path) */
    int32_t fitness;         /* ORC_FITNESS_*: compute_posterior! or evaluate_fun! (loglike only, no prior) */
    int32_t reserved;
    /* blocking_on(de) per iteration (main.jl:137,162: a function of the sampler, evaluated every
     * iteration): block_on[it] != 0 => block_update! in iteration it, else update! (all parameters at
     * once).  NULL, or it >= n_block_on => on whenever n_blocks > 0.  A sweep of an unblocked
     * iteration is sweep slot it*B + 0 of the tape / trace arrays; slots 1..B-1 stay unused. */
    const uint8_t *block_on;
    int64_t n_block_on;
} orc_config;

/* Structured replay tape.  All indices are 0-based slot indices inside
 * the group; particle position p = g*Np + j; sweep s = iter0*B + block (B = max(1,n_blocks)).
 * Any pointer may be NULL when recording is not wanted.  In consume mode (orc_run with
 * tape_in) every non-NULL array REPLACES the corresponding generated draw. */
typedef struct {
    double  *mig_u;       /* [n_iter]        u for rand() <= alpha                       */
    int32_t *mig_n;       /* [n_iter]        number of migrating groups (0 = none)       */
    int32_t *mig_groups;  /* [n_iter][G]     ordered subset of groups, -1 padded         */
    double  *mig_pick_u;  /* [n_iter][G]     uniform used by select_particle             */
    int32_t *mig_slots;   /* [n_iter][G]     recorded picked slot (-1 padded)            */
    double  *mut_u;       /* [S][G]          u for rand() <= beta                        */
    uint8_t *kind;        /* [S][P]          ORC_KIND_*                                  */
    int32_t *idx;         /* [S][P][3]       DE: (base,m,n); snooker: (z,m,n)            */
    double  *u_snk;       /* [S][P]                                                      */
    double  *u_base;      /* [S][P]                                                      */
    double  *gamma1;      /* [S][P]          DE gamma_1 or snooker gamma                 */
    double  *gamma2;      /* [S][P]          DE gamma_2 (0 after burn-in)                */
    double  *u_acc;       /* [S][P]                                                      */
    double  *noise;       /* [S][P][d]       b_k (crossover) or N(0,sigma) (mutation)    */
    uint8_t *keep;        /* [S][P][d]       1 = recombination restores theta_t,k        */
    int32_t *idx_row;     /* [S][P][3]       resample: history row of each donor (idx then holds
                                             the donor's particle id); -1 where unused            */
} orc_tape;

/* Optional per-sweep trace for teacher-forced comparison.  NULL pointers are skipped. */
typedef struct {
    double  *prop_theta;  /* [S][P][d]  proposal after recombination!/reset!              */
    double  *prop_weight; /* [S][P]     proposal log posterior (-inf if out of bounds)    */
    double  *log_adj;     /* [S][P]     snooker adjustment (0 otherwise)                  */
    uint8_t *accepted;    /* [S][P]                                                       */
    double  *state_theta; /* [n_iter][P][d] theta by slot after each iteration            */
    double  *state_weight;/* [n_iter][P]                                                  */
    int32_t *state_id;    /* [n_iter][P]  0-based particle id by slot                     */
    double  *pre_theta;   /* [n_iter][P][d] theta by slot after migration, before update  */
    double  *pre_weight;  /* [n_iter][P]                                                  */
    int32_t *pre_id;      /* [n_iter][P]                                                  */
} orc_trace;

/* Runs n_iter iterations of step!/pstep! (main.jl:84-107).
 *  theta0[P][d]       initial state by slot (= id order, main.jl:263-271)
 *  samples            [n_rows][d][P] Fortran order exactly as utilities.jl:34 (row fastest),
 *                     n_rows = n_iter + n_initial; rows < n_initial are the caller's
 *                     initialize_samples prior draws (utilities.jl:29-41) and are left untouched;
 *                     with n_initial > 0 the initial state is samples[0, :, id]
 *                     (init_particle, utilities.jl:15) and theta0 is ignored
 *  accept [n_rows][P] (row fastest, column = particle id), lp likewise
 *  final_id[P]        particle id sitting at each slot after the run (0-based)
 *  final_theta[P][d], final_weight[P]
 * Returns 0, or a negative error code. */
int orc_run(const orc_config *cfg, const orc_model *model, const double *theta0, int64_t n_iter,
            const orc_tape *tape_in, orc_tape *tape_out, orc_trace *trace,
            double *samples, uint8_t *accept, double *lp,
            int32_t *final_id, double *final_theta, double *final_weight);

/* log posterior pieces, exposed for density tests */
double orc_loglike(const orc_model *m, const double *theta);
/* bench.py's CPU arm: plain (uncompensated) sums in the MVN likelihood, the speed a straightforward CPU code has */
void orc_set_plain_sums(int on);
double orc_prior_loglike(const orc_model *m, const double *theta);
/* compute_posterior! (utilities.jl:92-99) */
double orc_posterior(const orc_config *cfg, const orc_model *m, const double *theta);

/* Particle algebra known-answer hooks (utilities.jl:239-357, crossover.jl:268-273,336-352) */
void orc_project(const double *p1, const double *p2, int d, double *out);
double orc_adjust_loglike(const double *pt, const double *prop, const double *pz, int d);
void orc_reset(double *prop, const double *pt, const uint8_t *mask, int d);
void orc_de_proposal(const double *pt, const double *pm, const double *pn, const double *pb,
                     double g1, double g2, const double *b, int d, double *out);
void orc_snooker_proposal(const double *pt, const double *pz, const double *pm, const double *pn,
                          double g, const double *b, int d, double *out);
int orc_accept(double w_prop, double w_cur, double log_adj, double u);
/* select_base / select_particle (crossover.jl:282-289, migration.jl:89-95); *drew = 1 when the
 * uniform was consumed */
int orc_select_base(const double *w, int n, double u);
int orc_select_particle(const double *w, int n, double u, int *drew);
/* migration cyclic shift on a plain array of "particle tags" (migration.jl:109-116) */
void orc_shift(int32_t *tags, const int32_t *groups, const int32_t *slots, int n, int Np);

/* Philox4x32-10 (Salmon et al. 2011) and the uniform mapping shared with the B200 path */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_uniform2(uint64_t seed, uint32_t stream, uint32_t sweep, uint32_t unit, uint32_t k, double u[2]);

#ifdef __cplusplus
}
#endif
#endif
