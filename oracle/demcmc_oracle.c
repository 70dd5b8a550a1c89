/*
 * demcmc_oracle.c -- CPU restatement (plain C, fp64, no FMA contraction) of the population step
 * of itsdfish/DifferentialEvolutionMCMC.jl.  TEST INFRASTRUCTURE ONLY -- see demcmc_oracle.h for
 * the parity status ("chain-level parity unpinned": no Julia here, no golden chain upstream).
 *
 * The sweep inside a group is SEQUENTIAL and IN PLACE exactly as in the reference
 * (src/crossover.jl:12-17, src/utilities.jl:201-210); arithmetic association follows the Julia
 * expressions literally.  Build with -ffp-contract=off (oracle/Makefile).
 */
#include "demcmc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LOG2PI 1.8378770664093454835606594728112
#define LOGPI 1.1447298858494001741434273513531

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 and the draw map                                                              */
/* ------------------------------------------------------------------------------------------ */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* streams of the draw map (shared spec with the B200 path, DESIGN.md "RNG") */
enum { ST_MIG = 1, ST_MUT = 2, ST_PLAN = 3, ST_GAMMA = 4, ST_NOISE = 5, ST_KAPPA = 6, ST_ACC = 7 };

/* two uniforms in [0,1) with 53 random bits each, as Julia's rand(Float64) */
void orc_uniform2(uint64_t seed, uint32_t stream, uint32_t sweep, uint32_t unit, uint32_t k, double u[2])
{
    uint32_t ctr[4] = { k, unit, sweep, stream }, key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, o[4];
    orc_philox4x32(ctr, key, o);
    uint64_t a = ((uint64_t)o[1] << 32) | o[0], b = ((uint64_t)o[3] << 32) | o[2];
    u[0] = (double)(a >> 11) * 0x1.0p-53;
    u[1] = (double)(b >> 11) * 0x1.0p-53;
}

static int rand_index(double u, int n) { int i = (int)(u * (double)n); return i >= n ? n - 1 : i; }

/* ------------------------------------------------------------------------------------------ */
/* densities (third-party arithmetic restated from its published closed forms)   */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double s, c; } ksum;   /* Neumaier compensated sum: stands in for Julia's pairwise sum */
static void kadd(ksum *k, double x)
{
    double t = k->s + x;
    if (isfinite(t)) { if (fabs(k->s) >= fabs(x)) k->c += (k->s - t) + x; else k->c += (x - t) + k->s; }
    k->s = t;
}
static double kval(const ksum *k) { return isfinite(k->s) ? k->s + k->c : k->s; }

/* Distributions.normlogpdf: z = (x-mu)/sigma; -(z^2 + log2pi)/2 - log(sigma) */
static double normlogpdf(double mu, double sigma, double x)
{
    double z = (x - mu) / sigma;
    return -(z * z + LOG2PI) / 2.0 - log(sigma);
}
static double norm_cdf(double z) { return 0.5 * erfc(-z * M_SQRT1_2); }
static double norm_pdf(double z) { return exp(-0.5 * z * z) / sqrt(2.0 * M_PI); }
/* StatsFuns.normlogccdf: log(erfc(z/sqrt2)/2), asymptotic series once erfc underflows */
static double normlogccdf(double z)
{
    double x = z * M_SQRT1_2;
    if (x < 25.0) return log(0.5 * erfc(x));
    double x2 = x * x, s = 1.0, term = 1.0;
    for (int k = 1; k < 12; ++k) { term *= -(2.0 * k - 1.0) / (2.0 * x2); s += term; }
    return -x2 - log(x) - 0.5 * LOGPI + log(s) - M_LN2;
}

static double prior_one(const orc_prior *p, const double *theta, int k)
{
    double x = theta[k];
    switch (p->kind) {
    case ORC_PRIOR_FLAT: return 0.0;
    case ORC_PRIOR_NORMAL: return normlogpdf(p->a, p->b, x);
    case ORC_PRIOR_NORMAL_REF: return normlogpdf(p->a, theta[p->ref], x);
    case ORC_PRIOR_HALFCAUCHY: {
        /* logpdf(truncated(Cauchy(a,b),0,Inf),x) = -(log1p(z^2)+log(pi)+log(b)) - log(1-cdf(0)) */
        if (!(x >= 0.0)) return x != x ? NAN : -INFINITY;
        double z = (x - p->a) / p->b;
        double lcdf = atan((0.0 - p->a) / p->b) / M_PI + 0.5;
        return -(log1p(z * z) + LOGPI + log(p->b)) - log(1.0 - lcdf);
    }
    case ORC_PRIOR_UNIFORM:
        if (x != x) return NAN;
        return (x >= p->a && x <= p->b) ? -log(p->b - p->a) : -INFINITY;
    case ORC_PRIOR_BETA: {
        if (x != x) return NAN;
        if (!(x >= 0.0 && x <= 1.0)) return -INFINITY;
        double t1 = (p->a == 1.0) ? 0.0 : (p->a - 1.0) * log(x);
        double t2 = (p->b == 1.0) ? 0.0 : (p->b - 1.0) * log1p(-x);
        return t1 + t2 - (lgamma(p->a) + lgamma(p->b) - lgamma(p->a + p->b));
    }
    }
    return NAN;
}

/* model.prior_loglike(theta) -- e.g. Examples/Gaussian_Example.jl:11-16 (LL += ... in order) */
double orc_prior_loglike(const orc_model *m, const double *theta)
{
    double LL = 0.0;
    for (int k = 0; k < m->d; ++k) LL += prior_one(&m->prior[k], theta, k);
    return LL;
}

/* Examples/Gaussian_Example.jl:26-28: sum(logpdf.(Normal(mu,sigma), data)) */
static double ll_gaussian(const orc_model *m, const double *th)
{
    ksum s = { 0, 0 };
    for (int64_t i = 0; i < m->n_obs; ++i) kadd(&s, normlogpdf(th[0], th[1], m->x[i]));
    return kval(&s);
}

/* Examples/Multivariate_Guassian_Example.jl:31-33: sum(logpdf(MvNormal(mu, sigma^2 I), data));
 * per column -(d*log2pi + d*log(sigma^2))/2 - sqmahal/2, sqmahal = sum((x-mu)^2)/sigma^2 */
/* The timing variant (orc_set_plain_sums(1), bench.py's CPU arm only): plain fp64 accumulation in four independent
 * partial sums per row -- what Distributions' sqmahal + Julia's sum compile to -- instead of the compensated sums the
 * parity oracle uses to sit within 1e-12 of both Julia and the device.  Same formula, ~3x the speed. */
static int g_plain_sums = 0;
void orc_set_plain_sums(int on) { g_plain_sums = on; }
static double ll_mvnormal_plain(const orc_model *m, const double *th)
{
    int dm = m->n_dim;
    double sig = th[dm], s2 = sig * sig;
    double c0 = -((double)dm * LOG2PI + (double)dm * log(s2)) / 2.0;
    double total = 0.0;
    for (int64_t i = 0; i < m->n_obs; ++i) {
        const double *x = m->x + i * dm;
        double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
        int k = 0;
        for (; k + 3 < dm; k += 4) {
            double t0 = x[k] - th[k], t1 = x[k + 1] - th[k + 1], t2 = x[k + 2] - th[k + 2], t3 = x[k + 3] - th[k + 3];
            q0 += t0 * t0; q1 += t1 * t1; q2 += t2 * t2; q3 += t3 * t3;
        }
        for (; k < dm; ++k) { double t = x[k] - th[k]; q0 += t * t; }
        total += c0 - (((q0 + q1) + (q2 + q3)) / s2) / 2.0;
    }
    return total;
}

static double ll_mvnormal(const orc_model *m, const double *th)
{
    if (g_plain_sums) return ll_mvnormal_plain(m, th);
    int dm = m->n_dim;
    double sig = th[dm], s2 = sig * sig;
    double c0 = -((double)dm * LOG2PI + (double)dm * log(s2)) / 2.0;
    ksum s = { 0, 0 };
    for (int64_t i = 0; i < m->n_obs; ++i) {
        const double *x = m->x + i * dm;
        ksum q = { 0, 0 };
        for (int k = 0; k < dm; ++k) { double t = x[k] - th[k]; kadd(&q, t * t); }
        kadd(&s, c0 - (kval(&q) / s2) / 2.0);
    }
    return kval(&s);
}

/* sum(logpdf(MvNormal(mu, sigma^2 Sigma), data)), Sigma known: per column -(k log2pi + logdet(sigma^2 Sigma))/2 -
 * sqmahal/2 with sqmahal = z'z, L z = x - mu, sigma^2 Sigma = (sigma L)(sigma L)' (PDMats: whiten through the Cholesky
 * factor).  O(n k^2) per evaluation: the Cholesky factor is rebuilt per call like MvNormal's constructor does. */
static double ll_mvn_full(const orc_model *m, const double *th)
{
    int k = m->n_dim;
    double sig = th[k], s2 = sig * sig;
    double *L = (double *)calloc((size_t)k * k, sizeof(double)), *z = (double *)malloc(sizeof(double) * k);
    double logdet = 0.0;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j <= i; ++j) {
            long double s = m->cov[(size_t)i * k + j];
            for (int q = 0; q < j; ++q) s -= (long double)L[(size_t)i * k + q] * L[(size_t)j * k + q];
            if (i == j) { L[(size_t)i * k + i] = sqrt((double)s); logdet += 2.0 * log(L[(size_t)i * k + i]); }
            else L[(size_t)i * k + j] = (double)(s / L[(size_t)j * k + j]);
        }
    double c0 = -((double)k * LOG2PI + ((double)k * log(s2) + logdet)) / 2.0;
    ksum s = { 0, 0 };
    for (int64_t i = 0; i < m->n_obs; ++i) {
        const double *x = m->x + i * k;
        ksum q = { 0, 0 };
        for (int r = 0; r < k; ++r) {
            long double t = x[r] - th[r];
            for (int c = 0; c < r; ++c) t -= (long double)L[(size_t)r * k + c] * z[c];
            z[r] = (double)(t / L[(size_t)r * k + r]);
            kadd(&q, z[r] * z[r]);
        }
        kadd(&s, c0 - (kval(&q) / s2) / 2.0);
    }
    free(L); free(z);
    return kval(&s);
}

/* test/binomial_tests.jl:15-17: logpdf(Binomial(N,theta),k) */
static double ll_binomial(const orc_model *m, const double *th)
{
    double N = m->x[0], k = m->x[1], p = th[0];
    double lc = lgamma(N + 1.0) - lgamma(k + 1.0) - lgamma(N - k + 1.0);
    double a = (k == 0.0) ? 0.0 : k * log(p);
    double b = (N - k == 0.0) ? 0.0 : (N - k) * log1p(-p);
    return lc + a + b;
}

/* test/lognormal_race_tests.jl:9-12: sum(logpdf(LNR(;nu,tau), data)); winner LogNormal logpdf,
 * losers LogNormal logccdf, decision time t - tau */
static double ll_lnr(const orc_model *m, const double *th)
{
    int nr = m->n_dim;
    double tau = th[nr];
    ksum s = { 0, 0 };
    for (int64_t i = 0; i < m->n_obs; ++i) {
        double x = m->x[i] - tau;
        int c = m->choice[i] - 1;
        double LL = 0.0;
        if (!(x > 0.0)) { kadd(&s, x != x ? NAN : -INFINITY); continue; }
        double lx = log(x);
        for (int r = 0; r < nr; ++r) {
            double sg = m->sigma ? m->sigma[r] : 1.0;
            double z = (lx - th[r]) / sg;
            if (r == c) LL += -(z * z + LOG2PI) / 2.0 - log(sg) - lx;
            else LL += normlogccdf(z);
        }
        kadd(&s, LL);
    }
    return kval(&s);
}

/* Examples/Run_LBA.jl:34-37: sum(logpdf.(LBA(;nu,A,k,tau), choice, rt)), sigma = 1 (Brown &
 * Heathcote 2008 closed form as in SequentialSamplingModels: product of winner density and loser
 * survivors, divided by 1 - P(all drifts negative), floored, then log) */
static double ll_lba(const orc_model *m, const double *th)
{
    int na = m->n_dim;
    double A = th[na], kk = th[na + 1], tau = th[na + 2], b = A + kk;
    double pneg = 1.0;
    for (int r = 0; r < na; ++r) pneg *= norm_cdf(-th[r]);
    ksum s = { 0, 0 };
    for (int64_t i = 0; i < m->n_obs; ++i) {
        double rt = m->x[i];
        int c = m->choice[i] - 1;
        double den = 1.0;
        if (rt < tau) { kadd(&s, m->lba_floor > 0.0 ? log(m->lba_floor) : -INFINITY); continue; }
        double dt = rt - tau;
        for (int r = 0; r < na; ++r) {
            double v = th[r];
            double n1 = (b - A - dt * v) / dt, n2 = (b - dt * v) / dt;
            if (r == c) {
                double f = (-v * norm_cdf(n1) + norm_pdf(n1) + v * norm_cdf(n2) - norm_pdf(n2)) / A;
                den *= (f > 0.0 ? f : (f != f ? f : 0.0));
            } else {
                double F = 1.0 + ((b - A - dt * v) / A) * norm_cdf(n1) - ((b - dt * v) / A) * norm_cdf(n2)
                           + (dt / A) * norm_pdf(n1) - (dt / A) * norm_pdf(n2);
                F = F > 0.0 ? F : (F != F ? F : 0.0);
                den *= (1.0 - F);
            }
        }
        den = den / (1.0 - pneg);
        if (den != den) { kadd(&s, -INFINITY); continue; }
        if (den < m->lba_floor) den = m->lba_floor;
        kadd(&s, log(den));
    }
    return kval(&s);
}

/* Examples/Hierarchical_Example.jl:36-44: per subject sum(logpdf.(Normal(0,sigma), y_s .- (mu+b_s))) */
static double ll_hier(const orc_model *m, const double *th)
{
    int S = m->n_dim, n = m->n_per;
    double mub = th[0], sig = th[S + 2];
    double LL = 0.0;
    for (int s = 0; s < S; ++s) {
        double mu = mub + th[2 + s];
        ksum q = { 0, 0 };
        for (int j = 0; j < n; ++j) kadd(&q, normlogpdf(0.0, sig, m->x[(int64_t)s * n + j] - mu));
        LL += kval(&q);
    }
    return LL;
}

double orc_loglike(const orc_model *m, const double *theta)
{
    switch (m->kind) {
    case ORC_GAUSSIAN: return ll_gaussian(m, theta);
    case ORC_MVNORMAL: return ll_mvnormal(m, theta);
    case ORC_BINOMIAL: return ll_binomial(m, theta);
    case ORC_LNR: return ll_lnr(m, theta);
    case ORC_LBA: return ll_lba(m, theta);
    case ORC_HIER_NORMAL: return ll_hier(m, theta);
    case ORC_MVN_FULL: return ll_mvn_full(m, theta);
    case ORC_RASTRIGIN: {       /* test/optimization_tests.jl:15-23 */
        const double A = 10.0;
        double y = A * (double)m->d;
        for (int i = 0; i < m->d; ++i) y += +(theta[i] * theta[i]) - A * cos(2.0 * M_PI * theta[i]);
        return y;
    }
    }
    return NAN;
}

/* in_bounds (utilities.jl:70-78): inclusive; NaN fails */
static int in_bounds(const orc_config *cfg, const double *theta)
{
    for (int k = 0; k < cfg->d; ++k)
        if (!(theta[k] >= cfg->lo[k] && theta[k] <= cfg->hi[k])) return 0;
    return 1;
}

/* compute_posterior! (utilities.jl:92-99) */
double orc_posterior(const orc_config *cfg, const orc_model *m, const double *theta)
{
    if (cfg->fitness == ORC_FITNESS_FUN) {
        /* evaluate_fun! (utilities.jl:113-120): the objective alone; out of bounds loses every comparison */
        if (in_bounds(cfg, theta)) return orc_loglike(m, theta);
        return cfg->update == ORC_UPDATE_MAXIMIZE ? -INFINITY : INFINITY;
    }
    if (in_bounds(cfg, theta)) return orc_prior_loglike(m, theta) + orc_loglike(m, theta);
    return -INFINITY;
}

/* ------------------------------------------------------------------------------------------ */
/* particle algebra                                                                            */
/* ------------------------------------------------------------------------------------------ */
/* project (utilities.jl:239-246): v1 = sum(p1.*p2), v2 = sum(p2.^2); p2 * (v1/v2) */
void orc_project(const double *p1, const double *p2, int d, double *out)
{
    double v1 = 0.0, v2 = 0.0;
    for (int k = 0; k < d; ++k) { v1 += p1[k] * p2[k]; v2 += p2[k] * p2[k]; }
    double r = v1 / v2;
    for (int k = 0; k < d; ++k) out[k] = p2[k] * r;
}

static double norm2(const double *a, const double *b, int d)
{
    /* LinearAlgebra.norm of the flattened difference */
    long double s = 0.0L;
    for (int k = 0; k < d; ++k) { long double t = (long double)(a[k] - b[k]); s += t * t; }
    return (double)sqrtl(s);
}

/* adjust_loglike (crossover.jl:268-273): log(norm(prop-Pz)^(d-1) / norm(Pt-Pz)^(d-1)) */
double orc_adjust_loglike(const double *pt, const double *prop, const double *pz, int d)
{
    double adj1 = pow(norm2(prop, pz, d), (double)(d - 1));
    double adj2 = pow(norm2(pt, pz, d), (double)(d - 1));
    return log(adj1 / adj2);
}

/* reset! (crossover.jl:336-352): mask false => restore the previous value */
void orc_reset(double *prop, const double *pt, const uint8_t *mask, int d)
{
    for (int k = 0; k < d; ++k) if (!mask[k]) prop[k] = pt[k];
}

/* random_gamma body (crossover.jl:168): ((Pt + g1*(Pm-Pn)) + g2*(Pb-Pt)) + b, n-ary + folds left.
 * pb == NULL gives the fixed/variable gamma form (Pt + g*(Pm-Pn)) + b (crossover.jl:195,222). */
void orc_de_proposal(const double *pt, const double *pm, const double *pn, const double *pb,
                     double g1, double g2, const double *b, int d, double *out)
{
    for (int k = 0; k < d; ++k) {
        double t = pt[k] + (pm[k] - pn[k]) * g1;
        if (pb) t = t + (pb[k] - pt[k]) * g2;
        out[k] = t + b[k];
    }
}

/* snooker_update! (crossover.jl:239-257) */
void orc_snooker_proposal(const double *pt, const double *pz, const double *pm, const double *pn,
                          double g, const double *b, int d, double *out)
{
    double *pd = (double *)calloc(3 * (size_t)d, sizeof(double)), *r1 = pd + d, *r2 = pd + 2 * d;
    for (int k = 0; k < d; ++k) pd[k] = pt[k] - pz[k];
    orc_project(pm, pd, d, r1);
    orc_project(pn, pd, d, r2);
    for (int k = 0; k < d; ++k) out[k] = (pt[k] + (r1[k] - r2[k]) * g) + b[k];
    free(pd);
}

/* accept (utilities.jl:55-58): p = min(1, exp(w' - w + adj)); rand() <= p.  Julia's min
 * propagates NaN, so NaN => reject. */
int orc_accept(double w_prop, double w_cur, double log_adj, double u)
{
    double p = exp(w_prop - w_cur + log_adj);
    if (p > 1.0) p = 1.0;
    return u <= p ? 1 : 0;
}

/* StatsBase.sample(Weights(w)): t = rand()*sum(w); walk the cumulative sum */
static int weighted_walk(const double *w, int n, double u)
{
    ksum s = { 0, 0 };
    for (int i = 0; i < n; ++i) kadd(&s, w[i]);
    double t = u * kval(&s);
    int i = 0;
    double cw = w[0];
    while (cw < t && i < n - 1) { ++i; cw += w[i]; }
    return i;
}

/* select_base (crossover.jl:282-289): softmax without max shift; NaN => raw weights as Weights */
int orc_select_base(const double *w, int n, double u)
{
    double *th = (double *)malloc(sizeof(double) * (size_t)n);
    ksum s = { 0, 0 };
    for (int i = 0; i < n; ++i) { th[i] = exp(w[i]); kadd(&s, th[i]); }
    double tot = kval(&s);
    int bad = 0;
    for (int i = 0; i < n; ++i) { th[i] = th[i] / tot; if (th[i] != th[i]) bad = 1; }
    int r = weighted_walk(bad ? w : th, n, u);
    free(th);
    return r;
}

/* select_particle (migration.jl:89-95): p ~ exp(-w); NaN => findmin(w), no draw */
int orc_select_particle(const double *w, int n, double u, int *drew)
{
    double *th = (double *)malloc(sizeof(double) * (size_t)n);
    ksum s = { 0, 0 };
    for (int i = 0; i < n; ++i) { th[i] = exp(-w[i]); kadd(&s, th[i]); }
    double tot = kval(&s);
    int bad = 0, r;
    for (int i = 0; i < n; ++i) { th[i] = th[i] / tot; if (th[i] != th[i]) bad = 1; }
    if (bad) {
        /* findmin: first NaN wins, else first minimum */
        r = 0;
        for (int i = 0; i < n; ++i) {
            if (w[i] != w[i]) { r = i; break; }
            if (w[i] < w[r]) r = i;
        }
        if (drew) *drew = 0;
    } else {
        r = weighted_walk(th, n, u);
        if (drew) *drew = 1;
    }
    free(th);
    return r;
}

/* shift_particles! (migration.jl:109-116) on tags[g*Np + j]: selected group i slot j_i receives
 * the particle picked from selected group i-1 (group 0 from group n-1) */
void orc_shift(int32_t *tags, const int32_t *groups, const int32_t *slots, int n, int Np)
{
    int32_t *picked = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; ++i) picked[i] = tags[groups[i] * Np + slots[i]];
    for (int i = 0; i < n; ++i) tags[groups[i] * Np + slots[i]] = picked[(i + n - 1) % n];
    free(picked);
}

/* ------------------------------------------------------------------------------------------ */
/* the sampler                                                                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double *theta;   /* Particle.Theta, flattened */
    double weight;
    int id;
} particle;

typedef struct {
    const orc_config *cfg;
    const orc_model *model;
    const orc_tape *tin;
    orc_tape *tout;
    orc_trace *trace;
    int B, P;
    int64_t n_iter, n_rows;
    particle **slot;     /* [P] pointers: groups[g][j] = slot[g*Np+j] */
    uint8_t *accept;     /* [n_rows][P] row fastest, column = id */
    double *lp;
    const double *samples; /* de.samples, [n_rows][d][P] row fastest (resample reads it) */
} sampler;

#define TAPE_GET(field, i, gen) ((S->tin && S->tin->field) ? S->tin->field[i] : (gen))
#define TAPE_PUT(field, i, v) do { if (S->tout && S->tout->field) S->tout->field[i] = (v); } while (0)

/* mh_update! (utilities.jl:201-210) */
static void mh_update(sampler *S, particle *cur, const double *prop, double wprop, double log_adj,
                      double u, int64_t row, int64_t ti)
{
    int acc;
    if (S->cfg->update == ORC_UPDATE_MAXIMIZE) acc = wprop > cur->weight;         /* maximize! (utilities.jl:212-218) */
    else if (S->cfg->update == ORC_UPDATE_MINIMIZE) acc = wprop < cur->weight;    /* minimize! (utilities.jl:220-226) */
    else acc = orc_accept(wprop, cur->weight, log_adj, u);
    if (acc) { memcpy(cur->theta, prop, sizeof(double) * (size_t)S->cfg->d); cur->weight = wprop; }
    /* maximize! / minimize! never touch Particle.accept / Particle.lp (they stay false / 0.0) */
    if (S->cfg->update == ORC_UPDATE_MH) {
        S->accept[row + S->n_rows * cur->id] = (uint8_t)acc;
        S->lp[row + S->n_rows * cur->id] = cur->weight;
    }
    if (S->trace && S->trace->accepted) S->trace->accepted[ti] = (uint8_t)acc;
}

static int64_t rand_index64(double u, int64_t n) { int64_t i = (int64_t)(u * (double)n); return i >= n ? n - 1 : i; }

/* resample (crossover.jl:113-124): Theta = de.samples[row, :, id] */
static void history_theta(const sampler *S, int64_t row, int id, double *out)
{
    const int d = S->cfg->d;
    for (int k = 0; k < d; ++k) out[k] = S->samples[row + S->n_rows * (k + (int64_t)d * id)];
}

/* one group's mutate_or_crossover! (main.jl:199-207) for sweep s (block bl or -1) */
static void update_group(sampler *S, int g, int64_t it, int bl)
{
    const orc_config *c = S->cfg;
    int Np = c->Np, d = c->d;
    int64_t s = it * S->B + (bl < 0 ? 0 : bl);
    int64_t row = it + c->n_initial;       /* de.iter - 1 */
    int64_t de_iter = it + 1 + c->n_initial;
    particle **grp = S->slot + (size_t)g * Np;
    double u2[2];
    double *prop = (double *)malloc(sizeof(double) * (size_t)d * 5);
    double *noise = prop + d, *h0 = prop + 2 * d, *h1 = prop + 3 * d, *h2 = prop + 4 * d, *snap_theta = NULL, *snap_w = NULL;
    /* resample draws from the ub = de.iter - 1 rows stored so far, all P ids (crossover.jl:115-116,124) */
    const int64_t ub = row, n_cell = ub * (int64_t)S->P;
    double *wbuf = (double *)malloc(sizeof(double) * (size_t)Np);
    if (c->base_snapshot) {
        snap_theta = (double *)malloc(sizeof(double) * (size_t)Np * d);
        snap_w = (double *)malloc(sizeof(double) * (size_t)Np);
        for (int j = 0; j < Np; ++j) { memcpy(snap_theta + (size_t)j * d, grp[j]->theta, sizeof(double) * d); snap_w[j] = grp[j]->weight; }
    }

    orc_uniform2(c->seed, ST_MUT, (uint32_t)s, (uint32_t)g, 0, u2);
    double u_mut = TAPE_GET(mut_u, s * c->n_groups + g, u2[0]);
    TAPE_PUT(mut_u, s * c->n_groups + g, u_mut);
    int mutate = u_mut <= c->beta;
    if (S->tin && S->tin->kind) mutate = S->tin->kind[s * S->P + (int64_t)g * Np] == ORC_KIND_MUTATION;

    for (int j = 0; j < Np; ++j) {               /* crossover.jl:12-17 / mutation.jl:16-23: in order */
        particle *pt = grp[j];
        uint32_t unit = (uint32_t)(g * Np + j);
        int64_t ti = s * S->P + unit;
        int kind, i0 = -1, i1 = -1, i2 = -1, r0 = -1, r1 = -1, r2 = -1;
        double g1 = 0.0, g2 = 0.0, u_snk = 0.0, u_base = 0.0, log_adj = 0.0;

        if (mutate) {
            /* mutation! (mutation.jl:13-25): all elements, block ignored (main.jl:205) */
            kind = ORC_KIND_MUTATION;
            for (int k = 0; k < d; k += 2) {
                orc_uniform2(c->seed, ST_NOISE, (uint32_t)s, unit, (uint32_t)(k / 2), u2);
                double r = sqrt(-2.0 * log(1.0 - u2[0])), ang = 2.0 * M_PI * u2[1];
                noise[k] = 0.0 + c->sigma * (r * cos(ang));
                if (k + 1 < d) noise[k + 1] = 0.0 + c->sigma * (r * sin(ang));
            }
            for (int k = 0; k < d; ++k) {
                noise[k] = TAPE_GET(noise, ti * d + k, noise[k]);
                TAPE_PUT(noise, ti * d + k, noise[k]);
                prop[k] = pt->theta[k] + noise[k];
            }
        } else {
            /* crossover!(model,de,group,pt[,block]) (crossover.jl:30-47, 80-99) */
            orc_uniform2(c->seed, ST_PLAN, (uint32_t)s, unit, 0, u2);
            u_snk = TAPE_GET(u_snk, ti, u2[0]);
            u_base = TAPE_GET(u_base, ti, u2[1]);
            int snooker = u_snk <= c->theta_snooker;
            if (S->tin && S->tin->kind) snooker = S->tin->kind[ti] == ORC_KIND_SNOOKER;
            kind = snooker ? ORC_KIND_SNOOKER : ORC_KIND_DE;
            double ui[4];
            orc_uniform2(c->seed, ST_PLAN, (uint32_t)s, unit, 1, ui);
            orc_uniform2(c->seed, ST_PLAN, (uint32_t)s, unit, 2, ui + 2);
            orc_uniform2(c->seed, ST_GAMMA, (uint32_t)s, unit, 0, u2);
            /* noise b ~ Uniform(-eps, eps) per element (crossover.jl:166-168, utilities.jl:291-306) */
            for (int k = 0; k < d; k += 2) {
                double ub[2];
                orc_uniform2(c->seed, ST_NOISE, (uint32_t)s, unit, (uint32_t)(k / 2), ub);
                noise[k] = -c->eps + (c->eps - (-c->eps)) * ub[0];
                if (k + 1 < d) noise[k + 1] = -c->eps + (c->eps - (-c->eps)) * ub[1];
            }
            for (int k = 0; k < d; ++k) {
                noise[k] = TAPE_GET(noise, ti * d + k, noise[k]);
                TAPE_PUT(noise, ti * d + k, noise[k]);
            }
            if (!snooker) {
                const double *pb = NULL;
                if (c->proposal == ORC_RANDOM_GAMMA) {
                    /* select_base(group) (crossover.jl:156, 282-289) */
                    if (c->base_snapshot) i0 = orc_select_base(snap_w, Np, u_base);
                    else { for (int q = 0; q < Np; ++q) wbuf[q] = grp[q]->weight; i0 = orc_select_base(wbuf, Np, u_base); }
                    if (S->tin && S->tin->idx) i0 = S->tin->idx[ti * 3 + 0];
                    pb = c->base_snapshot ? snap_theta + (size_t)i0 * d : grp[i0]->theta;
                }
                /* Pm,Pn = sample(setdiff(group,[Pt]), 2; replace=false) (crossover.jl:158-160);
                 * StatsBase.samplepair: i1=rand(1:n); i2=rand(1:n-1); i2==i1 && (i2=n) */
                const double *pm, *pn;
                if (c->resample) {
                    /* sample(CartesianIndices(samples[1:ub, 1, :]), 2; replace=false): samplepair over
                     * the ub x P cells, column-major cell -> (row, id) */
                    int64_t a = rand_index64(ui[0], n_cell), b = rand_index64(ui[1], n_cell - 1);
                    if (b == a) b = n_cell - 1;
                    r1 = (int)(a % ub); i1 = (int)(a / ub); r2 = (int)(b % ub); i2 = (int)(b / ub);
                    if (S->tin && S->tin->idx) { i1 = S->tin->idx[ti * 3 + 1]; i2 = S->tin->idx[ti * 3 + 2]; r1 = S->tin->idx_row[ti * 3 + 1]; r2 = S->tin->idx_row[ti * 3 + 2]; }
                    history_theta(S, r1, i1, h1); history_theta(S, r2, i2, h2);
                    pm = h1; pn = h2;
                } else {
                int n = Np - 1;
                int a = rand_index(ui[0], n), b = rand_index(ui[1], n - 1);
                if (b == a) b = n - 1;
                i1 = a >= j ? a + 1 : a;
                i2 = b >= j ? b + 1 : b;
                if (S->tin && S->tin->idx) { i1 = S->tin->idx[ti * 3 + 1]; i2 = S->tin->idx[ti * 3 + 2]; }
                pm = grp[i1]->theta; pn = grp[i2]->theta;
                }
                if (c->proposal == ORC_RANDOM_GAMMA) {
                    g1 = 0.5 + (1.0 - 0.5) * u2[0];                       /* crossover.jl:162 */
                    g2 = de_iter > c->burnin ? 0.0 : 0.5 + (1.0 - 0.5) * u2[1]; /* :164 */
                } else if (c->proposal == ORC_FIXED_GAMMA) {
                    g1 = 2.38;                                            /* :191 */
                } else {
                    g1 = 2.38 / sqrt(2.0 * (double)d);                    /* :218 */
                }
                g1 = TAPE_GET(gamma1, ti, g1);
                g2 = TAPE_GET(gamma2, ti, g2);
                orc_de_proposal(pt->theta, pm, pn, pb, g1, g2, noise, d, prop);
            } else {
                /* Pz,Pm,Pn = sample(group, 3; replace=false): whole group incl. Pt (crossover.jl:241) */
                if (c->resample) {
                    /* three distinct cells of the history (crossover.jl:241 with de.sample = resample) */
                    int64_t a = rand_index64(ui[0], n_cell), b = rand_index64(ui[1], n_cell - 1);
                    if (b >= a) ++b;
                    int64_t cc = rand_index64(ui[2], n_cell - 2), lo = a < b ? a : b, hi = a < b ? b : a;
                    if (cc >= lo) ++cc;
                    if (cc >= hi) ++cc;
                    r0 = (int)(a % ub); i0 = (int)(a / ub); r1 = (int)(b % ub); i1 = (int)(b / ub); r2 = (int)(cc % ub); i2 = (int)(cc / ub);
                    if (S->tin && S->tin->idx) {
                        i0 = S->tin->idx[ti * 3]; i1 = S->tin->idx[ti * 3 + 1]; i2 = S->tin->idx[ti * 3 + 2];
                        r0 = S->tin->idx_row[ti * 3]; r1 = S->tin->idx_row[ti * 3 + 1]; r2 = S->tin->idx_row[ti * 3 + 2];
                    }
                    history_theta(S, r0, i0, h0); history_theta(S, r1, i1, h1); history_theta(S, r2, i2, h2);
                } else {
                int a = rand_index(ui[0], Np), b = rand_index(ui[1], Np - 1);
                if (b >= a) ++b;
                int cc = rand_index(ui[2], Np - 2), lo = a < b ? a : b, hi = a < b ? b : a;
                if (cc >= lo) ++cc;
                if (cc >= hi) ++cc;
                i0 = a; i1 = b; i2 = cc;
                if (S->tin && S->tin->idx) { i0 = S->tin->idx[ti * 3]; i1 = S->tin->idx[ti * 3 + 1]; i2 = S->tin->idx[ti * 3 + 2]; }
                memcpy(h0, grp[i0]->theta, sizeof(double) * d);   /* Pz as it is NOW (adjust_loglike reads it after the proposal) */
                memcpy(h1, grp[i1]->theta, sizeof(double) * d); memcpy(h2, grp[i2]->theta, sizeof(double) * d);
                }
                g1 = 1.2 + (2.2 - 1.2) * u2[0];                           /* crossover.jl:249 */
                g1 = TAPE_GET(gamma1, ti, g1);
                orc_snooker_proposal(pt->theta, h0, h1, h2, g1, noise, d, prop);
            }
            /* recombination! (crossover.jl:301-321) */
            if (c->kappa != 1.0) {
                for (int k = 0; k < d; ++k) {
                    double uk[2];
                    orc_uniform2(c->seed, ST_KAPPA, (uint32_t)s, unit, (uint32_t)(k / 2), uk);
                    int keep = uk[k & 1] <= (1.0 - c->kappa);
                    keep = TAPE_GET(keep, ti * d + k, keep);
                    TAPE_PUT(keep, ti * d + k, (uint8_t)keep);
                    if (keep) prop[k] = pt->theta[k];
                }
            }
            /* reset! (crossover.jl:84,93) then adjust_loglike (crossover.jl:85) */
            if (bl >= 0) orc_reset(prop, pt->theta, c->blocks + (size_t)bl * d, d);
            if (snooker) log_adj = orc_adjust_loglike(pt->theta, prop, h0, d);
        }
        TAPE_PUT(kind, ti, (uint8_t)kind);
        TAPE_PUT(u_snk, ti, u_snk);
        TAPE_PUT(u_base, ti, u_base);
        TAPE_PUT(gamma1, ti, g1);
        TAPE_PUT(gamma2, ti, g2);
        if (S->tout && S->tout->idx) { S->tout->idx[ti * 3] = i0; S->tout->idx[ti * 3 + 1] = i1; S->tout->idx[ti * 3 + 2] = i2; }
        if (S->tout && S->tout->idx_row) { S->tout->idx_row[ti * 3] = r0; S->tout->idx_row[ti * 3 + 1] = r1; S->tout->idx_row[ti * 3 + 2] = r2; }

        double wprop = orc_posterior(c, S->model, prop);          /* evaluate_fitness! */
        orc_uniform2(c->seed, ST_ACC, (uint32_t)s, unit, 0, u2);
        double u_acc = TAPE_GET(u_acc, ti, u2[0]);
        TAPE_PUT(u_acc, ti, u_acc);
        if (S->trace) {
            if (S->trace->prop_theta) memcpy(S->trace->prop_theta + ti * d, prop, sizeof(double) * d);
            if (S->trace->prop_weight) S->trace->prop_weight[ti] = wprop;
            if (S->trace->log_adj) S->trace->log_adj[ti] = log_adj;
        }
        mh_update(S, pt, prop, wprop, log_adj, u_acc, row, ti);   /* update_particle! */
    }
    free(prop); free(wbuf); free(snap_theta); free(snap_w);
}

/* migration! (migration.jl:11-19) */
static void migration(sampler *S, int64_t it)
{
    const orc_config *c = S->cfg;
    int G = c->n_groups, Np = c->Np;
    double u2[2];
    orc_uniform2(c->seed, ST_MIG, (uint32_t)it, 0, 0, u2);
    double u_mig = TAPE_GET(mig_u, it, u2[0]);
    TAPE_PUT(mig_u, it, u_mig);
    for (int i = 0; i < G; ++i) { TAPE_PUT(mig_groups, it * G + i, -1); TAPE_PUT(mig_slots, it * G + i, -1); TAPE_PUT(mig_pick_u, it * G + i, 0.0); }
    if (!(u_mig <= c->alpha)) { TAPE_PUT(mig_n, it, 0); return; }

    /* select_groups (migration.jl:56-60): N = rand(2:G); ordered subset without replacement */
    int N = 2 + rand_index(u2[1], G - 1);
    int *order = (int *)malloc(sizeof(int) * (size_t)G * 3), *arr = order + G, *slots = order + 2 * G;
    double *upick = (double *)malloc(sizeof(double) * (size_t)G);
    for (int i = 0; i < G; ++i) arr[i] = i;
    for (int i = 0; i < N; ++i) {
        orc_uniform2(c->seed, ST_MIG, (uint32_t)it, 0, (uint32_t)(1 + i), u2);
        int jj = i + rand_index(u2[0], G - i), t = arr[i];
        arr[i] = arr[jj]; arr[jj] = t;
        order[i] = arr[i];
        upick[i] = u2[1];
    }
    if (S->tin && S->tin->mig_n) {
        N = S->tin->mig_n[it];
        for (int i = 0; i < N; ++i) order[i] = S->tin->mig_groups[it * G + i];
    }
    TAPE_PUT(mig_n, it, N);
    /* select_particles (migration.jl:71-79): all picks before any shift */
    double *w = (double *)malloc(sizeof(double) * (size_t)Np);
    for (int i = 0; i < N; ++i) {
        upick[i] = TAPE_GET(mig_pick_u, it * G + i, upick[i]);
        for (int q = 0; q < Np; ++q) w[q] = S->slot[order[i] * Np + q]->weight;
        slots[i] = orc_select_particle(w, Np, upick[i], NULL);
        TAPE_PUT(mig_groups, it * G + i, order[i]);
        TAPE_PUT(mig_pick_u, it * G + i, upick[i]);
        TAPE_PUT(mig_slots, it * G + i, slots[i]);
    }
    /* shift_particles! (migration.jl:109-116): the Particle OBJECT moves (theta, weight, id, history) */
    particle **picked = (particle **)malloc(sizeof(particle *) * (size_t)N);
    for (int i = 0; i < N; ++i) picked[i] = S->slot[order[i] * Np + slots[i]];
    for (int i = 0; i < N; ++i) S->slot[order[i] * Np + slots[i]] = picked[(i + N - 1) % N];
    free(picked); free(w); free(upick); free(order);
}

int orc_run(const orc_config *cfg, const orc_model *model, const double *theta0, int64_t n_iter,
            const orc_tape *tape_in, orc_tape *tape_out, orc_trace *trace,
            double *samples, uint8_t *accept, double *lp,
            int32_t *final_id, double *final_theta, double *final_weight)
{
    if (!cfg || !model || !theta0 || cfg->Np < 3 || cfg->n_groups < 1 || cfg->d != model->d) return -1;
    /* resample needs stored rows to draw from: n_initial prior rows (utilities.jl:35-39), at least 3 cells */
    if (cfg->resample && (!samples || (int64_t)cfg->n_initial * cfg->n_groups * cfg->Np < 3)) return -2;
    if (cfg->resample && tape_in && tape_in->idx && !tape_in->idx_row) return -3;
    if (cfg->n_initial > 0 && !samples) return -2;
    sampler Sv, *S = &Sv;
    memset(S, 0, sizeof(*S));
    int G = cfg->n_groups, Np = cfg->Np, d = cfg->d, P = G * Np;
    S->cfg = cfg; S->model = model; S->tin = tape_in; S->tout = tape_out; S->trace = trace;
    S->B = cfg->n_blocks > 0 ? cfg->n_blocks : 1; S->P = P;
    S->n_iter = n_iter; S->n_rows = n_iter + cfg->n_initial; S->samples = samples;
    int own_acc = accept == NULL, own_lp = lp == NULL;
    S->accept = own_acc ? (uint8_t *)calloc((size_t)S->n_rows * P, 1) : accept;
    S->lp = own_lp ? (double *)calloc((size_t)S->n_rows * P, sizeof(double)) : lp;
    /* init_particle (utilities.jl:13-22): accept = falses, lp = zeros */
    if (!own_acc) memset(accept, 0, (size_t)S->n_rows * P);
    if (!own_lp) memset(lp, 0, sizeof(double) * (size_t)S->n_rows * P);

    particle *pool = (particle *)malloc(sizeof(particle) * (size_t)P);
    double *thetas = (double *)malloc(sizeof(double) * (size_t)P * d);
    S->slot = (particle **)malloc(sizeof(particle *) * (size_t)P);
    memcpy(thetas, theta0, sizeof(double) * (size_t)P * d);
    /* init_particle (utilities.jl:15): with n_initial > 0 the particle starts from samples[1, :, id] */
    if (cfg->n_initial > 0)
        for (int p = 0; p < P; ++p) for (int k = 0; k < d; ++k) thetas[(size_t)p * d + k] = samples[0 + S->n_rows * (k + (int64_t)d * p)];
    /* sample_init (main.jl:263-271): ids 1..P group-major; weight via evaluate_fitness! */
    #pragma omp parallel for schedule(static) num_threads(cfg->n_threads > 1 ? cfg->n_threads : 1)
    for (int p = 0; p < P; ++p) {
        pool[p].theta = thetas + (size_t)p * d;
        pool[p].id = p;
        pool[p].weight = orc_posterior(cfg, model, pool[p].theta);
        S->slot[p] = &pool[p];
    }

    for (int64_t it = 0; it < n_iter; ++it) {              /* _sample loop (main.jl:33-38) */
        int64_t row = it + cfg->n_initial;
        if (G > 1) migration(S, it);                        /* step! (main.jl:85); alpha forced 0 when G==1 (structs.jl:102-105) */
        else { TAPE_PUT(mig_u, it, 1.0); TAPE_PUT(mig_n, it, 0);
               for (int i = 0; i < G; ++i) { TAPE_PUT(mig_groups, it * G + i, -1); TAPE_PUT(mig_slots, it * G + i, -1); TAPE_PUT(mig_pick_u, it * G + i, 0.0); } }
        if (trace) {
            for (int p = 0; p < P; ++p) {
                if (trace->pre_theta) memcpy(trace->pre_theta + ((size_t)it * P + p) * d, S->slot[p]->theta, sizeof(double) * d);
                if (trace->pre_weight) trace->pre_weight[(size_t)it * P + p] = S->slot[p]->weight;
                if (trace->pre_id) trace->pre_id[(size_t)it * P + p] = S->slot[p]->id;
            }
        }
        /* update! / p_update! (main.jl:135-167): groups are independent between migrations */
        #pragma omp parallel for schedule(dynamic, 1) num_threads(cfg->n_threads > 1 ? cfg->n_threads : 1)
        for (int g = 0; g < G; ++g) {
            const int on = cfg->n_blocks > 0 && (!cfg->block_on || it >= cfg->n_block_on || cfg->block_on[it]);   /* de.blocking_on(de) (main.jl:137,162) */
            if (on) for (int bl = 0; bl < cfg->n_blocks; ++bl) update_group(S, g, it, bl); /* block_update! (main.jl:174-179) */
            else update_group(S, g, it, -1);
        }
        /* store_samples! (utilities.jl:161-180): samples[iter, :, p.id] = p.Theta */
        for (int p = 0; p < P; ++p) {
            particle *q = S->slot[p];
            if (samples) for (int k = 0; k < d; ++k) samples[row + S->n_rows * (k + (int64_t)d * q->id)] = q->theta[k];
            if (trace) {
                if (trace->state_theta) memcpy(trace->state_theta + ((size_t)it * P + p) * d, q->theta, sizeof(double) * d);
                if (trace->state_weight) trace->state_weight[(size_t)it * P + p] = q->weight;
                if (trace->state_id) trace->state_id[(size_t)it * P + p] = q->id;
            }
        }
    }
    for (int p = 0; p < P; ++p) {
        if (final_id) final_id[p] = S->slot[p]->id;
        if (final_theta) memcpy(final_theta + (size_t)p * d, S->slot[p]->theta, sizeof(double) * d);
        if (final_weight) final_weight[p] = S->slot[p]->weight;
    }
    if (own_acc) free(S->accept);
    if (own_lp) free(S->lp);
    free(S->slot); free(thetas); free(pool);
    return 0;
}
