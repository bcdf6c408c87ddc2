/* oracle/lbm_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or executed from
 * the product path (metalbm_b200/): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * reference legs may use it, and only as the checker.
 *
 * A plain-C, double-precision CPU restatement of the reference's fused collide-and-stream pull step,
 * written from the reference's semantics (evaluation order included) so that it
 * agrees BIT FOR BIT with the reference's own CPU code compiled without FMA contraction.
 * Parity pin: tests/test_oracle_vs_reference.py runs this file against oracle/_ref (the unmodified
 * reference compiled from /root/reference by oracle/refbuild.py) and against the committed golden
 * vectors in tests/golden/ that the same reference binaries produced.
 *
 * Layout: populations are a plain [Q][nx][ny][nz] array of the GLOBAL interior (z fastest, x slowest;
 * Domain.h:88-91, 205-207 without the halo cells).  The reference fills halo cells from the periodic
 * image before every step (Communication.h:134-180, Boundary.h:45-102, Algorithm.h:336-350); pulling
 * straight from the periodic image is the same thing.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "../include/metalbm_b200.h"

#define MAXQ 33

typedef struct {
  int D, Q;
  int c[MAXQ][3];
  double w[MAXQ];
  double inv_cs2;   /* L::inv_cs2 */
} lattice_t;

/* Lattice.h:80-143 (D2Q5), :145-210 (D2Q9), :460-532 (D3Q15), :535-612 (D3Q19), :614-703 (D3Q27) */
static int lattice_table(int lattice, lattice_t* L) {
  static const int c_d2q5[5][3] = {{0,0,0},{-1,0,0},{1,0,0},{0,-1,0},{0,1,0}};
  static const int c_d2q9[9][3] = {{0,0,0},{-1,1,0},{-1,0,0},{-1,-1,0},{1,-1,0},{1,0,0},{1,1,0},{0,-1,0},{0,1,0}};
  static const int c_d3q15[15][3] = {{0,0,0},{-1,0,0},{-1,-1,-1},{-1,-1,1},{-1,1,-1},{-1,1,1},{1,0,0},{1,1,1},
                                     {1,1,-1},{1,-1,1},{1,-1,-1},{0,-1,0},{0,0,-1},{0,1,0},{0,0,1}};
  static const int c_d3q19[19][3] = {{0,0,0},{-1,0,0},{-1,-1,0},{-1,1,0},{-1,0,-1},{-1,0,1},{1,0,0},{1,1,0},{1,-1,0},
                                     {1,0,1},{1,0,-1},{0,-1,0},{0,0,-1},{0,-1,-1},{0,-1,1},{0,1,0},{0,0,1},{0,1,1},{0,1,-1}};
  static const int c_d3q27[27][3] = {{0,0,0},{-1,0,0},{-1,-1,0},{-1,1,0},{-1,0,-1},{-1,0,1},{-1,-1,-1},{-1,-1,1},{-1,1,-1},
                                     {-1,1,1},{1,0,0},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{1,1,1},{1,1,-1},{1,-1,1},{1,-1,-1},
                                     {0,-1,0},{0,0,-1},{0,-1,-1},{0,-1,1},{0,1,0},{0,0,1},{0,1,1},{0,1,-1}};
  /* multi-speed lattices: Lattice.h:213-288 (D2Q13), :290-370 (D2Q17), :372-458 (D2Q21), :706-803 (D3Q33) */
  static const int c_d2q13[13][3] = {{0,0,0},{-1,0,0},{-1,-1,0},{-1,1,0},{-2,0,0},{1,0,0},{1,-1,0},{1,1,0},{2,0,0},{0,-1,0},
                                     {0,1,0},{0,-2,0},{0,2,0}};
  static const int c_d2q17[17][3] = {{0,0,0},{-1,-1,0},{-1,1,0},{-2,-2,0},{-2,2,0},{-3,0,0},{-3,-3,0},{-3,3,0},{1,-1,0},{1,1,0},
                                     {2,-2,0},{2,2,0},{3,0,0},{3,-3,0},{3,3,0},{0,-3,0},{0,3,0}};
  static const int c_d2q21[21][3] = {{0,0,0},{-1,0,0},{-1,-1,0},{-1,1,0},{-2,0,0},{-2,2,0},{-2,-2,0},{-3,0,0},{1,0,0},{1,-1,0},
                                     {1,1,0},{2,0,0},{2,-2,0},{2,2,0},{3,0,0},{0,-1,0},{0,1,0},{0,-2,0},{0,2,0},{0,-3,0},{0,3,0}};
  static const int c_d3q33[33][3] = {{0,0,0},{-1,0,0},{-1,-1,0},{-1,1,0},{-1,0,-1},{-1,0,1},{-1,-1,-1},{-1,-1,1},{-1,1,-1},
                                     {-1,1,1},{-2,0,0},{1,0,0},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{1,1,1},{1,1,-1},{1,-1,1},
                                     {1,-1,-1},{2,0,0},{0,-1,0},{0,0,-1},{0,-1,-1},{0,-1,1},{0,1,0},{0,0,1},{0,1,1},{0,1,-1},
                                     {0,2,0},{0,-2,0},{0,0,2},{0,0,-2}};
  const int (*c)[3] = NULL;
  L->inv_cs2 = 3.0;   /* Lattice.h:86,151,218,466,541,620 */
  switch (lattice) {
    case MLBM_D2Q13: L->D = 2; L->Q = 13; c = c_d2q13; break;
    case MLBM_D2Q17: L->D = 2; L->Q = 17; c = c_d2q17; L->inv_cs2 = 2.0 / 3.0; break;                 /* Lattice.h:294 */
    case MLBM_D2Q21: L->D = 2; L->Q = 21; c = c_d2q21; L->inv_cs2 = 1.0 / (2.0 / 3.0); break;         /* Lattice.h:377-378 */
    case MLBM_D3Q33: L->D = 3; L->Q = 33; c = c_d3q33; L->inv_cs2 = 1.0 / 0.4156023517935171; break;  /* Lattice.h:710-711 */
    case MLBM_D2Q5: L->D = 2; L->Q = 5; c = c_d2q5; break;
    case MLBM_D2Q9: L->D = 2; L->Q = 9; c = c_d2q9; break;
    case MLBM_D3Q15: L->D = 3; L->Q = 15; c = c_d3q15; break;
    case MLBM_D3Q19: L->D = 3; L->Q = 19; c = c_d3q19; break;
    case MLBM_D3Q27: L->D = 3; L->Q = 27; c = c_d3q27; break;
    default: return -1;
  }
  for (int q = 0; q < L->Q; ++q) {
    int n2 = 0;
    for (int d = 0; d < 3; ++d) { L->c[q][d] = c[q][d]; n2 += c[q][d] * c[q][d]; }
    double w;
    switch (lattice) {
      case MLBM_D2Q5: w = n2 == 0 ? 4.0 / 6.0 : 1.0 / 12.0; break;
      case MLBM_D2Q9: w = n2 == 0 ? 4.0 / 9.0 : (n2 == 1 ? 1.0 / 9.0 : 1.0 / 36.0); break;
      case MLBM_D3Q15: w = n2 == 0 ? 2.0 / 9.0 : (n2 == 1 ? 1.0 / 9.0 : 1.0 / 72.0); break;
      case MLBM_D3Q19: w = n2 == 0 ? 1.0 / 3.0 : (n2 == 1 ? 1.0 / 18.0 : 1.0 / 36.0); break;
      case MLBM_D2Q13: w = n2 == 0 ? 1.0 / 2.0 : (n2 == 1 ? 4.0 / 45.0 : (n2 == 2 ? 1.0 / 30.0 : 1.0 / 360.0)); break;
      case MLBM_D2Q17: w = n2 == 0 ? 0.121527777777777777777778 : (n2 == 2 ? 0.175781250000000000000000 :
                           (n2 == 8 ? 0.014062500000000000000000 : (n2 == 9 ? 0.027777777777777777777778 : 0.001996527777777777777778))); break;
      case MLBM_D2Q21: w = n2 == 0 ? 91. / 324. : (n2 == 1 ? 1. / 12. : (n2 == 2 ? 2. / 27. : (n2 == 4 ? 7. / 360. :
                           (n2 == 8 ? 1. / 432. : 1. / 1620.)))); break;
      case MLBM_D3Q33: w = n2 == 0 ? 0.177627658370520295649084 : (n2 == 1 ? 0.103315974899246818673111 :
                           (n2 == 2 ? 0.000513472406731114352456 : (n2 == 3 ? 0.021333928148672240120078 : 0.004273899693974583187026))); break;
      default: w = n2 == 0 ? 8.0 / 27.0 : (n2 == 1 ? 2.0 / 27.0 : (n2 == 2 ? 1.0 / 54.0 : 1.0 / 216.0)); break;
    }
    L->w[q] = w;
  }
  return 0;
}

int mlbm_oracle_lattice(int lattice, int* D, int* Q, int* celerity /*[Q][3]*/, double* weight /*[Q]*/) {
  lattice_t L;
  if (lattice_table(lattice, &L)) return -1;
  *D = L.D; *Q = L.Q;
  if (celerity) for (int q = 0; q < L.Q; ++q) for (int d = 0; d < 3; ++d) celerity[3 * q + d] = L.c[q][d];
  if (weight) for (int q = 0; q < L.Q; ++q) weight[q] = L.w[q];
  return 0;
}

/* c_q . v with the reference's evaluation order (MathVector::dot, MathVector.h:45-52) */
static double cdot(const lattice_t* L, int q, const double* v) {
  double r = (double)L->c[q][0] * v[0];
  for (int d = 1; d < L->D; ++d) r += (double)L->c[q][d] * v[d];
  return r;
}

static double norm2(const lattice_t* L, const double* v) {
  double r = v[0] * v[0];
  for (int d = 1; d < L->D; ++d) r += v[d] * v[d];
  return r;
}

/* PowerBase (Helpers.h:61-71) */
static double power_base(double arg, int power) {
  if (power == 1) return arg;
  if (power == 0) return 1.0;
  if (power == -1) return 1.0 / arg;
  return pow(arg, power);
}

/* Equilibrium::calculate -- TruncationMa3 (Equilibrium.h:17-34), Exact (Equilibrium.h:60-81, 106-126) */
static double equilibrium(const lattice_t* L, int type, double density, const double* u, double u2, int q) {
  const double s = L->inv_cs2;
  if (type == MLBM_TRUNCATION_MA3) {
    const double cu = cdot(L, q, u);
    const double s2 = s * s, s3 = s * s * s, s4 = s * s * s * s;
    double fEq = 1.0 + cu * s - 0.5 * u2 * s + 0.5 * s2 * cu * cu - 0.5 * s2 * cu * u2 +
                 (cu * cu * cu) * s3 / 6.0 + 0.125 * u2 * u2 * s2 - 0.25 * cu * cu * u2 * s3 +
                 (cu * cu * cu * cu) * s4 / 24.0;
    return density * L->w[q] * fEq;
  }
  double fEq = 1.0;
  for (int d = 0; d < L->D; ++d) {
    fEq *= (2.0 - sqrt(1.0 + 3.0 * u[d] * u[d])) *
           power_base((2 * u[d] + sqrt(1.0 + 3.0 * u[d] * u[d])) / (1.0 - u[d]), L->c[q][d]);
  }
  return density * L->w[q] * fEq;
}

/* ForcingScheme::calculateCollisionSource (ForcingScheme.h:68-78 None, :99-117 Guo, :141-151 ShanChen,
 * :184-197 ExactDifferenceMethod).  tau is the INPUT relaxation time, also under ELBM (Collision.h:44). */
static double collision_source(const lattice_t* L, const mlbm_config* cfg, const double* F, double density,
                               const double* u, double u2, double feq_q, int q) {
  (void)u2;
  switch (cfg->forcing_scheme) {
    case MLBM_GUO: {
      const double s = L->inv_cs2;
      const double cu = cdot(L, q, u);
      double t[3];
      for (int d = 0; d < L->D; ++d) t[d] = ((double)L->c[q][d] - u[d]) + (double)L->c[q][d] * (s * cu);
      double r = t[0] * F[0];
      for (int d = 1; d < L->D; ++d) r += t[d] * F[d];
      return (1.0 - 1.0 / (2.0 * cfg->tau)) * L->w[q] * s * r;
    }
    case MLBM_EXACT_DIFFERENCE: {
      double v[3] = {0, 0, 0};
      for (int d = 0; d < L->D; ++d) v[d] = u[d] + F[d] * (1.0 / density);
      return equilibrium(L, cfg->equilibrium, density, v, norm2(L, v), q) - feq_q;
    }
    default:
      return 0.0;
  }
}

/* Force::setForce at LOCAL interior coordinates (Collision.h:81-88; Force.h:120-125 None, :154-159 Constant,
 * :208-215 Sinusoidal, :262-267 Kolmogorov).  p[d]*2 is unsigned-int arithmetic in the reference. */
static void body_force(const lattice_t* L, const mlbm_config* cfg, const unsigned p[3], double* F) {
  for (int d = 0; d < 3; ++d) F[d] = 0.0;
  switch (cfg->force) {
    case MLBM_FORCE_CONSTANT:
      for (int d = 0; d < L->D; ++d) F[d] = cfg->force_amplitude[d];
      break;
    case MLBM_FORCE_SINUSOIDAL:
      for (int d = 0; d < L->D; ++d)
        F[d] = cfg->force_amplitude[d] * sin(p[d] * 2 * M_PI / cfg->force_wavelength[d]);
      break;
    case MLBM_FORCE_KOLMOGOROV:
      F[0] = cfg->force_amplitude[0] * sin(p[1] * 2 * M_PI / cfg->force_wavelength[0]);
      break;
    default:
      break;
  }
}

/* EntropicStepFunctor<T,false> (EntropicStep.h:31-62) */
static double entropic_function(const lattice_t* L, const double* f, const double* fNeq, double alpha) {
  double r = 0.0;
  for (int q = 0; q < L->Q; ++q) {
    double f_q = f[q];
    double g = f_q - alpha * fNeq[q];
    r += f_q * log(f_q / L->w[q]) - g * log(g / L->w[q]);
  }
  return r;
}

static double entropic_derivative(const lattice_t* L, const double* f, const double* fNeq, double alpha) {
  double r = 0.0;
  for (int q = 0; q < L->Q; ++q) {
    double g = f[q] - alpha * fNeq[q];
    r += fNeq[q] * (1 + log(g / L->w[q]));
  }
  return r;
}

/* NewtonRaphsonSolver (EntropicStep.h:111-140) */
static int newton_raphson(const lattice_t* L, const double* f, const double* fNeq, double tolerance,
                          int iterationMax, double* xR, double xMin, double xMax, int* iterations) {
  double xStep = 0.0;
  for (int iteration = 1; iteration <= iterationMax; ++iteration) {
    *xR = *xR - xStep;
    double fx = entropic_function(L, f, fNeq, *xR);
    double dfx = entropic_derivative(L, f, fNeq, *xR);
    xStep = fx / dfx;
    double error = fabs(xStep);
    if (iterations) *iterations = iteration;
    if (error <= tolerance) return (*xR > xMin && *xR < xMax) ? 1 : 0;
  }
  return 0;
}

/* Collision<ELBM>::calculateAlpha (Collision.h:351-375) with isDeviationSmall (:284-303),
 * calculateAlphaMax (:305-326) and solveAlpha (:328-349).  branch: 0 = small deviation, 1 = alphaMax < 2,
 * 2 = Newton converged in range, 3 = Newton failed (alpha = 2). */
static double calculate_alpha(const lattice_t* L, const double* f, const double* fNeq, double alphaGuess,
                              int* branch, int* iterations) {
  int small = 1;
  for (int q = 0; q < L->Q; ++q) {
    double deviation = fabs(fNeq[q]) / f[q];
    if (deviation > 1.0e-3) small = 0;
  }
  if (iterations) *iterations = 0;
  if (small) { if (branch) *branch = 0; return 2.0; }

  double alphaMax = 2.5;
  for (int q = 0; q < L->Q; ++q) {
    if (fNeq[q] > 0) {
      double t = fabs(f[q]) / fNeq[q];
      if (t < alphaMax) alphaMax = t;
    }
  }
  if (alphaMax < 2.) { if (branch) *branch = 1; return 0.95 * alphaMax; }

  double alpha = alphaGuess;
  int ok = newton_raphson(L, f, fNeq, 1e-8, 50, &alpha, 1., alphaMax, iterations);
  if (branch) *branch = ok ? 2 : 3;
  return ok ? alpha : 2.0;
}

/* Collision<ForcedNR_ELBM_Forcing>::calculateAlpha (Collision.h:792-808) with its own calculateAlphaMax (:810-832,
 * min over fNeq_q > 0 of |ff_q / fNeq_q|, start 2.5) and solveAlpha (:835-855) on the mirror functor
 * EntropicStepFunctor<T,true> (EntropicStep.h:65-108), i.e. the ELBM entropy condition written for the FORCED
 * populations ff = f + S.  No small-deviation shortcut.  branch: 1 = alphaMax < 2, 2 = Newton converged in range,
 * 3 = Newton failed (alpha = 2). */
static double calculate_alpha_forcing(const lattice_t* L, const double* ff, const double* fNeq, double alphaGuess,
                                      int* branch, int* iterations) {
  double alphaMax = 2.5;
  for (int q = 0; q < L->Q; ++q) {
    if (fNeq[q] > 0) {
      double t = fabs(ff[q] / fNeq[q]);
      if (t < alphaMax) alphaMax = t;
    }
  }
  if (iterations) *iterations = 0;
  if (alphaMax < 2.) { if (branch) *branch = 1; return 0.95 * alphaMax; }
  double alpha = alphaGuess;
  int ok = newton_raphson(L, ff, fNeq, 1e-8, 50, &alpha, 1., alphaMax, iterations);
  if (branch) *branch = ok ? 2 : 3;
  return ok ? alpha : 2.0;
}

/* Checker-side diagnostic (not in the reference): how far double rounding alone can move the Newton iterate.
 * F(alpha) is a difference of two sums of magnitude ~rho that cancel to O(fNeq^2) and F'(alpha) is O(fNeq^2) as
 * well (sum fNeq = 0), so one Newton step carries an absolute uncertainty of about
 *     eps * sum_q (|f_q| (1 + |ln f_q/w_q|) + |g_q| (1 + |ln g_q/w_q|)) / |F'(alpha)|,   g = f - alpha fNeq.
 * Two implementations that differ only in rounding (libm vs device log, FMA contraction) cannot agree on alpha
 * better than a small multiple of this number; the parity tests add it to the 1e-10 alpha tolerance. */
static double alpha_rounding_noise(const lattice_t* L, const double* f, const double* fNeq, double alpha) {
  double magnitude = 0.0;
  for (int q = 0; q < L->Q; ++q) {
    const double g = f[q] - alpha * fNeq[q];
    if (!(f[q] > 0.0) || !(g > 0.0)) return 1.0;
    magnitude += fabs(f[q]) * (1.0 + fabs(log(f[q] / L->w[q]))) + fabs(g) * (1.0 + fabs(log(g / L->w[q])));
  }
  const double derivative = fabs(entropic_derivative(L, f, fNeq, alpha));
  return derivative > 0.0 ? 2.220446049250313e-16 * magnitude / derivative : 1.0;
}

static size_t wrap(long i, long n) { return (size_t)((i % n + n) % n); }

/* One Algorithm::iterate (Algorithm.h:326-358) over the global domain = the per-node functor
 * Algorithm::operator() (Algorithm.h:97-126) at every node.
 *   prev, next : [Q][nx][ny][nz]      alpha : [nx][ny][nz] read (warm start) and written every step
 *   density, velocity[D], force[D] : written when is_stored (Algorithm::storeFields, Algorithm.h:150-194);
 *   force[D] is also READ, every step, when cfg->force == MLBM_FORCE_FIELD
 *   branch, iterations (optional, [nx][ny][nz] int32): which alpha branch each node took. */
int mlbm_oracle_step_ex(const mlbm_config* cfg, const double* prev, double* next, double* alpha, double* density,
                        double* velocity, double* force, int is_stored, int* branchOut, int* iterationsOut,
                        double* alphaNoiseOut, double* fNeqMaxOut) {
  lattice_t L;
  if (lattice_table(cfg->lattice, &L)) return -1;
  if (cfg->equilibrium == MLBM_EXACT && !(cfg->lattice == MLBM_D2Q9 || cfg->lattice == MLBM_D3Q27)) return -1;
  const long nx = cfg->global_length[0], ny = cfg->global_length[1], nz = L.D > 2 ? cfg->global_length[2] : 1;
  const size_t V = (size_t)nx * ny * nz;
  const long lx = nx / (cfg->nranks > 0 ? cfg->nranks : 1);
  const double beta = 1.0 / (2.0 * cfg->tau);
  const int entropic = cfg->collision != MLBM_BGK;

  for (long x = 0; x < nx; ++x)
    for (long y = 0; y < ny; ++y)
      for (long z = 0; z < nz; ++z) {
        const size_t idx = ((size_t)x * ny + y) * nz + z;
        double f[MAXQ], fNeq[MAXQ];
        for (int q = 0; q < L.Q; ++q) {
          size_t src = (wrap(x - L.c[q][0], nx) * ny + wrap(y - L.c[q][1], ny)) * nz + wrap(z - L.c[q][2], nz);
          f[q] = prev[(size_t)q * V + src];
        }
        /* Moment::calculateDensity / calculateVelocity (Moment.h:14-47) */
        double rho = f[0];
        for (int q = 1; q < L.Q; ++q) rho += f[q];
        double u[3] = {0, 0, 0};
        for (int d = 0; d < L.D; ++d) {
          double m = (double)L.c[0][d] * f[0];
          for (int q = 1; q < L.Q; ++q) m += (double)L.c[q][d] * f[q];
          u[d] = m / rho;
        }
        const double u2 = norm2(&L, u);

        unsigned p[3] = {(unsigned)(x % lx), (unsigned)y, (unsigned)z};
        double F[3];
        body_force(&L, cfg, p, F);
        if (cfg->force >= MLBM_FORCE_FIELD) {  /* array-type forces; the spectral ones are made by oracle.py */
          /* Force<T, ForceType::Generic>::setForce (Force.h:39-48): component iD of the force array at the node's index
           * (the array-type forces ConstantShell / EnergyRemoval / Turbulent2D, Force.h:296-623, fill it outside the step) */
          if (!force) return -1;
          for (int d = 0; d < L.D; ++d) F[d] = force[(size_t)d * V + idx];
        }

        double a = 2.0;
        if (cfg->collision == MLBM_FORCED_NR_ELBM_FORCING) {
          /* Collision<ForcedNR_ELBM_Forcing>::calculateRelaxationTime (Collision.h:757-778): next = f + S(feq), the pulled
           * population of `prev` is overwritten by fNeq = f - feq; then alpha on (next, fNeq) and
           * collideAndStream (:780-789): next -= 1/tau * fNeq */
          double ff[MAXQ];
          for (int q = 0; q < L.Q; ++q) {
            const double feq_q = equilibrium(&L, cfg->equilibrium, rho, u, u2, q);
            ff[q] = f[q] + collision_source(&L, cfg, F, rho, u, u2, feq_q, q);
            fNeq[q] = f[q] - feq_q;
          }
          int branch = 0, iterations = 0;
          a = calculate_alpha_forcing(&L, ff, fNeq, alpha[idx], &branch, &iterations);
          if (branchOut) branchOut[idx] = branch;
          if (iterationsOut) iterationsOut[idx] = iterations;
          if (alphaNoiseOut) alphaNoiseOut[idx] = branch >= 2 ? alpha_rounding_noise(&L, ff, fNeq, a) : 0.0;
          if (fNeqMaxOut) {
            double m = 0.0;
            for (int q = 0; q < L.Q; ++q) if (fabs(fNeq[q]) > m) m = fabs(fNeq[q]);
            fNeqMaxOut[idx] = m;
          }
          const double tau = 1.0 / (a * beta);
          for (int q = 0; q < L.Q; ++q) next[(size_t)q * V + idx] = ff[q] - 1.0 / tau * fNeq[q];
        } else if (entropic) {
          /* Collision<ELBM>::calculateRelaxationTime (Collision.h:227-241) */
          for (int q = 0; q < L.Q; ++q) fNeq[q] = f[q] - equilibrium(&L, cfg->equilibrium, rho, u, u2, q);
          int branch = 0, iterations = 0;
          a = calculate_alpha(&L, f, fNeq, alpha[idx], &branch, &iterations);
          if (branchOut) branchOut[idx] = branch;
          if (iterationsOut) iterationsOut[idx] = iterations;
          if (alphaNoiseOut) alphaNoiseOut[idx] = branch >= 2 ? alpha_rounding_noise(&L, f, fNeq, a) : 0.0;
          if (fNeqMaxOut) {
            double m = 0.0;
            for (int q = 0; q < L.Q; ++q) if (fabs(fNeq[q]) > m) m = fabs(fNeq[q]);
            fNeqMaxOut[idx] = m;
          }
          const double tau = 1.0 / (a * beta);
          /* Collision<ELBM>::collideAndStream (Collision.h:243-258) */
          for (int q = 0; q < L.Q; ++q) {
            const double feq_q = f[q] - fNeq[q];
            next[(size_t)q * V + idx] =
                f[q] - 1.0 / tau * fNeq[q] + collision_source(&L, cfg, F, rho, u, u2, feq_q, q);
          }
        } else {
          /* Collision<BGK>::collideAndStream (Collision.h:134-151) */
          for (int q = 0; q < L.Q; ++q) {
            const double feq_q = equilibrium(&L, cfg->equilibrium, rho, u, u2, q);
            next[(size_t)q * V + idx] =
                (1. - 2. * beta) * f[q] + 2. * beta * feq_q + collision_source(&L, cfg, F, rho, u, u2, feq_q, q);
          }
        }
        alpha[idx] = a; /* Algorithm.h:105-106 */

        if (is_stored) {
          density[idx] = rho;
          for (int d = 0; d < L.D; ++d) {
            /* ForcingScheme::calculateHydrodynamicVelocity (ForcingScheme.h:26-33; None :50-57) */
            velocity[(size_t)d * V + idx] =
                cfg->forcing_scheme == MLBM_SCHEME_NONE ? u[d] : u[d] + F[d] * (0.5 / rho);
            force[(size_t)d * V + idx] = F[d];
          }
        }
      }
  return 0;
}

int mlbm_oracle_step(const mlbm_config* cfg, const double* prev, double* next, double* alpha, double* density,
                     double* velocity, double* force, int is_stored, int* branchOut, int* iterationsOut) {
  return mlbm_oracle_step_ex(cfg, prev, next, alpha, density, velocity, force, is_stored, branchOut, iterationsOut,
                             NULL, NULL);
}

/* initDistribution, equilibrium branch (Initialize.h:106-117): f = feq(rho, u) */
int mlbm_oracle_init_equilibrium(const mlbm_config* cfg, const double* density, const double* velocity, double* f) {
  lattice_t L;
  if (lattice_table(cfg->lattice, &L)) return -1;
  const size_t V = (size_t)cfg->global_length[0] * cfg->global_length[1] * (L.D > 2 ? cfg->global_length[2] : 1);
  for (size_t i = 0; i < V; ++i) {
    double u[3] = {0, 0, 0};
    for (int d = 0; d < L.D; ++d) u[d] = velocity[(size_t)d * V + i];
    const double u2 = norm2(&L, u);
    for (int q = 0; q < L.Q; ++q) f[(size_t)q * V + i] = equilibrium(&L, cfg->equilibrium, density[i], u, u2, q);
  }
  return 0;
}

/* TotalEnergy (Analysis.h:53-61) + normalize (Analysis.h:30); mass = sum of density (Routine.h:117-118);
 * mach = max |u| / c_s (not in the reference; north_star observable). */
int mlbm_oracle_observables(const mlbm_config* cfg, const double* density, const double* velocity, double* out) {
  lattice_t L;
  if (lattice_table(cfg->lattice, &L)) return -1;
  const size_t V = (size_t)cfg->global_length[0] * cfg->global_length[1] * (L.D > 2 ? cfg->global_length[2] : 1);
  double energy = 0.0, mass = 0.0, mach2 = 0.0;
  for (size_t i = 0; i < V; ++i) {
    double v2 = 0.0;
    for (int d = 0; d < L.D; ++d) {
      const double v = velocity[(size_t)d * V + i];
      energy += 0.5 * density[i] * v * v;
      v2 += v * v;
    }
    if (v2 > mach2) mach2 = v2;
    mass += density[i];
  }
  out[0] = energy / (double)V;
  out[1] = 0.0;
  out[2] = sqrt(mach2 * L.inv_cs2);
  out[3] = mass;
  return 0;
}
