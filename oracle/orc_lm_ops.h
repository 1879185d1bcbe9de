// TEST INFRASTRUCTURE ONLY (see orc_common.h).  The operations g2o's Levenberg-Marquardt driver performs on its Solver and
// SparseOptimizer (optimization_algorithm_levenberg.cpp:61-194), as a table of C functions over an opaque problem: each of
// the oracle's optimisers (local BA, pose-only, inertial BA) provides one, runs its restated control law over it, and can
// be handed another driver instead -- oracle/_ref/libref_lm.so's, which is the reference's own object code
// (tests/test_ref_lm.py holds the two equal).
#pragma once
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_lm_ops {
  void (*compute_errors)(void* h);            /* SparseOptimizer::computeActiveErrors */
  double (*robust_chi2)(void* h);             /* SparseOptimizer::activeRobustChi2 */
  void (*build_system)(void* h);              /* Solver::buildSystem */
  int (*n_vertices)(void* h);                 /* indexMapping().size() */
  int (*vertex_dim)(void* h, int v);          /* Vertex::dimension */
  double (*hessian)(void* h, int v, int i, int j); /* Vertex::hessian(i, j) */
  int (*solve)(void* h, double lambda);       /* Solver::setLambda + solve; 0 = failed, x left as it was */
  double* (*x)(void* h);
  double* (*b)(void* h);
  size_t (*vector_size)(void* h);
  void (*update)(void* h, const double* x);   /* SparseOptimizer::update: oplus on every vertex */
  void (*push)(void* h);
  void (*pop)(void* h);
  void (*discard_top)(void* h);
  int (*terminate)(void* h);                  /* SparseOptimizer::terminate (the force-stop flag); may be NULL */
} orc_lm_ops;

typedef struct orc_lm_report {
  int iterations, trials;      /* optimize()'s return value; LM trials over all iterations */
  double lambda_final;
  double chi_first, chi_final; /* activeRobustChi2 before the first trial / of the last accepted state */
  double* trace;               /* optional, in: 4 doubles per trial from row trace_rows on (cap 128): lambda, chi2, rho, accepted */
  int trace_rows;              /* in / out */
} orc_lm_report;

/* optimizer.optimize(max_iters) over `ops`; lambda_init > 0 = setUserLambdaInit.  Fills *rep (trials and trace_rows are
 * added to); returns the iterations run. */
typedef int (*orc_lm_driver)(const orc_lm_ops* ops, void* h, int max_iters, double lambda_init, orc_lm_report* rep);

/* the restated control law (orc_pose.cpp) */
int orc_lm_restated(const orc_lm_ops* ops, void* h, int max_iters, double lambda_init, orc_lm_report* rep);

#ifdef __cplusplus
}
#endif
