// TEST INFRASTRUCTURE ONLY -- the reference's Levenberg-Marquardt driver as object code (oracle/_ref/libref_lm.so):
// /root/reference/Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h and .cpp piped UNMODIFIED into the compiler
// (oracle/Makefile) over the interface stand-ins of oracle/eigencompat/g2o_unit/core/.  This file is appended to the same
// header in a second pipe.  It implements those interfaces -- g2o::Solver, g2o::SparseOptimizer, the vertices
// computeLambdaInit scans -- by forwarding every call to the oracle's Stepper operations (orc_lba.cpp:
// orc_lba_stepper_*), and runs SparseOptimizer::optimize's outer loop (sparse_optimizer.cpp:354-419, restated below: it is
// the `for` around _algorithm->solve) on top.  So the control law -- lambda initialisation, the trial loop, rho, the
// accept / reject updates of lambda and ni, maxTrialsAfterFailure, the nBad stop added by ORB-SLAM3 -- is the reference's
// own object code, and the linear algebra underneath is the oracle's.  tests/test_ref_lm.py holds orc_lba_solve (the
// restated control law over the same operations) equal to it: iterations, trials, every lambda, every chi2, the result.
// Nothing in the product links this.
#include <cstring>

#include "../../../../include/orb_b200.h"

extern "C" {
void* orc_lba_stepper_open(const lba_graph_view* g);
void orc_lba_stepper_close(void* h);
void orc_lba_stepper_compute_errors(void* h);
double orc_lba_stepper_robust_chi2(void* h);
void orc_lba_stepper_build_system(void* h);
int orc_lba_stepper_n_vertices(void* h);
int orc_lba_stepper_vertex_dim(void* h, int v);
double orc_lba_stepper_hessian(void* h, int v, int i, int j);
int orc_lba_stepper_solve(void* h, double lambda);
double* orc_lba_stepper_x(void* h);
double* orc_lba_stepper_b(void* h);
size_t orc_lba_stepper_vector_size(void* h);
void orc_lba_stepper_update(void* h, const double* x);
void orc_lba_stepper_push(void* h);
void orc_lba_stepper_pop(void* h);
void orc_lba_stepper_discard_top(void* h);
void orc_lba_stepper_results(void* h, double* kf_pose_out, double* mp_pos_out);
}

namespace {

struct Trace {  // one row per trial, like orc_lba_solve's: lambda, tempChi, (rho is private to the driver: NaN), accepted
  double* rows;
  int n;
  double lambda;
  void trial_lambda(double l) { lambda = l; }
  void trial_chi(double chi) { if (rows && n < 128) { rows[4 * n] = lambda; rows[4 * n + 1] = chi; rows[4 * n + 2] = NAN; rows[4 * n + 3] = -1; } }
  void trial_end(bool accepted) { if (rows && n < 128) rows[4 * n + 3] = accepted ? 1 : 0; n++; }
};

class StepperVertex : public g2o::OptimizableGraph::Vertex {
 public:
  StepperVertex(void* h, int v) : _h(h), _v(v) {}
  int dimension() const override { return orc_lba_stepper_vertex_dim(_h, _v); }
  const double& hessian(int i, int j) const override { _tmp = orc_lba_stepper_hessian(_h, _v, i, j); return _tmp; }
 private:
  void* _h; int _v; mutable double _tmp = 0;
};

class StepperOptimizer : public g2o::SparseOptimizer {
 public:
  StepperOptimizer(void* h, Trace* t, const volatile uint8_t* stop) : _h(h), _t(t), _stop(stop), _pending(false) {
    const int n = orc_lba_stepper_n_vertices(h);
    for (int v = 0; v < n; v++) _iv.push_back(new StepperVertex(h, v));
  }
  ~StepperOptimizer() override { for (auto* v : _iv) delete v; }
  void computeActiveErrors() override { orc_lba_stepper_compute_errors(_h); }
  double activeRobustChi2() const override {
    const double chi = orc_lba_stepper_robust_chi2(_h);
    if (_pending) { _t->trial_chi(chi); _pending = false; }
    return chi;
  }
  void push() override { orc_lba_stepper_push(_h); }
  void pop() override { orc_lba_stepper_pop(_h); _t->trial_end(false); }
  void discardTop() override { orc_lba_stepper_discard_top(_h); _t->trial_end(true); }
  void update(const double* x) override { orc_lba_stepper_update(_h, x); _pending = true; }
  const g2o::OptimizableGraph::VertexContainer& indexMapping() const override { return _iv; }
  bool terminate() override { return _stop && *_stop; }
 private:
  void* _h; Trace* _t; const volatile uint8_t* _stop; mutable bool _pending;
  g2o::OptimizableGraph::VertexContainer _iv;
};

class StepperSolver : public g2o::Solver {  // BlockSolver<6,3> + LinearSolverEigen as the oracle restates them
 public:
  StepperSolver(void* h, g2o::SparseOptimizer* opt, Trace* t) : _h(h), _t(t), _lambda(0) {
    _optimizer = opt;
    _x = orc_lba_stepper_x(h); _b = orc_lba_stepper_b(h); _xSize = orc_lba_stepper_vector_size(h);
  }
  bool buildStructure(bool) override { return true; }
  bool buildSystem() override { orc_lba_stepper_build_system(_h); return true; }
  bool setLambda(double lambda, bool) override { _lambda = lambda; _t->trial_lambda(lambda); return true; }
  bool solve() override { return orc_lba_stepper_solve(_h, _lambda) != 0; }
  void restoreDiagonal() override {}   // the Stepper adds lambda to copies of the diagonal blocks
  bool schur() override { return true; }
 private:
  void* _h; Trace* _t; double _lambda;
};

}  // namespace

extern "C" {

// optimizer.optimize(max_iters) with OptimizationAlgorithmLevenberg (+ setUserLambdaInit(lambda_init) when > 0), the way
// Optimizer::LocalBundleAdjustment sets it up (Optimizer.cc:1142-1153, :1410-1411).  out4 = iterations, trials,
// final robust chi2, final lambda; trace = 128 rows of (lambda, tempChi, NaN, accepted).  Returns iterations.
int ref_lm_optimize(const lba_graph_view* g, const volatile uint8_t* stop, int max_iters, double lambda_init, double* kf_pose_out,
                    double* mp_pos_out, double* out4, double* trace) {
  void* h = orc_lba_stepper_open(g);
  Trace t{trace, 0, 0.0};
  int done;
  {
    StepperOptimizer optimizer(h, &t, stop);
    StepperSolver solver(h, &optimizer, &t);
    g2o::OptimizationAlgorithmLevenberg algorithm(&solver);
    algorithm.setOptimizer(&optimizer);
    if (lambda_init > 0) algorithm.setUserLambdaInit(lambda_init);
    // SparseOptimizer::optimize (sparse_optimizer.cpp:354-419)
    int cjIterations = 0;
    bool ok = true;
    g2o::OptimizationAlgorithm::SolverResult result = g2o::OptimizationAlgorithm::OK;
    for (int i = 0; i < max_iters && !optimizer.terminate() && ok; i++) {
      result = algorithm.solve(i, false);
      ok = (result == g2o::OptimizationAlgorithm::OK);
      ++cjIterations;
    }
    done = (result == g2o::OptimizationAlgorithm::Fail) ? 0 : cjIterations;
    orc_lba_stepper_compute_errors(h);
    out4[0] = done; out4[1] = t.n; out4[2] = orc_lba_stepper_robust_chi2(h); out4[3] = algorithm.currentLambda();
  }
  orc_lba_stepper_results(h, kf_pose_out, mp_pos_out);
  orc_lba_stepper_close(h);
  return done;
}

}  // extern "C"
