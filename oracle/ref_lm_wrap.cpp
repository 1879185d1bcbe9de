// TEST INFRASTRUCTURE ONLY -- the reference's Levenberg-Marquardt driver as object code (oracle/_ref/libref_lm.so):
// /root/reference/Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h and .cpp piped UNMODIFIED into the compiler
// (oracle/Makefile) over the interface stand-ins of oracle/eigencompat/g2o_unit/core/.  This file is appended to the same
// header in a second pipe.  It implements those interfaces -- g2o::Solver, g2o::SparseOptimizer, the vertices
// computeLambdaInit scans -- by forwarding every call to an orc_lm_ops table (orc_lm_ops.h: the operations of one of the
// oracle's optimisers), and runs SparseOptimizer::optimize's outer loop (sparse_optimizer.cpp:354-419, restated below: it
// is the `for` around _algorithm->solve) on top.  So the control law -- lambda initialisation, the trial loop, rho, the
// accept / reject updates of lambda and ni, maxTrialsAfterFailure, the nBad stop added by ORB-SLAM3 -- is the reference's
// own object code, and the linear algebra underneath is the oracle's.  ref_lm_driver has the orc_lm_driver signature: the
// oracle's optimisers take it in place of their restated control law (orc_lm_restated), and tests/test_ref_lm.py holds
// the two equal -- iterations, trials, every lambda, every chi2, the result.  Nothing in the product links this.
#include <cmath>
#include <cstring>

#include "../../../orc_lm_ops.h"

namespace {

struct Run {  // what the adapters observe of the driver's private state
  const orc_lm_ops* ops; void* h; orc_lm_report* rep;
  double lambda = 0, candidate = 0;
  bool pending = false, first = true;
  void trial_chi(double chi) {
    candidate = chi;
    if (rep->trace && rep->trace_rows < 128) { double* r = rep->trace + 4 * (size_t)rep->trace_rows; r[0] = lambda; r[1] = chi; r[2] = NAN; r[3] = -1; }
  }
  void trial_end(bool accepted) {
    if (rep->trace && rep->trace_rows < 128) rep->trace[4 * (size_t)rep->trace_rows + 3] = accepted ? 1 : 0;
    if (accepted) rep->chi_final = candidate;
    rep->trace_rows++; rep->trials++;
  }
};

class OpsVertex : public g2o::OptimizableGraph::Vertex {
 public:
  OpsVertex(Run* r, int v) : _r(r), _v(v) {}
  int dimension() const override { return _r->ops->vertex_dim(_r->h, _v); }
  const double& hessian(int i, int j) const override { _tmp = _r->ops->hessian(_r->h, _v, i, j); return _tmp; }
 private:
  Run* _r; int _v; mutable double _tmp = 0;
};

class OpsOptimizer : public g2o::SparseOptimizer {
 public:
  explicit OpsOptimizer(Run* r) : _r(r) {
    const int n = r->ops->n_vertices(r->h);
    for (int v = 0; v < n; v++) _iv.push_back(new OpsVertex(r, v));
  }
  ~OpsOptimizer() override { for (auto* v : _iv) delete v; }
  void computeActiveErrors() override { _r->ops->compute_errors(_r->h); }
  double activeRobustChi2() const override {
    const double chi = _r->ops->robust_chi2(_r->h);
    if (_r->pending) { _r->trial_chi(chi); _r->pending = false; }
    else { if (_r->first) { _r->rep->chi_first = chi; _r->first = false; } _r->rep->chi_final = chi; }  // currentChi of an iteration
    return chi;
  }
  void push() override { _r->ops->push(_r->h); }
  void pop() override { _r->ops->pop(_r->h); _r->trial_end(false); }
  void discardTop() override { _r->ops->discard_top(_r->h); _r->trial_end(true); }
  void update(const double* x) override { _r->ops->update(_r->h, x); _r->pending = true; }
  const g2o::OptimizableGraph::VertexContainer& indexMapping() const override { return _iv; }
  bool terminate() override { return _r->ops->terminate && _r->ops->terminate(_r->h); }
 private:
  Run* _r;
  g2o::OptimizableGraph::VertexContainer _iv;
};

class OpsSolver : public g2o::Solver {  // the block solver + linear solver of the optimiser, as its oracle restates them
 public:
  OpsSolver(Run* r, g2o::SparseOptimizer* opt) : _r(r) {
    _optimizer = opt;
    _x = r->ops->x(r->h); _b = r->ops->b(r->h); _xSize = r->ops->vector_size(r->h);
  }
  bool buildStructure(bool) override { return true; }
  bool buildSystem() override { _r->ops->build_system(_r->h); return true; }
  bool setLambda(double lambda, bool) override { _r->lambda = lambda; return true; }
  bool solve() override { return _r->ops->solve(_r->h, _r->lambda) != 0; }
  void restoreDiagonal() override {}   // the oracle adds lambda to copies of the diagonal blocks
  bool schur() override { return true; }
 private:
  Run* _r;
};

}  // namespace

// optimizer.optimize(max_iters) with OptimizationAlgorithmLevenberg (+ setUserLambdaInit(lambda_init) when > 0), the way the
// Optimizer functions set it up (Optimizer.cc:1142-1153, :1410-1411; :838-846; :2503-2520).  An orc_lm_driver.
extern "C" int ref_lm_driver(const orc_lm_ops* ops, void* h, int max_iters, double lambda_init, orc_lm_report* rep) {
  orc_lm_report local = {};
  if (!rep) rep = &local;
  Run run{ops, h, rep};
  OpsOptimizer optimizer(&run);
  OpsSolver solver(&run, &optimizer);
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
  const int done = (result == g2o::OptimizationAlgorithm::Fail) ? 0 : cjIterations;
  rep->iterations = done;
  rep->lambda_final = algorithm.currentLambda();
  return done;
}
