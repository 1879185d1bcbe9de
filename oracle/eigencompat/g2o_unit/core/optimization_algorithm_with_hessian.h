// TEST INFRASTRUCTURE -- NOT g2o.  The interfaces g2o's LM driver is written against, as far as
// Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.{h,cpp} use them: OptimizationAlgorithm(WithHessian), Solver,
// SparseOptimizer, OptimizableGraph::Vertex, Property / PropertyMap, G2OBatchStatistics, get_monotonic_time.  Member names,
// signatures and the SolverResult values follow core/optimization_algorithm.h:46-80, optimization_algorithm_with_hessian.h,
// solver.h:47-132, sparse_optimizer.h and stuff/property.h; there is no implementation here -- the driver's object code
// (oracle/_ref/libref_lm.so, oracle/Makefile) is linked against adapters (oracle/ref_lm_wrap.cpp) that forward every call
// to the oracle's Stepper operations.  The real headers need Eigen (SparseBlockMatrix<MatrixXd> in the signatures of
// computeMarginals, which the LM driver never calls).
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iomanip>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

#define g2o_isfinite(x) std::isfinite(x)
#define FIXED(s) std::fixed << s << std::resetiosflags(std::ios_base::fixed)

namespace g2o {

inline double get_monotonic_time() { return 0.0; }

struct G2OBatchStatistics {  // batch_stats.h: never active here
  static G2OBatchStatistics* globalStats() { return nullptr; }
  double timeResiduals = 0, timeQuadraticForm = 0, timeLinearSolution = 0, timeUpdate = 0;
  int levenbergIterations = 0;
};

template <typename T> class Property {  // stuff/property.h
 public:
  typedef T ValueType;
  Property(const std::string& name, const T& v) : _name(name), _value(v) {}
  void setValue(const T& v) { _value = v; }
  const T& value() const { return _value; }
 private:
  std::string _name;
  T _value;
};
class PropertyMap {
 public:
  ~PropertyMap() { for (Holder* h : _owned) delete h; }
  template <typename P> P* makeProperty(const std::string& name, const typename P::ValueType& v) {
    Typed<P>* h = new Typed<P>(name, v);
    _owned.push_back(h);
    return &h->p;
  }
 private:
  struct Holder { virtual ~Holder() {} };
  template <typename P> struct Typed : Holder { P p; Typed(const std::string& n, const typename P::ValueType& v) : p(n, v) {} };
  std::vector<Holder*> _owned;
};

class OptimizableGraph {
 public:
  class Vertex {
   public:
    virtual ~Vertex() {}
    virtual int dimension() const = 0;
    virtual const double& hessian(int i, int j) const = 0;
  };
  typedef std::vector<Vertex*> VertexContainer;
};

class SparseOptimizer {  // sparse_optimizer.h, the members the LM driver calls
 public:
  virtual ~SparseOptimizer() {}
  virtual void computeActiveErrors() = 0;
  virtual double activeRobustChi2() const = 0;
  virtual void push() = 0;
  virtual void pop() = 0;
  virtual void discardTop() = 0;
  virtual void update(const double* update) = 0;
  virtual const OptimizableGraph::VertexContainer& indexMapping() const = 0;
  virtual bool terminate() = 0;
};

class Solver {  // solver.h
 public:
  Solver() : _optimizer(nullptr), _x(nullptr), _b(nullptr), _xSize(0) {}
  virtual ~Solver() {}
  virtual bool buildStructure(bool zeroBlocks = false) = 0;
  virtual bool buildSystem() = 0;
  virtual bool solve() = 0;
  virtual bool setLambda(double lambda, bool backup = false) = 0;
  virtual void restoreDiagonal() = 0;
  virtual bool schur() = 0;
  double* x() { return _x; }
  const double* x() const { return _x; }
  double* b() { return _b; }
  const double* b() const { return _b; }
  size_t vectorSize() const { return _xSize; }
  SparseOptimizer* optimizer() const { return _optimizer; }
 protected:
  SparseOptimizer* _optimizer;
  double* _x;
  double* _b;
  size_t _xSize;
};

class OptimizationAlgorithm {  // optimization_algorithm.h:46-80
 public:
  enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
  OptimizationAlgorithm() : _optimizer(nullptr) {}
  virtual ~OptimizationAlgorithm() {}
  virtual SolverResult solve(int iteration, bool online = false) = 0;
  virtual void printVerbose(std::ostream& os) const { (void)os; }
  void setOptimizer(SparseOptimizer* optimizer) { _optimizer = optimizer; }
 protected:
  SparseOptimizer* _optimizer;
  PropertyMap _properties;
};

class OptimizationAlgorithmWithHessian : public OptimizationAlgorithm {  // optimization_algorithm_with_hessian.h
 public:
  explicit OptimizationAlgorithmWithHessian(Solver* solver) : _solver(solver) {}
  Solver* solver() { return _solver; }
 protected:
  Solver* _solver;
};

}  // namespace g2o
