// TEST INFRASTRUCTURE ONLY -- appended (oracle/Makefile) to the reference's own Thirdparty/g2o/g2o/core/robust_kernel.{h,cpp}
// and robust_kernel_impl.{h,cpp}, piped UNMODIFIED into the compiler: RobustKernelHuber::setDelta / robustify as the
// reference's object code (float dsqr, robust_kernel_impl.cpp:65-91), part of oracle/_ref/libref_g2o.so.
// tests/test_ref_edges.py holds the oracle's Huber (orc_se3.h) against it.  Nothing in the product links this.
extern "C" void ref_g2o_huber(double delta, double e, double* rho3) {
  g2o::RobustKernelHuber k;
  k.setDelta(delta);
  Eigen::Vector3d rho;
  k.robustify(e, rho);
  rho3[0] = rho[0]; rho3[1] = rho[1]; rho3[2] = rho[2];
}
