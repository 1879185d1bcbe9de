// TEST INFRASTRUCTURE ONLY (see orc_common.h).  Flat graph of Optimizer::LocalInertialBA
// (reference src/Optimizer.cc:2383-2958) as the oracle takes it.  This is the
// candidate C-ABI view for the row; it moves to include/orb_b200.h together with the CUDA path
// (round 2) -- nothing in the product uses it yet.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lia_graph_view {
  /* keyframes: vpOptimizableKFs (newest first), then lFixedKeyFrames */
  int32_t n_kf;
  const double* kf_Rwb;      /* n_kf x 9 row-major: GetImuRotation().cast<double>() */
  const double* kf_twb;      /* n_kf x 3: GetImuPosition() */
  const double* kf_Rcw;      /* n_kf x 9: GetRotation() (left camera) */
  const double* kf_tcw;      /* n_kf x 3: GetTranslation() */
  const uint8_t* kf_fixed;   /* VertexPose (and the IMU vertices) fixed */
  const uint8_t* kf_has_imu; /* KeyFrame::bImu: velocity / gyro-bias / acc-bias vertices exist */
  const double* kf_vel;      /* n_kf x 3: GetVelocity() */
  const double* kf_bg;       /* n_kf x 3: GetGyroBias() */
  const double* kf_ba;       /* n_kf x 3: GetAccBias() */
  double Rcb[9], tcb[3], tbc[3]; /* mImuCalib.mTcb / mTbc (one rig) */
  float fx, fy, cx, cy, bf;
  /* map points and visual edges (EdgeMono / EdgeStereo, left camera) */
  int32_t n_mp;
  const double* mp_pos;      /* n_mp x 3 */
  int32_t n_edges;
  const int32_t* e_kf;
  const int32_t* e_mp;
  const uint8_t* e_stereo;
  const double* e_obs;       /* n_edges x 3: u, v, uRight */
  const float* e_inv_sigma2; /* mvInvLevelSigma2[octave] / uncertainty2 */
  /* inertial edges: EdgeInertial + EdgeGyroRW + EdgeAccRW between kf1 (previous) and kf2 */
  int32_t n_inertial;
  const int32_t* i_kf1;
  const int32_t* i_kf2;
  const float* i_dR;         /* x 9: IMU::Preintegrated::dR, then dV, dP (x 3 each) */
  const float* i_dV;
  const float* i_dP;
  const float* i_JRg;        /* x 9 each: bias Jacobians of the preintegration */
  const float* i_JVg;
  const float* i_JVa;
  const float* i_JPg;
  const float* i_JPa;
  const float* i_bias;       /* x 6: linearisation bias b = (bax, bay, baz, bwx, bwy, bwz) */
  const float* i_dT;         /* integrated time */
  const float* i_C;          /* x 225: 15x15 covariance, row-major */
  const uint8_t* i_last;     /* i == N-1: Huber(sqrt(16.92)) and information * 1e-2 (:2585-2596) */
  double lambda_init;        /* setUserLambdaInit: 1e0, or 1e-2 when bLarge */
  int32_t iterations;        /* opt_it: 10, or 4 when bLarge */
} lia_graph_view;

#ifdef __cplusplus
}
#endif
