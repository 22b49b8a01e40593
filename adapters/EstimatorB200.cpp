// Reference-side adapter: okvis::Estimator::optimize / setOptimizationTimeLimit on top of libsvin_b200.so.
//
// Drop this file into okvis_ros/okvis/okvis_ceres/src/ IN PLACE of the two definitions at
// okvis_ceres/src/Estimator.cpp:876-929 (optimize) and :932-951 (setOptimizationTimeLimit) - e.g. wrap those two in
// `#ifndef SVIN_B200` - add it to okvis_ceres/CMakeLists.txt and link `svin_b200`.  Nothing else of the reference
// changes: ThreadedKFVio::optimizationLoop (okvis_multisensor_processing/src/ThreadedKFVio.cpp:1086) keeps calling
// estimator_.optimize(numIter, numThreads, verbose).  Needs adapters/reference_accessors.patch (eight getters on
// SonarError / DepthError / MarginalizationError, whose measurements are protected members).
//
// The function walks the okvis::ceres::Map exactly once per call:
//   parameter blocks  Map::id2parameterBlockMap()                       (okvis_ceres/include/okvis/ceres/Map.hpp:330)
//   residual blocks   Map::residualBlockId2ResidualBlockSpecMap()        (:332-334) + Map::parameters(id) (:311-312)
// and flattens them into one SvinBaWindow (include/svin_b200.h).  Not compiled in the svin_b200 repository (Eigen, Ceres,
// glog and OpenCV are not in its image); tests/c_abi_harness.c exercises the same C calls from a C program.
#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include <okvis/Estimator.hpp>
#include <okvis/cameras/PinholeCamera.hpp>
#include <okvis/cameras/RadialTangentialDistortion.hpp>
#include <okvis/ceres/DepthError.hpp>
#include <okvis/ceres/HomogeneousPointParameterBlock.hpp>
#include <okvis/ceres/ImuError.hpp>
#include <okvis/ceres/MarginalizationError.hpp>
#include <okvis/ceres/PoseError.hpp>
#include <okvis/ceres/PoseParameterBlock.hpp>
#include <okvis/ceres/RelativePoseError.hpp>
#include <okvis/ceres/ReprojectionError.hpp>
#include <okvis/ceres/SonarError.hpp>
#include <okvis/ceres/SpeedAndBiasError.hpp>
#include <okvis/ceres/SpeedAndBiasParameterBlock.hpp>

#include "svin_b200.h"

namespace okvis {
namespace {

struct B200State {             // one engine context per Estimator, created on first use
  svin_ba_ctx* ctx = nullptr;
  double timeLimit = -1.0;     // setOptimizationTimeLimit (the reference keeps these inside its CeresIterationCallback)
  int minIterations = 1;
  ~B200State() {
    if (ctx) svin_ba_destroy(ctx);
  }
};
std::mutex g_mutex;
std::unordered_map<const Estimator*, std::unique_ptr<B200State>> g_state;
B200State& stateOf(const Estimator* e) {
  std::lock_guard<std::mutex> l(g_mutex);
  std::unique_ptr<B200State>& s = g_state[e];
  if (!s) s.reset(new B200State());
  return *s;
}

template <int N>
void appendRowMajor(std::vector<double>& dst, const Eigen::Matrix<double, N, N>& M) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) dst.push_back(M(i, j));
}
void appendPose(std::vector<double>& dst, const kinematics::Transformation& T) {
  const Eigen::Vector3d r = T.r();
  const Eigen::Quaterniond q = T.q();
  dst.insert(dst.end(), {r[0], r[1], r[2], q.x(), q.y(), q.z(), q.w()});   // PoseParameterBlock layout
}

}  // namespace

void Estimator::optimize(size_t numIter, size_t /*numThreads*/, bool verbose) {
  typedef cameras::PinholeCamera<cameras::RadialTangentialDistortion> camera_t;
  B200State& st = stateOf(this);
  if (!st.ctx && svin_ba_create(0, &st.ctx) != SVIN_OK) OKVIS_THROW(Exception, svin_last_error());

  // ---------------------------------------------------------------- 1. parameter blocks (sorted ids: deterministic)
  std::vector<uint64_t> ids;
  for (const auto& kv : mapPtr_->id2parameterBlockMap()) ids.push_back(kv.first);
  std::sort(ids.begin(), ids.end());
  std::vector<double> poses, sbs, lms;
  std::vector<uint8_t> poseFixed, sbFixed, lmFixed;
  std::vector<std::shared_ptr<ceres::ParameterBlock>> poseBlk, sbBlk, lmBlk;
  std::unordered_map<uint64_t, int> poseIdx, sbIdx, lmIdx;
  for (uint64_t id : ids) {
    std::shared_ptr<ceres::ParameterBlock> pb = mapPtr_->parameterBlockPtr(id);
    const double* x = pb->parameters();
    if (std::dynamic_pointer_cast<ceres::PoseParameterBlock>(pb)) {            // T_WS and T_SCi blocks alike
      poseIdx[id] = static_cast<int>(poseBlk.size());
      poseBlk.push_back(pb);
      poseFixed.push_back(pb->fixed());
      poses.insert(poses.end(), x, x + 7);
    } else if (std::dynamic_pointer_cast<ceres::SpeedAndBiasParameterBlock>(pb)) {
      sbIdx[id] = static_cast<int>(sbBlk.size());
      sbBlk.push_back(pb);
      sbFixed.push_back(pb->fixed());
      sbs.insert(sbs.end(), x, x + 9);
    } else if (std::dynamic_pointer_cast<ceres::HomogeneousPointParameterBlock>(pb)) {
      lmIdx[id] = static_cast<int>(lmBlk.size());
      lmBlk.push_back(pb);
      lmFixed.push_back(pb->fixed());
      lms.insert(lms.end(), x, x + 4);
    } else {
      OKVIS_THROW(Exception, "svin_b200: unsupported parameter block type " << pb->typeInfo());
    }
  }

  // ---------------------------------------------------------------- 2. residual blocks
  std::vector<int32_t> obsPose, obsLm, obsExt, obsCam;
  std::vector<double> obsZ, obsInfo;
  std::vector<int32_t> imuPose0, imuSb0, imuPose1, imuSb1, imuOff(1, 0);
  std::vector<int64_t> imuT0, imuT1, imuT;
  std::vector<double> imuGyro, imuAcc;
  std::vector<int32_t> ppBlk, spBlk, rpBlk0, rpBlk1, soPose, dePose, mgKind, mgIndex;
  std::vector<double> ppMeas, ppInfo, spMeas, spInfo, rpInfo, soRange, soHeading, soInfo, soMean, deMeas, deFirst, deInfo;
  std::vector<double> mgLin, mgJ, mgE0, sonarT(7, 0.0);
  sonarT[6] = 1.0;
  SvinBaWindow w{};
  size_t maxCam = 0;
  for (const auto& kv : mapPtr_->residualBlockId2ResidualBlockSpecMap()) {
    const ::ceres::ResidualBlockId rid = kv.first;
    const std::shared_ptr<ceres::ErrorInterface>& e = kv.second.errorInterfacePtr;
    const ceres::Map::ParameterBlockCollection pars = mapPtr_->parameters(rid);
    if (auto re = std::dynamic_pointer_cast<ceres::ReprojectionErrorBase>(e)) {   // [T_WS, landmark, T_SCi]
      auto r2 = std::dynamic_pointer_cast<ceres::ReprojectionError<camera_t>>(e);
      OKVIS_ASSERT_TRUE(Exception, r2, "svin_b200: only PinholeCamera<RadialTangentialDistortion> is on the device path");
      obsPose.push_back(poseIdx.at(pars[0].first));
      obsLm.push_back(lmIdx.at(pars[1].first));
      obsExt.push_back(poseIdx.at(pars[2].first));
      obsCam.push_back(static_cast<int32_t>(re->cameraId()));
      maxCam = std::max<size_t>(maxCam, re->cameraId());
      const Eigen::Vector2d z = r2->measurement();
      const Eigen::Matrix2d I = r2->information();
      obsZ.insert(obsZ.end(), {z[0], z[1]});
      obsInfo.insert(obsInfo.end(), {I(0, 0), I(0, 1), I(1, 0), I(1, 1)});
    } else if (auto ie = std::dynamic_pointer_cast<ceres::ImuError>(e)) {          // [T_WS_0, sb_0, T_WS_1, sb_1]
      imuPose0.push_back(poseIdx.at(pars[0].first));
      imuSb0.push_back(sbIdx.at(pars[1].first));
      imuPose1.push_back(poseIdx.at(pars[2].first));
      imuSb1.push_back(sbIdx.at(pars[3].first));
      imuT0.push_back(static_cast<int64_t>(ie->t0().toNSec()));
      imuT1.push_back(static_cast<int64_t>(ie->t1().toNSec()));
      for (const ImuMeasurement& m : ie->imuMeasurements()) {
        imuT.push_back(static_cast<int64_t>(m.timeStamp.toNSec()));
        for (int k = 0; k < 3; ++k) imuGyro.push_back(m.measurement.gyroscopes[k]);
        for (int k = 0; k < 3; ++k) imuAcc.push_back(m.measurement.accelerometers[k]);
      }
      imuOff.push_back(static_cast<int32_t>(imuT.size()));
      const ImuParameters& ip = ie->imuParameters();
      w.imu_params.sigma_g_c = ip.sigma_g_c;
      w.imu_params.sigma_a_c = ip.sigma_a_c;
      w.imu_params.sigma_gw_c = ip.sigma_gw_c;
      w.imu_params.sigma_aw_c = ip.sigma_aw_c;
      w.imu_params.g = ip.g;
      w.imu_params.g_max = ip.g_max;
      w.imu_params.a_max = ip.a_max;
    } else if (auto pe = std::dynamic_pointer_cast<ceres::PoseError>(e)) {
      ppBlk.push_back(poseIdx.at(pars[0].first));
      appendPose(ppMeas, pe->measurement());
      appendRowMajor<6>(ppInfo, pe->information());
    } else if (auto se = std::dynamic_pointer_cast<ceres::SpeedAndBiasError>(e)) {
      spBlk.push_back(sbIdx.at(pars[0].first));
      const SpeedAndBias m = se->measurement();
      for (int k = 0; k < 9; ++k) spMeas.push_back(m[k]);
      appendRowMajor<9>(spInfo, se->information());
    } else if (auto rp = std::dynamic_pointer_cast<ceres::RelativePoseError>(e)) {
      rpBlk0.push_back(poseIdx.at(pars[0].first));
      rpBlk1.push_back(poseIdx.at(pars[1].first));
      appendRowMajor<6>(rpInfo, rp->information());
    } else if (auto so = std::dynamic_pointer_cast<ceres::SonarError>(e)) {       // getters: reference_accessors.patch
      soPose.push_back(poseIdx.at(pars[0].first));
      soRange.push_back(so->range());
      soHeading.push_back(so->heading());
      soInfo.push_back(so->information()(0, 0));
      Eigen::Vector3d mean = Eigen::Vector3d::Zero();                              // SonarError.cpp:125-131
      for (const Eigen::Vector3d& p : so->landmarkSubset()) mean += p;
      if (!so->landmarkSubset().empty()) mean /= static_cast<double>(so->landmarkSubset().size());
      soMean.insert(soMean.end(), {mean[0], mean[1], mean[2]});
      sonarT.clear();
      appendPose(sonarT, so->sonarParameters().T_SSo);
    } else if (auto de = std::dynamic_pointer_cast<ceres::DepthError>(e)) {
      dePose.push_back(poseIdx.at(pars[0].first));
      deMeas.push_back(de->depth());
      deFirst.push_back(de->firstDepth());
      deInfo.push_back(de->information()(0, 0));
    } else if (auto me = std::dynamic_pointer_cast<ceres::MarginalizationError>(e)) {
      OKVIS_ASSERT_TRUE(Exception, mgKind.empty(), "svin_b200: more than one MarginalizationError in the map");
      for (const auto& info : me->parameterBlockInfos()) {   // protected nested type: deduced
        const uint64_t id = info.parameterBlockId;
        if (poseIdx.count(id)) {
          mgKind.push_back(SVIN_BLOCK_POSE);
          mgIndex.push_back(poseIdx.at(id));
        } else if (sbIdx.count(id)) {
          mgKind.push_back(SVIN_BLOCK_SPEEDBIAS);
          mgIndex.push_back(sbIdx.at(id));
        } else {
          OKVIS_THROW(Exception, "svin_b200: landmark block inside the marginalisation prior");
        }
        mgLin.insert(mgLin.end(), info.linearizationPoint.get(), info.linearizationPoint.get() + info.dimension);
      }
      const Eigen::MatrixXd& J = me->J();
      const Eigen::VectorXd& e0 = me->e0();
      for (int i = 0; i < J.rows(); ++i)
        for (int j = 0; j < J.cols(); ++j) mgJ.push_back(J(i, j));
      for (int i = 0; i < e0.size(); ++i) mgE0.push_back(e0[i]);
    } else {
      OKVIS_THROW(Exception, "svin_b200: unsupported error term " << e->typeInfo());
    }
  }
  // The residual map is unordered: put the observations in (landmark, pose block, camera) order.  svin_ba_upload accepts
  // any order, but for this one it plans the window on the device (grouping, chunking, observation order:
  // csrc/ba_plan.cu) instead of on the host threads.
  {
    const size_t n = obsPose.size();
    std::vector<size_t> ord(n);
    for (size_t i = 0; i < n; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
      if (obsLm[a] != obsLm[b]) return obsLm[a] < obsLm[b];
      if (obsPose[a] != obsPose[b]) return obsPose[a] < obsPose[b];
      return obsCam[a] < obsCam[b];
    });
    std::vector<int32_t> p2(n), l2(n), e2(n), c2(n);
    std::vector<double> z2(2 * n), i2(4 * n);
    for (size_t k = 0; k < n; ++k) {
      const size_t s = ord[k];
      p2[k] = obsPose[s];
      l2[k] = obsLm[s];
      e2[k] = obsExt[s];
      c2[k] = obsCam[s];
      for (int j = 0; j < 2; ++j) z2[2 * k + j] = obsZ[2 * s + j];
      for (int j = 0; j < 4; ++j) i2[4 * k + j] = obsInfo[4 * s + j];
    }
    obsPose.swap(p2);
    obsLm.swap(l2);
    obsExt.swap(e2);
    obsCam.swap(c2);
    obsZ.swap(z2);
    obsInfo.swap(i2);
  }
  // intrinsics of every camera that carries observations: [fu fv cu cv k1 k2 p1 p2]
  std::vector<double> intrinsics;
  if (!obsCam.empty()) {
    OKVIS_ASSERT_TRUE(Exception, !multiFramePtrMap_.empty(), "no multiframe to take the camera geometry from");
    const std::shared_ptr<MultiFrame> mf = multiFramePtrMap_.rbegin()->second;
    for (size_t c = 0; c <= maxCam; ++c) {
      Eigen::VectorXd v;
      mf->geometryAs<camera_t>(c)->getIntrinsics(v);
      OKVIS_ASSERT_TRUE(Exception, v.size() == 8, "unexpected intrinsics vector");
      for (int k = 0; k < 8; ++k) intrinsics.push_back(v[k]);
    }
  }

  // ---------------------------------------------------------------- 3. the window struct (host pointers only)
  w.num_pose_blocks = static_cast<int32_t>(poseBlk.size());
  w.num_speedbias = static_cast<int32_t>(sbBlk.size());
  w.num_landmarks = static_cast<int32_t>(lmBlk.size());
  w.num_cameras = static_cast<int32_t>(intrinsics.size() / 8);
  w.pose_blocks = poses.data();
  w.speedbias = sbs.data();
  w.landmarks = lms.data();
  w.pose_fixed = poseFixed.data();
  w.speedbias_fixed = sbFixed.data();
  w.landmark_fixed = lmFixed.data();
  w.intrinsics = intrinsics.data();
  w.num_obs = static_cast<int32_t>(obsPose.size());
  w.loss_type = SVIN_LOSS_CAUCHY;                       // cauchyLossFunctionPtr_(new CauchyLoss(1)), Estimator.cpp:61
  w.loss_scale = 1.0;
  w.obs_pose = obsPose.data();
  w.obs_landmark = obsLm.data();
  w.obs_extrinsics = obsExt.data();
  w.obs_camera = obsCam.data();
  w.obs_measurement = obsZ.data();
  w.obs_information = obsInfo.data();
  w.num_imu = static_cast<int32_t>(imuPose0.size());
  w.imu_pose0 = imuPose0.data();
  w.imu_speedbias0 = imuSb0.data();
  w.imu_pose1 = imuPose1.data();
  w.imu_speedbias1 = imuSb1.data();
  w.imu_t0_ns = imuT0.data();
  w.imu_t1_ns = imuT1.data();
  w.imu_meas_offset = imuOff.data();
  w.imu_meas_t_ns = imuT.data();
  w.imu_meas_gyro = imuGyro.data();
  w.imu_meas_accel = imuAcc.data();
  w.num_pose_priors = static_cast<int32_t>(ppBlk.size());
  w.pose_prior_block = ppBlk.data();
  w.pose_prior_measurement = ppMeas.data();
  w.pose_prior_information = ppInfo.data();
  w.num_speedbias_priors = static_cast<int32_t>(spBlk.size());
  w.speedbias_prior_block = spBlk.data();
  w.speedbias_prior_measurement = spMeas.data();
  w.speedbias_prior_information = spInfo.data();
  w.num_relative_pose = static_cast<int32_t>(rpBlk0.size());
  w.relative_pose_block0 = rpBlk0.data();
  w.relative_pose_block1 = rpBlk1.data();
  w.relative_pose_information = rpInfo.data();
  w.num_sonar = static_cast<int32_t>(soPose.size());
  w.sonar_pose = soPose.data();
  w.sonar_range = soRange.data();
  w.sonar_heading = soHeading.data();
  w.sonar_information = soInfo.data();
  w.sonar_landmark_mean = soMean.data();
  w.sonar_T_SSo = sonarT.data();
  w.num_depth = static_cast<int32_t>(dePose.size());
  w.depth_pose = dePose.data();
  w.depth_measurement = deMeas.data();
  w.depth_first = deFirst.data();
  w.depth_information = deInfo.data();
  w.marg_num_blocks = static_cast<int32_t>(mgKind.size());
  w.marg_dim = static_cast<int32_t>(mgE0.size());
  w.marg_block_kind = mgKind.data();
  w.marg_block_index = mgIndex.data();
  w.marg_linearization_points = mgLin.data();
  w.marg_J = mgJ.data();
  w.marg_e0 = mgE0.data();

  // ---------------------------------------------------------------- 4. solve (Ceres defaults + Estimator.cpp:878-890)
  SvinBaOptions opt;
  svin_ba_default_options(&opt);
  opt.max_num_iterations = static_cast<int32_t>(numIter);
  opt.min_num_iterations = st.minIterations;
  opt.time_limit_seconds = st.timeLimit;                // < 0: no limit (CeresIterationCallback.hpp:55-80)
  opt.compute_landmark_quality = 1;
  SvinBaSummary summary;
  std::vector<double> quality(lmBlk.size(), 0.0);
  double* qptr = quality.data();
  if (svin_ba_optimize(st.ctx, &w, 1, &opt, &summary, &qptr) != SVIN_OK) OKVIS_THROW(Exception, svin_last_error());

  // ---------------------------------------------------------------- 5. write back
  for (size_t k = 0; k < poseBlk.size(); ++k)
    if (!poseFixed[k]) poseBlk[k]->setParameters(&poses[7 * k]);
  for (size_t k = 0; k < sbBlk.size(); ++k)
    if (!sbFixed[k]) sbBlk[k]->setParameters(&sbs[9 * k]);
  for (size_t k = 0; k < lmBlk.size(); ++k)
    if (!lmFixed[k]) lmBlk[k]->setParameters(&lms[4 * k]);
  for (auto it = landmarksMap_.begin(); it != landmarksMap_.end(); ++it) {      // Estimator.cpp:903-922
    const auto f = lmIdx.find(it->first);
    if (f == lmIdx.end()) continue;
    it->second.quality = quality[f->second];
    it->second.point = Eigen::Vector4d(lms[4 * f->second], lms[4 * f->second + 1], lms[4 * f->second + 2],
                                       lms[4 * f->second + 3]);
  }
  if (verbose)
    LOG(INFO) << "svin_b200: " << summary.iterations << " iterations, cost " << summary.initial_cost << " -> "
              << summary.final_cost << ", termination " << summary.termination;
}

// Estimator.cpp:932-951.  The engine applies the limit per svin_ba_solve call on the device clock; with one window per
// call (this adapter) that is the reference's per-optimize() budget.
bool Estimator::setOptimizationTimeLimit(double timeLimit, int minIterations) {
  B200State& st = stateOf(this);
  st.timeLimit = timeLimit;
  st.minIterations = timeLimit < 0.0 ? static_cast<int>(mapPtr_->options.max_num_iterations) : minIterations;
  return true;
}

}  // namespace okvis
