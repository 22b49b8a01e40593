// Reference-side adapters for hot path (A): detector/extractor, dense matcher and RANSAC on top of libsvin_b200.so.
// Header-only; include from okvis_frontend/src/Frontend.cpp.  Not compiled in the svin_b200 repository (OpenCV, Eigen, brisk
// and OpenGV are not in its image).  Reference seams:
//   B200Brisk          cv::FeatureDetector / cv::DescriptorExtractor installed at Frontend.cpp:101-102, called by
//                      Frame::detect / Frame::describe (okvis_cv/include/okvis/implementation/Frame.hpp:93-135)
//   matchOnB200        matcher_->match<MATCHING_ALGORITHM>(matchingAlgorithm) in matchToKeyframes / matchToLastFrame /
//                      matchStereo (Frontend.cpp:336-614) = DenseMatcher::match (DenseMatcher.hpp impl:195-203)
//   ransac3d2dOnB200   the opengv::sac::Ransac block of Frontend::runRansac3d2d (Frontend.cpp:632-648)
//   ransac2d2dOnB200   the two Ransac blocks of runRansac2d2d (:860-895) and runRansac2d2dToRefineScale (:706-735)
#pragma once
#include <cstdlib>
#include <memory>
#include <vector>

#include <opencv2/features2d/features2d.hpp>

#include <okvis/Estimator.hpp>
#include <okvis/VioKeyframeWindowMatchingAlgorithm.hpp>
#include <opengv/absolute_pose/FrameNoncentralAbsoluteAdapter.hpp>
#include <opengv/relative_pose/FrameRelativeAdapter.hpp>

#include "svin_b200.h"

namespace okvis {
namespace b200 {

// ------------------------------------------------------------------------------------------------ detect + describe
// One object serves as detector AND extractor of a camera (Frame::setDetector / setExtractor): detect() runs Harris +
// uniformity + orientation + description on the device and keeps the descriptors for the compute() that follows.
class B200Brisk : public cv::Feature2D {
 public:
  B200Brisk(int width, int height, const Eigen::VectorXd& intrinsics /*fu fv cu cv k1 k2 p1 p2*/, double uniformityRadius,
            double absoluteThreshold, int maxNoKeypoints) {
    SvinFeOptions o;
    svin_fe_default_options(&o);
    o.image_width = width;
    o.image_height = height;
    o.max_images = 1;
    o.detection_threshold = uniformityRadius;     // detection_options.threshold (config_fpga_p2_euroc.yaml:66)
    o.absolute_threshold = absoluteThreshold;     // 800 (Frontend.cpp:75)
    o.max_keypoints = maxNoKeypoints;             // detection_options.maxNoKeypoints (:68)
    maxKp_ = maxNoKeypoints;
    if (svin_fe_create(0, &o, &ctx_) != SVIN_OK) OKVIS_THROW(std::runtime_error, svin_last_error());
    for (int k = 0; k < 8; ++k) intr_[k] = intrinsics[k];
  }
  ~B200Brisk() override { svin_fe_destroy(ctx_); }
  // extractionDirection = T_WC^-1.C() * (0, 0, -1), Frontend.cpp:107-108; set before detect()
  void setExtractionDirection(const Eigen::Vector3d& g) {
    for (int k = 0; k < 3; ++k) g_[k] = g[k];
  }
  void detect(cv::InputArray image, std::vector<cv::KeyPoint>& keypoints, cv::InputArray /*mask*/) override {
    static_assert(sizeof(cv::KeyPoint) == sizeof(SvinKeypoint), "cv::KeyPoint and SvinKeypoint share the 28-byte layout");
    const cv::Mat img = image.getMat();
    const uint8_t* p = img.data;
    int32_t n = 0;
    keypoints.resize(maxKp_);
    desc_.create(maxKp_, 48, CV_8U);
    if (svin_fe_detect_describe(ctx_, 1, &p, static_cast<int32_t>(img.step), intr_, g_,
                                reinterpret_cast<SvinKeypoint*>(keypoints.data()), desc_.data, &n) != SVIN_OK)
      OKVIS_THROW(std::runtime_error, svin_last_error());
    keypoints.resize(n);
    desc_ = desc_.rowRange(0, n);
  }
  // Frame::describe first overwrites kp.angle with the same gravity-aligned angle (Frame.hpp impl:113-129) and then calls
  // this; the descriptors were extracted with exactly that angle, so they are returned as they are.
  void compute(cv::InputArray /*image*/, std::vector<cv::KeyPoint>& /*keypoints*/, cv::OutputArray descriptors) override {
    desc_.copyTo(descriptors);
  }

 private:
  svin_fe_ctx* ctx_ = nullptr;
  int maxKp_ = 400;
  double intr_[8], g_[3] = {0, 0, -1};
  cv::Mat desc_;
};

// ------------------------------------------------------------------------------------------------ dense matching
// Replaces matcher_->match<ALG>(alg) for ALG = VioKeyframeWindowMatchingAlgorithm<CAM>.  The graph-state part of doSetup
// (which landmarks exist / are initialised / are already observed) stays here on the host - it needs the Estimator - and
// is handed over as skip masks; projections, covariances, Hamming distances, verifyMatch and the best-4 assignment run on
// the device; setBestMatch (graph mutation, VKWMA.cpp:352-498) is called here in B-index order like matchBody.
template <class CAM>
void matchOnB200(svin_fe_ctx* fe, Estimator& estimator, VioKeyframeWindowMatchingAlgorithm<CAM>& alg, uint64_t mfIdA,
                 uint64_t mfIdB, size_t camIdA, size_t camIdB, bool match3d2d) {
  alg.doSetup();   // keeps the algorithm's own state (triangulator, counters) valid for setBestMatch
  std::shared_ptr<MultiFrame> fA = estimator.multiFrame(mfIdA), fB = estimator.multiFrame(mfIdB);
  const int nA = static_cast<int>(fA->numKeypoints(camIdA)), nB = static_cast<int>(fB->numKeypoints(camIdB));
  if (nA == 0 || nB == 0) return;
  std::vector<uint8_t> skipA(nA, 0), skipB(nB, 0);
  std::vector<double> lmA(4 * static_cast<size_t>(nA), 0.0);
  std::vector<SvinKeypoint> kpA(nA), kpB(nB);
  auto fill = [](const std::shared_ptr<MultiFrame>& f, size_t cam, std::vector<SvinKeypoint>& out) {
    for (size_t k = 0; k < out.size(); ++k) {
      const cv::KeyPoint* kp = f->keypoint(cam, k);
      out[k] = SvinKeypoint{kp->pt.x, kp->pt.y, kp->size, kp->angle, kp->response, kp->octave, kp->class_id};
    }
  };
  fill(fA, camIdA, kpA);
  fill(fB, camIdB, kpB);
  for (int k = 0; k < nA; ++k) {                          // VKWMA.cpp:160-185 (3D-2D) / :208-221 (2D-2D)
    const uint64_t id = fA->landmarkId(camIdA, k);
    if (match3d2d) {
      if (id == 0 || !estimator.isLandmarkAdded(id) || !estimator.isLandmarkInitialized(id)) {
        skipA[k] = 1;
        continue;
      }
      MapPoint lm;
      estimator.getLandmark(id, lm);
      if (lm.observations.size() < 2) {
        estimator.setLandmarkInitialized(id, false);
        skipA[k] = 1;
        continue;
      }
      for (int c = 0; c < 4; ++c) lmA[4 * k + c] = lm.point[c];
    } else if (id != 0 && estimator.isLandmarkAdded(id) && estimator.isLandmarkInitialized(id)) {
      skipA[k] = 1;
    }
  }
  for (int k = 0; k < nB; ++k) {                          // VKWMA.cpp:228-262
    const uint64_t id = fB->landmarkId(camIdB, k);
    if (id == 0 || !estimator.isLandmarkAdded(id)) continue;
    if (match3d2d) {
      MapPoint lm;
      estimator.getLandmark(id, lm);
      skipB[k] = lm.observations.find(KeypointIdentifier(mfIdB, camIdB, k)) != lm.observations.end();
    } else {
      skipB[k] = estimator.isLandmarkInitialized(id);
    }
  }
  kinematics::Transformation T_WSa, T_WSb, T_SCa, T_SCb;
  estimator.get_T_WS(mfIdA, T_WSa);
  estimator.get_T_WS(mfIdB, T_WSb);
  estimator.getCameraSensorStates(mfIdA, camIdA, T_SCa);
  estimator.getCameraSensorStates(mfIdB, camIdB, T_SCb);
  const kinematics::Transformation T_WCa = T_WSa * T_SCa, T_WCb = T_WSb * T_SCb;
  const kinematics::Transformation T_CbW = T_WCb.inverse(), T_CaCb = T_WCa.inverse() * T_WCb;
  auto pose7 = [](const kinematics::Transformation& T, double* o) {
    const Eigen::Vector3d r = T.r();
    const Eigen::Quaterniond q = T.q();
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = q.x(); o[4] = q.y(); o[5] = q.z(); o[6] = q.w();
  };
  double t_cbw[7], t_cacb[7], intrA[8], intrB[8];
  pose7(T_CbW, t_cbw);
  pose7(T_CaCb, t_cacb);
  Eigen::VectorXd v;
  fA->template geometryAs<CAM>(camIdA)->getIntrinsics(v);
  for (int k = 0; k < 8; ++k) intrA[k] = v[k];
  fB->template geometryAs<CAM>(camIdB)->getIntrinsics(v);
  for (int k = 0; k < 8; ++k) intrB[k] = v[k];
  // UOplus translation variance, VKWMA.cpp:132-144
  double poseUnc = 4e-8;
  const uint64_t cur = estimator.currentFrameId();
  if (estimator.isInImuWindow(cur) && mfIdA != mfIdB) {
    SpeedAndBias sb;
    estimator.getSpeedAndBias(cur, 0, sb);
    const double scale = std::max(1.0, sb.head<3>().norm());
    poseUnc = scale * scale * 1.0e-2;
  }
  SvinMatchProblem p{};
  p.type = match3d2d ? SVIN_MATCH_3D2D : SVIN_MATCH_2D2D;
  p.nA = nA;
  p.nB = nB;
  p.descA = fA->keypointDescriptor(camIdA, 0);           // N x 48 contiguous (Frame.hpp impl:212-221)
  p.descB = fB->keypointDescriptor(camIdB, 0);
  p.skipA = skipA.data();
  p.skipB = skipB.data();
  p.kpA = kpA.data();
  p.kpB = kpB.data();
  p.distance_threshold = alg.distanceThreshold();        // 60, Frontend.cpp:79
  p.landmarksA = lmA.data();
  p.T_CbW = t_cbw;
  p.pose_uncertainty = poseUnc;
  p.intrA = intrA;
  p.intrB = intrB;
  p.T_CaCb = t_cacb;
  p.image_width = static_cast<int32_t>(fB->template geometryAs<CAM>(camIdB)->imageWidth());
  p.image_height = static_cast<int32_t>(fB->template geometryAs<CAM>(camIdB)->imageHeight());
  std::vector<int32_t> matchOfB(nB);
  std::vector<float> dist(nB);
  SvinMatchResult r{nullptr, nullptr, matchOfB.data(), dist.data(), nullptr};
  if (svin_match(fe, 1, &p, &r) != SVIN_OK) OKVIS_THROW(std::runtime_error, svin_last_error());
  alg.reserveMatches(nB);
  for (int b = 0; b < nB; ++b)                            // matchBody tail, DenseMatcher.hpp impl:95-119
    if (matchOfB[b] >= 0) alg.setBestMatch(matchOfB[b], b, dist[b]);
}

// ------------------------------------------------------------------------------------------------ RANSAC
// Sample index sets drawn the way OpenGV's SampleConsensusProblem::getSamples does (rand() without repetition); for the
// absolute problem the first three indices come from one camera (svin_ransac_absolute solves central P3P per camera).
inline std::vector<int32_t> drawSamples(int n, int size, int count) {
  std::vector<int32_t> s(static_cast<size_t>(size) * count);
  for (int j = 0; j < count; ++j)
    for (int k = 0; k < size; ++k) {
      int32_t v;
      bool fresh;
      do {
        v = std::rand() % n;
        fresh = true;
        for (int q = 0; q < k; ++q) fresh = fresh && s[static_cast<size_t>(j) * size + q] != v;
      } while (!fresh);
      s[static_cast<size_t>(j) * size + k] = v;
    }
  return s;
}

// -> inlier mask over the adapter's correspondences, like ransac.inliers_ (empty when no model was found)
inline std::vector<bool> ransac3d2dOnB200(svin_ransac_ctx* ctx, opengv::absolute_pose::FrameNoncentralAbsoluteAdapter& adapter,
                                          Eigen::Matrix<double, 3, 4>* model = nullptr) {
  const int n = static_cast<int>(adapter.getNumberCorrespondences());
  std::vector<double> pts(3 * n), brs(3 * n), sig(n);
  std::vector<int32_t> cam(n);
  int numCams = 0;
  for (int k = 0; k < n; ++k) {
    const opengv::point_t p = adapter.getPoint(k);
    const opengv::bearingVector_t f = adapter.getBearingVector(k);
    for (int c = 0; c < 3; ++c) {
      pts[3 * k + c] = p[c];
      brs[3 * k + c] = f[c];
    }
    sig[k] = adapter.getSigmaAngle(k);
    cam[k] = static_cast<int32_t>(adapter.camIndex(k));
    numCams = std::max(numCams, cam[k] + 1);
  }
  std::vector<double> camR(9 * numCams), camT(3 * numCams);
  std::vector<std::vector<int32_t>> perCam(numCams);
  for (int k = 0; k < n; ++k) {
    perCam[cam[k]].push_back(k);
    const opengv::rotation_t R = adapter.getCamRotation(k);
    const opengv::translation_t t = adapter.getCamOffset(k);
    for (int i = 0; i < 3; ++i) {
      camT[3 * cam[k] + i] = t[i];
      for (int j = 0; j < 3; ++j) camR[9 * cam[k] + 3 * i + j] = R(i, j);
    }
  }
  const int numSamples = 64;                              // >= max_iterations + 1 plus room for skipped models
  std::vector<int32_t> samples;
  for (int j = 0; j < numSamples; ++j) {
    int c = std::rand() % numCams;
    for (int tries = 0; tries < numCams && perCam[c].size() < 3; ++tries) c = (c + 1) % numCams;
    if (perCam[c].size() < 3) break;
    const std::vector<int32_t> three = drawSamples(static_cast<int>(perCam[c].size()), 3, 1);
    int32_t fourth;
    do {
      fourth = std::rand() % n;
    } while (fourth == perCam[c][three[0]] || fourth == perCam[c][three[1]] || fourth == perCam[c][three[2]]);
    samples.insert(samples.end(), {perCam[c][three[0]], perCam[c][three[1]], perCam[c][three[2]], fourth});
  }
  SvinRansacAbsProblem prob{};
  prob.num_correspondences = n;
  prob.points = pts.data();
  prob.bearings = brs.data();
  prob.camera_index = cam.data();
  prob.sigma_angle = sig.data();
  prob.num_cameras = numCams;
  prob.camera_rotation = camR.data();
  prob.camera_offset = camT.data();
  prob.num_samples = static_cast<int32_t>(samples.size() / 4);
  prob.samples = samples.data();
  prob.threshold = 9;                                     // Frontend.cpp:643
  prob.max_iterations = 50;                               // :644
  std::vector<uint8_t> mask(n, 0);
  SvinRansacResult res{};
  res.inliers = mask.data();
  if (svin_ransac_absolute(ctx, 1, &prob, &res) != SVIN_OK) OKVIS_THROW(std::runtime_error, svin_last_error());
  if (res.best_sample < 0) return std::vector<bool>();
  if (model)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) (*model)(i, j) = res.model[4 * i + j];
  return std::vector<bool>(mask.begin(), mask.end());
}

struct Ransac2d2dResult {
  std::vector<bool> rotationOnlyInliers, relPoseInliers;
  int rotationOnlyCount = 0, relPoseCount = 0;
  Eigen::Matrix3d rotationOnlyModel = Eigen::Matrix3d::Identity();           // rotation_only_ransac.model_coefficients_
  Eigen::Matrix<double, 3, 4> relPoseModel = Eigen::Matrix<double, 3, 4>::Zero();   // rel_pose_ransac.model_coefficients_
};
// Both RANSACs of runRansac2d2d / runRansac2d2dToRefineScale; the ratio rule (Frontend.cpp:876-905) stays in the caller.
inline Ransac2d2dResult ransac2d2dOnB200(svin_ransac_ctx* ctx, opengv::relative_pose::FrameRelativeAdapter& adapter) {
  const int n = static_cast<int>(adapter.getNumberCorrespondences());
  std::vector<double> f1(3 * n), f2(3 * n), s1(n), s2(n);
  for (int k = 0; k < n; ++k) {
    const opengv::bearingVector_t a = adapter.getBearingVector1(k), b = adapter.getBearingVector2(k);
    for (int c = 0; c < 3; ++c) {
      f1[3 * k + c] = a[c];
      f2[3 * k + c] = b[c];
    }
    s1[k] = adapter.getSigmaAngle1(k);
    s2[k] = adapter.getSigmaAngle2(k);
  }
  const int numSamples = 64;
  const std::vector<int32_t> smpRot = drawSamples(n, 2, numSamples), smpRel = drawSamples(n, 8, numSamples);
  SvinRansacRelProblem prob{};
  prob.num_correspondences = n;
  prob.bearings1 = f1.data();
  prob.bearings2 = f2.data();
  prob.sigma_angle1 = s1.data();
  prob.sigma_angle2 = s2.data();
  prob.num_samples = numSamples;
  prob.samples_rotation = smpRot.data();
  prob.samples_relative = smpRel.data();
  prob.threshold = 9;                                     // Frontend.cpp:866,881
  prob.max_iterations = 50;
  std::vector<uint8_t> mRot(n, 0), mRel(n, 0);
  SvinRansacResult rRot{}, rRel{};
  rRot.inliers = mRot.data();
  rRel.inliers = mRel.data();
  if (svin_ransac_relative(ctx, 1, &prob, &rRot, &rRel) != SVIN_OK) OKVIS_THROW(std::runtime_error, svin_last_error());
  Ransac2d2dResult out;
  out.rotationOnlyInliers.assign(mRot.begin(), mRot.end());
  out.relPoseInliers.assign(mRel.begin(), mRel.end());
  out.rotationOnlyCount = rRot.num_inliers;
  out.relPoseCount = rRel.num_inliers;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      if (j < 3) out.rotationOnlyModel(i, j) = rRot.model[4 * i + j];
      out.relPoseModel(i, j) = rRel.model[4 * i + j];
    }
  return out;
}

}  // namespace b200
}  // namespace okvis
