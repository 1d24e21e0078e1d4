// b200Adapter.h -- the translation unit a bioslam maintainer adds to put libbioslam_b200 behind fastOptimize()
// (for real).  NOT compiled in this repository's image: GTSAM 4.2 / Eigen / Boost are absent there.
//
// It flattens the estimator's gtsam::NonlinearFactorGraph + gtsam::Values (built by the UNCHANGED setup(),
// reference src/lowerBodyPoseEstimator.cpp:29-505, src/imuPoseEstimator.cpp:47-315) into a b200_problem_desc, runs the LM loop on the
// GPU through the C ABI of include/bioslam_b200.h and writes the optimised values back into a gtsam::Values.
//
// Upstream patch this needs (constants of bioslam's factors are private members): in each of
//   AngularVelocityFactor, HingeJointConstraintVecErrEstAngVel, HingeJointConstraintNormErrEstAngVel, ConstrainedJointCenterPositionFactor,
//   ConstrainedJointCenterNormPositionFactor, ConstrainedJointCenterVelocityFactor, SegmentLengthMagnitudeFactor,
//   SegmentLengthMaxMagnitudeFactor, SegmentLengthMinMagnitudeFactor, SegmentLengthDiscrepancyFactor, AngleBetweenAxisAndSegmentFactor,
//   Pose3TranslationPrior, Pose3CompassPrior, MagPose3Factor            (include/factors/*.h)
// add     #ifdef BIOSLAM_B200_ADAPTER
//         friend struct b200::FactorAccess;
//         #endif
#pragma once
#include <gtsam/nonlinear/NonlinearFactorGraph.h>
#include <gtsam/nonlinear/Values.h>
#include <map>
#include <string>
#include <vector>
#include "bioslam_b200.h"

namespace b200 {

struct FlatGraph {
    b200_problem_desc desc{};
    std::vector<int32_t> timeVarType, staticVarType;
    std::vector<unsigned char> timeVarChr, staticVarChr;      // gtsam::Symbol::chr() of every time / static variable, table order
    std::vector<b200_factor_slot> slots;
    std::vector<double> consts;                               // [K][const_stride]
    std::vector<double> timeValues, staticValues;             // b200_set_values layout
    uint64_t kMin = 0;                                        // Symbol::index() of keyframe 0
};

// Walk graph / values.  A variable whose Symbol::chr() appears with more than one index is a time variable (one per keyframe),
// otherwise a static one; factors with the same type and the same key characters form one slot.  Throws std::runtime_error on a
// factor type outside include/bioslam_b200.h's enum.
FlatGraph flatten(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& values);

// values -> gtsam::Values (same keys as `like`)
gtsam::Values unflatten(const FlatGraph& fg, const gtsam::Values& like);

struct RunResult { int iterations = 0, trials = 0, status = 0; double finalError = 0, deviceMs = 0; std::vector<double> errorHistory; };

// == the while(...) optimizer.iterate() loop of lowerBodyPoseEstimator::fastOptimize (reference src/lowerBodyPoseEstimator.cpp:651-667)
RunResult optimizeLowerBody(FlatGraph& fg, double lambdaInitial, double lambdaFactor, double lambdaLower, double lambdaUpper, double absErrDecreaseLimit,
                            double relErrDecreaseLimit, double absErrLimit, int maxIterations, int convergeAtConsecSmallIter, int device = 0);

}  // namespace b200
