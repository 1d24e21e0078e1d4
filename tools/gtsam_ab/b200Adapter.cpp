// See b200Adapter.h.  NOT compiled in this repository's image (GTSAM 4.2 absent): written against the GTSAM 4.2 and bioslam 1.2 headers.
#include "b200Adapter.h"
#include <gtsam/base/Vector.h>
#include <gtsam/geometry/Pose3.h>
#include <gtsam/geometry/Unit3.h>
#include <gtsam/inference/Symbol.h>
#include <gtsam/navigation/CombinedImuFactor.h>
#include <gtsam/navigation/ImuFactor.h>
#include <gtsam/nonlinear/PriorFactor.h>
#include <Eigen/Cholesky>
#include <set>
#include <stdexcept>
#include "factors/AngleBetweenAxisAndSegmentFactor.h"
#include "factors/AngularVelocityFactor.h"
#include "factors/ConstrainedJointCenterPositionFactor.h"
#include "factors/ConstrainedJointCenterVelocityFactor.h"
#include "factors/HingeJointFactors.h"
#include "factors/MagPose3Factor.h"
#include "factors/Pose3Priors.h"
#include "factors/SegmentLengthDiscrepancyFactor.h"
#include "factors/SegmentLengthMagnitudeFactor.h"

namespace b200 {
using gtsam::Symbol;

// friend of bioslam's factor classes (see the header): read-only access to their constants
struct FactorAccess {
    static gtsam::Vector3 measured(const bioslam::AngularVelocityFactor& f) { return f.m_measuredAngVel; }
    static double ideal(const bioslam::SegmentLengthMagnitudeFactor& f) { return f.m_idealSegmentLength; }
    static double maxLen(const bioslam::SegmentLengthMaxMagnitudeFactor& f) { return f.m_maxSegmentLength; }
    static double minLen(const bioslam::SegmentLengthMinMagnitudeFactor& f) { return f.m_minSegmentLength; }
    static gtsam::Vector3 prior(const bioslam::Pose3TranslationPrior& f) { return f.m_priorTranslation; }
    static void compass(const bioslam::Pose3CompassPrior& f, double* p) {   // params[1..10]
        p[1] = f.m_expectedAng;
        for (int i = 0; i < 3; i++) { p[2 + i] = f.m_refVecLocal[i]; p[5 + i] = f.m_refVecNav[i]; p[8 + i] = f.m_upVecNav[i]; }
    }
    static void mag(const bioslam::MagPose3Factor& f, double* params3to5, double* consts3) {
        for (int i = 0; i < 3; i++) { params3to5[i] = f.B_Global[i]; consts3[i] = f.m_measurement[i]; }
    }
    // expected distance / angle of the norm-type and angle factors: same pattern (m_expectedDist, m_ang ...)
};

static void sigmasOf(const gtsam::NoiseModelFactor& f, double* out, int n) {
    auto diag = boost::dynamic_pointer_cast<gtsam::noiseModel::Diagonal>(f.noiseModel());
    if (!diag) throw std::runtime_error("b200Adapter: factor with a non-diagonal noise model");
    for (int i = 0; i < n; i++) out[i] = diag->sigmas()[i];
}

static int varTypeOf(const gtsam::Values& v, gtsam::Key k) {
    if (v.exists<gtsam::Pose3>(k)) return B200_VAR_POSE3;
    if (v.exists<gtsam::imuBias::ConstantBias>(k)) return B200_VAR_BIAS6;
    if (v.exists<gtsam::Unit3>(k)) return B200_VAR_UNIT3;
    return B200_VAR_VEC3;   // gtsam::Vector3 / gtsam::Point3
}

static void packValue(const gtsam::Values& v, gtsam::Key k, int type, double* out) {
    switch (type) {
    case B200_VAR_POSE3: {
        const gtsam::Pose3 T = v.at<gtsam::Pose3>(k); const gtsam::Matrix3 R = T.rotation().matrix();
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) out[3 * r + c] = R(r, c);       // row-major, body -> nav
        for (int i = 0; i < 3; i++) out[9 + i] = T.translation()[i];
        break; }
    case B200_VAR_BIAS6: { const auto b = v.at<gtsam::imuBias::ConstantBias>(k).vector(); for (int i = 0; i < 6; i++) out[i] = b[i]; break; }   // [acc, gyro]
    case B200_VAR_UNIT3: { const auto u = v.at<gtsam::Unit3>(k).unitVector(); for (int i = 0; i < 3; i++) out[i] = u[i]; break; }
    default: { const gtsam::Vector3 x = v.at<gtsam::Vector3>(k); for (int i = 0; i < 3; i++) out[i] = x[i]; }
    }
}
static const int kValDim[4] = {12, 3, 6, 3};

// per-keyframe constant record of a CombinedImuFactor: bit-identical to what GTSAM integrated in setup()
static void packCombinedImu(const gtsam::CombinedImuFactor& f, double* rec /*190*/) {
    const auto& pim = f.preintegratedMeasurements();
    const gtsam::Vector9 z = pim.preintegrated();
    for (int i = 0; i < 9; i++) rec[i] = z[i];
    const gtsam::Matrix93 Ha = pim.preintegrated_H_biasAcc(), Hg = pim.preintegrated_H_biasOmega();
    for (int r = 0; r < 9; r++) for (int c = 0; c < 3; c++) { rec[9 + 3 * r + c] = Ha(r, c); rec[36 + 3 * r + c] = Hg(r, c); }
    const auto bh = pim.biasHat().vector();
    for (int i = 0; i < 6; i++) rec[63 + i] = bh[i];
    rec[69] = pim.deltaTij();
    // sqrt information: upper-triangular R with R^T R = preintMeasCov^-1 (what noiseModel::Gaussian::Covariance whitens with)
    const Eigen::Matrix<double, 15, 15> info = pim.preintMeasCov().inverse();
    const Eigen::Matrix<double, 15, 15> R = info.llt().matrixU();
    int off = 70;
    for (int r = 0; r < 15; r++) for (int c = r; c < 15; c++) rec[off++] = R(r, c);
}

FlatGraph flatten(const gtsam::NonlinearFactorGraph& graph, const gtsam::Values& values) {
    FlatGraph fg;
    // ---- variable tables: characters seen with several indices are per-keyframe variables
    std::map<unsigned char, std::set<uint64_t>> idx;
    for (const gtsam::Key k : values.keys()) idx[Symbol(k).chr()].insert(Symbol(k).index());
    uint64_t kMin = UINT64_MAX, kMax = 0;
    std::map<unsigned char, int> timeIdx, staticIdx;
    for (const auto& [chr, s] : idx) {
        const gtsam::Key any = Symbol(chr, *s.begin());
        if (s.size() > 1) {
            timeIdx[chr] = (int)fg.timeVarType.size(); fg.timeVarType.push_back(varTypeOf(values, any)); fg.timeVarChr.push_back(chr);
            kMin = std::min(kMin, *s.begin()); kMax = std::max(kMax, *s.rbegin());
        } else { staticIdx[chr] = (int)fg.staticVarType.size(); fg.staticVarType.push_back(varTypeOf(values, any)); fg.staticVarChr.push_back(chr); }
    }
    const int K = (int)(kMax - kMin + 1);
    fg.kMin = kMin;
    // ---- slots: (factor type, key characters) -> slot; per-keyframe constants appended to the keyframe record
    struct SlotKey { int type; std::vector<unsigned char> chrs; bool operator<(const SlotKey& o) const { return std::tie(type, chrs) < std::tie(o.type, o.chrs); } };
    std::map<SlotKey, int> slotOf;
    std::vector<int> constOffset;   // per slot
    int constStride = 0;
    auto slotFor = [&](int type, const gtsam::NonlinearFactor& f, int nconst, const double* params, int nparams) -> b200_factor_slot& {
        SlotKey sk{type, {}};
        for (const gtsam::Key k : f.keys()) sk.chrs.push_back(Symbol(k).chr());
        auto it = slotOf.find(sk);
        if (it == slotOf.end()) {
            b200_factor_slot s{};
            s.type = type; s.nvars = (int)f.keys().size(); s.k_begin = K; s.k_end = 0;
            s.const_offset = nconst ? constStride : -1; constStride += nconst;
            for (int i = 0; i < nparams; i++) s.params[i] = params[i];
            it = slotOf.emplace(sk, (int)fg.slots.size()).first;
            fg.slots.push_back(s);
        }
        return fg.slots[it->second];
    };
    // two passes: the constant stride must be known before the records are written
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) fg.consts.assign((size_t)K * std::max(constStride, 1), 0.0);
        for (const auto& fp : graph) {
            const auto* nf = dynamic_cast<const gtsam::NoiseModelFactor*>(fp.get());
            if (!nf) throw std::runtime_error("b200Adapter: factor without a noise model");
            double prm[B200_MAX_FACTOR_PARAMS] = {0};
            double rec[190];
            int type = -1, nconst = 0, nparams = 0;
            if (auto f = dynamic_cast<const gtsam::CombinedImuFactor*>(nf)) {
                type = B200_F_IMU_COMBINED; nconst = 190; nparams = 3;
                const gtsam::Vector3 g = f->preintegratedMeasurements().p().n_gravity; for (int i = 0; i < 3; i++) prm[i] = g[i];
                if (pass == 1) packCombinedImu(*f, rec);
            } else if (auto f = dynamic_cast<const bioslam::AngularVelocityFactor*>(nf)) {
                type = B200_F_ANGVEL; nconst = 3; nparams = 3; sigmasOf(*nf, prm, 3);
                const gtsam::Vector3 m = FactorAccess::measured(*f); for (int i = 0; i < 3; i++) rec[i] = m[i];
            } else if (dynamic_cast<const bioslam::HingeJointConstraintVecErrEstAngVel*>(nf)) { type = B200_F_HINGE_VEC; nparams = 3; sigmasOf(*nf, prm, 3);
            } else if (dynamic_cast<const bioslam::HingeJointConstraintNormErrEstAngVel*>(nf)) { type = B200_F_HINGE_NORM; nparams = 1; sigmasOf(*nf, prm, 1);
            } else if (dynamic_cast<const bioslam::ConstrainedJointCenterPositionFactor*>(nf)) { type = B200_F_JOINTPOS_VEC; nparams = 3; sigmasOf(*nf, prm, 3);
            } else if (dynamic_cast<const bioslam::ConstrainedJointCenterVelocityFactor*>(nf)) { type = B200_F_JOINTVEL_VEC; nparams = 3; sigmasOf(*nf, prm, 3);
            } else if (auto f = dynamic_cast<const bioslam::SegmentLengthMagnitudeFactor*>(nf)) { type = B200_F_SEGLEN_MAG; nparams = 2; sigmasOf(*nf, prm, 1); prm[1] = FactorAccess::ideal(*f);
            } else if (auto f = dynamic_cast<const bioslam::SegmentLengthMaxMagnitudeFactor*>(nf)) { type = B200_F_SEGLEN_MAX; nparams = 3; sigmasOf(*nf, prm, 1); prm[1] = FactorAccess::maxLen(*f); prm[2] = 4.000000000000001;
            } else if (auto f = dynamic_cast<const bioslam::SegmentLengthMinMagnitudeFactor*>(nf)) { type = B200_F_SEGLEN_MIN; nparams = 3; sigmasOf(*nf, prm, 1); prm[1] = FactorAccess::minLen(*f); prm[2] = 4.000000000000001;
            } else if (dynamic_cast<const bioslam::SegmentLengthDiscrepancyFactor*>(nf)) { type = B200_F_SEGLEN_DISCREPANCY; nparams = 1; sigmasOf(*nf, prm, 1);
            } else if (auto f = dynamic_cast<const bioslam::Pose3TranslationPrior*>(nf)) { type = B200_F_POSE3_TRANSLATION_PRIOR; nparams = 6; sigmasOf(*nf, prm, 3); const gtsam::Vector3 t = FactorAccess::prior(*f); for (int i = 0; i < 3; i++) prm[3 + i] = t[i];
            } else if (auto f = dynamic_cast<const bioslam::Pose3CompassPrior*>(nf)) { type = B200_F_POSE3_COMPASS_PRIOR; nparams = 11; sigmasOf(*nf, prm, 1); FactorAccess::compass(*f, prm);
            } else if (auto f = dynamic_cast<const gtsam::PriorFactor<gtsam::Vector3>*>(nf)) { type = B200_F_PRIOR_VEC3; nparams = 6; sigmasOf(*nf, prm, 3); for (int i = 0; i < 3; i++) prm[3 + i] = f->prior()[i];
            } else if (auto f = dynamic_cast<const gtsam::PriorFactor<gtsam::imuBias::ConstantBias>*>(nf)) { type = B200_F_PRIOR_BIAS6; nparams = 12; sigmasOf(*nf, prm, 6); const auto b = f->prior().vector(); for (int i = 0; i < 6; i++) prm[6 + i] = b[i];
            } else if (auto f = dynamic_cast<const gtsam::PriorFactor<gtsam::Unit3>*>(nf)) { type = B200_F_PRIOR_UNIT3; nparams = 5; sigmasOf(*nf, prm, 2); const auto u = f->prior().unitVector(); for (int i = 0; i < 3; i++) prm[2 + i] = u[i];
            } else throw std::runtime_error("b200Adapter: factor type outside include/bioslam_b200.h (add a branch: AngleBetweenAxisAndSegment, norm-type joint factors, MagPose3Factor, ImuFactor follow the same pattern)");
            b200_factor_slot& s = slotFor(type, *nf, nconst, prm, nparams);
            // keyframe of this instance: the smallest index among its time variables
            uint64_t k = UINT64_MAX;
            for (const gtsam::Key key : nf->keys()) if (timeIdx.count(Symbol(key).chr())) k = std::min<uint64_t>(k, Symbol(key).index());
            const int kk = (k == UINT64_MAX) ? 0 : (int)(k - kMin);
            if (pass == 0) {
                s.k_begin = std::min(s.k_begin, kk); s.k_end = std::max(s.k_end, kk + 1);
                int i = 0;
                for (const gtsam::Key key : nf->keys()) {
                    const Symbol sym(key);
                    if (staticIdx.count(sym.chr())) { s.var_kind[i] = B200_AT_STATIC; s.var_index[i] = staticIdx[sym.chr()]; }
                    else { s.var_kind[i] = (sym.index() == k) ? B200_AT_K : B200_AT_K1; s.var_index[i] = timeIdx[sym.chr()]; }
                    i++;
                }
            } else if (nconst) std::copy(rec, rec + nconst, fg.consts.begin() + (size_t)kk * constStride + s.const_offset);
        }
    }
    // ---- values
    int tvd = 0, svd = 0;
    for (int t : fg.timeVarType) tvd += kValDim[t];
    for (int t : fg.staticVarType) svd += kValDim[t];
    fg.timeValues.assign((size_t)K * tvd, 0.0); fg.staticValues.assign(svd, 0.0);
    for (int k = 0; k < K; k++) {
        int off = 0;
        for (size_t v = 0; v < fg.timeVarType.size(); v++) { packValue(values, Symbol(fg.timeVarChr[v], kMin + k), fg.timeVarType[v], &fg.timeValues[(size_t)k * tvd + off]); off += kValDim[fg.timeVarType[v]]; }
    }
    { int off = 0; for (size_t v = 0; v < fg.staticVarType.size(); v++) { packValue(values, Symbol(fg.staticVarChr[v], *idx[fg.staticVarChr[v]].begin()), fg.staticVarType[v], &fg.staticValues[off]); off += kValDim[fg.staticVarType[v]]; } }
    fg.desc.K = K; fg.desc.n_time_vars = (int)fg.timeVarType.size(); fg.desc.time_var_type = fg.timeVarType.data();
    fg.desc.n_static_vars = (int)fg.staticVarType.size(); fg.desc.static_var_type = fg.staticVarType.data();
    fg.desc.n_slots = (int)fg.slots.size(); fg.desc.slots = fg.slots.data();
    fg.desc.const_stride = constStride; fg.desc.consts = fg.consts.data();
    return fg;
}

gtsam::Values unflatten(const FlatGraph& fg, const gtsam::Values& like) {
    gtsam::Values out;
    int tvd = 0; for (int t : fg.timeVarType) tvd += kValDim[t];
    auto put = [&](gtsam::Key key, int type, const double* x) {
        switch (type) {
        case B200_VAR_POSE3: { gtsam::Matrix3 R; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R(r, c) = x[3 * r + c];
                               out.insert(key, gtsam::Pose3(gtsam::Rot3(R), gtsam::Point3(x[9], x[10], x[11]))); break; }
        case B200_VAR_BIAS6: out.insert(key, gtsam::imuBias::ConstantBias(gtsam::Vector3(x[0], x[1], x[2]), gtsam::Vector3(x[3], x[4], x[5]))); break;
        case B200_VAR_UNIT3: out.insert(key, gtsam::Unit3(x[0], x[1], x[2])); break;
        default: out.insert(key, gtsam::Vector3(x[0], x[1], x[2]));
        }
    };
    for (int k = 0; k < fg.desc.K; k++) { int off = 0; for (size_t v = 0; v < fg.timeVarType.size(); v++) { put(Symbol(fg.timeVarChr[v], fg.kMin + k), fg.timeVarType[v], &fg.timeValues[(size_t)k * tvd + off]); off += kValDim[fg.timeVarType[v]]; } }
    { int off = 0; for (size_t v = 0; v < fg.staticVarType.size(); v++) { gtsam::Key key = 0; for (const gtsam::Key k : like.keys()) if (Symbol(k).chr() == fg.staticVarChr[v]) key = k; put(key, fg.staticVarType[v], &fg.staticValues[off]); off += kValDim[fg.staticVarType[v]]; } }
    return out;
}

RunResult optimizeLowerBody(FlatGraph& fg, double lambdaInitial, double lambdaFactor, double lambdaLower, double lambdaUpper, double absErrDecreaseLimit,
                            double relErrDecreaseLimit, double absErrLimit, int maxIterations, int convergeAtConsecSmallIter, int device) {
    b200_problem* prob = nullptr;
    if (b200_problem_create(&fg.desc, device, &prob) != B200_OK) throw std::runtime_error(std::string("b200_problem_create: ") + b200_last_error());
    b200_set_values(prob, fg.timeValues.data(), fg.staticValues.empty() ? nullptr : fg.staticValues.data());
    b200_lm_params lm{lambdaInitial, lambdaFactor, lambdaLower, lambdaUpper, 1.0e-20, 1.0e-20, 1.0e-3, 0, 0};
    RunResult r; r.errorHistory.resize((size_t)maxIterations + 2);
    b200_outer_params outer{absErrDecreaseLimit, relErrDecreaseLimit, absErrLimit, maxIterations, convergeAtConsecSmallIter, 0, (int)r.errorHistory.size()};
    b200_optimize_result res{};
    r.status = b200_optimize(prob, &lm, &outer, r.errorHistory.data(), &res);
    r.errorHistory.resize(res.n_history);
    r.iterations = res.iterations; r.trials = res.total_trials; r.finalError = res.final_error; r.deviceMs = res.device_ms;
    b200_get_values(prob, fg.timeValues.data(), fg.staticValues.empty() ? nullptr : fg.staticValues.data());
    b200_problem_destroy(prob);
    return r;
}

}  // namespace b200
