// Flat time-chain graph description builder (host side of include/bioslam_b200.h's b200_problem_desc).
#pragma once
#include <stdexcept>
#include <utility>
#include <vector>
#include "../../include/bioslam_b200.h"

namespace bioslam_b200 {

struct VarRef { int kind, index; };   // kind: B200_AT_K / B200_AT_K1 / B200_AT_STATIC

class GraphBuilder {
  public:
    explicit GraphBuilder(int K) : K_(K) {}
    int K() const { return K_; }
    static int valDim(int t) { return t == B200_VAR_POSE3 ? 12 : (t == B200_VAR_BIAS6 ? 6 : 3); }
    static int tanDim(int t) { return t == B200_VAR_POSE3 ? 6 : (t == B200_VAR_BIAS6 ? 6 : (t == B200_VAR_UNIT3 ? 2 : 3)); }
    // values: K x valDim(type), row-major
    int addTimeVar(int type, std::vector<double> values) {
        if ((int)values.size() != K_ * valDim(type)) throw std::runtime_error("GraphBuilder: wrong value count for time variable");
        tvType_.push_back(type); tvVals_.push_back(std::move(values));
        return (int)tvType_.size() - 1;
    }
    int addStaticVar(int type, const double* value) {
        svType_.push_back(type); svVals_.emplace_back(value, value + valDim(type));
        return (int)svType_.size() - 1;
    }
    // per-keyframe constants: K x n (rows beyond the factor's range may be zero); returns const_offset
    int addConstBlock(int n, std::vector<double> perKeyframe) {
        if ((int)perKeyframe.size() != K_ * n) throw std::runtime_error("GraphBuilder: wrong size of constant block");
        int off = cstride_; cstride_ += n; cBlocks_.emplace_back(n, std::move(perKeyframe));
        return off;
    }
    void addSlot(int type, const std::vector<VarRef>& vars, const std::vector<double>& params, int k_begin, int k_end, int const_offset = -1) {
        b200_factor_slot s{};
        s.type = type; s.k_begin = k_begin; s.k_end = k_end; s.nvars = (int)vars.size(); s.const_offset = const_offset;
        for (size_t i = 0; i < vars.size(); i++) { s.var_kind[i] = vars[i].kind; s.var_index[i] = vars[i].index; }
        if (params.size() > B200_MAX_FACTOR_PARAMS) throw std::runtime_error("GraphBuilder: too many factor parameters");
        for (size_t i = 0; i < params.size(); i++) s.params[i] = params[i];
        slots_.push_back(s);
    }
    void finalize() {
        tvd_ = 0; svd_ = 0; tvOff_.clear(); svOff_.clear();
        for (int t : tvType_) { tvOff_.push_back(tvd_); tvd_ += valDim(t); }
        for (int t : svType_) { svOff_.push_back(svd_); svd_ += valDim(t); }
        timeValues_.assign((size_t)K_ * tvd_, 0.0);
        for (size_t v = 0; v < tvType_.size(); v++) {
            const int d = valDim(tvType_[v]);
            for (int k = 0; k < K_; k++) for (int c = 0; c < d; c++) timeValues_[(size_t)k * tvd_ + tvOff_[v] + c] = tvVals_[v][(size_t)k * d + c];
        }
        staticValues_.assign(svd_, 0.0);
        for (size_t v = 0; v < svType_.size(); v++) for (int c = 0; c < valDim(svType_[v]); c++) staticValues_[svOff_[v] + c] = svVals_[v][c];
        consts_.assign((size_t)K_ * (cstride_ > 0 ? cstride_ : 1), 0.0);
        int off = 0;
        for (auto& cb : cBlocks_) {
            for (int k = 0; k < K_; k++) for (int c = 0; c < cb.first; c++) consts_[(size_t)k * cstride_ + off + c] = cb.second[(size_t)k * cb.first + c];
            off += cb.first;
        }
        desc_.K = K_; desc_.n_time_vars = (int)tvType_.size(); desc_.time_var_type = tvType_.data();
        desc_.n_static_vars = (int)svType_.size(); desc_.static_var_type = svType_.empty() ? nullptr : svType_.data();
        desc_.n_slots = (int)slots_.size(); desc_.slots = slots_.data();
        desc_.const_stride = cstride_; desc_.consts = consts_.data();
    }
    const b200_problem_desc& desc() const { return desc_; }
    std::vector<double>& timeValues() { return timeValues_; }
    std::vector<double>& staticValues() { return staticValues_; }
    int timeValueDim() const { return tvd_; }
    int staticValueDim() const { return svd_; }
    const double* timeValue(int k, int var) const { return timeValues_.data() + (size_t)k * tvd_ + tvOff_[var]; }
    const double* staticValue(int var) const { return staticValues_.data() + svOff_[var]; }

  private:
    int K_, cstride_ = 0, tvd_ = 0, svd_ = 0;
    std::vector<int32_t> tvType_, svType_;
    std::vector<std::vector<double>> tvVals_, svVals_;
    std::vector<std::pair<int, std::vector<double>>> cBlocks_;
    std::vector<b200_factor_slot> slots_;
    std::vector<int> tvOff_, svOff_;
    std::vector<double> timeValues_, staticValues_, consts_;
    b200_problem_desc desc_{};
};

}  // namespace bioslam_b200
