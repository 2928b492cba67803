// viwb_marginalization_impl.h -- MarginalizationInfo / MarginalizationFactor over viwb_marginalize / viwb_prior_evaluate.
#pragma once

#include "viwb_marginalization_lower.h"

inline void MarginalizationInfo::marginalize() {
    // Lower the collected factors (they ARE the marginalization set of estimator.cpp:1670-1776) to a viwb_problem that
    // contains only them, with the dropped pose as frame 0 (MARGIN_OLD) or the prior alone (MARGIN_SECOND_NEW).
    std::vector<viwb_shim::MargFactor> fs;
    for (auto *f : factors) fs.push_back(viwb_shim::MargFactor{f->cost_function, f->loss_function, &f->parameter_blocks, &f->drop_set});
    viwb_shim::MargResult res;
    viwb_shim::marginalize_factors(fs, res);
    if (!res.valid) { valid = false; return; }
    x0_ = res.x0; J_ = res.J; r_ = res.r;
    prior_ = res.prior; prior_.x0 = x0_.data(); prior_.J = J_.data(); prior_.r = r_.data();
    kept_addr_ = res.kept_addr;
    n = prior_.n; m = res.m;
}

inline std::vector<double *> MarginalizationInfo::getParameterBlocks(std::unordered_map<long, double *> &addr_shift) {
    std::vector<double *> keep_block_addr;
    keep_block_size.clear(); keep_block_idx.clear(); keep_block_data.clear();
    for (int i = 0; i < prior_.num_blocks; i++) {
        const int bid = prior_.block_id[i];
        keep_block_size.push_back(viwb_block_size(bid));
        keep_block_idx.push_back(prior_.block_idx[i] + m);
        keep_block_data.push_back(x0_.data() + viwb_block_offset(bid));
        keep_block_addr.push_back(kept_addr_[i] ? addr_shift[reinterpret_cast<long>(kept_addr_[i])] : nullptr);
    }
    sum_block_size = 0; for (int s : keep_block_size) sum_block_size += s;
    return keep_block_addr;
}

inline bool MarginalizationFactor::Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
    const viwb_prior *p = marginalization_info->prior();
    std::vector<double> state(VIWB_STATE_FIXED, 0.0), jac;
    for (int i = 0; i < p->num_blocks; i++) std::memcpy(state.data() + viwb_block_offset(p->block_id[i]), parameters[i], sizeof(double) * viwb_block_size(p->block_id[i]));
    if (jacobians) jac.resize((size_t)p->n * VIWB_STATE_FIXED);
    viwb_context *ctx = viwb_shim::context();
    if (!ctx || viwb_prior_evaluate(ctx, p, state.data(), residuals, jacobians ? jac.data() : nullptr) != VIWB_OK) return false;
    if (jacobians) for (int i = 0; i < p->num_blocks; i++) if (jacobians[i]) {
        const int gs = viwb_block_size(p->block_id[i]), off = viwb_block_offset(p->block_id[i]);
        for (int r = 0; r < p->n; r++) for (int c = 0; c < gs; c++) jacobians[i][r * gs + c] = jac[(size_t)r * VIWB_STATE_FIXED + off + c];
    }
    return true;
}
