// marginalization_factor_device.cpp -- drop-in replacement for the reference's factor/marginalization_factor.cpp (same header, same class layout):
// MarginalizationInfo::marginalize() runs on the GPU (viwb_marginalize) instead of the four pthreads + Eigen of marginalization_factor.cpp:98-312.
//
// Integration (adapter route): build vins_estimator with viw-fusion_b200/host first on the include path (its <ceres/ceres.h>),
// take factor/marginalization_factor.cpp OUT of the source list and put this file in.  estimator.cpp stays untouched: it keeps calling
//   new MarginalizationInfo(); addResidualBlockInfo(...); preMarginalize(); marginalize(); getParameterBlocks(addr_shift); new MarginalizationFactor(info)
// (estimator.cpp:1669-1893) on the reference's own classes.  The factor classes are recognised through viwb_reference_adapter.h (install it once).
#include "factor/marginalization_factor.h"          // the REFERENCE'S header (vins_estimator/src on the include path)
#include "factor/viwb_marginalization_lower.h"

namespace viwb_shim {
// exact block ids of the priors this translation unit produced (the reference's class has no member for them); the adapter reads them when it
// lowers a MarginalizationFactor (viwb_reference_adapter.h), so that scalar blocks (td, td_wheel, sx, sy, sw, plane_Z) are never guessed by size
std::map<const void *, std::vector<int>> &device_prior_ids() { static std::map<const void *, std::vector<int>> m; return m; }
}

void ResidualBlockInfo::Evaluate() {}               // the factors are linearised on the GPU inside viwb_marginalize; nothing is evaluated on the host

MarginalizationInfo::~MarginalizationInfo() {
    viwb_shim::device_prior_ids().erase(this);
    for (auto it = parameter_block_data.begin(); it != parameter_block_data.end(); ++it) delete[] it->second;
    for (int i = 0; i < (int)factors.size(); i++) { delete factors[i]->cost_function; delete factors[i]; }
}
int MarginalizationInfo::localSize(int size) const { return size == 7 ? 6 : size; }
int MarginalizationInfo::globalSize(int size) const { return size == 6 ? 7 : size; }

void MarginalizationInfo::addResidualBlockInfo(ResidualBlockInfo *info) {
    factors.emplace_back(info);
    const std::vector<int> sizes(info->cost_function->parameter_block_sizes().begin(), info->cost_function->parameter_block_sizes().end());
    for (int i = 0; i < (int)info->parameter_blocks.size(); i++) parameter_block_size[reinterpret_cast<long>(info->parameter_blocks[i])] = sizes[i];
    for (int d : info->drop_set) parameter_block_idx[reinterpret_cast<long>(info->parameter_blocks[d])] = 0;
}

// the linearisation point of every block (marginalization_factor.cpp:160-181 keeps a copy per block; the prior's keep_block_data point into them)
void MarginalizationInfo::preMarginalize() {
    for (auto *f : factors) {
        const auto &sizes = f->cost_function->parameter_block_sizes();
        for (int i = 0; i < (int)sizes.size(); i++) {
            const long addr = reinterpret_cast<long>(f->parameter_blocks[i]);
            if (parameter_block_data.find(addr) != parameter_block_data.end()) continue;
            double *copy = new double[sizes[i]];
            std::memcpy(copy, f->parameter_blocks[i], sizeof(double) * sizes[i]);
            parameter_block_data[addr] = copy;
        }
    }
}

void MarginalizationInfo::marginalize() {
    std::vector<viwb_shim::MargFactor> fs;
    for (auto *f : factors) fs.push_back(viwb_shim::MargFactor{f->cost_function, f->loss_function, &f->parameter_blocks, &f->drop_set});
    viwb_shim::MargResult res;
    viwb_shim::marginalize_factors(fs, res);
    if (!res.valid) { valid = false; return; }
    const viwb_prior &pr = res.prior;
    m = res.m; n = pr.n;
    // index map in the reference's convention: dropped blocks in [0, m), kept block i at m + (its column in the prior)
    int pos = 0;
    for (auto &kv : parameter_block_idx) { kv.second = pos; pos += localSize(parameter_block_size[kv.first]); }
    std::vector<int> ids;
    for (int i = 0; i < pr.num_blocks; i++) {
        if (!res.kept_addr[i]) { valid = false; return; }
        parameter_block_idx[reinterpret_cast<long>(res.kept_addr[i])] = m + pr.block_idx[i];
        ids.push_back(pr.block_id[i]);
    }
    linearized_jacobians.resize(n, n); linearized_residuals.resize(n);
    for (int i = 0; i < n; i++) { linearized_residuals(i) = pr.r[i]; for (int j = 0; j < n; j++) linearized_jacobians(i, j) = pr.J[(size_t)i * n + j]; }
    viwb_shim::device_prior_ids()[this] = ids;
}

// kept blocks in the order of the device prior (marginalization_factor.cpp:314-334 walks its hash map; the order is free, the consumer pairs
// keep_block_* with the returned addresses)
std::vector<double *> MarginalizationInfo::getParameterBlocks(std::unordered_map<long, double *> &addr_shift) {
    std::vector<double *> keep_block_addr;
    keep_block_size.clear(); keep_block_idx.clear(); keep_block_data.clear();
    std::vector<std::pair<int, long>> kept;          // (column, address)
    for (const auto &kv : parameter_block_idx) if (kv.second >= m) kept.push_back({kv.second, kv.first});
    std::sort(kept.begin(), kept.end());
    for (const auto &k : kept) {
        keep_block_size.push_back(parameter_block_size[k.second]);
        keep_block_idx.push_back(k.first);
        keep_block_data.push_back(parameter_block_data[k.second]);
        keep_block_addr.push_back(addr_shift[k.second]);
    }
    sum_block_size = 0; for (int s : keep_block_size) sum_block_size += s;
    return keep_block_addr;
}

MarginalizationFactor::MarginalizationFactor(MarginalizationInfo *_marginalization_info) : marginalization_info(_marginalization_info) {
    for (auto s : marginalization_info->keep_block_size) mutable_parameter_block_sizes()->push_back(s);
    set_num_residuals(marginalization_info->n);
}

// MarginalizationFactor::Evaluate (marginalization_factor.cpp:349-397) on the GPU: ceres::Solve lowers the prior from the info's members, so this is
// only reached by code that evaluates the factor by hand
bool MarginalizationFactor::Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
    const MarginalizationInfo *info = marginalization_info;
    const auto it = viwb_shim::device_prior_ids().find(info);
    if (it == viwb_shim::device_prior_ids().end()) return false;
    const std::vector<int> &ids = it->second;
    const int n = info->n, nb = (int)info->keep_block_size.size();
    if ((int)ids.size() != nb) return false;
    viwb_prior pr; std::memset(&pr, 0, sizeof pr);
    std::vector<double> x0(VIWB_STATE_FIXED, 0.0), J((size_t)n * n), r(n), state(VIWB_STATE_FIXED, 0.0), jac;
    pr.valid = 1; pr.n = n; pr.num_blocks = nb;
    // the ids were stored in the device prior's block order; keep_block_* in column order: match them through the column
    std::vector<int> id_of_col(n + 1, -1);
    { std::vector<int> cols(info->keep_block_idx); std::vector<int> sorted_ids(ids);
      // both lists enumerate the same blocks; the device prior's block order IS column order (the kernel assigns columns in block order)
      for (int i = 0; i < nb; i++) id_of_col[cols[i] - info->m] = sorted_ids[i]; }
    for (int i = 0; i < nb; i++) {
        const int bid = id_of_col[info->keep_block_idx[i] - info->m];
        if (bid < 0) return false;
        pr.block_id[i] = bid; pr.block_idx[i] = info->keep_block_idx[i] - info->m;
        std::memcpy(x0.data() + viwb_block_offset(bid), info->keep_block_data[i], sizeof(double) * viwb_block_size(bid));
        std::memcpy(state.data() + viwb_block_offset(bid), parameters[i], sizeof(double) * viwb_block_size(bid));
    }
    for (int i = 0; i < n; i++) { r[i] = info->linearized_residuals(i); for (int j = 0; j < n; j++) J[(size_t)i * n + j] = info->linearized_jacobians(i, j); }
    pr.x0 = x0.data(); pr.J = J.data(); pr.r = r.data();
    if (jacobians) jac.resize((size_t)n * VIWB_STATE_FIXED);
    viwb_context *ctx = viwb_shim::context();
    if (!ctx || viwb_prior_evaluate(ctx, &pr, state.data(), residuals, jacobians ? jac.data() : nullptr) != VIWB_OK) return false;
    if (jacobians) for (int i = 0; i < nb; i++) if (jacobians[i]) {
        const int gs = viwb_block_size(pr.block_id[i]), off = viwb_block_offset(pr.block_id[i]);
        for (int rr = 0; rr < n; rr++) for (int c = 0; c < gs; c++) jacobians[i][rr * gs + c] = jac[(size_t)rr * VIWB_STATE_FIXED + off + c];
    }
    return true;
}
