// Balanced plans of the two-applications-per-pass kernels (cheb_pair.cu, cheb_cube.cu): the (panel, patch column, x) space of
// a launch cut into one contiguous chunk per CTA slot and flattened into the device-side work lists the `LISTED` kernel
// instantiations read.  Host code, internal.
#pragma once

#include <algorithm>
#include <array>
#include <vector>

#include "bdg_internal.h"

namespace {

using WorkPiece = std::array<int, 4>;  // panel, patch column, x0, len

struct WorkPlan {
    std::vector<std::vector<WorkPiece>> per_cta;
    double longest = 0.0;  // iterations of the longest CTA: sum over its pieces of (len + piece_cost)
};

// `n_columns` patch columns of `Lx` planes per panel, `n_panels` panels, `slots` CTAs resident at a time.
inline WorkPlan balanced_plan(int n_panels, int n_columns, int Lx, int64_t slots, int piece_cost) {
    WorkPlan plan;
    const int64_t total = (int64_t)n_panels * n_columns * Lx;  // plane units, (panel, column)-major
    const int64_t chunk = std::max<int64_t>(16, ceil_div(total, slots));
    const int64_t n_chunks = ceil_div(total, chunk);
    int64_t begin = 0;
    for (int64_t c = 1; c <= n_chunks; ++c) {
        int64_t end = std::min(total, total * c / n_chunks);
        const int64_t into = end % Lx;  // a cut close to a column boundary moves onto it (no sliver pieces)
        if (c < n_chunks && into > 0 && into < 8) end -= into;
        else if (c < n_chunks && into > Lx - 8) end += Lx - into;
        if (end <= begin) continue;
        std::vector<WorkPiece> mine;
        int64_t cost = 0;
        for (int64_t u = begin; u < end;) {
            const int64_t col = u / Lx;
            const int x0 = (int)(u - col * Lx), len = (int)std::min<int64_t>(Lx - x0, end - u);
            mine.push_back({(int)(col / n_columns), (int)(col % n_columns), x0, len});
            cost += len + piece_cost;
            u += len;
        }
        plan.longest = std::max(plan.longest, (double)cost);
        plan.per_cta.push_back(std::move(mine));
        begin = end;
    }
    return plan;
}

struct WorkLists {  // device pointers into one buffer
    const int4 *pieces = nullptr;
    const int *cta_begin = nullptr, *cta_run0 = nullptr, *panel_runs = nullptr;
    int n_ctas = 0, n_runs = 0;
};

// Upload: pieces (16-byte aligned: first), cta_begin[n_ctas + 1], cta_run0[n_ctas], panel_runs[n_panels + 1].  A run = a maximal
// sequence of pieces of one panel inside a CTA; runs are numbered in (panel, CTA) order (the plan is panel-major).
inline int upload_work_lists(bdg_system *sys, DevBuf &buf, const WorkPlan &plan, int n_panels, WorkLists &out) {
    std::vector<int> flat, cta_begin{0}, cta_run0, panel_runs(n_panels + 1, 0);
    int n_runs = 0;
    for (const auto &mine : plan.per_cta) {
        cta_run0.push_back(n_runs);
        int last_panel = -1;
        for (const WorkPiece &pc : mine) {
            flat.insert(flat.end(), pc.begin(), pc.end());
            if (pc[0] != last_panel) {
                last_panel = pc[0];
                panel_runs[pc[0] + 1] += 1;
                n_runs += 1;
            }
        }
        cta_begin.push_back((int)(flat.size() / 4));
    }
    for (int p = 0; p < n_panels; ++p) panel_runs[p + 1] += panel_runs[p];
    const size_t o_begin = flat.size(), o_run0 = o_begin + cta_begin.size(), o_panel = o_run0 + cta_run0.size();
    flat.insert(flat.end(), cta_begin.begin(), cta_begin.end());
    flat.insert(flat.end(), cta_run0.begin(), cta_run0.end());
    flat.insert(flat.end(), panel_runs.begin(), panel_runs.end());
    BDG_TRY(dev_alloc(sys, buf, flat.size() * sizeof(int)));
    BDG_CUDA(cudaMemcpyAsync(buf.ptr, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice, sys->stream));
    BDG_CUDA(cudaStreamSynchronize(sys->stream));  // `flat` goes out of scope
    const int *base = buf.as<int>();
    out.pieces = reinterpret_cast<const int4 *>(base);
    out.cta_begin = base + o_begin;
    out.cta_run0 = base + o_run0;
    out.panel_runs = base + o_panel;
    out.n_ctas = (int)plan.per_cta.size();
    out.n_runs = n_runs;
    return BDG_OK;
}

}  // namespace
