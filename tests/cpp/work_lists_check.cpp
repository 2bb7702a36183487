// Host-side check of the balanced planner (bodge_b200/csrc/work_lists.h), built and run by tests/test_work_lists.py:
// every (panel, patch column, plane) unit is covered exactly once, pieces stay inside their column, the plan fits the CTA
// slots, and `longest` is the cost of the longest CTA.  No device code runs.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "work_lists.h"

static int check(int n_panels, int n_columns, int Lx, long slots, int piece_cost, bool grouped_wanted) {
    const WorkPlan plan = balanced_plan(n_panels, n_columns, Lx, slots, piece_cost, grouped_wanted);
    std::vector<int> seen((size_t)n_panels * n_columns * Lx, 0);
    double longest = 0.0;
    if ((long)plan.per_cta.size() > slots || plan.per_cta.empty()) return 1;
    for (const auto &cta : plan.per_cta) {
        if (cta.empty()) return 2;
        double cost = 0.0;
        for (size_t k = 0; k < cta.size(); ++k) {
            const WorkPiece &pc = cta[k];
            if (pc[0] < 0 || pc[0] >= n_panels || pc[1] < 0 || pc[1] >= n_columns) return 3;
            if (pc[3] < 1 || pc[2] < 0 || pc[2] + pc[3] > Lx) return 4;
            // a CTA walks the space in ascending (panel, column, x) order: its dot-product runs are contiguous per panel
            if (k && (pc[0] < cta[k - 1][0])) return 5;
            for (int x = pc[2]; x < pc[2] + pc[3]; ++x) seen[((size_t)pc[0] * n_columns + pc[1]) * Lx + x] += 1;
            cost += pc[3] + piece_cost;
        }
        if (cost > longest) longest = cost;
    }
    for (int v : seen)
        if (v != 1) return 6;
    if (longest != plan.longest) return 7;
    return 0;
}

int main() {
    long cases = 0;
    for (int n_panels : {1, 2, 3, 7, 32, 64, 512})
        for (int n_columns : {1, 2, 7, 8, 9, 72, 300})
            for (int Lx : {3, 4, 7, 8, 9, 15, 16, 17, 64, 100, 250, 1000})
                for (long slots : {1L, 2L, 7L, 148L, 296L, 444L})
                    for (int piece_cost : {0, 5})
                        for (int grouped = 0; grouped < 2; ++grouped) {
                            if ((long)n_panels * n_columns * Lx > 2000000L) continue;  // keep the run to seconds
                            const int rc = check(n_panels, n_columns, Lx, slots, piece_cost, grouped != 0);
                            if (rc) {
                                std::printf("FAIL rc=%d panels=%d columns=%d Lx=%d slots=%ld cost=%d grouped=%d\n", rc, n_panels,
                                            n_columns, Lx, slots, piece_cost, grouped);
                                return 1;
                            }
                            ++cases;
                        }
    // the better-of-two choice never returns a plan longer than the grouped one
    for (int n_panels : {1, 8, 64})
        for (int n_columns : {7, 8, 72}) {
            const WorkPlan best = best_balanced_plan(n_panels, n_columns, 100, 296, 5, 1.08);
            const WorkPlan grouped = balanced_plan(n_panels, n_columns, 100, 296, 5, true);
            if (best.longest > grouped.longest) {
                std::printf("FAIL best plan longer than the grouped one: panels=%d columns=%d\n", n_panels, n_columns);
                return 1;
            }
        }
    std::printf("ok %ld plans\n", cases);
    return 0;
}
