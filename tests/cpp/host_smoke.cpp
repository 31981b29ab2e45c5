// Drives the engine from C++ through include/smcb200.hpp (the reference's exported function names on std::vector).
// Prints one line per call; exit code 10 + status when the library reports an error (12 = no CUDA device).
#include <cmath>
#include <cstdio>

#include "smcb200.hpp"

int main() {
    std::vector<double> y(24);
    for (size_t i = 0; i < y.size(); ++i) y[i] = 3.0 * std::sin(0.7 * (double)i) + 0.1 * (double)i;
    try {
        smcb200::pfNonlinBS_result a = smcb200::pfNonlinBS_impl(y, 4096, /*seed=*/7, SMCB200_SYSTEMATIC);
        std::printf("pfNonlinBS mean");
        for (double m : a.mean) std::printf(" %.17g", m);
        std::printf("\n");
        smcb200::blockpfGaussianOpt_result b = smcb200::blockpfGaussianOpt_impl(y, 512, 3, /*seed=*/5);
        double wsum = 0.0;
        for (double w : b.weight) wsum += w;
        std::printf("blockpfGaussianOpt logNC %.17g weight_sum %.12f values %zu\n", b.logNC, wsum, b.values.size());
        smcb200::nonLinPMMH_result p = smcb200::nonLinPMMH_impl(y, 512, 3);
        std::printf("nonLinPMMH loglike0 %.17g n %zu\n", p.loglike[0], p.loglike.size());
        std::vector<double> ref(y.size() * 2, 0.0);
        smcb200::compareNCestimates_result c = smcb200::compareNCestimates_imp(y, 200, 2, smcb200::parInits{0.9, 1.0, 0.0, 1.0}, ref);
        std::printf("compareNCestimates smcOut[0] %.17g csmcOut[0] %.17g\n", c.smcOut[0], c.csmcOut[0]);
        try {
            smcb200::blockpfGaussianOpt_impl(y, 512, 0);
            return 1;
        } catch (const smcb200::error& e) {
            std::printf("argument error reported: status %d\n", e.status);
        }
        return 0;
    } catch (const smcb200::error& e) {
        std::fprintf(stderr, "smcb200 error %d: %s\n", e.status, e.what());
        return 10 + e.status;
    }
}
