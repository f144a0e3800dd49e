// statistics.cpp -- jackknife error estimates for the measurement history (host post-processing).
//
// Semantics follow the reference (src/statistics.cpp:6-45) exactly, including its corner case: the data
// are cut into `nbins` bins of floor(N / nbins) entries, entries beyond nbins * floor(N / nbins) belong to
// no bin, yet the divisor of every leave-one-out mean is N - binsize and the plain mean runs over all N.
#include "statistics.h"

#include <algorithm>
#include <cmath>

namespace {

// sum of each bin and of all binned entries
struct Binned {
    std::vector<double> sum;
    double total = 0.0;
    std::size_t binsize = 0;
};

Binned bin_sums(const std::vector<double>& data, int nbins) {
    Binned b;
    b.binsize = data.size() / nbins;
    b.sum.assign(nbins, 0.0);
    for (int k = 0; k < nbins; ++k) {
        const auto first = data.begin() + k * b.binsize;
        for (auto it = first; it != first + b.binsize; ++it) b.sum[k] += *it;
    }
    return b;
}

}  // namespace

std::vector<double> samples_mean(std::vector<double> dat, int bin) {
    const Binned b = bin_sums(dat, bin);
    std::vector<double> without(bin);
    for (int leave = 0; leave < bin; ++leave) {
        // add the other bins in index order (the order fixes the rounding, keep it)
        double rest = 0.0;
        for (int k = 0; k < bin; ++k)
            if (k != leave) rest += b.sum[k];
        without[leave] = rest / (dat.size() - b.binsize);
    }
    return without;
}

double Jackknife_error(std::vector<double> dat, int bin) {
    const double full = mean(dat);
    double spread = 0.0;
    for (double m : samples_mean(dat, bin)) spread += (m - full) * (m - full);
    return std::sqrt(spread * (bin - 1) / bin);
}

double Jackknife(std::vector<double> dat, std::vector<int> bins) {
    double worst = 0.0;
    for (int nb : bins) worst = std::max(worst, Jackknife_error(dat, nb));
    return worst;
}
