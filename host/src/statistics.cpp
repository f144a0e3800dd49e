// statistics.cpp -- jackknife errors (reference: src/statistics.cpp:6-45), host post-processing.
#include "statistics.h"

#include <algorithm>
#include <cmath>
#include <numeric>

std::vector<double> samples_mean(std::vector<double> dat, int bin) {
    // mean of the data with bin i left out; only the first bin*per entries belong to a bin
    const int per = (int)dat.size() / bin;
    std::vector<double> bin_sum(bin, 0.0), out(bin);
    for (int k = 0; k < bin; k++)
        for (int j = k * per; j < (k + 1) * per; j++) bin_sum[k] += dat[j];
    for (int i = 0; i < bin; i++) {
        double s = 0.0;
        for (int k = 0; k < bin; k++)
            if (k != i) s += bin_sum[k];
        out[i] = s / (dat.size() - per);
    }
    return out;
}

double Jackknife_error(std::vector<double> dat, int bin) {
    const std::vector<double> sm = samples_mean(dat, bin);
    const double m = mean(dat);
    double err = 0.0;
    for (double s : sm) err += (s - m) * (s - m);
    return std::sqrt(err * (bin - 1) / bin);
}

double Jackknife(std::vector<double> dat, std::vector<int> bins) {
    double worst = 0.0;
    for (int b : bins) worst = std::max(worst, Jackknife_error(dat, b));
    return worst;
}
