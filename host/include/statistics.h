// statistics.h -- host-side post-processing of the measurement history.
// Same entry points as the reference (include/statistics.h, src/statistics.cpp): mean(), rand_range(),
// samples_mean(), Jackknife_error(), Jackknife(), linspace().  None of this is on the GPU path.
#ifndef SM_HOST_STATISTICS_H
#define SM_HOST_STATISTICS_H
#include <cstdlib>
#include <numeric>
#include <vector>

// arithmetic mean; the accumulator is a double whatever T is
template <typename T>
double mean(std::vector<T> x) {
    return std::accumulate(x.begin(), x.end(), 0.0, [](double acc, const T& v) { return acc + v * 1.0; }) / x.size();
}

// uniform double in [a, b] drawn from rand(): the generator of the reference's Metropolis step
inline double rand_range(double a, double b) {
    const double u = static_cast<double>(rand()) / (RAND_MAX);
    return a + (b - a) * u;
}

// leave-one-bin-out means, jackknife error for a given bin count, and the worst case over several
std::vector<double> samples_mean(std::vector<double> dat, int bin);
double Jackknife_error(std::vector<double> dat, int bin);
double Jackknife(std::vector<double> dat, std::vector<int> bins);

// n equally spaced values from min to max inclusive
template <typename T>
std::vector<double> linspace(T min, T max, int n) {
    std::vector<double> grid;
    grid.reserve(n);
    const double lo = 1.0 * min, step = (1.0 * max - lo) / (n - 1);
    for (int i = 0; i < n; ++i) grid.push_back(lo + i * step);
    return grid;
}
#endif
