// statistics.h -- host-side post-processing (reference: include/statistics.h, src/statistics.cpp).
#ifndef SM_HOST_STATISTICS_H
#define SM_HOST_STATISTICS_H
#include <cstdlib>
#include <vector>

template <typename T>
double mean(std::vector<T> x) {
    double acc = 0;
    for (const T& v : x) acc += v * 1.0;
    return acc / x.size();
}

// uniform double in [a,b] from rand(), the generator the reference's Metropolis step uses
inline double rand_range(double a, double b) { return (b - a) * ((double)rand() / (RAND_MAX)) + a; }

std::vector<double> samples_mean(std::vector<double> dat, int bin);   // leave-one-bin-out means
double Jackknife_error(std::vector<double> dat, int bin);
double Jackknife(std::vector<double> dat, std::vector<int> bins);     // worst case over bin counts

template <typename T>
std::vector<double> linspace(T min, T max, int n) {
    std::vector<double> out(n);
    const double h = (1.0 * max - 1.0 * min) / (n - 1);
    for (int i = 0; i < n; ++i) out[i] = min * 1.0 + i * h;
    return out;
}
#endif
