// Gauss-Legendre nodes / weights for the host mirror (reference: gauss_quadrature_points, glq.rs:179-222, which runs
// Golub-Welsch through nalgebra's SymmetricEigen).  Nodes and weights are *inputs* of the C-ABI (include/fem2d.h), so a
// Rust caller passes nalgebra's exact values; this mirror produces them to ~1 ulp with Newton iteration on P_n in
// extended precision -- the reference's own test pins them to 1e-9 only (glq.rs:255-343).
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace fem2d {

inline void gauss_quadrature_points(size_t n, bool include_endpoints, std::vector<double>& points, std::vector<double>& weights) {
    points.assign(n, 0.0); weights.assign(n, 0.0);
    const long double PI = 3.141592653589793238462643383279502884L;
    for (size_t k = 0; k < (n + 1) / 2; k++) {
        long double x = std::cos(PI * ((long double)k + 0.75L) / ((long double)n + 0.5L));   // k-th largest root
        long double dp = 1;
        for (int it = 0; it < 100; it++) {
            long double p0 = 1, p1 = x;
            for (size_t j = 2; j <= n; j++) { long double p2 = ((2 * (long double)j - 1) * x * p1 - ((long double)j - 1) * p0) / (long double)j; p0 = p1; p1 = p2; }
            if (n == 0) { p1 = 1; p0 = 0; }
            dp = (long double)n * (x * p1 - p0) / (x * x - 1);
            const long double dx = p1 / dp;
            x -= dx;
            if (std::fabs((double)dx) < 1e-19) break;
        }
        {   // derivative at the converged root
            long double p0 = 1, p1 = x;
            for (size_t j = 2; j <= n; j++) { long double p2 = ((2 * (long double)j - 1) * x * p1 - ((long double)j - 1) * p0) / (long double)j; p0 = p1; p1 = p2; }
            dp = (long double)n * (x * p1 - p0) / (x * x - 1);
        }
        const long double w = 2 / ((1 - x * x) * dp * dp);
        points[n - 1 - k] = (double)x; weights[n - 1 - k] = (double)w;
        points[k] = (double)(-x); weights[k] = (double)w;
    }
    if (n % 2 == 1) points[n / 2] = 0.0;
    if (include_endpoints) {   // glq.rs:213-219
        points.insert(points.begin(), -1.0); points.push_back(1.0);
        weights.insert(weights.begin(), 1.0); weights.push_back(1.0);
    }
}

// 4 * max_order rounded up to a power of two (basis.rs:172-177; f32 arithmetic as in the reference).
inline size_t default_ngq(size_t max_order) {
    const float conv = (float)(max_order * 4);
    const int p2 = (int)std::ceil(std::log2(conv));
    return (size_t)std::lround(std::pow(2.0f, (float)p2));
}

// glq.rs:238-249
inline double scale_gauss_quad_points(const std::vector<double>& pts, double mn, double mx, std::vector<double>& out) {
    const double s = (mx - mn) / 2.0, o = (mx + mn) / 2.0;
    out.resize(pts.size());
    for (size_t k = 0; k < pts.size(); k++) out[k] = pts[k] * s + o;
    return s;
}

}  // namespace fem2d
