// Host harness for csrc/nrldpc_bp_math.cuh (tests/test_bp_math.py): the float64 tanh(q/2) and 2 atanh(x) of the
// sum-product kernel against glibc's long double functions on 4 x 10^6 points; prints the worst errors in ulp.
#include "nrldpc_bp_math.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
using namespace nrldpc::bpmath;
static double ulp_err(double got, long double ref) {
    if (ref == 0) return got == 0 ? 0 : 1e9;
    int ex; frexpl(ref, &ex);
    long double ulp = ldexpl(1.0L, ex - 53);
    return (double)(fabsl((long double)got - ref) / ulp);
}
int main() {
    std::mt19937_64 g(1);
    std::uniform_real_distribution<double> U(0, 1);
    double worst_t = 0, wt_at = 0, worst_a = 0, wa_at = 0, worst_lt=0, worst_la=0;
    for (int i = 0; i < 4000000; ++i) {
        // q: log-uniform magnitudes 1e-12 ... 90, both signs
        double q = exp(log(1e-12) + U(g) * (log(90.0) - log(1e-12))) * (U(g) < 0.5 ? -1 : 1);
        double e = ulp_err(tanh_half(q), tanhl((long double)q / 2));
        if (e > worst_t) { worst_t = e; wt_at = q; }
        double el = ulp_err(tanh(q/2), tanhl((long double)q / 2)); if (el > worst_lt) worst_lt = el;
        // x: |x| in (0,1): mix of log-uniform small, uniform, and 1 - log-uniform
        double r = U(g), x;
        if (r < 0.33) x = exp(log(1e-12) + U(g) * (0 - log(1e-12)));
        else if (r < 0.66) x = U(g);
        else x = 1.0 - exp(log(1.2e-16) + U(g) * (0 - log(1.2e-16)));
        if (x >= 1.0 || x <= 0) continue;
        if (U(g) < 0.5) x = -x;
        e = ulp_err(atanh_twice(x), 2 * atanhl((long double)x));
        if (e > worst_a) { worst_a = e; wa_at = x; }
        el = ulp_err(2*atanh(x), 2 * atanhl((long double)x)); if (el > worst_la) worst_la = el;
    }
    // the ends of the domain: the largest double below 1, tiny arguments (the series' first term alone)
    const double top = nextafter(1.0, 0.0), tops[4] = {top, -top, 1e-300, -3e-9};
    for (double x : tops) {
        double e = ulp_err(atanh_twice(x), 2 * atanhl((long double)x));
        if (e > worst_a) { worst_a = e; wa_at = x; }
        e = ulp_err(tanh_half(x), tanhl((long double)x / 2));
        if (e > worst_t) { worst_t = e; wt_at = x; }
    }
    printf("tanh_half max ulp %.3f at %.17g (libm %.3f)\natanh_twice max ulp %.3f at %.17g (libm %.3f)\n", worst_t, wt_at, worst_lt, worst_a, wa_at, worst_la);
    printf("%.17g %.17g %.17g\n", tanh_half(0.0), tanh_half(1e300), atanh_twice(0.0));
    return 0;
}
