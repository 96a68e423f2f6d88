// Host accuracy check of csrc/tamc_math.cuh against libm in long double (test infrastructure).
// Prints max / mean error in ulps for each function over random and edge arguments.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>

#include "../tissue-ablation-mc_b200/csrc/tamc_math.cuh"

static double ulp_err(double got, long double want)
{
    if (want == 0.0L) return got == 0.0 ? 0.0 : 1e9;
    int e;
    frexpl(want, &e);
    const long double ulp = ldexpl(1.0L, e - 53);
    return (double)(fabsl((long double)got - want) / ulp);
}

int main(int argc, char **argv)
{
    const long n = argc > 1 ? atol(argv[1]) : 4000000;
    std::mt19937_64 rng(12345);
    const long double PI = 3.14159265358979323846264338327950288L;
    double ms = 0, mc = 0, ml = 0, mq = 0, mr = 0, as = 0, al = 0;
    for (long i = 0; i < n; ++i) {
        const uint32_t x = i < 64 ? (uint32_t)i : (i < 128 ? 0xffffffffu - (uint32_t)(i - 64) : (uint32_t)rng());
        // optical depth
        const double tl = tamc::fm::neglog_u32(x);
        const long double wl = -logl(((long double)x + 0.5L) / 4294967296.0L);
        const double el = ulp_err(tl, wl);
        ml = fmax(ml, el); al += el;
        // sincospi on [0, 2]
        double a = (double)(rng() >> 11) * (1.0 / 9007199254740992.0) * 2.0;
        if (i < 9) a = 0.25 * (double)i;
        double s, c;
        tamc::fm::sincospi_0_2(a, &s, &c);
        // reference: exact reduction t = a - q/2 first (sinl(PI*a) itself loses relative accuracy near the zeros)
        const long qq = lrintl(2.0L * (long double)a);
        const long double tt = (long double)a - 0.5L * (long double)qq;
        const long double s0 = sinl(PI * tt), c0 = cosl(PI * tt);
        long double ws = (qq & 1) ? c0 : s0, wc = (qq & 1) ? s0 : c0;
        if (qq & 2) ws = -ws;
        if ((qq + 1) & 2) wc = -wc;
        const double es = ulp_err(s, ws), ec = ulp_err(c, wc);
        ms = fmax(ms, es); mc = fmax(mc, ec); as += es;
        // sqrt and reciprocal on the launch / scattering ranges
        const double r = ldexp((double)(rng() >> 11) * (1.0 / 9007199254740992.0) + 1e-12, (int)(rng() % 60) - 50);
        mq = fmax(mq, ulp_err(tamc::fm::sqrt_normal(r), sqrtl((long double)r)));
        mr = fmax(mr, ulp_err(tamc::fm::rcp_normal(r), 1.0L / (long double)r));
    }
    printf("{\"n\": %ld, \"neglog_max_ulp\": %.3f, \"neglog_mean_ulp\": %.3f, \"sinpi_max_ulp\": %.3f, \"cospi_max_ulp\": %.3f, "
           "\"sinpi_mean_ulp\": %.3f, \"sqrt_max_ulp\": %.3f, \"rcp_max_ulp\": %.3f}\n",
           n, ml, al / n, ms, mc, as / n, mq, mr);
    return 0;
}
