// Discrete-gamma category rates (Yang 1994), host-only.
//
// C++ restatement of the numerical recipe the reference compiles from PAML
// (/root/reference/src/c_discrete_gamma.c): equiprobable categories of Gamma(alpha, beta),
// represented by the category mean (Yang 1994 eq. 10) or median.  The building blocks are
// classic published algorithms, re-expressed here with structured control flow:
//
//   normal_quantile      Odeh & Evans 1974, AS 70            (ref c_discrete_gamma.c:27-57)
//   ln_gamma_stirling    Pike & Hill 1966, CACM Alg. 291      (ref :59-83)
//   ln_gamma             same, exact for small integers       (ref :85-128)
//   chi2_quantile        Best & Roberts 1975, AS 91           (ref :130-202)
//   incomplete_gamma     Bhattacharjee 1970, AS 32, 1e-8      (ref :204-283)
//   phb_discrete_gamma   Yang 1994 discretisation             (ref :285-321)
//
// The floating-point expressions keep the operation order of the published algorithms so
// the rates agree with the reference's to the last bit (tests/test_gamma.py checks this
// against oracle/_ref, the reference's own C file compiled by oracle/Makefile).
// Build with -ffp-contract=off: a fused multiply-add would change the rounding.
#include <cmath>
#include <cstdint>

#include "../../include/phylo_b200.h"

namespace {

double normal_quantile(double prob) {
    const double a[5] = {-.322232431088, -1.0, -.342242088547, -.0204231210245, -.453642210148e-4};
    const double b[5] = {.0993484626060, .588581570495, .531103462366, .103537752850, .0038560700634};
    const double tail = prob < 0.5 ? prob : 1 - prob;
    double z;
    if (tail < 1e-20) {
        z = 999;
    } else {
        const double y = std::sqrt(std::log(1 / (tail * tail)));
        const double num = (((y * a[4] + a[3]) * y + a[2]) * y + a[1]) * y + a[0];
        const double den = (((y * b[4] + b[3]) * y + b[2]) * y + b[1]) * y + b[0];
        z = y + num / den;
    }
    return prob < 0.5 ? -z : z;
}

// Stirling series correction term c(x) / x shared by both log-gamma flavours
inline double stirling_correction(double x) {
    const double z = 1 / (x * x);
    return (((-.000595238095238 * z + .000793650793651) * z - .002777777777778) * z + .083333333333333) / x;
}

// push x above 7 with the recurrence Gamma(x+1) = x Gamma(x); returns -log(product)
inline double shift_above_seven(double& x) {
    if (!(x < 7)) return 0.0;
    double prod = 1;
    double z = x - 1;
    while (++z < 7) prod *= z;
    x = z;
    return -std::log(prod);
}

double ln_gamma_stirling(double alpha) {
    double x = alpha;
    const double f = shift_above_seven(x);
    // left-to-right sum, as published: the association decides the last bit
    return f + (x - 0.5) * std::log(x) - x + .918938533204673 + stirling_correction(x);
}

double ln_gamma(double x) {
    const int nx = static_cast<int>(x);
    if (static_cast<double>(nx) == x && nx >= 0 && nx <= 11) {
        long fact = 1;
        for (long i = 2; i <= static_cast<long>(nx - 1); ++i) fact *= i;
        return std::log(static_cast<double>(fact));
    }
    double fneg = 0;
    if (x <= 0) {
        if (static_cast<int>(x) - x == 0) return -1;   // pole
        for (fneg = 1; x < 0; x++) fneg /= x;
        if (fneg < 0) return -1;
        fneg = std::log(fneg);
    }
    const double f = shift_above_seven(x);
    return fneg + f + (x - 0.5) * std::log(x) - x + .918938533204673 + stirling_correction(x);
}

double incomplete_gamma(double x, double alpha, double ln_gamma_alpha) {
    const double accurate = 1e-8, overflow = 1e30;
    const double p = alpha;
    if (x == 0) return 0;
    if (x < 0 || p <= 0) return -1;

    const double factor = std::exp(p * std::log(x) - x - ln_gamma_alpha);
    if (!(x > 1 && x >= p)) {
        // series expansion
        double gin = 1, term = 1, rn = p;
        do {
            rn++;
            term *= x / rn;
            gin += term;
        } while (term > accurate);
        return gin * (factor / p);
    }
    // continued fraction
    double a = 1 - p;
    double b = a + x + 1;
    double term = 0;
    double pn[6] = {1, x, x + 1, x * b, 0, 0};
    double gin = pn[2] / pn[3];
    for (;;) {
        a++;
        b += 2;
        term++;
        const double an = a * term;
        pn[4] = b * pn[2] - an * pn[0];
        pn[5] = b * pn[3] - an * pn[1];
        if (pn[5] != 0) {
            const double rn = pn[4] / pn[5];
            const double dif = std::fabs(gin - rn);
            if (dif <= accurate && dif <= accurate * rn) break;
            gin = rn;
        }
        for (int i = 0; i < 4; ++i) pn[i] = pn[i + 2];
        if (!(std::fabs(pn[4]) < overflow))
            for (int i = 0; i < 4; ++i) pn[i] /= overflow;
    }
    return 1 - factor * gin;
}

double chi2_quantile(double prob, double v) {
    const double e = .5e-6, aa = .6931471805, p = prob;
    if (p < .000002 || p > .999998 || v <= 0) return -1;

    const double g = ln_gamma_stirling(v / 2);
    const double xx = v / 2;
    const double c = xx - 1;
    double ch;

    if (v < -1.24 * std::log(p)) {
        // small chi-squared starting value
        ch = std::pow(p * xx * std::exp(g + xx * aa), 1 / xx);
        if (ch - e < 0) return ch;
    } else if (v > .32) {
        // Wilson-Hilferty start
        const double x = normal_quantile(p);
        const double p1 = 0.222222 / v;
        ch = v * std::pow(x * std::sqrt(p1) + 1 - p1, 3.0);
        if (ch > 2.2 * v + 6) ch = -2 * (std::log(1 - p) - c * std::log(.5 * ch) + g);
    } else {
        // v <= 0.32: Newton-type iteration on an approximation
        ch = 0.4;
        const double a = std::log(1 - p);
        double q;
        do {
            q = ch;
            const double p1 = 1 + ch * (4.67 + ch);
            const double p2 = ch * (6.73 + ch * (6.66 + ch));
            const double t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2;
            ch -= (1 - std::exp(a + g + .5 * ch + c * aa) * p2 / p1) / t;
        } while (std::fabs(q / ch - 1) - .01 > 0);
    }

    // seven-term Taylor refinement
    double q;
    do {
        q = ch;
        const double p1 = .5 * ch;
        double t = incomplete_gamma(p1, xx, g);
        if (t < 0) return -1;
        const double p2 = p - t;
        t = p2 * std::exp(xx * aa + g + p1 - c * std::log(ch));
        const double b = t / ch;
        const double a = 0.5 * t - b * c;
        const double s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420;
        const double s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520;
        const double s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520;
        const double s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040;
        const double s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520;
        const double s6 = (120 + c * (346 + 127 * c)) / 5040;
        ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
    } while (std::fabs(q / ch - 1) > e);
    return ch;
}

inline double gamma_quantile(double prob, double alpha, double beta) {
    return chi2_quantile(prob, 2.0 * alpha) / (2.0 * beta);
}

}  // namespace

extern "C" int phb_discrete_gamma(double alpha, double beta, int ncat, int use_median, double* rates,
                                  double* weights) {
    if (ncat < 1 || rates == nullptr || weights == nullptr || !(alpha > 0) || !(beta > 0)) return PHB_ERR_INVALID;
    const int K = ncat;
    const double mean = alpha / beta;
    if (use_median) {
        double total = 0;
        for (int i = 0; i < K; ++i) rates[i] = gamma_quantile((i * 2. + 1) / (2. * K), alpha, beta);
        for (int i = 0; i < K; ++i) total += rates[i];
        for (int i = 0; i < K; ++i) rates[i] *= mean * K / total;
    } else if (K == 1) {
        rates[0] = mean;
    } else {
        const double lnga1 = ln_gamma(alpha + 1);
        double* cut = weights;  // scratch, overwritten with the weights at the end
        for (int i = 0; i < K - 1; ++i) cut[i] = gamma_quantile((i + 1.0) / K, alpha, beta);
        for (int i = 0; i < K - 1; ++i) cut[i] = incomplete_gamma(cut[i] * beta, alpha + 1, lnga1);
        rates[0] = cut[0] * mean * K;
        for (int i = 1; i < K - 1; ++i) rates[i] = (cut[i] - cut[i - 1]) * mean * K;
        rates[K - 1] = (1 - cut[K - 2]) * mean * K;
    }
    for (int i = 0; i < K; ++i) weights[i] = 1.0 / K;
    return PHB_OK;
}
