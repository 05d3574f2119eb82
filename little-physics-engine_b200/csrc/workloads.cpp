// workloads.cpp — deterministic synthetic body distributions (host side of liblpe_bh.so).
//
// std::mt19937_64, uniform u = (g() >> 11) * 2^-53 only (no std::*_distribution, which differ
// between standard libraries), bodies created in index order. Masses m_i = 1e6*(0.5+u) unless stated, so that the
// reference's first-occupant double count is visible in the results.
//   kind 0  C2  uniform disk            r = 0.45 U sqrt(u1), phi = 2 pi u2, centre (U/2, U/2), v = 0
//   kind 1  C3  Plummer sphere          r = a / sqrt(u1^(-2/3) - 1), a = U/40, isotropic direction, projected on
//                                       x-y about (U/2, U/2); rejected outside [0.05 U, 0.95 U]^2, v = 0
//   kind 2  C4  two-galaxy collision    two disks of n/2 bodies centred (0.35 U, 0.5 U) and (0.65 U, 0.5 U), radial
//                                       pdf ~ (r_in/r)^(15/8) between r_in = U/60 and r_out = U/6 (the law of
//                                       reference keplerian_disk.cpp:72-75,99-106), a central body of half the disk
//                                       mass created first, circular Kepler speed plus a bulk approach velocity
//   kind 4      Keplerian disk, counter-based (kepler_gen.h): body i from its own stream, identical law; the device
//                                       makes the same bodies without any host array (lpe_bh_generate)
//   kind 3  C1  Keplerian disk          restates reference src/scenarios/keplerian_disk.cpp:45-146 (central mass 1e36,
//                                       density/height/mass power laws, velocity dispersion) with this RNG; the
//                                       reference itself seeds from time() and is not reproducible
#include <cmath>
#include <cstdint>
#include <random>

#include "../../include/lpe_bh.h"
#include "kepler_gen.h"

namespace {

struct Rng {
    std::mt19937_64 g;
    explicit Rng(uint64_t seed) : g(seed) {}
    double u() { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); }
    // Box-Muller on two uniforms: deterministic across standard libraries
    double normal(double mean, double sd) {
        double u1 = u();
        if (u1 < 1e-300) u1 = 1e-300;
        const double u2 = u();
        return mean + sd * std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

constexpr double kTwoPi = 6.283185307179586;
constexpr double kG = 6.674e-11;  // SimulatorConstants::RealG, reference src/core/constants.cpp:8

void uniform_disk(uint64_t n, uint64_t seed, double U, double* x, double* y, double* vx, double* vy, double* m) {
    Rng r(seed);
    for (uint64_t i = 0; i < n; ++i) {
        const double rad = 0.45 * U * std::sqrt(r.u());
        const double phi = kTwoPi * r.u();
        x[i] = 0.5 * U + rad * std::cos(phi);
        y[i] = 0.5 * U + rad * std::sin(phi);
        m[i] = 1e6 * (0.5 + r.u());
        vx[i] = vy[i] = 0.0;
    }
}

void plummer(uint64_t n, uint64_t seed, double U, double* x, double* y, double* vx, double* vy, double* m) {
    Rng r(seed);
    const double a = U / 40.0;
    for (uint64_t i = 0; i < n;) {
        double u1 = r.u();
        if (u1 < 1e-12) u1 = 1e-12;
        const double rad = a / std::sqrt(std::pow(u1, -2.0 / 3.0) - 1.0);
        const double ct = 2.0 * r.u() - 1.0;
        const double phi = kTwoPi * r.u();
        const double st = std::sqrt(1.0 - ct * ct);
        const double px = 0.5 * U + rad * st * std::cos(phi);
        const double py = 0.5 * U + rad * st * std::sin(phi);
        const double mm = 1e6 * (0.5 + r.u());
        if (!(px >= 0.05 * U && px <= 0.95 * U && py >= 0.05 * U && py <= 0.95 * U)) continue;
        x[i] = px; y[i] = py; m[i] = mm; vx[i] = vy[i] = 0.0;
        ++i;
    }
}

void one_galaxy(uint64_t n, uint64_t seed, double U, double cx, double cy, double bulkVx, double* x, double* y,
                double* vx, double* vy, double* m) {
    if (n == 0) return;
    Rng r(seed);
    const double rin = U / 60.0, rout = U / 6.0;
    const double meanMass = 1e6;
    const double central = 0.5 * meanMass * (double)(n - 1);
    x[0] = cx; y[0] = cy; vx[0] = bulkVx; vy[0] = 0.0; m[0] = central;  // created first
    for (uint64_t i = 1; i < n;) {
        const double rad = rin + (rout - rin) * r.u();
        const double t = r.u();
        if (t > std::pow(rin / rad, 15.0 / 8.0)) continue;
        const double phi = kTwoPi * r.u();
        const double mm = meanMass * (0.5 + r.u());
        const double speed = std::sqrt(kG * central / rad);
        x[i] = cx + rad * std::cos(phi);
        y[i] = cy + rad * std::sin(phi);
        vx[i] = -speed * std::sin(phi) + bulkVx;
        vy[i] = speed * std::cos(phi);
        m[i] = mm;
        ++i;
    }
}

void two_galaxies(uint64_t n, uint64_t seed, double U, double* x, double* y, double* vx, double* vy, double* m) {
    const uint64_t n0 = n / 2, n1 = n - n0;
    const double central = 0.5 * 1e6 * (double)(n0 > 0 ? n0 - 1 : 0);
    const double vb = 0.5 * std::sqrt(kG * 2.0 * central / (0.3 * U));  // half the mutual circular speed at separation
    one_galaxy(n0, seed, U, 0.35 * U, 0.5 * U, +vb, x, y, vx, vy, m);
    one_galaxy(n1, seed + 1, U, 0.65 * U, 0.5 * U, -vb, x + n0, y + n0, vx + n0, vy + n0, m + n0);
}

void keplerian(uint64_t n, uint64_t seed, double U, double* x, double* y, double* vx, double* vy, double* m) {
    if (n == 0) return;
    Rng r(seed);
    // KeplerianDiskConfig defaults, reference include/scenarios/keplerian_disk.hpp:17-41
    const double centralMass = 1e36, innerRpix = 100.0, outerFactor = 2.5, heightScale = 20.0, heightPow = 1.25;
    const double densPow = 15.0 / 8.0, massMean = 1e22, massSd = 1e21, massRadPow = 0.5;
    const double velDisp = 0.01, radVel = 0.001;
    const double screen = 600.0;              // SimulatorConstants::ScreenLength, constants.cpp:12
    const double mpp = U / screen;            // MetersPerPixel (1e7 when U = 6e9, keplerian_disk.cpp:16-17)
    const double cx = 0.5 * screen * mpp, cy = 0.5 * screen * mpp;
    x[0] = cx; y[0] = cy; vx[0] = vy[0] = 0.0; m[0] = centralMass;     // createCentralBody, keplerian_disk.cpp:45-53
    const double minRpix = innerRpix, maxRpix = screen / outerFactor, minRm = minRpix * mpp;
    for (uint64_t i = 1; i < n;) {
        const double rpix = minRpix + (maxRpix - minRpix) * r.u();
        const double thresh = r.u();
        if (thresh > std::pow(innerRpix / rpix, densPow)) continue;    // keplerian_disk.cpp:99-106
        const double rm = rpix * mpp;
        const double ang = kTwoPi * r.u();
        const double maxH = (innerRpix / heightScale) * std::pow(rpix / innerRpix, heightPow) * mpp;
        const double hOff = r.normal(0.0, maxH / 3.0);
        const double px = cx + rm * std::cos(ang);
        const double py = cy + rm * std::sin(ang) + hOff;
        const double speed = std::sqrt(kG * centralMass / rm) * r.normal(1.0, velDisp);
        double vxx = -speed * std::sin(ang), vyy = speed * std::cos(ang);
        const double rv = r.normal(0.0, speed * radVel);
        vxx += rv * std::cos(ang);
        vyy += rv * std::sin(ang);
        const double mm = r.normal(std::pow(minRm / rm, massRadPow) * massMean, massSd);
        x[i] = px; y[i] = py; vx[i] = vxx; vy[i] = vyy; m[i] = mm;
        ++i;
    }
}

}  // namespace

extern "C" int lpe_bh_workload(int kind, uint64_t n, uint64_t seed, double U, double* x, double* y, double* vx,
                               double* vy, double* m) {
    if (!x || !y || !vx || !vy || !m || !(U > 0.0)) return 1;
    switch (kind) {
        case 0: uniform_disk(n, seed, U, x, y, vx, vy, m); return 0;
        case 1: plummer(n, seed, U, x, y, vx, vy, m); return 0;
        case 2: two_galaxies(n, seed, U, x, y, vx, vy, m); return 0;
        case 3: keplerian(n, seed, U, x, y, vx, vy, m); return 0;
        case 4:   // the same law, one independent stream per body: what lpe_bh_generate makes on the device
            for (uint64_t i = 0; i < n; ++i) lpe_keplerian_body(i, seed, U, x + i, y + i, vx + i, vy + i, m + i);
            return 0;
        default: return 1;
    }
}
