// kepler_gen.h — the Keplerian-disk scenario's entity law (reference src/scenarios/keplerian_disk.cpp:45-146, defaults
// of include/scenarios/keplerian_disk.hpp:17-41) as a COUNTER-BASED generator: body i draws from its own splitmix64
// stream seeded by (seed, i), so any body can be made anywhere — one thread per body on the device
// (lpe_bh_generate), a plain loop on the host (lpe_bh_workload kind 4) — and the two agree to libm rounding.
// The reference itself seeds std::mt19937 from time() and uses std::normal_distribution, i.e. it is not reproducible
// (SURVEY.md D8); this is SURVEY.md 8(f) N3.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LPE_HD __host__ __device__ inline
#else
#define LPE_HD inline
#endif

struct LpeStream {
    uint64_t s;
    LPE_HD uint64_t next() {   // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    LPE_HD double u() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    LPE_HD double normal(double mean, double sd) {   // Box-Muller on two uniforms
        double u1 = u();
        if (u1 < 1e-300) u1 = 1e-300;
        const double u2 = u();
        return mean + sd * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
};

// body 0 is the central mass (keplerian_disk.cpp:45-53); bodies 1.. are disk particles (:95-146)
LPE_HD void lpe_keplerian_body(uint64_t i, uint64_t seed, double U, double* x, double* y, double* vx, double* vy, double* m) {
    const double kG = 6.674e-11;                // SimulatorConstants::RealG, constants.cpp:8
    const double centralMass = 1e36, innerRpix = 100.0, outerFactor = 2.5, heightScale = 20.0, heightPow = 1.25;
    const double densPow = 15.0 / 8.0, massMean = 1e22, massSd = 1e21, massRadPow = 0.5;
    const double velDisp = 0.01, radVel = 0.001;
    const double screen = 600.0;                // SimulatorConstants::ScreenLength, constants.cpp:12
    const double mpp = U / screen;              // MetersPerPixel (1e7 when U = 6e9, keplerian_disk.cpp:16-17)
    const double cx = 0.5 * screen * mpp, cy = 0.5 * screen * mpp;
    if (i == 0) { *x = cx; *y = cy; *vx = 0.0; *vy = 0.0; *m = centralMass; return; }
    LpeStream r;
    r.s = seed * 0xD1342543DE82EF95ull + i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    r.next();
    const double minRpix = innerRpix, maxRpix = screen / outerFactor, minRm = minRpix * mpp;
    double rpix;
    for (;;) {                                  // density ~ (r_in / r)^(15/8) by rejection, keplerian_disk.cpp:99-106
        rpix = minRpix + (maxRpix - minRpix) * r.u();
        const double thresh = r.u();
        if (!(thresh > pow(innerRpix / rpix, densPow))) break;
    }
    const double rm = rpix * mpp;
    const double ang = 6.283185307179586 * r.u();
    const double maxH = (innerRpix / heightScale) * pow(rpix / innerRpix, heightPow) * mpp;
    const double hOff = r.normal(0.0, maxH / 3.0);
    const double sa = sin(ang), ca = cos(ang);
    const double speed = sqrt(kG * centralMass / rm) * r.normal(1.0, velDisp);
    const double rv = r.normal(0.0, speed * radVel);
    *x = cx + rm * ca;
    *y = cy + rm * sa + hOff;
    *vx = -speed * sa + rv * ca;
    *vy = speed * ca + rv * sa;
    *m = r.normal(pow(minRm / rm, massRadPow) * massMean, massSd);
}
