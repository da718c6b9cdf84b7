/* kepler_core.h -- two-body (Kepler) drift in universal variables, shared verbatim by the device kernel
 * (kernels_kepler.cuh, compiled by nvcc with -fmad=false) and the host integrator harness of the REBOUND
 * stand-in (rebound_shim.c, gcc without FMA contraction), so that both evaluate the same IEEE operations in
 * the same order and a device-resident Wisdom-Holman step reproduces the host one bit for bit.
 *
 * It stands where the reference's rebx_kepler_step (src/steppers.c:79-87) calls REBOUND's WHFast Kepler
 * solver -- REBOUND code that is neither in the reference tree nor in this image; the algorithm here is
 * the textbook one (Danby 1988 / Mikkola & Innanen 1999): solve  r0 G1 + eta G2 + mu G3 = dt  for the
 * universal anomaly X by Newton iteration, G_n = X^n c_n(beta X^2) with Stumpff functions c_n evaluated by
 * argument quartering + series + the duplication formulae, then advance with the f, g functions.
 * Plain C99 / CUDA C++.  Reciprocals of the series constants are literal quotients (folded identically
 * by both compilers), so no division is executed for them. */
#ifndef REBXB_KEPLER_CORE_H
#define REBXB_KEPLER_CORE_H
#include <math.h>

#ifdef __CUDACC__
#define REBXB_HD __host__ __device__ __forceinline__
#else
#define REBXB_HD static inline
#endif

/* c0..c3 of z (c0 = cos sqrt z, c1 = sin sqrt z / sqrt z, c2 = (1 - c0)/z, c3 = (1 - c1)/z for z > 0) */
REBXB_HD void rebxb_stumpff(double z, double* c0, double* c1, double* c2, double* c3) {
    int n = 0;
    while (fabs(z) > 0.1) { z *= 0.25; n++; }
    double C2 = (1. - z*(1. - z*(1. - z*(1. - z*(1. - z*(1. - z*(1./182.))*(1./132.))*(1./90.))*(1./56.))*(1./30.))*(1./12.))*0.5;
    double C3 = (1. - z*(1. - z*(1. - z*(1. - z*(1. - z*(1. - z*(1./210.))*(1./156.))*(1./110.))*(1./72.))*(1./42.))*(1./20.))*(1./6.);
    double C1 = 1. - z*C3, C0 = 1. - z*C2;
    for (; n > 0; n--) {
        C3 = (C2 + C0*C3)*0.25;
        C2 = C1*C1*0.5;
        C1 = C0*C1;
        C0 = 2.*C0*C0 - 1.;
    }
    *c0 = C0; *c1 = C1; *c2 = C2; *c3 = C3;
}

/* Advances the relative state (x, v) along its Kepler orbit with gravitational parameter mu for dt. */
REBXB_HD void rebxb_kepler_drift(double mu, double dt, double* x, double* v) {
    const double r0 = sqrt(x[0]*x[0] + x[1]*x[1] + x[2]*x[2]);
    const double v2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
    const double eta = x[0]*v[0] + x[1]*v[1] + x[2]*v[2];
    const double beta = 2.*mu/r0 - v2;
    /* first guess: the Kepler equation to second order in X (dt = r0 X + eta X^2/2 + ...) */
    const double X1 = dt/r0;
    double X = X1*(1. - 0.5*eta*X1/r0), c0 = 1., c1 = 1., c2 = 0.5, c3 = 1./6., G1 = 0., G2 = 0., G3 = 0., rr = r0;
    double dX = 0.;
    int converged = 0;
    for (int it = 0; it < 60; it++) {
        rebxb_stumpff(beta*X*X, &c0, &c1, &c2, &c3);
        G1 = c1*X; G2 = c2*X*X; G3 = c3*X*X*X;
        const double f = r0*G1 + eta*G2 + mu*G3 - dt;
        rr = r0*c0 + eta*G1 + mu*G2;
        dX = f/rr;
        X -= dX;
        /* Newton converges quadratically: a correction below 1e-9 |X| leaves an error below round-off */
        if (fabs(dX) <= 1e-9*fabs(X)) { converged = 1; break; }
    }
    /* G-functions at the final X = (X + dX) - dX to first order in the last (tiny) correction, using
     * dG3/dX = G2, dG2/dX = G1, dG1/dX = G0 = c0, dG0/dX = -beta G1; the neglected terms are O(dX^2) <= 1e-18
     * relative.  Saves re-evaluating the Stumpff series. */
    if (!converged) {                  /* not reached for the step sizes an integrator uses; stay exact anyway */
        rebxb_stumpff(beta*X*X, &c0, &c1, &c2, &c3);
        G1 = c1*X; G2 = c2*X*X; G3 = c3*X*X*X;
    } else {
        const double G0 = c0;
        G3 = G3 - dX*G2;
        G2 = G2 - dX*G1;
        const double G1n = G1 - dX*G0;
        c0 = G0 + dX*beta*G1;
        G1 = G1n;
    }
    rr = r0*c0 + eta*G1 + mu*G2;
    const double f = 1. - mu*G2/r0, g = dt - mu*G3;
    const double fd = -mu*G1/(r0*rr), gd = 1. - mu*G2/rr;
    for (int c = 0; c < 3; c++) {
        const double xn = f*x[c] + g*v[c], vn = fd*x[c] + gd*v[c];
        x[c] = xn; v[c] = vn;
    }
}
#endif /* REBXB_KEPLER_CORE_H */
