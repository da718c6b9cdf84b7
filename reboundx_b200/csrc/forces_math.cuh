// forces_math.cuh -- per-particle FP64 arithmetic of the REBOUNDx force effects, for sm_100a.
//
// Every function here evaluates ONE particle against ONE body record and is written in the
// association order of the reference's C expressions, so that with
// -fmad=false the +,-,*,/ and sqrt results are bit-identical to the reference built without
// FMA contraction.  libdevice pow/atan/sin/cos/acos differ from glibc by <= 1-2 ulp.
//
// Compiled only into kernels.cu.  No host fallbacks exist: these are __device__ only.
#pragma once
#include <cstdint>
#include <cmath>
#include "rebx_b200.h"
#include "pass_common.cuh"
#include "exact_math.cuh"

#ifndef REBXB_YARK_TRIG_IDENTITY
#define REBXB_YARK_TRIG_IDENTITY 1
#endif

namespace rebxb {

constexpr double kPi     = 3.14159265358979323846;
constexpr double kTwoPi  = 2*kPi;
constexpr double kPi5    = kPi*kPi*kPi*kPi*kPi;          // M_PI*M_PI*M_PI*M_PI*M_PI, left-assoc (yarkovsky_effect.c:156)
constexpr double kTiny   = 1.E-308;

// ---------------------------------------------------------------------------------------
// Arithmetic policies.  The light effects (radiation, J2/J4, gr_potential) are written once against
// a policy M that supplies reciprocal / division / square root:
//   IeeeMath  the hardware operators (each ends in a conditional slow-path call);
//   FastMath  the straight-line, correctly rounded sequences of exact_math.cuh, which flag operands
//             outside their domain in `bad` instead of branching.  Bit-identical results; the caller
//             redoes a flagged particle with IeeeMath.
// ---------------------------------------------------------------------------------------
struct IeeeMath {
    static constexpr bool straight_line = false;
    struct R { double b; };
    static __device__ __forceinline__ R rcp(double b, unsigned&) { return R{b}; }
    static __device__ __forceinline__ double div(double a, const R& r, unsigned&) { return a/r.b; }
    static __device__ __forceinline__ double sqrt_(double x, unsigned&) { return sqrt(x); }
};
struct FastMath {
    static constexpr bool straight_line = true;
    using R = ExactRecip;
    static __device__ __forceinline__ R rcp(double b, unsigned& bad) { return exact_rcp(b, bad); }
    static __device__ __forceinline__ double div(double a, const R& r, unsigned& bad) { return exact_div(a, r, bad); }
    static __device__ __forceinline__ double sqrt_(double x, unsigned& bad) { return exact_sqrt(x, bad); }
};

// Per-thread constants of one effect entry, computed once before the sweep loop (all of
// them particle-independent sub-expressions of the reference formulas, same association).
struct Pre { ExactRecip rc; unsigned rc_bad; double c, k0, k1; };

template <int K>
__device__ __forceinline__ Pre prepare(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body) {
    Pre pre;
    pre.rc.b = 1.0; pre.rc.y = 1.0; pre.rc_bad = 0u; pre.c = 1.0; pre.k0 = 0.0; pre.k1 = 0.0;
    const double bm = org.m;
    if constexpr (K == REBX_B200_RADIATION) {
        pre.c = fx.scal[1];
        pre.rc = exact_rcp(pre.c, pre.rc_bad);           // RN(1/c), shared by the four divisions by c
        pre.k0 = fx.scal[0]*bm;                          // mu = G*m_source           (radiation_forces.c:71)
    } else if constexpr (K == REBX_B200_HARMONICS) {
        const double G = fx.scal[0], J2 = body[REBX_B200_BODY_J2], J4 = body[REBX_B200_BODY_J4], Req = body[REBX_B200_BODY_REQ];
        pre.k0 = 1.5*G*bm*J2*Req*Req;                    // numerator of f1 in j2_func (gravitational_harmonics.c:77)
        pre.k1 = 0.625*G*bm*J4*Req*Req*Req*Req;          // numerator of f1 in j4_func (:93)
    } else if constexpr (K == REBX_B200_GR_POTENTIAL) {
        const double Gm = fx.scal[0]*bm;
        pre.k0 = 6.*Gm*Gm/fx.scal[1];                    // prefac1                    (gr_potential.c:62)
    } else if constexpr (K == REBX_B200_YARKOVSKY) {
        pre.c = fx.scal[2];
    }
    return pre;
}

// m_particle / m_body.  For a massless test particle and a body of positive mass the quotient is
// exactly +0, which is what the division would return.
__device__ __forceinline__ double mass_ratio(double m, double bm) {
    return (m == 0.0 && bm > 0.0) ? 0.0 : m/bm;
}

// ---------------------------------------------------------------------------------------
// Two-body elements relative to `body` -- restates REBOUND's reb_orbit_from_particle_err
// (third-party, not under /root/reference; same restatement as rebound_shim.c).
// Only the members the force path reads: a, e, inc, P.
// ---------------------------------------------------------------------------------------
struct Orbit { double a, e, inc, P; };

template <bool WANT_E, bool WANT_INC, bool WANT_P>
__device__ __forceinline__ Orbit orbit_from_particle(double G, const Particle& p, const Origin& org, const double* __restrict__ body) {
    Orbit o;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    o.a = nan; o.e = nan; o.inc = nan; o.P = nan;
    const double bm = org.m;
    if (bm <= kTiny) return o;                          // primary has no mass
    const double mu = G*(p.m + bm);
    const double dx = p.x - org.x;
    const double dy = p.y - org.y;
    const double dz = p.z - org.z;
    const double dvx = p.vx - org.vx;
    const double dvy = p.vy - org.vy;
    const double dvz = p.vz - org.vz;
    const double d = sqrt(dx*dx + dy*dy + dz*dz);
    const double vsquared = dvx*dvx + dvy*dvy + dvz*dvz;
    const double vcircsquared = mu/d;
    const double a = -mu/(vsquared - 2.*vcircsquared);
    if (d <= kTiny) return o;                           // particle on top of primary
    o.a = a;
    if (WANT_E) {
        const double vdiffsquared = vsquared - vcircsquared;
        const double vr = (dx*dvx + dy*dvy + dz*dvz)/d;
        const double rvr = d*vr;
        const double muinv = 1./mu;
        const double ex = muinv*(vdiffsquared*dx - rvr*dvx);
        const double ey = muinv*(vdiffsquared*dy - rvr*dvy);
        const double ez = muinv*(vdiffsquared*dz - rvr*dvz);
        o.e = sqrt(ex*ex + ey*ey + ez*ez);
    }
    if (WANT_P) {
        const double n = a/fabs(a)*sqrt(fabs(mu/(a*a*a)));
        o.P = kTwoPi/n;
    }
    if (WANT_INC) {
        const double hx = (dy*dvz - dz*dvy);
        const double hy = (dz*dvx - dx*dvz);
        const double hz = (dx*dvy - dy*dvx);
        const double h = sqrt(hx*hx + hy*hy + hz*hz);
        const double cosine = hz/h;
        double val;
        if (cosine > -1. && cosine < 1.) val = acos(cosine);
        else                             val = (cosine <= -1.) ? kPi : 0.;
        o.inc = val;
    }
    return o;
}

// src/inner_disk_edge.c:61-77
__device__ __forceinline__ double planet_trap(double r, double dedge, double hedge) {
    if (r > dedge*(1.0 + hedge)) return 1.0;
    if (dedge*(1.0 - hedge) < r) return 5.5*cos(((dedge*(1.0 + hedge) - r)*2*kPi)/(4*hedge*dedge)) - 4.5;
    return -10.0;
}

// ---------------------------------------------------------------------------------------
// Effects.  Each returns through `a` (running acceleration of the particle, already
// holding gravity + earlier effects) and `red` (signed back-reaction sum onto the body).
// `gi` is the particle's global index, `k` its index inside the shard's tables.
// ---------------------------------------------------------------------------------------

// Division by the speed of light: FastMath reuses the per-thread reciprocal of c.
template <class M> __device__ __forceinline__ double div_c(double a, const Pre& pre, unsigned& bad);
template <> __device__ __forceinline__ double div_c<IeeeMath>(double a, const Pre& pre, unsigned&) { return a/pre.c; }
template <> __device__ __forceinline__ double div_c<FastMath>(double a, const Pre& pre, unsigned& bad) { bad |= pre.rc_bad; return exact_div(a, pre.rc, bad); }

// src/radiation_forces.c:69-98
template <class M>
__device__ __forceinline__ void radiation(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body, const Pre& pre,
                                          const Particle& p, double beta, Vec3& a, unsigned& bad) {
    const bool absent = is_absent(beta);
    if constexpr (M::straight_line) beta = absent ? 0.0 : beta;     // keep the marker NaN out of the arithmetic; result discarded below
    else if (absent) return;                                        // a particle without beta feels nothing (radiation_forces.c:78)
    const double mu = pre.k0;
    const double dx = p.x - org.x;
    const double dy = p.y - org.y;
    const double dz = p.z - org.z;
    const double dr = M::sqrt_(dx*dx + dy*dy + dz*dz, bad);
    const typename M::R rdr = M::rcp(dr, bad);
    const double dvx = p.vx - org.vx;
    const double dvy = p.vy - org.vy;
    const double dvz = p.vz - org.vz;
    const double rdot = M::div(dx*dvx + dy*dvy + dz*dvz, rdr, bad);
    const double a_rad = M::div(beta*mu, M::rcp(dr*dr, bad), bad);
    const double k1 = 1. - div_c<M>(rdot, pre, bad);
    const double fx_ = a_rad*(M::div(k1*dx, rdr, bad) - div_c<M>(dvx, pre, bad));
    const double fy_ = a_rad*(M::div(k1*dy, rdr, bad) - div_c<M>(dvy, pre, bad));
    const double fz_ = a_rad*(M::div(k1*dz, rdr, bad) - div_c<M>(dvz, pre, bad));
    if constexpr (M::straight_line) {       // select instead of branching
        a.x = absent ? a.x : a.x + fx_;
        a.y = absent ? a.y : a.y + fy_;
        a.z = absent ? a.z : a.z + fz_;
    } else {
        a.x += fx_;  a.y += fy_;  a.z += fz_;
    }
}

// src/gravitational_harmonics.c:184-224 with j2_func :72-86 and j4_func :88-102
template <class M>
__device__ __forceinline__ void harmonics(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body, const Pre& pre,
                                          const Particle& p, Vec3& a, Vec3& red, unsigned& bad) {
    const double bm  = org.m;
    const double J4  = body[REBX_B200_BODY_J4];
    const double ux = body[REBX_B200_BODY_U+0], uy = body[REBX_B200_BODY_U+1], uz = body[REBX_B200_BODY_U+2];
    const double vx = body[REBX_B200_BODY_V+0], vy = body[REBX_B200_BODY_V+1], vz = body[REBX_B200_BODY_V+2];
    const double wx = body[REBX_B200_BODY_W+0], wy = body[REBX_B200_BODY_W+1], wz = body[REBX_B200_BODY_W+2];

    const double dx = p.x - org.x;
    const double dy = p.y - org.y;
    const double dz = p.z - org.z;
    const double r2 = dx*dx + dy*dy + dz*dz;
    const double r  = M::sqrt_(r2, bad);
    const typename M::R rr2 = M::rcp(r2, bad), rr = M::rcp(r, bad);
    const double du = ux*dx + uy*dy + uz*dz;
    const double dv = vx*dx + vy*dy + vz*dz;
    const double dw = wx*dx + wy*dy + wz*dz;
    const double costheta  = M::div(dw, rr, bad);
    const double costheta2 = costheta*costheta;

    double au = 0.0, av = 0.0, aw = 0.0;
    {   // J2 != 0 is guaranteed by the table builder (body skipped otherwise)
        const double f1 = M::div(M::div(M::div(pre.k0, rr2, bad), rr2, bad), rr, bad);      // 3/2 G m J2 R^2 /r2/r2/r
        const double f2 = 5.0*costheta2 - 1.0;
        const double f3 = f2 - 2.0;
        au += f1*f2*du;
        av += f1*f2*dv;
        aw += f1*f3*dw;
    }
    if (J4 != 0.0) {    // same for every thread of the launch (a property of the oblate body)
        const double f1 = M::div(M::div(M::div(M::div(pre.k1, rr2, bad), rr2, bad), rr2, bad), rr, bad);   // 5/8 G m J4 R^4 /r2/r2/r2/r
        const double f2 = 63.0*costheta2*costheta2 - 42.0*costheta2 + 3.0;
        const double f3 = f2 - 28.0*costheta2 + 12.0;
        au += f1*f2*du;
        av += f1*f2*dv;
        aw += f1*f3*dw;
    }
    const double ax = ux*au + vx*av + wx*aw;
    const double ay = uy*au + vy*av + wy*aw;
    const double az = uz*au + vz*av + wz*aw;
    a.x += ax;  a.y += ay;  a.z += az;
    const double fac = mass_ratio(p.m, bm);
    red.x -= fac*ax;  red.y -= fac*ay;  red.z -= fac*az;
}

// src/gr_potential.c:60-78
template <class M>
__device__ __forceinline__ void gr_potential(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body, const Pre& pre,
                                             const Particle& p, Vec3& a, Vec3& red, unsigned& bad) {
    const double prefac1 = pre.k0;
    const double dx = p.x - org.x;
    const double dy = p.y - org.y;
    const double dz = p.z - org.z;
    const double r2 = dx*dx + dy*dy + dz*dz;
    const double prefac = M::div(prefac1, M::rcp(r2*r2, bad), bad);
    a.x -= prefac*dx;  a.y -= prefac*dy;  a.z -= prefac*dz;
    const double mr = mass_ratio(p.m, org.m);
    red.x += mr*prefac*dx;  red.y += mr*prefac*dy;  red.z += mr*prefac*dz;
}

// src/yarkovsky_effect.c:76-216.  `code` is the table builder's ye_flag code:
// -1 / 0 / +1 as in the reference, 2 = particle not eligible (a required parameter is
// missing, or a flag value the reference leaves undefined).
struct YarkParams { double density, rot_period, Gamma, albedo, emissivity, k, sx, sy, sz; };

__device__ __forceinline__ void yarkovsky(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body, const Pre& pre,
                                          const Particle& p, const YarkParams& yp, int code, Vec3& a) {
    if (code == 2 || p.r == 0) return;
    const double G = fx.scal[0], lstar = fx.scal[1], c = fx.scal[2], stef_boltz = fx.scal[3];
    const double radius = p.r;
    const double q_yar = 1.0 - yp.albedo;
    const double dx = p.x - org.x;
    const double dy = p.y - org.y;
    const double dz = p.z - org.z;
    const double dvx = p.vx - org.vx;
    const double dvy = p.vy - org.vy;
    const double dvz = p.vz - org.vz;
    const double distance = sqrt((dx*dx) + (dy*dy) + (dz*dz));
    const double rdotv = ((dx*dvx) + (dy*dvy) + (dz*dvz))/(c*distance);
    const double i0 = ((1 - rdotv)*(dx/distance)) - (dvx/c);
    const double i1 = ((1 - rdotv)*(dy/distance)) - (dvy/c);
    const double i2 = ((1 - rdotv)*(dz/distance)) - (dvz/c);

    double mag;
    double Y[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
    if (code != 0) {
        mag = (3*q_yar*lstar)/(64*kPi*radius*yp.density*c*distance*distance);
        if (code == 1) Y[1][0] = 1.0; else Y[0][1] = 1.0;
    } else {
        const Orbit o = orbit_from_particle<false, false, true>(G, p, org, body);
        mag = (3*yp.k*q_yar*lstar)/(16*kPi*radius*yp.density*c*distance*distance);
        const double sx = yp.sx, sy = yp.sy, sz = yp.sz;
        const double Smag = sqrt((sx*sx) + sy*sy + sz*sz);
        const double hx = (dy*dvz) - (dz*dvy);
        const double hy = (dz*dvx) - (dx*dvz);
        const double hz = (dx*dvy) - (dy*dvx);
        const double Hmag = sqrt((hx*hx) + (hy*hy) + (hz*hz));
        const double inv_smag = 1.0/Smag;
        const double inv_mag_sqrd = 1.0/(Smag*Smag);
        const double inv_hmag = 1.0/Hmag;
        const double inv_hmag_sqrd = 1.0/(Hmag*Hmag);
        const double R1s[3][3] = {{0.0, -sz*inv_smag, sy*inv_smag}, {sz*inv_smag, 0.0, -sx*inv_smag}, {-sy*inv_smag, sx*inv_smag, 0.0}};
        const double R2s[3][3] = {{sx*sx*inv_mag_sqrd, sx*sy*inv_mag_sqrd, sx*sz*inv_mag_sqrd},
                                  {sx*sy*inv_mag_sqrd, sy*sy*inv_mag_sqrd, sy*sz*inv_mag_sqrd},
                                  {sx*sz*inv_mag_sqrd, sy*sz*inv_mag_sqrd, sz*sz*inv_mag_sqrd}};
        const double R1h[3][3] = {{0.0, -hz*inv_hmag, hy*inv_hmag}, {hz*inv_hmag, 0.0, -hx*inv_hmag}, {-hy*inv_hmag, hx*inv_hmag, 0.0}};
        const double R2h[3][3] = {{hx*hx*inv_hmag_sqrd, hx*hy*inv_hmag_sqrd, hx*hz*inv_hmag_sqrd},
                                  {hx*hy*inv_hmag_sqrd, hy*hy*inv_hmag_sqrd, hy*hz*inv_hmag_sqrd},
                                  {hx*hz*inv_hmag_sqrd, hy*hz*inv_hmag_sqrd, hz*hz*inv_hmag_sqrd}};
#if REBXB_YARK_TRIG_IDENTITY
        // x^(1/4) = sqrt(sqrt(x)) and y^(3/4) = sqrt(y) sqrt(sqrt(y)): correctly rounded square roots
        // (<= 1.5 ulp in all) instead of two libdevice pow calls (1-2 ulp, ~250 FP64 instructions)
        const double pre  = .5*sqrt(sqrt((stef_boltz*yp.emissivity)/kPi5));
        const double fsq  = sqrt((lstar*q_yar)/(distance*distance));
        const double flux = fsq*sqrt(fsq);
#else
        const double pre  = .5*pow((stef_boltz*yp.emissivity)/kPi5, .25);
        const double flux = pow((lstar*q_yar)/(distance*distance), .75);
#endif
        const double tanPhi     = 1.0/(1.0 + pre*sqrt(yp.rot_period/(yp.Gamma*yp.Gamma))*flux);
        const double tanEpsilon = 1.0/(1.0 + pre*sqrt(o.P/(yp.Gamma*yp.Gamma))*flux);
#if REBXB_YARK_TRIG_IDENTITY
        // cos(atan t) = 1/sqrt(1+t^2), sin(atan t) = t/sqrt(1+t^2): three correctly rounded operations
        // (<= 1.5 ulp) instead of libdevice atan + sin + cos (each 1-2 ulp, ~150 FP64 instructions together).
        // The reference composes glibc's atan/cos/sin (yarkovsky_effect.c:160-166); neither form is
        // bit-reproducible against it, both agree with it to a few ulp.
        const double inv_phi = 1.0/sqrt(1.0 + tanPhi*tanPhi), inv_eps = 1.0/sqrt(1.0 + tanEpsilon*tanEpsilon);
        const double cos_phi = inv_phi, sin_phi = tanPhi*inv_phi;
        const double cos_epsilon = inv_eps, sin_epsilon = tanEpsilon*inv_eps;
#else
        const double Phi = atan(tanPhi);
        const double Epsilon = atan(tanEpsilon);
        const double cos_phi = cos(Phi), sin_phi = sin(Phi);
        const double cos_epsilon = cos(Epsilon), sin_epsilon = sin(Epsilon);
#endif
        double Rys[3][3], Ryh[3][3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double unit = (i == j) ? 1.0 : 0.0;
                Rys[i][j] = (cos_phi*unit) + (sin_phi*R1s[i][j]) + ((1.0 - cos_phi)*R2s[i][j]);
                Ryh[i][j] = (cos_epsilon*unit) - (sin_epsilon*R1h[i][j]) + ((1 - cos_epsilon)*R2h[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                Y[i][j] = (Rys[i][0]*Ryh[0][j]) + (Rys[i][1]*Ryh[1][j]) + (Rys[i][2]*Ryh[2][j]);
            }
        }
    }
    const double d0 = (Y[0][0]*i0) + (Y[0][1]*i1) + (Y[0][2]*i2);
    const double d1 = (Y[1][0]*i0) + (Y[1][1]*i1) + (Y[1][2]*i2);
    const double d2 = (Y[2][0]*i0) + (Y[2][1]*i1) + (Y[2][2]*i2);
    a.x += mag*d0;  a.y += mag*d1;  a.z += mag*d2;
}

// Shared tail of rebx_com_force for one particle (src/rebxtools.c:101-141):
// PARTICLE coordinates  : p.a += f ; p.a -= mr f ; body.a -= mr f   with mr = m/(M+m)
// BARYCENTRIC           : p.a += f ;                all.a  -= mr f   with mr = m/M  (second sweep)
__device__ __forceinline__ void com_force_tail(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body,
                                               const Particle& p, const Vec3& f, Vec3& a, Vec3& red) {
    a.x += f.x;  a.y += f.y;  a.z += f.z;
    const double M = org.m;
    if (fx.flags & REBX_B200_FLAG_BARYCENTRIC) {
        const double mr = p.m/M;
        red.x -= mr*f.x;  red.y -= mr*f.y;  red.z -= mr*f.z;
    } else {
        const double mr = p.m/(M + p.m);
        a.x -= mr*f.x;  a.y -= mr*f.y;  a.z -= mr*f.z;
        red.x -= mr*f.x;  red.y -= mr*f.y;  red.z -= mr*f.z;
    }
}

struct Rel { double dx, dy, dz, dvx, dvy, dvz, r2; };
__device__ __forceinline__ Rel relative(const Particle& p, const Origin& org, const double* __restrict__ body) {
    Rel q;
    q.dvx = p.vx - org.vx;
    q.dvy = p.vy - org.vy;
    q.dvz = p.vz - org.vz;
    q.dx = p.x - org.x;
    q.dy = p.y - org.y;
    q.dz = p.z - org.z;
    q.r2 = q.dx*q.dx + q.dy*q.dy + q.dz*q.dz;
    return q;
}

// src/modify_orbits_forces.c:77-128
__device__ __forceinline__ Vec3 modify_orbits(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body,
                                              const Particle& p, double tau_a_tab, double tau_e_tab, double tau_inc_tab,
                                              const Orbit* so = nullptr) {
    double invtau_a = 0.0;
    double tau_e = INFINITY, tau_inc = INFINITY;
    const Rel q = relative(p, org, body);
    if (!is_absent(tau_a_tab)) {
        invtau_a = 1.0/tau_a_tab;
        if (fx.flags & REBX_B200_FLAG_TRAP) {
            const Orbit o = so ? *so : orbit_from_particle<false, false, false>(fx.scal[0], p, org, body);
            invtau_a *= planet_trap(o.a, fx.scal[1], fx.scal[2]);
        }
    }
    if (!is_absent(tau_e_tab))   tau_e = tau_e_tab;
    if (!is_absent(tau_inc_tab)) tau_inc = tau_inc_tab;
    Vec3 f;
    f.x = q.dvx*invtau_a/(2.);
    f.y = q.dvy*invtau_a/(2.);
    f.z = q.dvz*invtau_a/(2.);
    if (tau_e < INFINITY || tau_inc < INFINITY) {
        const double vdotr = q.dx*q.dvx + q.dy*q.dvy + q.dz*q.dvz;
        const double prefac = 2*vdotr/q.r2/tau_e;
        f.x += prefac*q.dx;
        f.y += prefac*q.dy;
        f.z += prefac*q.dz + 2.*q.dvz/tau_inc;
    }
    return f;
}

// src/gas_damping_timescale.c:67-130.  d_factor absent => zero force (the host raises the error).
__device__ __forceinline__ Vec3 gas_damping(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body,
                                            const Particle& p, double d_factor, const Orbit* so = nullptr) {
    Vec3 f = {0.0, 0.0, 0.0};
    if (is_absent(d_factor)) return f;
    const double G = fx.scal[0], cs_coeff = fx.scal[1], tau_coeff = fx.scal[2];
    const Orbit o = so ? *so : orbit_from_particle<true, true, false>(G, p, org, body);
    const Rel q = relative(p, org, body);
    const double a0 = o.a, e0 = o.e, inc0 = o.inc;
    const double starMass = org.m;
    const double planetMass = p.m;
    double coeff;
    const double vk = sqrt(G*starMass/a0);
    const double v = sqrt(e0*e0 + inc0*inc0)*vk;
    const double cs = cs_coeff/sqrt(sqrt(a0));
    const double v_over_cs = v/cs;
    if (v <= cs) {
        coeff = 1.;
    } else if (inc0 < cs/vk) {
        coeff = v_over_cs*v_over_cs*v_over_cs;
    } else {
        coeff = v_over_cs*v_over_cs*v_over_cs*v_over_cs;
    }
    const double tau_e = -tau_coeff*d_factor*a0*a0*(starMass/planetMass)*coeff;
    const double tau_inc = 2.*tau_e;
    if (tau_e < INFINITY || tau_inc < INFINITY) {
        const double vdotr = q.dx*q.dvx + q.dy*q.dvy + q.dz*q.dvz;
        const double prefac = 2*vdotr/q.r2/tau_e;
        f.x += prefac*q.dx;
        f.y += prefac*q.dy;
        f.z += prefac*q.dz + 2.*q.dvz/tau_inc;
    }
    return f;
}

// x^p for an exponent that is a FORCE-LEVEL scalar (uniform over the sweep, so the branch never diverges):
// the exponents disc models actually use are served by correctly rounded operators -- 1/x, sqrt chains
// (each further sqrt halves the relative error carried in: <= 1 ulp in all), x*sqrt(x) -- instead of
// libdevice pow (1-2 ulp, ~150 FP64 instructions); any other exponent goes to pow as before.
__device__ __forceinline__ double pow_uniform(double x, double p) {
    if (p == 0.0)    return 1.0;
    if (p == 1.0)    return x;
    if (p == -1.0)   return 1.0/x;
    if (p == 0.5)    return sqrt(x);
    if (p == -0.5)   return 1.0/sqrt(x);
    if (p == 0.25)   return sqrt(sqrt(x));
    if (p == 0.125)  return sqrt(sqrt(sqrt(x)));
    if (p == 1.5)    return x*sqrt(x);
    if (p == -1.5)   return 1.0/(x*sqrt(x));
    if (p == 2.0)    return x*x;
    if (p == -2.0)   return 1.0/(x*x);
    return pow(x, p);
}

// src/type_I_migration.c:69-113 (time-scale fits) and :115-203
__device__ __forceinline__ Vec3 type_I_migration(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body,
                                                 const Particle& p, const Orbit* so = nullptr) {
    const double G = fx.scal[0], beta = fx.scal[1], h0 = fx.scal[2], sd0 = fx.scal[3], s = fx.scal[4];
    const double dedge = fx.scal[5], hedge = fx.scal[6];
    const Orbit o = so ? *so : orbit_from_particle<true, true, false>(G, p, org, body);
    const double a0 = o.a, e0 = o.e, inc0 = o.inc;
    const double mp = p.m, ms = org.m;
    const Rel q = relative(p, org, body);

    const double h  = h0*pow_uniform(q.r2, beta/2);
    const double h2 = h*h;
    const double eh = e0/h;
    const double ih = inc0/h;

    const double sd   = sd0*pow_uniform(sqrt(q.r2), -s);
    const double wave = (sqrt(ms*ms*ms)*h2*h2)/(mp*sd*sqrt(a0*G));

    const double term  = (eh/2.25);
    const double term2 = (eh/2.84)*(eh/2.84);
    const double term3 = (eh/2.02)*(eh/2.02);
    const double Pe = (1. + pow(term, 1.2) + term2*term2*term2)/(1. - term3*term3);
    const double t_mig = ((2.*wave)/(2.7 + 1.1*s))*(1/h2)*(Pe + (Pe/fabs(Pe))*((0.070*ih) + (0.085*ih*ih*ih*ih) - (0.080*eh*ih*ih)));
    const double invtau_mig = planet_trap(a0, dedge, hedge)/t_mig;
    const double tau_e   = (wave/0.780)*(1. - (0.14*eh*eh) + (0.06*eh*eh*eh) + (0.18*eh*ih*ih));
    const double tau_inc = (wave/0.544)*(1 - (0.30*ih*ih) + (0.24*ih*ih*ih) + (0.14*eh*eh*ih));

    Vec3 f = {0.0, 0.0, 0.0};
    if (invtau_mig != 0.0) {
        f.x = -q.dvx*invtau_mig;
        f.y = -q.dvy*invtau_mig;
        f.z = -q.dvz*invtau_mig;
    }
    if (tau_e < INFINITY || tau_inc < INFINITY) {
        const double vdotr = q.dx*q.dvx + q.dy*q.dvy + q.dz*q.dvz;
        const double prefac = -2*vdotr/q.r2/tau_e;
        f.x += prefac*q.dx;
        f.y += prefac*q.dy;
        f.z += prefac*q.dz - 2*q.dvz/tau_inc;
    }
    return f;
}

// src/central_force.c:63-85: a += A r^(gamma-1) d ; the central body recoils by m/m_source of it
__device__ __forceinline__ void central_force(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body,
                                              const Particle& p, Vec3& a, Vec3& red) {
    const double A = body[REBX_B200_BODY_J2], gamma = body[REBX_B200_BODY_J4];     // slots 8, 9 of the record
    const double dx = p.x - org.x;
    const double dy = p.y - org.y;
    const double dz = p.z - org.z;
    const double r2 = dx*dx + dy*dy + dz*dz;
    const double prefac = A*pow_uniform(r2, (gamma - 1.)/2.);
    a.x += prefac*dx;  a.y += prefac*dy;  a.z += prefac*dz;
    const double mr = p.m/org.m;
    red.x -= mr*prefac*dx;  red.y -= mr*prefac*dy;  red.z -= mr*prefac*dz;
}

// src/lense_thirring.c:65-100: gravitomagnetic acceleration 2 Omega_LT x v from the spin of particle 0
__device__ __forceinline__ void lense_thirring(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body,
                                               const Particle& p, Vec3& a, Vec3& red) {
    const double G = fx.scal[0], C2 = fx.scal[1];
    const double gamma = 1.000021;                               // Eddington-Robertson-Schiff parameter, hard-coded in the reference
    const double dx = p.x - org.x;
    const double dy = p.y - org.y;
    const double dz = p.z - org.z;
    const double r2 = dx*dx + dy*dy + dz*dz;
    const double r = sqrt(r2);
    const double r3 = r2*r;
    const double dvx = p.vx - org.vx;
    const double dvy = p.vy - org.vy;
    const double dvz = p.vz - org.vz;
    const double Jx = body[8], Jy = body[9], Jz = body[10];
    const double ms = org.m;
    const double mt = p.m;
    const double mtot = ms + mt;
    const double mratio = mt/ms;
    const double Omega_fac = ms/mtot*(1. + gamma)*G/2/C2/r3;
    const double Jdotr = Jx*dx + Jy*dy + Jz*dz;
    const double Omega_x = Omega_fac*(-Jx + 3.*Jdotr*dx/r2);
    const double Omega_y = Omega_fac*(-Jy + 3.*Jdotr*dy/r2);
    const double Omega_z = Omega_fac*(-Jz + 3.*Jdotr*dz/r2);
    a.x += 2.*(Omega_y*dvz - Omega_z*dvy);
    a.y += 2.*(Omega_z*dvx - Omega_x*dvz);
    a.z += 2.*(Omega_x*dvy - Omega_y*dvx);
    red.x -= mratio*2.*(Omega_y*dvz - Omega_z*dvy);
    red.y -= mratio*2.*(Omega_z*dvx - Omega_x*dvz);
    red.z -= mratio*2.*(Omega_x*dvy - Omega_y*dvx);
}

// src/gas_dynamical_friction.c:66-137: Ostriker dynamical friction + geometric drag in a gaseous disc
// around particle 0 (no back-reaction)
__device__ __forceinline__ void gas_dynamical_friction(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body,
                                                       const Particle& p, Vec3& a) {
    const double G = fx.scal[0], rhog = fx.scal[1], alpha_rhog = fx.scal[2], cs = fx.scal[3], alpha_cs = fx.scal[4];
    const double xmin = fx.scal[5], hr = fx.scal[6], Qd = fx.scal[7];
    const double dx = p.x - org.x, dy = p.y - org.y, dz = p.z - org.z;
    const double dvx = p.vx - org.vx, dvy = p.vy - org.vy, dvz = p.vz - org.vz;
    const double rcyl = sqrt(dx*dx + dy*dy);
    // get_vrel_disk (:90-100): velocity relative to the sub-Keplerian gas
    const double sin_phi = dy/rcyl, cos_phi = dx/rcyl;
    const double vk = sqrt(G*org.m/rcyl)*(1.0 - hr*hr);
    const double v0 = dvx + vk*sin_phi, v1 = dvy - vk*cos_phi, v2 = dvz;
    const double vrel_norm = sqrt(v0*v0 + v1*v1 + v2*v2);
    const double mach = vrel_norm/(cs*pow_uniform(rcyl, alpha_cs));
    // calculate_pre_factor (:79-88) with mach_piece_sub (:67-76)
    const double coul = log(1.0/xmin);
    double integ;
    if (mach >= 1.0) integ = coul;
    else {
        const double sub = (mach < 0.02) ? mach*mach*mach/3. + mach*mach*mach*mach*mach/5. : 0.5*log((1.0 + mach)/(1.0 - mach)) - mach;
        integ = fmin(coul, sub);
    }
    const double h = hr*rcyl;
    const double vert = (fabs(dz) < (10*h)) ? exp(-dz*dz/(2.0*h*h)) : 0;
    const double rhog_loc = rhog*pow_uniform(rcyl, alpha_rhog)*vert;
    const double mp = p.m, rstar = p.r;
    const double fc = 4.*kPi*(G*G)*mp*(rhog_loc)/(vrel_norm*vrel_norm*vrel_norm)*integ + kPi*rhog_loc*rstar*rstar*vrel_norm*Qd/mp;
    a.x -= fc*v0;  a.y -= fc*v1;  a.z -= fc*v2;
}

// src/exponential_migration.c:61-100 (absent parameters: tau = inf, a_ini = 24, a_fin = 30, :65-67)
__device__ __forceinline__ Vec3 exponential_migration(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body,
                                                      const Particle& p, double tau_tab, double aini_tab, double afin_tab) {
    const Orbit o = orbit_from_particle<false, false, false>(fx.scal[0], p, org, body);
    const double em_tau_a = is_absent(tau_tab) ? INFINITY : tau_tab;
    const double em_aini = is_absent(aini_tab) ? 24. : aini_tab;
    const double em_afin = is_absent(afin_tab) ? 30. : afin_tab;
    const double dvx = p.vx - org.vx;
    const double dvy = p.vy - org.vy;
    const double dvz = p.vz - org.vz;
    const double t = fx.scal[1];
    Vec3 f;
    f.x = (dvx/(2.*em_tau_a))*((em_afin - em_aini)/(o.a))*exp(-(t)/em_tau_a);
    f.y = (dvy/(2.*em_tau_a))*((em_afin - em_aini)/(o.a))*exp(-(t)/em_tau_a);
    f.z = (dvz/(2.*em_tau_a))*((em_afin - em_aini)/(o.a))*exp(-(t)/em_tau_a);
    return f;
}

}  // namespace rebxb
