// kernels_tol.cu -- TOLERANCE-MODE sweeps: the same effects as kernels.cu, evaluated for speed instead of for
// bit identity with the reference.  Selected per pass with REBX_B200_PASS_TOLERANCE (dispatcher:
// REBX_B200_MATH=tolerance; device.Shard(math="tolerance")); the default stays the bit-exact build.
//
// What changes against forces_math.cuh (reference src/radiation_forces.c:69-98, gravitational_harmonics.c:72-224,
// gr_potential.c:60-78, yarkovsky_effect.c:76-216, modify_orbits_forces.c:77-128, gas_damping_timescale.c:67-130,
// type_I_migration.c:69-203, rebxtools.c:131-141):
//   * this translation unit is compiled WITH FMA contraction (the exact build: -fmad=false);
//   * one reciprocal square root of r^2 per particle (MUFU seed + two coupled Newton steps, <= 2 ulp) serves every
//     power of the distance -- r^-1, r^-2, r^-4, r^-5, r^-7 by multiplication -- where the reference divides
//     12 times and takes a square root (the exact build repeats those IEEE divisions one for one);
//   * every other quotient is a product with a Newton reciprocal (MUFU seed + two steps, <= 1 ulp), particle-
//     independent quotients are hoisted into per-CTA constants in shared memory;
//   * stacked effects that refer to the same central body share the geometry and the two-body elements;
//   * the two Rodrigues matrices of the full Yarkovsky version are applied as two rotations of the vector instead
//     of being built and multiplied out; pow(x, 1.2) of type_I_migration is exp(1.2 log x).
// Accuracy: a few ulp per operation instead of 0.5, i.e. ~1e-15 relative on a force vector -- two orders inside the
// north-star tolerance (per-component relative error <= 1e-13 on the accelerations REBOUND integrates); measured
// against the reference-compiled oracle in tests/test_gpu_parity.py::test_tolerance_mode_* and reported by
// bench.py next to the bit-exact numbers.
//
// Only stacks whose effects all refer to ONE central body in PARTICLE coordinates have an instantiation here
// (every BASELINE configuration); anything else falls back to the exact kernels.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <type_traits>
#include "rebx_b200.h"
#include "pass_common.cuh"
#include "exact_math.cuh"       // rcp_seed / rsqrt_seed (MUFU.RCP64H / MUFU.RSQ64H)
#include "jacobi_common.cuh"    // CTA-wide scan + mass grid of the centre-of-mass pipeline
#include "jacobi_fused.cuh"     // the whole Jacobi pipeline in one cooperative launch

namespace rebxb {
namespace tol {

constexpr double kPi    = 3.14159265358979323846;
constexpr double kTwoPi = 2*kPi;
constexpr double kPi5   = kPi*kPi*kPi*kPi*kPi;

// 1/b, <= 1 ulp for finite non-zero b away from the exponent limits (seed ~2^-20, two Newton steps)
__device__ __forceinline__ double frcp(double b) {
    double y = rcp_seed(b);
    double e = fma(-b, y, 1.0);  y = fma(y, e, y);
    e = fma(-b, y, 1.0);         y = fma(y, e, y);
    return y;
}
// s = sqrt(x), rs = 1/sqrt(x), <= 2 ulp each (Goldschmidt: two coupled steps on g ~ sqrt x, h ~ 1/(2 sqrt x))
__device__ __forceinline__ void fsqrt2(double x, double& s, double& rs) {
    const double y = rsqrt_seed(x);
    double g = x*y, h = 0.5*y;
    double r = fma(-g, h, 0.5);  g = fma(g, r, g);  h = fma(h, r, h);
    r = fma(-g, h, 0.5);         g = fma(g, r, g);  h = fma(h, r, h);
    s = (x == 0.0) ? 0.0 : g;
    rs = h + h;
}
__device__ __forceinline__ double fsqrt(double x) { double s, rs; fsqrt2(x, s, rs); return s; }
__device__ __forceinline__ double frsqrt(double x) { double s, rs; fsqrt2(x, s, rs); return rs; }

// x^p for a force-level (uniform) exponent, through the cheap roots where disc models put it; s1 = sqrt(x), rs1 = 1/sqrt(x)
// are passed in when the caller has them (x = r: the distance).  Returns x^p and, in `inv`, x^-p.
__device__ __forceinline__ double pow_uniform2(double x, double xinv, double p, double& inv) {
    if (p == 0.0) { inv = 1.0; return 1.0; }
    if (p == 1.0) { inv = xinv; return x; }
    if (p == -1.0) { inv = x; return xinv; }
    if (p == 2.0) { inv = frcp(x*x); return x*x; }
    if (p == -2.0) { inv = x*x; return frcp(x*x); }
    double s, rs;
    if (p == 0.5 || p == -0.5 || p == 0.25 || p == -0.25 || p == 1.5 || p == -1.5 || p == 0.125 || p == -0.125) {
        fsqrt2(x, s, rs);
        if (p == 0.5) { inv = rs; return s; }
        if (p == -0.5) { inv = s; return rs; }
        if (p == 1.5) { inv = rs*rs*rs; return x*s; }
        if (p == -1.5) { inv = x*s; return rs*rs*rs; }
        double s2, rs2;
        fsqrt2(s, s2, rs2);
        if (p == 0.25) { inv = rs2; return s2; }
        if (p == -0.25) { inv = s2; return rs2; }
        double s3, rs3;
        fsqrt2(s2, s3, rs3);
        if (p == 0.125) { inv = rs3; return s3; }
        inv = s3; return rs3;
    }
    const double v = pow(x, p);
    inv = frcp(v);
    return v;
}

// ---- per-CTA constants of one effect entry (shared memory; particle-independent sub-expressions) ------------
constexpr int kC = 24;
// common: c[0] = G, c[1] = body mass, c[2] = 1/body mass, c[kC-1] = 1/G, c[kC-2] = sqrt(G)
// RADIATION     c[3] = G m_s (mu), c[4] = 1/c
// HARMONICS     c[3] = 3/2 G m J2 R^2, c[4] = 5/8 G m J4 R^4 (0 without J4), c[5..13] = u, v, w
// GR_POTENTIAL  c[3] = 6 (G m)^2/c^2
// YARKOVSKY     c[3] = 1/c, c[4] = 3 L/(64 pi c), c[5] = 3 L/(16 pi c), c[6] = 0.5 (sb/pi^5)^(1/4) [times emissivity^(1/4) per particle], c[7] = L
// MODIFY_ORBITS c[3] = ide_position, c[4] = ide_width
// GAS_DAMPING   c[3] = sqrt(G M), c[4] = 1/sqrt(G M), c[5] = cs_coeff, c[6] = 1/cs_coeff, c[7] = -tau_coeff M, c[8] = -tau_coeff
// TYPE_I        c[3] = beta, c[4] = h0, c[5] = 1/h0, c[6] = sd0, c[7] = s, c[8] = dedge, c[9] = hedge,
//               c[10] = sqrt(M^3), c[11] = sqrt(G), c[12] = (2.7 + 1.1 s)/2, c[13], c[14] = trap band, c[15], c[16] = beta, s in
//               quarters (or 99), c[17] = 1/sqrt(M^3)
// Exponent p of a uniform power law as an integer number of quarters, |p| <= 2.75, or 99: not on that grid.
__device__ __forceinline__ double quarter_code(double p) {
    const double k = 4.0*p;
    return (k == rint(k) && fabs(k) <= 11.0) ? k : 99.0;
}

__device__ void fill_constants(const rebx_b200_effect& fx, const double* __restrict__ body, double* __restrict__ c) {
    const double G = fx.scal[0], bm = body[REBX_B200_BODY_M];
    c[0] = G; c[1] = bm; c[2] = frcp(bm); c[kC - 1] = frcp(G); c[kC - 2] = fsqrt(G);
    switch (fx.kind) {
    case REBX_B200_RADIATION: c[3] = G*bm; c[4] = 1.0/fx.scal[1]; break;
    case REBX_B200_HARMONICS: {
        const double J2 = body[REBX_B200_BODY_J2], J4 = body[REBX_B200_BODY_J4], R = body[REBX_B200_BODY_REQ];
        c[3] = 1.5*G*bm*J2*R*R; c[4] = 0.625*G*bm*J4*R*R*R*R;
        for (int j = 0; j < 9; j++) c[5 + j] = body[REBX_B200_BODY_U + j];
        break;
    }
    case REBX_B200_GR_POTENTIAL: c[3] = 6.*(G*bm)*(G*bm)/fx.scal[1]; break;
    case REBX_B200_YARKOVSKY: {
        const double L = fx.scal[1], cc = fx.scal[2], sb = fx.scal[3];
        c[3] = 1.0/cc; c[4] = 3.0*L/(64.0*kPi*cc); c[5] = 3.0*L/(16.0*kPi*cc); c[6] = 0.5*sqrt(sqrt(sb/kPi5)); c[7] = L;
        break;
    }
    case REBX_B200_MODIFY_ORBITS:
        c[3] = fx.scal[1]; c[4] = fx.scal[2];
        c[13] = fx.scal[1]*(1.0 + fx.scal[2])*(1.0 + 1e-12); c[14] = fx.scal[1]*(1.0 - fx.scal[2])*(1.0 - 1e-12);    // trap band (trap_factor)
        break;
    case REBX_B200_GAS_DAMPING: { double sq, rsq; fsqrt2(G*bm, sq, rsq); c[3] = sq; c[4] = rsq; c[5] = fx.scal[1]; c[6] = frcp(fx.scal[1]); c[7] = -fx.scal[2]*bm; c[8] = -fx.scal[2]; break; }
    case REBX_B200_TYPE_I:
        c[3] = fx.scal[1]; c[4] = fx.scal[2]; c[5] = frcp(fx.scal[2]); c[6] = fx.scal[3]; c[7] = fx.scal[4]; c[8] = fx.scal[5]; c[9] = fx.scal[6];
        { double rq; fsqrt2(bm*bm*bm, c[10], rq); c[17] = rq; }
        c[11] = fsqrt(G); c[12] = 0.5*(2.7 + 1.1*fx.scal[4]);
        c[13] = fx.scal[5]*(1.0 + fx.scal[6])*(1.0 + 1e-12); c[14] = fx.scal[5]*(1.0 - fx.scal[6])*(1.0 - 1e-12);    // trap band (trap_factor)
        c[15] = quarter_code(fx.scal[1]); c[16] = quarter_code(fx.scal[4]);          // exponents in quarters (pow_quarters)
        break;
    default: break;
    }
}

// Geometry of one particle relative to the central body, shared by the whole stack.
struct Geo { double dx, dy, dz, r2, r, rinv, rinv2, dvx, dvy, dvz, rv; };
struct Elements { double a, e, inc, sa, rsa; };       // sa = sqrt(a), rsa = 1/sqrt(a)

__device__ __forceinline__ double planet_trap(double r, double dedge, double hedge) {      // src/inner_disk_edge.c:61-77
    if (r > dedge*(1.0 + hedge)) return 1.0;
    if (dedge*(1.0 - hedge) < r) return 5.5*cos(((dedge*(1.0 + hedge) - r)*2*kPi)/(4*hedge*dedge)) - 4.5;
    return -10.0;
}

// Two-body elements a, e, inc relative to the central body (REBOUND's reb_orbit_from_particle_err as restated in
// forces_math.cuh).  Two of their uses are ill conditioned on the nearly planar, nearly circular orbits of a
// planetesimal disc: the inclination is acos(h_z/h), and the planet trap is a cosine of ~80 (a - a_edge)/a_edge --
// both amplify the rounding of their argument by 10^2..10^3 (measured with fast reciprocals throughout: 2e-12
// norm-wise on the pure damping force).  Those two arguments therefore follow the reference's own operation
// sequence with IEEE operators that the compiler may not contract (__dmul_rn & co.), which makes them bit-identical
// to the exact build's: h_z/h always, the semi-major axis only for a particle inside the trap's transition band
// (trap_factor).  Everything else -- a elsewhere, the eccentricity, all the forces -- is well conditioned and uses the
// tolerance arithmetic.
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double xdot(double ax, double ay, double az, double bx, double by, double bz) {
    return xadd(xadd(xmul(ax, bx), xmul(ay, by)), xmul(az, bz));
}

__device__ __noinline__ double exact_semimajor_axis(double G, double bm, double m, double dx, double dy, double dz, double dvx, double dvy, double dvz) {
    const double mu = xmul(G, xadd(m, bm));
    const double d = __dsqrt_rn(xdot(dx, dy, dz, dx, dy, dz));
    const double v2 = xdot(dvx, dvy, dvz, dvx, dvy, dvz);
    const double vc2 = __ddiv_rn(mu, d);
    return __ddiv_rn(-mu, xsub(v2, xmul(2., vc2)));
}

// acos(c) for the rounded cosine c of the inclination.  Disc models live at c -> 1, where acos(c) = 2 asin(s) with
// s = sqrt((1 - c)/2) (1 - c is exact there, Sterbenz) and the Maclaurin series of asin converges in a few terms:
// for 1 - c < 5e-3 (inc < 0.1 rad) eight terms leave < 2e-21 s.  Anything else goes to libdevice.
__device__ __forceinline__ double acos_near_one(double c) {
    const double t = 1.0 - c;
    if (!(t < 5e-3) || !(t > 0.0)) return acos(c);
    const double s2 = 0.5*t, s = fsqrt(s2);
    // (2k)!/(4^k (k!)^2 (2k+1)), k = 1..8
    double p = 6435.0/557056.0;
    p = fma(p, s2, 143.0/10240.0);
    p = fma(p, s2, 231.0/13312.0);
    p = fma(p, s2, 63.0/2816.0);
    p = fma(p, s2, 35.0/1152.0);
    p = fma(p, s2, 5.0/112.0);
    p = fma(p, s2, 3.0/40.0);
    p = fma(p, s2, 1.0/6.0);
    const double as = fma(s*s2, p, s);
    return as + as;
}

template <bool WANT_E_INC>
__device__ __forceinline__ Elements elements(double G, double inv_G, double bm, double m, double inv_mb, const Geo& g) {
    Elements o;
    const double mu = G*(m + bm);
    const double v2 = g.dvx*g.dvx + g.dvy*g.dvy + g.dvz*g.dvz;
    const double vc2 = mu*g.rinv;
    o.a = -mu*frcp(v2 - 2.*vc2);
    o.e = 0.0; o.inc = 0.0; o.sa = 0.0; o.rsa = 0.0;
    if (WANT_E_INC) {
        const double inv_mu = inv_mb*inv_G, vd2 = v2 - vc2;        // 1/(G (m + M))
        const double ex = inv_mu*(vd2*g.dx - g.rv*g.dvx), ey = inv_mu*(vd2*g.dy - g.rv*g.dvy), ez = inv_mu*(vd2*g.dz - g.rv*g.dvz);
        o.e = fsqrt(ex*ex + ey*ey + ez*ez);
        const double hx = xsub(xmul(g.dy, g.dvz), xmul(g.dz, g.dvy));
        const double hy = xsub(xmul(g.dz, g.dvx), xmul(g.dx, g.dvz));
        const double hz = xsub(xmul(g.dx, g.dvy), xmul(g.dy, g.dvx));
        const double cosine = __ddiv_rn(hz, __dsqrt_rn(xdot(hx, hy, hz, hx, hy, hz)));
        o.inc = (cosine > -1. && cosine < 1.) ? acos_near_one(cosine) : ((cosine <= -1.) ? kPi : 0.);
        fsqrt2(o.a, o.sa, o.rsa);
    }
    return o;
}

// planet trap of src/inner_disk_edge.c:61-77 evaluated at the particle's semi-major axis: outside the transition band
// it is a constant, inside it the axis is recomputed with the reference's exact operation sequence first
__device__ __forceinline__ double trap_factor(double a, double dedge, double hedge, double G, double bm, double m, const Geo& g) {
    const double hi = dedge*(1.0 + hedge), lo = dedge*(1.0 - hedge);
    if (a > hi*(1.0 + 1e-12)) return 1.0;
    if (!(a > lo*(1.0 - 1e-12))) return -10.0;
    return planet_trap(exact_semimajor_axis(G, bm, m, g.dx, g.dy, g.dz, g.dvx, g.dvy, g.dvz), dedge, hedge);
}

// x^1.2 = (x z^2)^2 with z = x^(-1/5): single-precision seed (MUFU lg2 / ex2), two division-free Newton steps
// z <- z (6 - x z^5)/5 in FP64 (quadratic: 1e-6 -> 3e-12 -> below eps), ~16 FP64 operations against ~150 for pow().
__device__ __forceinline__ double pow_1p2(double x) {
    if (!(x > 0.0)) return 0.0;
    if (x < 1e-30) return pow(x, 1.2);                   // below the seed's comfortable single-precision range
    double z = (double)__powf((float)x, -0.2f);
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const double z2 = z*z;
        z = z*fma(-x*z2, z2*z, 6.0)*0.2;
    }
    const double t = x*z*z;
    return t*t;
}

// ---- effects (each adds to `a`; `red` = signed back-reaction sum onto the body) ------------------------------
__device__ __forceinline__ void radiation(const double* __restrict__ c, const Geo& g, double beta, Vec3& a) {
    if (is_absent(beta)) return;
    const double rdot = g.rv*g.rinv;
    const double a_rad = beta*c[3]*g.rinv2;
    const double k1 = (1. - rdot*c[4])*g.rinv;
    a.x += a_rad*(k1*g.dx - g.dvx*c[4]);
    a.y += a_rad*(k1*g.dy - g.dvy*c[4]);
    a.z += a_rad*(k1*g.dz - g.dvz*c[4]);
}

__device__ __forceinline__ void harmonics(const double* __restrict__ c, const Geo& g, double m, Vec3& a, Vec3& red) {
    const double du = c[5]*g.dx + c[6]*g.dy + c[7]*g.dz;
    const double dv = c[8]*g.dx + c[9]*g.dy + c[10]*g.dz;
    const double dw = c[11]*g.dx + c[12]*g.dy + c[13]*g.dz;
    const double ct = dw*g.rinv, c2 = ct*ct;
    const double r5 = g.rinv2*g.rinv2*g.rinv;
    const double f1 = c[3]*r5, f2 = 5.0*c2 - 1.0;
    double au = f1*f2*du, av = f1*f2*dv, aw = f1*(f2 - 2.0)*dw;
    if (c[4] != 0.0) {
        const double g1 = c[4]*r5*g.rinv2, g2 = 63.0*c2*c2 - 42.0*c2 + 3.0;
        au += g1*g2*du; av += g1*g2*dv; aw += g1*(g2 - 28.0*c2 + 12.0)*dw;
    }
    const double ax = c[5]*au + c[8]*av + c[11]*aw;
    const double ay = c[6]*au + c[9]*av + c[12]*aw;
    const double az = c[7]*au + c[10]*av + c[13]*aw;
    a.x += ax; a.y += ay; a.z += az;
    const double fac = m*c[2];
    red.x -= fac*ax; red.y -= fac*ay; red.z -= fac*az;
}

__device__ __forceinline__ void gr_potential(const double* __restrict__ c, const Geo& g, double m, Vec3& a, Vec3& red) {
    const double prefac = c[3]*g.rinv2*g.rinv2;
    a.x -= prefac*g.dx; a.y -= prefac*g.dy; a.z -= prefac*g.dz;
    const double mr = m*c[2]*prefac;
    red.x += mr*g.dx; red.y += mr*g.dy; red.z += mr*g.dz;
}

struct YarkParams { double density, rot_period, Gamma, albedo, emissivity, k, sx, sy, sz; };

// v' = cos v + sin (k x v) + (1 - cos)(k.v) k : rotation of v about the UNIT vector k (sin < 0 rotates the other way)
__device__ __forceinline__ void rotate(double kx, double ky, double kz, double cs, double sn, double& vx, double& vy, double& vz) {
    const double cx = ky*vz - kz*vy, cy = kz*vx - kx*vz, cz = kx*vy - ky*vx;
    const double kd = (1.0 - cs)*(kx*vx + ky*vy + kz*vz);
    const double nx = cs*vx + sn*cx + kd*kx, ny = cs*vy + sn*cy + kd*ky, nz = cs*vz + sn*cz + kd*kz;
    vx = nx; vy = ny; vz = nz;
}

__device__ __forceinline__ void yarkovsky(const double* __restrict__ c, const Geo& g, double m, double radius, const YarkParams& yp, int code, Vec3& a) {
    if (code == 2 || radius == 0) return;
    const double q = 1.0 - yp.albedo, inv_c = c[3];
    const double k1 = (1.0 - g.rv*inv_c*g.rinv)*g.rinv;
    double i0 = k1*g.dx - g.dvx*inv_c, i1 = k1*g.dy - g.dvy*inv_c, i2 = k1*g.dz - g.dvz*inv_c;
    const double geom = q*frcp(radius*yp.density*g.r2);          // (1 - A)/(R rho r^2)
    if (code != 0) {
        const double mag = c[4]*geom;
        if (code == 1) a.y += mag*i0; else a.x += mag*i1;        // Y[1][0] = 1 resp. Y[0][1] = 1 (yarkovsky_effect.c:109-121)
        return;
    }
    const double mag = c[5]*yp.k*geom;
    // orbital period: n = sign(a) sqrt(|mu/a^3|), P = 2 pi/n; only sqrt(P) enters (NaN for an unbound orbit, like the reference)
    const double mu = c[0]*(m + c[1]);
    const double v2 = g.dvx*g.dvx + g.dvy*g.dvy + g.dvz*g.dvz;
    const double inv_a = -(v2 - 2.*mu*g.rinv)*frcp(mu);
    const double n = copysign(fsqrt(fabs(mu*inv_a*inv_a*inv_a)), inv_a);
    const double sqrtP = frsqrt(n*(1.0/kTwoPi));
    const double inv_gamma = fabs(frcp(yp.Gamma));
    // X = 0.5 (sb eps/pi^5)^(1/4) sqrt(P_x/Gamma^2) ((1-A) L/r^2)^(3/4)
    const double fsq = fsqrt(c[7]*q)*g.rinv;
    const double flux = fsq*fsqrt(fsq);
    const double pre = c[6]*fsqrt(fsqrt(yp.emissivity))*inv_gamma*flux;
    const double tanPhi = frcp(1.0 + pre*fsqrt(yp.rot_period));
    const double tanEps = frcp(1.0 + pre*sqrtP);
    const double cphi = frsqrt(1.0 + tanPhi*tanPhi), sphi = tanPhi*cphi;
    const double ceps = frsqrt(1.0 + tanEps*tanEps), seps = tanEps*ceps;
    const double inv_s = frsqrt(yp.sx*yp.sx + yp.sy*yp.sy + yp.sz*yp.sz);
    const double hx = g.dy*g.dvz - g.dz*g.dvy, hy = g.dz*g.dvx - g.dx*g.dvz, hz = g.dx*g.dvy - g.dy*g.dvx;
    const double inv_h = frsqrt(hx*hx + hy*hy + hz*hz);
    rotate(hx*inv_h, hy*inv_h, hz*inv_h, ceps, -seps, i0, i1, i2);       // Ryh = cos I - sin [h]x + (1 - cos) h h^T   (:176-182)
    rotate(yp.sx*inv_s, yp.sy*inv_s, yp.sz*inv_s, cphi, sphi, i0, i1, i2);   // Rys = cos I + sin [s]x + (1 - cos) s s^T   (:170-175)
    a.x += mag*i0; a.y += mag*i1; a.z += mag*i2;
}

__device__ __forceinline__ Vec3 modify_orbits(const double* __restrict__ c, int flags, const Geo& g, const Elements& o, double m,
                                              double tau_a, double tau_e, double tau_inc) {
    double invtau_a = 0.0;
    if (!is_absent(tau_a)) {
        invtau_a = frcp(tau_a);
        if (flags & REBX_B200_FLAG_TRAP) invtau_a *= trap_factor(o.a, c[3], c[4], c[0], c[1], m, g);
    }
    const double ite = is_absent(tau_e) ? 0.0 : frcp(tau_e), iti = is_absent(tau_inc) ? 0.0 : frcp(tau_inc);
    const double prefac = 2.*g.rv*g.rinv2*ite;
    Vec3 f;
    f.x = 0.5*g.dvx*invtau_a + prefac*g.dx;
    f.y = 0.5*g.dvy*invtau_a + prefac*g.dy;
    f.z = 0.5*g.dvz*invtau_a + prefac*g.dz + 2.*g.dvz*iti;
    return f;
}

__device__ __forceinline__ Vec3 gas_damping(const double* __restrict__ c, const Geo& g, const Elements& o, double m, double d_factor) {
    Vec3 f = {0.0, 0.0, 0.0};
    if (is_absent(d_factor)) return f;
    const double vk = c[3]*o.rsa;                         // sqrt(G M/a)
    const double v = fsqrt(o.e*o.e + o.inc*o.inc)*vk;
    double s4, rs4;
    fsqrt2(o.sa, s4, rs4);                                // a^(1/4)
    const double cs = c[5]*rs4;
    const double voc = v*s4*c[6];                         // v/cs
    double coeff = 1.;
    if (!(v <= cs)) coeff = (o.inc < cs*o.sa*c[4]) ? voc*voc*voc : voc*voc*voc*voc;
    // tau_e = -tau_coeff d a^2 (M/m) coeff ; tau_inc = 2 tau_e
    const double ite = m*frcp(c[7]*d_factor*o.a*o.a*coeff);
    const double prefac = 2.*g.rv*g.rinv2*ite;
    f.x = prefac*g.dx; f.y = prefac*g.dy; f.z = prefac*g.dz + g.dvz*ite;
    return f;
}

__device__ __forceinline__ Vec3 type_I_migration(const double* __restrict__ c, const Geo& g, const Elements& o, double m) {
    double inv_hr, inv_sd;
    const double hr = pow_uniform2(g.r, g.rinv, c[3], inv_hr);              // r^beta   (h = h0 r2^(beta/2))
    const double h = c[4]*hr, inv_h = c[5]*inv_hr, h2 = h*h;
    const double eh = o.e*inv_h, ih = o.inc*inv_h;
    const double rs = pow_uniform2(g.r, g.rinv, c[7], inv_sd);              // r^s ; sd = sd0 r^-s
    (void)rs;
    const double sd = c[6]*inv_sd;
    // 1/wave = m sd sqrt(a G)/(sqrt(M^3) h^4)
    const double inv_wave = m*sd*c[11]*o.sa*frcp(c[10]*h2*h2);
    const double x1 = eh*(1.0/2.25), x2 = eh*(1.0/2.84), x3 = eh*(1.0/2.02);
    const double x22 = x2*x2, x32 = x3*x3;
    const double p12 = pow_1p2(x1);
    const double Pe = (1. + p12 + x22*x22*x22)*frcp(1. - x32*x32);
    const double ih2 = ih*ih;
    const double corr = 0.070*ih + 0.085*ih2*ih2 - 0.080*eh*ih2;
    const double invtau_mig = trap_factor(o.a, c[8], c[9], c[0], c[1], m, g)*inv_wave*c[12]*h2*frcp(Pe + (Pe < 0.0 ? -corr : corr));
    const double ite = 0.780*inv_wave*frcp(1. - 0.14*eh*eh + 0.06*eh*eh*eh + 0.18*eh*ih2);
    const double iti = 0.544*inv_wave*frcp(1. - 0.30*ih2 + 0.24*ih2*ih + 0.14*eh*eh*ih);
    const double prefac = -2.*g.rv*g.rinv2*ite;
    Vec3 f;
    f.x = -g.dvx*invtau_mig + prefac*g.dx;
    f.y = -g.dvy*invtau_mig + prefac*g.dy;
    f.z = -g.dvz*invtau_mig + prefac*g.dz - 2.*g.dvz*iti;
    return f;
}

// ---- the damping effects in ONE basic block ---------------------------------------------------------------------
// ncu on the functions above (profiles/r02_kernel_tolerance_planetesimal_disk.txt): 820 warp instructions per particle,
// 30 of them branches, 4 warps per scheduler issuing 46 % of the cycles with `stall_wait` on top -- every warp a serial
// FP64 chain, because each quotient's Newton steps, each root and each effect sits in its own basic block (IEEE
// division / sqrt slow-path calls, acos and pow fall-backs, is_absent and trap tests) and the scheduler cannot
// interleave across them.  damping_fast evaluates the same formulas without a single branch: rare cases are not
// handled inline but FLAGGED in `bad`, and the caller redoes such a particle with the functions above
// (damping_generic); the planet trap's transition band, the one rare case that is not THAT rare (0.2 % of a disc),
// only marks its particle, and the caller recomputes the two trap factors -- the forces are linear in them.
//   flagged: inclination outside acos_near_one's series, e/h below pow_1p2's single-precision seed, an exponent that
//   is not a multiple of 1/4 (uniform: then every particle takes the generic path), operands outside exact_math's domain.
struct DampCoef { double cv_mo, cv_ti, cr, cz; };     // fd = dv (cv_mo T_mo + cv_ti T_ti) + d cr ; fd.z += dvz cz
// What the formulas need of the reference body's MASS: per CTA in PARTICLE coordinates (the central body, from the
// constants in shared memory), per particle in JACOBI coordinates (the mass interior to the particle).
struct BodyK { double bm, sqrtGM, rsqrtGM, rsqrtM3; };

// x^(k/4) and x^(-k/4) from x, 1/x and the two roots, k = c (an integer in [-11, 11], uniform): selects and products only
__device__ __forceinline__ double pow_quarters(double kq, double x, double xinv, double s1, double rs1, double s2, double rs2, double& inv) {
    const int k = (int)kq, ka = k < 0 ? -k : k, q = ka >> 2;
    double P = (q == 0) ? 1.0 : ((q == 1) ? x : x*x);
    double N = (q == 0) ? 1.0 : ((q == 1) ? xinv : xinv*xinv);
    const double P1 = P*s1, N1 = N*rs1;
    P = (ka & 2) ? P1 : P;  N = (ka & 2) ? N1 : N;
    const double P2 = P*s2, N2 = N*rs2;
    P = (ka & 1) ? P2 : P;  N = (ka & 1) ? N2 : N;
    inv = (k < 0) ? P : N;
    return (k < 0) ? N : P;
}

template <bool MO, bool GD, bool TI>
__device__ __forceinline__ DampCoef damping_fast(const double* __restrict__ cmo, const double* __restrict__ cgd, const double* __restrict__ cti,
                                                 int mo_flags, const BodyK& bk, const Geo& g, double m, double inv_mb,
                                                 double tau_a, double tau_e, double tau_inc, double d_factor,
                                                 double& a_out, unsigned& bad, unsigned& band) {
    const double* c0 = MO ? cmo : (GD ? cgd : cti);            // G and 1/G are the same in every effect of the stack
    const double G = c0[0], bm = bk.bm, inv_G = c0[kC - 1];
    // -- two-body elements (elements<> above), straight-line
    const double mu = G*(m + bm);
    const double v2 = g.dvx*g.dvx + g.dvy*g.dvy + g.dvz*g.dvz;
    const double vc2 = mu*g.rinv;
    const double a = -mu*frcp(v2 - 2.*vc2);
    a_out = a;
    double e = 0.0, inc = 0.0, sa = 0.0, rsa = 0.0;
    if constexpr (GD || TI) {
        const double inv_mu = inv_mb*inv_G, vd2 = v2 - vc2;
        const double ex = inv_mu*(vd2*g.dx - g.rv*g.dvx), ey = inv_mu*(vd2*g.dy - g.rv*g.dvy), ez = inv_mu*(vd2*g.dz - g.rv*g.dvz);
        e = fsqrt(ex*ex + ey*ey + ez*ez);
        // cos(inc) = h_z/h with the reference's own operation sequence, correctly rounded (exact_math.cuh: same bits as / and sqrt)
        const double hx = xsub(xmul(g.dy, g.dvz), xmul(g.dz, g.dvy));
        const double hy = xsub(xmul(g.dz, g.dvx), xmul(g.dx, g.dvz));
        const double hz = xsub(xmul(g.dx, g.dvy), xmul(g.dy, g.dvx));
        const double hh = exact_sqrt(xdot(hx, hy, hz, hx, hy, hz), bad);
        const double cosine = exact_div(hz, exact_rcp(hh, bad), bad);
        // acos_near_one's series; anything else is flagged
        const double t = 1.0 - cosine;
        const bool inside = (cosine > -1. && cosine < 1.);
        bad |= (inside && !(t < 5e-3)) ? 1u : 0u;
        const double s2h = 0.5*t, sh = fsqrt(s2h);
        // Estrin on the 8 coefficients (2k)!/(4^k (k!)^2 (2k+1)), k = 1..8: depth 3 instead of 8
        const double z2 = s2h*s2h, z4 = z2*z2;
        const double p01 = fma(3.0/40.0, s2h, 1.0/6.0), p23 = fma(35.0/1152.0, s2h, 5.0/112.0);
        const double p45 = fma(231.0/13312.0, s2h, 63.0/2816.0), p67 = fma(6435.0/557056.0, s2h, 143.0/10240.0);
        const double pl = fma(p23, z2, p01), ph = fma(p67, z2, p45);
        const double poly = fma(ph, z4, pl);
        const double as = fma(sh*s2h, poly, sh);
        inc = inside ? (as + as) : ((cosine <= -1.) ? kPi : 0.);
        fsqrt2(a, sa, rsa);
    }
    DampCoef o = {0.0, 0.0, 0.0, 0.0};
    const double two_rv_r2 = 2.*g.rv*g.rinv2;
    double ite_sum = 0.0;
    if constexpr (MO) {
        const double ita = is_absent(tau_a) ? 0.0 : frcp(tau_a);
        const double ite = is_absent(tau_e) ? 0.0 : frcp(tau_e), iti = is_absent(tau_inc) ? 0.0 : frcp(tau_inc);
        o.cv_mo = 0.5*ita;
        ite_sum += ite;
        o.cz += 2.*iti;
        if (mo_flags & REBX_B200_FLAG_TRAP) band |= (!(a > cmo[13]) && (a > cmo[14]) && !is_absent(tau_a)) ? 1u : 0u;
    }
    if constexpr (GD) {
        const double vk = bk.sqrtGM*rsa;
        const double v = fsqrt(e*e + inc*inc)*vk;
        double s4, rs4;
        fsqrt2(sa, s4, rs4);
        const double cs = cgd[5]*rs4;
        const double voc = v*s4*cgd[6];
        const double voc3 = voc*voc*voc;
        const double coeff = (v <= cs) ? 1. : ((inc < cs*sa*bk.rsqrtGM) ? voc3 : voc3*voc);
        const double ite = m*frcp(cgd[8]*bk.bm*d_factor*a*a*coeff);
        const bool on = !is_absent(d_factor);
        ite_sum += on ? ite : 0.0;
        o.cz += on ? ite : 0.0;
    }
    if constexpr (TI) {
        double s1, rs1, s2, rs2, inv_hr, inv_sd;
        fsqrt2(g.r, s1, rs1);
        fsqrt2(s1, s2, rs2);
        bad |= (cti[15] == 99.0 || cti[16] == 99.0) ? 1u : 0u;
        const double hr = pow_quarters(cti[15], g.r, g.rinv, s1, rs1, s2, rs2, inv_hr);
        (void)pow_quarters(cti[16], g.r, g.rinv, s1, rs1, s2, rs2, inv_sd);
        const double h = cti[4]*hr, inv_h = cti[5]*inv_hr, h2 = h*h;
        const double eh = e*inv_h, ih = inc*inv_h;
        const double sd = cti[6]*inv_sd;
        const double ih4 = inv_h*inv_h;                                  // 1/h^4 from the inverse power, no quotient
        const double inv_wave = m*sd*cti[11]*sa*bk.rsqrtM3*(ih4*ih4);
        const double x1 = eh*(1.0/2.25), x2 = eh*(1.0/2.84), x3 = eh*(1.0/2.02);
        const double x22 = x2*x2, x32 = x3*x3;
        // pow_1p2, straight-line: x = 0 gives 0 by the select, a positive x below the seed's range is flagged
        bad |= (x1 > 0.0 && x1 < 1e-30) ? 1u : 0u;
        double z = (double)__powf((float)x1, -0.2f);
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const double zz = z*z;
            z = z*fma(-x1*zz, zz*z, 6.0)*0.2;
        }
        const double t12 = x1*z*z;
        const double p12 = (x1 > 0.0) ? t12*t12 : 0.0;
        const double Pe = (1. + p12 + x22*x22*x22)*frcp(1. - x32*x32);
        const double ih2 = ih*ih;
        const double corr = 0.070*ih + 0.085*ih2*ih2 - 0.080*eh*ih2;
        o.cv_ti = -inv_wave*cti[12]*h2*frcp(Pe + (Pe < 0.0 ? -corr : corr));
        const double ite = 0.780*inv_wave*frcp(1. - 0.14*eh*eh + 0.06*eh*eh*eh + 0.18*eh*ih2);
        const double iti = 0.544*inv_wave*frcp(1. - 0.30*ih2 + 0.24*ih2*ih + 0.14*eh*eh*ih);
        ite_sum -= ite;
        o.cz -= 2.*iti;
        band |= (!(a > cti[13]) && (a > cti[14])) ? 2u : 0u;
    }
    o.cr = two_rv_r2*ite_sum;
    return o;
}

template <int K> __host__ __device__ constexpr bool is_damping() { return K == REBX_B200_MODIFY_ORBITS || K == REBX_B200_GAS_DAMPING || K == REBX_B200_TYPE_I; }

template <int K, int K0, int K1, int K2> __host__ __device__ constexpr int slot_of() { return K0 == K ? 0 : (K1 == K ? 1 : (K2 == K ? 2 : -1)); }
template <int K0, int K1, int K2> __host__ __device__ constexpr bool all_damping() {
    return is_damping<K0>() && (K1 == 0 || is_damping<K1>()) && (K2 == 0 || is_damping<K2>());
}

// The flagged particles of damping_fast: the branching functions above, effects in stack order.  Out of line (rare).
template <int K0, int K1, int K2>
__device__ __noinline__ void damping_generic(const rebx_b200_pass* P, const double (*s_c)[kC], const Geo* gp, double m, double inv_mb,
                                             double tau_a, double tau_e, double tau_inc, double d_factor, Vec3* out) {
    constexpr bool EINC = slot_of<REBX_B200_GAS_DAMPING, K0, K1, K2>() >= 0 || slot_of<REBX_B200_TYPE_I, K0, K1, K2>() >= 0;
    const Geo g = *gp;
    const Elements o = elements<EINC>(s_c[0][0], s_c[0][kC - 1], s_c[0][1], m, inv_mb, g);
    Vec3 fd = {0., 0., 0.};
    auto one = [&](auto kc, int j) {
        constexpr int K = decltype(kc)::value;
        Vec3 f = {0., 0., 0.};
        if constexpr (K == REBX_B200_MODIFY_ORBITS) f = modify_orbits(s_c[j], P->fx[j].flags, g, o, m, tau_a, tau_e, tau_inc);
        else if constexpr (K == REBX_B200_GAS_DAMPING) f = gas_damping(s_c[j], g, o, m, d_factor);
        else if constexpr (K == REBX_B200_TYPE_I) f = type_I_migration(s_c[j], g, o, m);
        fd.x += f.x; fd.y += f.y; fd.z += f.z;
    };
    one(std::integral_constant<int, K0>{}, 0);
    if constexpr (K1 != 0) one(std::integral_constant<int, K1>{}, 1);
    if constexpr (K2 != 0) one(std::integral_constant<int, K2>{}, 2);
    *out = fd;
}

template <int K, int V>
__device__ __forceinline__ void apply(const rebx_b200_effect& fx, const double* __restrict__ c, const Geo& g, const Elements& o, double mr,
                                      double m, double rad, const Params<K, V>& q, int v, Vec3& a, Vec3& red) {
    if constexpr (K == REBX_B200_RADIATION) radiation(c, g, q.t[0][v], a);
    else if constexpr (K == REBX_B200_HARMONICS) harmonics(c, g, m, a, red);
    else if constexpr (K == REBX_B200_GR_POTENTIAL) gr_potential(c, g, m, a, red);
    else if constexpr (K == REBX_B200_YARKOVSKY) {
        const YarkParams yp{q.t[0][v], q.t[1][v], q.t[2][v], q.t[3][v], q.t[4][v], q.t[5][v], q.t[6][v], q.t[7][v], q.t[8][v]};
        yarkovsky(c, g, m, rad, yp, q.it[v], a);
    } else if constexpr (is_damping<K>()) {
        // the damping effects of a stack share one rebx_com_force tail (pass_kernel_tol): their forces are summed in `a`
        // here, which the caller hands in as a separate accumulator
        Vec3 f;
        if constexpr (K == REBX_B200_MODIFY_ORBITS) f = modify_orbits(c, fx.flags, g, o, m, q.t[0][v], q.t[1][v], q.t[2][v]);
        else if constexpr (K == REBX_B200_GAS_DAMPING) f = gas_damping(c, g, o, m, q.t[0][v]);
        else f = type_I_migration(c, g, o, m);
        a.x += f.x; a.y += f.y; a.z += f.z;
        (void)mr; (void)red;
    }
}

// CTA geometry per stack class (tools/tol_variants.sh sweeps them on the GPU): threads per CTA and resident CTAs per
// SM asked of the register allocator.  "light" = the streaming stacks (radiation, gr_potential, J2/J4), "heavy" =
// the ones with transcendentals (yarkovsky, damping trio).
#ifndef REBXB_TOL_FAST_DAMP
#define REBXB_TOL_FAST_DAMP 1
#endif
#ifndef REBXB_TOL_PREFETCH_ALL
#define REBXB_TOL_PREFETCH_ALL 0
#endif
#ifndef REBXB_TOL_LIGHT_THREADS
#define REBXB_TOL_LIGHT_THREADS 256
#endif
#ifndef REBXB_TOL_LIGHT_BLOCKS
#define REBXB_TOL_LIGHT_BLOCKS 2
#endif
#ifndef REBXB_TOL_YARK_THREADS
#define REBXB_TOL_YARK_THREADS 256
#endif
#ifndef REBXB_TOL_YARK_BLOCKS
#define REBXB_TOL_YARK_BLOCKS 3
#endif
#ifndef REBXB_TOL_TRIO_THREADS
#define REBXB_TOL_TRIO_THREADS 256
#endif
#ifndef REBXB_TOL_TRIO_BLOCKS
#define REBXB_TOL_TRIO_BLOCKS 2
#endif
template <int K0, int K1, int K2>
__host__ __device__ constexpr bool trio_stack() { return K0 >= REBX_B200_MODIFY_ORBITS || K1 >= REBX_B200_MODIFY_ORBITS || K2 >= REBX_B200_MODIFY_ORBITS; }
template <int K0, int K1, int K2>
__host__ __device__ constexpr bool yark_stack() { return !trio_stack<K0, K1, K2>() && (K0 == REBX_B200_YARKOVSKY || K1 == REBX_B200_YARKOVSKY || K2 == REBX_B200_YARKOVSKY); }
template <int K0, int K1, int K2>
__host__ __device__ constexpr int tol_blocks() {
    return trio_stack<K0, K1, K2>() ? REBXB_TOL_TRIO_BLOCKS : (yark_stack<K0, K1, K2>() ? REBXB_TOL_YARK_BLOCKS : REBXB_TOL_LIGHT_BLOCKS);
}
template <int K0, int K1, int K2>
__host__ __device__ constexpr int tol_threads() {
    return trio_stack<K0, K1, K2>() ? REBXB_TOL_TRIO_THREADS : (yark_stack<K0, K1, K2>() ? REBXB_TOL_YARK_THREADS : REBXB_TOL_LIGHT_THREADS);
}

// ---- operands of the NEXT iteration staged through shared memory (cp.async) ------------------------------------------
// ncu on the trio sweep (profiles/r02_kernel_tolerance_planetesimal_disk.txt and the source page): 22 % of all stall
// samples sit on the first use of the ~14 operands loaded at the top of an iteration (4 warps per scheduler, 128
// registers: nobody to hide an HBM round trip behind, and no registers for a second set of operands); an L2 / L1
// prefetch one iteration ahead only took 30 % off that wait (pass_common.cuh).  Here every thread copies ITS OWN next
// operands global -> shared with cp.async (8 bytes each, no registers, no barrier: a thread only ever reads what it
// copied itself, after cp.async.wait_group), one full iteration (~8000 cycles) ahead, and the top of the loop reads
// them back with LDS (29 cycles).  Two stages of (operand arrays) x (CTA threads) doubles of dynamic shared memory.
#ifndef REBXB_TOL_LOCKSTEP
#define REBXB_TOL_LOCKSTEP 0
#endif
#ifndef REBXB_TOL_STAGED
#define REBXB_TOL_STAGED 1
#endif
template <int K0, int K1, int K2>
__host__ __device__ constexpr int staged_doubles() {        // per thread and stage; 0 = this stack is not staged
    constexpr bool VEL  = Traits<K0>::vel  || Traits<K1>::vel  || Traits<K2>::vel;
    constexpr bool MASS = Traits<K0>::mass || Traits<K1>::mass || Traits<K2>::mass;
    constexpr bool RAD  = Traits<K0>::rad  || Traits<K1>::rad  || Traits<K2>::rad;
    return 6 + (VEL ? 3 : 0) + (MASS ? 1 : 0) + (RAD ? 1 : 0) + Traits<K0>::ntab + Traits<K1>::ntab + Traits<K2>::ntab +
           ((Traits<K0>::itab || Traits<K1>::itab || Traits<K2>::itab) ? 1 : 0);      // an int table takes a double's slot
}
template <int V, int K0, int K1, int K2>
__host__ __device__ constexpr bool staged_stack() { return REBXB_TOL_STAGED && V == 1 && trio_stack<K0, K1, K2>(); }
template <int V, int K0, int K1, int K2>
__host__ __device__ constexpr int staged_bytes() { return staged_stack<V, K0, K1, K2>() ? 2*staged_doubles<K0, K1, K2>()*tol_threads<K0, K1, K2>()*8 : 0; }

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(int* smem_dst, const int* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int K, int T>
__device__ __forceinline__ void stage_params(const rebx_b200_effect& fx, int64_t k, double* base, int& slot) {
#pragma unroll
    for (int j = 0; j < Traits<K>::ntab; j++) cp_async8(base + (slot++)*T, fx.tab[j] + k);
    if constexpr (Traits<K>::itab) cp_async4(reinterpret_cast<int*>(base + (slot++)*T), fx.itab + k);
}
template <int K, int T>
__device__ __forceinline__ void unstage_params(const double* base, int& slot, Params<K, 1>& q) {
#pragma unroll
    for (int j = 0; j < Traits<K>::ntab; j++) q.t[j][0] = base[(slot++)*T];
    if constexpr (Traits<K>::itab) q.it[0] = *reinterpret_cast<const int*>(base + (slot++)*T);
}

template <int V, int K0, int K1, int K2>
__global__ void __launch_bounds__(tol_threads<K0, K1, K2>(), tol_blocks<K0, K1, K2>())
pass_kernel_tol(const __grid_constant__ rebx_b200_pass P, double* __restrict__ partials) {
    constexpr bool VEL  = Traits<K0>::vel  || Traits<K1>::vel  || Traits<K2>::vel;
    constexpr bool MASS = Traits<K0>::mass || Traits<K1>::mass || Traits<K2>::mass;
    constexpr bool RAD  = Traits<K0>::rad  || Traits<K1>::rad  || Traits<K2>::rad;
    constexpr bool DAMP = is_damping<K0>() || is_damping<K1>() || is_damping<K2>();
    constexpr bool EINC = K0 == REBX_B200_GAS_DAMPING || K1 == REBX_B200_GAS_DAMPING || K2 == REBX_B200_GAS_DAMPING ||
                          K0 == REBX_B200_TYPE_I || K1 == REBX_B200_TYPE_I || K2 == REBX_B200_TYPE_I;
    constexpr bool FAST_DAMP = REBXB_TOL_FAST_DAMP && all_damping<K0, K1, K2>();
    constexpr int T = tol_threads<K0, K1, K2>();
    // prefetch the next iteration's operands in the sweeps with transcendentals (few warps, long chains per particle)
    constexpr bool PREFETCH = REBXB_TOL_PREFETCH_ALL || trio_stack<K0, K1, K2>() || yark_stack<K0, K1, K2>();
    __shared__ double s_c[3][kC];
    __shared__ double s_org[8];
    const int64_t n = P.st.n;
    const int64_t ntiles = (n + V - 1)/V;
    const int64_t stride = (int64_t)gridDim.x*T;
    constexpr bool STAGED = staged_stack<V, K0, K1, K2>();
    constexpr bool LOCKSTEP = REBXB_TOL_LOCKSTEP && (trio_stack<K0, K1, K2>() || yark_stack<K0, K1, K2>());
    constexpr int ND = staged_doubles<K0, K1, K2>();
    extern __shared__ __align__(16) double s_stage[];                    // [2][ND][T] when STAGED
    auto stage = [&](int buf, int64_t k) {                               // this thread's operands of particle k -> stage `buf`
        double* base = s_stage + (size_t)buf*ND*T + threadIdx.x;
        int slot = 0;
        cp_async8(base + (slot++)*T, P.st.x + k); cp_async8(base + (slot++)*T, P.st.y + k); cp_async8(base + (slot++)*T, P.st.z + k);
        if constexpr (VEL)  { cp_async8(base + (slot++)*T, P.st.vx + k); cp_async8(base + (slot++)*T, P.st.vy + k); cp_async8(base + (slot++)*T, P.st.vz + k); }
        if constexpr (MASS) cp_async8(base + (slot++)*T, P.st.m + k);
        if constexpr (RAD)  cp_async8(base + (slot++)*T, P.st.r + k);
        cp_async8(base + (slot++)*T, P.st.ax + k); cp_async8(base + (slot++)*T, P.st.ay + k); cp_async8(base + (slot++)*T, P.st.az + k);
        stage_params<K0, T>(P.fx[0], k, base, slot);
        if constexpr (K1 != 0) stage_params<K1, T>(P.fx[1], k, base, slot);
        if constexpr (K2 != 0) stage_params<K2, T>(P.fx[2], k, base, slot);
        cp_async_commit();
    };
    int buf = 0;
    // the first particle's operands travel while the CTA's constants are worked out
    if constexpr (STAGED) { if ((int64_t)blockIdx.x*T + threadIdx.x < ntiles) stage(0, (int64_t)blockIdx.x*T + threadIdx.x); }
    if (threadIdx.x < P.n_effects) fill_constants(P.fx[threadIdx.x], P.fx[threadIdx.x].body, s_c[threadIdx.x]);
    if (threadIdx.x < 8) s_org[threadIdx.x] = P.fx[0].body[threadIdx.x];
    __syncthreads();
    const long long b0 = (long long)s_org[0];
    const double ox = s_org[REBX_B200_BODY_X], oy = s_org[REBX_B200_BODY_X+1], oz = s_org[REBX_B200_BODY_X+2];
    Vec3 red0 = {0., 0., 0.}, red1 = {0., 0., 0.}, red2 = {0., 0., 0.};

    // LOCKSTEP (an option that did NOT pay, off): the warps of a CTA start every iteration together, one barrier per
    // ~8000-cycle iteration, in case the schedulers let the warps of an SM drift iterations apart and the sweep ended
    // in a tail of a few warps finishing alone (ncu: 14 % of all warp time waits at the final barrier).  Measured
    // (profiles/r02_tol_ab.txt): trio 0.0575 -> 0.0640 ms, asteroids 0.0396 -> 0.0415 -- warps in phase hide each
    // other's latencies worse than warps out of phase.  The trip count stays uniform over the CTA for it.
    const int64_t t_cta = (int64_t)blockIdx.x*T;
    const int64_t iters = t_cta < ntiles ? (ntiles - t_cta + stride - 1)/stride : 0;        // uniform over the CTA
    for (int64_t it = 0; it < iters; it++) {
        const int64_t t = t_cta + threadIdx.x + it*stride;
        if constexpr (LOCKSTEP) __syncthreads();
        if (t >= ntiles) continue;
        const int64_t k = t*V;
        const bool full = (k + V <= n);
        if constexpr (PREFETCH && !STAGED) {
            if (t + stride < ntiles) {
                const int64_t kn = (t + stride)*V;
                prefetch_next(P.st.x + kn); prefetch_next(P.st.y + kn); prefetch_next(P.st.z + kn);
                if constexpr (VEL)  { prefetch_next(P.st.vx + kn); prefetch_next(P.st.vy + kn); prefetch_next(P.st.vz + kn); }
                if constexpr (MASS) prefetch_next(P.st.m + kn);
                if constexpr (RAD)  prefetch_next(P.st.r + kn);
                prefetch_next(P.st.ax + kn); prefetch_next(P.st.ay + kn); prefetch_next(P.st.az + kn);
                prefetch_params<K0>(P.fx[0], kn);
                if constexpr (K1 != 0) prefetch_params<K1>(P.fx[1], kn);
                if constexpr (K2 != 0) prefetch_params<K2>(P.fx[2], kn);
            }
        }
        double x[V], y[V], z[V], vx[V], vy[V], vz[V], m[V], r[V], ax[V], ay[V], az[V];
        Params<K0, V> q0; Params<K1, V> q1; Params<K2, V> q2;
        if constexpr (STAGED) {
            // start the copies of the next particle, then wait for this one's (issued one iteration ago)
            if (t + stride < ntiles) { stage(buf ^ 1, t + stride); cp_async_wait<1>(); } else cp_async_wait<0>();
            const double* base = s_stage + (size_t)buf*ND*T + threadIdx.x;
            int slot = 0;
            x[0] = base[(slot++)*T]; y[0] = base[(slot++)*T]; z[0] = base[(slot++)*T];
            if constexpr (VEL)  { vx[0] = base[(slot++)*T]; vy[0] = base[(slot++)*T]; vz[0] = base[(slot++)*T]; }
            if constexpr (MASS) m[0] = base[(slot++)*T];
            if constexpr (RAD)  r[0] = base[(slot++)*T];
            ax[0] = base[(slot++)*T]; ay[0] = base[(slot++)*T]; az[0] = base[(slot++)*T];
            unstage_params<K0, T>(base, slot, q0);
            if constexpr (K1 != 0) unstage_params<K1, T>(base, slot, q1);
            if constexpr (K2 != 0) unstage_params<K2, T>(base, slot, q2);
            buf ^= 1;
        } else {
        ld<V>(P.st.x, k, full, x); ld<V>(P.st.y, k, full, y); ld<V>(P.st.z, k, full, z);
        if constexpr (VEL)  { ld<V>(P.st.vx, k, full, vx); ld<V>(P.st.vy, k, full, vy); ld<V>(P.st.vz, k, full, vz); }
        if constexpr (MASS) { ld<V>(P.st.m, k, full, m); }
        if constexpr (RAD)  { ld<V>(P.st.r, k, full, r); }
        ld<V>(P.st.ax, k, full, ax); ld<V>(P.st.ay, k, full, ay); ld<V>(P.st.az, k, full, az);
        load_params<K0, V>(P.fx[0], k, full, q0);
        if constexpr (K1 != 0) load_params<K1, V>(P.fx[1], k, full, q1);
        if constexpr (K2 != 0) load_params<K2, V>(P.fx[2], k, full, q2);
        }
#pragma unroll
        for (int v = 0; v < V; v++) {
            if (v > 0 && !full) break;
            if (P.st.index_base + k + v == b0) continue;            // the central body itself
            Geo g;
            g.dx = x[v] - ox; g.dy = y[v] - oy; g.dz = z[v] - oz;
            g.r2 = g.dx*g.dx + g.dy*g.dy + g.dz*g.dz;
            fsqrt2(g.r2, g.r, g.rinv);
            g.rinv2 = g.rinv*g.rinv;
            if constexpr (VEL) {
                g.dvx = vx[v] - s_org[REBX_B200_BODY_VX]; g.dvy = vy[v] - s_org[REBX_B200_BODY_VX+1]; g.dvz = vz[v] - s_org[REBX_B200_BODY_VX+2];
                g.rv = g.dx*g.dvx + g.dy*g.dvy + g.dz*g.dvz;
            } else { g.dvx = g.dvy = g.dvz = g.rv = 0.0; }
            const double mp = MASS ? m[v] : 0.0, rp = RAD ? r[v] : 0.0;
            Elements o = {0., 0., 0., 0., 0.};
            double mr = 0.0, inv_mb = 0.0;
            if constexpr (DAMP) {
                inv_mb = frcp(s_c[0][1] + mp);                      // 1/(M + m): serves mu^-1 and the mass ratio
                if constexpr (!FAST_DAMP) o = elements<EINC>(s_c[0][0], s_c[0][kC - 1], s_c[0][1], mp, inv_mb, g);
                mr = mp*inv_mb;
            }
            Vec3 a = {ax[v], ay[v], az[v]};
            Vec3 fd = {0., 0., 0.};                                 // sum of the damping forces of the stack
            if constexpr (FAST_DAMP) {
                // the whole stack in one basic block (damping_fast), flagged particles redone by the branching functions
                constexpr int SMO = slot_of<REBX_B200_MODIFY_ORBITS, K0, K1, K2>(), SGD = slot_of<REBX_B200_GAS_DAMPING, K0, K1, K2>(),
                              STI = slot_of<REBX_B200_TYPE_I, K0, K1, K2>();
                auto tabv = [&](int slot, int j) -> double {
                    return slot == 0 ? q0.t[j < Traits<K0>::ntab ? j : 0][v] : (slot == 1 ? q1.t[j < Traits<K1>::ntab ? j : 0][v] : q2.t[j < Traits<K2>::ntab ? j : 0][v]);
                };
                const double tau_a = SMO >= 0 ? tabv(SMO, 0) : 0.0, tau_e = SMO >= 0 ? tabv(SMO, 1) : 0.0, tau_inc = SMO >= 0 ? tabv(SMO, 2) : 0.0;
                const double d_factor = SGD >= 0 ? tabv(SGD, 0) : 0.0;
                const int mo_flags = SMO >= 0 ? P.fx[SMO >= 0 ? SMO : 0].flags : 0;
                const double* cmo = s_c[SMO >= 0 ? SMO : 0]; const double* cgd = s_c[SGD >= 0 ? SGD : 0]; const double* cti = s_c[STI >= 0 ? STI : 0];
                unsigned bad = 0, band = 0;
                double a_el;
                const BodyK bk = {s_c[0][1], cgd[3], cgd[4], cti[17]};      // entries of an absent effect are never read
                const DampCoef dc = damping_fast<(SMO >= 0), (SGD >= 0), (STI >= 0)>(cmo, cgd, cti, mo_flags, bk, g, mp, inv_mb, tau_a, tau_e, tau_inc, d_factor, a_el, bad, band);
                double T_mo = 1.0, T_ti = 1.0;
                if constexpr (SMO >= 0) { if (mo_flags & REBX_B200_FLAG_TRAP) T_mo = (a_el > cmo[13]) ? 1.0 : -10.0; }
                if constexpr (STI >= 0) T_ti = (a_el > cti[13]) ? 1.0 : -10.0;
                if (band) {             // inside a trap's transition band: the axis with the reference's exact operation sequence
                    const double ax_exact = exact_semimajor_axis(s_c[0][0], s_c[0][1], mp, g.dx, g.dy, g.dz, g.dvx, g.dvy, g.dvz);
                    if (band & 1u) T_mo = planet_trap(ax_exact, cmo[3], cmo[4]);
                    if (band & 2u) T_ti = planet_trap(ax_exact, cti[8], cti[9]);
                }
                const double cv = dc.cv_mo*T_mo + dc.cv_ti*T_ti;
                fd.x = g.dvx*cv + g.dx*dc.cr;
                fd.y = g.dvy*cv + g.dy*dc.cr;
                fd.z = g.dvz*(cv + dc.cz) + g.dz*dc.cr;
                if (bad) damping_generic<K0, K1, K2>(&P, s_c, &g, mp, inv_mb, tau_a, tau_e, tau_inc, d_factor, &fd);
            } else {
            apply<K0, V>(P.fx[0], s_c[0], g, o, mr, mp, rp, q0, v, is_damping<K0>() ? fd : a, red0);
            if constexpr (K1 != 0) apply<K1, V>(P.fx[1], s_c[1], g, o, mr, mp, rp, q1, v, is_damping<K1>() ? fd : a, red1);
            if constexpr (K2 != 0) apply<K2, V>(P.fx[2], s_c[2], g, o, mr, mp, rp, q2, v, is_damping<K2>() ? fd : a, red2);
            }
            if constexpr (DAMP) {
                // rebx_com_force, PARTICLE coordinates (src/rebxtools.c:131-141), once for the summed force:
                // p.a += f - mr f ; body.a -= mr f with mr = m/(M + m).  The back-reaction of the whole stack is
                // carried by its first damping effect's sum (the host adds the effects' sums to the body anyway).
                const double kx = mr*fd.x, ky = mr*fd.y, kz = mr*fd.z;
                a.x += fd.x - kx; a.y += fd.y - ky; a.z += fd.z - kz;
                Vec3& red = is_damping<K0>() ? red0 : (is_damping<K1>() ? red1 : red2);
                red.x -= kx; red.y -= ky; red.z -= kz;
            }
            ax[v] = a.x; ay[v] = a.y; az[v] = a.z;
        }
        st<V>(P.st.ax, k, full, ax); st<V>(P.st.ay, k, full, ay); st<V>(P.st.az, k, full, az);
    }
    finish_sums<T, K0, K1, K2, 0>(red0, red1, red2, Vec3{0., 0., 0.}, P, partials, ApplyTail{});
}

// ---- JACOBI coordinates: the sweep of the centre-of-mass pipeline in tolerance arithmetic ----------------------------
// Same interface and the same scans as jacobi_sweep_kernel (kernels_jacobi.cuh; reference src/rebxtools.c:92-128 as two
// O(N) scans): the interior MASS of particle k still carries the reference's own rounding, bit for bit (mass grid,
// dip, literal recurrences) -- mu = G (m + M_k) enters v^2 - mu/d, a difference of nearly equal numbers.  What
// changes is everything behind it: COM_k = S_k * (1/M_k) by a Newton reciprocal instead of six IEEE divisions, and the
// stacked damping forces by damping_fast with the body-mass constants (BodyK) worked out per particle.
template <int K0, int K1, int K2>
__global__ void __launch_bounds__(kScanThreads, 2)
jacobi_sweep_tol_kernel(const __grid_constant__ rebx_b200_pass P, const double* __restrict__ tile_prefix,
                        double* __restrict__ tvec, double* __restrict__ tile_tsums, const double* __restrict__ carry7,
                        const MassGrid* __restrict__ grid, const double* __restrict__ status, const double* __restrict__ mref, int ntiles) {
    constexpr int SMO = slot_of<REBX_B200_MODIFY_ORBITS, K0, K1, K2>(), SGD = slot_of<REBX_B200_GAS_DAMPING, K0, K1, K2>(),
                  STI = slot_of<REBX_B200_TYPE_I, K0, K1, K2>();
    __shared__ double s_warp[kScanThreads/32][7];
    __shared__ double s_c[3][kC];
    __shared__ double s_unit[REBX_B200_BODY_DOUBLES];        // a unit-mass body at rest: the body-independent constants only
    if (threadIdx.x < REBX_B200_BODY_DOUBLES) s_unit[threadIdx.x] = (threadIdx.x == REBX_B200_BODY_M) ? 1.0 : 0.0;
    __syncthreads();
    if (threadIdx.x < P.n_effects) fill_constants(P.fx[threadIdx.x], s_unit, s_c[threadIdx.x]);
    __syncthreads();
    const rebx_b200_state& st = P.st;
    const MassGrid mg = *grid;
    const bool literal = mref != nullptr && status[0] != 0.0;
    const double G = s_c[0][0];
    // Resident grid, a CTA walks the tiles (one tile = kScanThreads consecutive particles, as in the scans around this
    // kernel): the constants are worked out once per CTA, not once per tile (measured with one CTA per tile: 107 us at
    // 10^6 particles against 48 us for the same forces in PARTICLE coordinates).
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t k = (int64_t)tile*kScanThreads + threadIdx.x;
        const bool live = k < st.n;
        double px = 0, py = 0, pz = 0, pvx = 0, pvy = 0, pvz = 0, pm = 0;
        double tau_a = 0.0, tau_e = 0.0, tau_inc = 0.0, d_factor = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0;
        if (live) {
            px = __ldcs(st.x + k); py = __ldcs(st.y + k); pz = __ldcs(st.z + k); pvx = __ldcs(st.vx + k); pvy = __ldcs(st.vy + k); pvz = __ldcs(st.vz + k); pm = __ldcs(st.m + k);
            if constexpr (SMO >= 0) { tau_a = __ldcs(P.fx[SMO >= 0 ? SMO : 0].tab[0] + k); tau_e = __ldcs(P.fx[SMO >= 0 ? SMO : 0].tab[1] + k); tau_inc = __ldcs(P.fx[SMO >= 0 ? SMO : 0].tab[2] + k); }
            if constexpr (SGD >= 0) d_factor = __ldcs(P.fx[SGD >= 0 ? SGD : 0].tab[0] + k);
            a0 = __ldcs(st.ax + k); a1 = __ldcs(st.ay + k); a2 = __ldcs(st.az + k);
        }
        bool unused_flag = false;
        double v[7] = {live ? quantise_mass(pm, mg, unused_flag) : 0.0, px*pm, py*pm, pz*pm, pvx*pm, pvy*pm, pvz*pm};
        double total[7];
        block_exclusive_scan<7>(v, total, s_warp);
        double M = tile_prefix[tile] + v[0];
        if (carry7) M = carry7[0] + M;
        if (mg.ok && M == mg.lead && st.index_base + k >= 1) M = mg.dip;
        if (literal && live) M = mref[k];
        double S[6];
#pragma unroll
        for (int j = 0; j < 6; j++) {
            S[j] = tile_prefix[(size_t)(1 + j)*ntiles + tile] + v[1 + j];
            if (carry7) S[j] = carry7[1 + j] + S[j];
        }
        double t[3] = {0.0, 0.0, 0.0};
        const long long gi = st.index_base + k;
        if (live && gi != 0) {
            const double inv_M = (M > 0.) ? frcp(M) : 1.0;
            Geo g;
            g.dx = px - S[0]*inv_M; g.dy = py - S[1]*inv_M; g.dz = pz - S[2]*inv_M;
            g.r2 = g.dx*g.dx + g.dy*g.dy + g.dz*g.dz;
            fsqrt2(g.r2, g.r, g.rinv);
            g.rinv2 = g.rinv*g.rinv;
            g.dvx = pvx - S[3]*inv_M; g.dvy = pvy - S[4]*inv_M; g.dvz = pvz - S[5]*inv_M;
            g.rv = g.dx*g.dvx + g.dy*g.dvy + g.dz*g.dvz;
            const double inv_mb = frcp(M + pm), mr = pm*inv_mb;
            BodyK bk;
            bk.bm = M;
            fsqrt2(G*M, bk.sqrtGM, bk.rsqrtGM);
            const double rsm = bk.rsqrtGM*s_c[0][kC - 2];           // 1/sqrt(M) = sqrt(G)/sqrt(G M)
            bk.rsqrtM3 = rsm*rsm*rsm;
            const int mo_flags = SMO >= 0 ? P.fx[SMO >= 0 ? SMO : 0].flags : 0;
            const double* cmo = s_c[SMO >= 0 ? SMO : 0]; const double* cgd = s_c[SGD >= 0 ? SGD : 0]; const double* cti = s_c[STI >= 0 ? STI : 0];
            unsigned bad = 0, band = 0;
            double a_el;
            const DampCoef dc = damping_fast<(SMO >= 0), (SGD >= 0), (STI >= 0)>(cmo, cgd, cti, mo_flags, bk, g, pm, inv_mb, tau_a, tau_e, tau_inc, d_factor, a_el, bad, band);
            double T_mo = 1.0, T_ti = 1.0;
            if constexpr (SMO >= 0) { if (mo_flags & REBX_B200_FLAG_TRAP) T_mo = (a_el > cmo[13]) ? 1.0 : -10.0; }
            if constexpr (STI >= 0) T_ti = (a_el > cti[13]) ? 1.0 : -10.0;
            if (band) {
                const double ax_exact = exact_semimajor_axis(G, M, pm, g.dx, g.dy, g.dz, g.dvx, g.dvy, g.dvz);
                if (band & 1u) T_mo = planet_trap(ax_exact, cmo[3], cmo[4]);
                if (band & 2u) T_ti = planet_trap(ax_exact, cti[8], cti[9]);
            }
            const double cv = dc.cv_mo*T_mo + dc.cv_ti*T_ti;
            Vec3 fd;
            fd.x = g.dvx*cv + g.dx*dc.cr;
            fd.y = g.dvy*cv + g.dy*dc.cr;
            fd.z = g.dvz*(cv + dc.cz) + g.dz*dc.cr;
            if (bad) {
                // flagged: the branching functions, with this particle's own body-mass constants
                double body[REBX_B200_BODY_DOUBLES];
                for (int j = 0; j < REBX_B200_BODY_DOUBLES; j++) body[j] = 0.0;
                body[REBX_B200_BODY_M] = M;
                double c_loc[3][kC];
                for (int j = 0; j < P.n_effects; j++) fill_constants(P.fx[j], body, c_loc[j]);
                damping_generic<K0, K1, K2>(&P, c_loc, &g, pm, inv_mb, tau_a, tau_e, tau_inc, d_factor, &fd);
            }
            __stcs(st.ax + k, a0 + fd.x); __stcs(st.ay + k, a1 + fd.y); __stcs(st.az + k, a2 + fd.z);
            t[0] = mr*fd.x; t[1] = mr*fd.y; t[2] = mr*fd.z;
        }
        if (live) { __stcs(tvec + k, t[0]); __stcs(tvec + st.n + k, t[1]); __stcs(tvec + 2*st.n + k, t[2]); }
        double tt[3], tot3[3];
        tt[0] = t[0]; tt[1] = t[1]; tt[2] = t[2];
        block_exclusive_scan<3>(tt, tot3, reinterpret_cast<double (*)[3]>(&s_warp[0][0]));
#pragma unroll
        for (int c = 0; c < 3; c++) if (threadIdx.x == c) tile_tsums[(size_t)c*ntiles + tile] = tot3[c];
    }
}

// ---- fused Wisdom-Holman kick + Kepler drift in tolerance arithmetic ---------------------------------------------------
// The tolerance counterpart of wh_kick_drift_kernel (kernels_kepler.cuh; the work reference src/steppers.c:79-87 hands to
// REBOUND's WHFast): a = ONE effect without back-reaction at the current state, v += dt_kick a, Kepler drift about the body
// for dt_drift.  Same algorithm as kepler_core.h -- universal variables, Newton on r0 G1 + eta G2 + mu G3 = dt, Stumpff
// functions by quartering + series + duplication -- evaluated with FMA contraction, Newton reciprocals instead of IEEE
// divisions, and the distance / reciprocal distance the force already has (one reciprocal square root per particle and step).
__device__ __forceinline__ void stumpff_tol(double z, double& c0, double& c1, double& c2, double& c3) {
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
    c0 = C0; c1 = C1; c2 = C2; c3 = C3;
}

__device__ __forceinline__ void kepler_drift_tol(double mu, double dt, double r0, double inv_r0, double (&x)[3], double (&v)[3]) {
    const double v2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
    const double eta = x[0]*v[0] + x[1]*v[1] + x[2]*v[2];
    const double beta = 2.*mu*inv_r0 - v2;
    const double X1 = dt*inv_r0;
    double X = X1*(1. - 0.5*eta*X1*inv_r0), c0 = 1., c1 = 1., c2 = 0.5, c3 = 1./6., G1 = 0., G2 = 0., G3 = 0., rr = r0;
    double dX = 0.;
    bool converged = false;
    for (int it = 0; it < 60; it++) {
        stumpff_tol(beta*X*X, c0, c1, c2, c3);
        G1 = c1*X; G2 = c2*X*X; G3 = c3*X*X*X;
        const double f = r0*G1 + eta*G2 + mu*G3 - dt;
        rr = r0*c0 + eta*G1 + mu*G2;
        dX = f*frcp(rr);
        X -= dX;
        if (fabs(dX) <= 1e-9*fabs(X)) { converged = true; break; }
    }
    if (!converged) {
        stumpff_tol(beta*X*X, c0, c1, c2, c3);
        G1 = c1*X; G2 = c2*X*X; G3 = c3*X*X*X;
    } else {                                   // first order in the last (tiny) correction, as in kepler_core.h
        const double G0 = c0;
        G3 = G3 - dX*G2;
        G2 = G2 - dX*G1;
        const double G1n = G1 - dX*G0;
        c0 = G0 + dX*beta*G1;
        G1 = G1n;
    }
    rr = r0*c0 + eta*G1 + mu*G2;
    const double inv_rr = frcp(rr);
    const double f = 1. - mu*G2*inv_r0, g = dt - mu*G3;
    const double fd = -mu*G1*inv_r0*inv_rr, gd = 1. - mu*G2*inv_rr;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double xn = f*x[c] + g*v[c], vn = fd*x[c] + gd*v[c];
        x[c] = xn; v[c] = vn;
    }
}

#ifndef REBXB_TOL_WH_BLOCKS
#define REBXB_TOL_WH_BLOCKS 3
#endif
template <int K>
__global__ void __launch_bounds__(256, K == REBX_B200_RADIATION ? REBXB_TOL_WH_BLOCKS : 1)
wh_kick_drift_tol_kernel(const __grid_constant__ rebx_b200_pass P, double G, double dt_kick, double dt_drift) {
    static_assert(!Traits<K>::red, "the fused kick+drift needs an effect without back-reaction");
    __shared__ double s_c[kC];
    __shared__ double s_org[8];
    if (threadIdx.x == 0) fill_constants(P.fx[0], P.fx[0].body, s_c);
    if (threadIdx.x < 8) s_org[threadIdx.x] = P.fx[0].body[threadIdx.x];
    __syncthreads();
    const rebx_b200_state& st = P.st;
    const long long bi = (long long)s_org[REBX_B200_BODY_INDEX];
    const double ox = s_org[REBX_B200_BODY_X], oy = s_org[REBX_B200_BODY_X+1], oz = s_org[REBX_B200_BODY_X+2];
    const double ovx = s_org[REBX_B200_BODY_VX], ovy = s_org[REBX_B200_BODY_VX+1], ovz = s_org[REBX_B200_BODY_VX+2];
    const double om = s_org[REBX_B200_BODY_M];
    const double nx = ox + dt_drift*ovx, ny = oy + dt_drift*ovy, nz = oz + dt_drift*ovz;      // the body after its inertial drift
    double* const X = const_cast<double*>(st.x);  double* const Y = const_cast<double*>(st.y);  double* const Z = const_cast<double*>(st.z);
    double* const VX = const_cast<double*>(st.vx); double* const VY = const_cast<double*>(st.vy); double* const VZ = const_cast<double*>(st.vz);
    const int64_t stride = (int64_t)gridDim.x*256;
    for (int64_t k = (int64_t)blockIdx.x*256 + threadIdx.x; k < st.n; k += stride) {
        if (st.index_base + k == bi) {          // the body: no force on itself, no back-reaction => inertial
            X[k] = nx; Y[k] = ny; Z[k] = nz;
            continue;
        }
        const double px = __ldcs(X + k), py = __ldcs(Y + k), pz = __ldcs(Z + k);
        const double pvx = __ldcs(VX + k), pvy = __ldcs(VY + k), pvz = __ldcs(VZ + k);
        const double pm = st.m ? __ldcs(st.m + k) : 0.0;
        double pr = 0.0;
        if constexpr (Traits<K>::rad) pr = __ldcs(st.r + k);
        Params<K, 1> q;
        load_params<K, 1>(P.fx[0], k, true, q);
        Geo g;
        g.dx = px - ox; g.dy = py - oy; g.dz = pz - oz;
        g.r2 = g.dx*g.dx + g.dy*g.dy + g.dz*g.dz;
        fsqrt2(g.r2, g.r, g.rinv);
        g.rinv2 = g.rinv*g.rinv;
        g.dvx = pvx - ovx; g.dvy = pvy - ovy; g.dvz = pvz - ovz;
        g.rv = g.dx*g.dvx + g.dy*g.dvy + g.dz*g.dvz;
        Vec3 a = {0.0, 0.0, 0.0}, red = {0.0, 0.0, 0.0};
        const Elements o = {0., 0., 0., 0., 0.};
        apply<K, 1>(P.fx[0], s_c, g, o, 0.0, pm, pr, q, 0, a, red);
        double v[3] = {g.dvx + dt_kick*a.x, g.dvy + dt_kick*a.y, g.dvz + dt_kick*a.z};
        double x[3] = {g.dx, g.dy, g.dz};
        kepler_drift_tol(G*(om + pm), dt_drift, g.r, g.rinv, x, v);
        __stcs(X + k, nx + x[0]); __stcs(Y + k, ny + x[1]); __stcs(Z + k, nz + x[2]);
        __stcs(VX + k, ovx + v[0]); __stcs(VY + k, ovy + v[1]); __stcs(VZ + k, ovz + v[2]);
    }
}

// ---- fused drift-kick-drift step in tolerance arithmetic -----------------------------------------------------------------
// The tolerance counterpart of step_kernel (kernels.cu) for ONE effect without back-reaction (radiation_forces,
// yarkovsky_effect): x += dt/2 v; a = pull of the central body + the effect; v += dt a; x += dt/2 v -- one sweep, the
// acceleration lives in registers.  One reciprocal square root per particle serves the central gravity and the effect.
// The central body itself is skipped here and advanced by step_body_kernel, exactly as in the bit-exact path.
template <int K>
__global__ void __launch_bounds__(256, K == REBX_B200_RADIATION ? 3 : 1)
step_tol_kernel(const __grid_constant__ rebx_b200_pass P, double dt, double hdt, double G, double soft2) {
    static_assert(!Traits<K>::red, "the tolerance step needs an effect without back-reaction");
    __shared__ double s_c[kC];
    __shared__ double s_org[8];
    if (threadIdx.x < 8) s_org[threadIdx.x] = P.fx[0].body[threadIdx.x];
    __syncthreads();
    if (threadIdx.x < 3) s_org[REBX_B200_BODY_X + threadIdx.x] += hdt*s_org[REBX_B200_BODY_VX + threadIdx.x];     // the body's own first half drift
    __syncthreads();
    if (threadIdx.x == 0) fill_constants(P.fx[0], s_org, s_c);
    __syncthreads();
    const rebx_b200_state& st = P.st;
    const long long bi = (long long)s_org[REBX_B200_BODY_INDEX];
    const double ox = s_org[REBX_B200_BODY_X], oy = s_org[REBX_B200_BODY_X+1], oz = s_org[REBX_B200_BODY_X+2];
    const double ovx = s_org[REBX_B200_BODY_VX], ovy = s_org[REBX_B200_BODY_VX+1], ovz = s_org[REBX_B200_BODY_VX+2];
    const double GM = G*s_org[REBX_B200_BODY_M];
    double* const X = const_cast<double*>(st.x);  double* const Y = const_cast<double*>(st.y);  double* const Z = const_cast<double*>(st.z);
    double* const VX = const_cast<double*>(st.vx); double* const VY = const_cast<double*>(st.vy); double* const VZ = const_cast<double*>(st.vz);
    const int64_t stride = (int64_t)gridDim.x*256;
    for (int64_t k = (int64_t)blockIdx.x*256 + threadIdx.x; k < st.n; k += stride) {
        if (st.index_base + k == bi) continue;                       // the central body: step_body_kernel
        const double pvx = __ldcs(VX + k), pvy = __ldcs(VY + k), pvz = __ldcs(VZ + k);
        const double px = __ldcs(X + k) + hdt*pvx, py = __ldcs(Y + k) + hdt*pvy, pz = __ldcs(Z + k) + hdt*pvz;
        const double pm = st.m ? __ldcs(st.m + k) : 0.0;
        double pr = 0.0;
        if constexpr (Traits<K>::rad) pr = __ldcs(st.r + k);
        Params<K, 1> q;
        load_params<K, 1>(P.fx[0], k, true, q);
        Geo g;
        g.dx = px - ox; g.dy = py - oy; g.dz = pz - oz;
        g.r2 = g.dx*g.dx + g.dy*g.dy + g.dz*g.dz;
        fsqrt2(g.r2, g.r, g.rinv);
        g.rinv2 = g.rinv*g.rinv;
        g.dvx = pvx - ovx; g.dvy = pvy - ovy; g.dvz = pvz - ovz;
        g.rv = g.dx*g.dvx + g.dy*g.dvy + g.dz*g.dvz;
        Vec3 a = {0.0, 0.0, 0.0}, red = {0.0, 0.0, 0.0};
        if (GM != 0.0) {                                             // central gravity
            double ri3;
            if (soft2 == 0.0) ri3 = g.rinv2*g.rinv;
            else { const double rs = frsqrt(g.r2 + soft2); ri3 = rs*rs*rs; }
            const double pref = -GM*ri3;
            a.x = pref*g.dx; a.y = pref*g.dy; a.z = pref*g.dz;
        }
        const Elements o = {0., 0., 0., 0., 0.};
        apply<K, 1>(P.fx[0], s_c, g, o, 0.0, pm, pr, q, 0, a, red);
        const double nvx = pvx + dt*a.x, nvy = pvy + dt*a.y, nvz = pvz + dt*a.z;
        __stcs(VX + k, nvx); __stcs(VY + k, nvy); __stcs(VZ + k, nvz);
        __stcs(X + k, px + hdt*nvx); __stcs(Y + k, py + hdt*nvy); __stcs(Z + k, pz + hdt*nvz);
    }
}

// ---- the whole Jacobi pipeline in one cooperative launch (jacobi_fused.cuh), tolerance arithmetic --------------------
// The per-particle part of jacobi_sweep_tol_kernel as the Force functor of jacobi_fused_body.
template <int K0, int K1, int K2>
struct JacobiTolForce {
    static constexpr int SMO = slot_of<REBX_B200_MODIFY_ORBITS, K0, K1, K2>(), SGD = slot_of<REBX_B200_GAS_DAMPING, K0, K1, K2>(),
                         STI = slot_of<REBX_B200_TYPE_I, K0, K1, K2>();
    struct Ops { double tau_a, tau_e, tau_inc, d_factor; };
    double (*s_c)[kC];
    __device__ __forceinline__ void init(const rebx_b200_pass& P) {
        __shared__ double sh_c[3][kC];
        __shared__ double sh_unit[REBX_B200_BODY_DOUBLES];      // a unit-mass body at rest: the body-independent constants only
        if (threadIdx.x < REBX_B200_BODY_DOUBLES) sh_unit[threadIdx.x] = (threadIdx.x == REBX_B200_BODY_M) ? 1.0 : 0.0;
        __syncthreads();
        if (threadIdx.x < P.n_effects) fill_constants(P.fx[threadIdx.x], sh_unit, sh_c[threadIdx.x]);
        s_c = sh_c;
    }
    __device__ __forceinline__ Ops load(const rebx_b200_pass& P, int64_t k) {
        Ops o = {0.0, 0.0, 0.0, 0.0};
        if constexpr (SMO >= 0) { o.tau_a = __ldcs(P.fx[SMO >= 0 ? SMO : 0].tab[0] + k); o.tau_e = __ldcs(P.fx[SMO >= 0 ? SMO : 0].tab[1] + k); o.tau_inc = __ldcs(P.fx[SMO >= 0 ? SMO : 0].tab[2] + k); }
        if constexpr (SGD >= 0) o.d_factor = __ldcs(P.fx[SGD >= 0 ? SGD : 0].tab[0] + k);
        return o;
    }
    __device__ __forceinline__ void eval(const rebx_b200_pass& P, const Ops& o, int64_t, double px, double py, double pz, double pvx, double pvy, double pvz,
                                         double pm, double M, const double (&S)[6], Vec3& a, double (&t)[3]) {
        const double G = s_c[0][0];
        const double inv_M = (M > 0.) ? frcp(M) : 1.0;
        Geo g;
        g.dx = px - S[0]*inv_M; g.dy = py - S[1]*inv_M; g.dz = pz - S[2]*inv_M;
        g.r2 = g.dx*g.dx + g.dy*g.dy + g.dz*g.dz;
        fsqrt2(g.r2, g.r, g.rinv);
        g.rinv2 = g.rinv*g.rinv;
        g.dvx = pvx - S[3]*inv_M; g.dvy = pvy - S[4]*inv_M; g.dvz = pvz - S[5]*inv_M;
        g.rv = g.dx*g.dvx + g.dy*g.dvy + g.dz*g.dvz;
        const double inv_mb = frcp(M + pm), mr = pm*inv_mb;
        BodyK bk;
        bk.bm = M;
        fsqrt2(G*M, bk.sqrtGM, bk.rsqrtGM);
        const double rsm = bk.rsqrtGM*s_c[0][kC - 2];           // 1/sqrt(M) = sqrt(G)/sqrt(G M)
        bk.rsqrtM3 = rsm*rsm*rsm;
        const int mo_flags = SMO >= 0 ? P.fx[SMO >= 0 ? SMO : 0].flags : 0;
        const double* cmo = s_c[SMO >= 0 ? SMO : 0]; const double* cgd = s_c[SGD >= 0 ? SGD : 0]; const double* cti = s_c[STI >= 0 ? STI : 0];
        unsigned bad = 0, band = 0;
        double a_el;
        const DampCoef dc = damping_fast<(SMO >= 0), (SGD >= 0), (STI >= 0)>(cmo, cgd, cti, mo_flags, bk, g, pm, inv_mb, o.tau_a, o.tau_e, o.tau_inc, o.d_factor, a_el, bad, band);
        double T_mo = 1.0, T_ti = 1.0;
        if constexpr (SMO >= 0) { if (mo_flags & REBX_B200_FLAG_TRAP) T_mo = (a_el > cmo[13]) ? 1.0 : -10.0; }
        if constexpr (STI >= 0) T_ti = (a_el > cti[13]) ? 1.0 : -10.0;
        if (band) {
            const double ax_exact = exact_semimajor_axis(G, M, pm, g.dx, g.dy, g.dz, g.dvx, g.dvy, g.dvz);
            if (band & 1u) T_mo = planet_trap(ax_exact, cmo[3], cmo[4]);
            if (band & 2u) T_ti = planet_trap(ax_exact, cti[8], cti[9]);
        }
        const double cv = dc.cv_mo*T_mo + dc.cv_ti*T_ti;
        Vec3 fd;
        fd.x = g.dvx*cv + g.dx*dc.cr;
        fd.y = g.dvy*cv + g.dy*dc.cr;
        fd.z = g.dvz*(cv + dc.cz) + g.dz*dc.cr;
        if (bad) {
            // flagged: the branching functions, with this particle's own body-mass constants
            double body[REBX_B200_BODY_DOUBLES];
            for (int j = 0; j < REBX_B200_BODY_DOUBLES; j++) body[j] = 0.0;
            body[REBX_B200_BODY_M] = M;
            double c_loc[3][kC];
            for (int j = 0; j < P.n_effects; j++) fill_constants(P.fx[j], body, c_loc[j]);
            damping_generic<K0, K1, K2>(&P, c_loc, &g, pm, inv_mb, o.tau_a, o.tau_e, o.tau_inc, o.d_factor, &fd);
        }
        a.x += fd.x; a.y += fd.y; a.z += fd.z;
        t[0] = mr*fd.x; t[1] = mr*fd.y; t[2] = mr*fd.z;
    }
};

template <int K0, int K1, int K2>
__global__ void __launch_bounds__(kScanThreads, 2)
jacobi_fused_tol_kernel(const __grid_constant__ rebx_b200_pass P, const FusedScratch F, const int wtiles) {
    jacobi_fused_body<JacobiTolForce<K0, K1, K2>>(P, F, wtiles);
}

typedef void (*jacobi_tol_fn)(const rebx_b200_pass, const double*, double*, double*, const double*, const MassGrid*, const double*, const double*, int);
typedef void (*jacobi_fused_tol_fn)(const rebx_b200_pass, const FusedScratch, const int);
struct JacobiTolCombo { int k[3]; jacobi_tol_fn fn; jacobi_fused_tol_fn fused; };
#define REBXB_JTCOMBO(a, b, c) { {a, b, c}, jacobi_sweep_tol_kernel<a, b, c>, jacobi_fused_tol_kernel<a, b, c> },
static const JacobiTolCombo kJacobiTolCombos[] = {
    REBXB_JTCOMBO(5,0,0) REBXB_JTCOMBO(6,0,0) REBXB_JTCOMBO(7,0,0)
    REBXB_JTCOMBO(5,6,7) REBXB_JTCOMBO(7,6,5)
};

typedef void (*tol_fn)(const rebx_b200_pass, double*);
struct TolCombo { int k[3]; tol_fn fn[2]; int threads; int smem[2]; };
#define REBXB_TCOMBO(a, b, c) { {a, b, c}, { pass_kernel_tol<1, a, b, c>, pass_kernel_tol<2, a, b, c> }, tol_threads<a, b, c>(), \
                                { staged_bytes<1, a, b, c>(), staged_bytes<2, a, b, c>() } },
static const TolCombo kTolCombos[] = {
    REBXB_TCOMBO(1,0,0) REBXB_TCOMBO(2,0,0) REBXB_TCOMBO(3,0,0) REBXB_TCOMBO(4,0,0) REBXB_TCOMBO(5,0,0) REBXB_TCOMBO(6,0,0) REBXB_TCOMBO(7,0,0)
    REBXB_TCOMBO(2,3,0) REBXB_TCOMBO(3,2,0) REBXB_TCOMBO(4,3,0) REBXB_TCOMBO(3,4,0) REBXB_TCOMBO(1,3,0) REBXB_TCOMBO(3,1,0)
    REBXB_TCOMBO(7,6,5) REBXB_TCOMBO(5,6,7)
};

}  // namespace tol

// Launches the tolerance-mode sweep of `sub` (effects 0..take-1, all referring to the same central body in PARTICLE
// coordinates).  Returns false when the stack has no tolerance instantiation (the caller then uses the exact
// kernels); on true *grid_out is the grid (the back-reaction partials are laid out as by pass_kernel).
bool tol_launch(const rebx_b200_pass& sub, int take, double* workspace, cudaStream_t s, int* grid_out) {
    if (take < 1 || take > 3) return false;
    for (int j = 0; j < take; j++) if (sub.fx[j].flags & REBX_B200_FLAG_BARYCENTRIC) return false;
    const tol::TolCombo* combo = nullptr;
    for (const tol::TolCombo& c : tol::kTolCombos) {
        bool ok = true;
        for (int j = 0; j < 3; j++) if (c.k[j] != (j < take ? sub.fx[j].kind : 0)) ok = false;
        if (ok) { combo = &c; break; }
    }
    if (!combo) return false;
    bool heavy = false;
    for (int j = 0; j < take; j++) if (sub.fx[j].kind >= REBX_B200_YARKOVSKY) heavy = true;
    // two particles per thread (double2 access) for the streaming stacks, one for the register-heavy ones;
    // REBX_B200_TOL_V=1|2 overrides (tuning)
    static const int forced = [] { const char* e = getenv("REBX_B200_TOL_V"); return e ? atoi(e) : 0; }();
    // REBX_B200_TOL_GRID=full: one tile per thread (as many CTAs as it takes, up to the partials' capacity) instead of
    // a resident grid with a grid-stride loop: the hardware scheduler then balances the tail of a short heavy sweep
    static const int full_grid = [] { const char* e = getenv("REBX_B200_TOL_GRID"); return (e && e[0] == 'f') ? 1 : 0; }();
    // Measured (tools/tol_sweep.sh, profiles/r02_tol_variants.txt): the streaming stacks want two particles per thread
    // (double2 access) and all 128 registers -- 256 x 2 CTAs: J2/J4 (+) gr_potential 0.134 ms against 0.165 with one
    // particle per thread at 80 registers; the stacks with transcendentals want one particle per thread.
    int vsel = (pass_is_vectorizable(sub) && !heavy) ? 1 : 0;
    if (forced == 1) vsel = 0;
    if (forced == 2 && pass_is_vectorizable(sub)) vsel = 1;
    const int V = vsel ? 2 : 1;
    tol::tol_fn fn = combo->fn[vsel];
    const int64_t tiles = (sub.st.n + V - 1)/V;
    const int smem = combo->smem[vsel];
    if (smem > 48*1024) {
        // opt in to more than 48 KB of dynamic shared memory: once per kernel and device
        static thread_local std::vector<std::pair<const void*, int>> done;
        int dev = -1; cudaGetDevice(&dev);
        bool seen = false;
        for (const auto& d : done) if (d.first == reinterpret_cast<const void*>(fn) && d.second == dev) seen = true;
        if (!seen) {
            if (cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { cudaGetLastError(); return false; }
            done.push_back({reinterpret_cast<const void*>(fn), dev});
        }
    }
    int grid = grid_for(reinterpret_cast<const void*>(fn), tiles, combo->threads, smem, combo->threads);
    if (full_grid) { const int64_t want = (tiles + combo->threads - 1)/combo->threads; grid = (int)(want < kMaxBlocks ? want : kMaxBlocks); }
    fn<<<grid, combo->threads, smem, s>>>(sub, workspace);
    *grid_out = grid;
    return true;
}

// The tolerance-mode Jacobi sweep for the stack of `P` (kinds 5, 6, 7), or false when it has no instantiation here (the
// caller then launches the exact jacobi_sweep_kernel).  Arguments as for that kernel.
bool tol_jacobi_launch(const rebx_b200_pass& P, unsigned ntiles, const double* tile_prefix, double* tvec, double* tile_tsums, const double* carry7,
                       const void* grid, const double* status, const double* mref, cudaStream_t s) {
    for (const tol::JacobiTolCombo& c : tol::kJacobiTolCombos) {
        bool ok = true;
        for (int j = 0; j < 3; j++) if (c.k[j] != (j < P.n_effects ? P.fx[j].kind : 0)) ok = false;
        if (!ok) continue;
        const int grid_ctas = grid_for(reinterpret_cast<const void*>(c.fn), (int64_t)ntiles*kScanThreads, kScanThreads, 0, kScanThreads);
        c.fn<<<grid_ctas, kScanThreads, 0, s>>>(P, tile_prefix, tvec, tile_tsums, carry7, static_cast<const MassGrid*>(grid), status, mref, (int)ntiles);
        return true;
    }
    return false;
}

// The fused Wisdom-Holman kick + Kepler drift in tolerance arithmetic; false when the effect has no instantiation here.
bool tol_wh_kick_drift_launch(const rebx_b200_pass& P, double G, double dt_kick, double dt_drift, cudaStream_t s) {
    if (P.n_effects != 1 || (P.fx[0].flags & REBX_B200_FLAG_BARYCENTRIC)) return false;
    if (P.fx[0].kind == REBX_B200_RADIATION) {
        const int grid = grid_for(reinterpret_cast<const void*>(tol::wh_kick_drift_tol_kernel<REBX_B200_RADIATION>), P.st.n, 256, 0, 256);
        tol::wh_kick_drift_tol_kernel<REBX_B200_RADIATION><<<grid, 256, 0, s>>>(P, G, dt_kick, dt_drift);
        return true;
    }
    if (P.fx[0].kind == REBX_B200_YARKOVSKY) {
        const int grid = grid_for(reinterpret_cast<const void*>(tol::wh_kick_drift_tol_kernel<REBX_B200_YARKOVSKY>), P.st.n, 256, 0, 256);
        tol::wh_kick_drift_tol_kernel<REBX_B200_YARKOVSKY><<<grid, 256, 0, s>>>(P, G, dt_kick, dt_drift);
        return true;
    }
    return false;
}

// The fused drift-kick-drift sweep in tolerance arithmetic (the particles only; the body follows in step_body_kernel);
// false when the stack has no instantiation here (anything but one effect without back-reaction).
bool tol_leapfrog_sweep_launch(const rebx_b200_pass& P, double dt, double G, double soft2, cudaStream_t s) {
    if (P.n_effects != 1 || (P.fx[0].flags & REBX_B200_FLAG_BARYCENTRIC)) return false;
    if (P.fx[0].kind == REBX_B200_RADIATION) {
        const int grid = grid_for(reinterpret_cast<const void*>(tol::step_tol_kernel<REBX_B200_RADIATION>), P.st.n, 256, 0, 256);
        tol::step_tol_kernel<REBX_B200_RADIATION><<<grid, 256, 0, s>>>(P, dt, 0.5*dt, G, soft2);
        return true;
    }
    if (P.fx[0].kind == REBX_B200_YARKOVSKY) {
        const int grid = grid_for(reinterpret_cast<const void*>(tol::step_tol_kernel<REBX_B200_YARKOVSKY>), P.st.n, 256, 0, 256);
        tol::step_tol_kernel<REBX_B200_YARKOVSKY><<<grid, 256, 0, s>>>(P, dt, 0.5*dt, G, soft2);
        return true;
    }
    return false;
}

// The kernel of the fused Jacobi pipeline in tolerance arithmetic for the stack of `P`, or nullptr.
const void* tol_jacobi_fused_kernel(const rebx_b200_pass& P) {
    for (const tol::JacobiTolCombo& c : tol::kJacobiTolCombos) {
        bool ok = true;
        for (int j = 0; j < 3; j++) if (c.k[j] != (j < P.n_effects ? P.fx[j].kind : 0)) ok = false;
        if (ok) return reinterpret_cast<const void*>(c.fused);
    }
    return nullptr;
}

}  // namespace rebxb
