/*
 * oracle_forces.c -- CPU restatement of the REBOUNDx force path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's library; the product (reboundx_b200/) never does.
 *
 * Each function restates one reference loop on plain SoA arrays, sequentially and in the
 * reference's evaluation order, so that it reproduces the reference
 * bit for bit where only + - * / sqrt are involved.  Pinning: tests/test_oracle.py checks
 * every function here against oracle/_ref/libreboundx_ref.so (the reference's own sources
 * compiled in place) and against the committed fixtures in tests/golden/ that were generated
 * from it.  The orbit-element routine restates REBOUND's reb_orbit_from_particle_err
 * (third-party; REBOUND >= 5.0.0 is not under /root/reference): parity is UNPINNED at that
 * boundary for the effects that use it (yarkovsky full version, gas_damping_timescale,
 * type_I_migration, modify_orbits_forces with the planet trap).
 *
 * Presence of an optional per-particle parameter is a byte mask (NULL = set everywhere).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { double x, y, z, vx, vy, vz, m; } body_t;
typedef struct { double a, e, inc, P; } orbit_t;

static int has(const unsigned char* mask, int64_t i) { return mask == NULL || mask[i]; }

/* REBOUND reb_orbit_from_particle_err restated (see header note). */
static orbit_t orbit_of(double G, double px, double py, double pz, double pvx, double pvy, double pvz, double pm, const body_t* b) {
    orbit_t o = {NAN, NAN, NAN, NAN};
    if (b->m <= 1.E-308) return o;
    const double mu = G*(pm + b->m);
    const double dx = px - b->x, dy = py - b->y, dz = pz - b->z;
    const double dvx = pvx - b->vx, dvy = pvy - b->vy, dvz = pvz - b->vz;
    const double d = sqrt(dx*dx + dy*dy + dz*dz);
    const double vsquared = dvx*dvx + dvy*dvy + dvz*dvz;
    const double vcircsquared = mu/d;
    const double a = -mu/(vsquared - 2.*vcircsquared);
    const double hx = (dy*dvz - dz*dvy), hy = (dz*dvx - dx*dvz), hz = (dx*dvy - dy*dvx);
    const double h = sqrt(hx*hx + hy*hy + hz*hz);
    const double vdiffsquared = vsquared - vcircsquared;
    if (d <= 1.E-308) return o;
    const double vr = (dx*dvx + dy*dvy + dz*dvz)/d;
    const double rvr = d*vr;
    const double muinv = 1./mu;
    const double ex = muinv*(vdiffsquared*dx - rvr*dvx);
    const double ey = muinv*(vdiffsquared*dy - rvr*dvy);
    const double ez = muinv*(vdiffsquared*dz - rvr*dvz);
    o.a = a;
    o.e = sqrt(ex*ex + ey*ey + ez*ez);
    const double n = a/fabs(a)*sqrt(fabs(mu/(a*a*a)));
    o.P = 2*M_PI/n;
    const double cosine = hz/h;
    if (cosine > -1. && cosine < 1.) o.inc = acos(cosine);
    else o.inc = (cosine <= -1.) ? M_PI : 0.;
    return o;
}

/* reference src/inner_disk_edge.c:61-77 */
double oracle_planet_trap(double r, double dedge, double hedge) {
    if (r > dedge*(1.0 + hedge)) return 1.0;
    else if (dedge*(1.0 - hedge) < r) return 5.5*cos(((dedge*(1.0 + hedge) - r)*2*M_PI)/(4*hedge*dedge)) - 4.5;
    return -10.0;
}

/* reference src/radiation_forces.c:69-118.  sources == NULL or nsrc == 0 -> index 0. */
void oracle_radiation_forces(int64_t n, const double* x, const double* y, const double* z,
                             const double* vx, const double* vy, const double* vz, const double* m,
                             const double* beta, const unsigned char* has_beta,
                             double G, double c, const int64_t* sources, int64_t nsrc,
                             double* ax, double* ay, double* az) {
    const int64_t zero = 0;
    if (sources == NULL || nsrc == 0) { sources = &zero; nsrc = 1; }
    for (int64_t s = 0; s < nsrc; s++) {
        const int64_t si = sources[s];
        const double sx = x[si], sy = y[si], sz = z[si], svx = vx[si], svy = vy[si], svz = vz[si];
        const double mu = G*m[si];
        for (int64_t i = 0; i < n; i++) {
            if (i == si || !has(has_beta, i)) continue;
            const double dx = x[i] - sx, dy = y[i] - sy, dz = z[i] - sz;
            const double dr = sqrt(dx*dx + dy*dy + dz*dz);
            const double dvx = vx[i] - svx, dvy = vy[i] - svy, dvz = vz[i] - svz;
            const double rdot = (dx*dvx + dy*dvy + dz*dvz)/dr;
            const double a_rad = beta[i]*mu/(dr*dr);
            ax[i] += a_rad*((1. - rdot/c)*dx/dr - dvx/c);
            ay[i] += a_rad*((1. - rdot/c)*dy/dr - dvy/c);
            az[i] += a_rad*((1. - rdot/c)*dz/dr - dvz/c);
        }
    }
}

/* reference src/gravitational_harmonics.c:136-226 for ONE oblate body `bi`
 * (J2 != 0 and R_eq set are the caller's responsibility; J4 == 0 means unset).
 * br[6] (may be NULL) receives the back-reaction sum accumulated in long double (br[0..2])
 * and the sum of the terms' absolute values (br[3..5], the scale a reduction's rounding
 * error is measured against). */
void oracle_gravitational_harmonics(int64_t n, const double* x, const double* y, const double* z, const double* m,
                                    int64_t bi, double J2, double J4, double R_eq, const double Omega[3], double G,
                                    double* ax, double* ay, double* az, long double* br) {
    const double omega = sqrt(Omega[0]*Omega[0] + Omega[1]*Omega[1] + Omega[2]*Omega[2]);
    const double sx = Omega[0]/omega, sy = Omega[1]/omega, sz = Omega[2]/omega;
    double ux, uy, uz;
    const double fac0 = sqrt(sx*sx + sy*sy);
    if (fac0 != 0.0) { ux = -sy/fac0; uy = sx/fac0; uz = 0.0; } else { ux = 1.0; uy = 0.0; uz = 0.0; }
    const double wx = sx, wy = sy, wz = sz;
    const double vx_ = -(uy*wz - uz*wy), vy_ = -(uz*wx - ux*wz), vz_ = -(ux*wy - uy*wx);
    const double bx = x[bi], by = y[bi], bz = z[bi], bm = m[bi];
    long double bsx = 0, bsy = 0, bsz = 0, bax = 0, bay = 0, baz = 0;
    for (int64_t j = 0; j < n; j++) {
        if (j == bi) continue;
        const double dx = x[j] - bx, dy = y[j] - by, dz = z[j] - bz;
        const double r2 = dx*dx + dy*dy + dz*dz;
        const double r = sqrt(r2);
        const double du = ux*dx + uy*dy + uz*dz;
        const double dv = vx_*dx + vy_*dy + vz_*dz;
        const double dw = wx*dx + wy*dy + wz*dz;
        const double costheta = dw/r;
        const double costheta2 = costheta*costheta;
        double au = 0.0, av = 0.0, aw = 0.0;
        if (J2 != 0.0) {
            const double f1 = 3.0/2.0*G*bm*J2*R_eq*R_eq/r2/r2/r;
            const double f2 = 5.0*costheta2 - 1.0;
            const double f3 = f2 - 2.0;
            au += f1*f2*du; av += f1*f2*dv; aw += f1*f3*dw;
        }
        if (J4 != 0.0) {
            const double f1 = 5.0/8.0*G*bm*J4*R_eq*R_eq*R_eq*R_eq/r2/r2/r2/r;
            const double f2 = 63.0*costheta2*costheta2 - 42.0*costheta2 + 3.0;
            const double f3 = f2 - 28.0*costheta2 + 12.0;
            au += f1*f2*du; av += f1*f2*dv; aw += f1*f3*dw;
        }
        const double fx = ux*au + vx_*av + wx*aw;
        const double fy = uy*au + vy_*av + wy*aw;
        const double fz = uz*au + vz_*av + wz*aw;
        ax[j] += fx; ay[j] += fy; az[j] += fz;
        const double fac = m[j]/bm;
        ax[bi] -= fac*fx; ay[bi] -= fac*fy; az[bi] -= fac*fz;
        bsx -= (long double)(fac*fx); bsy -= (long double)(fac*fy); bsz -= (long double)(fac*fz);
        bax += fabsl((long double)(fac*fx)); bay += fabsl((long double)(fac*fy)); baz += fabsl((long double)(fac*fz));
    }
    if (br) { br[0] = bsx; br[1] = bsy; br[2] = bsz; br[3] = bax; br[4] = bay; br[5] = baz; }
}

/* reference src/gr_potential.c:60-89 */
void oracle_gr_potential(int64_t n, const double* x, const double* y, const double* z, const double* m,
                         double G, double c, double* ax, double* ay, double* az, long double* br) {
    const double C2 = c*c;
    const double sx = x[0], sy = y[0], sz = z[0], sm = m[0];
    const double prefac1 = 6.*(G*sm)*(G*sm)/C2;
    long double bsx = 0, bsy = 0, bsz = 0, bax = 0, bay = 0, baz = 0;
    for (int64_t i = 1; i < n; i++) {
        const double dx = x[i] - sx, dy = y[i] - sy, dz = z[i] - sz;
        const double r2 = dx*dx + dy*dy + dz*dz;
        const double prefac = prefac1/(r2*r2);
        ax[i] -= prefac*dx; ay[i] -= prefac*dy; az[i] -= prefac*dz;
        const double tx = m[i]/sm*prefac*dx, ty = m[i]/sm*prefac*dy, tz = m[i]/sm*prefac*dz;
        ax[0] += tx; ay[0] += ty; az[0] += tz;
        bsx += (long double)tx; bsy += (long double)ty; bsz += (long double)tz;
        bax += fabsl((long double)tx); bay += fabsl((long double)ty); baz += fabsl((long double)tz);
    }
    if (br) { br[0] = bsx; br[1] = bsy; br[2] = bsz; br[3] = bax; br[4] = bay; br[5] = baz; }
}

/* reference src/yarkovsky_effect.c:76-248.  flag values outside {-1,0,1} and particles whose
 * full-version parameter set is incomplete are skipped (the reference's behaviour for the
 * former is undefined, for the latter an error + skip).  par[9][n]: density, rotation_period,
 * thermal_inertia, albedo, emissivity, k, spin_x, spin_y, spin_z; present[9][n] byte masks. */
void oracle_yarkovsky_effect(int64_t n, const double* x, const double* y, const double* z,
                             const double* vx, const double* vy, const double* vz, const double* m, const double* rad,
                             const double* const par[9], const unsigned char* const present[9],
                             const int* flag, const unsigned char* has_flag,
                             double G, double lstar, double c, double stef_boltz, int has_stef_boltz,
                             double* ax, double* ay, double* az) {
    const body_t star = {x[0], y[0], z[0], vx[0], vy[0], vz[0], m[0]};
    for (int64_t i = 1; i < n; i++) {
        if (!has(present[0], i) || rad[i] == 0 || !has(present[3], i) || !has(has_flag, i)) continue;
        const int f = flag[i];
        if (f != 0 && f != 1 && f != -1) continue;
        const double density = par[0][i], albedo = par[3][i];
        const double radius = rad[i];
        const double q_yar = 1.0 - albedo;
        const double dx = x[i] - star.x, dy = y[i] - star.y, dz = z[i] - star.z;
        const double dvx = vx[i] - star.vx, dvy = vy[i] - star.vy, dvz = vz[i] - star.vz;
        const double distance = sqrt((dx*dx) + (dy*dy) + (dz*dz));
        const double rdotv = ((dx*dvx) + (dy*dvy) + (dz*dvz))/(c*distance);
        double iv[3];
        iv[0] = ((1 - rdotv)*(dx/distance)) - (dvx/c);
        iv[1] = ((1 - rdotv)*(dy/distance)) - (dvy/c);
        iv[2] = ((1 - rdotv)*(dz/distance)) - (dvz/c);
        double mag;
        double Y[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
        if (f == 1)  { Y[1][0] = 1.0; mag = (3*q_yar*lstar)/(64*M_PI*radius*density*c*distance*distance); }
        else if (f == -1) { Y[0][1] = 1.0; mag = (3*q_yar*lstar)/(64*M_PI*radius*density*c*distance*distance); }
        else {
            int complete = has_stef_boltz;
            const int need[7] = {1, 2, 4, 5, 6, 7, 8};
            for (int q = 0; q < 7; q++) if (!has(present[need[q]], i)) complete = 0;
            if (!complete) continue;
            const double rotation_period = par[1][i], Gamma = par[2][i], emissivity = par[4][i], k = par[5][i];
            const double sx = par[6][i], sy = par[7][i], sz = par[8][i];
            const orbit_t o = orbit_of(G, x[i], y[i], z[i], vx[i], vy[i], vz[i], m[i], &star);
            mag = (3*k*q_yar*lstar)/(16*M_PI*radius*density*c*distance*distance);
            const double Smag = sqrt((sx*sx) + sy*sy + sz*sz);
            const double hx = (dy*dvz) - (dz*dvy), hy = (dz*dvx) - (dx*dvz), hz = (dx*dvy) - (dy*dvx);
            const double Hmag = sqrt((hx*hx) + (hy*hy) + (hz*hz));
            const double is = 1.0/Smag, is2 = 1.0/(Smag*Smag), ih = 1.0/Hmag, ih2 = 1.0/(Hmag*Hmag);
            const double R1s[3][3] = {{0.0, -sz*is, sy*is}, {sz*is, 0.0, -sx*is}, {-sy*is, sx*is, 0.0}};
            const double R2s[3][3] = {{sx*sx*is2, sx*sy*is2, sx*sz*is2}, {sx*sy*is2, sy*sy*is2, sy*sz*is2}, {sx*sz*is2, sy*sz*is2, sz*sz*is2}};
            const double R1h[3][3] = {{0.0, -hz*ih, hy*ih}, {hz*ih, 0.0, -hx*ih}, {-hy*ih, hx*ih, 0.0}};
            const double R2h[3][3] = {{hx*hx*ih2, hx*hy*ih2, hx*hz*ih2}, {hx*hy*ih2, hy*hy*ih2, hy*hz*ih2}, {hx*hz*ih2, hy*hz*ih2, hz*hz*ih2}};
            const double tanPhi = 1.0/(1.0 + (.5*pow((stef_boltz*emissivity)/(M_PI*M_PI*M_PI*M_PI*M_PI), .25))*sqrt(rotation_period/(Gamma*Gamma))*pow((lstar*q_yar)/(distance*distance), .75));
            const double tanEps = 1.0/(1.0 + (.5*pow((stef_boltz*emissivity)/(M_PI*M_PI*M_PI*M_PI*M_PI), .25))*sqrt(o.P/(Gamma*Gamma))*pow((lstar*q_yar)/(distance*distance), .75));
            const double Phi = atan(tanPhi), Eps = atan(tanEps);
            const double cp = cos(Phi), sp = sin(Phi), ce = cos(Eps), se = sin(Eps);
            double Rys[3][3], Ryh[3][3];
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
                const double unit = (a == b) ? 1.0 : 0.0;
                Rys[a][b] = (cp*unit) + (sp*R1s[a][b]) + ((1.0 - cp)*R2s[a][b]);
                Ryh[a][b] = (ce*unit) - (se*R1h[a][b]) + ((1 - ce)*R2h[a][b]);
            }
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++)
                Y[a][b] = (Rys[a][0]*Ryh[0][b]) + (Rys[a][1]*Ryh[1][b]) + (Rys[a][2]*Ryh[2][b]);
        }
        double dir[3];
        for (int a = 0; a < 3; a++) dir[a] = (Y[a][0]*iv[0]) + (Y[a][1]*iv[1]) + (Y[a][2]*iv[2]);
        ax[i] += mag*dir[0]; ay[i] += mag*dir[1]; az[i] += mag*dir[2];
    }
}

/* reference src/central_force.c:63-96 for ONE central body `bi` with (Acentral, gammacentral) */
void oracle_central_force(int64_t n, const double* x, const double* y, const double* z, const double* m,
                          int64_t bi, double A, double gamma, double* ax, double* ay, double* az, long double* br) {
    const double sx = x[bi], sy = y[bi], sz = z[bi], sm = m[bi];
    long double bsx = 0, bsy = 0, bsz = 0, bax = 0, bay = 0, baz = 0;
    for (int64_t i = 0; i < n; i++) {
        if (i == bi) continue;
        const double dx = x[i] - sx, dy = y[i] - sy, dz = z[i] - sz;
        const double r2 = dx*dx + dy*dy + dz*dz;
        const double prefac = A*pow(r2, (gamma - 1.)/2.);
        ax[i] += prefac*dx; ay[i] += prefac*dy; az[i] += prefac*dz;
        const double tx = m[i]/sm*prefac*dx, ty = m[i]/sm*prefac*dy, tz = m[i]/sm*prefac*dz;
        ax[bi] -= tx; ay[bi] -= ty; az[bi] -= tz;
        bsx -= (long double)tx; bsy -= (long double)ty; bsz -= (long double)tz;
        bax += fabsl((long double)tx); bay += fabsl((long double)ty); baz += fabsl((long double)tz);
    }
    if (br) { br[0] = bsx; br[1] = bsy; br[2] = bsz; br[3] = bax; br[4] = bay; br[5] = baz; }
}

/* reference src/lense_thirring.c:65-100 (source = particle 0 with spin angular momentum J = I*Omega) */
void oracle_lense_thirring(int64_t n, const double* x, const double* y, const double* z,
                           const double* vx, const double* vy, const double* vz, const double* m,
                           double G, double c, double I, const double Omega[3], double* ax, double* ay, double* az, long double* br) {
    const double C2 = c*c;
    const double gamma = 1.000021;
    const double sx = x[0], sy = y[0], sz = z[0], svx = vx[0], svy = vy[0], svz = vz[0], ms = m[0];
    long double bs[3] = {0, 0, 0}, ba[3] = {0, 0, 0};
    for (int64_t i = 1; i < n; i++) {
        const double dx = x[i] - sx, dy = y[i] - sy, dz = z[i] - sz;
        const double r2 = dx*dx + dy*dy + dz*dz;
        const double r = sqrt(r2);
        const double r3 = r2*r;
        const double dvx = vx[i] - svx, dvy = vy[i] - svy, dvz = vz[i] - svz;
        const double Jx = I*Omega[0], Jy = I*Omega[1], Jz = I*Omega[2];
        const double mt = m[i];
        const double mtot = ms + mt;
        const double mratio = mt/ms;
        const double Omega_fac = ms/mtot*(1. + gamma)*G/2/C2/r3;
        const double Jdotr = Jx*dx + Jy*dy + Jz*dz;
        const double Ox = Omega_fac*(-Jx + 3.*Jdotr*dx/r2);
        const double Oy = Omega_fac*(-Jy + 3.*Jdotr*dy/r2);
        const double Oz = Omega_fac*(-Jz + 3.*Jdotr*dz/r2);
        ax[i] += 2.*(Oy*dvz - Oz*dvy); ay[i] += 2.*(Oz*dvx - Ox*dvz); az[i] += 2.*(Ox*dvy - Oy*dvx);
        const double t[3] = {mratio*2.*(Oy*dvz - Oz*dvy), mratio*2.*(Oz*dvx - Ox*dvz), mratio*2.*(Ox*dvy - Oy*dvx)};
        ax[0] -= t[0]; ay[0] -= t[1]; az[0] -= t[2];
        for (int q = 0; q < 3; q++) { bs[q] -= (long double)t[q]; ba[q] += fabsl((long double)t[q]); }
    }
    if (br) { for (int q = 0; q < 3; q++) { br[q] = bs[q]; br[3 + q] = ba[q]; } }
}

/* reference src/gas_dynamical_friction.c:66-137 (gaseous disc around particle 0; par = rhog, alpha_rhog,
 * cs, alpha_cs, xmin, hr, Qd) */
void oracle_gas_dynamical_friction(int64_t n, const double* x, const double* y, const double* z,
                                   const double* vx, const double* vy, const double* vz, const double* m, const double* rad,
                                   double G, const double par[7], double* ax, double* ay, double* az) {
    const double rhog = par[0], alpha_rhog = par[1], cs = par[2], alpha_cs = par[3], xmin = par[4], hr = par[5], Qd = par[6];
    for (int64_t i = 1; i < n; i++) {
        const double dx = x[i] - x[0], dy = y[i] - y[0], dz = z[i] - z[0];
        const double dvx = vx[i] - vx[0], dvy = vy[i] - vy[0], dvz = vz[i] - vz[0];
        const double rcyl = sqrt(dx*dx + dy*dy);
        const double sin_phi = dy/rcyl, cos_phi = dx/rcyl;
        const double vk = sqrt(G*m[0]/rcyl)*(1.0 - hr*hr);
        const double v0 = dvx + vk*sin_phi, v1 = dvy - vk*cos_phi, v2 = dvz;
        const double vrel_norm = sqrt(v0*v0 + v1*v1 + v2*v2);
        const double mach = vrel_norm/(cs*pow(rcyl, alpha_cs));
        const double coul = log(1.0/xmin);
        double integ;
        if (mach >= 1.0) integ = coul;
        else {
            const double sub = (mach < 0.02) ? mach*mach*mach/3. + mach*mach*mach*mach*mach/5. : 0.5*log((1.0 + mach)/(1.0 - mach)) - mach;
            integ = fmin(coul, sub);
        }
        const double h = hr*rcyl;
        const double vert = (fabs(dz) < (10*h)) ? exp(-dz*dz/(2.0*h*h)) : 0;
        const double rhog_loc = rhog*pow(rcyl, alpha_rhog)*vert;
        const double mp = m[i], rstar = rad[i];
        const double fc = 4.*M_PI*(G*G)*mp*(rhog_loc)/(vrel_norm*vrel_norm*vrel_norm)*integ + M_PI*rhog_loc*rstar*rstar*vrel_norm*Qd/mp;
        ax[i] -= fc*v0; ay[i] -= fc*v1; az[i] -= fc*v2;
    }
}

/* ---- damping trio under rebx_com_force (reference src/rebxtools.c:66-147) ------------------ */
enum { ORACLE_MODIFY_ORBITS = 5, ORACLE_GAS_DAMPING = 6, ORACLE_TYPE_I = 7, ORACLE_EXP_MIGRATION = 9 };
enum { ORACLE_JACOBI = 0, ORACLE_BARYCENTRIC = 1, ORACLE_PARTICLE = 2 };

typedef struct {
    /* modify_orbits_forces */
    const double *tau_a, *tau_e, *tau_inc;
    const unsigned char *has_tau_a, *has_tau_e, *has_tau_inc;
    int trap; double dedge, hedge;
    /* gas_damping_timescale */
    const double* d_factor; const unsigned char* has_d_factor; double cs_coeff, tau_coeff;
    /* type_I_migration */
    double tIm_beta, tIm_h0, tIm_sd0, tIm_s, tIm_dedge, tIm_hedge;
    /* exponential_migration */
    const double *em_tau_a, *em_aini, *em_afin;
    const unsigned char *has_em_tau_a, *has_em_aini, *has_em_afin;
    double t;
} oracle_damping_params;

static void damping_calc(int kind, const oracle_damping_params* P, double G, int64_t i,
                         double px, double py, double pz, double pvx, double pvy, double pvz, double pm,
                         const body_t* src, double f[3]) {
    f[0] = f[1] = f[2] = 0.0;
    const double dvx = pvx - src->vx, dvy = pvy - src->vy, dvz = pvz - src->vz;
    const double dx = px - src->x, dy = py - src->y, dz = pz - src->z;
    const double r2 = dx*dx + dy*dy + dz*dz;
    if (kind == ORACLE_MODIFY_ORBITS) {            /* src/modify_orbits_forces.c:77-128 */
        double invtau_a = 0.0, tau_e = INFINITY, tau_inc = INFINITY;
        if (has(P->has_tau_a, i) && P->tau_a) {
            invtau_a = 1.0/P->tau_a[i];
            if (P->trap) {
                const orbit_t o = orbit_of(G, px, py, pz, pvx, pvy, pvz, pm, src);
                invtau_a *= oracle_planet_trap(o.a, P->dedge, P->hedge);
            }
        }
        if (has(P->has_tau_e, i) && P->tau_e) tau_e = P->tau_e[i];
        if (has(P->has_tau_inc, i) && P->tau_inc) tau_inc = P->tau_inc[i];
        f[0] = dvx*invtau_a/(2.); f[1] = dvy*invtau_a/(2.); f[2] = dvz*invtau_a/(2.);
        if (tau_e < INFINITY || tau_inc < INFINITY) {
            const double vdotr = dx*dvx + dy*dvy + dz*dvz;
            const double prefac = 2*vdotr/r2/tau_e;
            f[0] += prefac*dx; f[1] += prefac*dy; f[2] += prefac*dz + 2.*dvz/tau_inc;
        }
    } else if (kind == ORACLE_GAS_DAMPING) {       /* src/gas_damping_timescale.c:67-130 */
        if (!P->d_factor || !has(P->has_d_factor, i)) return;
        const orbit_t o = orbit_of(G, px, py, pz, pvx, pvy, pvz, pm, src);
        const double a0 = o.a, e0 = o.e, inc0 = o.inc;
        const double starMass = src->m, planetMass = pm;
        double coeff;
        const double vk = sqrt(G*starMass/a0);
        const double v = sqrt(e0*e0 + inc0*inc0)*vk;
        const double cs = P->cs_coeff/sqrt(sqrt(a0));
        const double v_over_cs = v/cs;
        if (v <= cs) coeff = 1.;
        else if (inc0 < cs/vk) coeff = v_over_cs*v_over_cs*v_over_cs;
        else coeff = v_over_cs*v_over_cs*v_over_cs*v_over_cs;
        const double tau_e = -P->tau_coeff*P->d_factor[i]*a0*a0*(starMass/planetMass)*coeff;
        const double tau_inc = 2.*tau_e;
        if (tau_e < INFINITY || tau_inc < INFINITY) {
            const double vdotr = dx*dvx + dy*dvy + dz*dvz;
            const double prefac = 2*vdotr/r2/tau_e;
            f[0] += prefac*dx; f[1] += prefac*dy; f[2] += prefac*dz + 2.*dvz/tau_inc;
        }
    } else if (kind == ORACLE_EXP_MIGRATION) {     /* src/exponential_migration.c:61-100 */
        const orbit_t o = orbit_of(G, px, py, pz, pvx, pvy, pvz, pm, src);
        double em_tau_a = INFINITY, em_aini = 24., em_afin = 30.;
        if (P->em_tau_a && has(P->has_em_tau_a, i)) em_tau_a = P->em_tau_a[i];
        if (P->em_aini && has(P->has_em_aini, i)) em_aini = P->em_aini[i];
        if (P->em_afin && has(P->has_em_afin, i)) em_afin = P->em_afin[i];
        f[0] = (dvx/(2.*em_tau_a))*((em_afin - em_aini)/(o.a))*exp(-(P->t)/em_tau_a);
        f[1] = (dvy/(2.*em_tau_a))*((em_afin - em_aini)/(o.a))*exp(-(P->t)/em_tau_a);
        f[2] = (dvz/(2.*em_tau_a))*((em_afin - em_aini)/(o.a))*exp(-(P->t)/em_tau_a);
    } else {                                       /* src/type_I_migration.c:69-203 */
        const orbit_t o = orbit_of(G, px, py, pz, pvx, pvy, pvz, pm, src);
        const double a0 = o.a, e0 = o.e, inc0 = o.inc, mp = pm, ms = src->m;
        const double h = P->tIm_h0*pow(r2, P->tIm_beta/2);
        const double h2 = h*h;
        const double eh = e0/h, ih = inc0/h;
        const double sd = P->tIm_sd0*pow(sqrt(r2), -P->tIm_s);
        const double wave = (sqrt(ms*ms*ms)*h2*h2)/(mp*sd*sqrt(a0*G));
        const double term = (eh/2.25), term2 = (eh/2.84)*(eh/2.84), term3 = (eh/2.02)*(eh/2.02);
        const double Pe = (1. + pow(term, 1.2) + term2*term2*term2)/(1. - term3*term3);
        const double t_mig = ((2.*wave)/(2.7 + 1.1*P->tIm_s))*(1/h2)*(Pe + (Pe/fabs(Pe))*((0.070*ih) + (0.085*ih*ih*ih*ih) - (0.080*eh*ih*ih)));
        const double invtau_mig = oracle_planet_trap(a0, P->tIm_dedge, P->tIm_hedge)/t_mig;
        const double tau_e = (wave/0.780)*(1. - (0.14*eh*eh) + (0.06*eh*eh*eh) + (0.18*eh*ih*ih));
        const double tau_inc = (wave/0.544)*(1 - (0.30*ih*ih) + (0.24*ih*ih*ih) + (0.14*eh*eh*ih));
        if (invtau_mig != 0.0) { f[0] = -dvx*invtau_mig; f[1] = -dvy*invtau_mig; f[2] = -dvz*invtau_mig; }
        if (tau_e < INFINITY || tau_inc < INFINITY) {
            const double vdotr = dx*dvx + dy*dvy + dz*dvz;
            const double prefac = -2*vdotr/r2/tau_e;
            f[0] += prefac*dx; f[1] += prefac*dy; f[2] += prefac*dz - 2*dvz/tau_inc;
        }
    }
}

/* COM of all particles, folded pairwise in index order (REBOUND reb_simulation_com restated). */
static body_t com_all(int64_t n, const double* x, const double* y, const double* z,
                      const double* vx, const double* vy, const double* vz, const double* m) {
    body_t c = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < n; i++) {
        c.x = c.x*c.m + x[i]*m[i]; c.y = c.y*c.m + y[i]*m[i]; c.z = c.z*c.m + z[i]*m[i];
        c.vx = c.vx*c.m + vx[i]*m[i]; c.vy = c.vy*c.m + vy[i]*m[i]; c.vz = c.vz*c.m + vz[i]*m[i];
        c.m += m[i];
        if (c.m > 0.) { c.x /= c.m; c.y /= c.m; c.z /= c.m; c.vx /= c.m; c.vy /= c.m; c.vz /= c.m; }
    }
    return c;
}

/* Returns 0, or 1 when PARTICLE coordinates are requested and `primary` < 0 (the reference
 * reports an error there).  Sequential, i = n-1 .. 0, O(n^2) in Jacobi/barycentric mode exactly
 * like the reference.  br[6] (may be NULL): long-double sum of the terms put on the primary and
 * the sum of their absolute values (PARTICLE mode only). */
int oracle_com_force(int kind, int coordinates, int64_t primary, const oracle_damping_params* P, double G,
                     int64_t n, const double* x, const double* y, const double* z,
                     const double* vx, const double* vy, const double* vz, const double* m,
                     double* ax, double* ay, double* az, long double* br) {
    body_t com = com_all(n, x, y, z, vx, vy, vz, m);
    int64_t refindex = -1;
    if (coordinates == ORACLE_JACOBI) refindex = 0;
    else if (coordinates == ORACLE_PARTICLE) {
        if (primary < 0) return 1;
        refindex = primary;
        com.x = x[primary]; com.y = y[primary]; com.z = z[primary];
        com.vx = vx[primary]; com.vy = vy[primary]; com.vz = vz[primary]; com.m = m[primary];
    }
    long double bsx = 0, bsy = 0, bsz = 0, bax = 0, bay = 0, baz = 0;
    for (int64_t i = n - 1; i >= 0; i--) {
        if (i == refindex) continue;
        if (coordinates == ORACLE_JACOBI) {         /* rebx_get_com_without_particle, rebxtools.c:6-30 */
            com.x = com.x*com.m - x[i]*m[i]; com.y = com.y*com.m - y[i]*m[i]; com.z = com.z*com.m - z[i]*m[i];
            com.vx = com.vx*com.m - vx[i]*m[i]; com.vy = com.vy*com.m - vy[i]*m[i]; com.vz = com.vz*com.m - vz[i]*m[i];
            com.m -= m[i];
            if (com.m > 0.) { com.x /= com.m; com.y /= com.m; com.z /= com.m; com.vx /= com.m; com.vy /= com.m; com.vz /= com.m; }
        }
        double f[3];
        damping_calc(kind, P, G, i, x[i], y[i], z[i], vx[i], vy[i], vz[i], m[i], &com, f);
        ax[i] += f[0]; ay[i] += f[1]; az[i] += f[2];
        double massratio;
        if (coordinates == ORACLE_BARYCENTRIC) {
            massratio = m[i]/com.m;
            for (int64_t j = 0; j < n; j++) { ax[j] -= massratio*f[0]; ay[j] -= massratio*f[1]; az[j] -= massratio*f[2]; }
        } else if (coordinates == ORACLE_JACOBI) {
            massratio = m[i]/(com.m + m[i]);
            for (int64_t j = 0; j < i + 1; j++) { ax[j] -= massratio*f[0]; ay[j] -= massratio*f[1]; az[j] -= massratio*f[2]; }
        } else {
            massratio = m[i]/(com.m + m[i]);
            ax[i] -= massratio*f[0]; ay[i] -= massratio*f[1]; az[i] -= massratio*f[2];
            ax[refindex] -= massratio*f[0]; ay[refindex] -= massratio*f[1]; az[refindex] -= massratio*f[2];
            bsx -= (long double)(massratio*f[0]); bsy -= (long double)(massratio*f[1]); bsz -= (long double)(massratio*f[2]);
            bax += fabsl((long double)(massratio*f[0])); bay += fabsl((long double)(massratio*f[1])); baz += fabsl((long double)(massratio*f[2]));
        }
    }
    if (br) { br[0] = bsx; br[1] = bsy; br[2] = bsz; br[3] = bax; br[4] = bay; br[5] = baz; }
    return 0;
}

/* O(n) restatement of the same function for ensembles the reference's O(n^2) loops cannot reach
 * (JACOBI at n = 4*10^4 already takes 4 s per force there; ~40 min at 10^6).
 *   - The reference frame of particle i is the SAME double-precision recurrence as above (forward fold of
 *     reb_simulation_com, then the backward peel of rebx_get_com_without_particle, rebxtools.c:6-30,92-99):
 *     it is O(n) in the reference too, so COM_i is bit-identical to the reference's.
 *   - Only the back-reaction is restated: the reference subtracts massratio*f_i from every j <= i (JACOBI) or
 *     from every j (BARYCENTRIC) one i at a time, i.e. n sequential roundings per particle; here the terms
 *     are accumulated in long double (suffix sums / one total) and applied once.  The two differ by the
 *     reference's own accumulated rounding, <= sqrt(n)*eps relative to |a_j| (1.6e-14 at n = 2*10^4):
 *     tests/test_oracle.py pins that against the O(n^2) function and against oracle/_ref.
 * com_mode 1 replaces the reference's COM recurrence by exact (long double) prefix sums -- used only to
 * measure how much of a disagreement stems from the reference's own rounding in COM_i. */
int oracle_com_force_linear(int kind, int coordinates, int64_t primary, const oracle_damping_params* P, double G,
                            int64_t n, const double* x, const double* y, const double* z,
                            const double* vx, const double* vy, const double* vz, const double* m,
                            double* ax, double* ay, double* az, int com_mode) {
    if (coordinates == ORACLE_PARTICLE)
        return oracle_com_force(kind, coordinates, primary, P, G, n, x, y, z, vx, vy, vz, m, ax, ay, az, NULL);
    body_t com = com_all(n, x, y, z, vx, vy, vz, m);
    long double Sm = 0, S[6] = {0, 0, 0, 0, 0, 0};
    if (com_mode == 1) {
        for (int64_t i = 0; i < n; i++) {
            Sm += m[i];
            S[0] += (long double)x[i]*m[i]; S[1] += (long double)y[i]*m[i]; S[2] += (long double)z[i]*m[i];
            S[3] += (long double)vx[i]*m[i]; S[4] += (long double)vy[i]*m[i]; S[5] += (long double)vz[i]*m[i];
        }
        if (coordinates == ORACLE_BARYCENTRIC) {
            com.m = (double)Sm;
            com.x = (double)(S[0]/Sm); com.y = (double)(S[1]/Sm); com.z = (double)(S[2]/Sm);
            com.vx = (double)(S[3]/Sm); com.vy = (double)(S[4]/Sm); com.vz = (double)(S[5]/Sm);
        }
    }
    long double tx = 0, ty = 0, tz = 0;      /* running sum of massratio*f over the particles already visited */
    if (coordinates == ORACLE_BARYCENTRIC) {
        for (int64_t i = n - 1; i >= 0; i--) {
            double f[3];
            damping_calc(kind, P, G, i, x[i], y[i], z[i], vx[i], vy[i], vz[i], m[i], &com, f);
            const double massratio = m[i]/com.m;
            tx += (long double)(massratio*f[0]); ty += (long double)(massratio*f[1]); tz += (long double)(massratio*f[2]);
            ax[i] += f[0]; ay[i] += f[1]; az[i] += f[2];
        }
        for (int64_t j = 0; j < n; j++) {
            ax[j] = (double)((long double)ax[j] - tx); ay[j] = (double)((long double)ay[j] - ty); az[j] = (double)((long double)az[j] - tz);
        }
        return 0;
    }
    for (int64_t i = n - 1; i >= 1; i--) {
        if (com_mode == 1) {
            Sm -= m[i];
            S[0] -= (long double)x[i]*m[i]; S[1] -= (long double)y[i]*m[i]; S[2] -= (long double)z[i]*m[i];
            S[3] -= (long double)vx[i]*m[i]; S[4] -= (long double)vy[i]*m[i]; S[5] -= (long double)vz[i]*m[i];
            com.m = (double)Sm;
            com.x = (double)(S[0]/Sm); com.y = (double)(S[1]/Sm); com.z = (double)(S[2]/Sm);
            com.vx = (double)(S[3]/Sm); com.vy = (double)(S[4]/Sm); com.vz = (double)(S[5]/Sm);
        } else {
            com.x = com.x*com.m - x[i]*m[i]; com.y = com.y*com.m - y[i]*m[i]; com.z = com.z*com.m - z[i]*m[i];
            com.vx = com.vx*com.m - vx[i]*m[i]; com.vy = com.vy*com.m - vy[i]*m[i]; com.vz = com.vz*com.m - vz[i]*m[i];
            com.m -= m[i];
            if (com.m > 0.) { com.x /= com.m; com.y /= com.m; com.z /= com.m; com.vx /= com.m; com.vy /= com.m; com.vz /= com.m; }
        }
        double f[3];
        damping_calc(kind, P, G, i, x[i], y[i], z[i], vx[i], vy[i], vz[i], m[i], &com, f);
        const double massratio = m[i]/(com.m + m[i]);
        tx += (long double)(massratio*f[0]); ty += (long double)(massratio*f[1]); tz += (long double)(massratio*f[2]);
        ax[i] = (double)((long double)ax[i] + f[0] - tx); ay[i] = (double)((long double)ay[i] + f[1] - ty); az[i] = (double)((long double)az[i] + f[2] - tz);
    }
    ax[0] = (double)((long double)ax[0] - tx); ay[0] = (double)((long double)ay[0] - ty); az[0] = (double)((long double)az[0] - tz);
    return 0;
}

/* ---- gr (reference src/gr.c:65-181) ----------------------------------------------------------
 * n_active < 0 means "unset" (all particles active).  The Jacobi transforms restate REBOUND's
 * reb_transformations_inertial_to_jacobi_posvelacc / _jacobi_to_inertial_acc (third-party, same
 * restatement as reboundx_b200/rebound_shim/rebound_shim.c: parity UNPINNED at that boundary).
 * Returns 1 if some fixed-point iteration hit 10 iterations (the reference's warning). */
#include <stdlib.h>
int oracle_gr(int64_t n, int64_t n_active, const double* x, const double* y, const double* z,
              const double* vx, const double* vy, const double* vz, const double* m,
              double G, double c, int max_iterations, double* ax, double* ay, double* az) {
    const int64_t A = (n_active < 0 || n_active > n) ? n : n_active;
    const double C2 = c*c;
    double* na = calloc(3*(size_t)n, sizeof(double));          /* Newtonian accelerations */
    double* J = malloc(9*(size_t)n*sizeof(double));             /* Jacobi pos, vel, acc */
    int warned = 0;
    for (int64_t i = 0; i < A; i++) {
        for (int64_t j = i + 1; j < n; j++) {
            const double dx = x[i] - x[j], dy = y[i] - y[j], dz = z[i] - z[j];
            const double r2 = dx*dx + dy*dy + dz*dz;
            const double r = sqrt(r2);
            const double prefac = G/(r2*r);
            na[3*i] -= prefac*m[j]*dx; na[3*i+1] -= prefac*m[j]*dy; na[3*i+2] -= prefac*m[j]*dz;
            na[3*j] += prefac*m[i]*dx; na[3*j+1] += prefac*m[i]*dy; na[3*j+2] += prefac*m[i]*dz;
        }
    }
    const double mu = G*m[0];
    /* inertial -> Jacobi (pos, vel, acc) */
    double eta = m[0];
    double s[9] = {eta*x[0], eta*y[0], eta*z[0], eta*vx[0], eta*vy[0], eta*vz[0], eta*na[0], eta*na[1], eta*na[2]};
    for (int64_t i = 1; i < A; i++) {
        const double ei = 1./eta;
        const double pi[9] = {x[i], y[i], z[i], vx[i], vy[i], vz[i], na[3*i], na[3*i+1], na[3*i+2]};
        eta += m[i];
        const double pme = eta*ei;
        for (int q = 0; q < 9; q++) { J[9*i+q] = pi[q] - s[q]*ei; }
        for (int q = 0; q < 9; q++) { s[q] = s[q]*pme + m[i]*J[9*i+q]; }
    }
    const double Mtotal = eta, Mtotali = 1./Mtotal;
    for (int64_t i = A; i < n; i++) {
        const double pi[9] = {x[i], y[i], z[i], vx[i], vy[i], vz[i], na[3*i], na[3*i+1], na[3*i+2]};
        for (int q = 0; q < 9; q++) J[9*i+q] = pi[q] - s[q]*Mtotali;
    }
    for (int q = 0; q < 9; q++) J[q] = s[q]*Mtotali;
    /* 1PN term in Jacobi coordinates */
    for (int64_t i = 1; i < n; i++) {
        const double px = J[9*i], py = J[9*i+1], pz = J[9*i+2], pvx = J[9*i+3], pvy = J[9*i+4], pvz = J[9*i+5];
        const double pax = J[9*i+6], pay = J[9*i+7], paz = J[9*i+8];
        double vix = pvx, viy = pvy, viz = pvz;
        double vi2 = vix*vix + viy*viy + viz*viz;
        const double ri = sqrt(px*px + py*py + pz*pz);
        int q = 0;
        double Aa = (0.5*vi2 + 3.*mu/ri)/C2;
        for (q = 0; q < max_iterations; q++) {
            const double ox = vix, oy = viy, oz = viz;
            vix = pvx/(1. - Aa); viy = pvy/(1. - Aa); viz = pvz/(1. - Aa);
            vi2 = vix*vix + viy*viy + viz*viz;
            Aa = (0.5*vi2 + 3.*mu/ri)/C2;
            const double dvx = vix - ox, dvy = viy - oy, dvz = viz - oz;
            if ((dvx*dvx + dvy*dvy + dvz*dvz)/vi2 < 2.220446049250313e-16*2.220446049250313e-16) break;
        }
        if (q == 10) warned = 1;
        const double B = (mu/ri - 1.5*vi2)*mu/(ri*ri*ri)/C2;
        const double rdotrdot = px*pvx + py*pvy + pz*pvz;
        const double vdx = pax + B*px, vdy = pay + B*py, vdz = paz + B*pz;
        const double vdotvdot = vix*vdx + viy*vdy + viz*vdz;
        const double D = (vdotvdot - 3.*mu/(ri*ri*ri)*rdotrdot)/C2;
        J[9*i+6] = B*(1. - Aa)*px - Aa*pax - D*vix;
        J[9*i+7] = B*(1. - Aa)*py - Aa*pay - D*viy;
        J[9*i+8] = B*(1. - Aa)*pz - Aa*paz - D*viz;
    }
    J[6] = J[7] = J[8] = 0.;
    /* Jacobi -> inertial (acc) */
    eta = Mtotal;
    double sa[3] = {J[6]*eta, J[7]*eta, J[8]*eta};
    {
        const double ei = 1./eta;
        for (int64_t i = A; i < n; i++) for (int q = 0; q < 3; q++) na[3*i+q] = J[9*i+6+q] + sa[q]*ei;
    }
    for (int64_t i = A - 1; i > 0; i--) {
        const double ei = 1./eta;
        for (int q = 0; q < 3; q++) { sa[q] = (sa[q] - m[i]*J[9*i+6+q])*ei; na[3*i+q] = J[9*i+6+q] + sa[q]; }
        eta -= m[i];
        for (int q = 0; q < 3; q++) sa[q] *= eta;
    }
    { const double mi = 1./eta; for (int q = 0; q < 3; q++) na[q] = sa[q]*mi; }
    for (int64_t i = 0; i < n; i++) { ax[i] += na[3*i]; ay[i] += na[3*i+1]; az[i] += na[3*i+2]; }
    free(na); free(J);
    return warned;
}
