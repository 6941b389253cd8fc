/* TEST INFRASTRUCTURE ONLY — see gbp_oracle.h.  Plain-C restatement of the reference's extend path.
 * Every function cites the reference file:line it follows (paths under /root/reference).
 * fp64 throughout, expressions in the reference's source order, compiled with -ffp-contract=off. */
#include "gbp_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ====================================================================== Philox4x32-10 stream spec
 * (north-star item 1: counter-based stream shared by the CPU harness and the CUDA sampler.)
 * Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11), Philox4x32 with 10 rounds. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
	for (int r = 0; r < 10; ++r) {
		uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t) p1;
		uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t) p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Uniform number j of a (seed, stream, idx, purpose) cell lives in Philox block j/2:
 *   key = (seed lo, seed hi); ctr = (idx lo, idx hi, stream lo, (stream hi & 0xffffff) | block<<24 | purpose<<28)
 *   u   = ((w[2(j&1)] >> 5) * 2^26 + (w[2(j&1)+1] >> 6)) * 2^-53        in [0,1), 53 bits, exact in fp64 */
void orc_uniforms(uint64_t seed, uint64_t stream, uint64_t idx, int purpose, int first, int n, double *u) {
	uint32_t key[2] = {(uint32_t) seed, (uint32_t) (seed >> 32)};
	int have = -1;
	uint32_t w[4];
	for (int j = first; j < first + n; ++j) {
		int block = j >> 1;
		if (block != have) {
			uint32_t ctr[4] = {(uint32_t) idx, (uint32_t) (idx >> 32), (uint32_t) stream,
							   ((uint32_t) (stream >> 32) & 0x00ffffffu) | ((uint32_t) block << 24) | ((uint32_t) purpose << 28)};
			orc_philox4x32_10(ctr, key, w);
			have = block;
		}
		uint32_t hi = w[2 * (j & 1)], lo = w[2 * (j & 1) + 1];
		u[j - first] = ((double) (hi >> 5) * 67108864.0 + (double) (lo >> 6)) * (1.0 / 9007199254740992.0);
	}
}

/* Deterministic elementary functions: only + - * / and bit moves, so a CUDA build with -fmad=false
 * reproduces them bit for bit (libm's log/sin/cos differ between glibc and CUDA).  Accuracy ~1e-16,
 * reproducibility is what the spec needs.  Used ONLY by the samplers, never by validity checks. */
double orc_det_log(double x) { /* x normal, > 0 */
	uint64_t b;
	memcpy(&b, &x, 8);
	int e = (int) ((b >> 52) & 0x7ff) - 1023;
	b = (b & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
	double m;
	memcpy(&m, &b, 8);
	if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
	double f = m - 1.0, s = f / (2.0 + f), z = s * s;
	double p = 1.0 / 23.0;
	p = p * z + 1.0 / 21.0; p = p * z + 1.0 / 19.0; p = p * z + 1.0 / 17.0; p = p * z + 1.0 / 15.0;
	p = p * z + 1.0 / 13.0; p = p * z + 1.0 / 11.0; p = p * z + 1.0 / 9.0;  p = p * z + 1.0 / 7.0;
	p = p * z + 1.0 / 5.0;  p = p * z + 1.0 / 3.0;  p = p * z + 1.0;
	double lm = 2.0 * s * p;
	return (double) e * 6.93147180369123816490e-01 + ((double) e * 1.90821492927058770002e-10 + lm);
}
void orc_det_sincos(double x, double *sn, double *cs) { /* |x| < ~1e4 */
	double qf = floor(x * 0.63661977236758134308 + 0.5);
	int q = (int) qf;
	double r = (x - qf * 1.57079632673412561417e+00) - qf * 6.07710050650619224932e-11;
	double z = r * r;
	double ps = -1.0 / 355687428096000.0; /* -1/17! */
	ps = ps * z + 1.0 / 1307674368000.0; ps = ps * z - 1.0 / 6227020800.0; ps = ps * z + 1.0 / 39916800.0;
	ps = ps * z - 1.0 / 362880.0; ps = ps * z + 1.0 / 5040.0; ps = ps * z - 1.0 / 120.0; ps = ps * z + 1.0 / 6.0;
	double s = r - r * z * ps;
	double pc = 1.0 / 20922789888000.0; /* 1/16! */
	pc = pc * z - 1.0 / 87178291200.0; pc = pc * z + 1.0 / 479001600.0; pc = pc * z - 1.0 / 3628800.0;
	pc = pc * z + 1.0 / 40320.0; pc = pc * z - 1.0 / 720.0; pc = pc * z + 1.0 / 24.0; pc = pc * z - 0.5;
	double c = 1.0 + z * pc;
	switch (q & 3) {
	case 0: *sn = s; *cs = c; break;
	case 1: *sn = c; *cs = -s; break;
	case 2: *sn = -s; *cs = -c; break;
	default: *sn = -c; *cs = s; break;
	}
}

/* ====================================================================== terrain
 * Cell search of fast_terrain_map.cpp:101-117: first i with x_data[i] <= x < x_data[i+1].  For
 * strictly increasing axes that i is unique, found here by bisection.  DEFINED semantics where the
 * reference has undefined behaviour: the scan is restricted to i <= n-2 (the
 * reference's last iteration reads x_data[n] past the end), and when no cell matches the cell
 * (0)-anchored extrapolation x1 = x_data[0], x2 = x_data[1] is used (the reference leaves x1/x2
 * uninitialised with ix = 0) and ORC_FLAG_OOG is raised. */
static int find_cell(const double *ax, int n, double v, unsigned *flags) {
	if (!(v >= ax[0]) || !(v < ax[n - 1])) { /* also catches NaN */
		if (flags) *flags |= ORC_FLAG_OOG;
		return 0;
	}
	int lo = 0, hi = n - 1; /* ax[lo] <= v < ax[hi] */
	while (hi - lo > 1) {
		int mid = (lo + hi) >> 1;
		if (ax[mid] <= v) lo = mid; else hi = mid;
	}
	return lo;
}

/* bilinear form of fast_terrain_map.cpp:120-126, left-to-right as written */
static double bilinear(const double *layer, int ny, int ix, int iy, double x1, double x2, double y1, double y2,
					   double x, double y) {
	double f11 = layer[(size_t) ix * ny + iy], f12 = layer[(size_t) ix * ny + iy + 1];
	double f21 = layer[(size_t) (ix + 1) * ny + iy], f22 = layer[(size_t) (ix + 1) * ny + iy + 1];
	return 1.0 / ((x2 - x1) * (y2 - y1)) *
		   (f11 * (x2 - x) * (y2 - y) + f21 * (x - x1) * (y2 - y) + f12 * (x2 - x) * (y - y1) + f22 * (x - x1) * (y - y1));
}

double orc_ground_height(const orc_terrain *t, double x, double y, unsigned *flags) { /* fast_terrain_map.cpp:94-132 */
	int ix = find_cell(t->x, t->nx, x, flags), iy = find_cell(t->y, t->ny, y, flags);
	return bilinear(t->z, t->ny, ix, iy, t->x[ix], t->x[ix + 1], t->y[iy], t->y[iy + 1], x, y);
}
int orc_height_is_nan(const orc_terrain *t, double x, double y, unsigned *flags) { /* fast_terrain_map.cpp:135-157 */
	int ix = find_cell(t->x, t->nx, x, flags), iy = find_cell(t->y, t->ny, y, flags);
	const double *z = t->z;
	int ny = t->ny;
	return isnan(z[(size_t) ix * ny + iy]) || isnan(z[(size_t) ix * ny + iy + 1]) ||
		   isnan(z[(size_t) (ix + 1) * ny + iy]) || isnan(z[(size_t) (ix + 1) * ny + iy + 1]);
}
void orc_surface_normal(const orc_terrain *t, double x, double y, double n[3], unsigned *flags) { /* :160-213, not renormalised */
	int ix = find_cell(t->x, t->nx, x, flags), iy = find_cell(t->y, t->ny, y, flags);
	double x1 = t->x[ix], x2 = t->x[ix + 1], y1 = t->y[iy], y2 = t->y[iy + 1];
	n[0] = bilinear(t->dx, t->ny, ix, iy, x1, x2, y1, y2, x, y);
	n[1] = bilinear(t->dy, t->ny, ix, iy, x1, x2, y1, y2, x, y);
	n[2] = bilinear(t->dz, t->ny, ix, iy, x1, x2, y1, y2, x, y);
}

/* ====================================================================== primitives */
void orc_apply_stance(const double s[8], const double a[10], double t, double o[8]) { /* planning_utils.cpp:237-274 */
	double ts = a[6];
	/* position rows: p + v t + 0.5 a_td t t + (a_to - a_td)(t t t)/(6 ts); velocity rows: v + a_td t + (a_to-a_td) t t/(2 ts) */
	for (int d = 0; d < 3; ++d) {
		o[d] = s[d] + s[3 + d] * t + 0.5 * a[d] * t * t + (a[3 + d] - a[d]) * (t * t * t) / (6.0 * ts);
		o[3 + d] = s[3 + d] + a[d] * t + (a[3 + d] - a[d]) * t * t / (2.0 * ts);
	}
	o[6] = s[6] + s[7] * t + 0.5 * a[8] * t * t + (a[9] - a[8]) * (t * t * t) / (6.0 * ts);
	o[7] = s[7] + a[8] * t + (a[9] - a[8]) * t * t / (2.0 * ts);
}
void orc_apply_flight(const double s[8], double t, double o[8]) { /* planning_utils.cpp:282-306, literal g = 9.81 */
	double g = 9.81;
	o[0] = s[0] + s[3] * t;
	o[1] = s[1] + s[4] * t;
	o[2] = s[2] + s[5] * t - 0.5 * g * t * t;
	o[3] = s[3];
	o[4] = s[4];
	o[5] = s[5] - g * t;
	o[6] = s[6] + s[7] * t;
	o[7] = s[7];
}
void orc_apply_stance_reverse(const double s[8], const double a[10], double t, double o[8]) { /* planning_utils.cpp:324-367 */
	double ts = a[6];
	for (int d = 0; d < 4; ++d) {
		int ip = d < 3 ? d : 6, iv = d < 3 ? 3 + d : 7, itd = d < 3 ? d : 8, ito = d < 3 ? 3 + d : 9;
		double c = s[iv] - a[itd] * ts - 0.5 * (a[ito] - a[itd]) * ts;
		o[ip] = s[ip] - c * (ts - t) - 0.5 * a[itd] * (ts * ts - t * t) - (a[ito] - a[itd]) * (ts * ts * ts - t * t * t) / (6.0 * ts);
		o[iv] = s[iv] - a[itd] * (ts - t) - (a[ito] - a[itd]) * (ts * ts - t * t) / (2.0 * ts);
	}
}

/* planning_utils.cpp:198-231.  3x3 algebra written out in the order a plain (non-vectorised)
 * evaluation of `I + vskew + vskew*vskew*(1-c)/(s*s)` and `R*f` performs it.  Parity vs real Eigen
 * is unpinned (Eigen is not in /root/reference); it is pinned vs oracle/shim/Eigen/Dense. */
void orc_rotate_grf(const double n[3], const double f[3], double out[3]) {
	double zs[3] = {0.0, 0.0, 1.0};
	double v[3] = {n[1] * zs[2] - n[2] * zs[1], n[2] * zs[0] - n[0] * zs[2], n[0] * zs[1] - n[1] * zs[0]};
	double s = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	double c = n[0] * zs[0] + n[1] * zs[1] + n[2] * zs[2];
	double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	if (!(s < 0.000001)) {
		double K[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0}, KK[9];
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j)
				KK[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
		for (int i = 0; i < 9; ++i) R[i] = (R[i] + K[i]) + KK[i] * (1 - c) / (s * s);
	}
	for (int i = 0; i < 3; ++i) out[i] = R[3 * i] * f[0] + R[3 * i + 1] * f[1] + R[3 * i + 2] * f[2];
}

int orc_is_valid_action(const double a[10]) { /* planning_utils.cpp:519-556 */
	if (a[6] <= 0 || a[7] < 0) return 0;
	double m = ORC_M_CONST, g = ORC_G_CONST, mu = ORC_MU;
	double fxd = m * a[0], fyd = m * a[1], fzd = m * (a[2] + g);
	double fxo = m * a[3], fyo = m * a[4], fzo = m * (a[5] + g);
	if (sqrt(fxd * fxd + fyd * fyd + fzd * fzd) >= ORC_F_MAX || sqrt(fxo * fxo + fyo * fyo + fzo * fzo) >= ORC_F_MAX ||
		fzd < 0 || fzo < 0 || a[8] >= ORC_F_MAX || a[9] >= ORC_F_MAX) /* sic: pitch accel vs F_MAX, no abs (:543) */
		return 0;
	if (sqrt(fxd * fxd + fyd * fyd) >= mu * fzd || sqrt(fxo * fxo + fyo * fyo) >= mu * fzo) return 0;
	return 1;
}

int orc_is_valid_state(const orc_terrain *t, const double s[8], int phase, orc_counters *c) { /* planning_utils.cpp:562-635 */
	orc_counters local = {0, 0, 0, 0};
	if (!c) c = &local;
	c->substates++;
	c->nanprobes++;
	if (orc_height_is_nan(t, s[0], s[1], &c->flags)) return 0;
	if (s[0] < t->x[0] || s[0] > t->x[t->nx - 1] || s[1] < t->y[0] || s[1] > t->y[t->ny - 1] || fabs(s[6]) >= ORC_P_MAX) return 0;
	if (sqrt(s[3] * s[3] + s[4] * s[4]) > ORC_V_MAX) return 0;
	double yaw = atan2(s[4], s[3]);
	double cy = cos(yaw), sy = sin(yaw), pitch = s[6], cp = cos(pitch), sp = sin(pitch);
	double R11 = cy * cp, R12 = -sy, R13 = cy * sp, R21 = sy * cp, R22 = cy, R23 = sy * sp, R31 = -sp, R32 = 0, R33 = cp;
	double tx[2] = {-0.5 * ORC_ROBOT_L, 0.5 * ORC_ROBOT_L}, ty[2] = {-0.5 * ORC_ROBOT_W, 0.5 * ORC_ROBOT_W}, zb = -ORC_ROBOT_H;
	for (int i = 0; i < 2; ++i)
		for (int j = 0; j < 2; ++j) {
			double xb = tx[i], yb = ty[j];
			double xl = s[0] + R11 * xb + R12 * yb, yl = s[1] + R21 * xb + R22 * yb, zl = s[2] + R31 * xb + R32 * yb;
			double xc = xl + R13 * zb, yc = yl + R23 * zb, zc = zl + R33 * zb;
			c->nanprobes++;
			if (orc_height_is_nan(t, xl, yl, &c->flags)) return 0;
			c->lookups += 2;
			double leg_h = zl - orc_ground_height(t, xl, yl, &c->flags);
			double cor_h = zc - orc_ground_height(t, xc, yc, &c->flags);
			if (cor_h < ORC_H_MIN || (phase == ORC_STANCE && leg_h > ORC_H_MAX)) return 0;
		}
	c->lookups += 1;
	double h = (s[2] + R33 * zb) - orc_ground_height(t, s[0] + R13 * zb, s[1] + R23 * zb, &c->flags);
	return h < ORC_H_MIN ? 0 : 1;
}

/* forward fixed step planning_utils.cpp:713-753; adaptive :651-712 */
static int pair_forward(const orc_terrain *T, const double s[8], const double a[10], int adaptive, double s_new[8],
						double *t_new, orc_counters *c) {
	double ts = a[6], tf = a[7], step = ORC_KINEMATICS_RES, t_ok = 0, chk[8];
	for (double t = 0; t <= ts; t += step) {
		orc_apply_stance(s, a, t, chk);
		if (!orc_is_valid_state(T, chk, ORC_STANCE, c)) {
			if (!adaptive || (ORC_KINEMATICS_RES - 0.01 <= step && step <= ORC_KINEMATICS_RES + 0.01)) {
				orc_apply_stance(s, a, (1.0 - ORC_BACKUP_RATIO) * t, s_new);
				return 0;
			}
			step = ORC_KINEMATICS_RES; /* :672-675 rewind to the last success */
			t = t_ok;
		} else {
			memcpy(s_new, chk, sizeof chk);
			*t_new = t;
			if (adaptive) { step += ORC_KINEMATICS_RES; t_ok = t; }
		}
	}
	double takeoff[8];
	orc_apply_stance(s, a, ts, takeoff);
	step = ORC_KINEMATICS_RES;
	for (double t = 0; t < tf; t += step) {
		orc_apply_flight(takeoff, t, chk);
		if (!orc_is_valid_state(T, chk, ORC_FLIGHT, c)) return 0;
		if (adaptive) step += ORC_KINEMATICS_RES;
	}
	orc_apply_flight(takeoff, tf, chk);
	if (!orc_is_valid_state(T, chk, ORC_STANCE, c)) return 0;
	memcpy(s_new, chk, sizeof chk);
	*t_new = ts + tf;
	return 1;
}
/* reverse fixed step planning_utils.cpp:837-876; adaptive :774-836.  `s` is the END state. */
static int pair_reverse(const orc_terrain *T, const double s[8], const double a[10], int adaptive, double s_new[8],
						double *t_new, orc_counters *c) {
	double ts = a[6], tf = a[7], step = ORC_KINEMATICS_RES, t_ok = 0, chk[8];
	for (double t = 0; t < tf; t += step) {
		orc_apply_flight(s, -t, chk);
		if (!orc_is_valid_state(T, chk, ORC_FLIGHT, c)) return 0;
		if (adaptive) step += ORC_KINEMATICS_RES;
	}
	double takeoff[8];
	orc_apply_flight(s, -tf, takeoff);
	step = ORC_KINEMATICS_RES;
	for (double t = ts; t >= 0; t -= step) {
		orc_apply_stance_reverse(takeoff, a, t, chk);
		if (!orc_is_valid_state(T, chk, ORC_STANCE, c)) {
			if (!adaptive || (ORC_KINEMATICS_RES - 0.01 <= step && step <= ORC_KINEMATICS_RES + 0.01)) {
				orc_apply_stance(s, a, t + ORC_BACKUP_RATIO * (ts - t), s_new); /* sic: forward stance from the end state (:862) */
				return 0;
			}
			step = ORC_KINEMATICS_RES;
			t = t_ok;
		} else {
			memcpy(s_new, chk, sizeof chk);
			*t_new = ts - t;
			if (adaptive) { step += ORC_KINEMATICS_RES; t_ok = t; }
		}
	}
	orc_apply_stance_reverse(takeoff, a, 0, chk);
	if (!orc_is_valid_state(T, chk, ORC_STANCE, c)) return 0;
	memcpy(s_new, chk, sizeof chk);
	*t_new = ts;
	return 1;
}
/* Outputs the reference leaves unwritten are DEFINED here: s_new starts as s, t_new as 0. */
int orc_validate_pair(const orc_terrain *t, const double s[8], const double a[10], int direction, int adaptive,
					  double s_new[8], double *t_new, orc_counters *c) {
	orc_counters local = {0, 0, 0, 0};
	if (!c) c = &local;
	memcpy(s_new, s, 8 * sizeof(double));
	*t_new = 0.0;
	return direction == ORC_FORWARD ? pair_forward(t, s, a, adaptive, s_new, t_new, c) : pair_reverse(t, s, a, adaptive, s_new, t_new, c);
}

double orc_pose_distance(const double a[8], const double b[8]) { /* planning_utils.cpp:106-115 */
	double sum = 0;
	for (int i = 0; i < 3; ++i) sum = sum + (b[i] - a[i]) * (b[i] - a[i]);
	return sqrt(sum);
}
double orc_state_distance(const double a[8], const double b[8]) { /* planning_utils.cpp:116-127 (weights all 1) */
	double sum = 0;
	for (int i = 0; i < 8; ++i) sum = sum + 1.0 * (b[i] - a[i]) * (b[i] - a[i]);
	return sqrt(sum);
}
double orc_yaw_distance(const double a[8], const double b[8]) { /* planning_utils.h:133-145 */
	double y1 = atan2(a[4], a[3]), y2 = atan2(b[4], b[3]);
	double lo = y1 < y2 ? y1 : y2, hi = y1 < y2 ? y2 : y1; /* std::min/std::max */
	double d1 = hi - lo, d2 = lo + 2 * ORC_MY_PI - hi;
	return d2 < d1 ? d2 : d1;
}

/* ====================================================================== samplers on the Philox stream
 * Draw layout of an ACTION cell (purpose 1): u0 f_z_td, u1 f_z_to, u2 f_x_td, u3 f_x_to, u4 f_y_td,
 * u5 f_y_to, u6 t_f, (u7,u8) Box-Muller pair -> a[8], a[9], u9 direction-sampling probability.
 * Recipe: planning_utils.cpp:392-442 (plain) and :443-515 (directional). */
static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static void box_muller(double ua, double ub, double *z0, double *z1) {
	double r = sqrt(-2.0 * orc_det_log(1.0 - ua)), sn, cs;
	orc_det_sincos(6.283185307179586 * ub, &sn, &cs);
	*z0 = r * cs;
	*z1 = r * sn;
}
void orc_sample_action(uint64_t seed, uint64_t stream, uint64_t idx, const double normal[3], int dir_flag,
					   double dir_thresh, const double s_from[8], const double s_to[8], double a[10]) {
	double u[10];
	orc_uniforms(seed, stream, idx, 1, 0, 10, u);
	double fzd = ORC_F_MAX * u[0], fzo = ORC_F_MAX * u[1], fxd, fxo, fyd, fyo;
	if (dir_flag && u[9] <= dir_thresh) { /* :383 and :450-487 */
		double frd = ORC_MU * fzd, fro = ORC_MU * fzo;
		if (s_to[3] > s_from[3]) { fxd = frd * u[2]; fxo = fro * u[3]; } else { fxd = frd * u[2] - frd; fxo = fro * u[3] - fro; }
		if (s_to[4] > s_from[4]) { fyd = frd * u[4]; fyo = fro * u[5]; } else { fyd = frd * u[4] - frd; fyo = fro * u[5] - fro; }
	} else { /* :400-403 */
		fxd = 2 * ORC_MU * fzd * u[2] - ORC_MU * fzd;
		fxo = 2 * ORC_MU * fzo * u[3] - ORC_MU * fzo;
		fyd = 2 * ORC_MU * fzd * u[4] - ORC_MU * fzd;
		fyo = 2 * ORC_MU * fzo * u[5] - ORC_MU * fzo;
	}
	double ftd[3] = {fxd, fyd, fzd}, fto[3] = {fxo, fyo, fzo}, rtd[3], rto[3];
	orc_rotate_grf(normal, ftd, rtd);
	orc_rotate_grf(normal, fto, rto);
	a[0] = rtd[0] / ORC_M_CONST;
	a[1] = rtd[1] / ORC_M_CONST;
	a[2] = rtd[2] / ORC_M_CONST - ORC_G_CONST;
	a[3] = rto[0] / ORC_M_CONST;
	a[4] = rto[1] / ORC_M_CONST;
	a[5] = rto[2] / ORC_M_CONST - ORC_G_CONST;
	a[6] = 0.3; /* :417 */
	a[7] = (ORC_T_F_MAX - ORC_T_F_MIN) * u[6] + ORC_T_F_MIN;
	double z0, z1, sd = ORC_ANG_ACC_MAX / 4.0; /* :437-439 */
	box_muller(u[7], u[8], &z0, &z1);
	a[8] = clampd(sd * z0, -ORC_ANG_ACC_MAX, ORC_ANG_ACC_MAX);
	a[9] = clampd(sd * z1, -ORC_ANG_ACC_MAX, ORC_ANG_ACC_MAX);
}
/* STATE cell (purpose 2): u0 x, u1 y, (u2,u3) Box-Muller -> height (z0 only), u4 phi, u5 cos(theta),
 * u6 speed, u7 pitch, u8 direction-sampling probability.  Recipe: planner_class.cpp:38-76 and
 * :82-148; sin(acos(c)) is restated as sqrt(1 - c*c) and cos(acos(c)) as c (deterministic). */
void orc_sample_state(const orc_terrain *t, uint64_t seed, uint64_t stream, uint64_t idx, int dir_flag,
					  double dir_thresh, int speed_dir_flag, const double s_from[8], const double s_to[8], double q[8]) {
	double u[9];
	orc_uniforms(seed, stream, idx, 2, 0, 9, u);
	int directional = dir_flag && u[8] <= dir_thresh;
	double x_min = t->x[0], x_max = t->x[t->nx - 1], y_min = t->y[0], y_max = t->y[t->ny - 1];
	if (directional) { /* :89-92 */
		x_min = s_from[0] < s_to[0] ? s_from[0] : s_to[0];
		x_max = s_from[0] < s_to[0] ? s_to[0] : s_from[0];
		y_min = s_from[1] < s_to[1] ? s_from[1] : s_to[1];
		y_max = s_from[1] < s_to[1] ? s_to[1] : s_from[1];
	}
	double z_min_rel = ORC_H_MIN + ORC_ROBOT_H, z_max_rel = ORC_H_MAX + ORC_ROBOT_H;
	double mean = 0.5 * (z_max_rel + z_min_rel), sd = (z_max_rel - z_min_rel) * (1.0 / (2 * 3.0));
	double z0, z1, sn, cs;
	box_muller(u[2], u[3], &z0, &z1);
	q[0] = (x_max - x_min) * u[0] + x_min;
	q[1] = (y_max - y_min) * u[1] + y_min;
	q[2] = clampd(mean + sd * z0, z_min_rel, z_max_rel) + orc_ground_height(t, q[0], q[1], 0);
	double cos_theta = 2.0 * u[5] - 1.0, sin_theta = sqrt(1.0 - cos_theta * cos_theta), v = u[6] * ORC_V_MAX;
	if (directional && speed_dir_flag) { /* :115-119: phi = atan2(dy, dx) restated as the unit vector */
		double ddx = s_to[0] - s_from[0], ddy = s_to[1] - s_from[1], nrm = sqrt(ddx * ddx + ddy * ddy);
		if (nrm > 0) { cs = ddx / nrm; sn = ddy / nrm; } else { cs = 1.0; sn = 0.0; }
	} else {
		orc_det_sincos((2.0 * ORC_MY_PI) * u[4], &sn, &cs);
	}
	q[3] = v * sin_theta * cs;
	q[4] = v * sin_theta * sn;
	q[5] = v * cos_theta;
	q[6] = 2 * ORC_P_MAX * u[7] - ORC_P_MAX;
	q[7] = 0.0;
}

/* ====================================================================== tree queries
 * planner_class.cpp:185-200: argmin of stateDistance with strict '<'.  Iteration order of the
 * reference is that of std::unordered_map; DEFINED here: ascending id, so the
 * lowest id wins ties.  *unique tells whether the minimum is attained once (only then must the
 * index equal the reference's). */
int orc_nearest(const double *verts, long long nv, const double q[8], double *dist, int *unique) {
	int best = 0, uniq = 1;
	double bd = INFINITY;
	for (long long i = 0; i < nv; ++i) {
		double d = orc_state_distance(q, verts + 8 * i);
		if (d < bd) { bd = d; best = (int) i; uniq = 1; }
		else if (d == bd) uniq = 0;
	}
	if (dist) *dist = bd;
	if (unique) *unique = uniq;
	return best;
}
long long orc_near(const double *verts, long long nv, const double q[8], double radius, int *ids, long long cap) { /* :173-182 */
	long long n = 0;
	for (long long i = 0; i < nv; ++i) {
		double d = orc_state_distance(q, verts + 8 * i);
		if (d <= radius && d > 0) { if (n < cap) ids[n] = (int) i; ++n; }
	}
	return n;
}

/* ====================================================================== connect
 * rrt_connect.cpp:20-91.  The tail recursion (:78) is unrolled: every level recomputes the
 * closed-form action toward the latest partial state with t_s := t_new; an invalid action or
 * t_s <= KINEMATICS_RES at any depth makes the whole call TRAPPED; success at depth 0 is REACHED,
 * deeper is ADVANCED. */
int orc_attempt_connect(const orc_terrain *T, const double s_existing[8], const double s_in[8], int direction,
						int adaptive, double s_new[8], double a_new[10], orc_counters *c) {
	double target[8], ts = orc_pose_distance(s_in, s_existing) / ORC_V_NOM; /* :89 */
	memcpy(target, s_in, sizeof target);
	for (int depth = 0;; ++depth) {
		if (ts <= ORC_KINEMATICS_RES) return ORC_TRAPPED;
		const double *st = direction == ORC_FORWARD ? s_existing : target;
		const double *go = direction == ORC_FORWARD ? target : s_existing;
		for (int d = 0; d < 4; ++d) { /* :53-63 */
			int ip = d < 3 ? d : 6, iv = d < 3 ? 3 + d : 7, itd = d < 3 ? d : 8, ito = d < 3 ? 3 + d : 9;
			a_new[itd] = -(2.0 * (3.0 * st[ip] - 3.0 * go[ip] + 2.0 * st[iv] * ts + go[iv] * ts)) / (ts * ts);
			a_new[ito] = (2.0 * (3.0 * st[ip] - 3.0 * go[ip] + st[iv] * ts + 2.0 * go[iv] * ts)) / (ts * ts);
		}
		a_new[6] = ts;
		a_new[7] = 0;
		if (!orc_is_valid_action(a_new)) return ORC_TRAPPED;
		double out[8], tn;
		int ok = orc_validate_pair(T, direction == ORC_FORWARD ? st : go, a_new, direction, adaptive, out, &tn, c);
		memcpy(s_new, out, sizeof out);
		if (ok) return depth == 0 ? ORC_REACHED : ORC_ADVANCED;
		memcpy(target, out, sizeof out);
		ts = tn;
	}
}

/* ====================================================================== batch drivers */
typedef struct {
	const orc_terrain *t; long long lo, hi; const double *s, *a; const unsigned char *dir; int adaptive;
	unsigned char *verdict, *flags; double *s_new, *t_new; orc_counters cnt;
} vp_job;
static void *vp_run(void *arg) {
	vp_job *j = (vp_job *) arg;
	for (long long i = j->lo; i < j->hi; ++i) {
		orc_counters c = {0, 0, 0, 0};
		double sn[8], tn;
		int ok = orc_validate_pair(j->t, j->s + 8 * i, j->a + 10 * i, j->dir[i], j->adaptive, sn, &tn, &c);
		j->verdict[i] = (unsigned char) ok;
		if (j->flags) j->flags[i] = (unsigned char) c.flags;
		if (j->s_new) memcpy(j->s_new + 8 * i, sn, sizeof sn);
		if (j->t_new) j->t_new[i] = tn;
		j->cnt.substates += c.substates; j->cnt.lookups += c.lookups; j->cnt.nanprobes += c.nanprobes;
	}
	return 0;
}
void orc_validate_pairs(const orc_terrain *t, long long n, const double *s, const double *a, const unsigned char *dir,
						int adaptive, unsigned char *verdict, unsigned char *flags, double *s_new, double *t_new,
						long long *counters3, int nthreads) {
	if (nthreads < 1) nthreads = 1;
	vp_job *jobs = (vp_job *) calloc((size_t) nthreads, sizeof(vp_job));
	pthread_t *th = (pthread_t *) calloc((size_t) nthreads, sizeof(pthread_t));
	for (int k = 0; k < nthreads; ++k) {
		vp_job j = {t, n * k / nthreads, n * (k + 1) / nthreads, s, a, dir, adaptive, verdict, flags, s_new, t_new, {0, 0, 0, 0}};
		jobs[k] = j;
		if (nthreads > 1) pthread_create(&th[k], 0, vp_run, &jobs[k]); else vp_run(&jobs[k]);
	}
	long long tot[3] = {0, 0, 0};
	for (int k = 0; k < nthreads; ++k) {
		if (nthreads > 1) pthread_join(th[k], 0);
		tot[0] += jobs[k].cnt.substates; tot[1] += jobs[k].cnt.lookups; tot[2] += jobs[k].cnt.nanprobes;
	}
	if (counters3) memcpy(counters3, tot, sizeof tot);
	free(jobs);
	free(th);
}
void orc_sample_actions(uint64_t seed, uint64_t stream, uint64_t idx0, long long n, const double normal[3], double *a) {
	for (long long i = 0; i < n; ++i) orc_sample_action(seed, stream, idx0 + (uint64_t) i, normal, 0, 0.0, 0, 0, a + 10 * i);
}
void orc_sample_states(const orc_terrain *t, uint64_t seed, uint64_t stream, uint64_t idx0, long long n, double *q) {
	for (long long i = 0; i < n; ++i) orc_sample_state(t, seed, stream, idx0 + (uint64_t) i, 0, 0.0, 0, 0, 0, q + 8 * i);
}
void orc_valid_states(const orc_terrain *t, long long n, const double *s, const unsigned char *phase,
					  unsigned char *verdict, unsigned char *flags) {
	for (long long i = 0; i < n; ++i) {
		orc_counters c = {0, 0, 0, 0};
		verdict[i] = (unsigned char) orc_is_valid_state(t, s + 8 * i, phase[i], &c);
		if (flags) flags[i] = (unsigned char) c.flags;
	}
}

/* ====================================================================== Tier-2 planner */
typedef struct {
	int n, cap;
	double *v, *act, *g, *y; /* [cap][8], [cap][10], [cap], [cap] */
	int *parent, *child, *sibling; /* tree links; child/sibling lists replace GraphClass::successors */
} orc_tree;

static void tree_alloc(orc_tree *T, int cap) {
	T->n = 0; T->cap = cap;
	T->v = (double *) malloc(sizeof(double) * 8 * cap);
	T->act = (double *) calloc((size_t) 10 * cap, sizeof(double));
	T->g = (double *) calloc((size_t) cap, sizeof(double));
	T->y = (double *) calloc((size_t) cap, sizeof(double));
	T->parent = (int *) malloc(sizeof(int) * cap);
	T->child = (int *) malloc(sizeof(int) * cap);
	T->sibling = (int *) malloc(sizeof(int) * cap);
}
static void tree_free(orc_tree *T) { free(T->v); free(T->act); free(T->g); free(T->y); free(T->parent); free(T->child); free(T->sibling); }
static void tree_init(orc_tree *T, const double s[8]) { /* graph_class.cpp:140-152 */
	T->n = 1;
	memcpy(T->v, s, 64);
	T->g[0] = 0; T->y[0] = 0; T->parent[0] = -1; T->child[0] = -1; T->sibling[0] = -1;
}
static void tree_link(orc_tree *T, int p, int c) { T->parent[c] = p; T->sibling[c] = T->child[p]; T->child[p] = c; }
static void tree_unlink(orc_tree *T, int p, int c) { /* graph_class.cpp:44-59 */
	int *it = &T->child[p];
	while (*it != -1 && *it != c) it = &T->sibling[*it];
	if (*it == c) *it = T->sibling[c];
	T->sibling[c] = -1;
}
static void tree_update_gy(orc_tree *T, int i, double g, double y) { /* graph_class.cpp:131-138 */
	T->g[i] = g; T->y[i] = y;
	for (int c = T->child[i]; c != -1; c = T->sibling[c])
		tree_update_gy(T, c, T->g[i] + orc_pose_distance(T->v + 8 * i, T->v + 8 * c), T->y[i] + orc_yaw_distance(T->v + 8 * i, T->v + 8 * c));
}
/* addVertex + addEdge + addAction + updateGYValue as rrt.cpp:87-92 / rrt_connect.cpp:110-116 */
static int tree_append(orc_tree *T, int parent, const double s[8], const double a[10]) {
	int i = T->n++;
	memcpy(T->v + 8 * i, s, 64);
	memcpy(T->act + 10 * i, a, 80);
	T->child[i] = -1; T->sibling[i] = -1;
	tree_link(T, parent, i);
	T->g[i] = T->g[parent] + orc_pose_distance(T->v + 8 * parent, s);
	T->y[i] = T->y[parent] + orc_yaw_distance(T->v + 8 * parent, s);
	return i;
}

typedef struct { const orc_terrain *t; const orc_plan_params *p; uint64_t seed, query; orc_plan_stats *st; } plan_ctx;

static int checked_pair(plan_ctx *C, const double s[8], const double a[10], int dir, double sn[8], double *tn) {
	C->st->pair_checks++;
	return orc_validate_pair(C->t, s, a, dir, C->p->adaptive, sn, tn, 0);
}
static int checked_connect(plan_ctx *C, const double s_existing[8], const double s[8], int dir, double sn[8], double an[10]) {
	/* same as orc_attempt_connect but counting pair checks */
	double target[8], ts = orc_pose_distance(s, s_existing) / ORC_V_NOM;
	memcpy(target, s, sizeof target);
	for (int depth = 0;; ++depth) {
		if (ts <= ORC_KINEMATICS_RES) return ORC_TRAPPED;
		const double *st = dir == ORC_FORWARD ? s_existing : target, *go = dir == ORC_FORWARD ? target : s_existing;
		for (int d = 0; d < 4; ++d) {
			int ip = d < 3 ? d : 6, iv = d < 3 ? 3 + d : 7, itd = d < 3 ? d : 8, ito = d < 3 ? 3 + d : 9;
			an[itd] = -(2.0 * (3.0 * st[ip] - 3.0 * go[ip] + 2.0 * st[iv] * ts + go[iv] * ts)) / (ts * ts);
			an[ito] = (2.0 * (3.0 * st[ip] - 3.0 * go[ip] + st[iv] * ts + 2.0 * go[iv] * ts)) / (ts * ts);
		}
		an[6] = ts; an[7] = 0;
		if (!orc_is_valid_action(an)) return ORC_TRAPPED;
		double out[8], tn;
		int ok = checked_pair(C, dir == ORC_FORWARD ? st : go, an, dir, out, &tn);
		memcpy(sn, out, sizeof out);
		if (ok) return depth == 0 ? ORC_REACHED : ORC_ADVANCED;
		memcpy(target, out, sizeof out);
		ts = tn;
	}
}

/* rrt.cpp:20-70 (only the first outer pass ever runs — SURVEY Appendix B-3), generalised to K
 * candidates: candidate j of this extend is ACTION cell idx = cell*K + j. */
static int new_config(plan_ctx *C, uint64_t cell, const double s[8], const double s_near[8], int dir, double s_new[8], double a_new[10]) {
	double best = orc_state_distance(s_near, s), normal[3];
	orc_surface_normal(C->t, s[0], s[1], normal, 0); /* rrt.cpp:25 — at the TARGET sample */
	int K = C->p->k_candidates, found = 0;
	for (int j = 0; j < K; ++j) {
		double a[10], st[8], tn;
		/* getRandomAction(surf_norm, direction, flag, threshold, s, s_near): FORWARD samples from s_near towards s,
		 * REVERSE from s towards s_near (planning_utils.cpp:385-388) */
		orc_sample_action(C->seed, C->query, cell * (uint64_t) K + (uint64_t) j, normal, C->p->action_direction_sampling,
						  C->p->action_direction_threshold, dir == ORC_FORWARD ? s_near : s, dir == ORC_FORWARD ? s : s_near, a);
		if (!checked_pair(C, s_near, a, dir, st, &tn)) continue;
		double d = orc_state_distance(st, s);
		if (d < best) { best = d; memcpy(s_new, st, 64); memcpy(a_new, a, 80); found = 1; }
		if (!C->p->best_of_k) break; /* first valid action decides (rrt.cpp:44-47) */
	}
	return found;
}

static int extend_plain(plan_ctx *C, orc_tree *T, uint64_t cell, const double s[8], int dir) { /* rrt.cpp:77-102 */
	C->st->nn_queries++;
	int near = orc_nearest(T->v, T->n, s, 0, 0);
	double s_new[8], a_new[10];
	if (!new_config(C, cell, s, T->v + 8 * near, dir, s_new, a_new)) return ORC_TRAPPED;
	tree_append(T, near, s_new, a_new);
	return orc_state_distance(s_new, s) <= ORC_GOAL_BOUNDS ? ORC_REACHED : ORC_ADVANCED;
}

static int extend_star(plan_ctx *C, orc_tree *T, uint64_t cell, const double s[8], int dir) { /* rrt_star_connect.cpp:12-75 */
	C->st->nn_queries++;
	int nearest = orc_nearest(T->v, T->n, s, 0, 0);
	double s_new[8], a_new[10], a_con[10], dummy[8];
	if (!new_config(C, cell, s, T->v + 8 * nearest, dir, s_new, a_new)) return ORC_TRAPPED;
	int inew = T->n++;
	memcpy(T->v + 8 * inew, s_new, 64);
	T->child[inew] = -1; T->sibling[inew] = -1; T->parent[inew] = -1;
	int *nb = (int *) malloc(sizeof(int) * (size_t) T->n);
	long long nnb = orc_near(T->v, T->n, s_new, ORC_RRT_STAR_DELTA, nb, T->n); /* ascending id (defined order) */
	int imin = nearest;
	double g_new = T->g[nearest] + orc_pose_distance(s_new, T->v + 8 * nearest);
	double y_new = T->y[nearest] + orc_yaw_distance(s_new, T->v + 8 * nearest);
	for (long long i = 0; i < nnb; ++i) {
		const double *sn = T->v + 8 * nb[i];
		if (checked_connect(C, sn, s_new, dir, dummy, a_con) == ORC_REACHED) {
			double g = T->g[nb[i]] + orc_pose_distance(sn, s_new), y = T->y[nb[i]] + orc_yaw_distance(sn, s_new);
			if (g < g_new) { memcpy(a_new, a_con, 80); imin = nb[i]; g_new = g; y_new = y; }
		}
	}
	tree_link(T, imin, inew);
	tree_update_gy(T, inew, g_new, y_new);
	memcpy(T->act + 10 * inew, a_new, 80);
	for (long long i = 0; i < nnb; ++i) {
		int k = nb[i];
		if (k == imin) continue;
		const double *sn = T->v + 8 * k;
		if (checked_connect(C, s_new, sn, dir, dummy, a_con) == ORC_REACHED && T->g[k] > T->g[inew] + orc_pose_distance(sn, s_new)) {
			tree_unlink(T, T->parent[k], k);
			tree_link(T, inew, k);
			tree_update_gy(T, k, T->g[inew] + orc_pose_distance(sn, s_new), T->y[inew] + orc_yaw_distance(sn, s_new));
			memcpy(T->act + 10 * k, a_con, 80);
		}
	}
	free(nb);
	return orc_state_distance(s_new, s) <= ORC_GOAL_BOUNDS ? ORC_REACHED : ORC_ADVANCED;
}

static int connect_tree(plan_ctx *C, orc_tree *T, const double s[8], int dir) { /* rrt_connect.cpp:98-120 */
	C->st->nn_queries++;
	int near = orc_nearest(T->v, T->n, s, 0, 0);
	double s_new[8], a_new[10];
	int r = checked_connect(C, T->v + 8 * near, s, dir, s_new, a_new);
	if (r != ORC_TRAPPED) tree_append(T, near, s_new, a_new);
	return r;
}

/* postProcessPath, rrt_connect.cpp:139-227 (FORWARD shortcutting; quirk kept: the fallback branch
 * adds to path_cost_ only, not to path_length_/path_yaw_). */
int orc_post_process_path(const orc_terrain *t, int ns, double *states, double *actions, int adaptive, double stats3[3]) {
	return orc_post_process_path_w(t, ns, states, actions, adaptive, 0, 1.0, 1.0, stats3);
}
int orc_post_process_path_w(const orc_terrain *t, int ns, double *states, double *actions, int adaptive, int cost_add_yaw,
							double w_length, double w_yaw, double stats3[3]) {
	double *ns_states = (double *) malloc(sizeof(double) * 8 * (size_t) (ns + 1)), *ns_actions = (double *) malloc(sizeof(double) * 10 * (size_t) (ns + 1));
	int m = 1;
	double cur[8], goal[8], len = 0, yaw = 0, cost = 0;
	memcpy(cur, states, 64);
	memcpy(goal, states + 8 * (ns - 1), 64);
	memcpy(ns_states, cur, 64);
	while (memcmp(cur, goal, 64) != 0 && m <= ns) {
		int j = ns - 1; /* s_next = states[j]; a_next = actions[j-1] */
		double a_new[10], dummy[8];
		int have_old = 0, jold = 0;
		while (orc_attempt_connect(t, cur, states + 8 * j, ORC_FORWARD, adaptive, dummy, a_new, 0) != ORC_REACHED &&
			   memcmp(cur, states + 8 * j, 64) != 0) {
			jold = j; have_old = 1;
			--j;
		}
		const double *nxt;
		if (memcmp(cur, states + 8 * j, 64) != 0) {
			nxt = states + 8 * j;
			memcpy(ns_actions + 10 * (m - 1), a_new, 80);
			double dl = orc_pose_distance(cur, nxt), dy = orc_yaw_distance(cur, nxt);
			len += dl; yaw += dy;
			if (cost_add_yaw) cost += dl * w_length + dy * w_yaw; else cost += dl;
		} else {
			if (!have_old) break; /* cannot happen for a well-formed path */
			nxt = states + 8 * jold;
			memcpy(ns_actions + 10 * (m - 1), actions + 10 * (jold - 1), 80);
			double dl = orc_pose_distance(cur, nxt), dy = orc_yaw_distance(cur, nxt);
			if (cost_add_yaw) cost += dl * w_length + dy * w_yaw; else cost += dl;
		}
		memcpy(ns_states + 8 * m, nxt, 64);
		memcpy(cur, nxt, 64);
		++m;
	}
	memcpy(states, ns_states, sizeof(double) * 8 * (size_t) m);
	memcpy(actions, ns_actions, sizeof(double) * 10 * (size_t) (m - 1));
	if (stats3) { stats3[0] = len; stats3[1] = yaw; stats3[2] = cost; }
	free(ns_states);
	free(ns_actions);
	return m;
}

/* runRRTConnect (rrt_connect.cpp:230-314) with an iteration budget; the RRT* main loop
 * (rrt_star_connect.cpp:130-165) has the same body, so rrt_star only swaps the extend.
 * Philox cells: STATE cell idx = 2*iter + half; ACTION cells idx = (2*iter + half)*K + j; stream = query. */
static void tree_dump(const orc_tree *T, orc_tree_dump *d) {
	if (!d) return;
	d->n = T->n;
	for (int i = 0; i < T->n && i < d->cap; ++i) {
		memcpy(d->states + 8 * i, T->v + 8 * i, 64);
		if (i == 0) memset(d->actions, 0, 80); else memcpy(d->actions + 10 * i, T->act + 10 * i, 80);
		d->parent[i] = T->parent[i];
		d->g[i] = T->g[i];
		d->y[i] = T->y[i];
	}
}
int orc_plan(const orc_terrain *t, const double start[8], const double goal[8], uint64_t seed, uint64_t query,
			 const orc_plan_params *p, orc_plan_stats *st, double *path_states, double *path_actions, int path_cap) {
	return orc_plan_ex(t, start, goal, seed, query, p, st, path_states, path_actions, path_cap, 0, 0);
}
int orc_plan_ex(const orc_terrain *t, const double start[8], const double goal[8], uint64_t seed, uint64_t query,
				const orc_plan_params *p, orc_plan_stats *st, double *path_states, double *path_actions, int path_cap,
				orc_tree_dump *dump_a, orc_tree_dump *dump_b) {
	memset(st, 0, sizeof *st);
	plan_ctx C = {t, p, seed, query, st};
	orc_tree Ta, Tb;
	tree_alloc(&Ta, p->max_vertices);
	tree_alloc(&Tb, p->max_vertices);
	tree_init(&Ta, start);
	tree_init(&Tb, goal);
	int solved = 0, full = 0, it = 0;
	/* `iters` counts started iterations; a query stops unsolved when a tree is full at the start of a half */
	for (; it < p->max_iters && !solved && !full; ++it) {
		for (int half = 0; half < 2 && !solved; ++half) {
			orc_tree *Tx = half == 0 ? &Ta : &Tb, *Ty = half == 0 ? &Tb : &Ta;
			int dir_ext = half == 0 ? ORC_FORWARD : ORC_REVERSE, dir_con = half == 0 ? ORC_REVERSE : ORC_FORWARD;
			if (Tx->n >= Tx->cap || Ty->n >= Ty->cap) { full = 1; break; }
			uint64_t cell = 2 * (uint64_t) it + (uint64_t) half;
			double s_rand[8];
			/* directional state sampling between the growing tree's newest vertex and the other tree's root
			 * (rrt_connect.cpp:246-251, :281-286): s_from on the start side, s_to on the goal side */
			const double *s_from = half == 0 ? Ta.v + 8 * (Ta.n - 1) : Ta.v, *s_to = half == 0 ? Tb.v : Tb.v + 8 * (Tb.n - 1);
			orc_sample_state(t, seed, query, cell, p->state_direction_sampling, p->state_direction_threshold, p->state_direction_speed,
							 s_from, s_to, s_rand);
			if (!orc_is_valid_state(t, s_rand, ORC_STANCE, 0)) continue;
			int r = p->rrt_star ? extend_star(&C, Tx, cell, s_rand, dir_ext) : extend_plain(&C, Tx, cell, s_rand, dir_ext);
			if (r == ORC_TRAPPED) continue;
			if (connect_tree(&C, Ty, Tx->v + 8 * (Tx->n - 1), dir_con) == ORC_REACHED) solved = 1;
		}
	}
	st->solved = solved; st->iters = it; st->nv_a = Ta.n; st->nv_b = Tb.n;
	if (solved) {
		st->path_length = Ta.g[Ta.n - 1] + Tb.g[Tb.n - 1]; /* rrt_connect.cpp:269-270 */
		st->path_yaw = Ta.y[Ta.n - 1] + Tb.y[Tb.n - 1];
		st->path_cost = p->cost_add_yaw ? st->path_length * p->cost_length_weight + st->path_yaw * p->cost_yaw_weight : st->path_length; /* :270-274 */
		/* stitch: rrt_connect.cpp:381-401 */
		int na = 0, nb = 0;
		for (int i = Ta.n - 1; i != -1; i = Ta.parent[i]) ++na;
		for (int i = Tb.n - 1; i != -1; i = Tb.parent[i]) ++nb;
		int total = na + nb - 1;
		double *ps = (double *) malloc(sizeof(double) * 8 * (size_t) total), *pa = (double *) malloc(sizeof(double) * 10 * (size_t) total);
		int k = na - 1;
		for (int i = Ta.n - 1; i != -1; i = Ta.parent[i], --k) {
			memcpy(ps + 8 * k, Ta.v + 8 * i, 64);
			if (k > 0) memcpy(pa + 10 * (k - 1), Ta.act + 10 * i, 80); /* action leading INTO vertex i (rrt.cpp:127-135) */
		}
		k = na - 1; /* Tb.last duplicates the shared state: its ACTION is kept, its STATE is dropped (:388-395) */
		for (int i = Tb.n - 1; Tb.parent[i] != -1; i = Tb.parent[i], ++k) {
			memcpy(pa + 10 * k, Tb.act + 10 * i, 80); /* action executed AT vertex i (rrt_connect.cpp:125-133) */
			memcpy(ps + 8 * (k + 1), Tb.v + 8 * Tb.parent[i], 64);
		}
		int nstates = total;
		if (p->post_process) {
			double s3[3];
			nstates = orc_post_process_path_w(t, total, ps, pa, p->adaptive, p->cost_add_yaw, p->cost_length_weight, p->cost_yaw_weight, s3);
			st->path_length = s3[0]; st->path_yaw = s3[1]; st->path_cost = s3[2];
		}
		st->path_states = nstates;
		for (int i = 0; i + 1 < nstates; ++i) st->path_duration += pa[10 * i + 6] + pa[10 * i + 7];
		if (path_states) for (int i = 0; i < nstates && i < path_cap; ++i) memcpy(path_states + 8 * i, ps + 8 * i, 64);
		if (path_actions) for (int i = 0; i + 1 < nstates && i < path_cap; ++i) memcpy(path_actions + 10 * i, pa + 10 * i, 80);
		free(ps);
		free(pa);
	}
	tree_dump(&Ta, dump_a);
	tree_dump(&Tb, dump_b);
	tree_free(&Ta);
	tree_free(&Tb);
	return solved;
}

typedef struct { const orc_terrain *t; long long lo, hi; const double *starts, *goals; uint64_t seed, q0; const orc_plan_params *p; orc_plan_stats *st; } pb_job;
static void *pb_run(void *arg) {
	pb_job *j = (pb_job *) arg;
	for (long long i = j->lo; i < j->hi; ++i) orc_plan(j->t, j->starts + 8 * i, j->goals + 8 * i, j->seed, j->q0 + (uint64_t) i, j->p, j->st + i, 0, 0, 0);
	return 0;
}
void orc_plan_batch(const orc_terrain *t, long long nq, const double *starts, const double *goals, uint64_t seed,
					uint64_t query0, const orc_plan_params *p, orc_plan_stats *st, int nthreads) {
	if (nthreads < 1) nthreads = 1;
	pb_job *jobs = (pb_job *) calloc((size_t) nthreads, sizeof(pb_job));
	pthread_t *th = (pthread_t *) calloc((size_t) nthreads, sizeof(pthread_t));
	for (int k = 0; k < nthreads; ++k) {
		pb_job j = {t, nq * k / nthreads, nq * (k + 1) / nthreads, starts, goals, seed, query0, p, st};
		jobs[k] = j;
		if (nthreads > 1) pthread_create(&th[k], 0, pb_run, &jobs[k]); else pb_run(&jobs[k]);
	}
	for (int k = 0; k < nthreads; ++k) if (nthreads > 1) pthread_join(th[k], 0);
	free(jobs);
	free(th);
}

/* ---- plan output: getInterpPath / interpStateActionPair (src/planning_utils.cpp:142-193).  Returns the number of
 * interpolated states; writes at most cap of them.  phase has one entry fewer than states (the closing state of the
 * sequence gets none, :189-191). */
long long orc_interp_path(int n_actions, const double *states, const double *actions, double dt, long long cap,
						  double *out_s, double *out_t, int *out_phase) {
	long long m = 0;
	double t0 = 0;
	for (int i = 0; i < n_actions; ++i) {
		const double *s = states + 8 * (size_t) i, *a = actions + 10 * (size_t) i;
		const double t_s = a[6], t_f = a[7];
		double takeoff[8];
		for (double t = 0; t < t_s; t += dt) {                                 /* :148-155 */
			if (m < cap) { out_t[m] = t + t0; orc_apply_stance(s, a, t, out_s + 8 * m); out_phase[m] = (t_f == 0) ? 2 : 1; }
			++m;
		}
		orc_apply_stance(s, a, t_s, takeoff);                                  /* :158 */
		for (double t = 0; t < t_f; t += dt) {                                 /* :161-165 */
			if (m < cap) { out_t[m] = t_s + t + t0; orc_apply_flight(takeoff, t, out_s + 8 * m); out_phase[m] = 0; }
			++m;
		}
		if (t_f > 0) {                                                         /* :168-172 */
			if (m < cap) { out_t[m] = t0 + t_s + t_f; orc_apply_flight(takeoff, t_f, out_s + 8 * m); out_phase[m] = 1; }
			++m;
		}
		t0 += (t_s + t_f);                                                     /* :186 */
	}
	if (m < cap) { out_t[m] = t0; memcpy(out_s + 8 * m, states + 8 * (size_t) n_actions, 64); }
	return m + 1;
}

/* calculateCurvature / calculateMaxCurvature (src/planning_utils.cpp:884-909) */
static double curvature3(double x1, double y1, double x2, double y2, double x3, double y3) {
	if ((x1 == x2 && x2 == x3) || (y1 == y2 && y2 == y3)) return 0;
	double dis12 = sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
	double dis13 = sqrt((x1 - x3) * (x1 - x3) + (y1 - y3) * (y1 - y3));
	double dis23 = sqrt((x2 - x3) * (x2 - x3) + (y2 - y3) * (y2 - y3));
	double dis = dis12 * dis12 + dis23 * dis23 - dis13 * dis13;
	double cosA = dis / (2 * dis12 * dis23);
	double sinA = sqrt(1 - cosA * cosA);
	double curvature = 0.5 * dis13 / sinA;
	return 1 / curvature;
}
double orc_max_curvature(long long n, const double *states) {
	double mx = 0;
	for (long long i = 0; i + 2 < n; ++i) {
		double c = curvature3(states[8 * i], states[8 * i + 1], states[8 * (i + 1)], states[8 * (i + 1) + 1], states[8 * (i + 2)], states[8 * (i + 2) + 1]);
		mx = (mx < c) ? c : mx;  /* std::max(mx, c): a NaN c is never taken */
	}
	return mx;
}

/* ====================================================================== terrain generators of the publisher node
 * TerrainMapPublisher::createOwnMap / changeOwnMapZData* / findXYIndex (terrain_map_publisher.cpp:34-231) and
 * createMap (:253-286).  PARITY UNPINNED against compiled reference code: that TU needs ROS, grid_map_ros and OpenCV;
 * the restatement follows the source line by line instead.  The reference draws the box heights from a
 * std::default_random_engine seeded with time(0) (:165), i.e. irreproducibly; here they come from the Philox stream:
 * TERRAIN cell (purpose 3), idx = iy * x_size + ix, stream = rectangle number, Box-Muller pair b = (u_2b, u_2b+1) gives
 * attempts 2b and 2b+1 of the rejection loop (:171-173): val = z * delta + mu, accepted iff mu - delta <= val <= mu + delta;
 * after 32 rejected attempts (p = 1e-16) the cell takes mu.  A rectangle with delta <= 0 or a NaN mu is the constant fill
 * of changeOwnMapZDataRectangle (:129-146). */
static int own_index_lo(const double *ax, int n, double v) { /* findXYIndex :188-197 */
	if (v <= ax[0]) return 0;
	for (int i = 0; i < n - 1; ++i)
		if (ax[i] <= v && v < ax[i + 1]) return i;
	return n - 1; /* v == ax[n-1]: the reference leaves the index uninitialised; defined as the last node */
}
static int own_index_hi(const double *ax, int n, double v) { /* :198-207 */
	if (v >= ax[n - 1]) return n;
	for (int i = n - 1; i > 0; --i)
		if (ax[i - 1] <= v && v < ax[i]) return i;
	return 0; /* unreachable: v < ax[0] returns before the search (:153-156) */
}
const double orc_own_map_default_rects[13][6] = { /* changeOwnMapZData :107-127 */
	{-1.7976931348623157e308, -1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308, 0, 0.01},
	{8.13, -4, 8.42, 4, 0.158, 0.01}, {8.42, -4, 8.71, 4, 0.316, 0.01}, {8.71, -4, 10.5, 4, 0.474, 0.01},
	{0.75, -3.15, 2.05, -2.35, 0.6, 0.1}, {4.25, -2.4, 5.4, -1.75, 0.5, 0.05}, {2.9, -0.6, 3.25, 1.15, 0.158, 0.01},
	{4.9, -0.5, 5.3, 0.35, -0.3, 0.01}, {6.5, 0.45, 7.2, 1.05, 0.7, 0.07}, {0.65, 2.95, 1.15, 3.75, 0.3, 0.08},
	{4.4, 2.8, 5.7, 3.55, 0.65, 0.04}, {7.5, -2.6, 9.45, -1.15, 0.68, 0.06}, {6.2, 1.1, 9.2, 2.3, -0.2, 0.06}};
void orc_own_map_axes(int n, double start, double res, double *ax) { /* :46-60: centimetre-rounded accumulation */
	double v = start;
	for (int i = 0; i < n; ++i) {
		ax[i] = round(v * 100) / 100;
		v = round((v + res) * 100) / 100;
	}
}
/* rectangle r -> [ix1, ix2) x [iy1, iy2) or an empty range when the reference returns early (:153-156) */
void orc_own_map_range(const double *xa, int nx, const double *ya, int ny, const double rect[6], int range[4]) {
	double x1 = rect[0], y1 = rect[1], x2 = rect[2], y2 = rect[3];
	range[0] = range[1] = range[2] = range[3] = 0;
	if (x1 > xa[nx - 1] || x2 < xa[0] || y1 > ya[ny - 1] || y2 < ya[0] || x1 >= x2 || y1 >= y2) return;
	range[0] = own_index_lo(xa, nx, x1); range[1] = own_index_lo(ya, ny, y1);
	range[2] = own_index_hi(xa, nx, x2); range[3] = own_index_hi(ya, ny, y2);
}
double orc_own_map_draw(uint64_t seed, int rect_no, uint64_t cell, double mu, double delta) {
	if (!(delta > 0.0) || mu != mu) return mu;
	double lo = mu - delta, hi = mu + delta;
	for (int b = 0; b < 16; ++b) {
		double u[2], z[2];
		orc_uniforms(seed, (uint64_t) rect_no, cell, 3, 2 * b, 2, u);
		box_muller(u[0], u[1], &z[0], &z[1]);
		for (int k = 0; k < 2; ++k) {
			double val = z[k] * delta + mu;
			if (!(val < lo || val > hi)) return val;
		}
	}
	return mu;
}
/* elevation: float layer in grid_map index order [i * y_size + j] (:88-93); geom = {resolution, centre x, centre y} (:76-78) */
void orc_own_map(uint64_t seed, int x_size, int y_size, double x_start, double y_start, double res, int n_rect,
				 const double *rects, float *elevation, double geom[3]) {
	double *xa = malloc(sizeof(double) * x_size), *ya = malloc(sizeof(double) * y_size);
	double *z = calloc((size_t) x_size * y_size, sizeof(double)); /* z_data[iy][ix], zeros (:63) */
	orc_own_map_axes(x_size, x_start, res, xa);
	orc_own_map_axes(y_size, y_start, res, ya);
	if (!rects) { rects = &orc_own_map_default_rects[0][0]; n_rect = 13; }
	for (int r = 0; r < n_rect; ++r) {
		int g[4];
		orc_own_map_range(xa, x_size, ya, y_size, rects + 6 * r, g);
		for (int i = g[1]; i < g[3]; ++i)
			for (int j = g[0]; j < g[2]; ++j)
				z[(size_t) i * x_size + j] = orc_own_map_draw(seed, r, (uint64_t) i * x_size + j, rects[6 * r + 4], rects[6 * r + 5]);
	}
	double x_end = x_start + res * (x_size - 1), y_end = y_start + res * (y_size - 1); /* :41-44 */
	double x_length = x_end - x_start + res, y_length = y_end - y_start + res;
	geom[0] = res;
	geom[1] = xa[0] - 0.5 * res + 0.5 * x_length;
	geom[2] = ya[0] - 0.5 * res + 0.5 * y_length;
	for (int i = 0; i < x_size; ++i)
		for (int j = 0; j < y_size; ++j)
			elevation[(size_t) i * y_size + j] = (float) z[(size_t) ((y_size - 1) - j) * x_size + ((x_size - 1) - i)];
	free(xa); free(ya); free(z);
}
/* createMap (:253-286): 12 x 5 m at 0.2 m centred on (4, 0) -> 60 x 25 cells; a disc of radius 0.5 m around (2, 0)
 * is 0.1 m high, everything else 0; normals (0, 0, 1).  Cell positions as in the grid_map stand-in of oracle/shim. */
void orc_default_map(float *elevation /* [60 * 25] */, double geom[3]) {
	const int nx = 60, ny = 25;
	const double res = 0.2, cx = 4.0, cy = 0.0;
	for (int i = 0; i < nx; ++i)
		for (int j = 0; j < ny; ++j) {
			double px = cx + (0.5 * (nx - 1) - i) * res, py = cy + (0.5 * (ny - 1) - j) * res;
			double xd = px - 2, yd = py - 0;
			elevation[i * ny + j] = (xd * xd + yd * yd <= 0.5 * 0.5) ? 0.1f : 0.0f;
		}
	geom[0] = res; geom[1] = cx; geom[2] = cy;
}
