// Device-side building blocks of the extend path (sm_100a).  fp64, no FMA contraction (the
// translation unit is compiled with -fmad=false), expressions in the reference's operation order so
// that propagated states, distances and indices are bit-identical to the reference's x86-64 results.
// Citations: file:line under the reference root (LiuShenLan/global_body_planner).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gbp_b200.h"

namespace gbp {

// ---- robot / planner constants (include/global_body_planner/planning_utils.h:21-54)
constexpr double H_MAX = 0.4, H_MIN = 0.075, V_MAX = 2.0, V_NOM = 0.75, P_MAX = 1.0, ANG_ACC_MAX = 7.0;
constexpr double ROBOT_L = 0.3, ROBOT_W = 0.3, ROBOT_H = 0.05;
constexpr double M_CONST = 13.0, G_CONST = 9.81, F_MAX = 637.0, MU = 1.0, T_F_MIN = 0.0, T_F_MAX = 0.5;
constexpr double KINEMATICS_RES = 0.05, BACKUP_RATIO = 0.5, GOAL_BOUNDS = 0.5, MY_PI = 3.14159;
constexpr double RRT_STAR_DELTA = 3.0;  // rrt_star_connect.h:59
constexpr double NEAR_MARGIN = 1e-11;   // GBP_FLAG_NEAR guard band, metres: 10x the 1e-12 m error bound of the fp64 evaluator (coordinates up to
                                        // 10 km, heights up to 1 km; measured error ~1e-13 m, the size of the reference's own rounding noise)
constexpr double HARD_MARGIN = 1e-12;   // guard band of the speed (m/s) and pitch (rad) comparisons of isValidState

// ---- device-resident FastTerrainMap (fast_terrain_map.h:97-118) as SoA grids.
// Height cells are CellT = float when the fp64 input converts losslessly, else double.
struct TerrainView {
	int nx, ny;
	const double *x, *y;  // axes, strictly increasing
	const void *z;        // [nx][ny] x-major, float or double (cell_f32)
	const float *nz3;     // [3][nx][ny] normal layers dx, dy, dz (fp32 when lossless) or null
	const double *nz3d;   // fp64 normal layers when not lossless
	int cell_f32;
	double x0, y0, x_last, y_last;  // axis end points (bounds test of isValidState, OOG test)
	double inv_dx, inv_dy;          // O(1) cell guess: i ~ (v - x0) * inv_dx
	int mixed_ok;                   // no NaN, uniform axes, cell pitch >= 1 cm, fp32 cells or an fp32-rounded texture copy (|z| <= 4 m): the mixed-precision evaluator applies
	int border;                     // cells a body probe can lie from the centre's cell (0.23 m / pitch, rounded up, + 1)
	int uniform;                    // both axes are x0 + i*step to within 1e-12 m: the fast path computes cell edges
	double step_x, step_y;          // instead of loading them (a probe within 1e-11 m of a grid line is flagged NEAR)
	unsigned long long ztex;        // cudaTextureObject_t over a block-linear copy of the fp32 height grid (0 = none): the mixed
	                                // evaluator fetches the 2x2 cells of a probe with ONE tex2Dgather (SASS TLD4) instead of 4 LDGs
};

// map traits: cell storage type x axis kind
template <typename C, bool U> struct MapKind { using cell = C; static constexpr bool uniform = U; };
using MapF32U = MapKind<float, true>;
using MapF32N = MapKind<float, false>;
using MapF64U = MapKind<double, true>;
using MapF64N = MapKind<double, false>;

struct Counters {  // work under the reference's early-exit semantics, per candidate
	unsigned substates, lookups, nanprobes, flags;
};

// ---- Philox4x32-10 stream spec (see include/gbp_b200.h)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
											  uint32_t w[4]) {
	// GBP_PHILOX_UNROLL = 1 keeps the ten rounds a LOOP: the samplers run once per extend, so their code is fetched, not
	// reused — 5 unrolled blocks are ~500 instructions (8 KB) of the planner's instruction-cache footprint
#ifndef GBP_PHILOX_UNROLL
#define GBP_PHILOX_UNROLL 10
#endif
#define GBP_PRAGMA_(x) _Pragma(#x)
#define GBP_UNROLL_(n) GBP_PRAGMA_(unroll n)
	GBP_UNROLL_(GBP_PHILOX_UNROLL)
	for (int r = 0; r < 10; ++r) {
		uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
		uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
		uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
		c0 = n0; c1 = l1; c2 = n2; c3 = l0;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
	return ((double) (hi >> 5) * 67108864.0 + (double) (lo >> 6)) * (1.0 / 9007199254740992.0);
}
// uniforms 2*block and 2*block+1 of cell (seed, stream, idx, purpose)
__device__ __forceinline__ void uniform_pair(uint64_t seed, uint64_t stream, uint64_t idx, int purpose, int block, double &ua,
											 double &ub) {
	uint32_t w[4];
	philox4x32_10((uint32_t) idx, (uint32_t) (idx >> 32), (uint32_t) stream,
				  ((uint32_t) (stream >> 32) & 0x00ffffffu) | ((uint32_t) block << 24) | ((uint32_t) purpose << 28),
				  (uint32_t) seed, (uint32_t) (seed >> 32), w);
	ua = u53(w[0], w[1]);
	ub = u53(w[2], w[3]);
}

// ---- deterministic elementary functions of the sampler spec (+ - * / only; oracle/gbp_oracle.c has
// the same operation sequence, so both sides produce identical bits).
__device__ __forceinline__ double det_log(double x) {
	long long b = __double_as_longlong(x);
	int e = (int) ((b >> 52) & 0x7ff) - 1023;
	double m = __longlong_as_double((b & 0x000fffffffffffffll) | 0x3ff0000000000000ll);
	if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
	double f = m - 1.0, s = f / (2.0 + f), z = s * s;
	double p = 1.0 / 23.0;
	p = p * z + 1.0 / 21.0; p = p * z + 1.0 / 19.0; p = p * z + 1.0 / 17.0; p = p * z + 1.0 / 15.0;
	p = p * z + 1.0 / 13.0; p = p * z + 1.0 / 11.0; p = p * z + 1.0 / 9.0;  p = p * z + 1.0 / 7.0;
	p = p * z + 1.0 / 5.0;  p = p * z + 1.0 / 3.0;  p = p * z + 1.0;
	double lm = 2.0 * s * p;
	return (double) e * 6.93147180369123816490e-01 + ((double) e * 1.90821492927058770002e-10 + lm);
}
__device__ __forceinline__ void det_sincos(double x, double &sn, double &cs) {
	double qf = floor(x * 0.63661977236758134308 + 0.5);
	int q = (int) qf;
	double r = (x - qf * 1.57079632673412561417e+00) - qf * 6.07710050650619224932e-11;
	double z = r * r;
	double ps = -1.0 / 355687428096000.0;
	ps = ps * z + 1.0 / 1307674368000.0; ps = ps * z - 1.0 / 6227020800.0; ps = ps * z + 1.0 / 39916800.0;
	ps = ps * z - 1.0 / 362880.0; ps = ps * z + 1.0 / 5040.0; ps = ps * z - 1.0 / 120.0; ps = ps * z + 1.0 / 6.0;
	double s = r - r * z * ps;
	double pc = 1.0 / 20922789888000.0;
	pc = pc * z - 1.0 / 87178291200.0; pc = pc * z + 1.0 / 479001600.0; pc = pc * z - 1.0 / 3628800.0;
	pc = pc * z + 1.0 / 40320.0; pc = pc * z - 1.0 / 720.0; pc = pc * z + 1.0 / 24.0; pc = pc * z - 0.5;
	double c = 1.0 + z * pc;
	switch (q & 3) {
	case 0: sn = s; cs = c; break;
	case 1: sn = c; cs = -s; break;
	case 2: sn = -s; cs = -c; break;
	default: sn = -c; cs = s; break;
	}
}
__device__ __forceinline__ double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ void box_muller(double ua, double ub, double &z0, double &z1) {
	double r = sqrt(-2.0 * det_log(1.0 - ua)), sn, cs;
	det_sincos(6.283185307179586 * ub, sn, cs);
	z0 = r * cs;
	z1 = r * sn;
}

// ---- terrain lookups (src/fast_terrain_map.cpp)
// Cell search of :101-117 — first i with ax[i] <= v < ax[i+1] — done as an O(1) guess plus a
// compare fix-up against the same fp64 axis values (bisection if the axis is not uniform), so the
// chosen cell is exactly the reference's.  Out of [ax[0], ax[n-1]) -> cell 0 + GBP_FLAG_OOG.
__device__ __forceinline__ int find_cell(const double *__restrict__ ax, int n, double v, double a0, double alast, double inv,
										 unsigned &flags, double &lo_v, double &hi_v) {
	int i = 0;
	if (!(v >= a0) || !(v < alast)) {
		flags |= GBP_FLAG_OOG;
	} else {
		double g = (v - a0) * inv;
		i = (int) g;
		i = max(0, min(i, n - 2));
		double a = __ldg(ax + i), b = __ldg(ax + i + 1);
		if (a <= v && v < b) { lo_v = a; hi_v = b; return i; }
		if (v < a && i > 0 && __ldg(ax + i - 1) <= v) { --i; }
		else if (v >= b && i < n - 2 && v < __ldg(ax + i + 2)) { ++i; }
		else {  // non-uniform axis: bisection on ax[lo] <= v < ax[hi]
			int lo = 0, hi = n - 1;
			while (hi - lo > 1) {
				int mid = (lo + hi) >> 1;
				if (__ldg(ax + mid) <= v) lo = mid; else hi = mid;
			}
			i = lo;
		}
	}
	lo_v = __ldg(ax + i);
	hi_v = __ldg(ax + i + 1);
	return i;
}

template <typename CellT>
__device__ __forceinline__ void load_quad(const CellT *__restrict__ layer, int ny, int ix, int iy, double &f11, double &f12,
										  double &f21, double &f22) {
	const CellT *p = layer + (size_t) ix * ny + iy;
	f11 = (double) __ldg(p);
	f12 = (double) __ldg(p + 1);
	f21 = (double) __ldg(p + ny);
	f22 = (double) __ldg(p + ny + 1);
}
// bilinear form of :120-126, left to right as written
__device__ __forceinline__ double bilinear(double f11, double f12, double f21, double f22, double x1, double x2, double y1,
										   double y2, double x, double y) {
	return 1.0 / ((x2 - x1) * (y2 - y1)) *
		   (f11 * (x2 - x) * (y2 - y) + f21 * (x - x1) * (y2 - y) + f12 * (x2 - x) * (y - y1) + f22 * (x - x1) * (y - y1));
}

template <typename M>
__device__ __forceinline__ double ground_height(const TerrainView &T, double x, double y, unsigned &flags) {  // :94-132
	double x1, x2, y1, y2, f11, f12, f21, f22;
	int ix = find_cell(T.x, T.nx, x, T.x0, T.x_last, T.inv_dx, flags, x1, x2);
	int iy = find_cell(T.y, T.ny, y, T.y0, T.y_last, T.inv_dy, flags, y1, y2);
	load_quad<typename M::cell>((const typename M::cell *) T.z, T.ny, ix, iy, f11, f12, f21, f22);
	return bilinear(f11, f12, f21, f22, x1, x2, y1, y2, x, y);
}
// heightIsNan (:135-157) and getGroundHeight of the SAME point share one cell search and one quad
template <typename M>
__device__ __forceinline__ bool nan_and_height(const TerrainView &T, double x, double y, unsigned &flags, double &h) {
	double x1, x2, y1, y2, f11, f12, f21, f22;
	int ix = find_cell(T.x, T.nx, x, T.x0, T.x_last, T.inv_dx, flags, x1, x2);
	int iy = find_cell(T.y, T.ny, y, T.y0, T.y_last, T.inv_dy, flags, y1, y2);
	load_quad<typename M::cell>((const typename M::cell *) T.z, T.ny, ix, iy, f11, f12, f21, f22);
	h = bilinear(f11, f12, f21, f22, x1, x2, y1, y2, x, y);
	return (f11 != f11) || (f12 != f12) || (f21 != f21) || (f22 != f22);
}
template <typename M>
__device__ __forceinline__ bool height_is_nan(const TerrainView &T, double x, double y, unsigned &flags) {
	double h;
	return nan_and_height<M>(T, x, y, flags, h);
}
__device__ __forceinline__ void surface_normal(const TerrainView &T, double x, double y, double n[3], unsigned &flags) {  // :160-213
	double x1, x2, y1, y2, f11, f12, f21, f22;
	int ix = find_cell(T.x, T.nx, x, T.x0, T.x_last, T.inv_dx, flags, x1, x2);
	int iy = find_cell(T.y, T.ny, y, T.y0, T.y_last, T.inv_dy, flags, y1, y2);
	size_t plane = (size_t) T.nx * T.ny;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		if (T.nz3) load_quad<float>(T.nz3 + k * plane, T.ny, ix, iy, f11, f12, f21, f22);
		else if (T.nz3d) load_quad<double>(T.nz3d + k * plane, T.ny, ix, iy, f11, f12, f21, f22);
		else { f11 = f12 = f21 = f22 = (k == 2) ? 1.0 : 0.0; }
		n[k] = bilinear(f11, f12, f21, f22, x1, x2, y1, y2, x, y);
	}
}

// ---- primitives (src/planning_utils.cpp)
__device__ __forceinline__ void apply_stance(const double s[8], const double a[10], double t, double o[8]) {  // :237-274
	double ts = a[6];
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		o[d] = s[d] + s[3 + d] * t + 0.5 * a[d] * t * t + (a[3 + d] - a[d]) * (t * t * t) / (6.0 * ts);
		o[3 + d] = s[3 + d] + a[d] * t + (a[3 + d] - a[d]) * t * t / (2.0 * ts);
	}
	o[6] = s[6] + s[7] * t + 0.5 * a[8] * t * t + (a[9] - a[8]) * (t * t * t) / (6.0 * ts);
	o[7] = s[7] + a[8] * t + (a[9] - a[8]) * t * t / (2.0 * ts);
}
__device__ __forceinline__ void apply_flight(const double s[8], double t, double o[8]) {  // :282-306 (literal g = 9.81)
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
__device__ __forceinline__ void apply_stance_reverse(const double s[8], const double a[10], double t, double o[8]) {  // :324-367
	double ts = a[6];
#pragma unroll
	for (int d = 0; d < 4; ++d) {
		const int ip = d < 3 ? d : 6, iv = d < 3 ? 3 + d : 7, itd = d < 3 ? d : 8, ito = d < 3 ? 3 + d : 9;
		double c = s[iv] - a[itd] * ts - 0.5 * (a[ito] - a[itd]) * ts;
		o[ip] = s[ip] - c * (ts - t) - 0.5 * a[itd] * (ts * ts - t * t) - (a[ito] - a[itd]) * (ts * ts * ts - t * t * t) / (6.0 * ts);
		o[iv] = s[iv] - a[itd] * (ts - t) - (a[ito] - a[itd]) * (ts * ts - t * t) / (2.0 * ts);
	}
}

// rotate_grf (:198-231): Rodrigues matrix taking +z to n, built once per extend.
__device__ __forceinline__ void grf_rotation(const double n[3], double R[9]) {
	double zs0 = 0.0, zs1 = 0.0, zs2 = 1.0;
	double v0 = n[1] * zs2 - n[2] * zs1, v1 = n[2] * zs0 - n[0] * zs2, v2 = n[0] * zs1 - n[1] * zs0;
	double s = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
	double c = n[0] * zs0 + n[1] * zs1 + n[2] * zs2;
	R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
	if (!(s < 0.000001)) {
		double K[9] = {0, -v2, v1, v2, 0, -v0, -v1, v0, 0}, KK[9];
#pragma unroll
		for (int i = 0; i < 3; ++i)
#pragma unroll
			for (int j = 0; j < 3; ++j) KK[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
#pragma unroll
		for (int i = 0; i < 9; ++i) R[i] = (R[i] + K[i]) + KK[i] * (1 - c) / (s * s);
	}
}
__device__ __forceinline__ void mat3_apply(const double R[9], const double f[3], double o[3]) {
#pragma unroll
	for (int i = 0; i < 3; ++i) o[i] = R[3 * i] * f[0] + R[3 * i + 1] * f[1] + R[3 * i + 2] * f[2];
}

__device__ __forceinline__ bool is_valid_action(const double a[10]) {  // :519-556
	if (a[6] <= 0 || a[7] < 0) return false;
	double m = M_CONST, g = G_CONST, mu = MU;
	double fxd = m * a[0], fyd = m * a[1], fzd = m * (a[2] + g), fxo = m * a[3], fyo = m * a[4], fzo = m * (a[5] + g);
	if (sqrt(fxd * fxd + fyd * fyd + fzd * fzd) >= F_MAX || sqrt(fxo * fxo + fyo * fyo + fzo * fzo) >= F_MAX || fzd < 0 ||
		fzo < 0 || a[8] >= F_MAX || a[9] >= F_MAX)  // sic (:543): pitch accel against F_MAX, no abs
		return false;
	if (sqrt(fxd * fxd + fyd * fyd) >= mu * fzd || sqrt(fxo * fxo + fyo * fyo) >= mu * fzo) return false;
	return true;
}

// =====================================================================================================
// Fast validity path.  Everything computed INSIDE isValidState only feeds threshold comparisons, so it
// needs guard-band accuracy, not bit equality: each clearance / reach comparison records whether its
// margin is below NEAR_MARGIN (1e-11 m, GBP_FLAG_NEAR), and the arithmetic below is accurate to
// ~1e-12 m (fp64 with explicit FMAs, IEEE reciprocal, algebraic yaw, polynomial sin/cos on |pitch|<1).
// A candidate that is not flagged NEAR therefore has the same verdict as the reference's libm-based
// evaluation, by an error bound instead of by instruction order.  Exact (bit-identical) arithmetic is
// kept for everything that is OUTPUT: s_new, t_new, the time grid, distances, tree values.
// Control flow: branch-free over the 10 terrain probes (all loads issued up front, no intra-warp
// divergence); the reference's early-exit order is reproduced afterwards, in integer logic, for the
// verdict, the OOG/NEAR flags and the k / L work counters.
struct Pose6 { double x, y, z, dx, dy, pitch; };  // the components isValidState reads

struct Probe {
	double h;       // bilinear ground height
	bool nan;       // any of the 4 cells NaN
	bool oog;       // outside [x0,x_last) x [y0,y_last)
	bool edge;      // (uniform axes only) within EDGE_MARGIN of a grid line: the cell choice is not certain
};
constexpr double EDGE_MARGIN = 1e-11;  // computed edges / probe positions are within ~1e-12 m of the reference's

__device__ __forceinline__ int find_cell_fast(const double *__restrict__ ax, int n, double v, double a0, double alast, double inv,
											  bool &oog, double &lo_v, double &hi_v) {
	oog = !(v >= a0) || !(v < alast);
	double g = (v - a0) * inv;
	int i = oog ? 0 : (int) g;
	i = max(0, min(i, n - 2));
	double a = __ldg(ax + i), b = __ldg(ax + i + 1);
	if (!oog && !(a <= v && v < b)) {  // rare: rounding of the guess, or a non-uniform axis
		if (v < a && i > 0 && __ldg(ax + i - 1) <= v) --i;
		else if (v >= b && i < n - 2 && v < __ldg(ax + i + 2)) ++i;
		else {
			int lo = 0, hi = n - 1;
			while (hi - lo > 1) {
				int mid = (lo + hi) >> 1;
				if (__ldg(ax + mid) <= v) lo = mid; else hi = mid;
			}
			i = lo;
		}
		a = __ldg(ax + i);
		b = __ldg(ax + i + 1);
	}
	lo_v = a;
	hi_v = b;
	return i;
}

// Uniform axis: cell index and cell-relative coordinate without touching the axis array.
// u = (v - a0)/step - i in [0,1); the edges are within ~1e-13 m of the stored axis values.
__device__ __forceinline__ int cell_uniform(int n, double v, double a0, double alast, double inv, double step, bool &oog, bool &edge,
											 double &u) {
	oog = !(v >= a0) || !(v < alast);
	const double g = oog ? 0.0 : (v - a0) * inv;
	int i = min((int) g, n - 2);
	const double fi = (double) i;
	u = oog ? (v - a0) * inv : g - fi;  // out of grid: extrapolate from cell 0 (defined semantics)
#ifdef GBP_EDGE_FMIN
	const double d = fmin(u, 1.0 - u) * step;
	edge = !oog && d < EDGE_MARGIN;
#else
	const double eps = EDGE_MARGIN * inv;  // guard band in cell units
	edge = !oog && (u < eps || u > 1.0 - eps);
#endif
	return oog ? 0 : i;
}

template <typename M>
__device__ __forceinline__ Probe probe_fast(const TerrainView &T, double x, double y) {
	Probe p;
	bool ox, oy;
	double f11, f12, f21, f22;
	if (M::uniform) {
		bool ex, ey;
		double ux, uy;
		const int ix = cell_uniform(T.nx, x, T.x0, T.x_last, T.inv_dx, T.step_x, ox, ex, ux);
		const int iy = cell_uniform(T.ny, y, T.y0, T.y_last, T.inv_dy, T.step_y, oy, ey, uy);
		load_quad<typename M::cell>((const typename M::cell *) T.z, T.ny, ix, iy, f11, f12, f21, f22);
		p.edge = ex || ey;
		const double lo = __fma_rn(f21 - f11, ux, f11), hi = __fma_rn(f22 - f12, ux, f12);
		p.h = __fma_rn(hi - lo, uy, lo);
	} else {
		double x1, x2, y1, y2;
		const int ix = find_cell_fast(T.x, T.nx, x, T.x0, T.x_last, T.inv_dx, ox, x1, x2);
		const int iy = find_cell_fast(T.y, T.ny, y, T.y0, T.y_last, T.inv_dy, oy, y1, y2);
		load_quad<typename M::cell>((const typename M::cell *) T.z, T.ny, ix, iy, f11, f12, f21, f22);
		p.edge = false;
		const double ax = x2 - x, bx = x - x1, ay = y2 - y, by = y - y1;
		const double w = __drcp_rn((x2 - x1) * (y2 - y1));
		const double lo = __fma_rn(f21, bx, f11 * ax), hi = __fma_rn(f22, bx, f12 * ax);
		p.h = w * __fma_rn(hi, by, lo * ay);
	}
	p.oog = ox || oy;
	p.nan = (f11 != f11) || (f12 != f12) || (f21 != f21) || (f22 != f22);
	return p;
}

// sin / cos on |x| < 1 (pitch is rejected at |pitch| >= P_MAX = 1 before it is used): Taylor to x^19 / x^18
__device__ __forceinline__ void sincos_small(double x, double &sn, double &cs) {
	const double z = x * x;
	double ps = -1.0 / 121645100408832000.0;  // -1/19!
	ps = __fma_rn(ps, z, 1.0 / 355687428096000.0);
	ps = __fma_rn(ps, z, -1.0 / 1307674368000.0);
	ps = __fma_rn(ps, z, 1.0 / 6227020800.0);
	ps = __fma_rn(ps, z, -1.0 / 39916800.0);
	ps = __fma_rn(ps, z, 1.0 / 362880.0);
	ps = __fma_rn(ps, z, -1.0 / 5040.0);
	ps = __fma_rn(ps, z, 1.0 / 120.0);
	ps = __fma_rn(ps, z, -1.0 / 6.0);
	sn = __fma_rn(x * z, ps, x);
	double pc = -1.0 / 6402373705728000.0;  // -1/18!
	pc = __fma_rn(pc, z, 1.0 / 20922789888000.0);
	pc = __fma_rn(pc, z, -1.0 / 87178291200.0);
	pc = __fma_rn(pc, z, 1.0 / 479001600.0);
	pc = __fma_rn(pc, z, -1.0 / 3628800.0);
	pc = __fma_rn(pc, z, 1.0 / 40320.0);
	pc = __fma_rn(pc, z, -1.0 / 720.0);
	pc = __fma_rn(pc, z, 1.0 / 24.0);
	pc = __fma_rn(pc, z, -0.5);
	cs = __fma_rn(z, pc, 1.0);
}
// cos / sin of yaw = atan2(dy, dx) without libm: the unit vector (dx, dy) / r
__device__ __forceinline__ void yaw_cs(double dx, double dy, double r, double &cy, double &sy) {
	if (r > 1e-150) {
		const double inv = 1.0 / r;
		cy = dx * inv;
		sy = dy * inv;
	} else if (dx == 0.0 && dy == 0.0) {  // atan2(+-0, +0) = +-0, atan2(+-0, -0) = +-pi
		const bool neg = __double_as_longlong(dx) < 0;
		cy = neg ? -1.0 : 1.0;
		sy = neg ? copysign(1.2246467991473532e-16, dy) : dy;
	} else {  // r underflowed: rescale
		const double m = fmax(fabs(dx), fabs(dy)), ux = dx / m, uy = dy / m, inv = 1.0 / sqrt(ux * ux + uy * uy);
		cy = ux * inv;
		sy = uy * inv;
	}
}

// Verdict bookkeeping shared by the fp64 and the mixed-precision evaluators: replays the reference's check order
// (:564-634) on precomputed per-probe results as straight-line predicate arithmetic (no branches: the compiler
// must not sink the probe loads behind early exits, and the warp must not diverge).
// Probe order: 0 centre, 1+2k leg k, 2+2k corner k (k = 0..3 in the reference's loop order), 9 belly.
struct ProbeBits {
	unsigned nan, oog, edge;  // bit p = probe p
};
__device__ __forceinline__ bool replay_checks(const ProbeBits &pb, bool pre_bad, bool speed_bad, bool stance, const double leg_m[4],
											  const double cor_m[4], double belly_m, double near_margin, Counters &c) {
	// leg_m = leg_height - H_MAX, cor_m = corner_height - H_MIN, belly_m = height - H_MIN
	unsigned flags = ((pb.oog & 1u) ? GBP_FLAG_OOG : 0u) | ((pb.edge & 1u) ? GBP_FLAG_NEAR : 0u), nanprobes = 1, lookups = 0;
	bool alive = !((pb.nan & 1u) || pre_bad || speed_bad);
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const unsigned lb = 1u << (1 + 2 * k), cb = 1u << (2 + 2 * k);
		nanprobes += alive ? 1u : 0u;
		flags |= (alive && (pb.oog & lb)) ? GBP_FLAG_OOG : 0u;
		flags |= (alive && (pb.edge & lb)) ? GBP_FLAG_NEAR : 0u;
		const bool reached = alive && !(pb.nan & lb);
		lookups += reached ? 2u : 0u;
		flags |= (reached && (pb.oog & cb)) ? GBP_FLAG_OOG : 0u;
		flags |= (reached && (pb.edge & cb)) ? GBP_FLAG_NEAR : 0u;
		const bool near = (fabs(cor_m[k]) < near_margin) || (stance && fabs(leg_m[k]) < near_margin);
		flags |= (reached && near) ? GBP_FLAG_NEAR : 0u;
		const bool bad = (cor_m[k] < 0.0) || (stance && leg_m[k] > 0.0);
		alive = reached && !bad;
	}
	lookups += alive ? 1u : 0u;
	flags |= (alive && (pb.oog & (1u << 9))) ? GBP_FLAG_OOG : 0u;
	flags |= (alive && (pb.edge & (1u << 9))) ? GBP_FLAG_NEAR : 0u;
	flags |= (alive && fabs(belly_m) < near_margin) ? GBP_FLAG_NEAR : 0u;
	c.substates += 1;
	c.nanprobes += nanprobes;
	c.lookups += lookups;
	c.flags |= flags;
	return alive && !(belly_m < 0.0);
}

template <typename M>
__device__ __forceinline__ bool is_valid_state_fast(const TerrainView &T, const Pose6 &s, int phase, Counters &c) {
	const bool pre_bad = (s.x < T.x0) || (s.x > T.x_last) || (s.y < T.y0) || (s.y > T.y_last) || (fabs(s.pitch) >= P_MAX);
	const double r = sqrt(s.dx * s.dx + s.dy * s.dy);  // the speed test is a hard comparison (:572)
	const bool speed_bad = r > V_MAX;
	// The poses checked here come from FMA forms of applyStance / applyFlight (rounding-level differences from the reference's
	// expressions, ~1e-15): the two hard comparisons get the same guard-band treatment as the clearances — a speed or pitch
	// within HARD_MARGIN of its threshold is flagged, not silently decided
	c.flags |= (fabs(r - V_MAX) < HARD_MARGIN || fabs(fabs(s.pitch) - P_MAX) < HARD_MARGIN) ? GBP_FLAG_NEAR : 0u;
	double cy, sy, sp, cp;
	yaw_cs(s.dx, s.dy, r, cy, sy);
	sincos_small(pre_bad ? 0.0 : s.pitch, sp, cp);
	const double R11 = cy * cp, R12 = -sy, R13 = cy * sp, R21 = sy * cp, R22 = cy, R23 = sy * sp, R31 = -sp, R33 = cp;
	const double zb = -ROBOT_H;
	// ---- stage A: the 10 probe points
	double px[10], py[10], zl[4], zc[4];
	px[0] = s.x; py[0] = s.y;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const double xb = (k & 2) ? 0.5 * ROBOT_L : -0.5 * ROBOT_L, yb = (k & 1) ? 0.5 * ROBOT_W : -0.5 * ROBOT_W;
		px[1 + 2 * k] = __fma_rn(R12, yb, __fma_rn(R11, xb, s.x));
		py[1 + 2 * k] = __fma_rn(R22, yb, __fma_rn(R21, xb, s.y));
		px[2 + 2 * k] = __fma_rn(R13, zb, px[1 + 2 * k]);
		py[2 + 2 * k] = __fma_rn(R23, zb, py[1 + 2 * k]);
		zl[k] = __fma_rn(R31, xb, s.z);
		zc[k] = __fma_rn(R33, zb, zl[k]);
	}
	px[9] = __fma_rn(R13, zb, s.x); py[9] = __fma_rn(R23, zb, s.y);
	// ---- stage B/C/D: cells, loads, heights
	ProbeBits pb = {0u, 0u, 0u};
	double h[10];
	if (M::uniform) {
		int cell[10];
		double ux[10], uy[10];
#pragma unroll
		for (int p = 0; p < 10; ++p) {
			bool ox, oy, ex, ey;
			const int ix = cell_uniform(T.nx, px[p], T.x0, T.x_last, T.inv_dx, T.step_x, ox, ex, ux[p]);
			const int iy = cell_uniform(T.ny, py[p], T.y0, T.y_last, T.inv_dy, T.step_y, oy, ey, uy[p]);
			cell[p] = ix * T.ny + iy;
			pb.oog |= (ox || oy) ? (1u << p) : 0u;
			pb.edge |= (ex || ey) ? (1u << p) : 0u;
		}
		typename M::cell f[10][4];
#pragma unroll
		for (int p = 0; p < 10; ++p) {  // all 40 loads are issued before any is consumed
			const typename M::cell *q = (const typename M::cell *) T.z + cell[p];
			f[p][0] = __ldg(q); f[p][1] = __ldg(q + 1); f[p][2] = __ldg(q + T.ny); f[p][3] = __ldg(q + T.ny + 1);
		}
#pragma unroll
		for (int p = 0; p < 10; ++p) {
			const double f11 = (double) f[p][0], f12 = (double) f[p][1], f21 = (double) f[p][2], f22 = (double) f[p][3];
			pb.nan |= ((f11 != f11) || (f12 != f12) || (f21 != f21) || (f22 != f22)) ? (1u << p) : 0u;
			const double lo = __fma_rn(f21 - f11, ux[p], f11), hi = __fma_rn(f22 - f12, ux[p], f12);
			h[p] = __fma_rn(hi - lo, uy[p], lo);
		}
	} else {
#pragma unroll
		for (int p = 0; p < 10; ++p) {
			const Probe q = probe_fast<M>(T, px[p], py[p]);
			h[p] = q.h;
			pb.nan |= q.nan ? (1u << p) : 0u;
			pb.oog |= q.oog ? (1u << p) : 0u;
		}
	}
	double leg_m[4], cor_m[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		leg_m[k] = (zl[k] - h[1 + 2 * k]) - H_MAX;
		cor_m[k] = (zc[k] - h[2 + 2 * k]) - H_MIN;
	}
	const double belly_m = (__fma_rn(R33, zb, s.z) - h[9]) - H_MIN;
	return replay_checks(pb, pre_bad, speed_bad, phase == GBP_STANCE, leg_m, cor_m, belly_m, NEAR_MARGIN, c);
}
// =====================================================================================================
// Mixed-precision evaluator (uniform axes; fp32 cells, or fp64 cells read through their fp32-rounded texture copy: that
// rounding, <= 2.4e-7 m for |z| <= 4 m, is part of the error budget below).  Same contract as is_valid_state_fast, one more
// level of the same idea: the centre cell and its in-cell fraction are found in fp64, everything that is a
// SMALL offset from it (body rotation, the 9 leg / corner / belly offsets in cell units, the bilinear
// increment over the cell's first corner) is fp32 on the full-rate FMA pipe, and each clearance margin is
// assembled in fp64 as (z - f11) - H + (offset - increment).  Error budget < 1e-6 m; any sub-state with a
// margin below MIXED_MARGIN (1e-5 m) or a probe
// outside the grid is NOT decided here: it is re-evaluated by is_valid_state_fast (fp64, 1e-11 m guard).
constexpr double MIXED_MARGIN = 1e-5;
// No guard is needed around grid lines here: the bilinear surface is continuous across a cell edge, so a probe that the fp32
// cell arithmetic puts in the neighbour of the reference's cell gets the height of the same surface extended by < 1e-6 cells
// (< 1e-8 m for any slope the |dz| <= 4 m per cell bound admits) — far inside MIXED_MARGIN; and these maps hold no NaN cell,
// the one thing a cell choice could change.

__device__ __forceinline__ void sincosf_small(float x, float &sn, float &cs) {  // |x| < 1, abs error < 3e-8
	const float z = x * x;
	float ps = 1.0f / 362880.0f;
	ps = fmaf(ps, z, -1.0f / 5040.0f);
	ps = fmaf(ps, z, 1.0f / 120.0f);
	ps = fmaf(ps, z, -1.0f / 6.0f);
	sn = fmaf(x * z, ps, x);
	float pc = -1.0f / 3628800.0f;
	pc = fmaf(pc, z, 1.0f / 40320.0f);
	pc = fmaf(pc, z, -1.0f / 720.0f);
	pc = fmaf(pc, z, 1.0f / 24.0f);
	pc = fmaf(pc, z, -0.5f);
	cs = fmaf(z, pc, 1.0f);
}

// Returns true when the sub-state was decided (verdict in `valid`, counters updated); false = needs the fp64 path.
// Preconditions checked here: the centre lies at least T.border cells inside the grid (so no probe can leave it:
// no OOG flags, no index clamping) and the map holds no NaN (so heightIsNan never fires; T.mixed_ok).
// TEX = true: the 4 cells of a probe come from one texture gather on T.ztex (array x = iy, array y = ix; sampling at the
// centre of the 2x2 footprint, (iy + 1, ix + 1), is exact in the unit's fixed-point coordinates; components w, z, x, y
// = cells (ix,iy), (ix,iy+1), (ix+1,iy), (ix+1,iy+1) — raw fp32 values, so every result is identical to the LDG form).
template <typename M, bool TEX = false>
__device__ __forceinline__ bool is_valid_state_mixed(const TerrainView &T, const Pose6 &s, int phase, Counters &c, bool &valid) {
	const bool pitch_bad = fabs(s.pitch) >= P_MAX;
	// sqrt(r2) > V_MAX  <=>  r2 > nextafter(V_MAX^2): the largest double whose correctly rounded root is still 2.0
	const double r2 = s.dx * s.dx + s.dy * s.dy;
	const bool speed_bad = r2 > __longlong_as_double(0x4010000000000001ll);
	static_assert(V_MAX == 2.0, "speed threshold constant is derived for V_MAX = 2");
	// speed or pitch within 1e-9 of its threshold (poses here come from the polynomial cursor, ~1e-14 from the reference's
	// expressions): not decided here, the fp64 evaluator flags it if it is closer than HARD_MARGIN
	const bool hard_near = fabs(r2 - V_MAX * V_MAX) < 4e-9 || fabs(fabs(s.pitch) - P_MAX) < 1e-9;
	// centre cell in fp64, in-cell fraction handed to fp32
	const double gx = (s.x - T.x0) * T.inv_dx, gy = (s.y - T.y0) * T.inv_dy;
	const int ixc = (int) gx, iyc = (int) gy;
	bool ok = (gx >= (double) T.border) && (gy >= (double) T.border) && (ixc <= T.nx - 2 - T.border) && (iyc <= T.ny - 2 - T.border);
	ok = ok && (s.z == s.z) && (pitch_bad || fabs(s.pitch) < P_MAX);  // a NaN pose is not decided here (NaN x, y, dx, dy fail the tests above / below)
	const float fux = (float) (gx - (double) ixc), fuy = (float) (gy - (double) iyc);
	// body orientation in fp32
	const float fdx = (float) s.dx, fdy = (float) s.dy, fr2 = fdx * fdx + fdy * fdy;
	float cy, sy;
	if (r2 == 0.0) {  // atan2(+-0, +0) = +-0, atan2(+-0, -0) = +-pi
		cy = __double_as_longlong(s.dx) < 0 ? -1.0f : 1.0f;
		sy = 0.0f;
	} else {
		ok = ok && (fr2 > 1e-30f) && (fr2 < 1e30f);
		const float rinv = rsqrtf(fr2);
		cy = fdx * rinv;
		sy = fdy * rinv;
	}
	float sp, cp;
	sincosf_small(pitch_bad ? 0.0f : (float) s.pitch, sp, cp);
	const float R11 = cy * cp, R12 = -sy, R13 = cy * sp, R21 = sy * cp, R22 = cy, R23 = sy * sp, R31 = -sp, R33 = cp;
	const float kx = (float) T.inv_dx, ky = (float) T.inv_dy, zb = -(float) ROBOT_H;
	// offsets of the 9 probes from the centre: cell units in x/y, metres in z.  p = 2k leg k, 2k+1 corner k, 8 belly
	float ox[9], oy[9], oz[9];
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const float xb = (k & 2) ? 0.5f * (float) ROBOT_L : -0.5f * (float) ROBOT_L, yb = (k & 1) ? 0.5f * (float) ROBOT_W : -0.5f * (float) ROBOT_W;
		const float lx = fmaf(R12, yb, R11 * xb), ly = fmaf(R22, yb, R21 * xb), lz = R31 * xb;
		ox[2 * k] = lx * kx; oy[2 * k] = ly * ky; oz[2 * k] = lz;
		ox[2 * k + 1] = fmaf(R13, zb, lx) * kx; oy[2 * k + 1] = fmaf(R23, zb, ly) * ky; oz[2 * k + 1] = fmaf(R33, zb, lz);
	}
	ox[8] = R13 * zb * kx; oy[8] = R23 * zb * ky; oz[8] = R33 * zb;
	// cells (no clamping needed: |offset| < T.border cells) and the distance of every probe to its nearest grid line
	const int base = ixc * T.ny + iyc;
	const float tbx = (float) (iyc + 1), tby = (float) (ixc + 1);  // texel-footprint centre of the centre cell (exact: < 2^24)
	int cell[9];
	float tx[9], ty[9];
	float ux[9], uy[9];
#pragma unroll
	for (int p = 0; p < 9; ++p) {
		const float pxf = fux + ox[p], pyf = fuy + oy[p], flx = floorf(pxf), fly = floorf(pyf);
		ux[p] = pxf - flx; uy[p] = pyf - fly;
		if (TEX) { tx[p] = tbx + fly; ty[p] = tby + flx; }  // clamp addressing: a lane that is not `ok` reads some cell and is discarded below
		else cell[p] = ok ? base + (int) flx * T.ny + (int) fly : 0;
	}
	// all 36 cells (36 loads or 9 gathers), then the heights as increments over the first corner
	float f[9][4];
#pragma unroll
	for (int p = 0; p < 9; ++p) {
		if (TEX) {
#ifndef GBP_NO_FLIGHT_SKIP
			// the 4 leg probes (p = 0, 2, 4, 6) only feed the STANCE reach test (:614-617): a lane checking a FLIGHT sub-state
			// does not fetch them (no NaN cell and no out-of-grid probe exists on this path, so nothing else depends on them)
			if (p < 8 && (p & 1) == 0 && phase != GBP_STANCE) { f[p][0] = f[p][1] = f[p][2] = f[p][3] = 0.0f; continue; }
#endif
			const float4 g = tex2Dgather<float4>((cudaTextureObject_t) T.ztex, tx[p], ty[p], 0);
			f[p][0] = g.w; f[p][1] = g.z; f[p][2] = g.x; f[p][3] = g.y;
		} else {
			const float *q = (const float *) T.z + cell[p];
			f[p][0] = __ldg(q); f[p][1] = __ldg(q + 1); f[p][2] = __ldg(q + T.ny); f[p][3] = __ldg(q + T.ny + 1);
		}
	}
	double m[9];
#pragma unroll
	for (int p = 0; p < 9; ++p) {
		const float f11 = f[p][0], f12 = f[p][1], f21 = f[p][2], f22 = f[p][3];
		const float lo = (f21 - f11) * ux[p], hi = fmaf(f22 - f12, ux[p], f12 - f11);
		const float inc = fmaf(hi - lo, uy[p], lo);  // ground(p) - f11
		const double H = (p & 1) ? H_MIN : (p == 8 ? H_MIN : H_MAX);
		m[p] = ((s.z - (double) f11) - H) + (double) (oz[p] - inc);  // clearance margin: leg - H_MAX, corner / belly - H_MIN
	}
	// the reference's check order (:564-634), with no NaN and no out-of-grid probe possible here
	const bool stance = phase == GBP_STANCE;
	bool near = false;
	unsigned bad = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		near = near || (fabs(m[2 * k + 1]) < MIXED_MARGIN) || (stance && fabs(m[2 * k]) < MIXED_MARGIN);
		bad |= ((m[2 * k + 1] < 0.0) || (stance && m[2 * k] > 0.0)) ? (1u << k) : 0u;
	}
	near = near || (fabs(m[8]) < MIXED_MARGIN);
	if (!ok || near || hard_near) return false;
	const bool alive0 = !(pitch_bad || speed_bad);
	const int corners = alive0 ? min(__ffs(bad | 16u), 4) : 0;  // corners the reference evaluates before it returns
	const bool all_ok = alive0 && bad == 0;
	c.substates += 1;
	c.nanprobes += 1 + corners;
	c.lookups += 2 * corners + (all_ok ? 1 : 0);
	valid = all_ok && !(m[8] < 0.0);
	return true;
}

// the evaluator every kernel calls
template <typename M>
__device__ __forceinline__ bool is_valid_state_auto(const TerrainView &T, const Pose6 &s, int phase, Counters &c) {
	if (M::uniform) {
		bool valid;
		// texture-gather fetch when the handle carries the block-linear fp32 copy (ztex implies mixed_ok; for fp64 maps it is
		// the only mixed form: the copy holds the heights rounded to fp32): +13 % plans/s in k_plan_batch
		if (T.ztex) { if (is_valid_state_mixed<M, true>(T, s, phase, c, valid)) return valid; }
		else if (sizeof(typename M::cell) == 4 && T.mixed_ok && is_valid_state_mixed<M>(T, s, phase, c, valid)) return valid;
	}
	return is_valid_state_fast<M>(T, s, phase, c);
}

__device__ __forceinline__ Pose6 pose6(const double s[8]) {
	Pose6 p;
	p.x = s[0]; p.y = s[1]; p.z = s[2]; p.dx = s[3]; p.dy = s[4]; p.pitch = s[6];
	return p;
}

// Fast propagation of the 6 components isValidState reads.  c3* = (a_to - a_td) / (6 ts), precomputed per candidate.
struct FastPrim {
	double inv6ts, inv2ts;
};
__device__ __forceinline__ Pose6 stance_fast(const double s[8], const double a[10], const FastPrim &f, double t) {
	Pose6 o;
	const double t2 = t * t;
	const double jx = (a[3] - a[0]), jy = (a[4] - a[1]), jz = (a[5] - a[2]), jp = (a[9] - a[8]);
	o.x = __fma_rn(jx * f.inv6ts, t2 * t, __fma_rn(0.5 * a[0], t2, __fma_rn(s[3], t, s[0])));
	o.y = __fma_rn(jy * f.inv6ts, t2 * t, __fma_rn(0.5 * a[1], t2, __fma_rn(s[4], t, s[1])));
	o.z = __fma_rn(jz * f.inv6ts, t2 * t, __fma_rn(0.5 * a[2], t2, __fma_rn(s[5], t, s[2])));
	o.pitch = __fma_rn(jp * f.inv6ts, t2 * t, __fma_rn(0.5 * a[8], t2, __fma_rn(s[7], t, s[6])));
	o.dx = __fma_rn(jx * f.inv2ts, t2, __fma_rn(a[0], t, s[3]));
	o.dy = __fma_rn(jy * f.inv2ts, t2, __fma_rn(a[1], t, s[4]));
	return o;
}
// full 8-component fast stance end state (needed as the take-off state of the flight phase)
__device__ __forceinline__ void stance_fast8(const double s[8], const double a[10], const FastPrim &f, double t, double o[8]) {
	const double t2 = t * t;
#pragma unroll
	for (int d = 0; d < 4; ++d) {
		const int ip = d < 3 ? d : 6, iv = d < 3 ? 3 + d : 7, itd = d < 3 ? d : 8, ito = d < 3 ? 3 + d : 9;
		const double j = a[ito] - a[itd];
		o[ip] = __fma_rn(j * f.inv6ts, t2 * t, __fma_rn(0.5 * a[itd], t2, __fma_rn(s[iv], t, s[ip])));
		o[iv] = __fma_rn(j * f.inv2ts, t2, __fma_rn(a[itd], t, s[iv]));
	}
}
__device__ __forceinline__ Pose6 flight_fast(const double s[8], double t) {
	Pose6 o;
	o.x = __fma_rn(s[3], t, s[0]);
	o.y = __fma_rn(s[4], t, s[1]);
	o.z = __fma_rn(-0.5 * 9.81 * t, t, __fma_rn(s[5], t, s[2]));
	o.dx = s[3];
	o.dy = s[4];
	o.pitch = __fma_rn(s[7], t, s[6]);
	return o;
}
// applyStanceReverse (:324-367) restricted to the 6 components, from the take-off state `s`
__device__ __forceinline__ Pose6 stance_reverse_fast(const double s[8], const double a[10], const FastPrim &f, double t) {
	Pose6 o;
	const double ts = a[6], dt = ts - t, d2 = ts * ts - t * t, d3 = ts * ts * ts - t * t * t;
	double p[4], v[2];
#pragma unroll
	for (int d = 0; d < 4; ++d) {
		const int ip = d < 3 ? d : 6, iv = d < 3 ? 3 + d : 7, itd = d < 3 ? d : 8, ito = d < 3 ? 3 + d : 9;
		const double j = a[ito] - a[itd];
		const double cc = s[iv] - a[itd] * ts - 0.5 * j * ts;
		p[d] = s[ip] - cc * dt - 0.5 * a[itd] * d2 - j * d3 * f.inv6ts;
		if (d < 2) v[d] = s[iv] - a[itd] * dt - j * d2 * f.inv2ts;
	}
	o.x = p[0]; o.y = p[1]; o.z = p[2]; o.pitch = p[3]; o.dx = v[0]; o.dy = v[1];
	return o;
}

// ---- distances (:106-127, planning_utils.h:133-145)
__device__ __forceinline__ double pose_distance(const double a[8], const double b[8]) {
	double sum = 0;
#pragma unroll
	for (int i = 0; i < 3; ++i) sum = sum + (b[i] - a[i]) * (b[i] - a[i]);
	return sqrt(sum);
}
__device__ __forceinline__ double state_distance(const double a[8], const double b[8]) {
	double sum = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) sum = sum + 1.0 * (b[i] - a[i]) * (b[i] - a[i]);
	return sqrt(sum);
}
__device__ __forceinline__ double yaw_distance(const double a[8], const double b[8]) {
	double y1 = atan2(a[4], a[3]), y2 = atan2(b[4], b[3]);
	double lo = y1 < y2 ? y1 : y2, hi = y1 < y2 ? y2 : y1;
	double d1 = hi - lo, d2 = lo + 2 * MY_PI - hi;
	return d2 < d1 ? d2 : d1;
}

// ---- samplers on the Philox stream (draw layout documented in oracle/gbp_oracle.c and DESIGN.md)
// getRandomAction (:392-442) / getRandomActionDirection (:443-515); R = grf_rotation(normal)
__device__ __forceinline__ void sample_action(uint64_t seed, uint64_t stream, uint64_t idx, const double R[9], bool dir_flag,
											  double dir_thresh, const double *s_from, const double *s_to, double a[10]) {
	double u0, u1, u2, u3, u4, u5, u6, u7, u8, u9;
	uniform_pair(seed, stream, idx, 1, 0, u0, u1);
	uniform_pair(seed, stream, idx, 1, 1, u2, u3);
	uniform_pair(seed, stream, idx, 1, 2, u4, u5);
	uniform_pair(seed, stream, idx, 1, 3, u6, u7);
	uniform_pair(seed, stream, idx, 1, 4, u8, u9);
	double fzd = F_MAX * u0, fzo = F_MAX * u1, fxd, fxo, fyd, fyo;
	if (dir_flag && u9 <= dir_thresh) {
		double frd = MU * fzd, fro = MU * fzo;
		if (s_to[3] > s_from[3]) { fxd = frd * u2; fxo = fro * u3; } else { fxd = frd * u2 - frd; fxo = fro * u3 - fro; }
		if (s_to[4] > s_from[4]) { fyd = frd * u4; fyo = fro * u5; } else { fyd = frd * u4 - frd; fyo = fro * u5 - fro; }
	} else {
		fxd = 2 * MU * fzd * u2 - MU * fzd;
		fxo = 2 * MU * fzo * u3 - MU * fzo;
		fyd = 2 * MU * fzd * u4 - MU * fzd;
		fyo = 2 * MU * fzo * u5 - MU * fzo;
	}
	double ftd[3] = {fxd, fyd, fzd}, fto[3] = {fxo, fyo, fzo}, rtd[3], rto[3];
	mat3_apply(R, ftd, rtd);
	mat3_apply(R, fto, rto);
	a[0] = rtd[0] / M_CONST;
	a[1] = rtd[1] / M_CONST;
	a[2] = rtd[2] / M_CONST - G_CONST;
	a[3] = rto[0] / M_CONST;
	a[4] = rto[1] / M_CONST;
	a[5] = rto[2] / M_CONST - G_CONST;
	a[6] = 0.3;
	a[7] = (T_F_MAX - T_F_MIN) * u6 + T_F_MIN;
	double z0, z1;
	const double sd = ANG_ACC_MAX / 4.0;
	box_muller(u7, u8, z0, z1);
	a[8] = clampd(sd * z0, -ANG_ACC_MAX, ANG_ACC_MAX);
	a[9] = clampd(sd * z1, -ANG_ACC_MAX, ANG_ACC_MAX);
}
// PlannerClass::randomState (planner_class.cpp:38-76) / randomStateDirection (:82-148)
template <typename M>
__device__ __forceinline__ void sample_state(const TerrainView &T, uint64_t seed, uint64_t stream, uint64_t idx, bool dir_flag,
											 double dir_thresh, bool speed_dir, const double *s_from, const double *s_to,
											 double q[8]) {
	double u0, u1, u2, u3, u4, u5, u6, u7, u8, u9;
	uniform_pair(seed, stream, idx, 2, 0, u0, u1);
	uniform_pair(seed, stream, idx, 2, 1, u2, u3);
	uniform_pair(seed, stream, idx, 2, 2, u4, u5);
	uniform_pair(seed, stream, idx, 2, 3, u6, u7);
	bool directional = false;
	if (dir_flag) { uniform_pair(seed, stream, idx, 2, 4, u8, u9); directional = u8 <= dir_thresh; }
	double x_min = T.x0, x_max = T.x_last, y_min = T.y0, y_max = T.y_last;
	if (directional) {
		x_min = s_from[0] < s_to[0] ? s_from[0] : s_to[0];
		x_max = s_from[0] < s_to[0] ? s_to[0] : s_from[0];
		y_min = s_from[1] < s_to[1] ? s_from[1] : s_to[1];
		y_max = s_from[1] < s_to[1] ? s_to[1] : s_from[1];
	}
	const double z_min_rel = H_MIN + ROBOT_H, z_max_rel = H_MAX + ROBOT_H;
	const double mean = 0.5 * (z_max_rel + z_min_rel), sd = (z_max_rel - z_min_rel) * (1.0 / (2 * 3.0));
	double z0, z1, sn, cs;
	box_muller(u2, u3, z0, z1);
	unsigned fl = 0;
	q[0] = (x_max - x_min) * u0 + x_min;
	q[1] = (y_max - y_min) * u1 + y_min;
	q[2] = clampd(mean + sd * z0, z_min_rel, z_max_rel) + ground_height<M>(T, q[0], q[1], fl);
	double cos_theta = 2.0 * u5 - 1.0, sin_theta = sqrt(1.0 - cos_theta * cos_theta), v = u6 * V_MAX;
	if (directional && speed_dir) {
		double ddx = s_to[0] - s_from[0], ddy = s_to[1] - s_from[1], nrm = sqrt(ddx * ddx + ddy * ddy);
		if (nrm > 0) { cs = ddx / nrm; sn = ddy / nrm; } else { cs = 1.0; sn = 0.0; }
	} else {
		det_sincos((2.0 * MY_PI) * u4, sn, cs);
	}
	q[3] = v * sin_theta * cs;
	q[4] = v * sin_theta * sn;
	q[5] = v * cos_theta;
	q[6] = 2 * P_MAX * u7 - P_MAX;
	q[7] = 0.0;
}

// ---- attemptConnect (src/rrt_connect.cpp:20-91), recursion unrolled into a loop (see oracle)
__device__ __forceinline__ void connect_action(const double st[8], const double go[8], double ts, double a[10]) {  // :53-63
#pragma unroll
	for (int d = 0; d < 4; ++d) {
		const int ip = d < 3 ? d : 6, iv = d < 3 ? 3 + d : 7, itd = d < 3 ? d : 8, ito = d < 3 ? 3 + d : 9;
		a[itd] = -(2.0 * (3.0 * st[ip] - 3.0 * go[ip] + 2.0 * st[iv] * ts + go[iv] * ts)) / (ts * ts);
		a[ito] = (2.0 * (3.0 * st[ip] - 3.0 * go[ip] + st[iv] * ts + 2.0 * go[iv] * ts)) / (ts * ts);
	}
	a[6] = ts;
	a[7] = 0;
}
}  // namespace gbp
