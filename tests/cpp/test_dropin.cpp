// Exercises the drop-in C++ API (include/global_body_planner/*.h) the way a caller of the reference
// would: same class names, same calls.  Prints "key v0 v1 ..." lines (%.17g) that
// tests/test_gpu_dropin.py compares with the oracle.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include <global_body_planner/rrt_star_connect.h>

using namespace planning_utils;

static std::vector<double> read_doubles(std::ifstream &f, size_t n) {
	std::vector<double> v(n);
	f.read((char *) v.data(), (std::streamsize) (n * sizeof(double)));
	return v;
}
static void line(const char *key, const double *v, int n) {
	std::printf("%s", key);
	for (int i = 0; i < n; ++i) std::printf(" %.17g", v[i]);
	std::printf("\n");
}

int main(int argc, char **argv) {
	std::setvbuf(stdout, nullptr, _IOLBF, 0);
	if (argc < 3) { std::fprintf(stderr, "usage: test_dropin terrain.bin cases.bin\n"); return 2; }
	std::ifstream ft(argv[1], std::ios::binary), fc(argv[2], std::ios::binary);
	double dims[2];
	ft.read((char *) dims, sizeof dims);
	const int nx = (int) dims[0], ny = (int) dims[1];
	std::vector<double> x = read_doubles(ft, nx), y = read_doubles(ft, ny);
	std::vector<std::vector<double>> L[4];
	for (int k = 0; k < 4; ++k) {
		std::vector<double> flat = read_doubles(ft, (size_t) nx * ny);
		L[k].assign(nx, std::vector<double>(ny));
		for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) L[k][i][j] = flat[(size_t) i * ny + j];
	}
	FastTerrainMap terrain;
	terrain.loadData(nx, ny, x, y, L[0], L[1], L[2], L[3]);
	FastTerrainMap terrain_copy = terrain;  // the reference copies terrains by value

	double nd;
	fc.read((char *) &nd, sizeof nd);
	const int n = (int) nd;
	std::vector<double> S = read_doubles(fc, (size_t) 8 * n), A = read_doubles(fc, (size_t) 10 * n), D = read_doubles(fc, n);
	std::vector<State> states(n);
	std::vector<Action> actions(n);
	for (int i = 0; i < n; ++i) {
		for (int d = 0; d < 8; ++d) states[i][d] = S[8 * i + d];
		for (int d = 0; d < 10; ++d) actions[i][d] = A[10 * i + d];
	}
	// --- scalar API, element by element
	for (int i = 0; i < n && i < 40; ++i) {
		double v[16];
		v[0] = terrain_copy.getGroundHeight(states[i][0], states[i][1]);
		std::array<double, 3> nn = terrain.getSurfaceNormal(states[i][0], states[i][1]);
		v[1] = nn[0]; v[2] = nn[1]; v[3] = nn[2];
		v[4] = terrain.heightIsNan(states[i][0], states[i][1]);
		v[5] = isValidState(states[i], terrain, STANCE);
		v[6] = isValidState(states[i], terrain, FLIGHT);
		v[7] = isValidAction(actions[i]);
		v[8] = poseDistance(states[i], states[(i + 1) % n]);
		v[9] = stateDistance(states[i], states[(i + 1) % n]);
		line("scalar", v, 10);
		State sn; double tn = 0;
		bool ok = D[i] == 0 ? isValidStateActionPair(states[i], actions[i], terrain, sn, tn, false)
							: isValidStateActionPairReverse(states[i], actions[i], terrain, sn, tn, false);
		double w[10] = {(double) ok, tn};
		for (int d = 0; d < 8; ++d) w[2 + d] = sn[d];
		line("pair", w, 10);
		State st = applyStance(states[i], actions[i], 0.1), fl = applyFlight(states[i], 0.2), rv = applyStanceReverse(states[i], actions[i], 0.1);
		line("stance", st.data(), 8);
		line("flight", fl.data(), 8);
		line("stancerev", rv.data(), 8);
	}
	// --- batched pair check
	{
		std::vector<unsigned char> dir(n);
		for (int i = 0; i < n; ++i) dir[i] = (unsigned char) D[i];
		std::vector<State> sn;
		std::vector<double> tn;
		std::vector<unsigned char> v = isValidStateActionPair(states, actions, dir, terrain, sn, tn);
		std::vector<double> out(n);
		for (int i = 0; i < n; ++i) out[i] = v[i];
		line("batchverdict", out.data(), n);
	}
	// --- tree: GraphClass / PlannerClass
	{
		PlannerClass T;
		T.init(states[0], false, 1, 1);
		const int nv = std::min(n, 200);
		for (int i = 1; i < nv; ++i) {
			T.addVertex(i, states[i]);
			T.addEdge((i - 1) / 2, i);
			T.addAction(i, actions[i]);
		}
		PlannerClass T2 = T;  // trees are copied wholesale (rrt_connect.cpp:410)
		std::vector<double> nnv, gv;
		for (int q = nv; q < n && q < nv + 40; ++q) nnv.push_back(T2.getNearestNeighbor(states[q]));
		line("nearest", nnv.data(), (int) nnv.size());
		for (int i = 0; i < nv; ++i) gv.push_back(T.getGValue(i));
		line("gvalues", gv.data(), nv);
		std::vector<int> nb = T.neighborhoodDist(states[nv], 3.0);
		std::vector<double> nbv(nb.begin(), nb.end());
		line("near", nbv.data(), (int) nbv.size());
		double misc[3] = {(double) T.getNumVertices(), (double) T.getPredecessor(7), (double) T.getSuccessors(3).size()};
		line("treemisc", misc, 3);
	}
	// --- connect + extend
	{
		RRTConnectClass P;
		for (int i = 0; i + 1 < n && i < 60; i += 2) {
			State sn = states[i + 1];
			Action an;
			an.fill(0.0);
			int st = P.attemptConnect(states[i], states[i + 1], sn, an, terrain, (int) D[i]);
			double w[19] = {(double) st};
			for (int d = 0; d < 8; ++d) w[1 + d] = sn[d];
			for (int d = 0; d < 10; ++d) w[9 + d] = an[d];
			line("connect", w, 19);
		}
		PlannerClass T;
		T.init(states[0], false, 1, 1);
		P.set_random_stream(5, 77);
		P.set_candidates_per_extend(64, true);
		int advanced = 0;
		for (int i = 1; i < 30; ++i) {
			int r = P.extend(T, states[i], terrain, FORWARD);
			advanced += r != TRAPPED;
			if (r != TRAPPED) {
				const int id = T.getNumVertices() - 1;
				State s_new = T.getVertex(id);
				State chk; double tn;
				bool ok = isValidStateActionPair(T.getVertex(T.getPredecessor(id)), T.getAction(id), terrain, chk, tn, false);
				double w[3] = {(double) ok, stateDistance(chk, s_new), (double) isValidState(s_new, terrain, STANCE)};
				line("extendcheck", w, 3);
			}
		}
		double w[2] = {(double) advanced, (double) T.getNumVertices()};
		line("extend", w, 2);
	}
	// --- whole planner: buildRRTConnect, checked primitive by primitive
	if (argc > 3) {
		State start, goal;
		std::ifstream fq(argv[3], std::ios::binary);
		std::vector<double> q = read_doubles(fq, 16);
		for (int d = 0; d < 8; ++d) { start[d] = q[d]; goal[d] = q[8 + d]; }
		RRTConnectClass P;
		P.set_parallel_attempts(1024, 600, 256);
		P.set_max_time_solve(20.0);
		std::vector<State> ss;
		std::vector<Action> aa;
		P.buildRRTConnect(terrain, start, goal, ss, aa, 0.2);
		double plan_time, ttf, dur;
		int succ, nvert;
		std::vector<double> lv, yv, cv, cvt;
		std::vector<std::vector<double>> all;
		P.getStatistics(plan_time, succ, nvert, ttf, lv, yv, cv, cvt, dur, all);
		int chain_ok = ss.size() >= 2 && ss.front() == start;
		double max_gap = 0;
		for (size_t i = 0; i + 1 < ss.size(); ++i) {
			State sn; double tn;
			bool ok = isValidStateActionPair(ss[i], aa[i], terrain, sn, tn, false);
			chain_ok = chain_ok && ok;
			max_gap = std::max(max_gap, stateDistance(sn, ss[i + 1]));
		}
		std::vector<State> ip; std::vector<double> it; std::vector<int> iph;
		if (!ss.empty()) getInterpPath(ss, aa, 0.05, ip, it, iph);
		double w[10] = {(double) ss.size(), (double) chain_ok, max_gap, ss.empty() ? -1 : stateDistance(ss.back(), goal), plan_time,
						(double) succ, (double) nvert, cv.empty() ? -1 : cv.back(), dur, (double) ip.size()};
		line("plan", w, 10);
	}
	return 0;
}
