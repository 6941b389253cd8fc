// extern "C" layer of libgbp_b200.so (include/gbp_b200.h).  Host side: handle lifetime, staging of
// HOST buffers through a chunked copy/compute pipeline, kernel selection and launch geometry.
// No torch types, no CPU fallback: without a CUDA device every compute entry point fails with
// GBP_E_CUDA.
#include <fstream>
#include <sstream>

#include "gbp_host.h"

std::string &gbp_err() {
	static thread_local std::string e;
	return e;
}

extern "C" {


const char *gbp_last_error(void) { return gbp_err().c_str(); }
const char *gbp_version(void) { return "gbp_b200 0.1 (sm_100a, fp64 exact path)"; }
int gbp_device_count(int *count) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) { *count = 0; return fail(GBP_E_CUDA, cudaGetErrorString(e)); }
	*count = n;
	return GBP_OK;
}
int gbp_set_device(int device) { CU(cudaSetDevice(device)); return GBP_OK; }

// ------------------------------------------------------------------------------------- terrain
static bool lossless_f32(const double *a, size_t n) {
	for (size_t i = 0; i < n; ++i) {
		double v = a[i];
		if (v != v) continue;  // NaN stays NaN
		if ((double) (float) v != v) return false;
	}
	return true;
}

int gbp_terrain_create(int nx, int ny, const double *x, const double *y, const double *z, const double *dx, const double *dy,
					   const double *dz, gbp_terrain **out) {
	if (!out) return fail(GBP_E_INVALID, "out is NULL");
	*out = nullptr;
	if (nx < 2 || ny < 2 || !x || !y || !z) return fail(GBP_E_INVALID, "terrain needs nx,ny >= 2 and x,y,z");
	for (int i = 0; i + 1 < nx; ++i) if (!(x[i] < x[i + 1])) return fail(GBP_E_INVALID, "x axis must be strictly increasing");
	for (int i = 0; i + 1 < ny; ++i) if (!(y[i] < y[i + 1])) return fail(GBP_E_INVALID, "y axis must be strictly increasing");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(GBP_E_CUDA, "no CUDA device (this library has no CPU fallback)");
	gbp_terrain *t = new gbp_terrain();
	const size_t cells = (size_t) nx * ny;
	t->hx.assign(x, x + nx);
	t->hy.assign(y, y + ny);
	TerrainView &v = t->view;
	memset(&v, 0, sizeof v);
	v.nx = nx; v.ny = ny;
	cudaError_t e;
#define TRY(call) if ((e = (call)) != cudaSuccess) { gbp_terrain_destroy(t); return fail(GBP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e)); }
	TRY(cudaMalloc(&t->d_x, nx * sizeof(double)));
	TRY(cudaMalloc(&t->d_y, ny * sizeof(double)));
	TRY(cudaMemcpy(t->d_x, x, nx * sizeof(double), cudaMemcpyHostToDevice));
	TRY(cudaMemcpy(t->d_y, y, ny * sizeof(double), cudaMemcpyHostToDevice));
	v.x = t->d_x; v.y = t->d_y;
	v.cell_f32 = lossless_f32(z, cells) ? 1 : 0;
	t->cell_bytes = v.cell_f32 ? 4 : 8;
	if (v.cell_f32) {
		std::vector<float> zf(cells);
		for (size_t i = 0; i < cells; ++i) zf[i] = (float) z[i];
		TRY(cudaMalloc(&t->d_z, cells * sizeof(float)));
		TRY(cudaMemcpy(t->d_z, zf.data(), cells * sizeof(float), cudaMemcpyHostToDevice));
	} else {
		TRY(cudaMalloc(&t->d_z, cells * sizeof(double)));
		TRY(cudaMemcpy(t->d_z, z, cells * sizeof(double), cudaMemcpyHostToDevice));
	}
	v.z = t->d_z;
	if (dx && dy && dz) {
		bool f32 = lossless_f32(dx, cells) && lossless_f32(dy, cells) && lossless_f32(dz, cells);
		const double *src[3] = {dx, dy, dz};
		if (f32) {
			std::vector<float> nf(3 * cells);
			for (int k = 0; k < 3; ++k) for (size_t i = 0; i < cells; ++i) nf[k * cells + i] = (float) src[k][i];
			TRY(cudaMalloc(&t->d_n, 3 * cells * sizeof(float)));
			TRY(cudaMemcpy(t->d_n, nf.data(), 3 * cells * sizeof(float), cudaMemcpyHostToDevice));
			v.nz3 = (const float *) t->d_n;
		} else {
			TRY(cudaMalloc(&t->d_n, 3 * cells * sizeof(double)));
			for (int k = 0; k < 3; ++k) TRY(cudaMemcpy((double *) t->d_n + k * cells, src[k], cells * sizeof(double), cudaMemcpyHostToDevice));
			v.nz3d = (const double *) t->d_n;
		}
	}
	TRY(cudaMalloc(&t->d_cnt, 6 * sizeof(unsigned long long)));
	TRY(cudaMemset(t->d_cnt, 0, 6 * sizeof(unsigned long long)));
#undef TRY
	{  // "hot map pinned in L2": reserve persisting L2 lines for the height grid (streaming inputs/outputs must not evict it)
		t->z_bytes = cells * (size_t) t->cell_bytes;
		int dev = 0, max_persist = 0, max_window = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
		cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
		if (max_persist > 0 && max_window > 0 && t->z_bytes > (1u << 20)) {
			size_t want = t->z_bytes < (size_t) max_persist ? t->z_bytes : (size_t) max_persist;
			if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
				size_t win = t->z_bytes < (size_t) max_window ? t->z_bytes : (size_t) max_window;
				t->l2_hit_ratio = (float) ((double) want / (double) win);
				if (t->l2_hit_ratio > 1.f) t->l2_hit_ratio = 1.f;
			}
		}
		(void) cudaGetLastError();
	}
	v.x0 = x[0]; v.y0 = y[0]; v.x_last = x[nx - 1]; v.y_last = y[ny - 1];
	v.inv_dx = (nx - 1) / (x[nx - 1] - x[0]);
	v.inv_dy = (ny - 1) / (y[ny - 1] - y[0]);
	// uniform axes (every shipped / synthetic / GridMap terrain): edges are recomputed instead of loaded
	v.step_x = (x[nx - 1] - x[0]) / (nx - 1);
	v.step_y = (y[ny - 1] - y[0]) / (ny - 1);
	v.uniform = 1;
	for (int i = 0; i < nx && v.uniform; ++i) if (fabs(x[i] - (x[0] + i * v.step_x)) > 1e-12) v.uniform = 0;
	for (int i = 0; i < ny && v.uniform; ++i) if (fabs(y[i] - (y[0] + i * v.step_y)) > 1e-12) v.uniform = 0;
	bool has_nan = false;
	double zmax = 0.0, dzmax = 0.0;  // largest |z| and largest height step between neighbouring cells (both directions)
	for (size_t i = 0; i < cells; ++i) {
		if (z[i] != z[i]) { has_nan = true; break; }
		if (fabs(z[i]) > zmax) zmax = fabs(z[i]);
		if ((i + 1) % (size_t) ny && fabs(z[i + 1] - z[i]) > dzmax) dzmax = fabs(z[i + 1] - z[i]);
		if (i + (size_t) ny < cells && fabs(z[i + ny] - z[i]) > dzmax) dzmax = fabs(z[i + ny] - z[i]);
	}
	// the mixed evaluator's error budget (< 1e-6 m against its 1e-5 m guard) carries ~1e-7 x the height step inside a cell
	// (fp32 bilinear increments): cliffs above 4 m per cell, like |z| above 4 m for the rounded copy, keep the fp64 walk
	bool mixed = v.uniform && !has_nan && dzmax <= 4.0 && v.step_x >= 0.01 && v.step_y >= 0.01 && !getenv("GBP_NO_MIXED");
	{  // farthest a leg / corner / belly probe can be from the centre: sqrt(0.15^2 + 0.15^2) + 0.05 < 0.27 m
		const double step = v.step_x < v.step_y ? v.step_x : v.step_y;
		v.border = (int) std::ceil(0.27 / step) + 1;
		if (nx < 4 * v.border + 2 || ny < 4 * v.border + 2) mixed = false;  // hardly any interior: fp64 walk throughout
	}
	// The mixed evaluator decides only sub-states whose margins exceed 1e-5 m, so it may read heights ROUNDED to fp32 as
	// long as the rounding stays inside its error budget: |z| <= 4 m keeps it below 2.4e-7 m per cell.  fp64 maps (the
	// shipped CSV maps: 0.1 m steps are not fp32 numbers) therefore get the mixed walk too — through the texture copy
	// only; everything exact (lookups, fp64 evaluator, outputs) keeps reading the fp64 grid.
	const bool rounded_copy = !v.cell_f32 && mixed && zmax <= 4.0;
	v.mixed_ok = (mixed && v.cell_f32) ? 1 : 0;
	if ((v.mixed_ok || rounded_copy) && nx <= 32768 && ny <= 32768 && !getenv("GBP_NO_TEX")) {
		// the mixed-precision walk fetches each probe's 2x2 cells with one texture gather: needs a CUDA array created
		// with the gather flag (array x = iy, array y = ix).  Optional for fp32 maps: on failure the walk keeps the 4-load form.
		cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
		cudaResourceDesc rd;
		cudaTextureDesc td;
		memset(&rd, 0, sizeof rd);
		memset(&td, 0, sizeof td);
		std::vector<float> zr;
		if (rounded_copy) { zr.resize(cells); for (size_t i = 0; i < cells; ++i) zr[i] = (float) z[i]; }
		if (cudaMallocArray(&t->z_arr, &fd, (size_t) ny, (size_t) nx, cudaArrayTextureGather) == cudaSuccess &&
			cudaMemcpy2DToArray(t->z_arr, 0, 0, rounded_copy ? (const void *) zr.data() : (const void *) t->d_z, (size_t) ny * sizeof(float),
								(size_t) ny * sizeof(float), (size_t) nx, rounded_copy ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice) == cudaSuccess) {
			rd.resType = cudaResourceTypeArray;
			rd.res.array.array = t->z_arr;
			td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
			td.filterMode = cudaFilterModePoint;
			td.readMode = cudaReadModeElementType;
			td.normalizedCoords = 0;
			if (cudaCreateTextureObject(&t->z_tex, &rd, &td, nullptr) == cudaSuccess) v.ztex = (unsigned long long) t->z_tex;
		}
		if (!v.ztex && t->z_arr) { cudaFreeArray(t->z_arr); t->z_arr = nullptr; }
		(void) cudaGetLastError();
		if (rounded_copy && v.ztex) v.mixed_ok = 1;
	}
	*out = t;
	return GBP_OK;
}

int gbp_terrain_create_gridmap(int nx, int ny, double res, double cx, double cy, const float *elev, const float *dxl,
							   const float *dyl, const float *dzl, gbp_terrain **out) {
	if (nx < 2 || ny < 2 || !elev || !(res > 0)) return fail(GBP_E_INVALID, "bad grid map");
	// fast_terrain_map.cpp:43-54: x_data[i] = position of index (nx-1-i, 0); layers flipped likewise
	std::vector<double> x(nx), y(ny), z((size_t) nx * ny), a, b, c;
	for (int i = 0; i < nx; ++i) x[i] = cx + (0.5 * (nx - 1) - ((nx - 1) - i)) * res;
	for (int j = 0; j < ny; ++j) y[j] = cy + (0.5 * (ny - 1) - ((ny - 1) - j)) * res;
	const bool normals = dxl && dyl && dzl;
	if (normals) { a.resize(z.size()); b.resize(z.size()); c.resize(z.size()); }
	for (int i = 0; i < nx; ++i)
		for (int j = 0; j < ny; ++j) {
			size_t src = (size_t) ((nx - 1) - i) * ny + ((ny - 1) - j), dst = (size_t) i * ny + j;
			z[dst] = (double) elev[src];
			if (normals) { a[dst] = (double) dxl[src]; b[dst] = (double) dyl[src]; c[dst] = (double) dzl[src]; }
		}
	return gbp_terrain_create(nx, ny, x.data(), y.data(), z.data(), normals ? a.data() : nullptr, normals ? b.data() : nullptr,
							  normals ? c.data() : nullptr, out);
}

// TerrainMapPublisher::loadCSV (terrain_map_publisher.cpp:289-327): '#' lines skipped, fields split on ',', std::stod
static bool load_csv(const std::string &path, std::vector<std::vector<double>> &data, std::string &err) {
	std::ifstream in(path);
	if (!in) { err = "cannot open " + path; return false; }
	std::string line;
	while (std::getline(in, line)) {
		if (!line.empty() && line[0] == '#') continue;
		std::vector<double> rec;
		std::istringstream ss(line);
		std::string num;
		while (std::getline(ss, num, ',')) {
			try { rec.push_back(std::stod(num)); } catch (const std::exception &) { /* the reference drops unparsable fields */ }
		}
		data.push_back(rec);
	}
	while (!data.empty() && data.back().empty()) data.pop_back();  // trailing blank lines
	return true;
}

int gbp_terrain_create_csv(const char *directory, int via_gridmap, gbp_terrain **out) {
	if (!out) return fail(GBP_E_INVALID, "out is NULL");
	*out = nullptr;
	if (!directory) return fail(GBP_E_INVALID, "directory is NULL");
	const char *names[6] = {"x", "y", "z", "dx", "dy", "dz"};
	std::vector<std::vector<double>> L[6];
	std::string err;
	for (int k = 0; k < 6; ++k)
		if (!load_csv(std::string(directory) + "/" + names[k] + "data.csv", L[k], err)) return fail(GBP_E_INVALID, err);
	const size_t ny = L[2].size(), nx = ny ? L[2][0].size() : 0;  // rows = y, columns = x (:343-344)
	if (nx < 2 || ny < 2) return fail(GBP_E_INVALID, "zdata.csv needs at least 2 x 2 values");
	for (int k = 0; k < 6; ++k) {
		if (L[k].size() != ny) return fail(GBP_E_INVALID, std::string(names[k]) + "data.csv: row count differs from zdata.csv");
		for (const auto &row : L[k]) if (row.size() != nx) return fail(GBP_E_INVALID, std::string(names[k]) + "data.csv: ragged rows");
	}
	if (!via_gridmap) {
		std::vector<double> x(nx), y(ny), lay[4];
		for (size_t i = 0; i < nx; ++i) x[i] = L[0][0][i];
		for (size_t j = 0; j < ny; ++j) y[j] = L[1][j][0];
		for (int k = 0; k < 4; ++k) {
			lay[k].resize(nx * ny);
			for (size_t i = 0; i < nx; ++i) for (size_t j = 0; j < ny; ++j) lay[k][i * ny + j] = L[2 + k][j][i];
		}
		return gbp_terrain_create((int) nx, (int) ny, x.data(), y.data(), lay[0].data(), lay[1].data(), lay[2].data(), lay[3].data(), out);
	}
	// the ROS path (:345-369): float resolution, map centred so that cell centres sit on the data points, float layers
	const float x_res = (float) (L[0][0][1] - L[0][0][0]), y_res = (float) (L[1][1][0] - L[1][0][0]);
	if (x_res != y_res) return fail(GBP_E_INVALID, "Map did not have square elements, make sure x and y resolution are equal.");
	const double x_length = L[0][0].back() - L[0][0].front() + x_res, y_length = L[1].back()[0] - L[1].front()[0] + y_res;
	const double cx = L[0][0].front() - 0.5 * x_res + 0.5 * x_length, cy = L[1].front()[0] - 0.5 * y_res + 0.5 * y_length;
	std::vector<float> lay[4];
	for (int k = 0; k < 4; ++k) {
		lay[k].resize(nx * ny);
		for (size_t i = 0; i < nx; ++i)
			for (size_t j = 0; j < ny; ++j) lay[k][i * ny + j] = (float) L[2 + k][(ny - 1) - j][(nx - 1) - i];  // grid_map index (i, j), :363-367
	}
	return gbp_terrain_create_gridmap((int) nx, (int) ny, (double) x_res, cx, cy, lay[0].data(), lay[1].data(), lay[2].data(), lay[3].data(), out);
}

// TerrainMapPublisher::findXYIndex (terrain_map_publisher.cpp:178-231): rectangle [x1, x2] x [y1, y2] -> node index
// range [lo, hi) on one axis.  v == last node leaves the reference's lower index uninitialised; defined as the last node.
static int own_index_lo(const std::vector<double> &ax, double v) {
	if (v <= ax.front()) return 0;
	for (size_t i = 0; i + 1 < ax.size(); ++i)
		if (ax[i] <= v && v < ax[i + 1]) return (int) i;
	return (int) ax.size() - 1;
}
static int own_index_hi(const std::vector<double> &ax, double v) {
	if (v >= ax.back()) return (int) ax.size();
	for (size_t i = ax.size() - 1; i > 0; --i)
		if (ax[i - 1] <= v && v < ax[i]) return (int) i;
	return 0;
}
static const double k_own_map_rects[13][6] = {  // changeOwnMapZData (:107-127): global 1 cm noise, then 12 boxes
	{-DBL_MAX, -DBL_MAX, DBL_MAX, DBL_MAX, 0, 0.01},
	{8.13, -4, 8.42, 4, 0.158, 0.01}, {8.42, -4, 8.71, 4, 0.316, 0.01}, {8.71, -4, 10.5, 4, 0.474, 0.01},
	{0.75, -3.15, 2.05, -2.35, 0.6, 0.1}, {4.25, -2.4, 5.4, -1.75, 0.5, 0.05}, {2.9, -0.6, 3.25, 1.15, 0.158, 0.01},
	{4.9, -0.5, 5.3, 0.35, -0.3, 0.01}, {6.5, 0.45, 7.2, 1.05, 0.7, 0.07}, {0.65, 2.95, 1.15, 3.75, 0.3, 0.08},
	{4.4, 2.8, 5.7, 3.55, 0.65, 0.04}, {7.5, -2.6, 9.45, -1.15, 0.68, 0.06}, {6.2, 1.1, 9.2, 2.3, -0.2, 0.06}};

int gbp_own_map_layer(uint64_t seed, int x_size, int y_size, double x_start, double y_start, double res, int n_rect,
					  const double *rects, float *elevation, double *geometry3) {
	if (x_size < 2 || y_size < 2 || (int64_t) x_size * y_size > (1ll << 30) || !(res > 0) || !elevation || !geometry3 || n_rect < 0 ||
		(n_rect && !rects))
		return fail(GBP_E_INVALID, "bad own-map arguments");
	if (!rects) { rects = &k_own_map_rects[0][0]; n_rect = 13; }
	// axes (:46-60): centimetre-rounded accumulation
	std::vector<double> xa(x_size), ya(y_size);
	double v = x_start;
	for (int i = 0; i < x_size; ++i) { xa[i] = std::round(v * 100) / 100; v = std::round((v + res) * 100) / 100; }
	v = y_start;
	for (int i = 0; i < y_size; ++i) { ya[i] = std::round(v * 100) / 100; v = std::round((v + res) * 100) / 100; }
	std::vector<OwnMapRect> rr(n_rect ? n_rect : 1);
	for (int r = 0; r < n_rect; ++r) {
		const double x1 = rects[6 * r], y1 = rects[6 * r + 1], x2 = rects[6 * r + 2], y2 = rects[6 * r + 3];
		OwnMapRect q = {0, 0, 0, 0, rects[6 * r + 4], rects[6 * r + 5]};
		if (!(x1 > xa.back() || x2 < xa.front() || y1 > ya.back() || y2 < ya.front() || x1 >= x2 || y1 >= y2)) {  // :153-156
			q.x1 = own_index_lo(xa, x1); q.y1 = own_index_lo(ya, y1);
			q.x2 = own_index_hi(xa, x2); q.y2 = own_index_hi(ya, y2);
		}
		rr[r] = q;
	}
	cudaStream_t st = lib_stream();
	Dev dr(st), de(st);
	int rc;
	if ((rc = upload(dr, rr.data(), rr.size(), st))) return rc;
	const size_t cells = (size_t) x_size * y_size;
	CU(de.alloc(cells * sizeof(float)));
	k_own_map<<<blocks_for((int64_t) cells, 256), 256, 0, st>>>(seed, x_size, y_size, n_rect, dr.as<OwnMapRect>(), de.as<float>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(elevation, de.p, cells * sizeof(float), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	const double x_end = x_start + res * (x_size - 1), y_end = y_start + res * (y_size - 1);  // :41-44
	const double x_length = x_end - x_start + res, y_length = y_end - y_start + res;
	geometry3[0] = res;
	geometry3[1] = xa.front() - 0.5 * res + 0.5 * x_length;  // :76-78
	geometry3[2] = ya.front() - 0.5 * res + 0.5 * y_length;
	return GBP_OK;
}

int gbp_terrain_create_own_map(uint64_t seed, int x_size, int y_size, double x_start, double y_start, double res, int n_rect,
							   const double *rects, gbp_terrain **out) {
	if (!out) return fail(GBP_E_INVALID, "out is NULL");
	*out = nullptr;
	if (x_size < 2 || y_size < 2 || (int64_t) x_size * y_size > (1ll << 30)) return fail(GBP_E_INVALID, "bad own-map arguments");
	std::vector<float> elev((size_t) x_size * y_size);
	double g[3];
	int rc = gbp_own_map_layer(seed, x_size, y_size, x_start, y_start, res, n_rect, rects, elev.data(), g);
	if (rc) return rc;
	// createOwnMap erases the dx / dy / dz layers (:83-85): loadDataFromGridMap then fills (0, 0, 1) (fast_terrain_map.cpp:68-72)
	return gbp_terrain_create_gridmap(x_size, y_size, g[0], g[1], g[2], elev.data(), nullptr, nullptr, nullptr, out);
}

// TerrainMapPublisher::createMap (terrain_map_publisher.cpp:253-286), the publisher's default source
int gbp_terrain_create_default_map(gbp_terrain **out) {
	if (!out) return fail(GBP_E_INVALID, "out is NULL");
	*out = nullptr;
	const int nx = 60, ny = 25;  // Length (12, 5) at 0.2 m
	const double res = 0.2, cx = 4.0, cy = 0.0;
	std::vector<float> elev((size_t) nx * ny), zero(elev.size(), 0.0f), one(elev.size(), 1.0f);
	for (int i = 0; i < nx; ++i)
		for (int j = 0; j < ny; ++j) {
			const double px = cx + (0.5 * (nx - 1) - i) * res, py = cy + (0.5 * (ny - 1) - j) * res;
			const double xd = px - 2, yd = py - 0;
			elev[(size_t) i * ny + j] = (xd * xd + yd * yd <= 0.5 * 0.5) ? 0.1f : 0.0f;
		}
	return gbp_terrain_create_gridmap(nx, ny, res, cx, cy, elev.data(), zero.data(), zero.data(), one.data(), out);
}

void gbp_terrain_destroy(gbp_terrain *t) {
	if (!t) return;
	if (t->z_tex) cudaDestroyTextureObject(t->z_tex);
	if (t->z_arr) cudaFreeArray(t->z_arr);
	cudaFree(t->d_x); cudaFree(t->d_y); cudaFree(t->d_z); cudaFree(t->d_n); cudaFree(t->d_cnt);
	for (int k = 0; k < HostPipe::NBUF; ++k) {
		cudaFree(t->pipe.in[k]); cudaFree(t->pipe.out[k]); cudaFree(t->pipe.redo[k]);
		if (t->pipe.st[k]) cudaStreamDestroy(t->pipe.st[k]);
	}
	if (t->pipe.ready) cudaEventDestroy(t->pipe.ready);
	delete t;
}
int gbp_terrain_dims(const gbp_terrain *t, int *nx, int *ny, int *cell_bytes) {
	if (!t) return fail(GBP_E_INVALID, "terrain is NULL");
	if (nx) *nx = t->view.nx;
	if (ny) *ny = t->view.ny;
	if (cell_bytes) *cell_bytes = t->cell_bytes;
	return GBP_OK;
}
int gbp_terrain_flags(const gbp_terrain *t, int *uniform_axes, int *mixed_precision) {
	if (!t) return fail(GBP_E_INVALID, "terrain is NULL");
	if (uniform_axes) *uniform_axes = t->view.uniform;
	if (mixed_precision) *mixed_precision = t->view.mixed_ok;
	return GBP_OK;
}
int gbp_terrain_fetch_path(const gbp_terrain *t, int *texture_gather) {
	if (!t) return fail(GBP_E_INVALID, "terrain is NULL");
	if (texture_gather) *texture_gather = t->view.ztex ? 1 : 0;
	return GBP_OK;
}
int gbp_terrain_axes(const gbp_terrain *t, double *x, double *y) {
	if (!t) return fail(GBP_E_INVALID, "terrain is NULL");
	if (x) memcpy(x, t->hx.data(), t->hx.size() * sizeof(double));
	if (y) memcpy(y, t->hy.data(), t->hy.size() * sizeof(double));
	return GBP_OK;
}

static int terrain_query(const gbp_terrain *t, int64_t n, const double *x, const double *y, int what, double *out, uint8_t *out8) {
	if (!t || n < 0 || (n && (!x || !y))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev dx(st), dy(st), dout(st), dout8(st);
	int rc;
	if ((rc = upload(dx, x, n, st)) || (rc = upload(dy, y, n, st))) return rc;
	const size_t per = what == 2 ? 3 : 1;
	if (out) CU(dout.alloc(n * per * sizeof(double)));
	if (out8) CU(dout8.alloc(n));
	GBP_DISPATCH(t->view, k_terrain_query, (blocks_for(n, 256), 256), st, t->view, n, dx.as<double>(), dy.as<double>(), what, dout.as<double>(), dout8.as<uint8_t>());
	CU(cudaGetLastError());
	if (out) CU(cudaMemcpyAsync(out, dout.p, n * per * sizeof(double), cudaMemcpyDeviceToHost, st));
	if (out8) CU(cudaMemcpyAsync(out8, dout8.p, n, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_ground_height(const gbp_terrain *t, int64_t n, const double *x, const double *y, double *h, uint8_t *flags) {
	if (!h) return fail(GBP_E_INVALID, "h is NULL");
	return terrain_query(t, n, x, y, 0, h, flags);
}
int gbp_height_is_nan(const gbp_terrain *t, int64_t n, const double *x, const double *y, uint8_t *is_nan) {
	if (!is_nan) return fail(GBP_E_INVALID, "is_nan is NULL");
	return terrain_query(t, n, x, y, 1, nullptr, is_nan);
}
int gbp_surface_normal(const gbp_terrain *t, int64_t n, const double *x, const double *y, double *normal3) {
	if (!normal3) return fail(GBP_E_INVALID, "normal3 is NULL");
	return terrain_query(t, n, x, y, 2, normal3, nullptr);
}

// ---------------------------------------------------------------------------------- primitives
int gbp_propagate(int kind, int64_t n, const double *states, const double *actions, const double *t, double *out) {
	if (kind < 0 || kind > 2 || n < 0 || (n && (!states || !t || !out || (kind != 1 && !actions)))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev ds(st), da(st), dt(st), dout(st);
	int rc;
	if ((rc = upload(ds, states, 8 * n, st)) || (rc = upload(dt, t, n, st))) return rc;
	if (kind != 1 && (rc = upload(da, actions, 10 * n, st))) return rc;
	CU(dout.alloc(8 * n * sizeof(double)));
	k_propagate<<<blocks_for(n, 256), 256, 0, st>>>(kind, n, ds.as<double>(), da.as<double>(), dt.as<double>(), dout.as<double>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(out, dout.p, 8 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_rotate_grf(int64_t n, const double *normal3, const double *force3, double *out3) {
	if (n < 0 || (n && (!normal3 || !force3 || !out3))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev dn(st), df(st), dout(st);
	int rc;
	if ((rc = upload(dn, normal3, 3 * n, st)) || (rc = upload(df, force3, 3 * n, st))) return rc;
	CU(dout.alloc(3 * n * sizeof(double)));
	k_rotate_grf<<<blocks_for(n, 256), 256, 0, st>>>(n, dn.as<double>(), df.as<double>(), dout.as<double>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(out3, dout.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_curvature(int64_t n, const double *points6, double *curvature) {
	if (n < 0 || (n && (!points6 || !curvature))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev dp(st), dout(st);
	int rc;
	if ((rc = upload(dp, points6, 6 * n, st))) return rc;
	CU(dout.alloc(n * sizeof(double)));
	k_curvature<<<blocks_for(n, 256), 256, 0, st>>>(n, dp.as<double>(), dout.as<double>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(curvature, dout.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_valid_actions(int64_t n, const double *actions, uint8_t *verdict) {
	if (n < 0 || (n && (!actions || !verdict))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev da(st), dv(st);
	int rc;
	if ((rc = upload(da, actions, 10 * n, st))) return rc;
	CU(dv.alloc(n));
	k_valid_actions<<<blocks_for(n, 256), 256, 0, st>>>(n, da.as<double>(), dv.as<uint8_t>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(verdict, dv.p, n, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_valid_states(const gbp_terrain *t, int64_t n, const double *states, const uint8_t *phase, uint8_t *verdict, uint8_t *flags) {
	if (!t || n < 0 || (n && (!states || !phase || !verdict))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev ds(st), dp(st), dv(st), df(st);
	int rc;
	if ((rc = upload(ds, states, 8 * n, st)) || (rc = upload(dp, phase, n, st))) return rc;
	CU(dv.alloc(n));
	if (flags) CU(df.alloc(n));
	GBP_DISPATCH(t->view, k_valid_states, (blocks_for(n, 128), 128), st, t->view, n, ds.as<double>(), dp.as<uint8_t>(), dv.as<uint8_t>(), df.as<uint8_t>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(verdict, dv.p, n, cudaMemcpyDeviceToHost, st));
	if (flags) CU(cudaMemcpyAsync(flags, df.p, n, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_valid_states_dev(const gbp_terrain *t, int64_t n, const double *states, const uint8_t *phase, uint8_t *verdict, uint8_t *flags,
						 void *stream) {
	if (!t || n < 0 || (n && (!states || !phase || !verdict))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = (cudaStream_t) stream;
	GBP_DISPATCH(t->view, k_valid_states, (blocks_for(n, 128), 128), st, t->view, n, states, phase, verdict, flags);
	CU(cudaGetLastError());
	return GBP_OK;
}
int gbp_distance(int kind, int64_t n, const double *q1, const double *q2, double *out) {
	if (kind < 0 || kind > 2 || n < 0 || (n && (!q1 || !q2 || !out))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev a(st), b(st), o(st);
	int rc;
	if ((rc = upload(a, q1, 8 * n, st)) || (rc = upload(b, q2, 8 * n, st))) return rc;
	CU(o.alloc(n * sizeof(double)));
	k_distance<<<blocks_for(n, 256), 256, 0, st>>>(kind, n, a.as<double>(), b.as<double>(), o.as<double>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(out, o.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}

// ---------------------------------------------------------------------------------- plan output
int gbp_interp_path(int n_actions, const double *states, const double *actions, double dt, int64_t cap, double *interp_states,
					double *interp_t, int *interp_phase, int64_t *count) {
	if (n_actions < 0 || !states || (n_actions && !actions) || !(dt > 0) || cap < 0 || !count || (cap && (!interp_states || !interp_t || !interp_phase)))
		return fail(GBP_E_INVALID, "bad arguments (dt must be positive)");
	// the sample grid, exactly as the reference's loops lay it out (planning_utils.cpp:142-193)
	std::vector<int> prim;
	std::vector<uint8_t> kind;
	std::vector<double> tloc, tabs;
	std::vector<int> phase;
	double t0 = 0;
	for (int i = 0; i < n_actions; ++i) {
		const double t_s = actions[10 * (size_t) i + 6], t_f = actions[10 * (size_t) i + 7];
		if (!(t_s < 1e6) || !(t_f < 1e6)) return fail(GBP_E_INVALID, "primitive duration out of range");
		for (double t = 0; t < t_s; t += dt) { prim.push_back(i); kind.push_back(0); tloc.push_back(t); tabs.push_back(t + t0); phase.push_back(t_f == 0 ? 2 : GBP_STANCE); }
		for (double t = 0; t < t_f; t += dt) { prim.push_back(i); kind.push_back(1); tloc.push_back(t); tabs.push_back(t_s + t + t0); phase.push_back(GBP_FLIGHT); }
		if (t_f > 0) { prim.push_back(i); kind.push_back(1); tloc.push_back(t_f); tabs.push_back(t0 + t_s + t_f); phase.push_back(GBP_STANCE); }
		t0 += (t_s + t_f);
	}
	prim.push_back(n_actions); kind.push_back(2); tloc.push_back(0.0); tabs.push_back(t0);
	const int64_t m = (int64_t) prim.size();
	*count = m;
	const int64_t w = m < cap ? m : cap;
	if (w == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev dp(st), dk(st), dt_(st), ds(st), da(st), dout(st);
	int rc;
	if ((rc = upload(dp, prim.data(), (size_t) w, st)) || (rc = upload(dk, kind.data(), (size_t) w, st)) || (rc = upload(dt_, tloc.data(), (size_t) w, st)) ||
		(rc = upload(ds, states, 8 * ((size_t) n_actions + 1), st)) || (rc = upload(da, actions, 10 * (size_t) n_actions, st)))
		return rc;
	CU(dout.alloc((size_t) w * 8 * sizeof(double)));
	k_interp_samples<<<blocks_for(w, 128), 128, 0, st>>>(w, dp.as<int>(), dk.as<uint8_t>(), dt_.as<double>(), ds.as<double>(), da.as<double>(), dout.as<double>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(interp_states, dout.p, (size_t) w * 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	memcpy(interp_t, tabs.data(), (size_t) w * sizeof(double));
	const int64_t np = (int64_t) phase.size() < w ? (int64_t) phase.size() : w;
	if (np) memcpy(interp_phase, phase.data(), (size_t) np * sizeof(int));
	return GBP_OK;
}
int gbp_max_curvature(int64_t n, const double *states, double *max_curvature) {
	if (n < 0 || (n && !states) || !max_curvature) return fail(GBP_E_INVALID, "bad arguments");
	*max_curvature = 0.0;
	if (n < 3) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev ds(st), dm(st);
	int rc;
	if ((rc = upload(ds, states, 8 * (size_t) n, st))) return rc;
	CU(dm.alloc(sizeof(unsigned long long)));
	CU(cudaMemsetAsync(dm.p, 0, sizeof(unsigned long long), st));
	k_max_curvature<<<blocks_for(n - 2, 128), 128, 0, st>>>(n, ds.as<double>(), dm.as<unsigned long long>());
	CU(cudaGetLastError());
	unsigned long long bits = 0;
	CU(cudaMemcpyAsync(&bits, dm.p, sizeof bits, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	memcpy(max_curvature, &bits, sizeof bits);
	return GBP_OK;
}

// ------------------------------------------------------------------------------------ samplers
int gbp_sample_actions_dev(uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n, const double *normal3_host, double *actions,
						   void *cuda_stream) {
	if (n < 0 || (n && !actions)) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	double n0 = 0, n1 = 0, n2 = 1;
	if (normal3_host) { n0 = normal3_host[0]; n1 = normal3_host[1]; n2 = normal3_host[2]; }
	k_sample_actions<<<blocks_for(n, 256), 256, 0, (cudaStream_t) cuda_stream>>>(seed, stream, idx0, n, n0, n1, n2, 0, 0.0, nullptr, nullptr, actions);
	CU(cudaGetLastError());
	return GBP_OK;
}
int gbp_sample_actions(uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n, const double *normal3, const double *s_from,
					   const double *s_to, double dir_threshold, double *actions) {
	if (n < 0 || (n && !actions)) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev da(st), df(st), dt(st);
	CU(da.alloc(10 * n * sizeof(double)));
	const int dir = (s_from && s_to) ? 1 : 0;
	int rc;
	if (dir && ((rc = upload(df, s_from, 8, st)) || (rc = upload(dt, s_to, 8, st)))) return rc;
	double n0 = 0, n1 = 0, n2 = 1;
	if (normal3) { n0 = normal3[0]; n1 = normal3[1]; n2 = normal3[2]; }
	k_sample_actions<<<blocks_for(n, 256), 256, 0, st>>>(seed, stream, idx0, n, n0, n1, n2, dir, dir_threshold, df.as<double>(), dt.as<double>(), da.as<double>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(actions, da.p, 10 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
static int sample_states_launch(const gbp_terrain *t, uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n, int dir, double thr,
								int speed, const double *d_from, const double *d_to, double *d_out, cudaStream_t st) {
	GBP_DISPATCH(t->view, k_sample_states, (blocks_for(n, 256), 256), st, t->view, seed, stream, idx0, n, dir, thr, speed, d_from, d_to, d_out);
	CU(cudaGetLastError());
	return GBP_OK;
}
int gbp_sample_states_dev(const gbp_terrain *t, uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n, double *states, void *cuda_stream) {
	if (!t || n < 0 || (n && !states)) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	return sample_states_launch(t, seed, stream, idx0, n, 0, 0.0, 0, nullptr, nullptr, states, (cudaStream_t) cuda_stream);
}
int gbp_sample_states(const gbp_terrain *t, uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n, const double *s_from,
					  const double *s_to, double dir_threshold, int speed_direction, double *states) {
	if (!t || n < 0 || (n && !states)) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev ds(st), df(st), dt(st);
	CU(ds.alloc(8 * n * sizeof(double)));
	const int dir = (s_from && s_to) ? 1 : 0;
	int rc;
	if (dir && ((rc = upload(df, s_from, 8, st)) || (rc = upload(dt, s_to, 8, st)))) return rc;
	if ((rc = sample_states_launch(t, seed, stream, idx0, n, dir, dir_threshold, speed_direction, df.as<double>(), dt.as<double>(), ds.as<double>(), st))) return rc;
	CU(cudaMemcpyAsync(states, ds.p, 8 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}

// --------------------------------------------------------------------------------- tree store
void gbp_tree_destroy(gbp_tree *T) {
	if (!T) return;
	cudaFree(T->d_n); cudaFree(T->d_v); cudaFree(T->d_act); cudaFree(T->d_g); cudaFree(T->d_y); cudaFree(T->d_parent);
	cudaFree(T->S.near_idx); cudaFree(T->S.near_dist); cudaFree(T->S.valid); cudaFree(T->S.dist); cudaFree(T->S.s_test); cudaFree(T->S.result);
	cudaFree(T->d_target); cudaFree(T->d_done);
	if (T->h_result) cudaFreeHost(T->h_result);
	delete T;
}
int gbp_tree_create(int capacity, gbp_tree **out) {
	if (!out || capacity < 1) return fail(GBP_E_INVALID, "bad arguments");
	*out = nullptr;
	gbp_tree *T = new gbp_tree();
	memset(&T->S, 0, sizeof T->S);
	cudaError_t e;
#define TRY(call) if ((e = (call)) != cudaSuccess) { gbp_tree_destroy(T); return fail(GBP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e)); }
	TRY(cudaMalloc(&T->d_n, sizeof(int)));
	TRY(cudaMemset(T->d_n, 0, sizeof(int)));
	TRY(cudaMalloc(&T->d_v, sizeof(double) * 8 * capacity));
	TRY(cudaMalloc(&T->d_act, sizeof(double) * 10 * capacity));
	TRY(cudaMalloc(&T->d_g, sizeof(double) * capacity));
	TRY(cudaMalloc(&T->d_y, sizeof(double) * capacity));
	TRY(cudaMalloc(&T->d_parent, sizeof(int) * capacity));
	TRY(cudaMalloc(&T->S.near_idx, sizeof(int)));
	TRY(cudaMalloc(&T->S.near_dist, sizeof(double)));
	TRY(cudaMalloc(&T->S.result, 4 * sizeof(int)));
	TRY(cudaMalloc(&T->d_target, 18 * sizeof(double)));
	TRY(cudaMalloc(&T->d_done, sizeof(unsigned)));
	TRY(cudaMemset(T->d_done, 0, sizeof(unsigned)));
	TRY(cudaHostAlloc((void **) &T->h_result, 4 * sizeof(int), cudaHostAllocMapped));
#undef TRY
	T->view.cap = capacity; T->view.n = T->d_n; T->view.v = T->d_v; T->view.act = T->d_act; T->view.parent = T->d_parent;
	T->view.g = T->d_g; T->view.y = T->d_y;
	*out = T;
	return GBP_OK;
}
int gbp_tree_init(gbp_tree *T, const double *root_state) {
	if (!T || !root_state) return fail(GBP_E_INVALID, "bad arguments");
	cudaStream_t st = lib_stream();
	CU(cudaMemcpyAsync(T->d_target, root_state, 8 * sizeof(double), cudaMemcpyHostToDevice, st));
	k_tree_init<<<1, 32, 0, st>>>(T->view, T->d_target);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_tree_size(const gbp_tree *T, int *n) {
	if (!T || !n) return fail(GBP_E_INVALID, "bad arguments");
	CU(cudaMemcpy(n, T->d_n, sizeof(int), cudaMemcpyDeviceToHost));
	return GBP_OK;
}
int gbp_tree_append(gbp_tree *T, int parent, const double *state, const double *action, int *new_id) {
	if (!T || !state || !action) return fail(GBP_E_INVALID, "bad arguments");
	cudaStream_t st = lib_stream();
	CU(cudaMemcpyAsync(T->d_target, state, 8 * sizeof(double), cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(T->d_target + 8, action, 10 * sizeof(double), cudaMemcpyHostToDevice, st));
	k_tree_append<<<1, 32, 0, st>>>(T->view, parent, T->d_target, T->d_target + 8, T->S.result);
	CU(cudaGetLastError());
	int id = -1;
	CU(cudaMemcpyAsync(&id, T->S.result, sizeof(int), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	if (id < 0) return fail(GBP_E_CAPACITY, "tree full or bad parent");
	if (new_id) *new_id = id;
	return GBP_OK;
}
int gbp_tree_load(gbp_tree *T, int n, const double *states, const double *actions, const int *parent) {
	if (!T || n < 1 || !states || (n > 1 && !parent)) return fail(GBP_E_INVALID, "bad arguments");
	if (n > T->view.cap) return fail(GBP_E_CAPACITY, "tree capacity exceeded");
	for (int i = 1; i < n; ++i) if (parent[i] < 0 || parent[i] >= i) return fail(GBP_E_INVALID, "parent ids must precede their children");
	cudaStream_t st = lib_stream();
	Dev ds(st), da(st), dp(st);
	int rc;
	if ((rc = upload(ds, states, (size_t) 8 * n, st))) return rc;
	if (actions && (rc = upload(da, actions, (size_t) 10 * n, st))) return rc;
	if (n > 1) { if ((rc = upload(dp, parent, (size_t) n, st))) return rc; }
	k_tree_load<<<blocks_for(n, 256), 256, 0, st>>>(T->view, n, ds.as<double>(), actions ? da.as<double>() : nullptr, dp.as<int>());
	k_tree_gy<<<1, 32, 0, st>>>(T->view, n);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_tree_read(const gbp_tree *T, int first, int n, double *states, double *actions, int *parent, double *g, double *yaw) {
	if (!T || first < 0 || n < 0) return fail(GBP_E_INVALID, "bad arguments");
	int have = 0;
	CU(cudaMemcpy(&have, T->d_n, sizeof(int), cudaMemcpyDeviceToHost));
	if (first + n > have) return fail(GBP_E_INVALID, "range exceeds tree size");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev ds(st), da(st), dp(st), dg(st), dy(st);
	if (states) CU(ds.alloc(sizeof(double) * 8 * n));
	if (actions) CU(da.alloc(sizeof(double) * 10 * n));
	if (parent) CU(dp.alloc(sizeof(int) * n));
	if (g) CU(dg.alloc(sizeof(double) * n));
	if (yaw) CU(dy.alloc(sizeof(double) * n));
	k_tree_read<<<blocks_for(n, 256), 256, 0, st>>>(T->view, first, n, ds.as<double>(), da.as<double>(), dp.as<int>(), dg.as<double>(), dy.as<double>());
	CU(cudaGetLastError());
	if (states) CU(cudaMemcpyAsync(states, ds.p, sizeof(double) * 8 * n, cudaMemcpyDeviceToHost, st));
	if (actions) CU(cudaMemcpyAsync(actions, da.p, sizeof(double) * 10 * n, cudaMemcpyDeviceToHost, st));
	if (parent) CU(cudaMemcpyAsync(parent, dp.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
	if (g) CU(cudaMemcpyAsync(g, dg.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
	if (yaw) CU(cudaMemcpyAsync(yaw, dy.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_nearest_dev(const gbp_tree *T, int64_t m, const double *queries, int *idx, double *dist, void *stream) {
	if (!T || m < 0 || (m && (!queries || !idx))) return fail(GBP_E_INVALID, "bad arguments");
	if (m == 0) return GBP_OK;
	if (m >= 4 * (int64_t) sm_count()) {  // enough queries to fill the GPU with 4-query tiles
		const int64_t tiles = (m + 3) / 4;
		k_nearest_tiled<4><<<(unsigned) (tiles < 65535 ? tiles : 65535), 256, 0, (cudaStream_t) stream>>>(T->view, m, queries, idx, dist);
	} else {
		unsigned grid = (unsigned) (m < 65535 ? m : 65535);
		k_nearest<<<grid, 256, 0, (cudaStream_t) stream>>>(T->view, m, queries, idx, dist);
	}
	CU(cudaGetLastError());
	return GBP_OK;
}
int gbp_nearest(const gbp_tree *T, int64_t m, const double *queries, int *idx, double *dist) {
	if (!T || m < 0 || (m && (!queries || !idx))) return fail(GBP_E_INVALID, "bad arguments");
	if (m == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev dq(st), di(st), dd(st);
	int rc;
	if ((rc = upload(dq, queries, (size_t) 8 * m, st))) return rc;
	CU(di.alloc(sizeof(int) * m));
	CU(dd.alloc(sizeof(double) * m));
	if ((rc = gbp_nearest_dev(T, m, dq.as<double>(), di.as<int>(), dd.as<double>(), st))) return rc;
	CU(cudaMemcpyAsync(idx, di.p, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
	if (dist) CU(cudaMemcpyAsync(dist, dd.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_near(const gbp_tree *T, const double *query, double radius, int *ids, int cap, int *count) {
	if (!T || !query || !count || cap < 0 || (cap && !ids)) return fail(GBP_E_INVALID, "bad arguments");
	cudaStream_t st = lib_stream();
	Dev dq(st), di(st), dc(st);
	int rc;
	if ((rc = upload(dq, query, 8, st))) return rc;
	CU(di.alloc(sizeof(int) * (cap ? cap : 1)));
	CU(dc.alloc(sizeof(int)));
	k_near<<<1, 32, 0, st>>>(T->view, dq.as<double>(), radius, di.as<int>(), cap, dc.as<int>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(count, dc.p, sizeof(int), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	int m = *count < cap ? *count : cap;
	if (m > 0) CU(cudaMemcpy(ids, di.p, sizeof(int) * m, cudaMemcpyDeviceToHost));
	return GBP_OK;
}

}  // extern "C"
