// Drop-in for the reference's include/global_body_planner/fast_terrain_map.h (class FastTerrainMap,
// same public signatures, :14-120), backed by a device-resident terrain (include/gbp_b200.h).
// Every query runs on the GPU through the C ABI; there is no host-side lookup code.
#ifndef GBP_DROPIN_FAST_TERRAIN_MAP_H
#define GBP_DROPIN_FAST_TERRAIN_MAP_H

#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../gbp_b200.h"

#if defined(__has_include)
#if __has_include(<grid_map_core/grid_map_core.hpp>)
#include <grid_map_core/grid_map_core.hpp>
#define GBP_HAVE_GRID_MAP 1
#endif
#endif

class FastTerrainMap {
public:
	FastTerrainMap();

	// reference: fast_terrain_map.cpp:10-28.  z/dx/dy/dz are [x][y] nested vectors.
	void loadData(int x_size, int y_size, std::vector<double> x_data, std::vector<double> y_data,
				  std::vector<std::vector<double>> z_data, std::vector<std::vector<double>> dx_data,
				  std::vector<std::vector<double>> dy_data, std::vector<std::vector<double>> dz_data);
#ifdef GBP_HAVE_GRID_MAP
	// reference: fast_terrain_map.cpp:31-91 (index flip, float layers "elevation", "dx", "dy", "dz")
	void loadDataFromGridMap(grid_map::GridMap map);
#endif
	// B200 addition: the reference's CSV terrain directory (data/<name>/{x,y,z,dx,dy,dz}data.csv, rows = y, columns = x,
	// TerrainMapPublisher::loadMapFromCSV, terrain_map_publisher.cpp:330-370).  via_gridmap = false: fp64 values through
	// loadData; true: float layers / float resolution through the GridMap index flip, as the ROS pipeline delivers them.
	void loadDataFromCSV(const std::string &directory, bool via_gridmap = false);
	// B200 addition: the publisher's procedural sources.  createOwnMap: the "create" source (TerrainMapPublisher::createOwnMap,
	// terrain_map_publisher.cpp:34-231: 221 x 161 box world at 5 cm, the reference's rectangle table, heights from the Philox
	// stream `seed` instead of the reference's time(0)-seeded engine); createMap: the default source (:253-286, 12 x 5 m with a
	// 0.1 m disc).  Both go through the GridMap ingest like the ROS pipeline.
	void createOwnMap(uint64_t seed);
	void createMap();
	double getGroundHeight(const double x, const double y);              // :94-132
	bool heightIsNan(const double x, const double y);                    // :135-157
	std::array<double, 3> getSurfaceNormal(const double x, const double y);  // :160-213
	std::vector<double> getXData();                                      // :216-218
	std::vector<double> getYData();                                      // :221-223

	// batched forms (no reference counterpart): one launch for n points
	std::vector<double> getGroundHeight(const std::vector<double> &x, const std::vector<double> &y);

	// the C-ABI handle, for the other drop-in classes; throws if no terrain was loaded
	const gbp_terrain *handle() const;

private:
	void adopt(gbp_terrain *t, const char *what);
	std::shared_ptr<gbp_terrain> dev_;  // copies of a FastTerrainMap share the device terrain: the map data is read-only after loadData; the
	                                    // handle's only mutable part is the host-pointer staging ring, which the library guards with a mutex
	std::vector<double> x_data_, y_data_;
};

namespace gbp_dropin {
[[noreturn]] void raise(const char *what);  // std::runtime_error carrying gbp_last_error()
inline void check(int rc, const char *what) { if (rc != GBP_OK) raise(what); }
}  // namespace gbp_dropin

#ifdef GBP_HAVE_GRID_MAP
inline void FastTerrainMap::loadDataFromGridMap(grid_map::GridMap map) {
	const int nx = map.getSize()(0), ny = map.getSize()(1);
	std::vector<double> x(nx), y(ny);
	std::vector<std::vector<double>> z(nx, std::vector<double>(ny)), a(nx, std::vector<double>(ny, 0.0)),
		b(nx, std::vector<double>(ny, 0.0)), c(nx, std::vector<double>(ny, 1.0));
	const bool normals = map.exists("dx");
	for (int i = 0; i < nx; ++i) { grid_map::Index k((nx - 1) - i, 0); grid_map::Position p; map.getPosition(k, p); x[i] = p.x(); }
	for (int j = 0; j < ny; ++j) { grid_map::Index k(0, (ny - 1) - j); grid_map::Position p; map.getPosition(k, p); y[j] = p.y(); }
	for (int i = 0; i < nx; ++i)
		for (int j = 0; j < ny; ++j) {
			grid_map::Index k((nx - 1) - i, (ny - 1) - j);
			z[i][j] = (double) map.at("elevation", k);
			if (normals) { a[i][j] = (double) map.at("dx", k); b[i][j] = (double) map.at("dy", k); c[i][j] = (double) map.at("dz", k); }
		}
	loadData(nx, ny, x, y, z, a, b, c);
}
#endif

#endif
