// TEST INFRASTRUCTURE ONLY (oracle/_ref build). Small stand-in for grid_map_core.
// Only FastTerrainMap::loadDataFromGridMap touches it (fast_terrain_map.cpp:31-91):
// getSize()(i), getPosition(Index, Position&), at(layer, Index), exists(layer).
// Conventions follow grid_map: float layers, index (0,0) is the cell with the LARGEST
// x and y, position(i,j) = centre + (0.5*(n-1) - i) * resolution.
#pragma once
#include <array>
#include <map>
#include <string>
#include <vector>
namespace grid_map {
struct Index {
	int v[2];
	Index() { v[0] = v[1] = 0; }
	Index(int a, int b) { v[0] = a; v[1] = b; }
	int operator()(int i) const { return v[i]; }
};
typedef Index Size;
struct Position {
	double p[2];
	Position() { p[0] = p[1] = 0.0; }
	double x() const { return p[0]; }
	double y() const { return p[1]; }
};
class GridMap {
public:
	GridMap() : nx_(0), ny_(0), res_(1.0), cx_(0.0), cy_(0.0) {}
	void setGeometry(int nx, int ny, double resolution, double centre_x, double centre_y) {
		nx_ = nx; ny_ = ny; res_ = resolution; cx_ = centre_x; cy_ = centre_y;
	}
	void add(const std::string &layer, float value = 0.0f) { layers_[layer].assign((size_t) nx_ * ny_, value); }
	bool exists(const std::string &layer) const { return layers_.count(layer) != 0; }
	Size getSize() const { return Size(nx_, ny_); }
	bool getPosition(const Index &idx, Position &pos) const {
		pos.p[0] = cx_ + (0.5 * (nx_ - 1) - idx(0)) * res_;
		pos.p[1] = cy_ + (0.5 * (ny_ - 1) - idx(1)) * res_;
		return true;
	}
	float &at(const std::string &layer, const Index &idx) { return layers_[layer][(size_t) idx(0) * ny_ + idx(1)]; }
private:
	int nx_, ny_;
	double res_, cx_, cy_;
	std::map<std::string, std::vector<float> > layers_;
};
}  // namespace grid_map
