// Drop-in for the reference's include/global_body_planner/graph_class.h (class GraphClass, :24-176).
// Host containers keep the bookkeeping the reference API exposes (arbitrary ids, parent/child lists);
// vertex states are mirrored into a device SoA store (gbp_tree) that serves the neighbour queries of
// PlannerClass.  Copies are deep on the host and share-then-rebuild on the device.
#ifndef GBP_DROPIN_GRAPH_CLASS_H
#define GBP_DROPIN_GRAPH_CLASS_H

#include <memory>
#include <unordered_map>
#include <vector>

#include "planning_utils.h"

using namespace planning_utils;

class GraphClass {
public:
	GraphClass();
	GraphClass(const GraphClass &other);
	GraphClass &operator=(const GraphClass &other);
	virtual ~GraphClass();

	void addVertex(int index, State s);
	State getVertex(int index);
	int getNumVertices();
	virtual void addEdge(int idx1, int idx2);
	void removeEdge(int idx1, int idx2);
	virtual int getPredecessor(int idx);
	std::vector<int> getSuccessors(int idx);
	void addAction(int idx, Action a);
	Action getAction(int idx);
	void updateGYValue(int idx, double g_val, double y_val);
	double getGValue(int idx);
	double getYValue(int idx);
	void printVertex(State s);
	void printVertices();
	void printIncomingEdges(int idx);
	virtual void printEdges();
	virtual void init(State s, bool cost_add_yaw_flag, double cost_add_yaw_length_weight, double cost_add_yaw_yaw_weight);

protected:
	struct Node {
		State state;
		Action action;
		std::vector<int> parents, children;
		double g = 0.0, y = 0.0;
		bool has_action = false;
	};
	Node &node(int idx);                 // inserts on miss, like the reference's operator[] reads
	gbp_tree *device_store();            // device mirror, synchronised lazily with `order_`
	std::unordered_map<int, Node> nodes_;
	std::vector<int> order_;             // ids in insertion order = device slot order
	bool cost_add_yaw_flag_ = false;
	double cost_add_yaw_length_weight_ = 1, cost_add_yaw_yaw_weight_ = 1;

private:
	struct Mirror;
	std::unique_ptr<Mirror> mirror_;
};

#endif
