// TEST INFRASTRUCTURE ONLY (oracle/_ref build). Empty stand-in for <ros/ros.h>.
// The reference core (planning_utils.h:15) includes it but uses no ROS symbol.
#pragma once
