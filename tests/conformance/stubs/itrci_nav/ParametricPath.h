// TEST STUB.  The reference's NMPCNavControl.h:7-8 includes two message headers of the private
// ROS package `itrci_nav`; the solver wrappers use nothing from them except the standard headers
// they pull in.  This stub stands in for them when the wrappers are compiled for the conformance test.
#pragma once
#include <list>
#include <string>
#include <vector>
#include <cmath>
