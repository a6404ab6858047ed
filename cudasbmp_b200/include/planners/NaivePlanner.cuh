/* planners/NaivePlanner.cuh — placeholder.  The reference's NaivePlanner is an early experiment that its own build does not
 * compile (CMakeLists.txt:25-33) and that is not on the KGMT path; demos/main.cu merely includes the header.
 * Nothing is declared here on purpose (out of scope). */
#pragma once
#include "planners/Planner.cuh"
