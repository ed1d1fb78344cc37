// Stand-in for /root/reference/tests/test_context.hpp when building the reference's TestCompressionBC7.cpp against the CUDA
// drop-in (integration/Makefile): the original pulls in the Vulkan test fixture (vulkan_test_context_t), which none of the
// compression tests uses.  Everything the test source itself needs from it is gtest and <cmath>.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <gtest/gtest.h>
