// perf_timer.hpp — host-visible timing of a registered function, same contract as the
// reference's perf_test / perf_test_prelu (cpp_impl/perf.h:6-7, cpp_impl/perf.cpp:298-339):
// fresh X (integers in [-512,512]), B = 2, Y = 0 (+10 floats of slack, perf.cpp:307-308), then a
// calibrated repeat loop; the return value is TSC cycles per call.  For a CUDA-backed function
// this is the DRIVER-VISIBLE cost (H2D + kernel + D2H + sync); device-only kernel time is what
// bench.py reports.
#pragma once
#include "tsg_host.hpp"

float perf_test(comp_func f, int M, int K, int N, int nonZero);
float perf_test_prelu(comp_func_prelu f, int M, int K, int N, int nonZero);
// wall-clock seconds per call of the last perf_test*/ run (for reporting next to cycles)
double perf_last_seconds_per_call();
