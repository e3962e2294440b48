/* compat.h -- force-included when the reference's headers are compiled host-side with g++
 * (TEST INFRASTRUCTURE ONLY). */
#pragma once
#include <cstring>
#include <cerrno>
#define __assume(x) __builtin_unreachable()
/* The reference targets MSVC, where krrmath/constants.h replaces M_PI by the FLOAT literal
 * (constants.h:8-17).  glibc's <math.h> defines a double M_PI first, which would silently promote
 * expressions such as `M_PI / 4 * x` (render/sampling.h:62-66) to double; restore the float one. */
#include <cmath>
#include <math.h>
#undef M_PI
#define M_PI 3.14159265358979323846f
