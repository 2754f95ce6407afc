// TEST INFRASTRUCTURE ONLY.  Included after glsl_prelude.h (and glsl_prelude_rt.h) and before the generated shader text.
#pragma once
// GLSL evaluates function arguments left to right; C++ leaves the order open (g++: right to left), which would swap the random
// numbers of `vec3(getRandomNegPos(), getRandomNegPos(), getRandomNegPos())`.  A braced initialiser list IS ordered left to right.
#define vec3(...) vec3{__VA_ARGS__}
#define vec2(...) vec2{__VA_ARGS__}      // the camera jitter: vec2(getRandomNegPos(), getRandomNegPos()), raytrace.rgen:1488
