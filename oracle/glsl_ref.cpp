// TEST INFRASTRUCTURE ONLY.  The reference's OWN shader include files compiled as C++ (see glsl_prelude.h) into
// oracle/_ref/libglsl_ref.so: the RNG (tea / lcg / rnd), the direction samplers and frames of random.glsl / transform.glsl
// and the von Mises-Fisher sampling and densities of guiding.glsl — SURVEY.md §8(a) rows a1, a2 and the shader half of a13.
// tests/test_oracle_cpu.py runs the same inputs through oracle/tracer_oracle.cpp's restatement (oracle_unit_eval, same
// function numbers) and compares: the integer RNG bit for bit, the float functions within a few ulp (the restatement uses the
// deterministic elementary functions of include/b200pt_detmath.h, this build the C library's).
#include "glsl_prelude.h"
#include <string.h>
#include "glsl_macros.h"
namespace glsl {
#include "_ref/glsl/limits.inc"
#include "_ref/glsl/wavefront.inc"
#include "_ref/glsl/raycommon.inc"
#include "_ref/glsl/random_fwd.inc"
#include "_ref/glsl/transform_fwd.inc"
#include "_ref/glsl/transform.inc"
#include "_ref/glsl/random.inc"
#include "_ref/glsl/guiding.inc"
static pushConstant pushC;                 // layout(push_constant) of raytrace.rgen:96
#include "_ref/glsl/rgen_functions.inc"
}  // namespace glsl
using namespace glsl;

static vec3 v3in(const float *p) { return vec3(p[0], p[1], p[2]); }
static void v3out(float *p, vec3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
static Material matIn(const float *p) {      // b200pt_material (std430 layout of binding 3) -> the struct of wavefront.glsl as C++ lays it out
    Material m;
    m.lightColor = v3in(p); m.diffuse = v3in(p + 4); m.specular = v3in(p + 8);
    m.specularHighlight = p[11]; m.transparency = p[12]; m.refractionIndex = p[13]; m.refractionIndexInv = p[14]; m.eta = p[15]; m.k = p[16]; m.roughness = p[17];
    int32_t t[3]; memcpy(t, p + 18, 12);
    m.textureIdDiffuse = t[0]; m.textureIdSpecular = t[1]; m.type = t[2];
    return m;
}
static_assert(sizeof(VMM_Theta) == 720 && sizeof(VMF_Theta) == 40, "guiding.glsl's structs must have the scalar block layout of binding 16");

extern "C" {
uint32_t glsl_ref_tea(uint32_t a, uint32_t b) { return tea(a, b); }
uint32_t glsl_ref_lcg(uint32_t *prev) { return lcg(*prev); }
float glsl_ref_rnd(uint32_t *prev) { return rnd(*prev); }

// fn: see oracle_unit_eval in oracle/tracer_oracle.cpp (same numbering, same argument layout)
int glsl_ref_unit_eval(int fn, uint32_t *seed_io, const float *in, float *out) {
    seed = *seed_io;
    switch (fn) {
        case 0: v3out(out, randomOnUnitSphere()); break;
        case 1: v3out(out, randomInHemisphere(v3in(in))); break;
        case 2: v3out(out, randomInHemisphereCosine(v3in(in))); break;
        case 3: v3out(out, randomInHemisphereCosinePower(v3in(in), in[3])); break;
        case 4: { sphere s; s.center = v3in(in); s.radius = in[3]; s.materialIndex = 0; s.iLight = 0; vec3 n; v3out(out, randomOnSphere(s, n)); v3out(out + 3, n); break; }
        case 5: { sphere s; s.center = v3in(in); s.radius = in[3]; s.materialIndex = 0; s.iLight = 0; vec3 n; v3out(out, randomOnSphereVisible(s, v3in(in + 4), n)); v3out(out + 3, n); break; }
        case 6: { Material m; memset(&m, 0, sizeof(m)); m.roughness = in[3]; v3out(out, randomBeckmannNormal(m, v3in(in))); break; }
        case 7: v3out(out, toWorld(v3in(in), v3in(in + 3))); break;
        case 8: v3out(out, toLocal(v3in(in), v3in(in + 3))); break;
        case 9: { VMF_Theta t; memcpy(&t, in, sizeof(t)); v3out(out, sampleVMF(t, v3in(in + 10), in[13] != 0.0f)); break; }
        case 10: { VMF_Theta t; memcpy(&t, in, sizeof(t)); out[0] = vMF(v3in(in + 14), t, v3in(in + 10), in[13] != 0.0f); break; }
        case 11: { VMM_Theta t; memcpy(&t, in, sizeof(t)); v3out(out, sampleVMM(t, v3in(in + 180), in[183] != 0.0f)); break; }
        case 12: { VMM_Theta t; memcpy(&t, in, sizeof(t)); out[0] = VMM(v3in(in + 184), t, v3in(in + 180), in[183] != 0.0f); break; }
        case 13: { VMF_Theta t; memcpy(&t, in, sizeof(t)); t = updateK(t, in[10]); memcpy(out, &t, sizeof(t)); break; }
        // raytrace.rgen.  Material = b200pt_material (24 floats) at in[0]; see oracle_unit_eval for the argument layouts
        case 20: out[0] = fresnelFn(in[0], in[1]); break;
        case 21: out[0] = fresnelConductor(in[0], in[1], in[2]); break;
        case 22: v3out(out, evalBsdf(matIn(in), vec2(0.0f, 0.0f), v3in(in + 24), v3in(in + 27), v3in(in + 30), in[33] != 0.0f)); break;
        case 23: out[0] = pdfBSDF(matIn(in), v3in(in + 24), v3in(in + 27), v3in(in + 30)); break;
        case 24: { vec3 d = vec3(0.0f); out[0] = sampleBSDF(matIn(in), v3in(in + 27), v3in(in + 24), in[33] != 0.0f, d); v3out(out + 1, d); break; }
        case 25: out[0] = powerHeuristic(in[0], in[1]); out[1] = balanceHeuristic(in[0], in[1]); break;
        case 26: { estimate = v3in(in + 6); pushC.adrrsS = in[9]; int n = 0; out[0] = applyWeightWindow(v3in(in), v3in(in + 3), n); out[1] = float(n); break; }
        case 27: v3out(out, approxDiffuse(matIn(in), v3in(in + 24), v3in(in + 27), vec2(0.0f, 0.0f))); break;
        case 28: { Light l; memset(&l, 0, sizeof(l)); l.sampleProb = in[0]; l.area = in[1]; out[0] = pdfLight(l, v3in(in + 2), v3in(in + 5), in[8]); break; }
        case 29: { Material m = matIn(in); out[0] = float(hasDiscreteDirection(m)); out[1] = float(isMatAlmostDiscrete(m));
                   pushC.useIrradianceCacheOnGlossy = false; out[2] = float(isICCapable(m)); pushC.useIrradianceCacheOnGlossy = true; out[3] = float(isICCapable(m)); break; }
        default: return -1;
    }
    *seed_io = seed;
    return 0;
}
}
