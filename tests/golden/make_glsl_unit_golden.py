"""Regenerates tests/golden/glsl_unit_golden.npz: outputs of the REFERENCE's own shader include files (shaders/random.glsl,
transform.glsl, guiding.glsl compiled as C++ into oracle/_ref/libglsl_ref.so by oracle/Makefile) on the seeded inputs of
tests/test_oracle_cpu.py.  Run in the build container (needs /root/reference); the vectors travel, the reference does not."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
import test_oracle_cpu as T  # noqa: E402

R = C.CDLL(T.GLSL_REF)
R.glsl_ref_tea.restype = C.c_uint32; R.glsl_ref_tea.argtypes = [C.c_uint32, C.c_uint32]
R.glsl_ref_lcg.restype = C.c_uint32; R.glsl_ref_lcg.argtypes = [C.c_void_p]
R.glsl_ref_rnd.restype = C.c_float; R.glsl_ref_rnd.argtypes = [C.c_void_p]
rng = np.random.default_rng(1)
pairs = rng.integers(0, 2 ** 32, (300, 2), dtype=np.uint64).astype(np.uint32)
out = {"tea": np.array([R.glsl_ref_tea(int(a), int(b)) for a, b in pairs], np.uint32)}
s1, s2 = np.array([12345], np.uint32), np.array([12345], np.uint32)
out["lcg"] = np.array([R.glsl_ref_lcg(s1.ctypes.data) for _ in range(300)], np.uint32)
out["rnd"] = np.array([R.glsl_ref_rnd(s2.ctypes.data) for _ in range(300)], np.float32)
for fn, name, nout, inputs in T._unit_cases():
    vals, seeds = T._run_unit(R, "glsl_ref_unit_eval", fn, nout, inputs)
    out[name], out[name + "_seed"] = vals, seeds
np.savez_compressed(T.GLSL_GOLDEN, **out)
print({k: v.shape for k, v in out.items()})
