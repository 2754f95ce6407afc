"""Regenerates tests/golden/shader_ref_frames.npz: the frames (and recorded samples) the REFERENCE's own shader source produces when
compiled as C++ (oracle/_ref/libshader_ref.so, see oracle/shader_ref.cpp) for every case of tests/test_shader_ref.py.  Run in the build
container (needs /root/reference); the frames travel, the reference does not."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import test_shader_ref as T  # noqa: E402

out = {}
T._check = lambda key, ours, ref_value: out.__setitem__(key, T.compact(ref_value).copy())     # record the live build's value
for scene_name, kw in T.PLAIN:
    T.test_path_tracing_frames_are_bit_equal_to_the_reference_shaders(scene_name, kw)
for name in dir(T):
    if name.startswith("test_") and name != "test_path_tracing_frames_are_bit_equal_to_the_reference_shaders":
        getattr(T, name)()
np.savez_compressed(T.GOLDEN, **out)
print(len(out), "arrays,", os.path.getsize(T.GOLDEN) // 1024, "KB")
