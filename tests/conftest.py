import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden_walk(name):
    """(scene, golden dict) for a walk fixture; the scene is regenerated from the stored recipe and
    checked against the stored input digest, or rebuilt from stored inputs when present."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    recipe = ast.literal_eval(str(g["recipe"]))
    design = None
    if recipe.get("design"):
        from gpu_ray_tracing_for_waveguide_based_ar_display_b200.couplers_coor import WaveguideDesign
        design = WaveguideDesign(**recipe["design"])
    scene = si.make_scene(recipe["num_FOV_x"], recipe["num_FOV_y"], recipe["num_rays_per_FoV"],
                          seed=recipe["seed"], lmd_subset=recipe.get("lmd_subset"), eff=recipe.get("eff"),
                          eb=tuple(recipe.get("eb", (80, 120))), design=design, ray_pol=recipe.get("ray_pol"))
    return scene, g


def golden_bins(g):
    eb = np.zeros(int(np.prod(g["eb_shape"])), dtype=np.float32)
    eb[g["eb_index"]] = g["eb_value"]
    return eb.reshape(tuple(g["eb_shape"]))


def input_digest(scene):
    import hashlib
    h = hashlib.sha256()
    for a in scene.kernel_args(scene.new_matrix_EB()):
        if isinstance(a, np.ndarray):
            h.update(str(a.dtype).encode()); h.update(str(a.shape).encode())
            h.update(np.ascontiguousarray(a).tobytes())
        else:
            h.update(repr(float(a)).encode())
    return h.hexdigest()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def small_scene():
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
    return si.make_scene(4, 3, 48, seed=5)
