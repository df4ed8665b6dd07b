"""CPU: the pure-PyTorch project + composite baseline (oracle/torch_baseline.py, the north star's CPU baseline) against
the C oracle and a golden vector of the unmodified reference."""
import math

import numpy as np
import torch

import scenes as SC
import util as U
from oracle import torch_baseline as TB


def _run(sc):
    return TB.render(sc["means3D"], sc["shs"], sc["opacities"], sc["scales"], sc["rotations"], sc["camera"], sc["bg"],
                     sh_degree=sc["sh_degree"], scale_modifier=sc["scale_modifier"])


def test_torch_baseline_matches_the_c_oracle():
    sc = SC._scene("tb", 1500, 96, 64, 5150, sh_degree=2, bg=(0.2, 0.4, 0.9), log_scale=math.log(0.03), opacity_mean=0.5)
    st, _ = U.run_oracle(sc)
    color, radii, depth, alpha, R = _run(sc)
    ok_g = st.ambiguous_gauss == 0
    assert np.array_equal(radii.numpy()[ok_g], st.radii[ok_g])
    if ok_g.all():
        assert R == st.num_rendered
    ok = st.ambiguous_pix == 0
    for mine, ref in ((color, st.color), (depth, st.depth), (alpha, st.alpha)):
        assert np.abs(mine.numpy() - ref)[:, ok].max() <= 1e-4


def test_torch_baseline_matches_a_reference_golden_vector():
    g = U.load_golden("deg1_white")  # captured from the unmodified reference on a B200 (tests/golden/make_golden.py)
    sc = U.scene_from_golden(g)
    color, radii, depth, alpha, R = _run(sc)
    assert np.array_equal(radii.numpy(), g["radii"])
    assert R == int(g["num_rendered"])
    # pixels that sit on a cut (alpha = 1/255, T = 1e-4) may flip with the host's exp(): allow one in a thousand
    for mine, ref in ((color, g["color"]), (depth, g["depth"]), (alpha, g["alpha"])):
        assert np.quantile(np.abs(mine.numpy() - ref), 0.999) <= 1e-4
