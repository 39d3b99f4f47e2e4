"""GPU test of the drop-in claim itself: the UNMODIFIED reference network, config class and
CustomBatch run on the B200 kernels after ``dropin.install()`` and agree with the reference's stock operator chain on
the same CUDA batch. The reference travels to the GPU box as the git-ignored ``baseline/_ref`` install
(tools/install_reference.py); without it the test is skipped."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_harness

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-3


@pytest.mark.skipif(ref_harness.find_root() is None, reason="no copy of the reference on this machine")
@pytest.mark.parametrize("cfg_name,in_radius,batch_num", [("vaihingen_pl", 9.0, 3), ("dales_pl", 7.0, 2)])
def test_unmodified_reference_kpfcnn_on_the_dropin(cfg_name, in_radius, batch_num):
    out = subprocess.run([sys.executable, os.path.join(HERE, "ref_dropin_script.py"), cfg_name, str(in_radius),
                          str(batch_num)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    print(r)
    assert r["layers"] == 5 and r["pyramid_tensors_compared"] == 25
    assert r["pyramid_paths_identical"], "device pyramid != reference walk on the numpy drop-ins"
    assert r["kpconv_modules_swapped"] == 10
    assert r["logits_rel"] < TOL, r
    assert abs(r["loss_new"] - r["loss_ref"]) < TOL * abs(r["loss_ref"]), r
    assert r["n_grads"] >= 40
    # parameter gradients: 1e-3 on the whole gradient vector (norm-wise). The worst single parameter is not a property
    # of the operator: the STOCK network, fed features perturbed by one TF32 rounding (2^-11 relative), moves its own
    # per-parameter gradients by 1-6e-2 (a LeakyReLU / max-pool decision flips on one of the few deep-layer points), so
    # that measured sensitivity, not 1e-3, bounds the per-parameter maximum. Without activation kinks
    # (negative_slope = 1 on both networks) the same comparison is reported as linear_*.
    assert r["grad_rel_l2"] < TOL, r
    assert r["linear_grad_rel_l2"] < TOL and r["linear_logits_rel"] < TOL, r
    assert r["grad_rel_max"] < 3.0 * max(r["stock_perturbed_grad_rel_max"], TOL), r
