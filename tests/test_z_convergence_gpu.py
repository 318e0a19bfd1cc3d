"""End-metric parity on the GPU (north_star: PSNR within 0.1 dB, IoU within 0.005): the training trajectory of the
CUDA path against the torch restatement of the reference from the same initial weights, at a reduced step count so
that the whole file runs in about a minute (the full 10 000-iteration runs are `python tests/convergence.py`,
results in profiles/r01f_convergence.jsonl).  SIREN training is chaotic, so every case also trains the reference
twice (second copy: initial weights perturbed by 1e-7 relative) and the tolerance is widened by twice that measured
reference-vs-reference difference.  Image cases compare the PSNR of the mean training loss over the last 100
iterations (the full-image MSE), which is far less noisy than one snapshot."""
import pytest

from tests import convergence as CV

pytestmark = pytest.mark.gpu
CASES = {c[0]: c for c in CV.cases()}


@pytest.mark.parametrize('name,steps', [('img256_sum_cfg1', 1000), ('img512_prod_bf16', 1500), ('img512_prod_fp32', 600)])
def test_image_psnr_tracks_reference(name, steps):
    _, kw, kind, shape, prec, _, lr, clip = CASES[name]
    rec = CV.run(name, kw, kind, shape, prec, steps, lr, clip)
    print(rec)
    assert rec['tail_ref'] > 20.0, rec                 # the run actually learned something
    assert abs(rec['tail_diff']) <= 0.1 + 2 * abs(rec['tail_noise']), rec


def test_occupancy_iou_tracks_reference():
    _, kw, kind, shape, prec, _, lr, clip = CASES['occ_torus_prod_bf16']
    rec = CV.run('occ_torus_prod_bf16', kw, kind, shape, prec, 2000, lr, clip)
    print(rec)
    assert rec['ref'] > 0.8, rec
    assert abs(rec['diff']) <= 0.005 + 2 * abs(rec['noise']), rec
