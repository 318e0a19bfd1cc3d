"""Batch producer of the video / volume path: dataio.py:26-49 get_mgrid and dataio.py:546-600
Implicit3DWrapper.  CPU: the oracle against the fixture made by executing the reference's own definitions
(oracle/gen_golden_dataio.py).  GPU: coordx_b200.dataio (device-resident data, coordx_gather_rows) against the oracle,
bit for bit, with the reference's index draws; and properties of the device-side draws."""
import os

import numpy as np
import pytest
import torch

from oracle import dataio_np as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'dataio_batches.npz')
CASES = [('split', True), ('split_b', True), ('flat', False), ('full', False)]


class _Video:                          # dataio.Video's surface (dataio.py:353-368)
    def __init__(self, vid):
        self.vid, self.shape, self.channels = vid, vid.shape[:-1], vid.shape[-1]

    def __len__(self):
        return 1

    def __getitem__(self, idx):
        return self.vid


def test_oracle_mgrid_matches_reference():
    g = np.load(GOLD)
    assert np.array_equal(D.get_mgrid(5, 1), g['mgrid1'])
    assert np.array_equal(D.get_mgrid((4, 7), 2), g['mgrid2'])
    assert np.array_equal(D.get_mgrid((3, 4, 5), 3), g['mgrid3'])


@pytest.mark.parametrize('name,split', CASES)
def test_oracle_batches_match_reference_bit_for_bit(name, split):
    g = np.load(GOLD)
    seed = int(g[name + '.seed'])
    np.random.seed(seed)
    torch.manual_seed(seed)
    for it in range(2):
        i, o = D.volume_batch(g['vid'], float(g[name + '.sample_fraction']), split)
        assert np.array_equal(o['img'], g[f'{name}.{it}.img'])
        if split:
            assert np.array_equal(i['coords_split'][0], g[f'{name}.{it}.coords_z'])
            assert np.array_equal(i['coords_split'][1], g[f'{name}.{it}.coords_xy'])
        else:
            assert np.array_equal(i['coords'], g[f'{name}.{it}.coords'])


def test_host_mgrid_of_the_package_matches_reference():
    from coordx_b200 import dataio as G
    g = np.load(GOLD)
    assert np.array_equal(G.get_mgrid(5, 1).numpy(), g['mgrid1'])
    assert np.array_equal(G.get_mgrid((4, 7), 2).numpy(), g['mgrid2'])
    assert np.array_equal(G.get_mgrid((3, 4, 5), 3).numpy(), g['mgrid3'])
    with pytest.raises(NotImplementedError):
        G.get_mgrid(4, dim=4)


def test_package_refuses_a_cpu_device():
    from coordx_b200 import dataio as G, _lib
    g = np.load(GOLD)
    with pytest.raises(_lib.CoordxError):
        G.Implicit3DWrapper(_Video(g['vid']), sidelength=g['vid'].shape[:-1], split_coord=True, device='cpu')


@pytest.mark.gpu
@pytest.mark.parametrize('name,split', CASES)
def test_gpu_batches_match_reference_bit_for_bit(name, split):
    """same seeds, same generators, same order of draws as the reference => identical batches (fixture of the reference)"""
    from coordx_b200 import dataio as G
    g = np.load(GOLD)
    ds = G.Implicit3DWrapper(_Video(g['vid']), sidelength=g['vid'].shape[:-1], split_coord=split,
                             sample_fraction=float(g[name + '.sample_fraction']))
    seed = int(g[name + '.seed'])
    np.random.seed(seed)
    torch.manual_seed(seed)
    for it in range(2):
        i, o = ds[0]
        assert o['img'].is_cuda
        assert np.array_equal(o['img'].cpu().numpy(), g[f'{name}.{it}.img'])
        if split:
            assert np.array_equal(i['coords_split'][0].cpu().numpy(), g[f'{name}.{it}.coords_z'])
            assert np.array_equal(i['coords_split'][1].cpu().numpy(), g[f'{name}.{it}.coords_xy'])
        else:
            assert np.array_equal(i['coords'].cpu().numpy(), g[f'{name}.{it}.coords'])


@pytest.mark.gpu
def test_gpu_video_batch_at_the_named_size_against_the_oracle_and_device_rng_properties():
    """BASELINE configs[2] shape (300 x 512 x 512 is 943 MB: 64 frames here, same code path): the gathered batch equals the
    oracle's gather for the same indices; with device-side draws every gt row is the data row its coordinates name."""
    from coordx_b200 import dataio as G
    T, H, W, C = 64, 512, 512, 3
    vid = torch.rand(T, H, W, C, generator=torch.Generator().manual_seed(0))
    ds = G.Implicit3DWrapper(_Video(vid.numpy()), sidelength=(T, H, W), split_coord=True, sample_fraction=298844 / (300 * 512 * 512))
    np.random.seed(3)
    torch.manual_seed(3)
    i, o = ds[0]
    np.random.seed(3)
    torch.manual_seed(3)
    i_ref, o_ref = D.volume_batch(vid.numpy(), ds.sample_fraction, True)
    assert np.array_equal(o['img'].cpu().numpy(), o_ref['img'])
    assert np.array_equal(i['coords_split'][0].cpu().numpy(), i_ref['coords_split'][0])
    assert np.array_equal(i['coords_split'][1].cpu().numpy(), i_ref['coords_split'][1])
    # device-side draws: shapes as the DataLoader collates them; every row consistent with its coordinates
    mi, gt = ds.sample(device_rng=True)
    cz, cxy = mi['coords_split']
    assert cz.shape[0] == 1 and cz.shape[2:] == (1, 1) and cxy.shape[:2] == (1, 1) and cxy.shape[3] == 2
    z_num, m = cz.shape[1], cxy.shape[2]
    assert gt['img'].shape == (1, z_num * m, C)
    t_idx = torch.round((cz.reshape(-1) * 0.5 + 0.5) * (T - 1)).long()
    y_idx = torch.round((cxy[0, 0, :, 0] * 0.5 + 0.5) * (H - 1)).long()
    x_idx = torch.round((cxy[0, 0, :, 1] * 0.5 + 0.5) * (W - 1)).long()
    assert len(torch.unique(t_idx)) == z_num                      # a permutation prefix: frames are distinct
    want = (vid.cuda()[t_idx][:, y_idx, x_idx, :] - 0.5) / 0.5      # [z_num, m, C]
    assert torch.equal(gt['img'].reshape(z_num, m, C), want)
