"""CPU restatement of the reference's batch producer for the video / volume path (TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU leg may import this; the product never does).

  get_mgrid          dataio.py:26-49    flattened pixel-centre grid in [-1, 1]^dim, row-major
  volume_batch       dataio.py:575-600  Implicit3DWrapper.__getitem__: the split (z x xy) and the flat sampling

Pinned to tests/golden/dataio_batches.npz, which oracle/gen_golden_dataio.py produced by executing the reference's own
two definitions.  The index draws use the SAME generators in the SAME order as the reference (numpy's global RandomState
for the frame count, torch's global CPU generator for the permutation and the pixel draw), so with the same seeds the
batches are identical bit for bit; the gathers are plain numpy.
"""
import numpy as np
import torch


def get_mgrid(sidelen, dim=2):
    """dataio.py:26-49"""
    if isinstance(sidelen, int):
        sidelen = dim * (sidelen,)
    axes = []
    for d in range(dim):
        n = sidelen[d]
        den = max(n - 1, 1) if (dim == 3 and d == 0) else (n - 1)      # dataio.py:37: only the first 3-D axis is guarded
        axes.append(np.arange(n, dtype=np.float32) / np.float32(den))
    grid = np.stack(np.meshgrid(*axes, indexing='ij'), axis=-1).astype(np.float32)
    grid -= 0.5
    grid *= 2.
    return grid.reshape(-1, dim)


def draw_split_indices(n_z, n_xy, n_samples):
    """dataio.py:577-579: frame count from numpy's global state, frame permutation and pixel draw from torch's."""
    z_num = np.random.randint(round(n_z * 0.1), round(n_z * 0.9))
    idz = torch.randperm(n_z)[:z_num].numpy()
    idxy = torch.randint(0, n_xy, (round(n_samples / z_num),)).numpy()
    return idz, idxy


def volume_batch(vid, sample_fraction=1., split_coord=False, indices=None):
    """Implicit3DWrapper(Video).__getitem__ (dataio.py:546-600) for one [T, H, W, C] array in [0, 1].
    Returns (in_dict, gt_dict) of numpy arrays WITHOUT the DataLoader's batch dimension."""
    T, H, W, C = vid.shape
    data = ((vid.astype(np.float32) - np.float32(0.5)) / np.float32(0.5)).reshape(-1, C)      # dataio.py:555-556
    if split_coord:
        n_z, n_xy = T, H * W
        mgrid_xy = get_mgrid((H, W), dim=2)[None]                # [1, n_xy, 2]   dataio.py:560
        mgrid_z = get_mgrid(T, dim=1)[:, None]                   # [n_z, 1, 1]    dataio.py:561
        n_samples = n_z * n_xy * sample_fraction                 # dataio.py:564 (a float)
        idz, idxy = indices if indices is not None else draw_split_indices(n_z, n_xy, n_samples)
        coord_id = (idz[:, None].astype(np.int64) * n_xy + idxy[None, :]).reshape(-1)
        return ({'coords_split': [mgrid_z[idz], mgrid_xy[:, idxy]]}, {'img': data[coord_id]})
    mgrid = get_mgrid((T, H, W), dim=3)
    if sample_fraction < 1.:
        n = int(sample_fraction * mgrid.shape[0])                # dataio.py:567
        idx = indices if indices is not None else torch.randint(0, data.shape[0], (n,)).numpy()
        return {'coords': mgrid[idx]}, {'img': data[idx]}
    return {'coords': mgrid}, {'img': data}
