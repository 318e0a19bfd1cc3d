"""GPU-side batch producer of the video / volume path: the reference's `dataio.Implicit3DWrapper` (dataio.py:546-600)
with the data array resident on the device.

The reference keeps the normalised video as a CPU tensor `[T * H * W, C]` (943 MB for the 300 x 512 x 512 RGB clip of
BASELINE configs[2]); every training item draws `z_num` frames and `M` pixels, gathers `data[z * n_xy + xy]` on the CPU
and the DataLoader copies the 3.6 MB batch to the GPU (train_video.py:61-63, training.py:92-93).  Once a training step
takes ~2 ms, that gather and copy are a visible share of the iteration.  Here the array is uploaded
once and a batch is one small index upload plus ONE gather kernel (`coordx_gather_rows`).

Same constructor arguments, attributes and `__getitem__` result as the reference class.  `__getitem__` draws its
indices from the SAME generators in the SAME order as the reference (numpy's global RandomState for the frame count,
torch's global CPU generator for the frame permutation and the pixel draw), so with equal seeds the batches are
identical to the reference's bit for bit; `sample(device_rng=True)` draws on the GPU instead (no host-side index
work, different random stream).  CUDA only, like the rest of the package: there is no CPU path.

A CUDA dataset cannot go through `DataLoader(pin_memory=True)`; iterate it directly:

    ds = Implicit3DWrapper(video, sidelength=video.shape, split_coord=True, sample_fraction=0.0038)
    for step in range(n):
        model_input, gt = ds.sample()            # what the reference's DataLoader(batch_size=1) yields, on the GPU
"""
import numpy as np
import torch

from . import _lib


def get_mgrid(sidelen, dim=2):
    """dataio.py:26-49: flattened pixel-centre grid in [-1, 1]^dim (CPU float32 tensor [prod(sidelen), dim])."""
    if isinstance(sidelen, int):
        sidelen = dim * (sidelen,)
    if dim not in (1, 2, 3):
        raise NotImplementedError('Not implemented for dim=%d' % dim)
    axes = []
    for d in range(dim):
        n = sidelen[d]
        den = max(n - 1, 1) if (dim == 3 and d == 0) else (n - 1)
        axes.append(np.arange(n, dtype=np.float32) / np.float32(den))
    grid = np.stack(np.meshgrid(*axes, indexing='ij'), axis=-1).astype(np.float32)
    grid -= 0.5
    grid *= 2.
    return torch.from_numpy(grid.reshape(-1, dim))


class Implicit3DWrapper:
    """dataio.py:546-600 with `self.data` and the coordinate grids on `device`."""

    def __init__(self, dataset, sidelength=None, sample_fraction=1., batch_size=1, split_coord=False,
                 frame_sample_fraction=1., pixel_sample_fraction=1., device='cuda'):
        if isinstance(sidelength, int):
            sidelength = 3 * (sidelength,)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.CoordxError('coordx_b200.dataio: a CUDA device is required (no CPU path; the reference class is the CPU path)')
        self.dataset = dataset
        self.batch_size = batch_size
        self.split_coord = split_coord
        vid = dataset[0]
        vid = vid if torch.is_tensor(vid) else torch.from_numpy(np.asarray(vid))
        data = (vid.to(torch.float32) - 0.5) / 0.5
        self.data = data.reshape(-1, self.dataset.channels).contiguous().to(self.device)
        self.sample_fraction = sample_fraction
        if split_coord:
            self.mgrid_xy = get_mgrid((sidelength[1], sidelength[2]), dim=2).unsqueeze(0).to(self.device)
            self.mgrid_z = get_mgrid(sidelength[0], dim=1).unsqueeze(1).to(self.device)
            self.n_z = sidelength[0]
            self.n_xy = sidelength[1] * sidelength[2]
            self.N_samples = self.n_z * self.n_xy * self.sample_fraction
        else:
            self.mgrid = get_mgrid(sidelength, dim=3).to(self.device)
            self.N_samples = int(self.sample_fraction * self.mgrid.shape[0])

    def __len__(self):
        return len(self.dataset) * self.batch_size

    # ------------------------------------------------------------------ gathers
    def _gather(self, idz, idx):
        """data[idz[:, None] * n_xy + idx[None, :]] (idz None: data[idx]) -> [len(idz) * len(idx), C]"""
        lib = _lib.load()
        nz = 1 if idz is None else int(idz.numel())
        m = int(idx.numel())
        out = torch.empty(nz * m, self.data.shape[1], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.coordx_gather_rows(self.data.data_ptr(), self.data.shape[1], None if idz is None else idz.data_ptr(), nz,
                                              self.n_xy if idz is not None else 0, idx.data_ptr(), m, out.data_ptr(),
                                              torch.cuda.current_stream(self.device).cuda_stream), 'coordx_gather_rows')
        return out

    def _draw(self, device_rng):
        """the index draws of dataio.py:577-579 / :589"""
        if self.split_coord:
            z_num = np.random.randint(round(self.n_z * 0.1), round(self.n_z * 0.9))
            m = round(self.N_samples / z_num)
            if device_rng:
                return torch.randperm(self.n_z, device=self.device)[:z_num].contiguous(), torch.randint(0, self.n_xy, (m,), device=self.device)
            idz = torch.randperm(self.n_z)[:z_num]
            idxy = torch.randint(0, self.n_xy, (m,))
            return idz.to(self.device, non_blocking=True), idxy.to(self.device, non_blocking=True)
        if device_rng:
            return None, torch.randint(0, self.data.shape[0], (self.N_samples,), device=self.device)
        return None, torch.randint(0, self.data.shape[0], (self.N_samples,)).to(self.device, non_blocking=True)

    def __getitem__(self, idx, device_rng=False):
        if self.split_coord:
            coord_idx, coord_idy = self._draw(device_rng)
            data = self._gather(coord_idx, coord_idy)
            coords_xy = self.mgrid_xy[:, coord_idy, :]
            coords_z = self.mgrid_z[coord_idx, :, :]
            in_dict = {'idx': idx, 'coords_split': [coords_z, coords_xy]}
            gt_dict = {'img': data}
        else:
            if self.sample_fraction < 1.:
                _, coord_idx = self._draw(device_rng)
                data = self._gather(None, coord_idx)
                coords = self.mgrid[coord_idx, :]
            else:
                coords = self.mgrid
                data = self.data
            in_dict = {'idx': idx, 'coords': coords}
            gt_dict = {'img': data}
        return in_dict, gt_dict

    def sample(self, idx=0, device_rng=False):
        """One batch as the reference's DataLoader(batch_size=1) collates it: a leading dimension of 1 on every tensor."""
        in_dict, gt_dict = self.__getitem__(idx, device_rng=device_rng)
        batched = {k: ([t.unsqueeze(0) for t in v] if isinstance(v, list) else (v.unsqueeze(0) if torch.is_tensor(v) else torch.tensor([v])))
                   for k, v in in_dict.items()}
        return batched, {k: v.unsqueeze(0) for k, v in gt_dict.items()}
