"""Golden fixtures for the batch producer of the video / volume path, made by RUNNING THE REFERENCE'S OWN CODE
(test infrastructure; runs only in the build container, where /root/reference exists).

dataio.py does not import here (skvideo, skimage, cmapy ... are not installed), so the two definitions on this path -
`get_mgrid` (dataio.py:26-49) and `Implicit3DWrapper` (dataio.py:546-600) - are taken out of its syntax tree and
executed UNCHANGED in a namespace that holds only numpy and torch.  Nothing of the reference is copied into the repo.

    python oracle/gen_golden_dataio.py     # writes tests/golden/dataio_batches.npz
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'dataio_batches.npz')
SRC = '/root/reference/dataio.py'


def reference_namespace():
    tree = ast.parse(open(SRC).read())
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in ('get_mgrid', 'Implicit3DWrapper')]
    assert len(keep) == 2
    ns = {'np': np, 'torch': torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), SRC, 'exec'), ns)
    return ns


class TinyVideo:                      # stands in for dataio.Video (dataio.py:353-368): one [T, H, W, C] array in [0, 1]
    def __init__(self, vid):
        self.vid = vid
        self.shape = vid.shape[:-1]
        self.channels = vid.shape[-1]

    def __len__(self):
        return 1

    def __getitem__(self, idx):
        return self.vid


def main():
    ns = reference_namespace()
    rng = np.random.RandomState(7)
    vid = rng.rand(20, 6, 9, 3).astype(np.float32)
    out = {'vid': vid, 'mgrid1': ns['get_mgrid'](5, dim=1).numpy(), 'mgrid2': ns['get_mgrid']((4, 7), dim=2).numpy(),
           'mgrid3': ns['get_mgrid']((3, 4, 5), dim=3).numpy()}
    cases = [('split', dict(split_coord=True, sample_fraction=0.05), 11), ('split_b', dict(split_coord=True, sample_fraction=0.2), 12),
             ('flat', dict(split_coord=False, sample_fraction=0.1), 13), ('full', dict(split_coord=False, sample_fraction=1.0), 14)]
    for name, kw, seed in cases:
        ds = ns['Implicit3DWrapper'](TinyVideo(vid), sidelength=vid.shape[:-1], **kw)
        np.random.seed(seed)
        torch.manual_seed(seed)
        for it in range(2):               # two consecutive items: the RNG streams continue
            in_dict, gt_dict = ds[0]
            out[f'{name}.{it}.img'] = gt_dict['img'].numpy()
            if 'coords_split' in in_dict:
                out[f'{name}.{it}.coords_z'] = in_dict['coords_split'][0].numpy()
                out[f'{name}.{it}.coords_xy'] = in_dict['coords_split'][1].numpy()
            else:
                out[f'{name}.{it}.coords'] = in_dict['coords'].numpy()
        out[f'{name}.seed'] = np.int64(seed)
        out[f'{name}.sample_fraction'] = np.float64(kw['sample_fraction'])
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, {k: v.shape for k, v in out.items() if hasattr(v, 'shape')})


if __name__ == '__main__':
    main()
