"""Helpers to read tests/golden/*.npz (test infrastructure, see splitmlp_np.py)."""
from __future__ import annotations

import glob
import json
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import splitmlp_np as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def golden_names() -> List[str]:
    # (dataio_*.npz are the batch producer's fixtures, oracle/dataio_np.py)
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))
                  if not n.startswith('dataio_'))


@dataclass
class Fixture:
    name: str
    kw: dict                      # SingleBVPNet ctor kwargs used by the reference run
    shape: List[int]
    loss: str
    cfg: O.SplitCfg
    params: Dict[str, np.ndarray]  # keys below `net.` (see splitmlp_np docstring)
    state_dict: Dict[str, np.ndarray]  # full reference state_dict
    coords_raw: List[np.ndarray]  # as fed to the model: [1, .., N_c, .., f_c]
    coords: List[np.ndarray]      # [N_c, f_c]
    freq_bands: Optional[np.ndarray]
    include_coord: bool
    data: dict

    def encode(self, x):
        """positional encoding of one axis, as SingleBVPNet.forward does
        (modules.py:450-451); identity in mlp mode."""
        if self.freq_bands is None:
            return x
        return O.pos_enc(x, self.freq_bands, self.include_coord, freq_last=True)


def load(name: str) -> Fixture:
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    meta = json.loads(str(z['cfg']))
    kw = meta['kw']
    sd = {k[3:]: z[k] for k in z.files if k.startswith('sd/')}
    params = {k[4:]: v for k, v in sd.items() if k.startswith('net.')}
    fb = sd.get('positional_encoding.freq_bands')
    in_features = kw.get('in_features', 2)
    if fb is not None:
        inc = kw.get('include_coord', True)
        enc_in = in_features * (2 * len(fb) + (1 if inc else 0))
    else:
        inc, enc_in = True, in_features
    cfg = O.SplitCfg(in_features=enc_in, coord_dim=in_features, out_features=kw['out_features'],
                     hidden_features=kw['hidden_features'], num_hidden_layers=kw.get('num_hidden_layers', 3),
                     nonlinearity=kw['type'], approx_layers=kw.get('approx_layers', 2),
                     split_rule=kw.get('split_rule'), fusion_operator=kw.get('fusion_operator', 'prod'),
                     act_scale=kw.get('act_scale', 1), fusion_before_act=kw.get('fusion_before_act', False),
                     learn_code=kw.get('learn_code', False), last_layer_features=kw.get('last_layer_features', -1),
                     fusion_size=kw.get('fusion_size', 1), reduced=kw.get('reduced', False))
    assert cfg.feat_per_channel == meta['feat_per_channel'], (cfg.feat_per_channel, meta)
    C = len(meta['shape'])
    raw = [z[f'coords/{c}'] for c in range(C)]
    flat = [r.reshape(-1, r.shape[-1]) for r in raw]
    return Fixture(name=name, kw=kw, shape=meta['shape'], loss=meta['loss'], cfg=cfg, params=params,
                   state_dict=sd, coords_raw=raw, coords=flat, freq_bands=fb, include_coord=inc,
                   data={k: z[k] for k in z.files})
