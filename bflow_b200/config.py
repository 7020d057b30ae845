"""Model-parameter dictionaries accepted by :class:`bflow_b200.RAFTSpline`.

The constructor takes the same plain ``dict`` the reference builds from its Hydra config
(keys read at models/raft_spline/raft.py:17-53 and models/raft_spline/update.py:53,61,103-105);
no Hydra/OmegaConf dependency exists on this path.  The two presets restate
config/experiment/dsec/raft_spline/E_LU4_BD2_lowpyramid.yaml and
config/experiment/multiflow/raft_spline/E_I_LU5_BD10_lowpyramid.yaml on top of
config/model/{base,raft_base,raft-spline}.yaml.
"""
from __future__ import annotations

import copy
from typing import Any, Dict, List, Tuple


def _base() -> Dict[str, Any]:
    return dict(
        name='raft-spline',
        detach_bezier=False,
        use_gma=False,
        correlation=dict(use_cosine_sim=False),
        hidden=dict(dim=128),
        context=dict(dim=128, norm='batch'),
        feature=dict(dim=256, norm='instance'),
        motion=dict(dim=128),
        num_iter=dict(train=12, test=12),
    )


def dsec_e_lu4_bd2() -> Dict[str, Any]:
    """Config "D": events only, 5+5-1 = 9 input bins, degree 2, 480x640."""
    c = _base()
    c.update(num_bins=dict(context=5, correlation=5), bezier_degree=2,
             use_boundary_images=False, use_events=True)
    c['correlation'].update(ev=dict(target_indices=[1, 2, 3, 4], levels=[1, 1, 1, 4], radius=[4, 4, 4, 4]),
                            img=dict(levels=None, radius=None))
    return c


def multiflow_e_i_lu5_bd10() -> Dict[str, Any]:
    """Config "M": events + boundary images, 41+25-1 = 65 input bins, degree 10, 384x512."""
    c = _base()
    c.update(num_bins=dict(context=41, correlation=25), bezier_degree=10,
             use_boundary_images=True, use_events=True)
    c['correlation'].update(ev=dict(target_indices=[8, 16, 24, 32, 40], levels=[1, 1, 1, 1, 4],
                                    radius=[4, 4, 4, 4, 4]),
                            img=dict(levels=4, radius=4))
    return c


PRESETS = {'E_LU4_BD2': dsec_e_lu4_bd2, 'E_I_LU5_BD10': multiflow_e_i_lu5_bd10}


def preset(name: str) -> Dict[str, Any]:
    return copy.deepcopy(PRESETS[name]())


# ---- derived quantities (all host-side bookkeeping) ------------------------------------------
def input_channels(cfg: Dict[str, Any]) -> int:
    return cfg['num_bins']['context'] + cfg['num_bins']['correlation'] - 1


def levels_per_target(cfg: Dict[str, Any]) -> List[int]:
    """Pyramid depth of every base target, event targets first, then the image target
    (CorrComputation.__add__, models/raft_utils/corr.py:223-227)."""
    out: List[int] = []
    if cfg['use_events']:
        out += [int(v) for v in cfg['correlation']['ev']['levels']]
    if cfg['use_boundary_images']:
        out.append(int(cfg['correlation']['img']['levels']))
    return out


def slot_table(levels: List[int]) -> List[Tuple[int, int]]:
    """(level, base target) of every lookup slot: level-major, ascending target
    (models/raft_utils/corr.py:322-346)."""
    return [(lvl, t) for lvl in range(max(levels)) for t, n in enumerate(levels) if n > lvl]


def lookup_timestamps(cfg: Dict[str, Any]) -> List[float]:
    """models/raft_spline/raft.py:156,170-177."""
    ts: List[float] = []
    if cfg['use_events']:
        dt = 1 / (cfg['num_bins']['context'] - 1)
        ts += [dt * i for i in cfg['correlation']['ev']['target_indices']]
    if cfg['use_boundary_images']:
        ts.append(1)
    return ts
